"""GPTQ-PB calibration on B200: drop-ins for the reference's gptq_pb/{gptq,low_quant,high_quant}.py with
the same class names, constructor signatures and method protocol -- `LowHighGPT(layer, lowq, highq, salient_metric)
.add_batch(inp, out) / .fasterquant(low_frac, blocksize, percdamp) / .free()` (gptq_pb/gptq.py:15-194) -- so the
reference's `quant_sequential` (gptq_pb/run.py:127-165) can use them unchanged.

What runs where:
  * Hessian accumulation X^T X and the cross-block update are fp32 library GEMMs (TF32 off, as gptq.py:11-12), the
    damped Cholesky-inverse factor is torch.linalg (cuSOLVER) -- library calls in the reference too;
  * the O(K) Python column loop of fasterquant (gptq.py:144-163: about ten small launches per column, 4096 columns per
    layer) is ONE kernel per 128-column block here (pbl_gptq_block, csrc/pbllm_gptq.cu);
  * the result is what the reference produces: `layer.weight.data` rewritten with fake-quant values in the layer's
    dtype (gptq.py:180-184) and the low mask saved to ./outputs/mask/mask_{low_frac}_{global_name}.pkl (gptq.py:108-114),
    i.e. exactly the input of pb.replace_from_fakequant.

Covered configuration: the one gptq_pb/run.py builds -- LowQuantizer(method="xnor"), HighQuantizer(bits, perchannel=True,
sym=False, mse=False). Other low methods raise NotImplementedError."""
from __future__ import annotations

import ctypes as C
import math
import os
import time

import torch
import torch.nn as nn

from . import _lib

OUTPUTMASK = 1


class LowQuantizer(nn.Module):
    """Reference gptq_pb/low_quant.py:6-82, method "xnor": per (group, row) mean and mean-absolute-deviation of the
    masked weights; quantize = mean + scale * sign(w - mean)."""

    def __init__(self, weight, method="xnor", groupsize=-1):
        super().__init__()
        if method != "xnor":
            raise NotImplementedError(f"LowQuantizer method {method!r}: the GPTQ-PB path (gptq_pb/run.py:134) uses 'xnor'")
        oc, ic = weight.shape
        if groupsize == -1:
            groupsize = ic
        self.groupsize = groupsize
        self.n_groups = math.ceil(ic / groupsize)
        self.register_buffer("scale", torch.zeros(self.n_groups, oc, 1))
        self.register_buffer("mean", torch.zeros(self.n_groups, oc, 1))
        self.method = method

    def calibrate(self, w, mask=None, groupi=0):
        if self.scale.device != w.device:
            self.scale, self.mean = self.scale.to(w.device), self.mean.to(w.device)
        w_mean = w.mean(-1).view(-1, 1)                       # low_quant.py:27 (over the masked-with-zeros row)
        self.mean[groupi] = w_mean
        w = w - w_mean
        self.scale[groupi] = w.abs().mean(-1, keepdim=True)   # :32

    def quantize(self, w, groupi=0):
        if w.device != self.scale.device:
            self.scale, self.mean = self.scale.to(w.device), self.mean.to(w.device)
        w_mean = self.mean[groupi]
        w = (w - w_mean).sign() * self.scale[groupi]
        return w + w_mean                                     # :76-82


def quantize(x, scale, zero, maxq):
    """Reference gptq_pb/high_quant.py:6-8."""
    q = torch.clamp(torch.round(x / scale) + zero, 0, maxq)
    return scale * (q - zero)


class HighQuantizer(nn.Module):
    """Reference gptq_pb/high_quant.py:10-122 for weights, per-channel min/max grid (mse=False)."""

    def __init__(self, bits, perchannel=False, sym=True, mse=False, norm=2.4, grid=100, maxshrink=.8, grouprows=1, shape=1):
        super().__init__()
        if mse or grouprows != 1:
            raise NotImplementedError("HighQuantizer: mse search / grouprows are not used by the GPTQ-PB path (gptq_pb/run.py:136-141)")
        self.register_buffer("maxq", torch.tensor(2 ** bits - 1))
        self.register_buffer("scale", torch.zeros(shape))
        self.register_buffer("zero", torch.zeros(shape))
        self.perchannel, self.sym = perchannel, sym

    def calibrate(self, x, weight=False):
        if not weight:
            raise NotImplementedError("HighQuantizer.calibrate: weights only")
        dev = x.device
        self.maxq = self.maxq.to(dev)
        shape = x.shape
        x = x.flatten(1) if self.perchannel else x.flatten().unsqueeze(0)
        tmp = torch.zeros(x.shape[0], device=dev)
        xmin = torch.minimum(x.min(1)[0], tmp)
        xmax = torch.maximum(x.max(1)[0], tmp)
        if self.sym:
            xmax = torch.maximum(torch.abs(xmin), xmax)
            neg = xmin < 0
            if torch.any(neg):
                xmin[neg] = -xmax[neg]
        flat = (xmin == 0) & (xmax == 0)
        xmin[flat] = -1
        xmax[flat] = +1
        self.scale = (xmax - xmin) / self.maxq
        self.zero = torch.full_like(self.scale, (self.maxq + 1) / 2) if self.sym else torch.round(-xmin / self.scale)
        if not self.perchannel:
            self.scale, self.zero = self.scale.repeat(shape[0]), self.zero.repeat(shape[0])
        shape = [-1] + [1] * (len(shape) - 1)
        self.scale, self.zero = self.scale.reshape(shape), self.zero.reshape(shape)

    def quantize(self, x, blocki=None):
        return quantize(x, self.scale, self.zero, self.maxq) if self.ready() else x

    def enabled(self):
        return self.maxq > 0

    def ready(self):
        return torch.all(self.scale != 0)


class LowHighGPT:
    """Reference gptq_pb/gptq.py:15-194 for nn.Linear layers."""

    def __init__(self, layer, low_quantizer, high_quantizer, salient_metric, disable_gptq=False):
        if not isinstance(layer, nn.Linear):
            raise NotImplementedError("LowHighGPT: nn.Linear layers (the OPT / LLaMA decoder linears of SURVEY.md 8)")
        self.layer = layer
        self.dev = layer.weight.device
        if self.dev.type != "cuda":
            raise RuntimeError("LowHighGPT runs on a CUDA device: pb-llm_b200 has no CPU path")
        self.rows, self.columns = layer.weight.shape
        self.H = torch.zeros((self.columns, self.columns), device=self.dev)
        self.nsamples = 0
        self.low_quantizer, self.high_quantizer = low_quantizer, high_quantizer
        self.salient_metric = salient_metric
        self.disable_gptq = disable_gptq
        self.mask = None

    def add_batch(self, inp, out=None, blocksize=1024):
        """gptq.py:35-51: running average of (2/n) X^T X in fp32, TF32 off."""
        if len(inp.shape) == 2:
            inp = inp.unsqueeze(0)
        tmp = inp.shape[0]
        if len(inp.shape) == 3:
            inp = inp.reshape((-1, inp.shape[-1]))
        inp = inp.t()
        self.H *= self.nsamples / (self.nsamples + tmp)
        self.nsamples += tmp
        inp = math.sqrt(2 / self.nsamples) * inp.float()
        tf32 = torch.backends.cuda.matmul.allow_tf32
        torch.backends.cuda.matmul.allow_tf32 = False
        try:
            self.H += inp.matmul(inp.t())
        finally:
            torch.backends.cuda.matmul.allow_tf32 = tf32

    @torch.no_grad()
    def fasterquant(self, low_frac, blocksize=128, percdamp=0.01):
        if blocksize != 128:
            raise NotImplementedError("fasterquant: blocksize 128 (the reference's default and only used value)")
        lowq, highq = self.low_quantizer, self.high_quantizer
        W = self.layer.weight.data.clone().float()
        if not highq.ready():
            highq.calibrate(W, weight=True)
        tick = time.time()
        H = self.H
        del self.H
        dead = torch.diag(H) == 0
        H[dead, dead] = 1
        W[:, dead] = 0
        losses = torch.zeros(self.rows, device=self.dev)
        damp = percdamp * torch.mean(torch.diag(H))
        diag = torch.arange(self.columns, device=self.dev)
        H[diag, diag] += damp
        H = torch.linalg.cholesky(H)
        H = torch.cholesky_inverse(H)
        H = torch.linalg.cholesky(H, upper=True)
        Hinv = H.contiguous()
        mask = torch.zeros_like(W, dtype=torch.bool)
        for groupi in range(lowq.n_groups):                       # gptq.py:83-105
            st = groupi * lowq.groupsize
            ed = min(st + lowq.groupsize, self.columns)
            if self.salient_metric == "magnitude":
                sal = torch.abs(W[:, st:ed])
            elif self.salient_metric == "hessian":
                sal = W[:, st:ed] ** 2 / (torch.diag(H[st:ed, st:ed]).reshape((1, -1))) ** 2
            else:
                raise NotImplementedError(self.salient_metric)
            thresh = torch.sort(sal.flatten())[0][int(sal.numel() * low_frac)]
            mask[:, st:ed] = sal <= thresh
            assert lowq.groupsize % blocksize == 0
            lowq.calibrate(W[:, st:ed] * mask[:, st:ed], mask[:, st:ed], groupi=groupi)
        self.mask = mask
        if OUTPUTMASK and getattr(self.layer, "global_name", None) is not None:
            os.makedirs("./outputs/mask", exist_ok=True)
            torch.save(mask, f"./outputs/mask/mask_{low_frac}_{self.layer.global_name.replace('/', '_')}.pkl")

        lib = _lib.load()
        stream = C.c_void_p(torch.cuda.current_stream(self.dev).cuda_stream)
        mask_u8 = mask.view(torch.uint8)
        hs = highq.scale.reshape(-1).float().contiguous()
        hz = highq.zero.reshape(-1).float().contiguous()
        if hs.numel() == 1:
            hs, hz = hs.expand(self.rows).contiguous(), hz.expand(self.rows).contiguous()
        maxq = float(highq.maxq)
        tf32 = torch.backends.cuda.matmul.allow_tf32
        torch.backends.cuda.matmul.allow_tf32 = False
        try:
            for col_st in range(0, self.columns, blocksize):
                col_ed = min(col_st + blocksize, self.columns)
                groupi = col_st // lowq.groupsize
                if self.disable_gptq:                              # RTN, gptq.py:119-127
                    w = W[:, col_st:col_ed]
                    q = highq.quantize(w) * ~mask[:, col_st:col_ed] + lowq.quantize(w, groupi) * mask[:, col_st:col_ed]
                    W[:, col_st:col_ed] = q
                    continue
                nc = col_ed - col_st
                W1 = W[:, col_st:col_ed]                           # updated in place: becomes Q1
                err1 = torch.empty((self.rows, nc), device=self.dev, dtype=torch.float32)
                lm = lowq.mean[groupi].reshape(-1).float().contiguous()
                ls = lowq.scale[groupi].reshape(-1).float().contiguous()
                with torch.cuda.device(self.dev):
                    rc = lib.pbl_gptq_block(W1.data_ptr(), W.stride(0), err1.data_ptr(), nc,
                                            Hinv.data_ptr() + 4 * (col_st * Hinv.stride(0) + col_st), Hinv.stride(0),
                                            mask_u8.data_ptr() + col_st, mask_u8.stride(0), lm.data_ptr(), ls.data_ptr(),
                                            hs.data_ptr(), hz.data_ptr(), maxq, self.rows, nc, losses.data_ptr(), stream)
                _lib.check(rc, "pbl_gptq_block")
                if col_ed < self.columns:
                    W[:, col_ed:] -= err1.matmul(Hinv[col_st:col_ed, col_ed:])      # gptq.py:168
        finally:
            torch.backends.cuda.matmul.allow_tf32 = tf32
        torch.cuda.synchronize(self.dev)
        error = float(torch.sum(losses).item())
        self.time = time.time() - tick
        self.layer.weight.data = W.reshape(self.layer.weight.shape).to(self.layer.weight.data.dtype)
        return {"error": error}

    def free(self):
        self.H = None
        torch.cuda.empty_cache()
