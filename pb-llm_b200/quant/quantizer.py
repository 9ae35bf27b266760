"""Packed drop-ins for quant/quantizer.py of the reference.

Each class keeps the reference's constructor signature `Cls(weight, bias)`, its `weight` /
`bias` parameters (fp32, reference quantizer.py:78-80,175-177), `BinaryInterface`, and the
value of `forward(x)`; what changes is HOW forward is evaluated: the effective weight w_sim is
built ONCE (with the same torch ops, in the same order and dtype as the reference, so the sign
bits and scales are the reference's own -- SURVEY.md fact 6), packed by libpbllm.so, and every
forward is a single pbl_linear_forward call. Inference only: the STE backward passes
(reference quantizer.py:8-67) are out of scope."""
from __future__ import annotations

from typing import Optional

import torch
import torch.nn as nn

from ..packing import PackedLinear


class BinaryInterface:
    """Reference quantizer.py:70-72."""

    def get_save_weight_dict(self):
        if getattr(self, "_latent_dropped", False):
            # pack(keep_latent=False) freed the latent weight: the reference's writer has nothing to write.
            raise RuntimeError(f"{type(self).__name__}: the latent weight was dropped after packing "
                               "(pack_model(keep_latent=False)); write this model with save_packed(), or keep the latent "
                               "weights (pack_model(model, keep_latent=True)) to use save_bnn / state_dict")
        return {"weight": self.weight.data.half().cpu(), "bias": self.bias}


class _PackedBase(nn.Module, BinaryInterface):
    """Shared lazy-pack machinery. Subclasses implement `_effective_weight()` returning
    (w_sim [N,K] on the weight's device, low_mask or None, groupsize)."""

    def _init_params(self, weight, bias, cast_fp32: bool):
        w = weight.data if isinstance(weight, nn.Parameter) else weight.detach()
        self.weight = nn.Parameter(w.to(torch.float32) if cast_fp32 else w, requires_grad=False)
        if bias is not None:
            b = bias.data if isinstance(bias, nn.Parameter) else bias.detach()
            self.bias = nn.Parameter(b.to(torch.float32) if cast_fp32 else b, requires_grad=False)
        else:
            self.bias = None
        self._packed: Optional[PackedLinear] = None
        self._packed_key = None
        self._packed_cast = {}       # autocast dtype -> PackedLinear of w_sim.to(dtype)
        self._latent_dropped = False
        self.global_name = None
        self.out_features, self.in_features = w.shape

    def _key(self):
        w = self.weight
        b = self.bias
        return (w.data_ptr(), w._version, w.dtype, w.device, None if b is None else (b.data_ptr(), b._version))

    def _effective_weight(self):
        raise NotImplementedError

    @torch.no_grad()
    def pack(self, keep_latent: bool = True, verify: bool = False) -> PackedLinear:
        """Build w_sim with the reference's arithmetic and pack it. keep_latent=False frees the
        latent weight afterwards (`.weight` becomes an empty parameter; `dense_weight()` still
        returns w_sim, unpacked on demand)."""
        if self._latent_dropped:
            return self._packed
        if not self.weight.is_cuda:
            raise RuntimeError(f"{type(self).__name__}: move the module to a CUDA device before packing "
                               "(pb-llm_b200 has no CPU fallback)")
        w_sim, low_mask, gs = self._effective_weight()
        self._packed = PackedLinear.from_dense(w_sim, self.bias, low_mask=low_mask, groupsize=gs, verify=verify)
        self._packed_key = self._key()
        if not keep_latent:
            self.weight = nn.Parameter(torch.empty(0, dtype=self.weight.dtype, device=self.weight.device),
                                       requires_grad=False)
            self._latent_dropped = True
        self._packed_cast = {}
        return self._packed

    def packed(self) -> PackedLinear:
        if self._latent_dropped:
            return self._packed
        if self._packed is None or self._packed_key != self._key():
            self.pack()
        return self._packed

    def dense_weight(self) -> torch.Tensor:
        """w_sim as a dense tensor, reconstructed bit-exactly from the packed form."""
        return self.packed().unpack()

    def packed_as(self, dtype: torch.dtype) -> PackedLinear:
        """The layer packed from w_sim.to(dtype): what F.linear multiplies under torch.autocast (the cast of the
        reference's materialised w_sim; casting is element-wise, so the two-level structure is kept)."""
        p = self.packed()
        if p.dtype == dtype:
            return p
        hit = self._packed_cast.get(dtype)
        if hit is not None and hit[0] is p:
            return hit[1]
        with torch.no_grad():
            w = p.unpack().to(dtype)
            low = None if p.salient_count() == 0 else p.low_mask_dense()
            q = PackedLinear.from_dense(w, p.bias, low_mask=low, groupsize=p.groupsize)
        self._packed_cast[dtype] = (p, q)
        return q

    def forward(self, x):
        # F.linear under autocast (reference qat/run_qat.py:120 trains under bf16 autocast) casts x and w_sim to the
        # autocast dtype and returns that dtype; without autocast mixed dtypes raise, like the reference.
        if x.is_cuda and torch.is_autocast_enabled("cuda"):
            dt = torch.get_autocast_dtype("cuda")
            return self.packed_as(dt).forward(x.to(dt))
        return self.packed().forward(x)

    def to_regular_linear(self):
        """Bake w_sim into a plain nn.Linear (reference outlier_quantizer.py:108-114)."""
        w = self.dense_weight()
        linear = nn.Linear(w.shape[1], w.shape[0], bias=self.bias is not None, device=w.device, dtype=w.dtype)
        linear.weight.data = w
        if self.bias is not None:
            linear.bias.data = self.bias.data.to(w.dtype)
        return linear

    def extra_repr(self):
        return f"in_features={self.in_features}, out_features={self.out_features}, bias={self.bias is not None}"


class BinaryLinear(_PackedBase):
    """Reference quantizer.py:75-86: y = F.linear(x, sign(W), b); W cast to fp32 by the ctor."""

    def __init__(self, weight, bias) -> None:
        super().__init__()
        self._init_params(weight, bias, cast_fp32=True)

    def _effective_weight(self):
        return self.weight.data.sign(), None, -1  # STEBinary.forward, quantizer.py:18-21


class FdaBinaryLinear(BinaryLinear):
    """Reference quantizer.py:112-128: forward value identical to BinaryLinear (FdaBinary.forward
    is torch.sign, :47-51); only the backward differs, which is out of scope."""


class XnorBinaryLinear(_PackedBase):
    """Reference quantizer.py:172-193: w = W - mean_row(W); alpha = mean_row|w|;
    y = F.linear(x, sign(w) * alpha, b) -- the row mean is NOT added back."""

    def __init__(self, weight, bias) -> None:
        super().__init__()
        self._init_params(weight, bias, cast_fp32=True)

    def quant_weight(self, outlier_mask=None):
        w = self.weight.data
        w = w - w.mean(-1).view(-1, 1)                       # :183
        if outlier_mask is not None:
            w = w * (~outlier_mask)                          # :184-185
        scaling_factor = w.abs().mean(-1).view(-1, 1)        # :186
        return w.sign() * scaling_factor                     # :187-188

    def _effective_weight(self):
        return self.quant_weight(), None, -1


class IrBinaryLinear(XnorBinaryLinear):
    """Reference quantizer.py:89-109: forward value identical to XnorBinaryLinear (IrNetBinary
    forward is torch.sign, :31-38)."""


class BiRealLinear(_PackedBase):
    """Reference quantizer.py:131-169 -- the reference's only true XNOR-popcount layer: activations
    are binarized as well (`sign(x)`, :153; the polynomial STE terms :154-165 cancel in the forward
    value), weights are `mean_row|W| * sign(W)` (:138-148, no mean subtraction), and the bias is
    DROPPED (`F.linear(input, w)`, :168). Output is fp32 like the reference's (its masks promote to
    float32). Forward runs pbl_bireal_forward: popcounts over the packed sign plane."""

    def __init__(self, weight, bias) -> None:
        super().__init__()
        self._init_params(weight, bias, cast_fp32=True)

    def quant_weight(self):
        real_weights = self.weight.data
        scaling_factor = torch.mean(abs(real_weights), dim=1, keepdim=True)      # :139
        return scaling_factor * torch.sign(real_weights)                         # :141 (forward value of :143-147)

    def _effective_weight(self):
        return self.quant_weight(), None, -1

    @torch.no_grad()
    def pack(self, keep_latent: bool = True, verify: bool = False):
        p = super().pack(keep_latent, verify)
        if p.nnz and bool((p.vals[: p.nnz] != 0).any()):
            raise RuntimeError("BiRealLinear: packed layer has non-zero salient values; XNOR-popcount path needs sign weights")
        return p

    def forward(self, input):
        return self.packed().bireal_forward(input)

    def to_regular_linear(self):
        raise NotImplementedError("BiRealLinear binarizes its activations; it has no dense nn.Linear equivalent")


class PackedFakeQuantLinear(_PackedBase):
    """A plain nn.Linear holding GPTQ-PB fake-quant weights (what gptq_pb/gptq.py:180-184 writes
    and the reference then evaluates as a dense fp16 GEMM), served from the packed form.
    `low_mask` is the mask file gptq.py:108-114 saves (True = binarized); dtype is kept."""

    def __init__(self, weight, bias, low_mask=None, groupsize: int = -1) -> None:
        super().__init__()
        self._init_params(weight, bias, cast_fp32=False)
        self.low_mask = low_mask
        self.groupsize = groupsize

    @classmethod
    def from_linear(cls, linear: nn.Linear, low_mask=None, groupsize: int = -1):
        return cls(linear.weight, linear.bias, low_mask, groupsize)

    def _effective_weight(self):
        return self.weight.data, self.low_mask, self.groupsize

    @torch.no_grad()
    def pack(self, keep_latent: bool = True, verify: bool = False):
        p = super().pack(keep_latent, verify)
        if not keep_latent:
            self.low_mask = None
        return p
