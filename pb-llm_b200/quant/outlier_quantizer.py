"""Packed drop-ins for quant/outlier_quantizer.py of the reference (the partially-binarized
layers proper). Same constructor signature, state (`outlier_mask`, `binary_scale`,
`outlier_nbits`, 8-bit `weight`), `gen_outlier_mask()` / `to_regular_linear()` API and forward
value; forward runs from the packed form through libpbllm.so."""
from __future__ import annotations

import os

import torch
import torch.nn as nn

from .quantizer import _PackedBase


def _kthvalue(w_flat, k):
    """torch.kthvalue(w_flat, k)[0] (reference :58-66); on the device through the radix-select kernel (pbl_kth_value) --
    the same element exactly, without the sort-sized temporary of the library call."""
    if w_flat.is_cuda:
        from ..packing import kth_value
        return kth_value(w_flat, k)
    return torch.kthvalue(w_flat, k)[0]


def weight_quant_8bit(w, simulated=True):
    """Reference outlier_quantizer.py:10-29, per-row asymmetric 8-bit fake-quant, INCLUDING its
    behaviour for negative rows: zp = round(min) is 0 for |w| < 0.5, so negative values reach the
    float->uint8 cast, which the reference's CPU path wraps modulo 256. The
    wrap is written out explicitly here (`& 255`) so CUDA and CPU tensors give the same codes;
    a bare `.type(torch.uint8)` saturates on CUDA and wraps on x86."""
    raw_type = w.dtype
    w_range = torch.max(w, dim=-1, keepdim=True)[0] - torch.min(w, dim=-1, keepdim=True)[0]
    w_range = w_range.type(torch.float32)
    w_zero_point = torch.round(torch.min(w, dim=-1, keepdim=True)[0])
    w_q = torch.round((w - w_zero_point) / w_range * 255)
    w_q = (w_q.to(torch.int32) & 255).to(torch.uint8)
    if simulated:
        w_q = w_q * (w_range / 255) + w_zero_point
        w_q = w_q.to(raw_type)
    return w_q


class BinaryXnorExceptOutliersLinear(_PackedBase):
    """Reference outlier_quantizer.py:33-123. dtype of weight/bias is kept (:38-40)."""

    def __init__(self, weight, bias, outlier_fraction, outlier_scale=1, train_outlier=False) -> None:
        super().__init__()
        self._init_params(weight, bias, cast_fp32=False)
        self.printed = False
        self.outlier_mask = None
        self.outlier_scale = outlier_scale
        self.outlier_fraction = outlier_fraction
        self.binary_scale = None
        self.train_outlier = train_outlier
        self.outlier_nbits = None

    # -- one-time state (reference :54-81) -----------------------------------------------------
    @torch.no_grad()
    def gen_outlier_mask(self):
        w = self.weight.data
        w_flat = w.reshape(-1)
        n = w_flat.numel()
        lower_threshold = _kthvalue(w_flat, int(n * self.outlier_fraction / 2))                  # :58-62
        upper_threshold = _kthvalue(w_flat, int(n * (1 - self.outlier_fraction / 2)))            # :63-66
        outliers = (w < lower_threshold) | (w > upper_threshold)                                  # :69
        self.outlier_mask = outliers
        self.binary_scale = w[~self.outlier_mask].abs().mean(-1).view(-1, 1)                     # :72-74, shape [1,1]
        self.weight.data = weight_quant_8bit(w)                                                   # :75
        self.calc_memory_consumption()

    @torch.no_grad()
    def calc_memory_consumption(self):
        """Reference :116-122: CSR of the uint8 codes under the mask, 8 bit per column index,
        value and row pointer, per weight. Counted without building the CSR tensor."""
        codes = weight_quant_8bit(self.weight.data, simulated=False)
        nnz = int(((codes != 0) & self.outlier_mask).sum())
        n_rows = codes.shape[0]
        self.outlier_nbits = (nnz * 8 + nnz * 8 + (n_rows + 1) * 8) / codes.numel()

    # -- effective weight (reference :83-99) ---------------------------------------------------
    @torch.no_grad()
    def binarize_except_outliers(self):
        if self.outlier_mask is None:
            self.gen_outlier_mask()
        if self.training:                                                                         # :90-93
            self.binary_scale = self.weight.data[~self.outlier_mask].abs().mean(-1).view(-1, 1)
        scaled_weight = self.weight.data * self.outlier_scale                                     # :94
        binary_weight = self.weight.data.sign() * self.binary_scale                               # :95 (TypeError if None)
        return torch.where(self.outlier_mask, scaled_weight, binary_weight)                       # :98

    def _effective_weight(self):
        w_sim = self.binarize_except_outliers()
        return w_sim, ~self.outlier_mask, -1

    def _key(self):
        # identity and version of binary_scale, never its value: reading it would be a device->host sync per forward
        # (and is illegal under CUDA-graph capture)
        bs = self.binary_scale
        return super()._key() + (self.training, None if bs is None else (id(bs), bs.data_ptr(), bs._version))

    def packed(self):
        if self._latent_dropped:
            return self._packed
        if self.training:
            self._packed = None       # train mode recomputes alpha every forward (:90-93)
        elif self.outlier_mask is not None and self.binary_scale is None:
            self.binarize_except_outliers()  # raises TypeError like the reference's eval forward (:95)
        return super().packed()

    @torch.no_grad()
    def pack(self, keep_latent: bool = True, verify: bool = False):
        p = super().pack(keep_latent, verify)
        if not keep_latent:
            self.outlier_mask = None
        return p


class BinaryXnorExceptOutliersLinearHessian(BinaryXnorExceptOutliersLinear):
    """Reference outlier_quantizer.py:126-143: the salient mask comes from the GPTQ-PB mask file
    `gptq_pb/outputs/mask/mask_{1-f}_{global_name}.pkl` (True = binarized); when the file is
    missing it falls back to the magnitude mask (:131-133). As in the reference, binary_scale
    is NOT set by the file path, so an eval-mode forward raises until one train-mode forward (or
    the fallback) has produced it."""

    mask_dir = "gptq_pb/outputs/mask"

    @torch.no_grad()
    def gen_outlier_mask(self):
        w = self.weight.data
        low_frac = 1 - self.outlier_fraction
        path = f"{self.mask_dir}/mask_{low_frac}_{str(self.global_name).replace('/', '_')}.pkl"
        if not os.path.exists(path):
            return super().gen_outlier_mask()
        mask = torch.load(path)
        self.outlier_mask = ~mask.to(w.device)                                                    # :138
        self.weight.data = weight_quant_8bit(w)                                                   # :142
        self.calc_memory_consumption()
