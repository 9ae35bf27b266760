"""PackedLinear: device buffers of one packed partially-binarized linear + its pbl_layer handle.

torch is used for device memory and the current stream only; all arithmetic on the packed form
happens inside libpbllm.so (include/pbllm.h)."""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional

import torch

from . import _lib

_DT = {torch.float16: _lib.PBL_F16, torch.bfloat16: _lib.PBL_BF16, torch.float32: _lib.PBL_F32}


def _stream(dev) -> C.c_void_p:
    return C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)


DECODE_MAX_M = 16          # calls of at most this many tokens take the decode kernel (pbl_select_kernel == 4)
_decode_ws = {}            # (device index, stream) -> zero-initialised workspace of the decode kernel's cross-CTA reduction
_decode_ws_retired = []    # outgrown workspaces stay alive: captured CUDA graphs may still point at them


def _decode_workspace(dev, stream_ptr: int, nbytes: int) -> torch.Tensor:
    """Persistent per-(device, stream) workspace for pbl_linear_forward_ws / pbl_bireal_forward_ws: zeroed once here,
    every kernel that uses it leaves it zero again, so all layers on that stream share it."""
    key = (dev.index, stream_ptr)
    ws = _decode_ws.get(key)
    if ws is None or ws.numel() < nbytes:
        if ws is not None:
            _decode_ws_retired.append(ws)
        ws = torch.zeros(max(nbytes, 4 << 20), dtype=torch.uint8, device=dev)
        _decode_ws[key] = ws
    return ws


def pack_sizes(N: int, K: int, groupsize: int, dtype: torch.dtype) -> _lib.PblSizes:
    sz = _lib.PblSizes()
    _lib.check(_lib.load().pbl_pack_sizes(N, K, groupsize, _DT[dtype], C.byref(sz)), "pbl_pack_sizes")
    return sz


class PackedLinear:
    """Packed form of a dense fake-quant weight w_sim [N, K] (the tensor the reference feeds to
    F.linear: quant/quantizer.py:86,193; quant/outlier_quantizer.py:105)."""

    def __init__(self):
        self.handle = None
        self._fwd = None

    @classmethod
    def from_dense(cls, w_sim: torch.Tensor, bias: Optional[torch.Tensor] = None,
                   low_mask: Optional[torch.Tensor] = None, groupsize: int = -1, verify: bool = False,
                   decode_index: Optional[bool] = None):
        """w_sim: CUDA [N,K] fp16/bf16/fp32. low_mask: bool [N,K], True = binarized position
        (the GPTQ-PB mask-file convention); None = every position may be binarized."""
        if not w_sim.is_cuda:
            raise RuntimeError("PackedLinear.from_dense needs a CUDA tensor: pb-llm_b200 has no CPU path")
        if w_sim.dtype not in _DT:
            raise RuntimeError(f"unsupported weight dtype {w_sim.dtype}")
        if w_sim.dim() != 2:
            raise RuntimeError("w_sim must be [out_features, in_features]")
        lib = _lib.load()
        dev = w_sim.device
        w = w_sim.detach()
        if w.stride(-1) != 1:
            w = w.contiguous()
        N, K = w.shape
        self = cls()
        self._want_decode_index = decode_index
        self.N, self.K, self.dtype, self.device = N, K, w.dtype, dev
        self.groupsize = K if (groupsize is None or groupsize <= 0 or groupsize >= K) else int(groupsize)
        sz = pack_sizes(N, K, self.groupsize, w.dtype)
        self.sizes = sz
        mptr = None
        if low_mask is not None:
            if low_mask.shape != w.shape:
                raise RuntimeError("low_mask shape must equal the weight shape")
            lm = low_mask.to(device=dev).contiguous()
            lm = lm.view(torch.uint8) if lm.dtype == torch.bool else (lm != 0).view(torch.uint8)
            mptr = C.c_void_p(lm.data_ptr())
        with torch.cuda.device(dev):
            st = _stream(dev)
            self.affine = torch.empty(sz.n_pad * sz.groups * 2, dtype=torch.float32, device=dev)
            self.planes = torch.empty(sz.planes_bytes // 4, dtype=torch.int32, device=dev)
            self.vptr = torch.empty(sz.vptr_bytes // 4, dtype=torch.int32, device=dev)
            wp, ldw, dt = C.c_void_p(w.data_ptr()), w.stride(0), _DT[w.dtype]
            _lib.check(lib.pbl_pack_affine(wp, ldw, mptr, N, K, self.groupsize, dt, C.c_void_p(self.affine.data_ptr()), st),
                       "pbl_pack_affine")
            _lib.check(lib.pbl_pack_planes(wp, ldw, mptr, C.c_void_p(self.affine.data_ptr()), N, K, self.groupsize, dt,
                                           C.c_void_p(self.planes.data_ptr()), C.c_void_p(self.vptr.data_ptr()), st),
                       "pbl_pack_planes")
            self.nnz = int(self.vptr[-1].item()) & 0xFFFFFFFF
            self.vals = torch.zeros(self.nnz + 8, dtype=w.dtype, device=dev)
            _lib.check(lib.pbl_pack_vals(wp, ldw, C.c_void_p(self.planes.data_ptr()), C.c_void_p(self.vptr.data_ptr()),
                                         N, K, dt, C.c_void_p(self.vals.data_ptr()), st), "pbl_pack_vals")
            self.bias = None if bias is None else bias.detach().to(device=dev, dtype=torch.float32).contiguous()
            self._create()
            if verify:
                back = self.unpack()
                if not torch.equal(back, w):
                    raise RuntimeError("pack invariant violated: unpack(pack(w_sim)) != w_sim")
        return self

    @classmethod
    def from_buffers(cls, N, K, groupsize, dtype, planes, vptr, vals, affine, bias=None, decode_index=None):
        """Re-create from previously packed device buffers (packed checkpoints / row shards)."""
        self = cls()
        self._want_decode_index = decode_index
        self.N, self.K, self.dtype, self.device = int(N), int(K), dtype, planes.device
        self.groupsize = K if groupsize <= 0 or groupsize >= K else int(groupsize)
        self.sizes = pack_sizes(N, K, self.groupsize, dtype)
        self.planes, self.vptr, self.vals, self.affine = planes, vptr, vals, affine
        self.bias = None if bias is None else bias.to(device=planes.device, dtype=torch.float32).contiguous()
        self.nnz = int(vptr[-1].item()) & 0xFFFFFFFF
        self._create()
        return self

    def _create(self):
        self.sign_planes = None
        if self.nnz == 0:   # pure binary layer: keep a compact copy of the sign words for the XNOR-popcount path
            self.sign_planes = self.planes.view(-1, 4)[:, :2].contiguous()
        d = _lib.PblLayerDesc(self.N, self.K, self.groupsize, _DT[self.dtype], 0, self.planes.data_ptr(),
                              self.vptr.data_ptr(), self.vals.data_ptr(), self.affine.data_ptr(),
                              0 if self.bias is None else self.bias.data_ptr(),
                              0 if self.sign_planes is None else self.sign_planes.data_ptr())
        h = C.c_void_p()
        _lib.check(_lib.load().pbl_layer_create(C.byref(d), C.byref(h)), "pbl_layer_create")
        self.handle = h
        self.dsign = self.eptr = self.ent = None
        want = getattr(self, "_want_decode_index", None)
        if want is None:
            want = os.environ.get("PBL_DECODE_INDEX", "1") != "0"
        if want and self.dtype in (torch.float16, torch.bfloat16):
            self.build_decode_index()

    def build_decode_index(self):
        """Second, row-group-major view of the layer with one positioned entry per salient weight
        (pbl_decode_index_*): what the decode kernel (M <= 16) streams. Derived from the packed buffers."""
        lib = _lib.load()
        dev = self.device
        ds = _lib.PblDecodeSizes()
        _lib.check(lib.pbl_decode_index_sizes(self.handle, C.byref(ds)), "pbl_decode_index_sizes")
        with torch.cuda.device(dev):
            st = _stream(dev)
            eptr = torch.empty(ds.eptr_bytes // 4, dtype=torch.int32, device=dev)
            _lib.check(lib.pbl_decode_index_count(self.handle, C.c_void_p(eptr.data_ptr()), st), "pbl_decode_index_count")
            units = int(eptr[-1].item()) & 0xFFFFFFFF
            dsign = torch.empty(ds.dsign_bytes // 4, dtype=torch.int32, device=dev)
            ent = torch.zeros(max(units, 1) * 4, dtype=torch.int32, device=dev)
            _lib.check(lib.pbl_decode_index_fill(self.handle, C.c_void_p(eptr.data_ptr()), C.c_void_p(dsign.data_ptr()),
                                                 C.c_void_p(ent.data_ptr()), st), "pbl_decode_index_fill")
            _lib.check(lib.pbl_layer_attach_decode_index(self.handle, C.c_void_p(dsign.data_ptr()), C.c_void_p(eptr.data_ptr()),
                                                         C.c_void_p(ent.data_ptr())), "pbl_layer_attach_decode_index")
        self.dsign, self.eptr, self.ent = dsign, eptr, ent
        # one-group passes (M <= 8) and two-group passes (9..16) use different grids: take the larger need
        self._dws_bytes = max(int(lib.pbl_decode_workspace_bytes(self.handle, 8)), int(lib.pbl_decode_workspace_bytes(self.handle, DECODE_MAX_M)))

    def drop_decode_index(self):
        _lib.check(_lib.load().pbl_layer_attach_decode_index(self.handle, None, None, None), "pbl_layer_attach_decode_index")
        self.dsign = self.eptr = self.ent = None

    def decode_index_bytes(self) -> int:
        if self.ent is None:
            return 0
        return int(self.dsign.numel() * 4 + self.eptr.numel() * 4 + self.ent.numel() * 4)

    def __deepcopy__(self, memo):
        b = None if self.bias is None else self.bias.clone()
        return PackedLinear.from_buffers(self.N, self.K, self.groupsize, self.dtype, self.planes.clone(), self.vptr.clone(),
                                         self.vals.clone(), self.affine.clone(), b)

    def __getstate__(self):
        raise RuntimeError("PackedLinear holds a native handle; save its buffers() and rebuild with from_buffers()")

    def __del__(self):
        h, self.handle = getattr(self, "handle", None), None
        if h is not None and _lib._lib is not None:
            _lib._lib.pbl_layer_destroy(h)

    # -- the hot path ----------------------------------------------------------------------
    def forward(self, x: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        """y = x @ w_sim.T + bias through pbl_linear_forward on the current stream."""
        if not x.is_cuda:
            raise RuntimeError("pb-llm_b200 forward needs CUDA activations (no CPU fallback)")
        if x.dtype != self.dtype:  # the reference raises on mixed dtypes too (SURVEY 8b "Call")
            raise RuntimeError(f"activation dtype {x.dtype} != packed weight dtype {self.dtype}")
        if x.shape[-1] != self.K:
            raise RuntimeError(f"last dim of x is {x.shape[-1]}, expected in_features={self.K}")
        x2 = x if x.dim() == 2 else x.reshape(-1, self.K)
        if x2.stride(-1) != 1 or (x2.shape[0] > 1 and x2.stride(0) < self.K):
            x2 = x2.contiguous()
        M = x2.shape[0]
        y = out if out is not None else torch.empty((M, self.N), dtype=self.dtype, device=x.device)
        if M:
            dev = x.device
            if dev.index != torch.cuda.current_device():
                with torch.cuda.device(dev):
                    self._launch(x2, y, M, dev)
            else:
                self._launch(x2, y, M, dev)
        if out is not None:
            return y
        return y if x.dim() == 2 else y.view(*x.shape[:-1], self.N)

    def _launch(self, x2, y, M, dev):
        fwd = self._fwd
        if fwd is None:
            fwd = self._fwd = _lib.load().pbl_linear_forward_ws
        st = torch.cuda.current_stream(dev).cuda_stream
        ws_ptr, ws_bytes = None, 0
        if M <= DECODE_MAX_M and self.ent is not None and self._dws_bytes:   # decode kernel: persistent workspace
            ws = _decode_workspace(dev, st, self._dws_bytes)
            ws_ptr, ws_bytes = ws.data_ptr(), ws.numel()
        rc = fwd(self.handle, x2.data_ptr(), x2.stride(0) if M > 1 else self.K, y.data_ptr(), y.stride(0), M,
                 ws_ptr, ws_bytes, st)
        if rc:
            _lib.check(rc, "pbl_linear_forward")

    def bireal_forward(self, x: torch.Tensor, out: Optional[torch.Tensor] = None,
                       workspace: Optional[torch.Tensor] = None) -> torch.Tensor:
        """XNOR-popcount forward (pbl_bireal_forward): y = sign(x) @ w_sim.T in fp32, no bias. The layer
        must have been packed from alpha*sign(W) (its salient values are all exactly zero)."""
        if not x.is_cuda:
            raise RuntimeError("pb-llm_b200 forward needs CUDA activations (no CPU fallback)")
        if x.dtype not in _DT:
            raise RuntimeError(f"unsupported activation dtype {x.dtype}")
        if x.shape[-1] != self.K:
            raise RuntimeError(f"last dim of x is {x.shape[-1]}, expected in_features={self.K}")
        x2 = x if x.dim() == 2 else x.reshape(-1, self.K)
        if x2.stride(-1) != 1 or (x2.shape[0] > 1 and x2.stride(0) < self.K):
            x2 = x2.contiguous()
        M = x2.shape[0]
        y = out if out is not None else torch.empty((M, self.N), dtype=torch.float32, device=x.device)
        if M:
            lib = _lib.load()
            with torch.cuda.device(x.device):
                ws = workspace if workspace is not None else \
                    torch.empty(int(lib.pbl_bireal_workspace(self.handle, M)), dtype=torch.uint8, device=x.device)
                st = torch.cuda.current_stream(x.device).cuda_stream
                fix_ptr, fix_bytes = None, int(lib.pbl_bireal_fixup_workspace(self.handle, M))
                if fix_bytes:                        # stream-K XNOR kernel: the persistent zeroed reduction workspace
                    fws = _decode_workspace(x.device, st, fix_bytes)
                    fix_ptr, fix_bytes = fws.data_ptr(), fws.numel()
                rc = lib.pbl_bireal_forward_ws(self.handle, x2.data_ptr(), x2.stride(0) if M > 1 else self.K, _DT[x.dtype],
                                               y.data_ptr(), y.stride(0), M, ws.data_ptr(), fix_ptr, fix_bytes, st)
            _lib.check(rc, "pbl_bireal_forward")
        if out is not None:
            return y
        return y if x.dim() == 2 else y.view(*x.shape[:-1], self.N)

    def bireal_workspace_bytes(self, M: int) -> int:
        return int(_lib.load().pbl_bireal_workspace(self.handle, M))

    def forward_host(self, x_host: torch.Tensor, y_host: torch.Tensor, workspace: torch.Tensor):
        """End-to-end form with HOST buffers (pbl_linear_forward_host): H2D, kernel, D2H, sync."""
        M = x_host.numel() // self.K
        rc = _lib.load().pbl_linear_forward_host(self.handle, C.c_void_p(x_host.data_ptr()), C.c_void_p(y_host.data_ptr()),
                                                 M, C.c_void_p(workspace.data_ptr()), _stream(self.device))
        _lib.check(rc, "pbl_linear_forward_host")

    def host_workspace_bytes(self, M: int) -> int:
        return int(_lib.load().pbl_forward_host_workspace(self.handle, M))

    def select_kernel(self, M: int) -> int:
        return int(_lib.load().pbl_select_kernel(self.handle, M))

    def unpack(self) -> torch.Tensor:
        """Dense w_sim [N,K] reconstructed bit-exactly from the packed form."""
        w = torch.empty((self.N, self.K), dtype=self.dtype, device=self.device)
        with torch.cuda.device(self.device):
            _lib.check(_lib.load().pbl_unpack(self.handle, C.c_void_p(w.data_ptr()), self.K, _stream(self.device)),
                       "pbl_unpack")
        return w

    def low_mask_dense(self) -> torch.Tensor:
        """bool [N,K], True = binarized position (the complement of the packed salient bitmap). One-time utility
        (re-packing in another dtype for autocast); decoded from the planes with torch bit ops."""
        sz = self.sizes
        pl = self.planes.view(sz.tiles_r, sz.tiles_c, _lib.TILE_ROWS, 4)[..., 2:4]            # salient words [TR,TC,128,2]
        sh = torch.arange(32, device=self.device, dtype=torch.int32)
        bits = ((pl.unsqueeze(-1) >> sh) & 1).to(torch.bool)                                     # [TR,TC,128,2,32]
        sal = bits.permute(0, 2, 1, 3, 4).reshape(sz.n_pad, sz.k_pad)
        return ~sal[: self.N, : self.K]

    # -- accounting --------------------------------------------------------------------------
    def packed_bytes(self) -> int:
        es = 4 if self.dtype == torch.float32 else 2
        return int(self.sizes.planes_bytes + self.sizes.vptr_bytes + self.sizes.affine_bytes + self.nnz * es
                   + (0 if self.bias is None else 4 * self.N))

    def bits_per_weight(self) -> float:
        return 8.0 * self.packed_bytes() / (self.N * self.K)

    def buffers(self) -> dict:
        return dict(planes=self.planes, vptr=self.vptr, vals=self.vals, affine=self.affine, bias=self.bias)
