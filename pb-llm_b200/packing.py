"""PackedLinear: device buffers of one packed partially-binarized linear + its pbl_layer handle.

torch is used for device memory and the current stream only; all arithmetic on the packed form
happens inside libpbllm.so (include/pbllm.h).

Two layouts, chosen by the weight dtype:
  * fp16 / bf16 -- "block stream": fragment-ordered sign words + one positioned 32-bit entry per salient weight.
    The ONE resident copy: the decode kernel streams it, unpack / the prefill expansion read it back.
  * fp32 -- "planes": sign plane + salient bitmap + packed fp32 values (CUDA-core kernel, BiReal XNOR-popcount)."""
from __future__ import annotations

import ctypes as C
from typing import Optional

import torch

from . import _lib

_DT = {torch.float16: _lib.PBL_F16, torch.bfloat16: _lib.PBL_BF16, torch.float32: _lib.PBL_F32}


def _stream(dev) -> C.c_void_p:
    return C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)


DECODE_MAX_M = 16          # calls of at most this many tokens are ONE pass of the decode kernel (persistent workspace)
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


def reset_decode_workspaces():
    """Re-zero every reduction workspace (call after a kernel fault / aborted launch: the kernels rely on finding them
    all-zero and leave them so)."""
    for ws in list(_decode_ws.values()) + _decode_ws_retired:
        ws.zero_()


@torch.no_grad()
def _free_levels(w: torch.Tensor, groupsize: int, n_pad: int, groups: int) -> torch.Tensor:
    """{lo, hi} per (row, group) for a 16-bit layer packed WITHOUT a low mask, where the packer is free to choose the
    two levels (any choice round-trips exactly; it decides how many weights are salient and how well the decode
    kernel's +-1 units fit them):
      * the two most frequent values of the (row, group), when together they cover a quarter of it -- this recovers
        the binarized structure of Binary / Xnor / GPTQ-PB weights without their mask file;
      * otherwise {-s, +s} with s = mean|w| of the (row, group): nothing is binarized, every weight becomes an entry
        tau = w / s with full relative precision (row min / max would put `mid` far from the bulk of the values).
    One-time torch ops (a sort per group); returns float32 [n_pad, groups, 2]."""
    N, K = w.shape
    out = torch.zeros(n_pad, groups, 2, dtype=torch.float32, device=w.device)
    wi = w.contiguous().view(torch.int16)
    for g in range(groups):
        c0, c1 = g * groupsize, min(K, (g + 1) * groupsize)
        seg, segf = wi[:, c0:c1], w[:, c0:c1].float()
        s, _ = torch.sort(seg, dim=-1)
        cnt = torch.searchsorted(s, s, right=True) - torch.searchsorted(s, s, right=False)      # occurrences of each element
        c1st, i1 = cnt.max(dim=-1)
        a = torch.gather(s, 1, i1[:, None])
        cnt2 = torch.where(s == a, torch.zeros_like(cnt), cnt)
        c2nd, i2 = cnt2.max(dim=-1)
        b = torch.where((c2nd > 0)[:, None], torch.gather(s, 1, i2[:, None]), a)
        af, bf = a.view(w.dtype).float()[:, 0], b.view(w.dtype).float()[:, 0]
        structured = (c1st + c2nd) * 4 >= (c1 - c0)
        sc = segf.abs().mean(dim=-1).to(w.dtype).float()
        lo = torch.where(structured, torch.minimum(af, bf), -sc)
        hi = torch.where(structured, torch.maximum(af, bf), sc)
        out[:N, g, 0], out[:N, g, 1] = lo, hi
    return out.view(-1)


def kth_value(x: torch.Tensor, k: int) -> torch.Tensor:
    """Exact k-th smallest element (k 1-based) of a CUDA tensor, as torch.kthvalue(x.flatten(), k)[0]: radix select in
    libpbllm.so (pbl_kth_value), stream-ordered, no host synchronisation. Returns a 0-dim tensor of x's dtype."""
    if not x.is_cuda or x.dtype not in _DT:
        raise RuntimeError("kth_value needs a CUDA fp16 / bf16 / fp32 tensor")
    xf = x.reshape(-1)
    if xf.stride(0) != 1:
        xf = xf.contiguous()
    lib = _lib.load()
    out = torch.empty((), dtype=x.dtype, device=x.device)
    ws = torch.empty(int(lib.pbl_kth_workspace()), dtype=torch.uint8, device=x.device)
    with torch.cuda.device(x.device):
        rc = lib.pbl_kth_value(C.c_void_p(xf.data_ptr()), xf.numel(), int(k), _DT[x.dtype], C.c_void_p(out.data_ptr()),
                               C.c_void_p(ws.data_ptr()), _stream(x.device))
    if rc == -3:
        raise IndexError(_lib.last_error())         # torch.kthvalue raises IndexError for k out of range too
    _lib.check(rc, "pbl_kth_value")
    return out


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
                   low_mask: Optional[torch.Tensor] = None, groupsize: int = -1, verify: bool = False):
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
            wp, ldw, dt = C.c_void_p(w.data_ptr()), w.stride(0), _DT[w.dtype]
            if w.dtype != torch.float32 and mptr is None:
                self.affine = _free_levels(w, self.groupsize, sz.n_pad, sz.groups)     # no mask: the levels are ours to choose
            else:                               # {min, max} of the positions the mask (or the fp32 planes layout) binarizes
                self.affine = torch.empty(sz.n_pad * sz.groups * 2, dtype=torch.float32, device=dev)
                _lib.check(lib.pbl_pack_affine(wp, ldw, mptr, N, K, self.groupsize, dt, C.c_void_p(self.affine.data_ptr()), st),
                           "pbl_pack_affine")
            self.bias = None if bias is None else bias.detach().to(device=dev, dtype=torch.float32).contiguous()
            if w.dtype == torch.float32:
                self._pack_planes(lib, wp, ldw, mptr, dt, st)
            else:
                self._pack_stream(lib, wp, ldw, mptr, dt, st)
            self._create()
            if verify:
                back = self.unpack()
                if not torch.equal(back, w):
                    raise RuntimeError("pack invariant violated: unpack(pack(w_sim)) != w_sim")
        return self

    def _pack_planes(self, lib, wp, ldw, mptr, dt, st):
        """fp32 layers: planes layout (sign plane + salient bitmap + packed values), pbl_pack_*."""
        sz, dev = self.sizes, self.device
        self.planes = torch.empty(sz.planes_bytes // 4, dtype=torch.int32, device=dev)
        self.vptr = torch.empty(sz.vptr_bytes // 4, dtype=torch.int32, device=dev)
        _lib.check(lib.pbl_pack_planes(wp, ldw, mptr, C.c_void_p(self.affine.data_ptr()), self.N, self.K, self.groupsize, dt,
                                       C.c_void_p(self.planes.data_ptr()), C.c_void_p(self.vptr.data_ptr()), st),
                   "pbl_pack_planes")
        self.nnz = int(self.vptr[-1].item()) & 0xFFFFFFFF
        self.vals = torch.zeros(self.nnz + 8, dtype=self.dtype, device=dev)
        _lib.check(lib.pbl_pack_vals(wp, ldw, C.c_void_p(self.planes.data_ptr()), C.c_void_p(self.vptr.data_ptr()),
                                     self.N, self.K, dt, C.c_void_p(self.vals.data_ptr()), st), "pbl_pack_vals")

    def _pack_stream(self, lib, wp, ldw, mptr, dt, st):
        """fp16 / bf16 layers: block-stream layout (fragment-ordered sign words + positioned salient entries), pbl_stream_*."""
        dev = self.device
        ss = _lib.PblStreamSizes()
        _lib.check(lib.pbl_stream_layout(self.N, self.K, self.groupsize, dt, C.byref(ss)), "pbl_stream_layout")
        aff = C.c_void_p(self.affine.data_ptr())
        self.eptr = torch.empty(ss.eptr_bytes // 4, dtype=torch.int32, device=dev)
        stats = torch.zeros(4, dtype=torch.int32, device=dev)
        _lib.check(lib.pbl_stream_count(wp, ldw, mptr, aff, self.N, self.K, self.groupsize, dt, C.c_void_p(self.eptr.data_ptr()),
                                        C.c_void_p(stats.data_ptr()), st), "pbl_stream_count")
        host = torch.cat([self.eptr[-1:], stats]).cpu()                 # one read-back: units, exceptions, flags
        units, n_exc, self.flags = int(host[0]) & 0xFFFFFFFF, int(host[1]) & 0xFFFFFFFF, int(host[2]) & 0xFFFFFFFF
        self.fsign = torch.empty(ss.fsign_bytes // 4, dtype=torch.int32, device=dev)
        self.ent = torch.zeros(max(units, 1) * 4, dtype=torch.int32, device=dev)
        self.exc = torch.zeros(max(n_exc, 1) * 2, dtype=torch.int32, device=dev)
        self.n_exc = n_exc
        _lib.check(lib.pbl_stream_fill(wp, ldw, mptr, aff, self.N, self.K, self.groupsize, dt, C.c_void_p(self.eptr.data_ptr()),
                                       C.c_void_p(self.fsign.data_ptr()), C.c_void_p(self.ent.data_ptr()),
                                       C.c_void_p(self.exc.data_ptr()), n_exc, C.c_void_p(stats.data_ptr()), st), "pbl_stream_fill")
        self.nnz = None          # counted lazily (salient_count()): the entry list includes padding copies

    @classmethod
    def from_buffers(cls, N, K, groupsize, dtype, buffers: dict, bias=None, flags: int = 0):
        """Re-create from previously packed device buffers (packed checkpoints / deep copies): `buffers` as returned by
        buffers() -- planes/vptr/vals/affine for fp32 layers, fsign/eptr/ent/exc/affine for fp16 / bf16 layers."""
        self = cls()
        dev = buffers["affine"].device
        self.N, self.K, self.dtype, self.device = int(N), int(K), dtype, dev
        self.groupsize = K if groupsize <= 0 or groupsize >= K else int(groupsize)
        self.sizes = pack_sizes(N, K, self.groupsize, dtype)
        self.affine = buffers["affine"]
        self.bias = None if bias is None else bias.to(device=dev, dtype=torch.float32).contiguous()
        if dtype == torch.float32:
            self.planes, self.vptr, self.vals = buffers["planes"], buffers["vptr"], buffers["vals"]
            self.nnz = int(self.vptr[-1].item()) & 0xFFFFFFFF
        else:
            self.fsign, self.eptr, self.ent = buffers["fsign"], buffers["eptr"], buffers["ent"]
            exc = buffers.get("exc")
            self.n_exc = 0 if exc is None else exc.numel() // 2
            self.exc = exc if self.n_exc else torch.zeros(2, dtype=torch.int32, device=dev)
            self.flags = int(flags)
            self.nnz = None
        self._create()
        return self

    @property
    def stream_layout(self) -> bool:
        return self.dtype != torch.float32

    def _create(self):
        d = _lib.PblLayerDesc()
        d.N, d.K, d.groupsize, d.dtype = self.N, self.K, self.groupsize, _DT[self.dtype]
        d.affine = self.affine.data_ptr()
        d.bias = 0 if self.bias is None else self.bias.data_ptr()
        self.sign_planes = None
        if self.stream_layout:
            d.flags = self.flags
            d.fsign, d.eptr, d.ent = self.fsign.data_ptr(), self.eptr.data_ptr(), self.ent.data_ptr()
            d.exc, d.n_exc = (self.exc.data_ptr() if self.n_exc else 0), self.n_exc
            self.planes = self.vptr = self.vals = None
        else:
            if self.nnz == 0:   # pure binary layer: keep a compact copy of the sign words for the XNOR-popcount path
                self.sign_planes = self.planes.view(-1, 4)[:, :2].contiguous()
            d.planes, d.vptr, d.vals = self.planes.data_ptr(), self.vptr.data_ptr(), self.vals.data_ptr()
            d.sign_planes = 0 if self.sign_planes is None else self.sign_planes.data_ptr()
            self.fsign = self.eptr = self.ent = self.exc = None
            self.n_exc, self.flags = 0, 0
        h = C.c_void_p()
        _lib.check(_lib.load().pbl_layer_create(C.byref(d), C.byref(h)), "pbl_layer_create")
        self.handle = h
        self._dws_bytes = 0
        if self.stream_layout:
            lib = _lib.load()
            # one-group passes (M <= 8) and two-group passes (9..16) use different grids: take the larger need
            self._dws_bytes = max(int(lib.pbl_decode_workspace_bytes(self.handle, 8)),
                                  int(lib.pbl_decode_workspace_bytes(self.handle, DECODE_MAX_M)))

    def salient_count(self) -> int:
        """Number of salient weights (values stored exactly instead of as a sign bit)."""
        if self.nnz is None:     # stream layout: distinct slots per block (padding copies repeat a real entry)
            units = self.eptr.to(torch.int64) & 0xFFFFFFFF
            n_real = int(units[-1].item()) * 4
            if n_real == 0:
                self.nnz = 0
            else:
                blk = torch.repeat_interleave(torch.arange(units.numel() - 1, device=self.device), (units[1:] - units[:-1]) * 4)
                slot = (self.ent[:n_real].to(torch.int64) & 0xFFFFFFFF) >> 21
                self.nnz = int(torch.unique(blk * 2048 + slot).numel())
        return self.nnz

    def __deepcopy__(self, memo):
        b = None if self.bias is None else self.bias.clone()
        bufs = {k: v.clone() for k, v in self.buffers().items() if v is not None and k != "bias"}
        return PackedLinear.from_buffers(self.N, self.K, self.groupsize, self.dtype, bufs, b, self.flags)

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
        if M <= DECODE_MAX_M and self._dws_bytes:   # decode kernel, one pass: persistent zeroed workspace
            ws = _decode_workspace(dev, st, self._dws_bytes)
            ws_ptr, ws_bytes = ws.data_ptr(), ws.numel()
        rc = fwd(self.handle, x2.data_ptr(), x2.stride(0) if M > 1 else self.K, y.data_ptr(), y.stride(0), M,
                 ws_ptr, ws_bytes, st)
        if rc:
            if ws_ptr is not None:
                reset_decode_workspaces()       # a failed launch may have left tagged slots behind
            _lib.check(rc, "pbl_linear_forward")

    def bireal_forward(self, x: torch.Tensor, out: Optional[torch.Tensor] = None,
                       workspace: Optional[torch.Tensor] = None) -> torch.Tensor:
        """XNOR-popcount forward (pbl_bireal_forward): y = sign(x) @ w_sim.T in fp32, no bias. The layer
        must have been packed from alpha*sign(W) in fp32 (planes layout; its salient values are all exactly zero)."""
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
            if rc and fix_ptr is not None:
                reset_decode_workspaces()
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

    def decode_variant(self, x: torch.Tensor) -> int:
        """2 = pair kernel, 1 = block kernel, 0 = not a decode call (pbl_decode_variant) for a 2-D activation view."""
        return int(_lib.load().pbl_decode_variant(self.handle, C.c_void_p(x.data_ptr()), x.stride(0), x.shape[0]))

    def unpack(self) -> torch.Tensor:
        """Dense w_sim [N,K] reconstructed bit-exactly from the packed form."""
        w = torch.empty((self.N, self.K), dtype=self.dtype, device=self.device)
        with torch.cuda.device(self.device):
            _lib.check(_lib.load().pbl_unpack(self.handle, C.c_void_p(w.data_ptr()), self.K, _stream(self.device)),
                       "pbl_unpack")
        return w

    def low_mask_dense(self) -> torch.Tensor:
        """bool [N,K], True = binarized position (the complement of the salient set). One-time utility (re-packing in
        another dtype for autocast); decoded with torch ops from the packed buffers."""
        sz = self.sizes
        if not self.stream_layout:
            pl = self.planes.view(sz.tiles_r, sz.tiles_c, _lib.TILE_ROWS, 4)[..., 2:4]            # salient words [TR,TC,128,2]
            sh = torch.arange(32, device=self.device, dtype=torch.int32)
            bits = ((pl.unsqueeze(-1) >> sh) & 1).to(torch.bool)                                     # [TR,TC,128,2,32]
            sal = bits.permute(0, 2, 1, 3, 4).reshape(sz.n_pad, sz.k_pad)
            return ~sal[: self.N, : self.K]
        units = self.eptr.to(torch.int64) & 0xFFFFFFFF
        n_real = int(units[-1].item()) * 4
        sal = torch.zeros(sz.n_pad * sz.k_pad, dtype=torch.bool, device=self.device)
        if n_real:
            blk = torch.repeat_interleave(torch.arange(units.numel() - 1, device=self.device), (units[1:] - units[:-1]) * 4)
            slot = (self.ent[:n_real].to(torch.int64) & 0xFFFFFFFF) >> 21
            r = slot >> 6
            pc = ((slot >> 3) & 7) ^ (r & 7)
            c = 16 * ((slot >> 1) & 3) + 2 * pc + (slot & 1)
            rg, kb = blk // sz.tiles_c, blk % sz.tiles_c
            sal[(rg * 32 + r) * sz.k_pad + kb * 64 + c] = True
        return ~sal.view(sz.n_pad, sz.k_pad)[: self.N, : self.K]

    # -- accounting --------------------------------------------------------------------------
    def packed_bytes(self) -> int:
        """Bytes of EVERY device buffer this layer keeps resident (there is no second copy of anything)."""
        return int(sum(v.numel() * v.element_size() for v in self.buffers().values() if v is not None)
                   + (0 if self.sign_planes is None else self.sign_planes.numel() * 4))

    def bits_per_weight(self) -> float:
        return 8.0 * self.packed_bytes() / (self.N * self.K)

    def buffers(self) -> dict:
        if self.stream_layout:
            return dict(fsign=self.fsign, eptr=self.eptr, ent=self.ent, exc=self.exc if self.n_exc else None, affine=self.affine,
                        bias=self.bias)
        return dict(planes=self.planes, vptr=self.vptr, vals=self.vals, affine=self.affine, bias=self.bias)
