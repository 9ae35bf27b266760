"""Model surgery: install the packed modules into HF OPT/LLaMA models the way the reference
does (qat/run_qat.py:45-80, utils.py:65-124), plus the GPTQ-PB checkpoint path."""
from __future__ import annotations

import json
import os
from typing import Callable, Optional

import torch
import torch.nn as nn

from . import quant as _quant
from .quant import BinaryInterface


def _walk_linears(root: nn.Module):
    names = {name: m for name, m in root.named_modules()}
    for name, m in names.items():
        if isinstance(m, nn.Linear):
            ind = name.rfind(".")
            father = names[""] if ind == -1 else names[name[:ind]]
            yield name, m, father, name[ind + 1:]


def replace_with_qlinear(root_module: nn.Module, binarization_method: str = "xnor_outlier",
                         outlier_fraction: float = 0.1, model_id: str = "", skip: Optional[Callable] = None):
    """Reference qat/run_qat.py:45-66: every nn.Linear (lm_head included) becomes a partially
    binarized layer; `global_name = model_id + name` (:66). Extra methods "xnor" / "binary"
    cover the quantizer.py classes the same way utils.load_bnn does."""
    for name, module, father, leaf in list(_walk_linears(root_module)):
        if skip is not None and skip(name, module):
            continue
        if binarization_method == "xnor_outlier":
            q = _quant.BinaryXnorExceptOutliersLinear(module.weight, module.bias, outlier_fraction)
        elif binarization_method == "xnor_outlier_hessian":
            q = _quant.BinaryXnorExceptOutliersLinearHessian(module.weight, module.bias, outlier_fraction)
        elif binarization_method == "xnor":
            q = _quant.XnorBinaryLinear(module.weight, module.bias)
        elif binarization_method == "binary":
            q = _quant.BinaryLinear(module.weight, module.bias)
        else:
            raise NotImplementedError(binarization_method)
        setattr(father, leaf, q)
        q.global_name = model_id + name
    return root_module


def to_regular_linear(root_module: nn.Module):
    """Reference qat/run_qat.py:69-80."""
    names = {name: m for name, m in root_module.named_modules()}
    for name, m in names.items():
        if isinstance(m, BinaryInterface):
            ind = name.rfind(".")
            father = names[""] if ind == -1 else names[name[:ind]]
            setattr(father, name[ind + 1:], m.to_regular_linear())
    return root_module


def replace_from_fakequant(root_module: nn.Module, mask_dir: Optional[str] = None, low_frac: float = 0.9,
                           model_id: str = "", groupsize: int = -1, skip: Optional[Callable] = None):
    """Serve a GPTQ-PB checkpoint (plain nn.Linear layers holding fake-quant fp16 weights,
    gptq_pb/gptq.py:180-184) from the packed form. The per-layer mask files are the ones
    gptq.py:108-114 writes: {mask_dir}/mask_{low_frac}_{global_name with / -> _}.pkl."""
    for name, module, father, leaf in list(_walk_linears(root_module)):
        if skip is not None and skip(name, module):
            continue
        low_mask = None
        if mask_dir is not None:
            path = os.path.join(mask_dir, f"mask_{low_frac}_{(model_id + name).replace('/', '_')}.pkl")
            if os.path.exists(path):
                low_mask = torch.load(path)
        q = _quant.PackedFakeQuantLinear(module.weight, module.bias, low_mask, groupsize)
        q.global_name = model_id + name
        setattr(father, leaf, q)
    return root_module


@torch.no_grad()
def from_reference(module: nn.Module, low_mask: Optional[torch.Tensor] = None, groupsize: int = -1):
    """Pack straight from an instance of the REFERENCE's own module (any class of its quant/ package, or a
    plain nn.Linear holding GPTQ-PB fake-quant weights): the effective weight is the tensor that module itself
    would feed to F.linear, materialised by its own method on its own device and dtype -- the preferred
    parity path (sign bits and scales are never recomputed here).
      * XnorBinaryLinear / IrBinaryLinear / FdaBinaryLinear / BiRealLinear: `module.quant_weight()`
      * BinaryXnorExceptOutliersLinear[Hessian]: `module.binarize_except_outliers()`, mask `~module.outlier_mask`
      * BinaryLinear: `sign(weight)` (quant/quantizer.py:85);  nn.Linear: `weight` as is (+ optional mask file)
    Returns a PackedFakeQuantLinear (BiReal keeps its XNOR-popcount forward through BiRealLinear)."""
    name = type(module).__name__
    bias = getattr(module, "bias", None)
    if hasattr(module, "binarize_except_outliers"):
        w_sim = module.binarize_except_outliers()
        if low_mask is None and getattr(module, "outlier_mask", None) is not None:
            low_mask = ~module.outlier_mask
    elif hasattr(module, "quant_weight"):
        w_sim = module.quant_weight()
    elif name == "BinaryLinear":
        w_sim = module.weight.data.sign()
    elif isinstance(module, nn.Linear):
        w_sim = module.weight.data
    else:
        raise TypeError(f"from_reference: don't know how {name} materialises its effective weight")
    w_sim = w_sim.detach()
    if name == "BiRealLinear":
        q = _quant.BiRealLinear.__new__(_quant.BiRealLinear)
        nn.Module.__init__(q)
        q._init_params(module.weight, None, cast_fp32=True)
        return q
    q = _quant.PackedFakeQuantLinear(w_sim, bias, low_mask, groupsize)
    q.global_name = getattr(module, "global_name", None)
    return q


@torch.no_grad()
def pack_model(root_module: nn.Module, keep_latent: bool = False, verify: bool = False):
    """Pack every BinaryInterface module now (instead of lazily at first forward) and, by
    default, free the latent weights so the model occupies its packed size in HBM."""
    n = 0
    for m in root_module.modules():
        if isinstance(m, BinaryInterface) and hasattr(m, "pack"):
            m.pack(keep_latent=keep_latent, verify=verify)
            n += 1
    return n


class _FusedGroup:
    """The packed concatenation of sibling linears that read the same input (q/k/v, gate/up): ONE launch per group."""

    def __init__(self, packed, splits):
        self.packed, self.splits = packed, splits
        self._key, self._y = None, None

    def output(self, x, index):
        # the FIRST member always runs the fused layer (a new activation tensor may reuse the address and version of the
        # previous step's); the others take their slice when they are called with that same tensor, as HF attention / MLP
        # blocks do, and fall back to running the layer otherwise
        key = (x.data_ptr(), tuple(x.shape), tuple(x.stride()), x._version, x.dtype)
        if index == 0 or key != self._key or self._y is None:
            self._y = self.packed.forward(x)
            self._key = key
        a, b = self.splits[index]
        return self._y[..., a:b]


class FusedSiblingLinear(nn.Module, BinaryInterface):
    """Stand-in for one member of a fused sibling group (installed by fuse_siblings): the first member called with a
    given input tensor runs the fused packed layer, the others return their slice of the same output."""

    def __init__(self, group: _FusedGroup, index: int, in_features: int, out_features: int, bias, global_name=None):
        super().__init__()
        self._group, self._index = [group], index           # in a list: not a submodule / not deep-copied per member
        self.in_features, self.out_features, self.bias, self.global_name = in_features, out_features, bias, global_name
        self.weight = nn.Parameter(torch.empty(0), requires_grad=False)
        self._latent_dropped = True

    def forward(self, x):
        return self._group[0].output(x, self._index)

    def packed(self):
        return self._group[0].packed

    def dense_weight(self):
        a, b = self._group[0].splits[self._index]
        return self._group[0].packed.unpack()[a:b]

    def to_regular_linear(self):
        w = self.dense_weight()
        linear = nn.Linear(w.shape[1], w.shape[0], bias=self.bias is not None, device=w.device, dtype=w.dtype)
        linear.weight.data = w.contiguous()
        if self.bias is not None:
            linear.bias.data = self.bias.data.to(w.dtype)
        return linear


@torch.no_grad()
def fuse_siblings(root_module: nn.Module, groups=(("q_proj", "k_proj", "v_proj"), ("gate_proj", "up_proj"))):
    """Pack sibling linears that consume the same activation (HF LLaMA / OPT: q/k/v; LLaMA: gate/up) as ONE packed layer
    each, rows concatenated: 4 launches per decoder layer instead of 7 in the per-token regime, where a launch costs as
    much as the weights it streams. Works on packed 16-bit modules (after replace_* / pack_model); the members are
    replaced by FusedSiblingLinear views. Returns the number of groups fused."""
    from .packing import PackedLinear
    n = 0
    for father in list(root_module.modules()):
        for names in groups:
            mods = [getattr(father, nm, None) for nm in names]
            if not all(isinstance(m, BinaryInterface) and hasattr(m, "packed") and not isinstance(m, FusedSiblingLinear) for m in mods):
                continue
            ps = [m.packed() for m in mods]
            if len({(p.K, p.dtype, p.groupsize) for p in ps}) != 1 or not ps[0].stream_layout:
                continue
            w = torch.cat([p.unpack() for p in ps])
            low = torch.cat([p.low_mask_dense() for p in ps])
            bias = None
            if any(p.bias is not None for p in ps):
                bias = torch.cat([p.bias if p.bias is not None else torch.zeros(p.N, device=p.device) for p in ps])
            fused = PackedLinear.from_dense(w, bias, low, ps[0].groupsize)
            del w, low
            splits, o = [], 0
            for p in ps:
                splits.append((o, o + p.N))
                o += p.N
            grp = _FusedGroup(fused, splits)
            for i, (nm, m) in enumerate(zip(names, mods)):
                setattr(father, nm, FusedSiblingLinear(grp, i, m.in_features, m.out_features, m.bias, getattr(m, "global_name", None)))
            n += 1
    return n


def get_bnn_meta(model):
    """Reference utils.py:65-70."""
    return {name: m.__class__.__name__ for name, m in model.named_modules() if isinstance(m, BinaryInterface)}


def get_bnn_weights(model):
    """Reference utils.py:73-84."""
    weights = {}
    for name, m in model.named_modules():
        if isinstance(m, BinaryInterface):
            weights.update({name + "_" + k: v for k, v in m.get_save_weight_dict().items()})
    return weights


def save_bnn(model, save_path):
    """Reference utils.py:87-94: meta.json (name -> class) + weights.pth (fp16 latent weights)."""
    os.makedirs(save_path, exist_ok=True)
    with open(os.path.join(save_path, "meta.json"), "w") as f:
        json.dump(get_bnn_meta(model), f)
    torch.save(get_bnn_weights(model), os.path.join(save_path, "weights.pth"))


def load_bnn(model, load_path):
    """Reference utils.py:97-124: two-argument constructors looked up by class name."""
    with open(os.path.join(load_path, "meta.json")) as f:
        meta = json.load(f)
    weights = torch.load(os.path.join(load_path, "weights.pth"))
    for name, module, father, leaf in list(_walk_linears(model)):
        if name in meta:
            q = getattr(_quant, meta[name])(weights[name + "_weight"], weights[name + "_bias"])
            q.to(module.weight.device)
            setattr(father, leaf, q)
    return model


# ---- packed on-disk format: what save_bnn / save_pretrained only account for ----------------
_PACKED_VERSION = 2


@torch.no_grad()
def save_packed(model: nn.Module, save_path: str):
    """Write every packed module's buffers (PackedLinear.buffers(): sign words / entries / affine / bias) instead of
    16-bit latent or fake-quant weights: the checkpoint is as small as the model is in HBM (~4.3 bit/weight at
    low_frac 0.9) where the reference's save_bnn (utils.py:87-94) and save_pretrained
    (gptq_pb/run.py:315-319) store 16 bit/weight. meta.json keeps the reference's name -> class map."""
    os.makedirs(save_path, exist_ok=True)
    meta, tensors = {"version": _PACKED_VERSION, "layers": {}}, {}
    for name, m in model.named_modules():
        if isinstance(m, BinaryInterface) and hasattr(m, "packed"):
            p = m.packed()
            meta["layers"][name] = {"cls": m.__class__.__name__, "N": p.N, "K": p.K, "groupsize": p.groupsize,
                                    "dtype": str(p.dtype).replace("torch.", ""), "flags": p.flags,
                                    "bias": p.bias is not None}
            for k, v in p.buffers().items():
                if v is not None:
                    tensors[f"{name}::{k}"] = v.cpu()
    with open(os.path.join(save_path, "packed_meta.json"), "w") as f:
        json.dump(meta, f)
    torch.save(tensors, os.path.join(save_path, "packed_weights.pth"))
    return meta


@torch.no_grad()
def load_packed(model: nn.Module, load_path: str, device=None):
    """Install packed modules from a save_packed() checkpoint into a freshly constructed model (its
    nn.Linear layers are replaced by name, like utils.load_bnn does, utils.py:102-122). No latent
    weights are materialised: the modules serve straight from the loaded buffers."""
    from .packing import PackedLinear
    with open(os.path.join(load_path, "packed_meta.json")) as f:
        meta = json.load(f)
    if meta.get("version") != _PACKED_VERSION:
        raise RuntimeError("unknown packed checkpoint version")
    tensors = torch.load(os.path.join(load_path, "packed_weights.pth"))
    names = {name: m for name, m in model.named_modules()}
    for name, info in meta["layers"].items():
        old = names[name]
        dev = device if device is not None else next(old.parameters()).device
        dt = getattr(torch, info["dtype"])
        bufs = {k.split("::", 1)[1]: v.to(dev) for k, v in tensors.items() if k.startswith(name + "::")}
        bias = bufs.pop("bias", None)
        p = PackedLinear.from_buffers(info["N"], info["K"], info["groupsize"], dt, bufs, bias if info["bias"] else None,
                                      info.get("flags", 0))
        cls = getattr(_quant, info["cls"])
        q = cls.__new__(cls)
        nn.Module.__init__(q)
        q.weight = nn.Parameter(torch.empty(0, dtype=dt, device=dev), requires_grad=False)
        q.bias = None if not info["bias"] else nn.Parameter(p.bias.to(dt), requires_grad=False)
        q._packed, q._packed_key, q._latent_dropped, q._packed_cast = p, None, True, {}
        q.global_name, q.out_features, q.in_features = name, info["N"], info["K"]
        for attr, val in (("outlier_mask", None), ("binary_scale", None), ("outlier_nbits", None), ("low_mask", None),
                          ("outlier_fraction", None), ("outlier_scale", 1), ("train_outlier", False), ("printed", False),
                          ("groupsize", info["groupsize"])):
            if not hasattr(q, attr):
                setattr(q, attr, val)
        ind = name.rfind(".")
        father = names[""] if ind == -1 else names[name[:ind]]
        setattr(father, name[ind + 1:], q)
    return model
