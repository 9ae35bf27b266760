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
def pack_model(root_module: nn.Module, keep_latent: bool = False, verify: bool = False):
    """Pack every BinaryInterface module now (instead of lazily at first forward) and, by
    default, free the latent weights so the model occupies its packed size in HBM."""
    n = 0
    for m in root_module.modules():
        if isinstance(m, BinaryInterface) and hasattr(m, "pack"):
            m.pack(keep_latent=keep_latent, verify=verify)
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
