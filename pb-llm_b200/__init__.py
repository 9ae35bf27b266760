"""pb-llm_b200: B200-native partially-binarized linear forward, a drop-in for the forward of
hahnyuan/PB-LLM's quant layer library (quant/quantizer.py, quant/outlier_quantizer.py).

Import as `importlib.import_module("pb-llm_b200")` or through the `pbllm_b200` alias package.
The CUDA library (lib/libpbllm.so, built from csrc/ for sm_100a) is mandatory: nothing here
falls back to PyTorch or CPU arithmetic for the forward."""
from . import _lib  # noqa: F401
from .packing import PackedLinear, pack_sizes  # noqa: F401
from .quant import (BinaryInterface, BinaryLinear, BiRealLinear, FdaBinaryLinear, IrBinaryLinear, XnorBinaryLinear,  # noqa: F401
                    BinaryXnorExceptOutliersLinear, BinaryXnorExceptOutliersLinearHessian, PackedFakeQuantLinear,
                    weight_quant_8bit)
from .surgery import (replace_with_qlinear, to_regular_linear, save_bnn, load_bnn, replace_from_fakequant,  # noqa: F401
                      pack_model, save_packed, load_packed, from_reference, fuse_siblings)

__all__ = ["PackedLinear", "pack_sizes", "BinaryInterface", "BinaryLinear", "BiRealLinear", "FdaBinaryLinear", "IrBinaryLinear",
           "XnorBinaryLinear", "BinaryXnorExceptOutliersLinear", "BinaryXnorExceptOutliersLinearHessian",
           "PackedFakeQuantLinear", "weight_quant_8bit", "replace_with_qlinear", "to_regular_linear", "save_bnn",
           "load_bnn", "replace_from_fakequant", "pack_model", "save_packed", "load_packed", "from_reference", "fuse_siblings"]
