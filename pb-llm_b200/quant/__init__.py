"""Drop-in replacements for the reference's quant layer library (`quant/__init__.py:1-2`):
same class names, constructor signatures, attributes and BinaryInterface mixin; forward goes
through libpbllm.so instead of re-materialising w_sim and calling F.linear."""
from .quantizer import (BinaryInterface, BinaryLinear, BiRealLinear, FdaBinaryLinear, IrBinaryLinear,  # noqa: F401
                        XnorBinaryLinear, PackedFakeQuantLinear)
from .outlier_quantizer import (BinaryXnorExceptOutliersLinear, BinaryXnorExceptOutliersLinearHessian,  # noqa: F401
                                weight_quant_8bit)
