"""Per-layer comparison of the packed modules inside a tiny HF OPT model against dense F.linear."""
import os
import sys

import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pbllm_b200 as pb  # noqa: E402
from tests.test_gpu_hf_models import tiny_opt, tiny_llama  # noqa: E402

DEV = "cuda:0"
for make in (tiny_opt, tiny_llama):
    torch.manual_seed(0)
    model = make().to(DEV).half().eval()
    pb.replace_with_qlinear(model, "xnor_outlier", 0.1, model_id="tiny/")
    ids = torch.randint(0, 512, (2, 64), device=DEV)

    def hook(name):
        def f(mod, inp, out):
            x = inp[0]
            w = mod.dense_weight()
            ref = F.linear(x.float(), w.float(), None if mod.bias is None else mod.bias.float())
            err = float((out.float() - ref).abs().max() / ref.abs().max().clamp_min(1e-9))
            M = x.numel() // x.shape[-1]
            print(f"  {name:34s} M={M:4d} x{tuple(x.shape)} strides{x.stride()} contig={x.is_contiguous()} ptr%16={x.data_ptr() % 16} "
                  f"kernel={mod.packed().select_kernel(M)} relerr={err:.2e} xmax={float(x.abs().max()):.2f} nan={bool(torch.isnan(out).any())}")
        return f

    for n, m in model.named_modules():
        if isinstance(m, pb.BinaryInterface):
            m.register_forward_hook(hook(n))
    with torch.no_grad():
        print(make.__name__, "M=128")
        model(ids)
        print(make.__name__, "M=6")
        model(ids[:, :3])
