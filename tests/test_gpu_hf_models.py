"""Drop-in check on real HF model classes (random-init, tiny configs; no checkpoints/network):
existing OPT / LLaMA models load unchanged, the surgery swaps every nn.Linear for a packed
module (reference qat/run_qat.py:45-66), and the logits equal those of the same model with the
modules baked back into dense nn.Linear layers (reference to_regular_linear, qat/run_qat.py:69-80),
i.e. the reference's own simulated-quant forward."""
import copy

import pytest
import torch

import pbllm_b200 as pb

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def tiny_llama():
    from transformers import LlamaConfig, LlamaForCausalLM
    cfg = LlamaConfig(hidden_size=256, intermediate_size=512, num_hidden_layers=2, num_attention_heads=4,
                      num_key_value_heads=4, vocab_size=512, max_position_embeddings=128)
    return LlamaForCausalLM(cfg)


def tiny_opt():
    from transformers import OPTConfig, OPTForCausalLM
    cfg = OPTConfig(hidden_size=256, ffn_dim=512, num_hidden_layers=2, num_attention_heads=4, vocab_size=512,
                    max_position_embeddings=128, word_embed_proj_dim=256)
    return OPTForCausalLM(cfg)


@pytest.mark.parametrize("make,method", [(tiny_llama, "xnor_outlier"), (tiny_opt, "xnor_outlier"), (tiny_llama, "xnor")])
def test_hf_model_drop_in(make, method):
    torch.manual_seed(0)
    model = make().to(DEV).half().eval()
    pb.replace_with_qlinear(model, method, 0.1, model_id="tiny/")
    if method == "xnor":
        model.half()                                        # quantizer.py classes cast to fp32 in the ctor
    n_q = sum(isinstance(m, pb.BinaryInterface) for m in model.modules())
    assert n_q >= 2 * 6 + 1                                 # every nn.Linear incl. lm_head (tied or not)
    ids = torch.randint(0, 512, (2, 64), device=DEV)
    layer_err = []

    def hook(mod, inp, out):      # every packed module against F.linear over its own effective weight, fp32
        x, w = inp[0], mod.dense_weight()
        ref = torch.nn.functional.linear(x.float(), w.float(), None if mod.bias is None else mod.bias.float())
        layer_err.append(float((out.float() - ref).abs().max() / ref.abs().max().clamp_min(1e-9)))

    hooks = [m.register_forward_hook(hook) for m in model.modules() if isinstance(m, pb.BinaryInterface)]
    with torch.no_grad():
        logits = model(ids).logits                          # M = 128 tokens: two-phase prefill path (tcgen05 GEMM)
        short = model(ids[:, :3]).logits                    # M = 6 tokens: decode kernel
    for h in hooks:
        h.remove()
    assert len(layer_err) == 2 * n_q and max(layer_err) <= 1e-3, max(layer_err)   # the parity bar, layer by layer
    with torch.no_grad():
        dense = copy.deepcopy(model)
        pb.to_regular_linear(dense)                         # reference's simulated-quant model (cuBLAS fp16 GEMMs)
        assert not any(isinstance(m, pb.BinaryInterface) for m in dense.modules())
        ref = dense(ids).logits
        ref_short = dense(ids[:, :3]).logits
    # logit level: a random-init tiny model amplifies the per-layer fp16 rounding differences between two
    # correct fp32-accumulate implementations (LayerNorm / softmax over 3 tokens), so this bound is a sanity check
    assert ((logits.float() - ref.float()).abs().max() / ref.float().abs().max()) < 5e-2
    assert ((short.float() - ref_short.float()).abs().max() / ref_short.float().abs().max()) < 5e-2
    assert pb.pack_model(model, keep_latent=False) == n_q  # free the latent weights, serve from packed form only
    with torch.no_grad():
        again = model(ids).logits
    assert torch.equal(again, logits)


def test_packed_checkpoint_roundtrip(tmp_path):
    """save_packed / load_packed: the on-disk model is the packed model."""
    torch.manual_seed(1)
    model = tiny_llama().to(DEV).half().eval()
    pb.replace_with_qlinear(model, "xnor_outlier", 0.1, model_id="tiny/")
    ids = torch.randint(0, 512, (2, 16), device=DEV)
    with torch.no_grad():
        ref = model(ids).logits
    meta = pb.save_packed(model, str(tmp_path / "packed"))
    assert len(meta["layers"]) == 15
    import os
    dense_bytes = sum(i["N"] * i["K"] * 2 for i in meta["layers"].values())
    assert os.path.getsize(tmp_path / "packed" / "packed_weights.pth") < 0.5 * dense_bytes   # tiny 256-wide layers: tables weigh more than at 4096
    torch.manual_seed(1)
    fresh = tiny_llama().to(DEV).half().eval()                     # same non-linear parameters (embeddings, norms)
    pb.load_packed(fresh, str(tmp_path / "packed"))
    assert all(m.weight.numel() == 0 for m in fresh.modules() if isinstance(m, pb.BinaryInterface))
    with torch.no_grad():
        out = fresh(ids).logits
    assert torch.equal(out, ref)


def test_autocast_matches_f_linear_under_autocast():
    """Reference qat/run_qat.py:120 runs the modules under bf16 autocast: F.linear casts x and w_sim to the autocast
    dtype. The drop-in must do the same instead of raising on the dtype mismatch."""
    torch.manual_seed(1)
    lin = torch.nn.Linear(256, 192, bias=True)
    m = pb.BinaryXnorExceptOutliersLinear(lin.weight, lin.bias, 0.1).to(DEV).eval()      # fp32 module, as prepared for QAT
    x = torch.randn(2, 5, 256, device=DEV)
    with torch.no_grad():
        w_sim = m.binarize_except_outliers()
        for dt in (torch.bfloat16, torch.float16):
            with torch.autocast("cuda", dtype=dt):
                y = m(x)
                ref = torch.nn.functional.linear(x, w_sim, m.bias)
            assert y.dtype == dt == ref.dtype and y.shape == ref.shape
            hi = torch.nn.functional.linear(x.to(dt).double(), w_sim.to(dt).double(), m.bias.double())
            tol = 6e-3 if dt == torch.bfloat16 else 1e-3
            assert float((y.double() - hi).abs().max() / hi.abs().max()) <= tol
            assert float((ref.double() - hi).abs().max() / hi.abs().max()) <= 2 * tol      # cuBLAS itself, for scale
        with pytest.raises(RuntimeError):
            m(x.half())                                          # no autocast: mixed dtypes raise, like the reference


def test_lazy_packed_forward_has_no_host_sync_and_is_graph_capturable():
    """A BinaryXnorExceptOutliersLinear that still holds its latent weight (what replace_with_qlinear installs) must not
    read binary_scale back to the host on every forward: that forward is captured into a CUDA graph here, which any
    device->host sync would abort."""
    torch.manual_seed(2)
    lin = torch.nn.Linear(256, 128, bias=False).half()
    m = pb.BinaryXnorExceptOutliersLinear(lin.weight, None, 0.1).to(DEV).eval()
    x = torch.randn(4, 256, device=DEV, dtype=torch.float16)
    with torch.no_grad():
        y0 = m(x)                                                # first call packs (syncs are fine here)
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            m(x)
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, stream=s):
                y1 = m(x)
        g.replay()
        torch.cuda.synchronize()
        assert torch.equal(y0, y1)
        key = m._key()
        assert all(not isinstance(k, float) for k in key)


def test_dropped_latent_refuses_the_reference_writer():
    torch.manual_seed(3)
    model = tiny_llama().to(DEV).half().eval()
    pb.replace_with_qlinear(model, "xnor_outlier", 0.1)
    pb.pack_model(model, keep_latent=True)
    pb.surgery.get_bnn_weights(model)                            # latent kept: the reference's save path works
    pb.pack_model(model, keep_latent=False)
    with pytest.raises(RuntimeError, match="save_packed"):
        pb.surgery.get_bnn_weights(model)
