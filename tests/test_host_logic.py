"""Host-side logic of the drop-in modules, on CPU: the one-time quantisation state they build
(torch ops mirroring the reference) must equal the reference's own, fixture by fixture; the
module API surface matches SURVEY.md 8b; surgery walks models like the reference does."""
import os

import numpy as np
import pytest
import torch
import torch.nn as nn

import pbllm_b200 as pb

G = os.path.join(os.path.dirname(__file__), "golden")


def load(name):
    return np.load(os.path.join(G, name + ".npz"), allow_pickle=False)


def t(a):
    return torch.from_numpy(np.ascontiguousarray(a))


def test_xnor_and_binary_effective_weight_match_reference_exactly():
    g = load("quantizer_small")
    m = pb.XnorBinaryLinear(t(g["W"]), t(g["b"]))
    assert m.weight.dtype == torch.float32 and m.bias.dtype == torch.float32
    assert torch.equal(m.quant_weight(), t(g["wsim_Xnor"]))
    assert torch.equal(pb.IrBinaryLinear(t(g["W"]), None)._effective_weight()[0], t(g["wsim_Xnor"]))
    b = pb.BinaryLinear(t(g["W"].astype(np.float16)), None)      # ctor casts to fp32 (quantizer.py:78)
    assert b.weight.dtype == torch.float32 and b.bias is None
    assert torch.equal(pb.BinaryLinear(t(g["W"]), None)._effective_weight()[0], t(g["wsim_Binary"]))
    assert torch.equal(pb.FdaBinaryLinear(t(g["W"]), None)._effective_weight()[0], t(g["wsim_Binary"]))
    assert sorted(m.state_dict().keys()) == ["bias", "weight"]   # mask/alpha are not serialised (8c item 9)
    d = m.get_save_weight_dict()
    assert d["weight"].dtype == torch.float16 and d["weight"].device.type == "cpu"
    assert isinstance(m, pb.BinaryInterface)


@pytest.mark.parametrize("tag", ["outlier_f32_small", "outlier_f16_small", "outlier_f32_heavy", "outlier_f16_heavy"])
def test_outlier_state_matches_reference_exactly(tag):
    g = load(tag)
    m = pb.BinaryXnorExceptOutliersLinear(t(g["W"]).clone(), t(g["b"]), float(g["frac"]))
    assert m.weight.dtype == t(g["W"]).dtype                      # dtype kept (outlier_quantizer.py:38)
    m.eval()
    ws = m.binarize_except_outliers()                             # lazily runs gen_outlier_mask (:84-85)
    assert torch.equal(m.outlier_mask, t(g["mask"]))
    assert tuple(m.binary_scale.shape) == (1, 1)
    assert torch.equal(m.weight.data.float(), t(g["w8"]))         # weight overwritten by its 8-bit version (:75)
    assert torch.equal(ws.float(), t(g["wsim"]))
    assert abs(m.outlier_nbits - float(g["nbits"])) < 1e-12
    m.train()
    assert torch.equal(m.binarize_except_outliers().float(), t(g["wsim_train"]))
    w_sim, low_mask, gs = m.eval()._effective_weight()
    assert torch.equal(low_mask, ~m.outlier_mask) and gs == -1


def test_weight_quant_8bit_wraps_like_the_reference_cpu_path():
    w = torch.tensor([[-100.4, -1.0, 255.6, 300.0]])
    # zp = round(min) = -100, range = 400.4: exercise the explicit mod-256
    q = pb.weight_quant_8bit(torch.tensor([[-0.3, -0.1, 0.0, 0.2]]), simulated=False)
    ref = torch.round((torch.tensor([[-0.3, -0.1, 0.0, 0.2]]) - 0.0) / 0.5 * 255)
    assert q.dtype == torch.uint8
    assert torch.equal(q.to(torch.int32), ref.to(torch.int32) & 255)
    assert pb.weight_quant_8bit(w).shape == w.shape


def test_hessian_subclass_mask_file_semantics(tmp_path, monkeypatch):
    g = load("hessian_mask")
    monkeypatch.chdir(tmp_path)
    os.makedirs("gptq_pb/outputs/mask")
    m = pb.BinaryXnorExceptOutliersLinearHessian(t(g["W"]).clone(), None, 0.1)
    m.global_name = "synthetic/model.layers.0.q_proj"
    torch.save(t(g["low_mask"]), "gptq_pb/outputs/mask/mask_0.9_synthetic_model.layers.0.q_proj.pkl")
    m.eval()
    m.gen_outlier_mask()
    assert torch.equal(m.outlier_mask, t(g["outlier_mask"])) and m.binary_scale is None
    with pytest.raises(TypeError):                                 # eval forward before any train forward
        m.binarize_except_outliers()
    m.train()
    ws = m.binarize_except_outliers()
    assert torch.equal(ws, t(g["wsim"])) and abs(float(m.binary_scale) - float(g["binary_scale"])) < 1e-9
    m2 = pb.BinaryXnorExceptOutliersLinearHessian(t(g["W"]).clone(), None, 0.1)
    m2.global_name = "synthetic/missing"                          # file missing -> magnitude fallback (:131-133)
    m2.eval()
    m2.binarize_except_outliers()
    m3 = pb.BinaryXnorExceptOutliersLinear(t(g["W"]).clone(), None, 0.1)
    m3.eval()
    assert torch.equal(m2.binarize_except_outliers(), m3.binarize_except_outliers())


class TinyBlock(nn.Module):
    def __init__(self):
        super().__init__()
        self.q_proj = nn.Linear(64, 64, bias=False)
        self.fc1 = nn.Linear(64, 128)


class TinyModel(nn.Module):
    def __init__(self):
        super().__init__()
        self.layers = nn.ModuleList([TinyBlock(), TinyBlock()])
        self.lm_head = nn.Linear(64, 100, bias=False)


def test_surgery_mirrors_reference(tmp_path):
    torch.manual_seed(0)
    model = TinyModel()
    pb.replace_with_qlinear(model, "xnor_outlier", 0.1, model_id="tiny/")
    kinds = {n: type(m).__name__ for n, m in model.named_modules() if isinstance(m, pb.BinaryInterface)}
    assert len(kinds) == 5 and set(kinds.values()) == {"BinaryXnorExceptOutliersLinear"}   # lm_head included
    assert model.layers[0].fc1.global_name == "tiny/layers.0.fc1"
    with pytest.raises(NotImplementedError):
        pb.replace_with_qlinear(TinyModel(), "nope")
    m2 = TinyModel()
    pb.replace_with_qlinear(m2, "xnor")
    pb.save_bnn(m2, str(tmp_path / "bnn"))
    m3 = pb.load_bnn(TinyModel(), str(tmp_path / "bnn"))
    assert isinstance(m3.lm_head, pb.XnorBinaryLinear)
    assert torch.equal(m3.lm_head.weight.data, m2.lm_head.weight.data.half().float())   # fp16 on disk (quantizer.py:72)
    m4 = TinyModel()
    pb.replace_from_fakequant(m4, mask_dir=None)
    assert isinstance(m4.layers[1].q_proj, pb.PackedFakeQuantLinear)
    assert m4.layers[1].q_proj.weight.dtype == torch.float32   # dtype kept


def test_from_reference_takes_the_modules_own_materialised_weight():
    """Duck-typed stand-ins for the reference classes (the real ones cannot travel to the GPU box): the packed
    module must be built from exactly the tensor the reference module itself returns."""
    g = load("quantizer_small")

    class XnorBinaryLinear(nn.Module):            # same method name / attributes as quant/quantizer.py:172-193
        def __init__(self, w, b):
            super().__init__()
            self.weight, self.bias = nn.Parameter(w), nn.Parameter(b)

        def quant_weight(self):
            return t(g["wsim_Xnor"])

    class BinaryXnorExceptOutliersLinear(nn.Module):   # quant/outlier_quantizer.py:33-106
        def __init__(self, w):
            super().__init__()
            self.weight, self.bias = nn.Parameter(w), None
            self.outlier_mask = t(g["W"]).abs() > 0.03
            self.global_name = "m/layer"

        def binarize_except_outliers(self):
            return torch.where(self.outlier_mask, self.weight.data, self.weight.data.sign() * 0.01)

    q = pb.from_reference(XnorBinaryLinear(t(g["W"]), t(g["b"])))
    assert isinstance(q, pb.PackedFakeQuantLinear) and torch.equal(q.weight.data, t(g["wsim_Xnor"]))
    assert torch.equal(q.bias.data, t(g["b"]))
    ref = BinaryXnorExceptOutliersLinear(t(g["W"]))
    q2 = pb.from_reference(ref)
    assert torch.equal(q2.weight.data, ref.binarize_except_outliers()) and torch.equal(q2.low_mask, ~ref.outlier_mask)
    assert q2.global_name == "m/layer"
    q3 = pb.from_reference(nn.Linear(8, 4))
    assert isinstance(q3, pb.PackedFakeQuantLinear)
    with pytest.raises(TypeError):
        pb.from_reference(nn.ReLU())
