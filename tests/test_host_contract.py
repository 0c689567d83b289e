"""CPU tests of the host-side mirror of the reference interface: registry keys, constructor / attribute contract,
state_dict key compatibility with the reference's trained checkpoint layout, and the loud failures (no CPU path,
unsupported options) that replace silent divergence."""
import types

import pytest
import torch

from transformerbasednavierstokesolver_b200 import model_dict, set_default_precision, get_default_precision
from transformerbasednavierstokesolver_b200.model import Physics_Attention as PA
from transformerbasednavierstokesolver_b200.model import Transolver_Irregular_Mesh, Transolver_Structured_Mesh_2D
from transformerbasednavierstokesolver_b200.model.SOL_Transolver_Structured_Mesh_2D import SOL_Transolver_Structured_Mesh_2D


def test_registry_keys_match_reference():
    from transformerbasednavierstokesolver_b200.model import Transolver_Structured_Mesh2D_Encoder
    for name, mod in (("Transolver_Irregular_Mesh", Transolver_Irregular_Mesh), ("Transolver_Structured_Mesh_2D", Transolver_Structured_Mesh_2D),
                      ("Transolver_Structured_Mesh2D_Encoder", Transolver_Structured_Mesh2D_Encoder)):
        assert model_dict.get_model(types.SimpleNamespace(model=name)) is mod
        assert hasattr(mod, "Model")
    from transformerbasednavierstokesolver_b200.model import Transolver_Structured_Mesh_3D
    assert model_dict.get_model(types.SimpleNamespace(model="Transolver_Structured_Mesh_3D")) is Transolver_Structured_Mesh_3D
    with pytest.raises(KeyError):
        model_dict.get_model(types.SimpleNamespace(model="Transolver_2D"))   # exp_ns.py's default is not a key in the reference either


def test_attention_module_contract(golden):
    fx = golden("pa_structured_small.pt")
    m = PA.Physics_Attention_Structured_Mesh_2D(**fx["kwargs"])
    assert set(m.state_dict().keys()) == set(fx["state"].keys())
    for k, v in m.state_dict().items():
        assert tuple(v.shape) == tuple(fx["state"][k].shape), k
    assert (m.heads, m.dim_head, m.H, m.W) == (4, 8, 6, 5)
    assert float(m.temperature.flatten()[0]) == 0.5 and tuple(m.temperature.shape) == (1, 4, 1, 1)
    fi = golden("pa_irregular_small.pt")
    mi = PA.Physics_Attention_Irregular_Mesh(**fi["kwargs"])
    assert set(mi.state_dict().keys()) == set(fi["state"].keys())
    assert isinstance(mi.in_project_x, torch.nn.Linear) and isinstance(m.in_project_x, torch.nn.Conv2d)


def test_trained_checkpoint_layout_loads_strict():
    """shape of checkpoints/ep400_sim100.pt: 8 layers, n_hidden 64, 8 heads, slice_num 32, fun_dim 10, unified_pos 1, ref 8"""
    m = Transolver_Structured_Mesh_2D.Model(space_dim=2, n_layers=8, n_hidden=64, n_head=8, fun_dim=10, out_dim=1, slice_num=32, ref=8,
                                            unified_pos=1, H=64, W=64, mlp_ratio=1)
    sd = m.state_dict()
    assert sum(v.numel() for v in sd.values()) == 714753
    assert "blocks.7.mlp2.weight" in sd and "blocks.0.Attn.to_out.0.bias" in sd and "preprocess.linear_pre.0.weight" in sd
    assert tuple(sd["preprocess.linear_pre.0.weight"].shape) == (128, 74) and "placeholder" in sd
    assert not hasattr(m, "pos") or "pos" not in sd          # the position table is a plain attribute, not a buffer


def test_loud_failures():
    mt = Transolver_Structured_Mesh_2D.Model(Time_Input=True, n_layers=1, n_hidden=16, n_head=2, slice_num=4, H=4, W=4)   # exp_plas.py:148
    assert {"time_fc.0.weight", "time_fc.0.bias", "time_fc.2.weight", "time_fc.2.bias"} <= set(mt.state_dict().keys())
    with pytest.raises(NotImplementedError):
        PA.Physics_Attention_Structured_Mesh_2D(16, heads=2, dim_head=8, kernel=5)
    m = PA.Physics_Attention_Irregular_Mesh(16, heads=2, dim_head=8, dropout=0.1, slice_num=4)
    with pytest.raises(RuntimeError, match="no CPU path"):
        m(torch.randn(1, 7, 16))
    with pytest.raises(ValueError):
        set_default_precision("fp16")
    assert get_default_precision() in ("bf16", "fp32")


def test_sol_wrapper_attributes():
    s = SOL_Transolver_Structured_Mesh_2D(space_dim=2, n_layers=1, n_hidden=16, n_head=2, fun_dim=3, slice_num=4, H=4, W=4, step=1, look_ahead=3)
    assert s.n == 3 and s.step == 1 and hasattr(s, "transolver_model")
    assert all(k.startswith("transolver_model.") for k in s.state_dict().keys())


def test_autoencoder_module_contract(golden):
    """Physics_Attention_Structured_Mesh_2D_Auto_Encoder / Transolver_Structured_Mesh2D_Encoder.Model: parameter names and
    shapes of the reference (model/Physics_Attention.py:122-151, model/Transolver_Structured_Mesh2D_Encoder.py:99-160), the
    slice-weight accessors in the reference layout [B, heads, N, slice_num], loud failure without a cache."""
    from transformerbasednavierstokesolver_b200.model import Transolver_Structured_Mesh2D_Encoder as E
    fx = golden("pa_autoencoder_small.pt")
    m = PA.Physics_Attention_Structured_Mesh_2D_Auto_Encoder(**fx["kwargs"])
    assert set(m.state_dict().keys()) == set(fx["state"].keys())
    for k, v in m.state_dict().items():
        assert tuple(v.shape) == tuple(fx["state"][k].shape), k
    assert m.slice_weights is None
    m.slice_weights = fx["w_enc"].float()
    assert tuple(m.slice_weights.shape) == tuple(fx["w_enc"].shape) and torch.equal(m.slice_weights, fx["w_enc"].float())
    m.slice_weights = None
    model = E.Model(space_dim=2, n_layers=2, n_hidden=32, n_head=1, fun_dim=1, out_dim=1, slice_num=16, ref=4, unified_pos=1, H=8, W=8)
    keys = set(model.state_dict().keys())
    assert {"blocks.1.Attn.project_slice.weight", "blocks.0.Attn.project_slice.bias", "blocks.1.mlp2.weight", "placeholder",
            "preprocess.linear_post.bias"} <= keys and "blocks.0.mlp2.weight" not in keys
    assert model.get_attention_slice() is None
    assert model.blocks[0].decode(torch.zeros(1)) is None          # reference prints and returns None for a non-last block


def test_3d_module_contract(golden):
    """Physics_Attention_Structured_Mesh_3D / Transolver_Structured_Mesh_3D.Model: parameter names and shapes of the reference
    (model/Physics_Attention.py:232-258, model/Transolver_Structured_Mesh_3D.py:78-145), grid-mismatch error."""
    from transformerbasednavierstokesolver_b200.model import Transolver_Structured_Mesh_3D as M3
    fx = golden("pa_structured3d_small.pt")
    m = PA.Physics_Attention_Structured_Mesh_3D(**fx["kwargs"])
    assert set(m.state_dict().keys()) == set(fx["state"].keys())
    for k, v in m.state_dict().items():
        assert tuple(v.shape) == tuple(fx["state"][k].shape), k
    assert (m.H, m.W, m.D, m.dim_head) == (4, 3, 5, 8) and isinstance(m.in_project_x, torch.nn.Conv3d)
    model = M3.Model(space_dim=3, n_layers=2, n_hidden=16, n_head=2, fun_dim=1, out_dim=1, slice_num=4, ref=2, unified_pos=1, H=3, W=4, D=2)
    sd = model.state_dict()
    assert tuple(sd["preprocess.linear_pre.0.weight"].shape) == (32, 1 + 8) and "blocks.1.mlp2.weight" in sd and "placeholder" in sd
    assert tuple(model.pos.shape) == (1, 3, 4, 2, 8)
