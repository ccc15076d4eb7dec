"""CPU-side checks: C-ABI library loads and exports every declared symbol, parameter layout, host value types,
drop-in surface.  No kernel is launched here."""
import os
import re

import pytest
import torch

from pmhc_diffusion_model_b200 import _lib
from pmhc_diffusion_model_b200.diffusion.model import Model
from pmhc_diffusion_model_b200.diffusion.optimizer import DiffusionModelOptimizer, linear_schedule
from pmhc_diffusion_model_b200.diffusion.tools import angle as ang
from pmhc_diffusion_model_b200.rigid import Rigid, Rotation
from tests.helpers import GOLDEN, load_case, rel_err

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

needs_lib = pytest.mark.skipif(not os.path.isfile(_lib.LIB_PATH), reason="libpmhc_b200.so not built (run __graft_entry__.build())")


@needs_lib
def test_library_exports_every_declared_symbol():
    header = open(os.path.join(ROOT, "include", "pmhc_b200.h")).read()
    declared = set(re.findall(r"\b(pmhc_[a-z_0-9]+)\s*\(", header))
    assert declared == set(_lib.EXPORTS), declared ^ set(_lib.EXPORTS)
    lib = _lib.load()
    for name in declared:
        assert hasattr(lib, name), name


@needs_lib
def test_flat_parameter_layout_matches_state_dict():
    lib = _lib.load()
    model = Model(16, 22, 1000)
    off = 0
    for idx, (name, p) in enumerate(model.named_parameters()):
        assert lib.pmhc_param_offset(idx) == off, name
        assert lib.pmhc_param_numel(idx) == p.numel(), name
        off += p.numel()
    assert off == _lib.NPARAM
    assert lib.pmhc_param_offset(48) == -1


def test_state_dict_is_drop_in_for_model_pth():
    sd = torch.load(os.path.join(GOLDEN, "shipped_params.pt"))
    model = Model(16, 22, 1000)
    res = model.load_state_dict(sd, strict=True)
    assert not res.missing_keys and not res.unexpected_keys
    assert list(model.state_dict().keys()) == list(sd.keys())
    flat = model._flat_params()
    off = 0
    for k, v in sd.items():
        assert torch.equal(flat[off:off + v.numel()].view(v.shape), v), k
        off += v.numel()
    # parameters stay views of the flat buffer through in-place updates and .to()
    with torch.no_grad():
        model.gnn1.message_mlp[0].weight.add_(1.0)
    assert torch.equal(model._flat_params()[model_offset(model, "gnn1.message_mlp.0.weight")], sd["gnn1.message_mlp.0.weight"].reshape(-1)[0] + 1.0)
    model = model.to(torch.float32)
    assert model.gnn2.torsion_mlp[2].bias.data_ptr() == model._flat_params().data_ptr() + 4 * (_lib.NPARAM - 7)


def model_offset(model, name):
    off = 0
    for n, p in model.named_parameters():
        if n == name:
            return off
        off += p.numel()
    raise KeyError(name)


def test_unsupported_configurations_fail_loudly():
    with pytest.raises(NotImplementedError):
        Model(20, 22, 100)
    model = Model(16, 22, 100)
    case = load_case("fwd_shipped_p80.pt")
    with pytest.raises((RuntimeError, ImportError)):
        model(case["batch"], 5)  # CPU tensors: there is no CPU fallback


def test_schedule_and_reverse_coefficients():
    dm = DiffusionModelOptimizer(1000, Model(16, 22, 1000), 1e-3)
    beta, alpha, sigma = dm.get_beta_alpha_sigma(250)
    assert beta == pytest.approx(0.2) and alpha == pytest.approx(0.8 ** 0.5) and sigma == pytest.approx(0.2 ** 0.5)
    assert linear_schedule(0, 10, 0.0, 0.8) == 0.0 and linear_schedule(10, 10, 0.0, 0.8) == pytest.approx(0.8)
    assert isinstance(dm.optimizer, torch.optim.Adam)


def test_rigid_value_types():
    g = torch.Generator().manual_seed(0)
    q = torch.nn.functional.normalize(torch.randn(5, 3, 4, generator=g), dim=-1)
    x = torch.randn(5, 3, 3, generator=g)
    r = Rigid(Rotation(quats=q), x)
    t7 = r.to_tensor_7()
    assert t7.shape == (5, 3, 7) and torch.allclose(t7[..., :4], q) and torch.equal(t7[..., 4:], x)
    back = Rigid.from_tensor_7(t7)
    assert torch.equal(back.get_rots().get_quats(), t7[..., :4]) and back.shape == (5, 3)
    # 4x4 round trip: same rotation up to quaternion sign
    r2 = Rigid.from_tensor_4x4(r.to_tensor_4x4())
    q2 = r2.get_rots().get_quats()
    assert torch.allclose((q2 * q).sum(-1).abs(), torch.ones(5, 3), atol=1e-5)
    assert torch.allclose(r2.get_rots().get_rot_mats(), r.get_rots().get_rot_mats(), atol=1e-5)
    pts = torch.randn(5, 3, 3, generator=g)
    assert torch.allclose(r.invert().apply(r.apply(pts)), pts, atol=1e-4)
    assert torch.allclose(r.compose(r.invert()).get_trans(), torch.zeros(5, 3, 3), atol=1e-4)
    assert r[1].shape == (3,) and r[1, :2].get_trans().shape == (2, 3)


def test_angle_tools_match_reference_known_answers():
    case = load_case("angle_tools.pt")
    assert rel_err(ang.shoemake_quat(case["u"]), case["shoemake"]) < 1e-6
    assert rel_err(ang.multiply_sin_cos(case["sc1"], case["sc2"]), case["multiply"]) < 1e-6
    assert rel_err(ang.inverse_sin_cos(case["sc1"]), case["inverse"]) < 1e-6
    for amt in (0.3, 0.8):
        assert rel_err(ang.partial_sin_cos(case["sc1"], amt), case[f"partial_{amt}"]) < 1e-6
        got = ang.partial_rot(Rotation(quats=case["q"], normalize_quats=False), amt).get_quats()
        assert rel_err(got, case[f"partial_rot_{amt}"]) < 1e-6
    q = ang.random_quat((10, 10), torch.device("cpu"))
    assert torch.all(((q ** 2).sum(-1).sqrt() - 1.0).abs() < 1e-6)
