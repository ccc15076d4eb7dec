#!/usr/bin/env python
"""Generate the golden fixtures in this directory from the UNMODIFIED reference.

Run in the build container only (needs /root/reference):

    python tests/golden/make_golden.py

Every output below is produced by the reference's own classes
(`diffusion.model.Model`, `diffusion.optimizer.DiffusionModelOptimizer`,
`diffusion.tools.angle`) imported through `oracle/ref_shim.py`; the oracle
(`oracle/egnn_oracle.py`) is used here ONLY to build seeded synthetic inputs.
The shipped weights (`/root/reference/model.pth`, a data fixture) are stored once
as `shipped_params.pt`, because the GPU box has no /root/reference; fixtures
name their weights by tag ("shipped" or ("random", seed) for
`oracle.egnn_oracle.random_params(seed)`).
"""
import os
import random
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from oracle import egnn_oracle as orc  # noqa: E402  (synthetic inputs only)
from oracle import ref_shim  # noqa: E402

torch.set_num_threads(1)  # fixtures must not depend on the thread count
ref_model, ref_opt, ref_angle = ref_shim.load_reference()
from openfold.utils.rigid_utils import Rigid, Rotation  # noqa: E402  (the shim's alias)


def shipped_params():
    return torch.load(os.path.join(ref_shim.REFERENCE_ROOT, "model.pth"), map_location="cpu")


def make_model(params, T):
    m = ref_model.Model(16, 22, T)
    m.load_state_dict(params, strict=True)
    return m


def to_ref_batch(b):
    """tensor_7 -> Rigid exactly as optimizer.py:201-202 does."""
    out = dict(b)
    out["frames"] = Rigid.from_tensor_7(b["frames"])
    out["pocket_frames"] = Rigid.from_tensor_7(b["pocket_frames"])
    return out


def noise_to_ref(n):
    return {"frames": Rigid(Rotation(quats=n["q"]), n["x"]), "torsions": n["tors"]}


def draw_noise(shape, seed):
    """Reference gen_noise under a fixed torch seed; returned as plain tensors."""
    torch.manual_seed(seed)
    n = ref_opt.DiffusionModelOptimizer.gen_noise(shape, torch.device("cpu"))
    return {"q": n["frames"].get_rots().get_quats().clone(), "x": n["frames"].get_trans().clone(),
            "tors": n["torsions"].clone()}


def save(name, obj):
    # weights are stored once (shipped_params.pt) or regenerated from a seed, not per fixture
    if "params" in obj:
        obj["params"] = PARAM_TAGS[id(obj["params"])]
    path = os.path.join(HERE, name)
    torch.save(obj, path)
    print(f"wrote {name}: {os.path.getsize(path) / 1024:.1f} KiB")


def case_forward(name, params, B, L, Pn, P_pad, t, T, seed):
    batch = orc.synthetic_batch(B, L, Pn, P_pad=P_pad, seed=seed)
    model = make_model(params, T)
    taps = {}

    def hook(mod, args, out):
        taps["l1_frames"] = out[0].to_tensor_7().detach().clone()
        taps["l1_torsions"] = out[1].detach().clone()
        taps["l1_features"] = out[2].detach().clone()

    h = model.gnn1.register_forward_hook(hook)
    with torch.no_grad():
        out = model(to_ref_batch(batch), t)
    h.remove()
    save(name, {
        "params": params, "batch": batch, "t": t, "T": T,
        "out_frames": out["frames"].to_tensor_7(), "out_torsions": out["torsions"], **taps,
    })


def case_train(name, params, B, L, Pn, P_pad, t, T, seed):
    batch = orc.synthetic_batch(B, L, Pn, P_pad=P_pad, seed=seed)
    noise = draw_noise([B, 16], seed + 1000)
    model = make_model(params, T)
    dm = ref_opt.DiffusionModelOptimizer(T, model, 1e-3)
    rb = to_ref_batch(batch)
    zt = dm.add_noise(rb, noise_to_ref(noise), t)
    zt_quats = zt["frames"].get_rots().get_quats().clone()  # eigh sign of THIS LAPACK (trap T2)
    pred = model(zt, t)
    losses = dm.get_loss(noise_to_ref(noise), pred, batch["mask"], batch["torsions_mask"])
    losses["total loss"].mean().backward()
    grads = {k: (v.grad.clone() if v.grad is not None else None) for k, v in model.named_parameters()}
    save(name, {
        "params": params, "batch": batch, "noise": noise, "t": t, "T": T,
        "zt_quats": zt_quats, "zt_rot_mats": zt["frames"].get_rots().get_rot_mats().clone(),
        "zt_trans": zt["frames"].get_trans().clone(), "zt_torsions": zt["torsions"].clone(),
        "pred_frames": pred["frames"].to_tensor_7().detach(), "pred_torsions": pred["torsions"].detach(),
        "losses": {k: v.detach() for k, v in losses.items()}, "grads": grads,
    })


def case_optimize(name, params, B, L, Pn, P_pad, T, seed, steps=2):
    """Genuine DiffusionModelOptimizer.optimize() calls (Adam included) with a pinned t and noise."""
    batch = orc.synthetic_batch(B, L, Pn, P_pad=P_pad, seed=seed)
    model = make_model(params, T)
    dm = ref_opt.DiffusionModelOptimizer(T, model, 1e-3)
    from diffusion.tools.metrics import MetricsRecord
    metrics = MetricsRecord()
    ts, noises, zt_quats = [], [], []
    orig = ref_opt.DiffusionModelOptimizer.gen_noise
    for k in range(steps):
        noise = draw_noise([B, 16], seed + 2000 + k)
        ref_opt.DiffusionModelOptimizer.gen_noise = staticmethod(lambda shape, device, n=noise: noise_to_ref(n))
        random.seed(seed + k)
        t = random.randint(0, T - 1)
        # sign tape of this step's z_t (same call optimize() makes at optimizer.py:208, deterministic)
        zt_quats.append(dm.add_noise(to_ref_batch(batch), noise_to_ref(noise), t)["frames"].get_rots().get_quats().clone())
        random.seed(seed + k)
        dm.optimize(dict(batch), metrics)
        ts.append(t)
        noises.append(noise)
    ref_opt.DiffusionModelOptimizer.gen_noise = orig
    save(name, {
        "params": params, "batch": batch, "T": T, "lr": 1e-3, "ts": ts, "noises": noises, "zt_quats": zt_quats,
        "params_after": {k: v.detach().clone() for k, v in model.state_dict().items()},
        "metrics_mean": metrics.mean(),
    })


def case_reverse_step(name, params, B, L, Pn, P_pad, t, T, seed):
    """One remove_noise() call with a rot-mat format z_t, as in every sampling step after the first."""
    batch = orc.synthetic_batch(B, L, Pn, P_pad=P_pad, seed=seed)
    model = make_model(params, T)
    dm = ref_opt.DiffusionModelOptimizer(T, model, 0.0)
    rb = to_ref_batch(batch)
    start = draw_noise([B, 16], seed + 1)
    zt = dm.add_noise(rb, noise_to_ref(start), t)
    fresh = draw_noise([B, 16], seed + 2)
    orig = ref_opt.DiffusionModelOptimizer.gen_noise
    ref_opt.DiffusionModelOptimizer.gen_noise = staticmethod(lambda shape, device: noise_to_ref(fresh))
    with torch.no_grad():
        pred = model(zt, t)
        zs = dm.remove_noise(zt, pred, t, t - 1)
    ref_opt.DiffusionModelOptimizer.gen_noise = orig
    save(name, {
        "params": params, "batch": batch, "t": t, "T": T, "fresh": fresh,
        "zt_quats": zt["frames"].get_rots().get_quats().clone(), "zt_trans": zt["frames"].get_trans().clone(),
        "zt_torsions": zt["torsions"].clone(),
        "pred_frames": pred["frames"].to_tensor_7(), "pred_torsions": pred["torsions"],
        "zs_quats": zs["frames"].get_rots().get_quats().clone(), "zs_rot_mats": zs["frames"].get_rots().get_rot_mats().clone(),
        "zs_trans": zs["frames"].get_trans().clone(), "zs_torsions": zs["torsions"].clone(),
    })


def case_trajectory(name, params, B, L, Pn, P_pad, T, seed):
    """Full DiffusionModelOptimizer.sample() with a recorded noise tape and the z_t entering every model call
    (quaternions as eigh returned them = the sign tape; translations and torsions for teacher forcing)."""
    batch = orc.synthetic_batch(B, L, Pn, P_pad=P_pad, seed=seed)
    model = make_model(params, T)
    dm = ref_opt.DiffusionModelOptimizer(T, model, 0.0)
    start = draw_noise([B, 16], seed + 1)
    tape = [draw_noise([B, 16], seed + 10 + k) for k in range(T)]
    calls = {"n": 0}

    def taped(shape, device):
        n = tape[calls["n"]]
        calls["n"] += 1
        return noise_to_ref(n)

    zt_quats, zt_trans, zt_tors = [], [], []
    orig_fwd = ref_model.Model.forward

    def spy(self, b, t):
        zt_quats.append(b["frames"].get_rots().get_quats().clone())
        zt_trans.append(b["frames"].get_trans().clone())
        zt_tors.append(b["torsions"].clone())
        return orig_fwd(self, b, t)

    orig = ref_opt.DiffusionModelOptimizer.gen_noise
    ref_opt.DiffusionModelOptimizer.gen_noise = staticmethod(taped)
    ref_model.Model.forward = spy
    inp = dict(batch)
    inp["frames"] = noise_to_ref(start)["frames"].to_tensor_7()
    inp["torsions"] = start["tors"]
    with torch.no_grad():
        out = dm.sample(inp)
    ref_model.Model.forward = orig_fwd
    ref_opt.DiffusionModelOptimizer.gen_noise = orig
    assert calls["n"] == T
    save(name, {
        "params": params, "batch": batch, "T": T, "start": start,
        "tape_q": torch.stack([n["q"] for n in tape]), "tape_x": torch.stack([n["x"] for n in tape]),
        "tape_tors": torch.stack([n["tors"] for n in tape]),
        "zt_quats": torch.stack(zt_quats), "zt_trans": torch.stack(zt_trans), "zt_torsions": torch.stack(zt_tors),
        "final_quats": out["frames"].get_rots().get_quats().clone(), "final_trans": out["frames"].get_trans().clone(),
        "final_torsions": out["torsions"].clone(),
    })


def case_angle_tools(name):
    """Known answers for the angle/quaternion helpers, incl. the reference's own two unit tests."""
    g = torch.Generator().manual_seed(7)
    u = torch.rand(5, 6, 3, generator=g)
    sc1 = torch.randn(4, 7, 2, generator=g)
    sc2 = torch.randn(4, 7, 2, generator=g)
    q = torch.randn(9, 4, generator=g)
    ang = torch.tensor([3.14159265, 1.5707963, 1.0471976, 0.0, -1.0471976, -1.5707963, -3.14159265])
    save(name, {
        "u": u, "shoemake": ref_angle.shoemake_quat(u),
        "sc1": sc1, "sc2": sc2, "multiply": ref_angle.multiply_sin_cos(sc1, sc2), "inverse": ref_angle.inverse_sin_cos(sc1),
        "partial_0.3": ref_angle.partial_sin_cos(sc1, 0.3), "partial_0.8": ref_angle.partial_sin_cos(sc1, 0.8),
        "q": q, "partial_rot_0.3": ref_angle.partial_rot(Rotation(quats=q, normalize_quats=False), 0.3).get_quats(),
        "partial_rot_0.8": ref_angle.partial_rot(Rotation(quats=q, normalize_quats=False), 0.8).get_quats(),
        "angles": ang, "angle_to_sin_cos": ref_angle.angle_to_sin_cos(ang),
    })


if __name__ == "__main__":
    shipped = shipped_params()
    rnd = orc.random_params(seed=11)
    PARAM_TAGS = {id(shipped): "shipped", id(rnd): ("random", 11)}
    save("shipped_params.pt", {k: v.clone() for k, v in shipped.items()})
    case_angle_tools("angle_tools.pt")
    case_forward("fwd_shipped_p80.pt", shipped, B=4, L=(8, 12), Pn=(50, 70), P_pad=80, t=500, T=1000, seed=1)
    case_forward("fwd_random_p96.pt", rnd, B=3, L=(8, 15), Pn=(30, 96), P_pad=96, t=3, T=100, seed=2)
    case_forward("fwd_shipped_p192.pt", shipped, B=2, L=9, Pn=180, P_pad=192, t=77, T=100, seed=3)
    case_train("train_shipped_p80.pt", shipped, B=4, L=(8, 11), Pn=(55, 65), P_pad=80, t=700, T=1000, seed=4)
    case_train("train_random_p80.pt", rnd, B=3, L=(8, 15), Pn=(20, 80), P_pad=80, t=150, T=1000, seed=5)
    case_optimize("optimize_shipped_p80.pt", shipped, B=4, L=9, Pn=60, P_pad=80, T=1000, seed=6)
    case_reverse_step("reverse_step_p80.pt", shipped, B=3, L=(8, 12), Pn=60, P_pad=80, t=40, T=100, seed=7)
    case_trajectory("trajectory_T100_p80.pt", shipped, B=2, L=9, Pn=60, P_pad=80, T=100, seed=8)
