"""TEST INFRASTRUCTURE ONLY — CPU restatement (torch fp32) of the reference hot path.

This is the parity oracle for the CUDA path: a from-scratch, functional
restatement of cmbi/pmhc-diffusion-model's denoiser, noising / reverse step and
loss.  Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s CPU-baseline
legs may import it; the product (`pmhc_diffusion_model_b200`) never does.

Parity status: PINNED.  `tests/golden/make_golden.py` runs the UNMODIFIED
reference (through `oracle/ref_shim.py`, build container only) and commits its
outputs under `tests/golden/`; `tests/test_oracle.py` checks this file against
those fixtures on every run and, when `/root/reference` is present, against the
live reference.  The reference's own two unit tests
(tests/unit/tools/test_angle.py:11-48) are restated there as well.

Citations: `model.py`, `optimizer.py`, `angle.py` are the reference's
`diffusion/model.py`, `diffusion/optimizer.py`, `diffusion/tools/angle.py`;
`RU:` is OpenFold's `rigid_utils` as shipped in
transformers==5.5.0:models/esm/openfold_utils/rigid_utils.py (the reference's
un-vendored `openfold` 0.0.1 dependency).

Frames are plain dicts here, not classes:
    {"quats": [*,4], "trans": [*,3]}     quaternion format (w first)
    {"rot_mats": [*,3,3], "trans": [*,3]} rotation-matrix format
mirroring the two storage formats of RU:Rotation, because the format decides
whether `get_quats` goes through `torch.linalg.eigh` (RU:168-202, trap T2).
"""
from __future__ import annotations

import math
import random
from typing import Dict, Optional, Tuple

import torch
import torch.nn.functional as F

Tensor = torch.Tensor
Params = Dict[str, Tensor]

N_TORSIONS = 7
MASK_SHIFT = 1e9  # model.py:11 `infinity`

# --------------------------------------------------------------------------
# quaternion / rotation algebra (RU:26-54, 145-250)
# --------------------------------------------------------------------------


def quat_mul(a: Tensor, b: Tensor) -> Tensor:
    """Hamilton product, w first (RU:205-232 `_QUAT_MULTIPLY` / `quat_multiply`)."""
    aw, ax, ay, az = a.unbind(-1)
    bw, bx, by, bz = b.unbind(-1)
    return torch.stack(
        (
            aw * bw - ax * bx - ay * by - az * bz,
            aw * bx + ax * bw + ay * bz - az * by,
            aw * by - ax * bz + ay * bw + az * bx,
            aw * bz + ax * by - ay * bx + az * bw,
        ),
        dim=-1,
    )


def quat_inv(q: Tensor) -> Tensor:
    """conj(q) / |q|^2 (RU:246-250 `invert_quat`)."""
    sign = q.new_tensor([1.0, -1.0, -1.0, -1.0])
    return q * sign / (q * q).sum(-1, keepdim=True)


def quat_to_rot(q: Tensor) -> Tensor:
    """Un-normalised quaternion -> matrix formula (RU:128-165, trap T9)."""
    a, b, c, d = q.unbind(-1)
    rows = (
        (a * a + b * b - c * c - d * d, 2 * b * c - 2 * a * d, 2 * b * d + 2 * a * c),
        (2 * b * c + 2 * a * d, a * a - b * b + c * c - d * d, 2 * c * d - 2 * a * b),
        (2 * b * d - 2 * a * c, 2 * c * d + 2 * a * b, a * a - b * b - c * c + d * d),
    )
    return torch.stack([torch.stack(r, dim=-1) for r in rows], dim=-2)


def rot_to_quat(rot: Tensor) -> Tensor:
    """Top eigenvector of the 4x4 K/3 matrix via eigh (RU:168-202); sign is LAPACK's."""
    xx, xy, xz = rot[..., 0, 0], rot[..., 0, 1], rot[..., 0, 2]
    yx, yy, yz = rot[..., 1, 0], rot[..., 1, 1], rot[..., 1, 2]
    zx, zy, zz = rot[..., 2, 0], rot[..., 2, 1], rot[..., 2, 2]
    k = torch.stack(
        (
            torch.stack((xx + yy + zz, zy - yz, xz - zx, yx - xy), -1),
            torch.stack((zy - yz, xx - yy - zz, xy + yx, xz + zx), -1),
            torch.stack((xz - zx, xy + yx, yy - xx - zz, yz + zy), -1),
            torch.stack((yx - xy, xz + zx, yz + zy, zz - xx - yy), -1),
        ),
        dim=-2,
    )
    _, vecs = torch.linalg.eigh((1.0 / 3.0) * k)
    return vecs[..., -1]


def rot_matmul(a: Tensor, b: Tensor) -> Tensor:
    """3x3 product written out term by term (RU:26-54)."""
    return torch.stack(
        [
            torch.stack(
                [a[..., i, 0] * b[..., 0, j] + a[..., i, 1] * b[..., 1, j] + a[..., i, 2] * b[..., 2, j] for j in range(3)],
                dim=-1,
            )
            for i in range(3)
        ],
        dim=-2,
    )


def frames_from_tensor7(t7: Tensor) -> Dict[str, Tensor]:
    """RU:1037-1045 `Rigid.from_tensor_7` — quats taken as they are (no normalisation).  fp32 is forced as in
    RU:283-287 / 780-781, except for float64 inputs (used by the tests to measure the fp32 noise floor)."""
    dt = torch.float64 if t7.dtype == torch.float64 else torch.float32
    return {"quats": t7[..., :4].to(dt), "trans": t7[..., 4:].to(dt)}


def frame_quats(fr: Dict[str, Tensor]) -> Tensor:
    """RU:471-485 `Rotation.get_quats`.

    Sign tape: eigh's eigenvector sign flips under 1-ulp changes of the matrix (measured: 1 of 64
    residues between this file and the reference on identical inputs), so a rot-mat frame may carry
    `quat_hint` — quaternions recorded from the reference run — and the eigh result is flipped to
    agree with it.  Without a hint the sign is whatever this LAPACK returns, as in the reference.
    """
    if "quats" in fr:
        return fr["quats"]
    q = rot_to_quat(fr["rot_mats"])
    hint = fr.get("quat_hint")
    if hint is not None:
        q = torch.where((q * hint).sum(-1, keepdim=True) < 0, -q, q)
    return q


def frame_rot_mats(fr: Dict[str, Tensor]) -> Tensor:
    """RU:457-469 `Rotation.get_rot_mats`."""
    if "rot_mats" in fr:
        return fr["rot_mats"]
    return quat_to_rot(fr["quats"])


def frames_to_tensor7(fr: Dict[str, Tensor]) -> Tensor:
    """RU:1022-1034 `Rigid.to_tensor_7`."""
    return torch.cat((frame_quats(fr), fr["trans"]), dim=-1)


# --------------------------------------------------------------------------
# angle tools (angle.py:33-186)
# --------------------------------------------------------------------------


def angle_to_sin_cos(angle: Tensor) -> Tensor:
    """angle.py:45-57."""
    return torch.stack((torch.sin(angle), torch.cos(angle)), dim=-1)


def sin_cos_mul(u: Tensor, v: Tensor) -> Tensor:
    """Complex product with sin = imaginary, cos = real (angle.py:139-152)."""
    us, uc = u[..., 0], u[..., 1]
    vs, vc = v[..., 0], v[..., 1]
    return torch.stack((us * vc + uc * vs, uc * vc - us * vs), dim=-1)


def sin_cos_inv(u: Tensor) -> Tensor:
    """angle.py:155-162."""
    n2 = (u * u).sum(-1, keepdim=True)
    return torch.stack((-u[..., 0], u[..., 1]), dim=-1) / n2


def sin_cos_partial(u: Tensor, amount: float) -> Tensor:
    """Scale the angle by `amount` (angle.py:165-174)."""
    u = F.normalize(u, dim=-1)
    a = torch.acos(torch.clamp(u[..., 1], -1.0, 1.0))
    a = torch.where(u[..., 0] < 0.0, -a, a)
    return torch.stack((torch.sin(a * amount), torch.cos(a * amount)), dim=-1)


def quat_partial(q: Tensor, amount: float) -> Tensor:
    """Scale the rotation angle about the normalised axis (angle.py:177-186)."""
    q = F.normalize(q, dim=-1)
    half = torch.acos(torch.clamp(q[..., :1], -1.0, 1.0))
    axis = F.normalize(q[..., 1:], dim=-1)
    return torch.cat((torch.cos(half * amount), torch.sin(half * amount) * axis), dim=-1)


def shoemake(u: Tensor) -> Tensor:
    """Uniform unit quaternions from U(0,1)^3 (angle.py:70-98)."""
    u = u.clamp(0.0, 1.0)
    th1 = 2 * math.pi * u[..., 1]
    th2 = 2 * math.pi * u[..., 2]
    r1 = torch.sqrt(1.0 - u[..., 0])
    r2 = torch.sqrt(u[..., 0])
    return torch.stack((r2 * torch.cos(th2), r1 * torch.sin(th1), r1 * torch.cos(th1), r2 * torch.sin(th2)), dim=-1)


# --------------------------------------------------------------------------
# denoiser (model.py:14-421)
# --------------------------------------------------------------------------


def _mlp2(p: Params, prefix: str, x: Tensor) -> Tensor:
    """Linear -> ReLU -> Linear, the shape of every MLP in model.py:39-81."""
    hid = F.relu(F.linear(x, p[prefix + ".0.weight"], p[prefix + ".0.bias"]))
    return F.linear(hid, p[prefix + ".2.weight"], p[prefix + ".2.bias"])


def egnn_layer(
    p: Params,
    prefix: str,
    frames: Dict[str, Tensor],
    torsions: Tensor,
    h: Tensor,
    edge: Tensor,
    mask: Tensor,
    pocket_h: Tensor,
    pocket_frames: Dict[str, Tensor],
    pocket_mask: Tensor,
    taps: Optional[dict] = None,
) -> Tuple[Dict[str, Tensor], Tensor, Tensor]:
    """One all-pairs message-passing layer (model.py:83-333).

    Shapes: frames [B,N], torsions [B,N,7,2], h [B,N,H], edge [B,N,N,E], mask [B,N],
    pocket_h [B,P,H], pocket_frames [B,P], pocket_mask [B,P].
    """
    B, N, H = h.shape
    P = pocket_h.shape[-2]
    K = N + P
    pfx = prefix + "."

    # pair mask (model.py:113-120)
    not_self = ~torch.eye(N, dtype=torch.bool)
    mask, pocket_mask = mask.bool(), pocket_mask.bool()
    pair_mask = torch.cat(
        (mask[:, :, None] & mask[:, None, :] & not_self[None], mask[:, :, None] & pocket_mask[:, None, :]), dim=-1
    )

    # node / neighbour geometry (model.py:125-133; eigh here when rot-mat format)
    q_nodes = frame_quats(frames)
    x_nodes = frames["trans"]
    q_nb = torch.cat((q_nodes, frame_quats(pocket_frames)), dim=-2)[:, None, :, :].expand(B, N, K, 4)
    x_nb = torch.cat((x_nodes, pocket_frames["trans"]), dim=-2)[:, None, :, :].expand(B, N, K, 3)

    # message (model.py:183-226): MLP over cat(h_i, h_j, e_ij), zero edge features towards the pocket
    h_i = h[:, :, None, :].expand(B, N, K, H)
    h_j = torch.cat((h, pocket_h), dim=-2)[:, None, :, :].expand(B, N, K, H)
    e_ij = torch.cat((edge.to(h.dtype), h.new_zeros(B, N, P, edge.shape[-1])), dim=-2)
    message = _mlp2(p, pfx + "message_mlp", torch.cat((h_i, h_j, e_ij), dim=-1))

    # attention over neighbours (model.py:228-245)
    qi = q_nodes[:, :, None, :]
    d2 = torch.square(x_nodes[:, :, None, :] - x_nb).sum(-1)
    qdot2 = torch.square((qi * q_nb).sum(-1))
    logits = _mlp2(p, pfx + "attention_mlp", torch.cat((message, -d2[..., None], qdot2[..., None]), dim=-1)).squeeze(-1)
    weights = torch.softmax(logits - (~pair_mask) * MASK_SHIFT, dim=-1)

    # node feature update (model.py:151) — UNMASKED sum over all N+P slots (trap T3)
    out_h = _mlp2(p, pfx + "feature_mlp", torch.cat((h, message.sum(dim=-2)), dim=-1))

    # rotation update (model.py:272-312)
    q_nb_inv = quat_inv(q_nb)
    local_q = quat_mul(q_nb_inv, quat_mul(qi, q_nb))
    local_delta = torch.sigmoid(_mlp2(p, pfx + "rotation_mlp", torch.cat((message, local_q), dim=-1)))  # T5: never normalised
    global_delta = quat_mul(q_nb, quat_mul(local_delta, q_nb_inv))
    delta = (global_delta * weights[..., None]).sum(dim=-2)
    has_nb = pair_mask.sum(dim=-1) > 0
    delta = torch.where(has_nb[..., None], delta, delta.new_tensor([1.0, 0.0, 0.0, 0.0]))
    delta = F.normalize(delta, dim=-1)
    new_q = quat_mul(delta, q_nodes)

    # torsion update (model.py:247-270)
    flat_t = torsions.reshape(B, N, 2 * N_TORSIONS)
    d_angle = _mlp2(p, pfx + "torsion_mlp", torch.cat((message, flat_t[:, :, None, :].expand(B, N, K, 2 * N_TORSIONS)), dim=-1))
    d_angle = (d_angle * weights[..., None]).sum(dim=-2)
    new_torsions = sin_cos_mul(angle_to_sin_cos(d_angle), torsions)

    # translation update (model.py:314-333); neighbour translations are the layer's inputs (model.py:163-177)
    scale = _mlp2(p, pfx + "translation_mlp", message)
    new_x = x_nodes + (scale * (x_nodes[:, :, None, :] - x_nb) * weights[..., None]).sum(dim=-2)

    if taps is not None:
        taps[prefix] = {"message": message, "weights": weights, "logits": logits, "pair_mask": pair_mask,
                        "delta": delta, "d_angle": d_angle, "q_in": q_nodes}

    # output quats are normalised for the next layer (model.py:181)
    new_q_unit = new_q / torch.linalg.norm(new_q, dim=-1, keepdim=True)
    return {"quats": new_q_unit, "trans": new_x}, new_torsions, out_h


def to_float64(p: Params, batch: dict) -> Tuple[Params, dict]:
    """Same problem in float64: the yardstick for how far fp32 rounding alone moves the reference's outputs."""
    def up(v):
        if isinstance(v, dict):
            return {k: up(x) for k, x in v.items()}
        return v.double() if isinstance(v, torch.Tensor) and v.is_floating_point() else v
    return {k: v.double() for k, v in p.items()}, {k: up(v) for k, v in batch.items()}


def model_forward(p: Params, batch: dict, t: int, T: int, max_len: int = 16, taps: Optional[dict] = None) -> dict:
    """Two-layer denoiser (model.py:377-421).  `batch['frames']`/`['pocket_frames']` are frame dicts."""
    feats = batch["features"]
    mask = batch["mask"]
    B, N = mask.shape
    # time feature on peptide nodes only (model.py:394-401)
    ft = torch.full((B, N, 1), t / T, dtype=feats.dtype)
    h = torch.cat((feats, ft), dim=-1)
    pocket_h = torch.cat((batch["pocket_features"], feats.new_zeros(list(batch["pocket_mask"].shape) + [1])), dim=-1)
    # one-hot relative position encoding, depth 2*max_len-1 (model.py:349-359)
    idx = torch.arange(max_len)
    rel = (max_len - 1) + (idx[:, None] - idx[None, :])
    edge = F.one_hot(rel, num_classes=2 * max_len - 1)[None].expand(B, -1, -1, -1)

    fr1, tors1, o1 = egnn_layer(p, "gnn1", batch["frames"], batch["torsions"], h, edge, mask,
                                pocket_h, batch["pocket_frames"], batch["pocket_mask"], taps)
    i1 = F.relu(o1)
    # pocket features zero-padded to the hidden width (model.py:411-412)
    pocket_i = F.pad(pocket_h, (0, i1.shape[-1] - pocket_h.shape[-1]))
    fr2, tors2, _ = egnn_layer(p, "gnn2", fr1, tors1, i1, edge, mask,
                               pocket_i, batch["pocket_frames"], batch["pocket_mask"], taps)
    return {"frames": fr2, "torsions": tors2}


# --------------------------------------------------------------------------
# diffusion process (optimizer.py:20-252)
# --------------------------------------------------------------------------

BETA_MIN = 0.0  # optimizer.py:35
BETA_MAX = 0.8  # optimizer.py:36


def beta_alpha_sigma(step: int, T: int) -> Tuple[float, float, float]:
    """Linear schedule in step/T (optimizer.py:20-21, 81-91)."""
    beta = BETA_MIN + (BETA_MAX - BETA_MIN) * (float(step) / T)
    return beta, math.sqrt(1.0 - beta), math.sqrt(beta)


def gen_noise(shape, generator: Optional[torch.Generator] = None) -> dict:
    """Draw order randn[*,3], rand[*,3], rand[*,7] (optimizer.py:93-108, angle.py:33-67)."""
    shape = list(shape)
    pos = torch.randn(shape + [3], generator=generator) * 5.0
    q = shoemake(torch.rand(shape + [3], generator=generator))
    q = q / torch.linalg.norm(q, dim=-1, keepdim=True)  # Rotation(quats=q) normalises (RU:283-287)
    tors = angle_to_sin_cos(torch.rand(shape + [N_TORSIONS], generator=generator) * 2 * math.pi)
    return {"frames": {"quats": q, "trans": pos}, "torsions": tors}


def add_noise(signal: dict, noise: dict, t: int, T: int, quat_hint: Optional[Tensor] = None) -> dict:
    """Forward noising (optimizer.py:110-138).  Result frames are ROT-MAT format (compose_r, RU:525-538)."""
    beta, alpha, sigma = beta_alpha_sigma(t, T)
    out = dict(signal)
    out["torsions"] = sin_cos_mul(sin_cos_partial(noise["torsions"], beta), signal["torsions"])
    pos = signal["frames"]["trans"] * alpha + noise["frames"]["trans"] * sigma
    rot = rot_matmul(quat_to_rot(quat_partial(frame_quats(noise["frames"]), beta)), frame_rot_mats(signal["frames"]))
    out["frames"] = {"rot_mats": rot, "trans": pos}
    if quat_hint is not None:
        out["frames"]["quat_hint"] = quat_hint
    return out


def remove_noise(zt: dict, pred: dict, t: int, s: int, T: int, fresh: dict, quat_hint: Optional[Tensor] = None) -> dict:
    """One reverse step t -> s (optimizer.py:140-193); `fresh` is the step's new noise (gen_noise at :151)."""
    beta_t, alpha_t, sigma_t = beta_alpha_sigma(t, T)
    beta_s, alpha_s, sigma_s = beta_alpha_sigma(s, T)
    alpha_ts = alpha_t / alpha_s
    var_ts = sigma_t ** 2 - sigma_s ** 2 * alpha_ts  # trap T8: alpha_ts not squared
    sigma_ts = math.sqrt(var_ts)
    sigma_t2s = sigma_ts * sigma_s / sigma_t

    pos = zt["frames"]["trans"] / alpha_ts - (pred["frames"]["trans"] * var_ts) / (alpha_ts * sigma_t) \
        + sigma_t2s * fresh["frames"]["trans"]

    undo = quat_to_rot(quat_inv(quat_partial(frame_quats(pred["frames"]), beta_t)))  # Rotation.invert on quats (RU:585-601)
    rot = rot_matmul(quat_to_rot(quat_partial(frame_quats(fresh["frames"]), beta_s)),
                     rot_matmul(undo, frame_rot_mats(zt["frames"])))

    tors = sin_cos_mul(
        sin_cos_partial(fresh["torsions"], beta_s),
        sin_cos_mul(sin_cos_inv(sin_cos_partial(pred["torsions"], beta_t)), zt["torsions"]),
    )
    out = dict(zt)
    out["frames"] = {"rot_mats": rot, "trans": pos}
    if quat_hint is not None:
        out["frames"]["quat_hint"] = quat_hint
    out["torsions"] = tors
    return out


def get_loss(true: dict, pred: dict, mask: Tensor, torsions_mask: Tensor) -> Dict[str, Tensor]:
    """Masked translation MSE + quaternion / torsion cosine deviations (optimizer.py:38-79)."""
    n_res = mask.sum(dim=-1)
    sq = torch.square(true["frames"]["trans"] - pred["frames"]["trans"]).sum(-1)
    pos_loss = (sq * mask).sum(-1) / n_res
    qt = F.normalize(frame_quats(true["frames"]), dim=-1)
    qp = F.normalize(frame_quats(pred["frames"]), dim=-1)
    rot_loss = ((1.0 - (qt * qp).sum(-1)) * mask).sum(-1) / n_res
    tt = F.normalize(true["torsions"], dim=-1)
    tp = F.normalize(pred["torsions"], dim=-1)
    tors_loss = ((1.0 - (tt * tp).sum(-1)) * torsions_mask).sum(dim=(-2, -1)) / torsions_mask.sum(dim=(-2, -1))
    return {
        "total loss": 0.1 * pos_loss + rot_loss + tors_loss,
        "positions loss": pos_loss,
        "rotations loss": rot_loss,
        "torsions loss": tors_loss,
        "rmsd": torch.sqrt(pos_loss),
    }


def train_step_loss(p: Params, batch: dict, noise: dict, t: int, T: int,
                    quat_hint: Optional[Tensor] = None) -> Tuple[Tensor, dict, dict]:
    """Noise -> denoiser -> loss of one optimize() call, without the Adam update (optimizer.py:195-222)."""
    zt = add_noise(batch, noise, t, T, quat_hint)
    pred = model_forward(p, zt, t, T)
    losses = get_loss(noise, pred, batch["mask"], batch["torsions_mask"])
    return losses["total loss"].mean(), losses, pred


def sample(p: Params, batch: dict, T: int, generator: Optional[torch.Generator] = None,
           noise_tape: Optional[list] = None, record: Optional[list] = None,
           quat_tape: Optional[Tensor] = None) -> dict:
    """T sequential reverse steps (optimizer.py:226-252).  `noise_tape[k]` (k = T - t) overrides the fresh
    noise; `quat_tape[k]` is the sign tape for the z_t entering model call k (see `frame_quats`)."""
    zt = dict(batch)
    t = T
    with torch.no_grad():
        while t > 0:
            s = t - 1
            pred = model_forward(p, zt, t, T)
            if noise_tape is not None:
                fresh = noise_tape[T - t]
            else:
                fresh = gen_noise(batch["mask"].shape, generator)
            if record is not None:
                record.append({"t": t, "zt_quats": frame_quats(zt["frames"]).clone(), "zt_trans": zt["frames"]["trans"].clone(),
                               "zt_torsions": zt["torsions"].clone()})
            hint = quat_tape[T - t + 1] if (quat_tape is not None and T - t + 1 < len(quat_tape)) else None
            zt = remove_noise(zt, pred, t, s, T, fresh, hint)
            t = s
    return zt


# --------------------------------------------------------------------------
# synthetic SwiftMHC-shaped complexes (mirrors data.py:53-117 padding rules)
# --------------------------------------------------------------------------

# number of chi angles per restype index 0..19 (ARNDCQEGHILKMFPSTWYV), used for torsions_mask[:, 3:]
_CHI_COUNT = (0, 4, 2, 2, 1, 3, 3, 0, 2, 2, 2, 4, 3, 2, 2, 1, 1, 2, 2, 1)


def synthetic_batch(B: int, peptide_len, pocket_n, P_pad: int = 80, N_pad: int = 16, seed: int = 0) -> dict:
    """Seeded synthetic complexes in the dataset's padded layout (frames as tensor_7).

    peptide_len / pocket_n: int or (lo, hi) inclusive range drawn per complex.
    Padded slots: identity frames, zero features, torsions (0, 1) where torsions_mask is False
    (data.py:65-66, 71-79, 99-102).
    """
    g = torch.Generator().manual_seed(seed)

    def draw(spec):
        if isinstance(spec, int):
            return torch.full((B,), spec, dtype=torch.long)
        lo, hi = spec
        return torch.randint(lo, hi + 1, (B,), generator=g)

    L = draw(peptide_len)
    Pn = draw(pocket_n)
    ident7 = torch.tensor([1.0, 0, 0, 0, 0, 0, 0])
    frames = ident7.repeat(B, N_pad, 1)
    pocket_frames = ident7.repeat(B, P_pad, 1)
    feats = torch.zeros(B, N_pad, 22)
    pocket_feats = torch.zeros(B, P_pad, 22)
    aatype = torch.zeros(B, N_pad, dtype=torch.long)
    pocket_aatype = torch.zeros(B, P_pad, dtype=torch.long)
    mask = torch.zeros(B, N_pad, dtype=torch.bool)
    pocket_mask = torch.zeros(B, P_pad, dtype=torch.bool)
    torsions = torch.zeros(B, N_pad, N_TORSIONS, 2)
    torsions[..., 1] = 1.0
    torsions_mask = torch.zeros(B, N_pad, N_TORSIONS, dtype=torch.bool)
    chi = torch.tensor(_CHI_COUNT)

    q_pep = F.normalize(torch.randn(B, N_pad, 4, generator=g), dim=-1)
    x_pep = torch.randn(B, N_pad, 3, generator=g) * 5.0
    q_poc = F.normalize(torch.randn(B, P_pad, 4, generator=g), dim=-1)
    x_poc = torch.randn(B, P_pad, 3, generator=g) * 10.0
    aa_pep = torch.randint(0, 20, (B, N_pad), generator=g)
    aa_poc = torch.randint(0, 20, (B, P_pad), generator=g)
    ang = torch.rand(B, N_pad, N_TORSIONS, generator=g) * 2 * math.pi

    for b in range(B):
        l, n = int(L[b]), int(Pn[b])
        frames[b, :l, :4] = q_pep[b, :l]
        frames[b, :l, 4:] = x_pep[b, :l]
        pocket_frames[b, :n, :4] = q_poc[b, :n]
        pocket_frames[b, :n, 4:] = x_poc[b, :n]
        mask[b, :l] = True
        pocket_mask[b, :n] = True
        aatype[b, :l] = aa_pep[b, :l]
        pocket_aatype[b, :n] = aa_poc[b, :n]
        feats[b, torch.arange(l), aa_pep[b, :l]] = 1.0
        pocket_feats[b, torch.arange(n), aa_poc[b, :n]] = 1.0
        tm = torch.zeros(N_pad, N_TORSIONS, dtype=torch.bool)
        tm[:l, 3:] = torch.arange(4)[None, :] < chi[aa_pep[b, :l]][:, None]
        tm[l - 1, 2] = True
        torsions_mask[b] = tm
        sc = angle_to_sin_cos(ang[b])
        torsions[b][tm] = sc[tm]

    return {
        "frames": frames, "torsions": torsions, "features": feats, "mask": mask, "aatype": aatype,
        "torsions_mask": torsions_mask, "pocket_frames": pocket_frames, "pocket_features": pocket_feats,
        "pocket_mask": pocket_mask, "pocket_aatype": pocket_aatype,
    }


def batch_to_frames(batch: dict) -> dict:
    """tensor_7 -> frame dicts, as optimize()/sample() do at optimizer.py:201-202 / :231-232."""
    out = dict(batch)
    out["frames"] = frames_from_tensor7(batch["frames"])
    out["pocket_frames"] = frames_from_tensor7(batch["pocket_frames"])
    return out


def random_params(seed: int = 0, node_input_size: int = 22, max_len: int = 16) -> Params:
    """nn.Linear-style uniform init with the reference's 48 state-dict keys/shapes (model.py:39-81, 362-371)."""
    g = torch.Generator().manual_seed(seed)
    H1, E, I, M, TR = node_input_size + 1, 2 * max_len - 1, 64, 64, 64
    p: Params = {}

    def lin(name, fan_out, fan_in):
        bound = 1.0 / math.sqrt(fan_in)
        p[name + ".weight"] = (torch.rand(fan_out, fan_in, generator=g) * 2 - 1) * bound
        p[name + ".bias"] = (torch.rand(fan_out, generator=g) * 2 - 1) * bound

    for layer, H, O in (("gnn1", H1, I), ("gnn2", I, 1)):
        lin(f"{layer}.feature_mlp.0", TR, H + M)
        lin(f"{layer}.feature_mlp.2", O, TR)
        lin(f"{layer}.message_mlp.0", TR, 2 * H + E)
        lin(f"{layer}.message_mlp.2", M, TR)
        lin(f"{layer}.attention_mlp.0", TR, M + 2)
        lin(f"{layer}.attention_mlp.2", 1, TR)
        lin(f"{layer}.translation_mlp.0", TR, M)
        lin(f"{layer}.translation_mlp.2", 1, TR)
        lin(f"{layer}.rotation_mlp.0", TR, M + 4)
        lin(f"{layer}.rotation_mlp.2", 4, TR)
        lin(f"{layer}.torsion_mlp.0", TR, M + 2 * N_TORSIONS)
        lin(f"{layer}.torsion_mlp.2", N_TORSIONS, TR)
    return p
