"""TEST INFRASTRUCTURE ONLY — CPU restatement (torch) of the loader / writer arithmetic either side of the denoising path.

  tensor7_from_4x4   Rigid.from_tensor_4x4(...).to_tensor_7() as data.py:107, :115 call it: rot_to_quat = top eigenvector
                     of the symmetric K matrix (RU:184-216, here `transformers==5.5.0:models/esm/openfold_utils/rigid_utils.py`);
                     the sign of the result is whatever eigh returns — compare up to sign.
  peptide_atoms      the coordinates tools/pdb.py:67-174 gives the peptide's atoms: torsion_angles_to_frames +
                     frames_and_literature_positions_to_atom14_pos (openfold.utils.feats, restated below), N / CA / C / CB
                     from the normalised frame (pdb.py:99-121), O from the next residue's N (:139-151), terminal O / OXT
                     from the psi frame (:153-174).
Pinned: tests/test_io.py checks both against `tests/golden/io_golden.pt`, produced by the unmodified reference
(tests/golden/make_golden_io.py).  Only tests/ may import this module.
"""
import torch


def _tables():
    try:
        from openfold.np import residue_constants as rc
    except ImportError:
        from transformers.models.esm.openfold_utils import residue_constants as rc
    return rc


def tensor7_from_4x4(m: torch.Tensor) -> torch.Tensor:
    rot, trans = m[..., :3, :3], m[..., :3, 3]
    (xx, xy, xz), (yx, yy, yz), (zx, zy, zz) = [[rot[..., i, j] for j in range(3)] for i in range(3)]
    k = torch.stack([
        torch.stack([xx + yy + zz, zy - yz, xz - zx, yx - xy], -1),
        torch.stack([zy - yz, xx - yy - zz, xy + yx, xz + zx], -1),
        torch.stack([xz - zx, xy + yx, yy - xx - zz, yz + zy], -1),
        torch.stack([yx - xy, xz + zx, yz + zy, zz - xx - yy], -1)], -2) / 3.0
    _, vectors = torch.linalg.eigh(k)
    return torch.cat((vectors[..., -1], trans), -1)


def _quat_to_rot(q: torch.Tensor) -> torch.Tensor:
    w, x, y, z = q.unbind(-1)
    return torch.stack([
        torch.stack([w * w + x * x - y * y - z * z, 2 * (x * y - w * z), 2 * (x * z + w * y)], -1),
        torch.stack([2 * (x * y + w * z), w * w - x * x + y * y - z * z, 2 * (y * z - w * x)], -1),
        torch.stack([2 * (x * z - w * y), 2 * (y * z + w * x), w * w - x * x - y * y + z * z], -1)], -2)


def _normalize(v: torch.Tensor) -> torch.Tensor:
    return v / v.norm(dim=-1, keepdim=True).clamp_min(1e-12)      # torch.nn.functional.normalize


def peptide_atoms(frames7: torch.Tensor, torsions: torch.Tensor, aatype: torch.Tensor, mask: torch.Tensor):
    """-> positions [B,16,15,3] (atom14 order, slot 14 = OXT), exists [B,16,15] bool; zeros on padded residues."""
    rc = _tables()
    dt = frames7.dtype
    default = torch.as_tensor(rc.restype_rigid_group_default_frame, dtype=dt)[aatype]          # [B,N,8,4,4]
    group = torch.as_tensor(rc.restype_atom14_to_rigid_group, dtype=torch.long)[aatype]          # [B,N,14]
    lit = torch.as_tensor(rc.restype_atom14_rigid_group_positions, dtype=dt)[aatype]             # [B,N,14,3]
    amask = torch.as_tensor(rc.restype_atom14_mask, dtype=torch.bool)[aatype]                    # [B,N,14]
    B, N = aatype.shape
    q, x = frames7[..., :4], frames7[..., 4:]
    R = _quat_to_rot(q)                                   # quaternion as stored (pdb.py:67-72)
    Rn = _quat_to_rot(_normalize(q))                      # normalised for the backbone atoms (pdb.py:99-102)
    # torsion_angles_to_frames: [bb = (0, 1)] + 7 torsions; rotation about x composed onto the default frames
    alpha = torch.cat((torch.tensor([0.0, 1.0], dtype=dt).expand(B, N, 1, 2), torsions), -2)     # [B,N,8,2] (sin, cos)
    s, c = alpha[..., 0], alpha[..., 1]
    one, zero = torch.ones_like(s), torch.zeros_like(s)
    rx = torch.stack([torch.stack([one, zero, zero], -1), torch.stack([zero, c, -s], -1), torch.stack([zero, s, c], -1)], -2)
    fr = default[..., :3, :3] @ rx                        # [B,N,8,3,3]
    ft = default[..., :3, 3]
    # chi2..chi4 chain through chi1
    rots, trans = [fr[:, :, k] for k in range(5)], [ft[:, :, k] for k in range(5)]
    cr, ct = fr[:, :, 4], ft[:, :, 4]
    for k in (5, 6, 7):
        ct = (cr @ ft[:, :, k].unsqueeze(-1)).squeeze(-1) + ct
        cr = cr @ fr[:, :, k]
        rots.append(cr)
        trans.append(ct)
    gr = torch.stack([R @ r for r in rots], 2)            # to global: [B,N,8,3,3]
    gt = torch.stack([(R @ t.unsqueeze(-1)).squeeze(-1) + x for t in trans], 2)
    # frames_and_literature_positions_to_atom14_pos
    idx = group[..., None, None].expand(B, N, 14, 3, 3)
    ar = torch.gather(gr, 2, idx)
    at = torch.gather(gt, 2, group[..., None].expand(B, N, 14, 3))
    atom14 = ((ar @ lit.unsqueeze(-1)).squeeze(-1) + at) * amask[..., None]
    bb = ((Rn.unsqueeze(2) @ lit.unsqueeze(-1)).squeeze(-1) + x.unsqueeze(2)) * amask[..., None]
    pos = torch.zeros(B, N, 15, 3, dtype=dt)
    pos[:, :, :14] = torch.where((group == 0)[..., None], bb, atom14)
    exists = torch.zeros(B, N, 15, dtype=torch.bool)
    exists[:, :, :14] = amask
    exists[:, :, 3] = True
    n_, ca, c_ = pos[:, :, 0], pos[:, :, 1], pos[:, :, 2]
    cac = _normalize(c_ - ca)
    nxt = torch.zeros_like(mask)
    nxt[:, :-1] = mask[:, 1:]                             # the next residue exists -> not a terminus
    n_next = torch.roll(n_, -1, dims=1)
    o_mid = c_ + _normalize(cac + _normalize(c_ - n_next)) * 1.24
    o_psi = (gr[:, :, 3] @ lit[:, :, 3].unsqueeze(-1)).squeeze(-1) + gt[:, :, 3]
    co = o_psi - c_
    proj = cac * (co * cac).sum(-1, keepdim=True)
    oxt = c_ + proj - (co - proj)
    pos[:, :, 3] = torch.where(nxt[..., None], o_mid, o_psi)
    pos[:, :, 14] = torch.where(nxt[..., None], torch.zeros_like(oxt), oxt)
    exists[:, :, 14] = ~nxt
    pos = pos * mask[..., None, None]
    exists = exists & mask[..., None]
    return pos, exists
