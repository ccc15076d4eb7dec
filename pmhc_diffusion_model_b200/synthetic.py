"""Seeded synthetic inputs of the reference's shapes for benchmarks and profiling runs: complexes in the dataset's padded
layout (data.py:105-117; SURVEY.md §8d "synthetic complex generator") and random-init weights with the reference's 48
state-dict keys (nn.Linear's uniform(-1/sqrt(fan_in), 1/sqrt(fan_in)), model.py:39-81, 362-371).  Host tensors; no kernel
is involved.  (The tests draw their inputs from the oracle's own generator, not from here.)"""
import math
from typing import Dict, Tuple, Union

import torch

N_TORSIONS = 7
_CHI_COUNT = (0, 4, 2, 2, 1, 3, 3, 0, 2, 2, 2, 4, 3, 2, 2, 1, 1, 2, 2, 1)     # chi angles per residue type, restypes order


def random_params(seed: int = 0, node_input_size: int = 22, max_len: int = 16) -> Dict[str, torch.Tensor]:
    g = torch.Generator().manual_seed(seed)
    h1, edge, width = node_input_size + 1, 2 * max_len - 1, 64
    p: Dict[str, torch.Tensor] = {}

    def linear(name: str, fan_out: int, fan_in: int) -> None:
        bound = 1.0 / math.sqrt(fan_in)
        p[name + ".weight"] = (torch.rand(fan_out, fan_in, generator=g) * 2 - 1) * bound
        p[name + ".bias"] = (torch.rand(fan_out, generator=g) * 2 - 1) * bound

    for layer, h, out in (("gnn1", h1, width), ("gnn2", width, 1)):
        linear(f"{layer}.feature_mlp.0", width, h + width)
        linear(f"{layer}.feature_mlp.2", out, width)
        linear(f"{layer}.message_mlp.0", width, 2 * h + edge)
        linear(f"{layer}.message_mlp.2", width, width)
        linear(f"{layer}.attention_mlp.0", width, width + 2)
        linear(f"{layer}.attention_mlp.2", 1, width)
        linear(f"{layer}.translation_mlp.0", width, width)
        linear(f"{layer}.translation_mlp.2", 1, width)
        linear(f"{layer}.rotation_mlp.0", width, width + 4)
        linear(f"{layer}.rotation_mlp.2", 4, width)
        linear(f"{layer}.torsion_mlp.0", width, width + 2 * N_TORSIONS)
        linear(f"{layer}.torsion_mlp.2", N_TORSIONS, width)
    return p


def synthetic_batch(B: int, peptide_len: Union[int, Tuple[int, int]], pocket_n: Union[int, Tuple[int, int]], P_pad: int = 80,
                    N_pad: int = 16, seed: int = 0) -> Dict[str, torch.Tensor]:
    """peptide_len / pocket_n: a number or an inclusive (lo, hi) range drawn per complex.  Real slots: unit quaternions,
    peptide positions ~ 5 N(0, I) A, pocket positions ~ 10 N(0, I) A, uniform residue types as one-hot-22, uniform torsions
    where the torsion mask is set (chi angles the residue type has + psi of the C-terminus, data.py:91-100).  Padded slots:
    identity frames, zero features, torsions (0, 1)."""
    g = torch.Generator().manual_seed(seed)

    def draw(spec):
        if isinstance(spec, int):
            return torch.full((B,), spec, dtype=torch.long)
        return torch.randint(spec[0], spec[1] + 1, (B,), generator=g)

    L, Pn = draw(peptide_len), draw(pocket_n)
    mask = torch.arange(N_pad)[None, :] < L[:, None]
    pocket_mask = torch.arange(P_pad)[None, :] < Pn[:, None]

    def frames(n, m, spread):
        q = torch.nn.functional.normalize(torch.randn(B, n, 4, generator=g), dim=-1)
        x = torch.randn(B, n, 3, generator=g) * spread
        ident = torch.tensor([1.0, 0, 0, 0, 0, 0, 0]).expand(B, n, 7)
        return torch.where(m[..., None], torch.cat((q, x), -1), ident)

    def residues(n, m):
        aa = torch.randint(0, 20, (B, n), generator=g) * m
        return aa, torch.nn.functional.one_hot(aa, 22).float() * m[..., None]

    pep_frames, pocket_frames = frames(N_pad, mask, 5.0), frames(P_pad, pocket_mask, 10.0)
    aatype, feats = residues(N_pad, mask)
    pocket_aatype, pocket_feats = residues(P_pad, pocket_mask)
    ang = torch.rand(B, N_pad, N_TORSIONS, generator=g) * 2 * math.pi
    tmask = torch.zeros(B, N_pad, N_TORSIONS, dtype=torch.bool)
    tmask[:, :, 3:] = torch.arange(4)[None, None, :] < torch.tensor(_CHI_COUNT)[aatype][..., None]
    tmask &= mask[..., None]
    tmask[torch.arange(B), L - 1, 2] = True
    torsions = torch.where(tmask[..., None], torch.stack((ang.sin(), ang.cos()), -1), torch.tensor([0.0, 1.0]).expand(B, N_pad, N_TORSIONS, 2))
    return {"frames": pep_frames, "torsions": torsions, "features": feats, "mask": mask, "aatype": aatype, "torsions_mask": tmask,
            "pocket_frames": pocket_frames, "pocket_features": pocket_feats, "pocket_mask": pocket_mask, "pocket_aatype": pocket_aatype}
