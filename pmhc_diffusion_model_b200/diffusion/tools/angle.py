"""Angle / quaternion helpers with the reference's names and semantics (diffusion/tools/angle.py:9-186).

Convenience API for callers and tests: the hot path does NOT go through these functions — the same math is
inlined in the CUDA kernels (csrc/pmhc_math.cuh).  They are plain elementwise torch expressions that run on
whatever device their inputs live on.
"""
from math import pi
from typing import List, Tuple, Union

import torch

from ...rigid import Rotation


def get_quat_conjugate(q: torch.Tensor) -> torch.Tensor:
    return q * q.new_tensor([1.0, -1.0, -1.0, -1.0])


def angle_to_sin_cos(angle: torch.Tensor) -> torch.Tensor:
    return torch.stack((torch.sin(angle), torch.cos(angle)), dim=-1)


def random_sin_cos(shape: Union[List[int], Tuple[int]], device: torch.device) -> torch.Tensor:
    return angle_to_sin_cos(torch.rand(list(shape), device=device) * 2 * pi)


def shoemake_quat(x: torch.Tensor) -> torch.Tensor:
    x = x.clamp(0.0, 1.0)
    th1, th2 = 2 * pi * x[..., 1], 2 * pi * x[..., 2]
    r1, r2 = torch.sqrt(1.0 - x[..., 0]), torch.sqrt(x[..., 0])
    return torch.stack((r2 * torch.cos(th2), r1 * torch.sin(th1), r1 * torch.cos(th1), r2 * torch.sin(th2)), dim=-1)


def random_quat(shape: Union[List[int], Tuple[int]], device: torch.device) -> torch.Tensor:
    return shoemake_quat(torch.rand(list(shape) + [3], device=device))


def multiply_sin_cos(sin_cos1: torch.Tensor, sin_cos2: torch.Tensor) -> torch.Tensor:
    s1, c1 = sin_cos1[..., 0], sin_cos1[..., 1]
    s2, c2 = sin_cos2[..., 0], sin_cos2[..., 1]
    return torch.stack((s1 * c2 + c1 * s2, c1 * c2 - s1 * s2), dim=-1)


def inverse_sin_cos(sin_cos: torch.Tensor) -> torch.Tensor:
    n2 = (sin_cos ** 2).sum(dim=-1, keepdim=True)
    return torch.stack((-sin_cos[..., 0], sin_cos[..., 1]), dim=-1) / n2


def partial_sin_cos(sin_cos: torch.Tensor, amount: float) -> torch.Tensor:
    u = torch.nn.functional.normalize(sin_cos, dim=-1)
    a = torch.acos(torch.clamp(u[..., 1], -1.0, 1.0))
    a = torch.where(u[..., 0] < 0.0, -a, a)
    return torch.stack((torch.sin(a * amount), torch.cos(a * amount)), dim=-1)


def partial_rot(rot: Rotation, amount: float) -> Rotation:
    q = torch.nn.functional.normalize(rot.get_quats(), dim=-1)
    half = torch.acos(torch.clamp(q[..., :1], -1.0, 1.0))
    axis = torch.nn.functional.normalize(q[..., 1:], dim=-1)
    return Rotation(quats=torch.cat((torch.cos(half * amount), torch.sin(half * amount) * axis), dim=-1), normalize_quats=False)
