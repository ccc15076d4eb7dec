"""Minimal Rigid / Rotation value types at the drop-in boundary.

The reference returns OpenFold `Rigid` objects from `Model.forward` (model.py:418-421) and feeds them to
`remove_noise`, `get_loss` and `tools/pdb.save`; OpenFold itself is not a dependency here.  These classes
are thin views over device tensors exposing exactly the methods the reference's callers use
(RU:253-601, 730-1045): from/to_tensor_7, from_tensor_4x4, get_trans/get_rots/get_quats/get_rot_mats,
compose_r, invert, apply, shape, device, __getitem__.  On the hot path rotations always stay quaternions
(w first), so `get_quats` never needs an eigen-decomposition (the reference's does).
"""
from __future__ import annotations

from typing import Any, Optional

import torch


def _quat_to_rot(q: torch.Tensor) -> torch.Tensor:
    """RU:145-165 (un-normalised formula: scales with |q|^2)."""
    a, b, c, d = q.unbind(-1)
    return torch.stack(
        (
            torch.stack((a * a + b * b - c * c - d * d, 2 * (b * c - a * d), 2 * (b * d + a * c)), -1),
            torch.stack((2 * (b * c + a * d), a * a - b * b + c * c - d * d, 2 * (c * d - a * b)), -1),
            torch.stack((2 * (b * d - a * c), 2 * (c * d + a * b), a * a - b * b - c * c + d * d), -1),
        ),
        dim=-2,
    )


def _rot_to_quat(r: torch.Tensor) -> torch.Tensor:
    """Rotation matrix -> unit quaternion by Shepperd's branch selection (largest of w,x,y,z first), then the
    sign is fixed so the largest-magnitude component is positive.  The reference takes the top eigenvector of
    a 4x4 matrix (RU:168-202), whose sign is arbitrary; both describe the same rotation."""
    m00, m01, m02 = r[..., 0, 0], r[..., 0, 1], r[..., 0, 2]
    m10, m11, m12 = r[..., 1, 0], r[..., 1, 1], r[..., 1, 2]
    m20, m21, m22 = r[..., 2, 0], r[..., 2, 1], r[..., 2, 2]
    cand = torch.stack(
        (
            torch.stack((1 + m00 + m11 + m22, m21 - m12, m02 - m20, m10 - m01), -1),
            torch.stack((m21 - m12, 1 + m00 - m11 - m22, m01 + m10, m02 + m20), -1),
            torch.stack((m02 - m20, m01 + m10, 1 - m00 + m11 - m22, m12 + m21), -1),
            torch.stack((m10 - m01, m02 + m20, m12 + m21, 1 - m00 - m11 + m22), -1),
        ),
        dim=-2,
    )  # row k is 4*q_k*q
    diag = torch.diagonal(cand, dim1=-2, dim2=-1)
    best = diag.argmax(dim=-1)
    q = torch.gather(cand, -2, best[..., None, None].expand(*best.shape, 1, 4)).squeeze(-2)
    return q / torch.linalg.norm(q, dim=-1, keepdim=True)


def _quat_mul(a: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
    aw, ax, ay, az = a.unbind(-1)
    bw, bx, by, bz = b.unbind(-1)
    return torch.stack(
        (aw * bw - ax * bx - ay * by - az * bz, aw * bx + ax * bw + ay * bz - az * by,
         aw * by - ax * bz + ay * bw + az * bx, aw * bz + ax * by - ay * bx + az * bw), -1)


class Rotation:
    """A batch of 3D rotations stored as quaternions (preferred) or rotation matrices (RU:253-727)."""

    def __init__(self, rot_mats: Optional[torch.Tensor] = None, quats: Optional[torch.Tensor] = None,
                 normalize_quats: bool = True):
        if (rot_mats is None) == (quats is None):
            raise ValueError("Exactly one input argument must be specified")
        if (rot_mats is not None and rot_mats.shape[-2:] != (3, 3)) or (quats is not None and quats.shape[-1] != 4):
            raise ValueError("Incorrectly shaped rotation matrix or quaternion")
        if quats is not None:
            quats = quats.to(dtype=torch.float32)
            if normalize_quats:
                quats = quats / torch.linalg.norm(quats, dim=-1, keepdim=True)
        if rot_mats is not None:
            rot_mats = rot_mats.to(dtype=torch.float32)
        self._rot_mats = rot_mats
        self._quats = quats

    @staticmethod
    def identity(shape, dtype=None, device=None, requires_grad: bool = False, fmt: str = "quat") -> "Rotation":
        q = torch.zeros((*shape, 4), dtype=torch.float32, device=device)
        q[..., 0] = 1.0
        return Rotation(quats=q, normalize_quats=False)

    def __getitem__(self, index: Any) -> "Rotation":
        if not isinstance(index, tuple):
            index = (index,)
        if self._quats is not None:
            return Rotation(quats=self._quats[index + (slice(None),)], normalize_quats=False)
        return Rotation(rot_mats=self._rot_mats[index + (slice(None), slice(None))])

    @property
    def shape(self) -> torch.Size:
        return self._quats.shape[:-1] if self._quats is not None else self._rot_mats.shape[:-2]

    @property
    def dtype(self) -> torch.dtype:
        return torch.float32

    @property
    def device(self) -> torch.device:
        return self._quats.device if self._quats is not None else self._rot_mats.device

    def get_quats(self) -> torch.Tensor:
        return self._quats if self._quats is not None else _rot_to_quat(self._rot_mats)

    def get_rot_mats(self) -> torch.Tensor:
        return self._rot_mats if self._rot_mats is not None else _quat_to_rot(self._quats)

    def get_cur_rot(self) -> torch.Tensor:
        return self._quats if self._quats is not None else self._rot_mats

    def compose_r(self, r: "Rotation") -> "Rotation":
        """self * r as rotations (RU:525-538).  Quaternion inputs stay quaternions (same rotation, no eigh)."""
        if self._quats is not None and r._quats is not None:
            return Rotation(quats=_quat_mul(self._quats, r._quats), normalize_quats=True)
        return Rotation(rot_mats=torch.matmul(self.get_rot_mats(), r.get_rot_mats()))

    def compose_q(self, r: "Rotation", normalize_quats: bool = True) -> "Rotation":
        return Rotation(quats=_quat_mul(self.get_quats(), r.get_quats()), normalize_quats=normalize_quats)

    def invert(self) -> "Rotation":
        if self._quats is not None:
            q = self._quats
            conj = q * q.new_tensor([1.0, -1.0, -1.0, -1.0])
            return Rotation(quats=conj / (q * q).sum(-1, keepdim=True), normalize_quats=False)
        return Rotation(rot_mats=self._rot_mats.transpose(-1, -2))

    def apply(self, pts: torch.Tensor) -> torch.Tensor:
        return torch.einsum("...ij,...j->...i", self.get_rot_mats(), pts)

    def invert_apply(self, pts: torch.Tensor) -> torch.Tensor:
        return torch.einsum("...ji,...j->...i", self.get_rot_mats(), pts)

    def map_tensor_fn(self, fn) -> "Rotation":
        if self._quats is not None:
            return Rotation(quats=torch.stack([fn(x) for x in self._quats.unbind(-1)], -1), normalize_quats=False)
        flat = self._rot_mats.reshape(self._rot_mats.shape[:-2] + (9,))
        flat = torch.stack([fn(x) for x in flat.unbind(-1)], -1)
        return Rotation(rot_mats=flat.reshape(flat.shape[:-1] + (3, 3)))

    def to(self, device=None, dtype=None) -> "Rotation":
        if self._quats is not None:
            return Rotation(quats=self._quats.to(device=device), normalize_quats=False)
        return Rotation(rot_mats=self._rot_mats.to(device=device))

    def cuda(self) -> "Rotation":
        return self.to(device="cuda")

    def detach(self) -> "Rotation":
        if self._quats is not None:
            return Rotation(quats=self._quats.detach(), normalize_quats=False)
        return Rotation(rot_mats=self._rot_mats.detach())


class Rigid:
    """Rotation + translation (RU:730-1243); `to_tensor_7` is (quat w x y z, translation x y z)."""

    def __init__(self, rots: Optional[Rotation], trans: Optional[torch.Tensor]):
        if rots is None and trans is None:
            raise ValueError("At least one input argument must be specified")
        if rots is None:
            rots = Rotation.identity(trans.shape[:-1], device=trans.device)
        if trans is None:
            trans = torch.zeros((*rots.shape, 3), dtype=torch.float32, device=rots.device)
        if rots.shape != trans.shape[:-1] or rots.device != trans.device:
            raise ValueError("Rots and trans incompatible")
        self._rots = rots
        self._trans = trans.to(dtype=torch.float32)

    @staticmethod
    def identity(shape, dtype=None, device=None, requires_grad: bool = False, fmt: str = "quat") -> "Rigid":
        return Rigid(Rotation.identity(shape, device=device), torch.zeros((*shape, 3), dtype=torch.float32, device=device))

    def __getitem__(self, index: Any) -> "Rigid":
        if not isinstance(index, tuple):
            index = (index,)
        return Rigid(self._rots[index], self._trans[index + (slice(None),)])

    @property
    def shape(self) -> torch.Size:
        return self._trans.shape[:-1]

    @property
    def device(self) -> torch.device:
        return self._trans.device

    def get_rots(self) -> Rotation:
        return self._rots

    def get_trans(self) -> torch.Tensor:
        return self._trans

    def compose(self, r: "Rigid") -> "Rigid":
        return Rigid(self._rots.compose_r(r._rots), self._rots.apply(r._trans) + self._trans)

    def apply(self, pts: torch.Tensor) -> torch.Tensor:
        return self._rots.apply(pts) + self._trans

    def invert_apply(self, pts: torch.Tensor) -> torch.Tensor:
        return self._rots.invert_apply(pts - self._trans)

    def invert(self) -> "Rigid":
        inv = self._rots.invert()
        return Rigid(inv, -1 * inv.apply(self._trans))

    def map_tensor_fn(self, fn) -> "Rigid":
        return Rigid(self._rots.map_tensor_fn(fn), torch.stack([fn(x) for x in self._trans.unbind(-1)], -1))

    def to_tensor_4x4(self) -> torch.Tensor:
        t = self._trans.new_zeros((*self.shape, 4, 4))
        t[..., :3, :3] = self._rots.get_rot_mats()
        t[..., :3, 3] = self._trans
        t[..., 3, 3] = 1
        return t

    @staticmethod
    def from_tensor_4x4(t: torch.Tensor) -> "Rigid":
        if t.shape[-2:] != (4, 4):
            raise ValueError("Incorrectly shaped input tensor")
        return Rigid(Rotation(rot_mats=t[..., :3, :3]), t[..., :3, 3])

    def to_tensor_7(self) -> torch.Tensor:
        return torch.cat((self._rots.get_quats(), self._trans), dim=-1)

    @staticmethod
    def from_tensor_7(t: torch.Tensor, normalize_quats: bool = False) -> "Rigid":
        if t.shape[-1] != 7:
            raise ValueError("Incorrectly shaped input tensor")
        return Rigid(Rotation(quats=t[..., :4], normalize_quats=normalize_quats), t[..., 4:])

    def unsqueeze(self, dim: int) -> "Rigid":
        return Rigid.from_tensor_7(self.to_tensor_7().unsqueeze(dim if dim >= 0 else dim - 1))

    def to(self, device=None, dtype=None) -> "Rigid":
        return Rigid(self._rots.to(device=device), self._trans.to(device=device))

    def cuda(self) -> "Rigid":
        return self.to(device="cuda")
