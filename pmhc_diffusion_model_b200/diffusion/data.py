"""SwiftMHC-format HDF5 input — drop-in for the reference's `diffusion.data.MhcpDataset` (diffusion/data.py:13-145), plus
a batched GPU-resident path for the fused kernels.

Per-entry surface (same class attributes, same 13 keys / shapes / dtypes as data.py:105-117, same ValueError on a missing
peptide, data.py:41-42): `MhcpDataset(hdf5_path, device)[i]`, `get_entry(name)`, `get_protein_positions(names)`.
The file is read with h5py when it is installed, else with `hdf5_lite` (this image has no h5py / libhdf5).

Batched path: the reference re-opens the file and runs a CPU `eigh` per entry for the 4x4 ->
quaternion conversion (data.py:38, :107, :115), which caps it at a few thousand complexes/s while the fused kernels
consume > 10^5.  `load_all()` parses the file once into padded pinned host arrays (frames still 4x4), `batches()` /
`device_batch()` copy slices to the GPU and convert the frames there (`pmhc_frames4x4_to_tensor7`, one thread per residue).
"""
from typing import Dict, Iterator, List, Optional, Sequence, Union

import numpy
import torch
from torch.utils.data import Dataset

from .. import _lib
from ..rigid import Rigid
from . import hdf5_lite


def frames4x4_to_tensor7(frames4x4: torch.Tensor) -> torch.Tensor:
    """[*, 4, 4] homogeneous matrices on the GPU -> [*, 7] (unit quaternion, w >= 0, then translation)."""
    t = _lib.f32c(frames4x4)
    dev = _lib.require_cuda(t)
    out = torch.empty(t.shape[:-2] + (7,), device=dev, dtype=torch.float32)
    with torch.cuda.device(dev):
        _lib.check(_lib.load().pmhc_frames4x4_to_tensor7(t.data_ptr(), t.numel() // 16, out.data_ptr(), _lib.stream_ptr(dev)),
                   "pmhc_frames4x4_to_tensor7")
    return out


class MhcpDataset(Dataset):

    peptide_maxlen = 16     # data.py:15
    pocket_maxlen = 80      # data.py:16

    def __init__(self, hdf5_path: str, device: Optional[torch.device] = None):
        self.device = torch.device(device) if device is not None else torch.device("cpu")
        self.hdf5_path = hdf5_path
        with hdf5_lite.open_file(hdf5_path) as f5:
            self.entry_names = list(f5.keys())
        self._host: Optional[Dict[str, torch.Tensor]] = None

    def __getitem__(self, index: int) -> Dict[str, torch.Tensor]:
        return self.get_entry(self.entry_names[index])

    def __len__(self) -> int:
        return len(self.entry_names)

    # ---- one entry, host arrays ---------------------------------------------------------------------------------------
    @classmethod
    def _entry_arrays(cls, f5, entry_name: str) -> Dict[str, numpy.ndarray]:
        """Padded numpy arrays of one complex (frames still 4x4): the masking / padding rules of data.py:44-102."""
        entry = f5[entry_name]
        if "peptide" not in entry:
            raise ValueError(f"no peptide in {entry_name}")
        peptide, mhc = entry["peptide"], entry["protein"]
        N, P = cls.peptide_maxlen, cls.pocket_maxlen

        sel = numpy.asarray(mhc["cross_residues_mask"][:]).astype(bool)
        n_pocket = int(sel.sum())
        pep_frames = numpy.asarray(peptide["backbone_rigid_tensor"][:], dtype=numpy.float32)
        L = pep_frames.shape[0]
        if L > N or n_pocket > P:
            raise ValueError(f"{entry_name}: peptide length {L} / pocket size {n_pocket} exceed ({N}, {P})")

        out: Dict[str, numpy.ndarray] = {}
        eye = numpy.eye(4, dtype=numpy.float32)
        frames = numpy.tile(eye, (N, 1, 1))                 # padded slots: identity frames (data.py:71-72)
        frames[:L] = pep_frames
        pocket_frames = numpy.tile(eye, (P, 1, 1))
        pocket_frames[:n_pocket] = numpy.asarray(mhc["backbone_rigid_tensor"][:], dtype=numpy.float32)[sel]
        out["frames4x4"], out["pocket_frames4x4"] = frames, pocket_frames

        out["mask"] = numpy.arange(N) < L
        out["pocket_mask"] = numpy.arange(P) < n_pocket
        aatype = numpy.zeros(N, dtype=numpy.int64)
        aatype[:L] = numpy.asarray(peptide["aatype"][:]).astype(numpy.int64)
        out["aatype"] = aatype
        pocket_aatype = numpy.zeros(P, dtype=numpy.int64)
        pocket_aatype[:n_pocket] = numpy.asarray(mhc["aatype"][:]).astype(numpy.int64)[sel]
        out["pocket_aatype"] = pocket_aatype

        feats = numpy.zeros((N, 22), dtype=numpy.float32)
        feats[:L] = numpy.asarray(peptide["sequence_onehot"][:], dtype=numpy.float32)
        out["features"] = feats
        pfeats = numpy.zeros((P, 22), dtype=numpy.float32)
        pfeats[:n_pocket] = numpy.asarray(mhc["sequence_onehot"][:], dtype=numpy.float32)[sel]
        out["pocket_features"] = pfeats

        xyz = numpy.zeros((P, 14, 3), dtype=numpy.float32)
        xyz[:n_pocket] = numpy.asarray(mhc["atom14_gt_positions"][:], dtype=numpy.float32)[sel]
        out["pocket_atom14_positions"] = xyz
        exists = numpy.zeros((P, 14), dtype=bool)
        exists[:n_pocket] = numpy.asarray(mhc["atom14_gt_exists"][:]).astype(bool)[sel]
        out["pocket_atom14_exists"] = exists

        # torsions: pre-omega, phi, psi, chi1..chi4.  The frames fix the backbone, so backbone torsions are disabled
        # except psi of the C-terminus (which places its oxygens); masked torsions are the identity (data.py:91-102)
        tmask = numpy.zeros((N, 7), dtype=bool)
        tmask[:L] = numpy.asarray(peptide["torsion_angles_mask"][:]).astype(bool)
        tmask[:, :3] = False
        tmask[L - 1, 2] = True
        tors = numpy.zeros((N, 7, 2), dtype=numpy.float32)
        tors[:L] = numpy.asarray(peptide["torsion_angles_sin_cos"][:], dtype=numpy.float32)
        tors[~tmask] = (0.0, 1.0)
        out["torsions"], out["torsions_mask"] = tors, tmask
        return out

    ORDER = ("mask", "frames", "features", "aatype", "torsions", "torsions_mask", "pocket_aatype", "pocket_features",
             "pocket_mask", "pocket_frames", "pocket_atom14_positions", "pocket_atom14_exists")

    def get_entry(self, entry_name: str) -> Dict[str, Union[List[str], torch.Tensor]]:
        with hdf5_lite.open_file(self.hdf5_path) as f5:
            arrays = self._entry_arrays(f5, entry_name)
        data: Dict[str, Union[List[str], torch.Tensor]] = {"name": [entry_name]}
        t = {k: torch.from_numpy(v).to(self.device) for k, v in arrays.items()}
        for key in ("frames", "pocket_frames"):
            m = t.pop(key + "4x4")
            # tensor_7 rows for collation (data.py:107, :115)
            t[key] = frames4x4_to_tensor7(m) if m.is_cuda else Rigid.from_tensor_4x4(m).to_tensor_7()
        for key in self.ORDER:
            data[key] = t[key]
        return data

    def get_protein_positions(self, entry_names: Sequence[str]) -> Dict[str, torch.Tensor]:
        """data.py:121-145: full-protein aatype / atom14 positions / existence, stacked over the entries."""
        data: Dict[str, list] = {"protein_aatype": [], "protein_atom14_positions": [], "protein_atom14_exists": []}
        with hdf5_lite.open_file(self.hdf5_path) as f5:
            for entry_name in entry_names:
                mhc = f5[entry_name]["protein"]
                data["protein_aatype"].append(torch.tensor(numpy.asarray(mhc["aatype"][:]), device=self.device))
                data["protein_atom14_positions"].append(torch.tensor(numpy.asarray(mhc["atom14_gt_positions"][:]), device=self.device))
                data["protein_atom14_exists"].append(torch.tensor(numpy.asarray(mhc["atom14_gt_exists"][:]), device=self.device))
        return {k: torch.stack(v) for k, v in data.items()}

    # ---- whole file, batched ------------------------------------------------------------------------------------------
    def load_all(self, pin: bool = True) -> Dict[str, torch.Tensor]:
        """Parses every entry once into stacked host tensors [n_entries, ...] (frames as 4x4), pinned for async copies."""
        if self._host is not None:
            return self._host
        cols: Dict[str, list] = {}
        with hdf5_lite.open_file(self.hdf5_path) as f5:
            for name in self.entry_names:
                for k, v in self._entry_arrays(f5, name).items():
                    cols.setdefault(k, []).append(v)
        host = {k: torch.from_numpy(numpy.stack(v)) for k, v in cols.items()}
        if pin and torch.cuda.is_available():
            host = {k: v.pin_memory() for k, v in host.items()}
        self._host = host
        return host

    def device_batch(self, indices: Union[slice, Sequence[int], torch.Tensor], device: Optional[torch.device] = None) -> Dict:
        """One collated batch on the GPU with the reference's keys; frames converted 4x4 -> tensor_7 on the device."""
        device = torch.device(device) if device is not None else self.device
        if device.type != "cuda":
            raise RuntimeError("device_batch builds GPU-resident batches; use the per-entry interface for CPU tensors")
        host = self.load_all()
        if isinstance(indices, slice):
            rows = {k: v[indices] for k, v in host.items()}
            names = self.entry_names[indices]
        else:
            idx = torch.as_tensor(indices, dtype=torch.long)
            rows = {k: v[idx] for k, v in host.items()}
            names = [self.entry_names[int(i)] for i in idx]
        dev = {k: v.to(device, non_blocking=True) for k, v in rows.items()}
        batch: Dict = {"name": [list(names)]}        # the DataLoader collates ['name'] per entry into [[...]] (test.py:63)
        frames = {key: frames4x4_to_tensor7(dev.pop(key + "4x4")) for key in ("frames", "pocket_frames")}
        dev.update(frames)
        for key in self.ORDER:
            batch[key] = dev[key]
        return batch

    def batches(self, batch_size: int, device: Optional[torch.device] = None, shuffle: bool = False,
                generator: Optional[torch.Generator] = None) -> Iterator[Dict]:
        """GPU-resident batches over the whole file: the DataLoader(dataset, batch_size, shuffle) of optimize.py:62-63 /
        test.py:57 without worker processes, per-entry file opens or per-entry eigh."""
        n = len(self)
        order = torch.randperm(n, generator=generator) if shuffle else torch.arange(n)
        for s in range(0, n, batch_size):
            yield self.device_batch(order[s:s + batch_size], device)


def write_synthetic_hdf5(path: str, n_complexes: int, peptide_len=9, protein_len: int = 180, pocket_n: int = 60, seed: int = 0) -> List[str]:
    """A synthetic file in the SwiftMHC layout (README.md:15-37) for tests and benchmarks: random proper rotations,
    one-hot sequences, chi masks from the amino-acid type.  Returns the entry names."""
    rng = numpy.random.default_rng(seed)
    chi_count = numpy.array([0, 4, 2, 2, 1, 3, 3, 0, 2, 2, 2, 4, 3, 2, 2, 1, 1, 2, 2, 1])   # restypes order ARNDCQEGHILKMFPSTWYV
    heavy_atoms = numpy.array([5, 11, 8, 8, 6, 9, 9, 4, 10, 8, 8, 9, 8, 11, 7, 6, 7, 14, 12, 7])   # filled atom14 slots

    def rigid(n, spread):
        q = rng.standard_normal((n, 4))
        q /= numpy.linalg.norm(q, axis=-1, keepdims=True)
        w, x, y, z = q.T
        R = numpy.stack([1 - 2 * (y * y + z * z), 2 * (x * y - w * z), 2 * (x * z + w * y),
                         2 * (x * y + w * z), 1 - 2 * (x * x + z * z), 2 * (y * z - w * x),
                         2 * (x * z - w * y), 2 * (y * z + w * x), 1 - 2 * (x * x + y * y)], -1).reshape(n, 3, 3)
        m = numpy.tile(numpy.eye(4), (n, 1, 1))
        m[:, :3, :3] = R
        m[:, :3, 3] = rng.standard_normal((n, 3)) * spread
        return m.astype(numpy.float32)

    def residues(n):
        aatype = rng.integers(0, 20, n)
        onehot = numpy.zeros((n, 22), dtype=numpy.float32)
        onehot[numpy.arange(n), aatype] = 1.0
        return aatype.astype(numpy.int64), onehot

    tree, names = {}, []
    for c in range(n_complexes):
        L = int(peptide_len) if numpy.isscalar(peptide_len) else int(rng.integers(peptide_len[0], peptide_len[1] + 1))
        aatype, onehot = residues(L)
        ang = rng.uniform(0, 2 * numpy.pi, (L, 7))
        tmask = numpy.zeros((L, 7), dtype=bool)
        tmask[:, :3] = True
        tmask[:, 3:] = numpy.arange(4)[None, :] < chi_count[aatype][:, None]
        paatype, ponehot = residues(protein_len)
        cross = numpy.zeros(protein_len, dtype=bool)
        cross[rng.choice(protein_len, size=min(pocket_n, protein_len), replace=False)] = True
        name = f"BA-{c:06d}"
        names.append(name)
        tree[name] = {
            "peptide": {"backbone_rigid_tensor": rigid(L, 5.0), "aatype": aatype, "sequence_onehot": onehot,
                        "torsion_angles_sin_cos": numpy.stack((numpy.sin(ang), numpy.cos(ang)), -1).astype(numpy.float32),
                        "torsion_angles_mask": tmask},
            "protein": {"backbone_rigid_tensor": rigid(protein_len, 10.0), "aatype": paatype, "sequence_onehot": ponehot,
                        "atom14_gt_positions": (rng.standard_normal((protein_len, 14, 3)) * 10.0).astype(numpy.float32),
                        "atom14_gt_exists": numpy.arange(14)[None, :] < heavy_atoms[paatype][:, None], "cross_residues_mask": cross},
        }
    hdf5_lite.write_file(path, tree)
    return names
