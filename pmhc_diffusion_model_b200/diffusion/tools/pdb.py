"""PDB output of sampled complexes — drop-in for the reference's `diffusion.tools.pdb.save` (tools/pdb.py:34-211).

The reference builds BioPython objects atom by atom on the host and runs OpenFold's torsion / literature-position
functions per call.  Here the peptide's heavy atoms of a whole batch come from ONE kernel launch (`pmhc_atom14`:
torsion frames, atom14 placement, backbone O from the next residue's N, terminal O / OXT — pdb.py:67-174) and the text is
formatted directly in the fixed PDB columns BioPython's PDBIO writes (ATOM records, serial numbers renumbered from 1, a
TER record per chain, END) — BioPython itself is not needed.

Chain P = the peptide built from frames and torsions, chain M = the protein from its stored atom14 coordinates
(pdb.py:177-204).  Amino-acid geometry tables come from OpenFold's residue_constants, here the copy shipped in
`transformers.models.esm.openfold_utils` (the reference imports the same tables from `openfold.np`, pdb.py:14-23).
"""
import logging
import os
from typing import Dict, List, Optional, Sequence, Tuple, Union

import numpy
import torch

from ... import _lib
from ...rigid import Rigid

_log = logging.getLogger(__name__)

_tables_host: Optional[dict] = None
_tables_dev: Dict[Tuple[str, Optional[int]], tuple] = {}


def _residue_constants():
    try:
        from openfold.np import residue_constants as rc   # the reference's import (pdb.py:14)
    except ImportError:
        try:
            from transformers.models.esm.openfold_utils import residue_constants as rc
        except ImportError as e:   # pragma: no cover
            raise ImportError("PDB output needs OpenFold's residue_constants (openfold, or transformers' "
                              "models.esm.openfold_utils copy)") from e
    return rc


def _host_tables() -> dict:
    global _tables_host
    if _tables_host is None:
        rc = _residue_constants()
        names3 = [rc.restype_1to3[r] for r in rc.restypes] + ["UNK"]
        atom_names = [list(rc.restype_name_to_atom14_names[n]) + ["OXT"] for n in names3]
        # PDB atom name column as PDBIO pads it: one-letter elements start in the second column
        padded = [[("" if not a else ((" " + a) if len(a) < 4 else a)).ljust(4) for a in row] for row in atom_names]
        _tables_host = {
            "default_frames": numpy.asarray(rc.restype_rigid_group_default_frame, dtype=numpy.float32),
            "group_idx": numpy.asarray(rc.restype_atom14_to_rigid_group, dtype=numpy.int32),
            "lit_positions": numpy.asarray(rc.restype_atom14_rigid_group_positions, dtype=numpy.float32),
            "atom_mask": numpy.asarray(rc.restype_atom14_mask, dtype=numpy.uint8),
            "names3": names3, "atom_names": atom_names, "atom_field": padded,
        }
    return _tables_host


def _device_tables(dev: torch.device) -> tuple:
    key = (dev.type, dev.index)
    if key not in _tables_dev:
        t = _host_tables()
        _tables_dev[key] = tuple(torch.from_numpy(t[k]).to(dev).contiguous() for k in ("default_frames", "group_idx", "lit_positions", "atom_mask"))
    return _tables_dev[key]


def peptide_atoms(batch: Dict[str, Union[Rigid, torch.Tensor]]) -> Tuple[torch.Tensor, torch.Tensor]:
    """Heavy-atom coordinates of every peptide in the batch: positions [B,16,15,3] (atom14 order, slot 14 = OXT) and
    exists [B,16,15] (bool).  One launch for the whole batch."""
    frames = batch["frames"]
    f7 = _lib.f32c(frames.to_tensor_7() if isinstance(frames, Rigid) else frames)
    tors = _lib.f32c(batch["torsions"])
    aatype = batch["aatype"].to(torch.int64).contiguous()
    mask = _lib.u8c(batch["mask"])
    dev = _lib.require_cuda(f7, tors, aatype, mask)
    B = f7.shape[0]
    if tuple(f7.shape[1:]) != (_lib.N, 7) or tuple(tors.shape[1:]) != (_lib.N, _lib.NTORS, 2):
        raise ValueError("frames must be [B,16,7] and torsions [B,16,7,2]")
    df, gi, lit, am = _device_tables(dev)
    pos = torch.empty(B, _lib.N, 15, 3, device=dev, dtype=torch.float32)
    exists = torch.empty(B, _lib.N, 15, device=dev, dtype=torch.uint8)
    with torch.cuda.device(dev):
        _lib.check(_lib.load().pmhc_atom14(f7.data_ptr(), tors.data_ptr(), aatype.data_ptr(), mask.data_ptr(), B, df.data_ptr(),
                                           gi.data_ptr(), lit.data_ptr(), am.data_ptr(), pos.data_ptr(), exists.data_ptr(),
                                           _lib.stream_ptr(dev)), "pmhc_atom14")
    return pos, exists.bool()


# atom order inside a peptide residue as the reference adds them (pdb.py:112-174): N, CA, C, CB, side chain, O [, OXT]
_PEPTIDE_ORDER = [0, 1, 2, 4] + list(range(5, 14)) + [3, 14]


def _atom_line(serial: int, name_field: str, res3: str, chain: str, resseq: int, x: float, y: float, z: float, element: str) -> str:
    return f"ATOM  {serial:5d} {name_field} {res3:>3s} {chain}{resseq:4d}    {x:8.3f}{y:8.3f}{z:8.3f}{1.0:6.2f}{0.0:6.2f}          {element:>2s}  \n"


def _ter_line(serial: int, res3: str, chain: str, resseq: int) -> str:
    return f"TER   {serial:5d}      {res3:>3s} {chain}{resseq:4d}".ljust(80) + "\n"


def _text_tables() -> Tuple[bytes, bytes, bytes]:
    t = _host_tables()
    if "c_fields" not in t:
        t["c_fields"] = "".join(f for row in t["atom_field"] for f in row).encode()
        t["c_elements"] = "".join((a[0] if a else " ") for row in t["atom_names"] for a in row).encode()
        t["c_res3"] = "".join(t["names3"]).encode()
    return t["c_fields"], t["c_elements"], t["c_res3"]


def format_pdb(pep_aatype: numpy.ndarray, pep_mask: numpy.ndarray, pep_pos: numpy.ndarray, pep_exists: numpy.ndarray,
               prot_aatype: numpy.ndarray, prot_pos: numpy.ndarray, prot_exists: numpy.ndarray) -> str:
    """The file text of one complex from host arrays (peptide [16,...], protein [M,...]), formatted by the library's host
    routine `pmhc_format_pdb_host` (per-atom Python string formatting caps at ~200 files/s)."""
    import ctypes
    fields, elements, res3 = _text_tables()
    c = lambda a, dt: numpy.ascontiguousarray(a, dtype=dt)
    pa, pm, pp, pe = c(pep_aatype, numpy.int64), c(pep_mask, numpy.uint8), c(pep_pos, numpy.float32), c(pep_exists, numpy.uint8)
    ra, rp, re_ = c(prot_aatype, numpy.int64), c(prot_pos, numpy.float32), c(prot_exists, numpy.uint8)
    cap = (16 * 15 + ra.shape[0] * 14 + 3) * 96 + 1
    out = ctypes.create_string_buffer(cap)
    n = _lib.load().pmhc_format_pdb_host(pa.ctypes.data, pm.ctypes.data, pp.ctypes.data, pe.ctypes.data, ra.shape[0], ra.ctypes.data,
                                         rp.ctypes.data, re_.ctypes.data, fields, elements, res3, ctypes.addressof(out), cap)
    if n < 0:
        raise RuntimeError("pmhc_format_pdb_host: output buffer too small")
    return out.raw[:n].decode()


def format_pdb_python(pep_aatype: numpy.ndarray, pep_mask: numpy.ndarray, pep_pos: numpy.ndarray, pep_exists: numpy.ndarray,
                      prot_aatype: numpy.ndarray, prot_pos: numpy.ndarray, prot_exists: numpy.ndarray) -> str:
    """The same text built line by line in Python: the readable statement of the layout (tests compare the two)."""
    t = _host_tables()
    lines: List[str] = []
    serial = 0
    last = None
    for i in numpy.nonzero(pep_mask)[0]:
        aa = min(int(pep_aatype[i]), 20)
        res3 = t["names3"][aa]
        for a in _PEPTIDE_ORDER:
            if pep_exists[i, a]:
                serial += 1
                x, y, z = pep_pos[i, a]
                lines.append(_atom_line(serial, t["atom_field"][aa][a], res3, "P", int(i) + 1, x, y, z, t["atom_names"][aa][a][0]))
        last = (res3, int(i) + 1)
    if last is not None:
        serial += 1
        lines.append(_ter_line(serial, last[0], "P", last[1]))
    last = None
    for i in range(prot_aatype.shape[0]):
        aa = min(int(prot_aatype[i]), 20)
        res3 = t["names3"][aa]
        for a in numpy.nonzero(prot_exists[i])[0]:
            serial += 1
            x, y, z = prot_pos[i, a]
            lines.append(_atom_line(serial, t["atom_field"][aa][a], res3, "M", i + 1, x, y, z, t["atom_names"][aa][a][0]))
        last = (res3, i + 1)
    if last is not None:
        serial += 1
        lines.append(_ter_line(serial, last[0], "M", last[1]))
    lines.append("END   \n")
    return "".join(lines)


def save_batch(batch: Dict[str, Union[Rigid, torch.Tensor]], names: Sequence[str], directory: str) -> List[str]:
    """Writes `<directory>/<name>.pdb` for every complex of the batch (the loop of test.py:82-84): one kernel launch and
    one device-to-host copy for all peptides, then text formatting."""
    pos, exists = peptide_atoms(batch)
    pos, exists = pos.cpu().numpy(), exists.cpu().numpy()
    aatype, mask = batch["aatype"].cpu().numpy(), batch["mask"].cpu().numpy().astype(bool)
    p_aa = batch["protein_aatype"].cpu().numpy()
    p_pos = batch["protein_atom14_positions"].cpu().numpy()
    p_ex = batch["protein_atom14_exists"].cpu().numpy().astype(bool)
    os.makedirs(directory, exist_ok=True)
    paths = []
    for b, name in enumerate(names):
        path = os.path.join(directory, f"{name}.pdb")
        with open(path, "w") as fh:
            fh.write(format_pdb(aatype[b], mask[b], pos[b], exists[b], p_aa[b], p_pos[b], p_ex[b]))
        paths.append(path)
    _log.debug("saved %d structures under %s", len(paths), directory)
    return paths


def save(batch: Dict[str, Union[Rigid, torch.Tensor]], batch_index: int, path: str) -> None:
    """tools/pdb.py:34-211: writes complex `batch_index` of the batch to `path`."""
    one = {}
    for k in ("frames", "torsions", "aatype", "mask", "protein_aatype", "protein_atom14_positions", "protein_atom14_exists"):
        v = batch[k]
        one[k] = v[batch_index:batch_index + 1]
    pos, exists = peptide_atoms(one)
    text = format_pdb(one["aatype"][0].cpu().numpy(), one["mask"][0].cpu().numpy().astype(bool), pos[0].cpu().numpy(),
                      exists[0].cpu().numpy(), one["protein_aatype"][0].cpu().numpy(),
                      one["protein_atom14_positions"][0].cpu().numpy(), one["protein_atom14_exists"][0].cpu().numpy().astype(bool))
    with open(path, "w") as fh:
        fh.write(text)
    _log.debug(f"saved {path}")
