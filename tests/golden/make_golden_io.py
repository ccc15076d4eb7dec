#!/usr/bin/env python
"""Golden fixtures for the loader / writer rows from the UNMODIFIED reference.

Run in the build container only (needs /root/reference):

    python tests/golden/make_golden_io.py

* `MhcpDataset.get_entry` / `get_protein_positions` (diffusion/data.py:35-145) run on synthetic SwiftMHC entries served
  by an in-memory stand-in for `h5py.File` (h5py is absent here; the stand-in only hands the reference the arrays).
* `diffusion.tools.pdb.save` (tools/pdb.py:34-211) runs with recording stand-ins for the BioPython containers
  (BioPython is absent): the fixture keeps every atom the reference adds — chain, residue number and name, atom name,
  coordinates — in the order it adds them.  The text PDBIO would write is therefore NOT pinned by this fixture.
Nothing of the reference is edited; its modules' imported names are pointed at the stand-ins at run time.
"""
import importlib
import os
import sys

import numpy
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from oracle import ref_shim  # noqa: E402
from pmhc_diffusion_model_b200.diffusion.data import write_synthetic_hdf5  # noqa: E402  (synthetic inputs only)
from pmhc_diffusion_model_b200.diffusion import hdf5_lite  # noqa: E402

torch.set_num_threads(1)
ref_shim.load_reference()
ref_data = importlib.import_module("diffusion.data")
ref_pdb = importlib.import_module("diffusion.tools.pdb")
from openfold.utils.rigid_utils import Rigid  # noqa: E402  (the shim's alias)


# ---- h5py stand-in: nested dicts of numpy arrays ---------------------------------------------------------------------
class _Node:
    def __init__(self, tree):
        self._t = tree

    def keys(self):
        return self._t.keys()

    def __contains__(self, k):
        return k in self._t

    def __getitem__(self, k):
        v = self._t[k]
        return _Node(v) if isinstance(v, dict) else v      # numpy arrays answer [:] themselves


TREES = {}


class FakeFile(_Node):
    def __init__(self, path, mode="r"):
        super().__init__(TREES[path])

    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False


# ---- BioPython stand-ins: record what is added -------------------------------------------------------------------------
class _Container:
    def __init__(self, ident=None, *rest):
        self.id, self.rest, self.children = ident, rest, []

    def add(self, child):
        self.children.append(child)


class FakeAtom:
    def __init__(self, name, coord, bfactor, occupancy, altloc, fullname, serial_number, element=None):
        self.name, self.coord, self.fullname, self.serial, self.element = name, torch.as_tensor(coord).clone(), fullname, serial_number, element
        self.bfactor, self.occupancy = bfactor, occupancy


class FakeIO:
    saved = {}

    def set_structure(self, s):
        self.s = s

    def save(self, path):
        atoms = []
        for model in self.s.children:
            for chain in model.children:
                for res in chain.children:
                    for atom in res.children:
                        atoms.append((chain.id, res.id[1], res.rest[0], atom.name, atom.coord.tolist(), atom.occupancy, atom.bfactor))
        FakeIO.saved[path] = atoms


def tree_of(path):
    """The arrays of a file written by hdf5_lite, as nested dicts (what the stand-in serves)."""
    def walk(g):
        return {k: (walk(g[k]) if isinstance(g[k], hdf5_lite.Group) else g[k][:]) for k in g.keys()}
    with hdf5_lite.File(path) as f:
        return walk(f)


def main():
    ref_data.h5py.File = FakeFile
    for name, cls in (("Structure", _Container), ("PDBModel", _Container), ("Chain", _Container), ("Residue", _Container),
                      ("Atom", FakeAtom), ("PDBIO", FakeIO)):
        setattr(ref_pdb, name, cls)

    path = os.path.join(HERE, "_tmp_io.h5")
    names = write_synthetic_hdf5(path, 4, peptide_len=(8, 13), protein_len=40, pocket_n=24, seed=11)
    raw = tree_of(path)
    os.remove(path)
    TREES["mem"] = raw

    ds = ref_data.MhcpDataset("mem")
    assert list(ds.entry_names) == names
    entries = [ds.get_entry(n) for n in names]
    protein = ds.get_protein_positions(names)

    # sampled-looking frames: random unit quaternions and torsions on the real rows
    g = torch.Generator().manual_seed(5)
    batch = {k: torch.stack([e[k] for e in entries]) for k in ("mask", "aatype")}
    B = len(names)
    q = torch.nn.functional.normalize(torch.randn(B, 16, 4, generator=g), dim=-1)
    x = torch.randn(B, 16, 3, generator=g) * 6.0
    ang = torch.rand(B, 16, 7, generator=g) * 2 * numpy.pi
    frames7 = torch.cat((q, x), -1)
    torsions = torch.stack((ang.sin(), ang.cos()), -1)
    batch.update({"frames": Rigid.from_tensor_7(frames7), "torsions": torsions})
    batch.update(protein)
    pdb_atoms = []
    for i, n in enumerate(names):
        ref_pdb.save(batch, i, n)
        pdb_atoms.append(FakeIO.saved[n])

    torch.save({"raw": raw, "names": names, "entries": entries, "protein": protein,
                "pdb": {"frames7": frames7, "torsions": torsions, "atoms": pdb_atoms}},
               os.path.join(HERE, "io_golden.pt"))
    print("wrote io_golden.pt:", len(names), "entries;", [len(a) for a in pdb_atoms], "atoms")


if __name__ == "__main__":
    main()
