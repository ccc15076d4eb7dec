"""TEST INFRASTRUCTURE ONLY — import shim for the UNMODIFIED reference.

Makes `/root/reference` (cmbi/pmhc-diffusion-model) importable in the build
container, where `openfold`, `h5py` and `Bio` are absent, without editing a
single reference file.  Used only by `tests/golden/make_golden.py` (fixture
generation) and by container-only tests that pin `oracle/egnn_oracle.py`
against the real reference.  `/root/reference` does not exist on the GPU box:
nothing in the product, `bench.py` or the `-m gpu` tests imports this module.

Pinned third-party dependency of the oracle: the reference
needs `openfold` 0.0.1 (README.md:9), un-vendored and un-pinned.  The only copy
of its `rigid_utils` in this image is the Apache-2.0 derivative shipped in
`transformers==5.5.0:models/esm/openfold_utils/rigid_utils.py`; the shim maps
`openfold.utils.rigid_utils` onto it.
"""
import importlib
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("PMHC_REFERENCE_ROOT", "/root/reference")


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "diffusion", "model.py"))


def _stub(name: str, **attrs) -> types.ModuleType:
    mod = types.ModuleType(name)
    mod.__dict__.update(attrs)
    sys.modules[name] = mod
    return mod


def install() -> None:
    """Register the module aliases/stubs and put the reference on sys.path."""
    if "openfold.utils.rigid_utils" in sys.modules:
        return
    if not reference_available():
        raise ImportError(f"reference tree not found at {REFERENCE_ROOT}")

    base = "transformers.models.esm.openfold_utils"
    rigid_utils = importlib.import_module(base + ".rigid_utils")
    feats = importlib.import_module(base + ".feats")
    residue_constants = importlib.import_module(base + ".residue_constants")

    def compute_fape(*args, **kwargs):  # imported by optimizer.py:8, never called
        raise NotImplementedError("compute_fape is not part of the hot path")

    openfold = _stub("openfold")
    utils = _stub("openfold.utils", rigid_utils=rigid_utils, feats=feats)
    loss = _stub("openfold.utils.loss", compute_fape=compute_fape)
    np_mod = _stub("openfold.np", residue_constants=residue_constants)
    sys.modules["openfold.utils.rigid_utils"] = rigid_utils
    sys.modules["openfold.utils.feats"] = feats
    sys.modules["openfold.np.residue_constants"] = residue_constants
    openfold.utils = utils
    openfold.np = np_mod
    utils.loss = loss

    # h5py / BioPython are only touched at import time on the hot path
    # (data.py:5, tools/pdb.py:4-9 via optimizer.py:12).
    if importlib.util.find_spec("h5py") is None:
        _stub("h5py", File=None)
    if importlib.util.find_spec("Bio") is None:
        _stub("Bio")
        _stub("Bio.PDB")
        for leaf, cls in (("Structure", "Structure"), ("Model", "Model"), ("Chain", "Chain"),
                          ("Residue", "Residue"), ("Atom", "Atom"), ("PDBIO", "PDBIO")):
            _stub("Bio.PDB." + leaf, **{cls: type(cls, (), {})})

    if REFERENCE_ROOT not in sys.path:
        sys.path.append(REFERENCE_ROOT)  # appended, not prepended: the reference has its own `tests` package


def load_reference():
    """Returns (model_module, optimizer_module, angle_module) of the real reference."""
    install()
    model = importlib.import_module("diffusion.model")
    optimizer = importlib.import_module("diffusion.optimizer")
    angle = importlib.import_module("diffusion.tools.angle")
    return model, optimizer, angle
