#!/usr/bin/env python
"""Install the UNMODIFIED reference into baseline/_ref (git-ignored; travels to the GPU box with the snapshot).

The reference (cmbi/pmhc-diffusion-model) ships no setup.py / pyproject.toml, so the tree is copied to a scratch directory
under /tmp (/root/reference is read-only), given a three-line setup.py that names its packages, and installed with

    python -m pip install --no-index --no-build-isolation --no-deps --find-links /opt/wheelhouse --target baseline/_ref <copy>

Not one reference file is edited; `diff -r /root/reference/diffusion baseline/_ref/diffusion` is empty (checked below).
`--no-deps`: its dependencies (openfold 0.0.1, h5py, BioPython) are absent from the image; bench.py's reference arm maps
`openfold.utils.rigid_utils` to the copy transformers ships (oracle/ref_shim.py) and never touches h5py / Bio.
Run by __graft_entry__.build() when /root/reference exists (the build container); a no-op elsewhere.
"""
import filecmp
import os
import shutil
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.environ.get("PMHC_REFERENCE_ROOT", "/root/reference")
DST = os.path.join(HERE, "_ref")

SETUP = ('from setuptools import setup\n'
         'setup(name="pmhc-diffusion-model-reference", version="0.0.0", packages=["diffusion", "diffusion.tools"],\n'
         '      py_modules=["optimize", "test"])\n')


def installed() -> bool:
    return os.path.isfile(os.path.join(DST, "diffusion", "optimizer.py"))


def install(force: bool = False) -> bool:
    if not os.path.isfile(os.path.join(SRC, "diffusion", "model.py")):
        return installed()
    if installed() and not force:
        return True
    tmp = tempfile.mkdtemp(prefix="pmhc_ref_")
    try:
        copy = os.path.join(tmp, "src")
        os.makedirs(copy)
        shutil.copytree(os.path.join(SRC, "diffusion"), os.path.join(copy, "diffusion"))
        for f in ("optimize.py", "test.py"):
            shutil.copy(os.path.join(SRC, f), copy)
        with open(os.path.join(copy, "setup.py"), "w") as f:
            f.write(SETUP)
        shutil.rmtree(DST, ignore_errors=True)
        subprocess.run([sys.executable, "-m", "pip", "install", "--quiet", "--no-index", "--no-build-isolation", "--no-deps",
                        "--find-links", "/opt/wheelhouse", "--target", DST, copy], check=True)
    finally:
        shutil.rmtree(tmp, ignore_errors=True)
    cmp = filecmp.dircmp(os.path.join(SRC, "diffusion"), os.path.join(DST, "diffusion"), ignore=["__pycache__"])
    if cmp.diff_files or cmp.left_only or cmp.funny_files:
        raise RuntimeError(f"baseline/_ref differs from the reference: {cmp.diff_files} {cmp.left_only}")
    return True


if __name__ == "__main__":
    print("installed" if install(force="--force" in sys.argv) else "reference tree not present; nothing installed")
