"""Loader / writer kernels on the GPU (through the C ABI) against the structure oracle and the reference's fixture:
4x4 -> tensor_7 conversion, GPU-resident batches, peptide heavy atoms, PDB files, and the test.py flow end to end."""
import os

import numpy
import pytest
import torch

from oracle import egnn_oracle as orc
from oracle import structure_oracle as sorc
from tests.helpers import GOLDEN
from tests.test_io import same_up_to_quat_sign

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


@pytest.fixture(scope="module")
def io():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from pmhc_diffusion_model_b200 import _lib
    from pmhc_diffusion_model_b200.diffusion import data, hdf5_lite
    from pmhc_diffusion_model_b200.diffusion.tools import pdb
    _lib.check(_lib.load().pmhc_check_device(), "pmhc_check_device")

    class Io:
        pass

    o = Io()
    o.lib, o.data, o.h5, o.pdb = _lib, data, hdf5_lite, pdb
    o.golden = torch.load(os.path.join(GOLDEN, "io_golden.pt"), map_location="cpu", weights_only=False)
    return o


def random_rotations(n, seed):
    g = torch.Generator().manual_seed(seed)
    q = torch.nn.functional.normalize(torch.randn(n, 4, generator=g), dim=-1)
    m = torch.eye(4).repeat(n, 1, 1)
    m[:, :3, :3] = sorc._quat_to_rot(q)
    m[:, :3, 3] = torch.randn(n, 3, generator=g) * 20.0
    return m


def test_frames4x4_kernel_matches_reference_conversion(io):
    m = random_rotations(5000, 1)
    # half-turns and near-half-turns exercise the three trace <= 0 branches
    special = torch.eye(4).repeat(7, 1, 1)
    for k, d in enumerate(([1, -1, -1], [-1, 1, -1], [-1, -1, 1])):
        special[k, :3, :3] = torch.diag(torch.tensor(d, dtype=torch.float32))
    special[3:6, :3, :3] = sorc._quat_to_rot(torch.nn.functional.normalize(torch.tensor(
        [[1e-3, 1.0, 0.2, 0.1], [1e-3, 0.1, 1.0, 0.2], [-1e-3, 0.2, 0.1, 1.0]]), dim=-1))
    m = torch.cat((m, special))
    got = io.data.frames4x4_to_tensor7(m.to(DEV)).cpu()
    ref = sorc.tensor7_from_4x4(m.double()).float()
    assert same_up_to_quat_sign(got, ref, tol=5e-6)
    assert (got[:, 0] >= 0).all() and ((got[:, :4].norm(dim=-1) - 1).abs() < 1e-6).all()
    assert io.data.frames4x4_to_tensor7(m[:0].to(DEV)).shape == (0, 7)
    with pytest.raises(RuntimeError):
        io.data.frames4x4_to_tensor7(m)       # CPU tensor: no host path


def test_gpu_dataset_and_batches_match_reference_entries(io, tmp_path):
    g = io.golden
    path = str(tmp_path / "g.h5")
    io.h5.write_file(path, g["raw"])
    ds = io.data.MhcpDataset(path, torch.device(DEV))
    for i, ref in enumerate(g["entries"]):
        got = ds[i]
        for k, v in ref.items():
            if k == "name":
                assert got[k] == v
            elif k in ("frames", "pocket_frames"):
                assert got[k].is_cuda and same_up_to_quat_sign(got[k].cpu(), v)
            else:
                assert got[k].is_cuda and got[k].dtype == v.dtype and torch.equal(got[k].cpu(), v), k
    batches = list(ds.batches(3))
    assert [b["mask"].shape[0] for b in batches] == [3, 1]
    assert batches[0]["name"] == [g["names"][:3]] and batches[1]["name"] == [g["names"][3:]]
    stacked = {k: torch.cat([b[k] for b in batches]) for k in io.data.MhcpDataset.ORDER}
    for k in io.data.MhcpDataset.ORDER:
        ref = torch.stack([e[k] for e in g["entries"]])
        if k in ("frames", "pocket_frames"):
            assert same_up_to_quat_sign(stacked[k].cpu(), ref)
        else:
            assert torch.equal(stacked[k].cpu(), ref), k
    perm = list(ds.batches(2, shuffle=True, generator=torch.Generator().manual_seed(0)))
    assert sorted(n for b in perm for n in b["name"][0]) == sorted(g["names"])


def test_atom14_kernel_matches_reference_atoms_and_oracle(io):
    g = io.golden
    aatype = torch.stack([e["aatype"] for e in g["entries"]])
    mask = torch.stack([e["mask"] for e in g["entries"]])
    batch = {"frames": g["pdb"]["frames7"].to(DEV), "torsions": g["pdb"]["torsions"].to(DEV), "aatype": aatype.to(DEV), "mask": mask.to(DEV)}
    pos, exists = io.pdb.peptide_atoms(batch)
    t = io.pdb._host_tables()
    for b in range(aatype.shape[0]):
        ref = [a for a in g["pdb"]["atoms"][b] if a[0] == "P"]
        mine = [(i + 1, t["atom_names"][int(aatype[b, i])][a], pos[b, i, a].cpu())
                for i in torch.nonzero(mask[b]).flatten().tolist() for a in io.pdb._PEPTIDE_ORDER if exists[b, i, a]]
        assert [(m[0], m[1]) for m in mine] == [(r[1], r[3]) for r in ref]
        assert max(float((m[2] - torch.tensor(r[4])).abs().max()) for m, r in zip(mine, ref)) < 1e-4

    # ragged batch incl. a full-length 16-mer (the reference's own save() raises IndexError there, pdb.py:153) and a 1-mer
    B = 64
    gen = torch.Generator().manual_seed(3)
    L = torch.randint(1, 17, (B,), generator=gen)
    L[0], L[1] = 16, 1
    mask = torch.arange(16)[None, :] < L[:, None]
    aatype = torch.randint(0, 20, (B, 16), generator=gen)
    q = torch.randn(B, 16, 4, generator=gen)             # deliberately not unit: the torsion frames use it as stored
    q = q / q.norm(dim=-1, keepdim=True) * (1.0 + 0.01 * torch.randn(B, 16, 1, generator=gen))
    frames7 = torch.cat((q, torch.randn(B, 16, 3, generator=gen) * 8.0), -1)
    ang = torch.rand(B, 16, 7, generator=gen) * 2 * numpy.pi
    tors = torch.stack((ang.sin(), ang.cos()), -1)
    pos, exists = io.pdb.peptide_atoms({"frames": frames7.to(DEV), "torsions": tors.to(DEV), "aatype": aatype.to(DEV), "mask": mask.to(DEV)})
    rpos, rex = sorc.peptide_atoms(frames7.double(), tors.double(), aatype, mask)
    assert torch.equal(exists.cpu(), rex)
    assert float((pos.cpu() - rpos.float()).abs().max()) < 2e-4
    assert bool(exists[0, 15, 14]) and bool(exists[1, 0, 14]) and not bool(exists[0, 14, 14])


def parse_atoms(path):
    out = []
    with open(path) as fh:
        for l in fh:
            if l.startswith("ATOM"):
                out.append((l[21], int(l[22:26]), l[17:20], l[12:16].strip(), [float(l[30:38]), float(l[38:46]), float(l[46:54])]))
    return out


def test_save_writes_the_reference_atoms(io, tmp_path):
    g = io.golden
    batch = {"frames": io.data.Rigid.from_tensor_7(g["pdb"]["frames7"].to(DEV)), "torsions": g["pdb"]["torsions"].to(DEV),
             "aatype": torch.stack([e["aatype"] for e in g["entries"]]).to(DEV),
             "mask": torch.stack([e["mask"] for e in g["entries"]]).to(DEV)}
    batch.update({k: v.to(DEV) for k, v in g["protein"].items()})
    io.pdb.save(batch, 2, str(tmp_path / "one.pdb"))
    paths = io.pdb.save_batch(batch, g["names"], str(tmp_path / "all"))
    for b, p in [(2, str(tmp_path / "one.pdb"))] + list(enumerate(paths)):
        got, ref = parse_atoms(p), g["pdb"]["atoms"][b]
        assert [(a[0], a[1], a[2], a[3]) for a in got] == [(r[0], r[1], r[2], r[3]) for r in ref]
        assert max(abs(x - y) for a, r in zip(got, ref) for x, y in zip(a[4], r[4])) < 6e-4


def test_sampling_flow_from_hdf5_to_pdb(io, tmp_path):
    """test.py:57-84 end to end on a synthetic file: batches from HDF5, noise as z_T, sample(), protein atoms, PDB files."""
    from pmhc_diffusion_model_b200.diffusion.model import Model
    from pmhc_diffusion_model_b200.diffusion.optimizer import DiffusionModelOptimizer
    path = str(tmp_path / "test_set.hdf5")
    names = io.data.write_synthetic_hdf5(path, 10, peptide_len=(8, 12), protein_len=60, pocket_n=30, seed=9)
    T = 4
    model = Model(16, 22, T)
    model.load_state_dict(orc.random_params(seed=2), strict=True)
    model = model.to(DEV)
    dm = DiffusionModelOptimizer(T, model, 0.0)
    ds = io.data.MhcpDataset(path, torch.device(DEV))
    out_dir = os.path.splitext(path)[0] + "-sampled"
    written = []
    with torch.no_grad():
        for true_batch in ds.batches(4):
            batch_names = list(true_batch["name"][0])
            noise = dm.gen_noise(true_batch["frames"].shape[:-1], device=torch.device(DEV))
            inp = dict(true_batch)
            inp["frames"], inp["torsions"] = noise["frames"].to_tensor_7(), noise["torsions"]
            pred = dm.sample(inp)
            pred.update(ds.get_protein_positions(batch_names))
            written += io.pdb.save_batch(pred, batch_names, out_dir)
    assert [os.path.basename(p) for p in written] == [n + ".pdb" for n in names]
    for p in written:
        atoms = parse_atoms(p)
        assert sum(a[0] == "P" for a in atoms) > 40 and sum(a[0] == "M" for a in atoms) > 200
        assert all(numpy.isfinite(a[4]).all() for a in atoms)


def test_cli_entry_points_keep_the_reference_interface(io, tmp_path):
    """optimize.py / test.py flags and outputs (optimize.py:24-82, test.py:20-84): <model>.pth with the reference's 48
    state-dict keys, <model>.csv with one row per epoch, <hdf5 stem>-sampled/<name>.pdb."""
    import csv
    from pmhc_diffusion_model_b200.cli import optimize as cli_optimize
    from pmhc_diffusion_model_b200.cli import test as cli_test
    train = str(tmp_path / "train_set.hdf5")
    names = io.data.write_synthetic_hdf5(train, 12, peptide_len=(8, 11), protein_len=50, pocket_n=25, seed=21)
    model_path = str(tmp_path / "model.pth")
    cli_optimize.main([train, "2", model_path, "-T", "20", "-b", "5", "--seed", "1"])
    state = torch.load(model_path, map_location="cpu")
    assert len(state) == 48 and "gnn1.message_mlp.0.weight" in state and all(torch.isfinite(v).all() for v in state.values())
    rows = list(csv.reader(open(str(tmp_path / "model.csv"))))
    assert rows[0] == ["epoch", "total loss", "positions loss", "rotations loss", "torsions loss", "rmsd"]
    assert [r[0] for r in rows[1:]] == ["0", "1"] and all(float(x) == float(x) for r in rows[1:] for x in r[1:])
    cli_test.main([model_path, train, "-T", "5", "-b", "7", "--precision", "bf16", "--seed", "3"])
    out = tmp_path / "train_set-sampled"
    assert sorted(os.listdir(out)) == sorted(n + ".pdb" for n in names)
    assert all(len(parse_atoms(str(out / f))) > 300 for f in os.listdir(out))


def test_cli_sampling_with_a_seed_does_not_depend_on_the_batching(io, tmp_path):
    """`--seed`: z_T and every reverse step's noise are keyed by the GLOBAL complex index, and a row's result does not depend on
    the kernel schedule — the same PDB files come out whether the set is sampled 3 complexes at a time or all at once."""
    import shutil
    from pmhc_diffusion_model_b200.cli import test as cli_test
    from pmhc_diffusion_model_b200.diffusion.model import Model
    a = str(tmp_path / "a.hdf5")
    names = io.data.write_synthetic_hdf5(a, 11, peptide_len=(8, 12), protein_len=50, pocket_n=25, seed=33)
    b = str(tmp_path / "b.hdf5")
    shutil.copy(a, b)
    model_path = str(tmp_path / "m.pth")
    m = Model(16, 22, 6)
    m.load_state_dict(orc.random_params(seed=8), strict=True)
    torch.save(m.state_dict(), model_path)
    cli_test.main([model_path, a, "-T", "6", "-b", "3", "--gpu-batch", "3", "--seed", "5"])
    cli_test.main([model_path, b, "-T", "6", "-b", "64", "--seed", "5"])
    for n in names:
        assert open(str(tmp_path / "a-sampled" / (n + ".pdb"))).read() == open(str(tmp_path / "b-sampled" / (n + ".pdb"))).read(), n


def test_checkpoint_resume_restores_the_training_state(io, tmp_path):
    """Two epochs in one run vs one epoch, stop, resume from --checkpoint for the second: weights, Adam moments, batch order,
    noise steps and noise keys are all restored, so both runs see the same batches, t and noise.  The backward's per-complex
    accumulators use shared-memory atomics, so two identical runs agree to fp32 rounding (which Adam turns into +-lr steps),
    not bitwise: the gate is the logged losses (3 decimals) and a few lr on the weights, the same as run-to-run."""
    from pmhc_diffusion_model_b200.cli import optimize as cli_optimize
    train = str(tmp_path / "train_set.hdf5")
    io.data.write_synthetic_hdf5(train, 9, peptide_len=(8, 11), protein_len=40, pocket_n=20, seed=31)
    a, b = str(tmp_path / "a.pth"), str(tmp_path / "b.pth")
    cli_optimize.main([train, "2", a, "-T", "20", "-b", "4", "--seed", "5", "--checkpoint", str(tmp_path / "a.ckpt")])
    cli_optimize.main([train, "1", b, "-T", "20", "-b", "4", "--seed", "5", "--checkpoint", str(tmp_path / "b.ckpt")])
    cli_optimize.main([train, "2", b, "-T", "20", "-b", "4", "--seed", "5", "--checkpoint", str(tmp_path / "b.ckpt")])
    sa, sb = torch.load(a, map_location="cpu"), torch.load(b, map_location="cpu")
    assert max(float((sa[k] - sb[k]).abs().max()) for k in sa) < 6e-3        # 6 Adam steps of lr = 1e-3
    rows_a, rows_b = open(str(tmp_path / "a.csv")).read().splitlines(), open(str(tmp_path / "b.csv")).read().splitlines()
    assert len(rows_b) == 3 and rows_a[0] == rows_b[0]
    for ra, rb in zip(rows_a[1:], rows_b[1:]):
        assert all(abs(float(x) - float(y)) <= 2e-3 * max(1.0, abs(float(x))) for x, y in zip(ra.split(","), rb.split(",")))
    # a different seed gives a different run: the equality above is not trivial
    c = str(tmp_path / "c.pth")
    cli_optimize.main([train, "2", c, "-T", "20", "-b", "4", "--seed", "6"])
    sc = torch.load(c, map_location="cpu")
    assert max(float((sa[k] - sc[k]).abs().max()) for k in sa) > 2e-2
