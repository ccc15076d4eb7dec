"""Loader / writer rows on the CPU: the HDF5 subset reader / writer, `MhcpDataset` against the fixture
produced by the unmodified reference (tests/golden/make_golden_io.py), the structure oracle pinned to the atoms the
reference's `pdb.save` builds, and the PDB text layout."""
import os
import struct
import zlib

import numpy
import pytest
import torch

from oracle import structure_oracle as sorc
from pmhc_diffusion_model_b200.diffusion import hdf5_lite
from pmhc_diffusion_model_b200.diffusion.data import MhcpDataset, write_synthetic_hdf5
from pmhc_diffusion_model_b200.diffusion.tools import pdb as pdbio
from tests.helpers import GOLDEN


@pytest.fixture(scope="module")
def golden():
    return torch.load(os.path.join(GOLDEN, "io_golden.pt"), map_location="cpu", weights_only=False)


def same_up_to_quat_sign(a, b, tol=2e-6):
    q = (a[..., :4] * b[..., :4]).sum(-1, keepdim=True).sign()
    return bool(((a[..., :4] * q - b[..., :4]).abs().max() < tol) and ((a[..., 4:] - b[..., 4:]).abs().max() < tol))


# ---- HDF5 subset ------------------------------------------------------------------------------------------------------

def test_hdf5_roundtrip_types_and_many_groups(tmp_path):
    rng = numpy.random.default_rng(0)
    tree = {}
    for i in range(300):       # > 256 entries: symbol nodes, two leaf-level B-tree nodes and a root above them
        tree[f"cx{i:04d}"] = {"peptide": {"f4": rng.standard_normal((9, 4, 4)).astype("f4"), "i8": rng.integers(0, 20, 9),
                                          "b": rng.random((9, 7)) > 0.5},
                              "protein": {"f8": rng.standard_normal((12, 14, 3)), "empty": numpy.zeros((0, 3), "f4"),
                                          "u1": rng.integers(0, 255, 5).astype("u1"), "i4": rng.integers(-5, 5, 4).astype("i4")}}
    path = str(tmp_path / "t.h5")
    hdf5_lite.write_file(path, tree)
    with hdf5_lite.File(path) as f:
        assert sorted(f.keys()) == sorted(tree) and len(f) == 300
        assert "cx0001/peptide/f4" in f and "nope" not in f["cx0001"]
        for k in ("cx0000", "cx0137", "cx0299"):
            for g, members in tree[k].items():
                assert sorted(f[k][g].keys()) == sorted(members)
                for d, v in members.items():
                    got = f[k][g][d][:]
                    assert got.dtype == numpy.asarray(v).dtype and got.shape == numpy.asarray(v).shape and (got == v).all()
        with pytest.raises(KeyError):
            f["missing"]
    with pytest.raises(ValueError):
        bad = tmp_path / "bad.h5"
        bad.write_bytes(b"not hdf5" * 100)
        hdf5_lite.File(str(bad))


def test_lzf_decoder_known_stream():
    # literal "abcab" then a back reference of length 7 at distance 5 (overlapping its own output), then literal "Z"
    stream = bytes([4]) + b"abcab" + bytes([(5 << 5) | 0, 4]) + bytes([0]) + b"Z"
    assert hdf5_lite._lzf_decompress(stream, 13) == b"abcab" + b"abcabab" + b"Z"
    long_ref = bytes([0]) + b"x" + bytes([(7 << 5) | 0, 3, 0])            # length 7 + 3 + 2 = 12 copies of 'x'
    assert hdf5_lite._lzf_decompress(long_ref, 13) == b"x" * 13
    with pytest.raises(ValueError):
        hdf5_lite._lzf_decompress(stream, 12)


def test_chunked_deflate_shuffle_dataset(tmp_path):
    """A chunked dataset (v1 chunk B-tree, shuffle + deflate, ragged edge chunks) laid out by hand per the format spec."""
    w = hdf5_lite._Writer()
    a = numpy.arange(7 * 5, dtype="<f4").reshape(7, 5) * 1.5
    cdims = (4, 3)
    entries = []
    for r0 in range(0, 7, 4):
        for c0 in range(0, 5, 3):
            chunk = numpy.zeros(cdims, "<f4")
            blk = a[r0:r0 + 4, c0:c0 + 3]
            chunk[:blk.shape[0], :blk.shape[1]] = blk
            raw = numpy.frombuffer(chunk.tobytes(), numpy.uint8).reshape(-1, 4).T.tobytes()      # shuffle
            comp = zlib.compress(raw)
            entries.append((len(comp), (r0, c0, 0), w.alloc(comp)))
    node = b"TREE" + struct.pack("<BBHQQ", 1, 0, len(entries), hdf5_lite.UNDEF, hdf5_lite.UNDEF)
    for size, offs, addr in entries:
        node += struct.pack("<II3Q", size, 0, *offs) + struct.pack("<Q", addr)
    node += struct.pack("<II3Q", 0, 0, 8, 6, 0)
    btree = w.alloc(node)
    space = struct.pack("<BBB5xQQ", 1, 2, 0, 7, 5)
    layout = struct.pack("<BBBQIII", 3, 2, 3, btree, 4, 3, 4)
    filt = struct.pack("<BB6x", 1, 2) + struct.pack("<HHHHI4x", 2, 0, 0, 1, 4) + struct.pack("<HHHHI4x", 1, 0, 0, 1, 6)
    ds = w.object_header([w.message(1, space), w.message(3, hdf5_lite._dtype_message(a.dtype), 1), w.message(0x0B, filt),
                          w.message(8, layout)])
    path = str(tmp_path / "c.h5")
    with open(path, "wb") as fh:
        fh.write(w.finish(w.group({"x": ds})))
    with hdf5_lite.File(path) as f:
        assert f["x"].shape == (7, 5)
        assert (f["x"][:] == a).all()


# ---- dataset ----------------------------------------------------------------------------------------------------------

def test_dataset_matches_reference_entries(golden, tmp_path):
    path = str(tmp_path / "g.h5")
    hdf5_lite.write_file(path, golden["raw"])
    ds = MhcpDataset(path)
    assert ds.entry_names == golden["names"] and len(ds) == len(golden["names"])
    for i, ref in enumerate(golden["entries"]):
        got = ds[i]
        assert list(got.keys()) == list(ref.keys())
        for k, v in ref.items():
            if k == "name":
                assert got[k] == v
            elif k in ("frames", "pocket_frames"):
                assert got[k].shape == v.shape and same_up_to_quat_sign(got[k], v)
            else:
                assert got[k].dtype == v.dtype and got[k].shape == v.shape and torch.equal(got[k], v), k
    prot = ds.get_protein_positions(golden["names"][1:3])
    for k, v in golden["protein"].items():
        assert torch.equal(prot[k], v[1:3]) and prot[k].dtype == v.dtype
    with pytest.raises(KeyError):
        ds.get_entry("nope")


def test_dataset_missing_peptide_raises(tmp_path):
    path = str(tmp_path / "m.h5")
    write_synthetic_hdf5(path, 1, seed=3)
    with hdf5_lite.File(path) as f:
        name = f.keys()[0]
        tree = {name: {"protein": {k: f[name]["protein"][k][:] for k in f[name]["protein"].keys()}}}
    hdf5_lite.write_file(path, tree)
    with pytest.raises(ValueError, match="no peptide"):
        MhcpDataset(path)[0]


def test_load_all_stacks_entries(tmp_path):
    path = str(tmp_path / "s.h5")
    names = write_synthetic_hdf5(path, 5, peptide_len=(8, 15), protein_len=30, pocket_n=20, seed=4)
    ds = MhcpDataset(path)
    host = ds.load_all(pin=False)
    assert host["frames4x4"].shape == (5, 16, 4, 4) and host["pocket_frames4x4"].shape == (5, 80, 4, 4)
    for i in range(5):
        e = ds.get_entry(names[i])
        assert torch.equal(host["torsions"][i], e["torsions"]) and torch.equal(host["pocket_mask"][i], e["pocket_mask"])
        assert int(host["mask"][i].sum()) == int(e["mask"].sum())


# ---- oracle pinned to the reference's pdb.save; text layout --------------------------------------------------------------

def golden_atoms(golden, i):
    return golden["pdb"]["atoms"][i]


def test_structure_oracle_matches_reference_atoms(golden):
    entries = golden["entries"]
    aatype = torch.stack([e["aatype"] for e in entries])
    mask = torch.stack([e["mask"] for e in entries])
    pos, exists = sorc.peptide_atoms(golden["pdb"]["frames7"], golden["pdb"]["torsions"], aatype, mask)
    t = pdbio._host_tables()
    for b in range(len(entries)):
        ref = [a for a in golden_atoms(golden, b) if a[0] == "P"]
        mine = []
        for i in torch.nonzero(mask[b]).flatten().tolist():
            aa = int(aatype[b, i])
            for a in pdbio._PEPTIDE_ORDER:
                if exists[b, i, a]:
                    mine.append(("P", i + 1, t["names3"][aa], t["atom_names"][aa][a], pos[b, i, a]))
        assert [(m[0], m[1], m[2], m[3]) for m in mine] == [(r[0], r[1], r[2], r[3]) for r in ref]
        err = max(float((m[4] - torch.tensor(r[4])).abs().max()) for m, r in zip(mine, ref))
        assert err < 1e-4, err


def test_tensor7_oracle_is_the_reference_conversion(golden):
    raw = golden["raw"]
    for name, ref in zip(golden["names"], golden["entries"]):
        m = torch.from_numpy(raw[name]["peptide"]["backbone_rigid_tensor"])
        L = m.shape[0]
        assert same_up_to_quat_sign(sorc.tensor7_from_4x4(m), ref["frames"][:L])


def test_pdb_text_layout(golden):
    entries = golden["entries"]
    b = 0
    aatype, mask = entries[b]["aatype"], entries[b]["mask"]
    pos, exists = sorc.peptide_atoms(golden["pdb"]["frames7"][b:b + 1], golden["pdb"]["torsions"][b:b + 1], aatype[None], mask[None])
    prot = golden["protein"]
    text = pdbio.format_pdb(aatype.numpy(), mask.numpy(), pos[0].numpy(), exists[0].numpy(), prot["protein_aatype"][b].numpy(),
                            prot["protein_atom14_positions"][b].numpy(), prot["protein_atom14_exists"][b].numpy().astype(bool))
    # the library's host-side formatter and the line-by-line Python statement of the layout give the same bytes
    assert text == pdbio.format_pdb_python(aatype.numpy(), mask.numpy(), pos[0].numpy(), exists[0].numpy(), prot["protein_aatype"][b].numpy(),
                                           prot["protein_atom14_positions"][b].numpy(), prot["protein_atom14_exists"][b].numpy().astype(bool))
    lines = text.splitlines()
    assert lines[-1] == "END   " and all(len(l) == 80 for l in lines[:-1])
    atoms = [l for l in lines if l.startswith("ATOM")]
    ref = golden_atoms(golden, b)
    assert len(atoms) == len(ref)
    assert [int(l[6:11]) for l in lines[:-1]] == list(range(1, len(lines)))          # serial numbers run through TER records
    for l, r in zip(atoms, ref):
        assert l[21] == r[0] and int(l[22:26]) == r[1] and l[17:20] == r[2] and l[12:16].strip() == r[3]
        assert l[76:78].strip() == r[3][0] and l[54:60] == "  1.00" and l[60:66] == "  0.00"
        xyz = [float(l[30:38]), float(l[38:46]), float(l[46:54])]
        assert max(abs(x - y) for x, y in zip(xyz, r[4])) < 6e-4
    ters = [l for l in lines if l.startswith("TER")]
    assert len(ters) == 2 and ters[0][21] == "P" and ters[1][21] == "M"
    assert atoms[0][12:16] == " N  " and any(l[12:16] == " CA " for l in atoms)


def test_pdb_number_formatting_matches_printf():
    """The host formatter writes "%8.3f" digits itself: ties (half-to-even on the exact value), negative zero, values that
    round to zero, column overflow and non-finite values must come out as Python's / printf's formatting does."""
    t = pdbio._host_tables()
    rng = numpy.random.default_rng(1)
    aa = rng.integers(0, 20, 16)
    mask = numpy.arange(16) < 9
    ex = numpy.concatenate([numpy.asarray(t["atom_mask"])[aa], numpy.zeros((16, 1), numpy.uint8)], 1).astype(bool)
    ex[:, 3], ex[8, 14] = True, True
    paa = rng.integers(0, 21, 40)
    pex = numpy.asarray(t["atom_mask"])[paa].astype(bool)
    ppos = (rng.standard_normal((40, 14, 3)) * 30).astype("f4")
    ppos[0, 0] = [0.0625, -0.0625, 0.1875]
    ppos[1, 0] = [-0.0004, 0.0005, -1e-9]
    ppos[2, 0] = [9999.9, -999.9, 123456.7]
    ppos[3, 0] = [2.5, -3.0005, 0.9995]
    ppos[4, 0] = [0.0, -0.0, 1234.5675]
    ppos[5, 0] = [float("nan"), float("inf"), -999.9995]
    for scale in (0.01, 1.0, 100.0, 3000.0):
        pos = (rng.standard_normal((16, 15, 3)) * scale).astype("f4")
        args = (aa, mask, pos, ex, paa, ppos * (scale if scale < 1000 else 1.0), pex)
        assert pdbio.format_pdb(*args) == pdbio.format_pdb_python(*args)


def test_hdf5_random_trees_round_trip(tmp_path):
    """Property test: any tree of nested groups with numeric / bool datasets of random shape written by the writer reads
    back identically (names incl. non-ASCII-free punctuation, empty groups, scalars-as-1-element arrays, group sizes
    around the symbol-node / B-tree fan-out boundaries 8 and 256)."""
    from hypothesis import given, settings, strategies as st
    dtypes = ["<f4", "<f8", "<i8", "<i4", "<i2", "|i1", "|u1", "<u2", "<u4", "bool"]
    names = st.text(alphabet="abcdefghijklmnopqrstuvwxyzABCDEFGHIJKLMNOPQRSTUVWXYZ0123456789_-. ", min_size=1, max_size=24).filter(
        lambda n: n.strip("/ ") == n and n not in (".", ".."))

    @st.composite
    def arrays(draw):
        dt = numpy.dtype(draw(st.sampled_from(dtypes)))
        shape = tuple(draw(st.lists(st.integers(0, 5), min_size=1, max_size=3)))
        n = int(numpy.prod(shape))
        seed = draw(st.integers(0, 2 ** 31 - 1))
        rng = numpy.random.default_rng(seed)
        if dt == numpy.dtype("bool"):
            return rng.random(n).reshape(shape) > 0.5
        if dt.kind == "f":
            return rng.standard_normal(n).astype(dt).reshape(shape)
        info = numpy.iinfo(dt)
        return rng.integers(info.min, info.max, n, dtype=dt, endpoint=True).reshape(shape)

    def trees(depth):
        leaf = arrays()
        if depth == 0:
            return st.dictionaries(names, leaf, max_size=4)
        return st.dictionaries(names, st.one_of(leaf, trees(depth - 1)), max_size=5)

    counter = {"n": 0}

    @settings(max_examples=40, deadline=None)
    @given(trees(2), st.sampled_from([0, 7, 8, 9, 17, 255, 256, 257]))
    def check(tree, extra):
        tree = dict(tree)
        for k in range(extra):        # pad the root group to sizes around the fan-out boundaries
            tree.setdefault(f"pad{k:04d}", numpy.arange(k % 3, dtype="<i4"))
        counter["n"] += 1
        path = str(tmp_path / f"r{counter['n']}.h5")
        hdf5_lite.write_file(path, tree)

        def compare(node, ref):
            assert sorted(node.keys()) == sorted(ref.keys())
            for k, v in ref.items():
                if isinstance(v, dict):
                    compare(node[k], v)
                else:
                    got = node[k][:]
                    assert got.dtype == v.dtype and got.shape == v.shape and numpy.array_equal(got, v), k

        with hdf5_lite.File(path) as f:
            compare(f, tree)
        os.remove(path)

    check()
