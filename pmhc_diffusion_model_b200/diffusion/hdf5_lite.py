"""hdf5_lite — a small reader / writer for the HDF5 subset the SwiftMHC-format input uses (README.md:11-37 of the
reference: a root group of complexes, each with `peptide` / `protein` groups of fixed-type numeric datasets).

The reference reads its input with h5py (diffusion/data.py:5, :28, :38).  This image has neither h5py nor libhdf5, so the
drop-in `MhcpDataset` (data.py in this package) parses the file itself:

  reader  superblock v0-v3; object headers v1 and v2 (with continuation blocks); old-style groups (symbol-table message ->
          v1 B-tree + local heap + symbol nodes, what h5py writes by default) and new-style groups with compact link
          messages; datasets with contiguous, compact or chunked (v1 chunk B-tree) layout; filters deflate (1),
          shuffle (2) and LZF (32000); datatypes: fixed-point, floating-point, enum (h5py's bool) -> numpy.
  writer  the plain "earliest" format: superblock v0, symbol-table groups, contiguous datasets — enough to produce
          synthetic SwiftMHC files that libhdf5 / h5py also open.

Anything else (dense link storage, variable-length strings, compound types, virtual datasets ...) raises
`Hdf5Unsupported` naming the feature.  `open_file(path)` returns an h5py.File when h5py is importable, else `File`.
Format reference: "HDF5 File Format Specification Version 3.0" (The HDF Group); nothing here derives from h5py sources.
"""
import struct
import zlib
from typing import Dict, Iterator, List, Optional, Tuple, Union

import numpy

UNDEF = 0xFFFFFFFFFFFFFFFF
SIGNATURE = b"\x89HDF\r\n\x1a\n"


class Hdf5Unsupported(NotImplementedError):
    pass


def open_file(path: str):
    """h5py.File(path, 'r') when h5py exists, else the reader below (same `keys()`, `[]`, `in`, `[:]` surface)."""
    try:
        import h5py  # noqa: F401
        if getattr(h5py, "File", None) is not None and getattr(h5py, "__version__", None):
            return h5py.File(path, "r")
    except ImportError:
        pass
    return File(path)


# ---------------------------------------------------------------------------------------------------------------------
# reader
# ---------------------------------------------------------------------------------------------------------------------

def _lzf_decompress(src: bytes, out_len: int) -> bytes:
    """LZF (Marc Lehmann's liblzf, HDF5 filter 32000): control byte < 32 -> literal run of ctrl + 1 bytes; else a back
    reference of length (ctrl >> 5) + 2 (7 -> one more length byte) at distance ((ctrl & 31) << 8 | next) + 1."""
    out = bytearray()
    i, n = 0, len(src)
    while i < n:
        ctrl = src[i]
        i += 1
        if ctrl < 32:
            out += src[i:i + ctrl + 1]
            i += ctrl + 1
        else:
            length = ctrl >> 5
            if length == 7:
                length += src[i]
                i += 1
            ref = len(out) - (((ctrl & 0x1F) << 8) | src[i]) - 1
            i += 1
            if ref < 0:
                raise ValueError("corrupt LZF stream")
            for _ in range(length + 2):   # may overlap its own output
                out.append(out[ref])
                ref += 1
    if len(out) != out_len:
        raise ValueError(f"LZF stream decoded to {len(out)} bytes, expected {out_len}")
    return bytes(out)


class _Buf:
    """Random access over the file image."""

    def __init__(self, data: bytes, base: int, so: int, sl: int):
        self.d, self.base, self.so, self.sl = data, base, so, sl

    def u(self, off: int, size: int) -> int:
        return int.from_bytes(self.d[off:off + size], "little")

    def addr(self, off: int) -> int:
        v = self.u(off, self.so)
        return UNDEF if v == (1 << (8 * self.so)) - 1 else v + self.base

    def length(self, off: int) -> int:
        return self.u(off, self.sl)

    def addr_from(self, body: bytes, off: int) -> int:
        v = int.from_bytes(body[off:off + self.so], "little")
        return UNDEF if v == (1 << (8 * self.so)) - 1 else v + self.base


def _parse_datatype(b: bytes) -> Tuple[numpy.dtype, int, Optional[dict]]:
    """datatype message -> (numpy dtype, bytes consumed, enum members or None)."""
    cls, ver = b[0] & 0x0F, b[0] >> 4
    bits = b[1] | (b[2] << 8) | (b[3] << 16)
    size = struct.unpack_from("<I", b, 4)[0]
    order = ">" if bits & 1 else "<"
    if cls == 0:    # fixed point: properties bit offset (2), precision (2)
        kind = "i" if bits & 0x08 else "u"
        return numpy.dtype(f"{order}{kind}{size}"), 8 + 4, None
    if cls == 1:    # floating point: 12 bytes of properties
        if size not in (2, 4, 8):
            raise Hdf5Unsupported(f"{size}-byte floating-point datatype")
        return numpy.dtype(f"{order}f{size}"), 8 + 12, None
    if cls == 8:    # enumeration over a fixed-point base (h5py stores numpy.bool_ as ENUM{FALSE = 0, TRUE = 1} of int8)
        n = bits & 0xFFFF
        base, used, _ = _parse_datatype(b[8:])
        p = 8 + used
        names = []
        for _ in range(n):
            end = b.index(b"\0", p)
            names.append(b[p:end].decode())
            p = end + 1 if ver >= 3 else p + ((end - p + 8) // 8) * 8
        values = [int(numpy.frombuffer(b[p + k * base.itemsize:p + (k + 1) * base.itemsize], dtype=base)[0]) for k in range(n)]
        p += n * base.itemsize
        members = dict(zip(names, values))
        is_bool = set(n_.upper() for n_ in names) == {"FALSE", "TRUE"} and base.itemsize == 1
        return (numpy.dtype("bool") if is_bool else base), p, members
    names = {2: "time", 3: "string", 4: "bitfield", 5: "opaque", 6: "compound", 7: "reference", 9: "variable-length", 10: "array"}
    raise Hdf5Unsupported(f"HDF5 datatype class {cls} ({names.get(cls, '?')})")


class _Object:
    """An object header parsed into its messages: list of (type, flags, body bytes)."""

    def __init__(self, f: "File", address: int):
        self.f, self.address = f, address
        self.messages: List[Tuple[int, int, bytes]] = []
        b = f._b
        if b.d[address:address + 4] == b"OHDR":
            self._parse_v2(address)
        else:
            self._parse_v1(address)

    def _parse_v1(self, a: int) -> None:
        b = self.f._b
        if b.d[a] != 1:
            raise Hdf5Unsupported(f"object header version {b.d[a]} at {a}")
        nmsg = b.u(a + 2, 2)
        size = b.u(a + 8, 4)
        blocks = [(a + 16, size)]
        while blocks and len(self.messages) < nmsg:
            p, left = blocks.pop(0)
            end = p + left
            while p + 8 <= end and len(self.messages) < nmsg:
                mtype, msize, flags = b.u(p, 2), b.u(p + 2, 2), b.d[p + 4]
                body = b.d[p + 8:p + 8 + msize]
                p += 8 + msize
                if mtype == 0x0010:     # continuation: address, length
                    blocks.append((b.addr_from(body, 0), int.from_bytes(body[b.so:b.so + b.sl], "little")))
                self.messages.append((mtype, flags, body))

    def _parse_v2(self, a: int) -> None:
        b = self.f._b
        flags = b.d[a + 5]
        p = a + 6
        if flags & 0x20:
            p += 16             # access, modification, change, birth times
        if flags & 0x10:
            p += 4              # max compact / min dense attributes
        csize = 1 << (flags & 3)
        chunk0 = b.u(p, csize)
        p += csize
        track = bool(flags & 0x04)
        blocks = [(p, chunk0)]
        while blocks:
            p, left = blocks.pop(0)
            end = p + left
            while p + 4 + (2 if track else 0) <= end:
                mtype, msize, mflags = b.d[p], b.u(p + 1, 2), b.d[p + 3]
                p += 4 + (2 if track else 0)
                body = b.d[p:p + msize]
                p += msize
                if mtype == 0x10:
                    ca, cl = b.addr_from(body, 0), int.from_bytes(body[b.so:b.so + b.sl], "little")
                    blocks.append((ca + 4, cl - 8))    # skip "OCHK", stop before the checksum
                if mtype != 0 or msize:
                    self.messages.append((mtype, mflags, body))

    def find(self, mtype: int) -> Optional[bytes]:
        for t, _, body in self.messages:
            if t == mtype:
                return body
        return None


class Group:
    def __init__(self, f: "File", obj: _Object, name: str):
        self._f, self._obj, self.name = f, obj, name
        self._links: Optional[Dict[str, int]] = None

    # -- link table ---------------------------------------------------------------------------------------------------
    def _load(self) -> Dict[str, int]:
        if self._links is not None:
            return self._links
        links: Dict[str, int] = {}
        b = self._f._b
        st = self._obj.find(0x0011)
        if st is not None:                                   # old style: v1 B-tree of symbol nodes + local heap
            btree, heap = b.addr_from(st, 0), b.addr_from(st, b.so)
            if b.d[heap:heap + 4] != b"HEAP":
                raise ValueError("bad local heap signature")
            heap_data = b.addr(heap + 8 + 2 * b.sl)
            self._walk_group_btree(btree, heap_data, links)
        else:
            for t, _, body in self._obj.messages:
                if t == 0x0006:                              # link message (compact new-style group)
                    name, addr = self._parse_link(body)
                    if addr is not None:
                        links[name] = addr
                elif t == 0x0002:                            # link info: dense storage -> fractal heap
                    fheap = b.addr_from(body, 2 + (8 if body[1] & 1 else 0))
                    if fheap != UNDEF:
                        raise Hdf5Unsupported("groups with dense link storage (fractal heap); re-save with libver='earliest'")
        self._links = links
        return links

    def _walk_group_btree(self, node: int, heap_data: int, links: Dict[str, int]) -> None:
        b = self._f._b
        if b.d[node:node + 4] == b"SNOD":
            n = b.u(node + 6, 2)
            p = node + 8
            for _ in range(n):
                name_off = b.u(p, b.so)
                addr = b.addr(p + b.so)
                end = b.d.index(b"\0", heap_data + name_off)
                links[b.d[heap_data + name_off:end].decode()] = addr
                p += 2 * b.so + 4 + 4 + 16
            return
        if b.d[node:node + 4] != b"TREE" or b.d[node + 4] != 0:
            raise ValueError("bad group B-tree node")
        used = b.u(node + 6, 2)
        p = node + 8 + 2 * b.so + b.sl                        # past the siblings and key 0
        for _ in range(used):
            self._walk_group_btree(b.addr(p), heap_data, links)
            p += b.so + b.sl

    def _parse_link(self, body: bytes) -> Tuple[str, Optional[int]]:
        b = self._f._b
        flags = body[1]
        p = 2
        ltype = 0
        if flags & 0x08:
            ltype = body[p]
            p += 1
        if flags & 0x04:
            p += 8
        if flags & 0x10:
            p += 1
        lsz = 1 << (flags & 3)
        nlen = int.from_bytes(body[p:p + lsz], "little")
        p += lsz
        name = body[p:p + nlen].decode()
        p += nlen
        return name, (b.addr_from(body, p) if ltype == 0 else None)   # soft / external links are skipped

    # -- mapping surface ----------------------------------------------------------------------------------------------
    def keys(self) -> List[str]:
        return list(self._load().keys())

    def __iter__(self) -> Iterator[str]:
        return iter(self.keys())

    def __len__(self) -> int:
        return len(self._load())

    def __contains__(self, name: str) -> bool:
        head, _, rest = name.strip("/").partition("/")
        links = self._load()
        if head not in links:
            return False
        return True if not rest else rest in self[head]

    def __getitem__(self, name: str) -> Union["Group", "Dataset"]:
        head, _, rest = name.strip("/").partition("/")
        links = self._load()
        if head not in links:
            raise KeyError(f"{head!r} not in {self.name!r}")
        child = self._f._open(links[head], (self.name.rstrip("/") + "/" + head))
        return child[rest] if rest else child


class Dataset:
    def __init__(self, f: "File", obj: _Object, name: str):
        self._f, self._obj, self.name = f, obj, name
        space = obj.find(0x0001)
        dtype = obj.find(0x0003)
        if space is None or dtype is None:
            raise ValueError(f"{name}: not a dataset")
        ver, rank = space[0], space[1]
        p = 8 if ver == 1 else 4
        self.shape = tuple(int.from_bytes(space[p + k * f._b.sl:p + (k + 1) * f._b.sl], "little") for k in range(rank))
        self.dtype, _, self.enum = _parse_datatype(dtype)
        self._storage = numpy.dtype("uint8") if self.dtype == numpy.dtype("bool") else self.dtype   # element type on disk

    def __len__(self) -> int:
        return self.shape[0]

    @property
    def size(self) -> int:
        return int(numpy.prod(self.shape, dtype=numpy.int64)) if self.shape else 1

    def _filters(self) -> List[Tuple[int, List[int]]]:
        body = self._obj.find(0x000B)
        if body is None:
            return []
        ver, n = body[0], body[1]
        p = 8 if ver == 1 else 2
        out = []
        for _ in range(n):
            fid = int.from_bytes(body[p:p + 2], "little")
            p += 2
            nlen = 0
            if ver == 1 or fid >= 256:
                nlen = int.from_bytes(body[p:p + 2], "little")
                p += 2
            p += 2   # flags
            ncv = int.from_bytes(body[p:p + 2], "little")
            p += 2
            if nlen:
                p += ((nlen + 7) // 8) * 8 if ver == 1 else nlen
            cd = [int.from_bytes(body[p + 4 * k:p + 4 * k + 4], "little") for k in range(ncv)]
            p += 4 * ncv
            if ver == 1 and ncv % 2:
                p += 4
            out.append((fid, cd))
        return out

    def _unfilter(self, raw: bytes, mask: int, filters, nbytes: int) -> bytes:
        for k in reversed(range(len(filters))):
            if mask & (1 << k):
                continue
            fid, cd = filters[k]
            if fid == 1:
                raw = zlib.decompress(raw)
            elif fid == 2:
                es = cd[0] if cd else self._storage.itemsize
                a = numpy.frombuffer(raw, dtype=numpy.uint8)
                n = len(a) // es
                raw = a[:n * es].reshape(es, n).T.tobytes() + a[n * es:].tobytes()
            elif fid == 32000:
                raw = _lzf_decompress(raw, nbytes) if len(raw) != nbytes else raw
            elif fid == 3:
                raw = raw[:-4]   # fletcher32: checksum dropped, not verified
            else:
                raise Hdf5Unsupported(f"HDF5 filter {fid}")
        return raw

    def _read(self) -> numpy.ndarray:
        b = self._f._b
        layout = self._obj.find(0x0008)
        if layout is None:
            raise ValueError(f"{self.name}: no data layout message")
        ver = layout[0]
        if ver not in (3, 4):
            raise Hdf5Unsupported(f"data layout message version {ver}")
        cls = layout[1]
        nbytes = self.size * self._storage.itemsize
        if cls == 0:      # compact
            n = int.from_bytes(layout[2:4], "little")
            raw = layout[4:4 + n]
        elif cls == 1:    # contiguous
            addr = b.addr_from(layout, 2)
            raw = b"\0" * nbytes if addr == UNDEF else b.d[addr:addr + nbytes]
        elif cls == 2:    # chunked
            if ver == 4:
                raise Hdf5Unsupported("version-4 chunk indexing (libver='latest'); re-save with the default libver")
            rank1 = layout[2]
            btree = b.addr_from(layout, 3)
            cdims = [int.from_bytes(layout[3 + b.so + 4 * k:3 + b.so + 4 * k + 4], "little") for k in range(rank1)]
            return self._read_chunked(btree, cdims[:-1])
        else:
            raise Hdf5Unsupported(f"data layout class {cls}")
        if len(raw) < nbytes:
            raise ValueError(f"{self.name}: file truncated")
        return numpy.frombuffer(raw[:nbytes], dtype=self._storage).reshape(self.shape)

    def _read_chunked(self, btree: int, cdims: List[int]) -> numpy.ndarray:
        b = self._f._b
        out = numpy.zeros(self.shape, dtype=self._storage)
        if btree == UNDEF:
            return out
        filters = self._filters()
        rank = len(self.shape)
        cbytes = int(numpy.prod(cdims, dtype=numpy.int64)) * self._storage.itemsize

        def walk(node: int) -> None:
            if b.d[node:node + 4] != b"TREE" or b.d[node + 4] != 1:
                raise ValueError("bad chunk B-tree node")
            level, used = b.d[node + 5], b.u(node + 6, 2)
            p = node + 8 + 2 * b.so
            ksize = 8 + 8 * (rank + 1)
            for _ in range(used):
                csize, mask = b.u(p, 4), b.u(p + 4, 4)
                offs = [b.u(p + 8 + 8 * k, 8) for k in range(rank)]
                child = b.addr(p + ksize)
                p += ksize + b.so
                if level > 0:
                    walk(child)
                    continue
                raw = self._unfilter(b.d[child:child + csize], mask, filters, cbytes)
                chunk = numpy.frombuffer(raw[:cbytes], dtype=self._storage).reshape(cdims)
                sl = tuple(slice(o, min(o + c, s)) for o, c, s in zip(offs, cdims, self.shape))
                out[sl] = chunk[tuple(slice(0, s.stop - s.start) for s in sl)]

        walk(btree)
        return out

    def __getitem__(self, key) -> numpy.ndarray:
        a = self._read()
        if self.dtype == numpy.dtype("bool"):
            a = a.astype(bool)
        elif not a.dtype.isnative:
            a = a.astype(a.dtype.newbyteorder("="))
        if key is Ellipsis or (isinstance(key, slice) and key == slice(None)) or (isinstance(key, tuple) and len(key) == 0):
            return numpy.array(a)
        return numpy.array(a[key])


class File(Group):
    """Read-only HDF5 file; a context manager like h5py.File."""

    def __init__(self, path: str, mode: str = "r"):
        if mode != "r":
            raise ValueError("hdf5_lite.File is read-only; use hdf5_lite.write_file to create files")
        with open(path, "rb") as fh:
            data = fh.read()
        start = 0
        while data[start:start + 8] != SIGNATURE:       # the superblock may sit at 0, 512, 1024, ...
            start = 512 if start == 0 else start * 2
            if start >= len(data):
                raise ValueError(f"{path}: not an HDF5 file")
        ver = data[start + 8]
        if ver in (0, 1):
            so, sl = data[start + 13], data[start + 14]
            p = start + 24 + (4 if ver == 1 else 0)
            self._b = _Buf(data, 0, so, sl)
            base = self._b.u(p, so)
            self._b.base = base
            root_entry = p + 4 * so
            root = self._b.addr(root_entry + so)
        elif ver in (2, 3):
            so, sl = data[start + 9], data[start + 10]
            self._b = _Buf(data, 0, so, sl)
            self._b.base = self._b.u(start + 12, so)
            root = self._b.addr(start + 12 + 3 * so)
        else:
            raise Hdf5Unsupported(f"superblock version {ver}")
        self._cache: Dict[int, Union[Group, Dataset]] = {}
        self.filename = path
        super().__init__(self, _Object(self, root), "/")

    def _open(self, address: int, name: str) -> Union[Group, Dataset]:
        hit = self._cache.get(address)
        if hit is not None:
            return hit
        obj = _Object(self, address)
        node: Union[Group, Dataset]
        if obj.find(0x0008) is not None or obj.find(0x0001) is not None:
            node = Dataset(self, obj, name)
        else:
            node = Group(self, obj, name)
        self._cache[address] = node
        return node

    def close(self) -> None:
        pass

    def __enter__(self) -> "File":
        return self

    def __exit__(self, *exc) -> None:
        self.close()


# ---------------------------------------------------------------------------------------------------------------------
# writer: superblock v0, symbol-table groups, contiguous datasets
# ---------------------------------------------------------------------------------------------------------------------

LEAF_K, INTERNAL_K = 4, 16      # libhdf5's defaults: <= 8 symbols per symbol node, <= 32 children per B-tree node


def _pad8(b: bytes) -> bytes:
    return b + b"\0" * (-len(b) % 8)


def _dtype_message(dt: numpy.dtype) -> bytes:
    dt = numpy.dtype(dt)
    if dt == numpy.dtype("bool"):
        base = _dtype_message(numpy.dtype("int8"))
        names = _pad8(b"FALSE\0") + _pad8(b"TRUE\0")
        return struct.pack("<BBBBI", 0x18, 2, 0, 0, 1) + base + names + bytes([0, 1])
    if dt.kind in "iu":
        return struct.pack("<BBBBIHH", 0x10, 0x08 if dt.kind == "i" else 0x00, 0, 0, dt.itemsize, 0, 8 * dt.itemsize)
    if dt.kind == "f" and dt.itemsize in (4, 8):
        exp_bits, man_bits = (8, 23) if dt.itemsize == 4 else (11, 52)
        return struct.pack("<BBBBIHHBBBBI", 0x11, 0x20, 8 * dt.itemsize - 1, 0, dt.itemsize, 0, 8 * dt.itemsize,
                           man_bits, exp_bits, 0, man_bits, (1 << (exp_bits - 1)) - 1)
    raise Hdf5Unsupported(f"writer: dtype {dt}")


class _Writer:
    def __init__(self):
        self.buf = bytearray(96)        # superblock (56 bytes) + root symbol table entry (40 bytes), filled in last

    def alloc(self, data: bytes) -> int:
        self.buf += b"\0" * (-len(self.buf) % 8)
        at = len(self.buf)
        self.buf += data
        return at

    @staticmethod
    def message(mtype: int, body: bytes, flags: int = 0) -> bytes:
        body = _pad8(body)
        return struct.pack("<HHB3x", mtype, len(body), flags) + body

    def object_header(self, messages: List[bytes]) -> int:
        data = b"".join(messages)
        return self.alloc(struct.pack("<BBHII4x", 1, 0, len(messages), 1, len(data)) + data)

    def dataset(self, a: numpy.ndarray) -> int:
        a = numpy.ascontiguousarray(a)
        if a.dtype.kind in "iuf" and not a.dtype.isnative:
            a = a.astype(a.dtype.newbyteorder("="))
        raw = a.astype(numpy.uint8).tobytes() if a.dtype == numpy.dtype("bool") else a.tobytes()
        data_at = self.alloc(raw) if raw else UNDEF
        space = struct.pack("<BBB5x", 1, a.ndim, 0) + b"".join(struct.pack("<Q", s) for s in a.shape)
        fill = struct.pack("<BBBB", 2, 2, 0, 0)      # version 2, allocate late, write at allocation, undefined
        layout = struct.pack("<BBQQ", 3, 1, data_at, len(raw))
        return self.object_header([self.message(0x0001, space), self.message(0x0003, _dtype_message(a.dtype), 1),
                                   self.message(0x0005, fill), self.message(0x0008, layout)])

    def group(self, children: Dict[str, int]) -> Tuple[int, int, int]:
        """-> (object header address, B-tree address, heap address)."""
        names = sorted(children)                      # symbol nodes are ordered by name (strcmp)
        heap = bytearray(b"\0" * 8)                   # offset 0: the empty string (key 0 of every B-tree)
        offs = {}
        for n in names:
            offs[n] = len(heap)
            heap += _pad8(n.encode() + b"\0")
        heap_data = self.alloc(bytes(heap))
        heap_at = self.alloc(b"HEAP" + struct.pack("<B3xQQQ", 0, len(heap), 1, heap_data))   # free-list head 1 = none
        # leaves: symbol nodes of <= 2 * LEAF_K entries, allocated at full size
        level: List[Tuple[int, int]] = []             # (address, heap offset of the largest name below)
        for s in range(0, max(len(names), 1), 2 * LEAF_K):
            part = names[s:s + 2 * LEAF_K]
            body = b"SNOD" + struct.pack("<BBH", 1, 0, len(part))
            for n in part:
                body += struct.pack("<QQII16x", offs[n], children[n], 0, 0)
            body += b"\0" * ((2 * LEAF_K - len(part)) * 40)
            level.append((self.alloc(body), offs[part[-1]] if part else 0))
        depth = 0
        while True:
            nxt: List[Tuple[int, int]] = []
            nodes = [level[s:s + 2 * INTERNAL_K] for s in range(0, len(level), 2 * INTERNAL_K)]
            addrs = []
            for part in nodes:
                body = bytearray(b"TREE" + struct.pack("<BBHQQ", 0, depth, len(part), UNDEF, UNDEF))
                body += struct.pack("<Q", 0)          # key 0: the empty string
                for addr, key in part:
                    body += struct.pack("<QQ", addr, key)
                body += b"\0" * ((2 * INTERNAL_K - len(part)) * 16)
                addrs.append(self.alloc(bytes(body)))
                nxt.append((addrs[-1], part[-1][1]))
            for k, at in enumerate(addrs):            # sibling links
                left = addrs[k - 1] if k > 0 else UNDEF
                right = addrs[k + 1] if k + 1 < len(addrs) else UNDEF
                self.buf[at + 8:at + 24] = struct.pack("<QQ", left, right)
            if len(nxt) == 1:
                btree = nxt[0][0]
                break
            level, depth = nxt, depth + 1
        header = self.object_header([self.message(0x0011, struct.pack("<QQ", btree, heap_at))])
        return header, btree, heap_at

    def tree(self, node: dict) -> Tuple[int, int, int]:
        children = {}
        for name, value in node.items():
            children[name] = self.tree(value)[0] if isinstance(value, dict) else self.dataset(numpy.asarray(value))
        return self.group(children)

    def finish(self, root: Tuple[int, int, int]) -> bytes:
        header, btree, heap = root
        eof = len(self.buf)
        sb = SIGNATURE + struct.pack("<BBBBBBBBHHI", 0, 0, 0, 0, 0, 8, 8, 0, LEAF_K, INTERNAL_K, 0)
        sb += struct.pack("<QQQQ", 0, UNDEF, eof, UNDEF)
        sb += struct.pack("<QQII", 0, header, 1, 0) + struct.pack("<QQ", btree, heap)   # root entry, cache type 1
        self.buf[0:len(sb)] = sb
        return bytes(self.buf)


def write_file(path: str, tree: dict) -> None:
    """tree: nested dicts (groups) of array-likes (datasets).  bool arrays become h5py-style FALSE/TRUE enums."""
    w = _Writer()
    data = w.finish(w.tree(tree))
    with open(path, "wb") as fh:
        fh.write(data)
