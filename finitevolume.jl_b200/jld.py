"""Minimal reader for the JLD (HDF5) mesh files of the reference's fracture examples.

`load(path, *names)` mirrors `JLD.load(joinpath(meshdir, "mesh.jld"), "xs", "ys", ...)`
(examples/fractures/ex.jl:9, examples/fractures/ex_comparison.jl:11): it returns the named
top-level variables as numpy arrays -- Float64 / Int64 vectors as they are and
`Array{Pair{Int64,Int64},1}` (the `neighbors` list) as an (F, 2) int64 array, the layout the C ABI
takes (2F interleaved 1-based int64, src/FiniteVolume.jl:157).

JLD 0.1.x writes HDF5 with a user block holding the "Julia data file (HDF5)" banner, superblock version 0
and version-1 object headers.  Groups come in three flavours, all read here: symbol tables (v1 B-tree +
local heap + SNOD nodes), compact link messages in the object header, and -- what the mesh files' root
group uses once it holds more than eight variables -- dense link storage (link-info message -> fractal
heap; the heap's direct blocks are walked directly, the v2 B-tree that only indexes them by name is not
needed).  Datasets must be contiguous or compact little-endian fixed-point / floating-point / compound
(no chunking, filters, variable-length or reference data: the mesh files contain none); anything else
raises `JLDFormatError`.
There is no h5py in the image; this is pure Python + numpy and runs on the host.
"""
from __future__ import annotations

import struct

import numpy as np

_SIG = b"\x89HDF\r\n\x1a\n"
_UNDEF = 0xFFFFFFFFFFFFFFFF


class JLDFormatError(ValueError):
    pass


class _File:
    def __init__(self, buf: bytes):
        self.buf = buf
        # the superblock sits at 0, 512, 1024, 2048, ... (JLD: after a 512-byte user block)
        off = 0
        while True:
            if buf[off:off + 8] == _SIG:
                break
            off = 512 if off == 0 else off * 2
            if off + 8 > len(buf):
                raise JLDFormatError("no HDF5 superblock found")
        self.sb = off
        ver = buf[off + 8]
        if ver != 0:
            raise JLDFormatError(f"superblock version {ver} not supported (JLD 0.1 writes version 0)")
        self.so, self.sl = buf[off + 13], buf[off + 14]
        if (self.so, self.sl) != (8, 8):
            raise JLDFormatError("only 8-byte offsets and lengths are supported")
        self.base = self.u64(off + 24)
        root = off + 24 + 4 * 8  # root group symbol-table entry
        self.root_header = self.u64(root + 8)
        cache_type = self.u32(root + 16)
        self.root_btree, self.root_heap = (self.u64(root + 24), self.u64(root + 32)) if cache_type == 1 else (None, None)

    # ---- raw access (addresses are relative to the base address) --------------------------------
    def u8(self, o): return self.buf[o]
    def u16(self, o): return struct.unpack_from("<H", self.buf, o)[0]
    def u32(self, o): return struct.unpack_from("<I", self.buf, o)[0]
    def u64(self, o): return struct.unpack_from("<Q", self.buf, o)[0]
    def at(self, addr): return self.base + addr

    # ---- object headers (version 1) --------------------------------------------------------------
    def messages(self, addr):
        """-> [(type, flags, absolute offset of the data, size)] of the object header at `addr`."""
        o = self.at(addr)
        if self.u8(o) != 1:
            raise JLDFormatError(f"object header version {self.u8(o)} not supported")
        nmsg = self.u16(o + 2)
        blocks = [(o + 16, self.u32(o + 8))]
        out = []
        while blocks and len(out) < nmsg:
            p, size = blocks.pop(0)
            end = p + size
            while p + 8 <= end and len(out) < nmsg:
                mtype, msize, flags = self.u16(p), self.u16(p + 2), self.u8(p + 4)
                data = p + 8
                if mtype == 0x0010:  # continuation block
                    blocks.append((self.at(self.u64(data)), self.u64(data + 8)))
                out.append((mtype, flags, data, msize))
                p = data + msize
        return out

    def message(self, addr, mtype):
        for t, flags, data, size in self.messages(addr):
            if t == mtype:
                if flags & 0x02:  # shared: the message lives in another object's header
                    ver = self.u8(data)
                    target = self.u64(data + (8 if ver == 1 else 2))
                    return self.message(target, mtype)
                return data, size
        return None

    # ---- groups (symbol tables) ---------------------------------------------------------------------
    def heap_string(self, heap_addr, offset):
        h = self.at(heap_addr)
        if self.buf[h:h + 4] != b"HEAP":
            raise JLDFormatError("bad local heap signature")
        seg = self.at(self.u64(h + 24))
        end = self.buf.index(b"\0", seg + offset)
        return self.buf[seg + offset:end].decode("utf-8")

    def group_entries(self, btree_addr, heap_addr):
        """-> {name: object header address} (walks the v1 B-tree down to the SNOD leaves)."""
        out = {}
        stack = [btree_addr]
        while stack:
            a = self.at(stack.pop())
            sig = self.buf[a:a + 4]
            if sig == b"TREE":
                if self.u8(a + 4) != 0:
                    raise JLDFormatError("not a group B-tree")
                used = self.u16(a + 6)
                p = a + 8 + 16  # past the sibling pointers; keys and children alternate
                for i in range(used):
                    stack.append(self.u64(p + 8 + i * 16))
            elif sig == b"SNOD":
                n = self.u16(a + 6)
                for i in range(n):
                    e = a + 8 + i * 40
                    out[self.heap_string(heap_addr, self.u64(e))] = self.u64(e + 8)
            else:
                raise JLDFormatError(f"unexpected node signature {sig!r}")
        return out

    def link_at(self, p):
        """Parse the link message at absolute offset p -> (name, object header address | None, next offset)."""
        if self.u8(p) != 1:
            raise JLDFormatError("link message version")
        flags = self.u8(p + 1)
        q = p + 2
        ltype = 0
        if flags & 0x08:
            ltype = self.u8(q); q += 1
        if flags & 0x04:
            q += 8  # creation order
        if flags & 0x10:
            q += 1  # character set
        nlen_size = 1 << (flags & 0x03)
        nlen = int.from_bytes(self.buf[q:q + nlen_size], "little"); q += nlen_size
        name = self.buf[q:q + nlen].decode("utf-8"); q += nlen
        if ltype == 0:      # hard link
            return name, self.u64(q), q + 8
        if ltype == 1:      # soft link: length + path (not followed)
            return name, None, q + 2 + self.u16(q)
        raise JLDFormatError(f"link type {ltype} not supported")

    def dense_links(self, heap_addr):
        """Links of a group with dense storage: every managed object of the fractal heap is a link message."""
        h = self.at(heap_addr)
        if self.buf[h:h + 4] != b"FRHP" or self.u8(h + 4) != 0:
            raise JLDFormatError("bad fractal heap header")
        if self.u16(h + 7) != 0:
            raise JLDFormatError("filtered fractal heaps are not supported")
        flags = self.u8(h + 9)
        nobj = self.u64(h + 14 + 8 * 7)              # number of managed objects
        q = h + 14 + 8 * 12                          # past the twelve 8-byte statistics fields
        width, start_size = self.u16(q), self.u64(q + 2)
        max_heap_bits, root_addr, cur_rows = self.u16(q + 18), self.u64(q + 22), self.u16(q + 30)
        off_bytes = (max_heap_bits + 7) // 8
        blocks = []
        if root_addr == _UNDEF:
            return {}
        if cur_rows == 0:
            blocks.append((self.at(root_addr), start_size))
        else:
            ib = self.at(root_addr)
            if self.buf[ib:ib + 4] != b"FHIB":
                raise JLDFormatError("bad fractal heap indirect block")
            e = ib + 5 + 8 + off_bytes
            for row in range(cur_rows):
                size = start_size if row < 2 else start_size << (row - 1)
                for _ in range(width):
                    a = self.u64(e); e += 8
                    if a != _UNDEF:
                        blocks.append((self.at(a), size))
        out = {}
        for b, size in blocks:
            if self.buf[b:b + 4] != b"FHDB":
                raise JLDFormatError("bad fractal heap direct block (nested indirect blocks are not supported)")
            p = b + 5 + 8 + off_bytes + (4 if flags & 0x02 else 0)
            end = b + size
            while p < end and len(out) < nobj and self.u8(p) == 1:
                name, addr, p = self.link_at(p)
                if addr is not None:
                    out[name] = addr
        return out

    def root_entries(self):
        msgs = self.messages(self.root_header)
        out = {}
        for t, flags, data, size in msgs:
            if t == 0x0011:    # symbol table
                out.update(self.group_entries(self.u64(data), self.u64(data + 8)))
            elif t == 0x0006:  # compact link
                name, addr, _ = self.link_at(data)
                if addr is not None:
                    out[name] = addr
            elif t == 0x0002:  # link info: dense storage if a fractal heap is attached
                q = data + 2 + (8 if self.u8(data + 1) & 0x01 else 0)
                heap = self.u64(q)
                if heap != _UNDEF:
                    out.update(self.dense_links(heap))
        if not out and self.root_btree is not None:
            out = self.group_entries(self.root_btree, self.root_heap)
        return out

    # ---- datasets ------------------------------------------------------------------------------------
    def dtype_at(self, p):
        """numpy dtype of the datatype message at absolute offset p -> (dtype, bytes consumed)."""
        cls, ver = self.u8(p) & 0x0F, self.u8(p) >> 4
        bits = self.u8(p + 1) | (self.u8(p + 2) << 8) | (self.u8(p + 3) << 16)
        size = self.u32(p + 4)
        if cls == 0:  # fixed point
            if bits & 1:
                raise JLDFormatError("big-endian integers not supported")
            return np.dtype(f"<{'i' if bits & 0x08 else 'u'}{size}"), 8 + 4
        if cls == 1:  # floating point
            if bits & 1:
                raise JLDFormatError("big-endian floats not supported")
            return np.dtype(f"<f{size}"), 8 + 12
        if cls == 6:  # compound
            nmemb = bits & 0xFFFF
            q = p + 8
            names, formats, offsets = [], [], []
            for _ in range(nmemb):
                end = self.buf.index(b"\0", q)
                name = self.buf[q:end].decode("utf-8")
                if ver < 3:
                    q += (end - q + 8) // 8 * 8  # name padded to a multiple of 8 (incl. terminator)
                    off = self.u32(q)
                    q += 4
                    if ver == 1:
                        q += 1 + 3 + 4 + 4 + 16  # dimensionality, reserved, permutation, reserved, 4 dims
                else:
                    q = end + 1
                    nb = max(1, (size.bit_length() + 7) // 8)
                    off = int.from_bytes(self.buf[q:q + nb], "little")
                    q += nb
                dt, used = self.dtype_at(q)
                q += used
                names.append(name); formats.append(dt); offsets.append(off)
            return np.dtype({"names": names, "formats": formats, "offsets": offsets, "itemsize": size}), q - p
        raise JLDFormatError(f"datatype class {cls} not supported")

    def dataset(self, addr):
        sp = self.message(addr, 0x0001)
        ty = self.message(addr, 0x0003)
        lay = self.message(addr, 0x0008)
        if sp is None or ty is None or lay is None:
            raise JLDFormatError("object is not a simple dataset")
        p = sp[0]
        ver, rank, flags = self.u8(p), self.u8(p + 1), self.u8(p + 2)
        dims_at = p + (8 if ver == 1 else 4)
        dims = [self.u64(dims_at + 8 * i) for i in range(rank)]
        dt, _ = self.dtype_at(ty[0])
        p = lay[0]
        lver = self.u8(p)
        if lver != 3:
            raise JLDFormatError(f"data layout version {lver} not supported")
        lclass = self.u8(p + 1)
        count = int(np.prod(dims)) if rank else 1
        if lclass == 1:  # contiguous
            daddr, dsize = self.u64(p + 2), self.u64(p + 10)
            if daddr == _UNDEF:
                return np.zeros(dims, dt)
            if dsize < count * dt.itemsize:
                raise JLDFormatError("dataset smaller than its dataspace")
            arr = np.frombuffer(self.buf, dtype=dt, count=count, offset=self.at(daddr))
        elif lclass == 0:  # compact
            dsize = self.u16(p + 2)
            arr = np.frombuffer(self.buf, dtype=dt, count=count, offset=p + 4)
        else:
            raise JLDFormatError("chunked datasets are not supported")
        # HDF5 dims are row-major; Julia wrote its column-major array with the dims reversed
        return arr.reshape(dims).T.copy() if rank > 1 else arr.copy()


def _simplify(a):
    """Pair{Int64,Int64} arrays (compound of two int64) -> (n, 2) int64; everything else as is."""
    if a.dtype.names and len(a.dtype.names) == 2 and all(a.dtype[n] == np.dtype("<i8") for n in a.dtype.names):
        return np.stack([a[a.dtype.names[0]], a[a.dtype.names[1]]], axis=-1).astype(np.int64)
    return a


def names(path):
    """Top-level variable names of a JLD file (JLD's own bookkeeping entries `_creator`, `_refs`, `_types`,
    `_require` excluded)."""
    with open(path, "rb") as f:
        F = _File(f.read())
    return sorted(n for n in F.root_entries() if not n.startswith("_"))


def load(path, *want):
    """`JLD.load(path, names...)`: the named variables in order (one name -> the array itself; no name ->
    a dict of everything readable)."""
    with open(path, "rb") as f:
        F = _File(f.read())
    entries = F.root_entries()
    if not want:
        out = {}
        for n, addr in entries.items():
            if n.startswith("_"):
                continue
            try:
                out[n] = _simplify(F.dataset(addr))
            except JLDFormatError:
                pass
        return out
    res = []
    for n in want:
        if n not in entries:
            raise KeyError(f"{n!r} not found in {path}")
        res.append(_simplify(F.dataset(entries[n])))
    return res[0] if len(res) == 1 else tuple(res)
