"""A small pure-Python reader (and fixture writer) for the HDF5 files of the reference's training set.

The reference stores every sample as one contiguous, uncompressed u8 dataset `(6, H, W)` under the group
`datum`, with the sample's meta data as a JSON string in the attribute `meta`
(training/generate_hdf5_coco2014.py:356-366) and reads it back through h5py
(py_rmpe_server/py_rmpe_data_iterator.py:17-19, 46-66).  h5py is not part of this image, so
`RawDataIterator` falls back to this module: the subset of the HDF5 File Format Specification that such
files use, exposed with the handful of h5py names the reference touches --

    f = File(path, "r"); g = f["datum"]; list(g.keys()); ds = g[key]
    ds.attrs["meta"]; "meta" in ds.attrs; ds[()] / ds.value; ds.shape; ds.dtype; f.close()

Supported on disk: superblock versions 0-3; version 1 and 2 object headers (with continuation blocks);
old-style groups (symbol table = v1 B-tree + local heap + symbol nodes) and compact new-style groups (link
messages); dataspace versions 1-2; fixed-point, floating-point, fixed-length string and variable-length
string datatypes; contiguous and compact layouts (a chunked or filtered dataset raises NotImplementedError
and names h5py as the way to read it); attribute message versions 1-3; the global heap (variable-length
strings).  Datasets are returned as numpy arrays read straight from the file (memory-mapped).

`write_datum_file` writes the same structures (superblock 0, symbol-table groups, version 1 headers,
contiguous data, variable-length string attributes in a global heap collection): enough for the committed
test fixture, laid out like the HDF5 library's "earliest" format.
"""
import json
import mmap
import struct

import numpy as np

SIGNATURE = b"\x89HDF\r\n\x1a\n"
UNDEF = 0xFFFFFFFFFFFFFFFF


class H5FormatError(ValueError):
    pass


def _pad8(n):
    return (n + 7) & ~7


# ------------------------------------------------------------------------------------------
# reader
# ------------------------------------------------------------------------------------------
class _Reader:
    def __init__(self, path):
        self.fh = open(path, "rb")
        try:
            self.buf = mmap.mmap(self.fh.fileno(), 0, access=mmap.ACCESS_READ)
        except ValueError:
            self.fh.close()
            raise H5FormatError("empty file")
        self._superblock()

    def close(self):
        if self.buf is not None:
            try:
                self.buf.close()
            except BufferError:        # numpy views of the mapping are still alive: the mapping goes with them
                pass
            self.buf = None
            self.fh.close()

    # -- primitive access (file addresses are relative to the superblock's base address) --
    def at(self, addr, n):
        a = self.base + addr
        if addr == self.undef or a + n > len(self.buf):
            raise H5FormatError("address 0x%x (+%d) outside the file" % (addr, n))
        return self.buf[a:a + n]

    def u(self, data, off, size):
        return int.from_bytes(data[off:off + size], "little")

    def _superblock(self):
        pos = 0
        while True:
            if self.buf[pos:pos + 8] == SIGNATURE:
                break
            pos = 512 if pos == 0 else pos * 2
            if pos + 8 > len(self.buf):
                raise H5FormatError("not an HDF5 file (no superblock signature)")
        sb = self.buf
        ver = sb[pos + 8]
        self.sb_version = ver
        if ver in (0, 1):
            self.O, self.L = sb[pos + 13], sb[pos + 14]
            p = pos + 24 + (4 if ver == 1 else 0)
            O = self.O
            self.undef = (1 << (8 * O)) - 1
            self.base = int.from_bytes(sb[p:p + O], "little")
            if self.base == 0 and pos:
                self.base = pos            # a user block in front of the superblock
            p += 4 * O                     # base, free-space, end-of-file, driver-info addresses
            # root group symbol table entry
            self.root_header = int.from_bytes(sb[p + O:p + 2 * O], "little")
        elif ver in (2, 3):
            self.O, self.L = sb[pos + 9], sb[pos + 10]
            O = self.O
            self.undef = (1 << (8 * O)) - 1
            p = pos + 12
            self.base = int.from_bytes(sb[p:p + O], "little")
            if self.base == 0 and pos:
                self.base = pos
            self.root_header = int.from_bytes(sb[p + 3 * O:p + 4 * O], "little")
        else:
            raise H5FormatError("unsupported superblock version %d" % ver)

    # -- object headers --
    def messages(self, addr):
        """[(type, flags, bytes)] of the object header at addr, continuation blocks followed."""
        head = self.at(addr, 16)
        out = []
        if head[:4] == b"OHDR":
            return self._messages_v2(addr)
        if head[0] != 1:
            raise H5FormatError("unsupported object header version %d at 0x%x" % (head[0], addr))
        nmsg = self.u(head, 2, 2)
        size = self.u(head, 8, 4)
        blocks = [(addr + 16, size)]
        while blocks and len(out) < nmsg:
            start, length = blocks.pop(0)
            data = self.at(start, length)
            p = 0
            while p + 8 <= length and len(out) < nmsg:
                mtype, msize, flags = self.u(data, p, 2), self.u(data, p + 2, 2), data[p + 4]
                body = bytes(data[p + 8:p + 8 + msize])
                p += 8 + msize
                if mtype == 0x10:
                    blocks.append((self.u(body, 0, self.O), self.u(body, self.O, self.L)))
                out.append((mtype, flags, body))
        return out

    def _messages_v2(self, addr):
        head = self.at(addr, 6)
        flags = head[5]
        p = 6
        if flags & 0x20:
            p += 16                         # access, modification, change, birth times
        if flags & 0x10:
            p += 4                          # max compact / min dense attributes
        csize_bytes = 1 << (flags & 3)
        chunk0 = self.u(self.at(addr + p, csize_bytes), 0, csize_bytes)
        p += csize_bytes
        track = bool(flags & 0x04)
        out = []
        blocks = [(addr + p, chunk0)]
        while blocks:
            start, length = blocks.pop(0)
            data = self.at(start, length)
            q = 0
            hdr = 4 + (2 if track else 0)
            while q + hdr <= length:
                mtype, msize, mflags = data[q], self.u(data, q + 1, 2), data[q + 3]
                body = bytes(data[q + hdr:q + hdr + msize])
                q += hdr + msize
                if mtype == 0x10:
                    coff, clen = self.u(body, 0, self.O), self.u(body, self.O, self.L)
                    blocks.append((coff + 4, clen - 8))     # "OCHK" signature in front, checksum behind
                elif mtype != 0:
                    out.append((mtype, mflags, body))
        return out

    # -- groups --
    def group_links(self, addr):
        """{name: object header address} of the group whose object header is at addr."""
        links = {}
        for mtype, _, body in self.messages(addr):
            if mtype == 0x11:               # symbol table: v1 B-tree + local heap
                btree, heap = self.u(body, 0, self.O), self.u(body, self.O, self.O)
                self._walk_btree(btree, self._local_heap(heap), links)
            elif mtype == 0x06:             # link message (compact new-style group)
                name, target = self._link(body)
                if target is not None:
                    links[name] = target
            elif mtype == 0x02:             # link info: dense storage needs fractal heaps
                fheap = self.u(body, 2 + (8 if body[1] & 1 else 0), self.O)
                if fheap != self.undef:
                    raise NotImplementedError("this group stores its links in a fractal heap (HDF5 'latest' format with "
                                              "many links); read the file with h5py")
        return links

    def _local_heap(self, addr):
        h = self.at(addr, 8 + 2 * self.L + self.O)
        if h[:4] != b"HEAP":
            raise H5FormatError("local heap signature missing at 0x%x" % addr)
        size = self.u(h, 8, self.L)
        data_addr = self.u(h, 8 + 2 * self.L, self.O)
        return self.at(data_addr, size)

    def _walk_btree(self, addr, heap, links):
        O, L = self.O, self.L
        h = self.at(addr, 8 + 2 * O)
        if h[:4] != b"TREE" or h[4] != 0:
            raise H5FormatError("group B-tree node expected at 0x%x" % addr)
        level, used = h[5], self.u(h, 6, 2)
        body = self.at(addr + 8 + 2 * O, used * (L + O) + L)
        for i in range(used):
            child = self.u(body, L + i * (L + O), O)
            if level > 0:
                self._walk_btree(child, heap, links)
            else:
                self._symbol_node(child, heap, links)

    def _symbol_node(self, addr, heap, links):
        O = self.O
        h = self.at(addr, 8)
        if h[:4] != b"SNOD":
            raise H5FormatError("symbol table node expected at 0x%x" % addr)
        n = self.u(h, 6, 2)
        esz = 2 * O + 24
        body = self.at(addr + 8, n * esz)
        for i in range(n):
            name_off = self.u(body, i * esz, O)
            header = self.u(body, i * esz + O, O)
            end = heap.find(b"\0", name_off)
            links[bytes(heap[name_off:end]).decode("utf-8")] = header

    def _link(self, body):
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
        nlen = self.u(body, p, lsz)
        p += lsz
        name = body[p:p + nlen].decode("utf-8")
        p += nlen
        if ltype != 0:
            return name, None               # soft / external links are not followed
        return name, self.u(body, p, self.O)

    # -- datasets and attributes --
    def dataspace(self, body):
        ver, rank, flags = body[0], body[1], body[2]
        if ver == 1:
            p = 8
        elif ver == 2:
            p = 4
            if body[3] == 2:                # null dataspace
                return None
        else:
            raise H5FormatError("unsupported dataspace version %d" % ver)
        return tuple(self.u(body, p + i * self.L, self.L) for i in range(rank))

    def datatype(self, body):
        """-> ("numpy", dtype) | ("string", size, charset) | ("vlen_string", charset)"""
        cls, ver = body[0] & 0x0F, body[0] >> 4
        b0, b1 = body[1], body[2]
        size = self.u(body, 4, 4)
        if cls == 0:
            order = ">" if b0 & 1 else "<"
            return ("numpy", np.dtype("%s%s%d" % (order, "i" if b0 & 8 else "u", size)))
        if cls == 1:
            order = ">" if b0 & 1 else "<"
            return ("numpy", np.dtype("%sf%d" % (order, size)))
        if cls == 3:
            return ("string", size, "utf-8" if (b0 >> 4) & 0xF == 1 else "ascii")
        if cls == 9:
            if b0 & 0x0F != 1:
                raise NotImplementedError("variable-length sequences (only variable-length strings are supported)")
            return ("vlen_string", "utf-8" if b1 & 0x0F == 1 else "ascii")
        raise NotImplementedError("HDF5 datatype class %d (version %d)" % (cls, ver))

    def global_heap_object(self, addr, index):
        L = self.L
        h = self.at(addr, 8 + L)
        if h[:4] != b"GCOL":
            raise H5FormatError("global heap collection expected at 0x%x" % addr)
        size = self.u(h, 8, L)
        data = self.at(addr, size)
        p = 8 + L
        while p + 8 + L <= size:
            idx, osz = self.u(data, p, 2), self.u(data, p + 8, L)
            if idx == index:
                return bytes(data[p + 8 + L:p + 8 + L + osz])
            if idx == 0:
                break
            p += 8 + L + _pad8(osz)
        raise H5FormatError("object %d not in the global heap collection at 0x%x" % (index, addr))

    def decode_values(self, dtype, shape, raw):
        n = 1
        for d in (shape or ()):
            n *= d
        if dtype[0] == "numpy":
            a = np.frombuffer(raw, dtype=dtype[1], count=n)
            return a.reshape(shape) if shape else a[0]
        if dtype[0] == "string":
            size = dtype[1]
            vals = [bytes(raw[i * size:(i + 1) * size]).split(b"\0")[0].decode(dtype[2], "replace") for i in range(n)]
        else:
            esz = 4 + self.O + 4
            vals = []
            for i in range(n):
                ln = self.u(raw, i * esz, 4)
                gaddr, gidx = self.u(raw, i * esz + 4, self.O), self.u(raw, i * esz + 4 + self.O, 4)
                vals.append("" if gaddr in (0, self.undef) and ln == 0 else
                            self.global_heap_object(gaddr, gidx)[:ln].decode(dtype[1], "replace"))
        if not shape:
            return vals[0]
        return np.array(vals, dtype=object).reshape(shape)

    def attribute(self, body):
        ver = body[0]
        nsz, tsz, ssz = self.u(body, 2, 2), self.u(body, 4, 2), self.u(body, 6, 2)
        if ver == 1:
            p = 8
            name = body[p:p + nsz].split(b"\0")[0].decode("utf-8")
            p += _pad8(nsz)
            dt = self.datatype(body[p:p + tsz])
            p += _pad8(tsz)
            shape = self.dataspace(body[p:p + ssz])
            p += _pad8(ssz)
        elif ver in (2, 3):
            if body[1] & 3:
                raise NotImplementedError("attributes with shared datatype / dataspace messages")
            p = 8 + (1 if ver == 3 else 0)
            name = body[p:p + nsz].split(b"\0")[0].decode("utf-8")
            p += nsz
            dt = self.datatype(body[p:p + tsz])
            p += tsz
            shape = self.dataspace(body[p:p + ssz])
            p += ssz
        else:
            raise H5FormatError("unsupported attribute message version %d" % ver)
        return name, self.decode_values(dt, shape, body[p:])


class _Attrs(dict):
    """dict with h5py's spelling (`'meta' in ds.attrs`, `ds.attrs['meta']`)."""


class Dataset:
    def __init__(self, reader, addr, name):
        self._r, self.name = reader, name
        self.attrs = _Attrs()
        self.shape, self._dtype, self._layout = None, None, None
        filtered = False
        for mtype, _, body in reader.messages(addr):
            if mtype == 0x01:
                self.shape = reader.dataspace(body)
            elif mtype == 0x03:
                self._dtype = reader.datatype(body)
            elif mtype == 0x08:
                self._layout = body
            elif mtype == 0x0B:
                filtered = True
            elif mtype == 0x0C:
                k, v = reader.attribute(body)
                self.attrs[k] = v
            elif mtype == 0x15:
                O = reader.O
                fheap = reader.u(body, 2 + (2 if body[1] & 1 else 0), O)
                if fheap != reader.undef:
                    raise NotImplementedError("attributes in dense storage (fractal heap); read the file with h5py")
        if self._layout is None or self._dtype is None:
            raise H5FormatError("%s is not a dataset" % name)
        self._filtered = filtered

    @property
    def dtype(self):
        return self._dtype[1] if self._dtype[0] == "numpy" else np.dtype(object)

    def _read(self):
        r, lay = self._r, self._layout
        n = 1
        for d in (self.shape or ()):
            n *= d
        if lay[0] in (1, 2):                # layout message of the 1.6 library: version, rank, class, 5 reserved, address
            if lay[2] != 1:
                raise NotImplementedError("dataset %s: only contiguous data of layout version %d is read here; use h5py" %
                                          (self.name, lay[0]))
            addr = r.u(lay, 8, r.O)
            a = np.frombuffer(r.buf, dtype=self._dtype[1], count=n, offset=r.base + addr) if self._dtype[0] == "numpy" \
                else r.decode_values(self._dtype, self.shape, r.at(addr, n * self._dtype[1]))
            return a.reshape(self.shape) if self._dtype[0] == "numpy" else a
        if lay[0] != 3:
            raise NotImplementedError("data layout message version %d; read the file with h5py" % lay[0])
        if lay[1] == 1:                     # contiguous
            addr, size = r.u(lay, 2, r.O), r.u(lay, 2 + r.O, r.L)
            if addr == r.undef:             # never written: fill value (zeros)
                return np.zeros(self.shape, self.dtype)
            if self._dtype[0] == "numpy":
                a = np.frombuffer(r.buf, dtype=self._dtype[1], count=n, offset=r.base + addr)
                return a.reshape(self.shape)
            return r.decode_values(self._dtype, self.shape, r.at(addr, size))
        if lay[1] == 0:                     # compact
            size = r.u(lay, 2, 2)
            return r.decode_values(self._dtype, self.shape, lay[4:4 + size])
        raise NotImplementedError("dataset %s is chunked%s: the reference writes contiguous data (chunks=None, "
                                  "generate_hdf5_coco2014.py:358); read this file with h5py" %
                                  (self.name, " and filtered" if self._filtered else ""))

    def __getitem__(self, key):
        a = self._read()
        if key == () or key is Ellipsis:
            return a
        return a[key]

    @property
    def value(self):                        # h5py < 3 spelling, used by the reference (py_rmpe_data_iterator.py:54)
        return self._read()


class Group:
    def __init__(self, reader, addr, name="/"):
        self._r, self._addr, self.name = reader, addr, name
        self._links = None

    def _load(self):
        if self._links is None:
            self._links = self._r.group_links(self._addr)
        return self._links

    def keys(self):
        return sorted(self._load().keys())

    def __iter__(self):
        return iter(self.keys())

    def __len__(self):
        return len(self._load())

    def __contains__(self, name):
        return name in self._load()

    def __getitem__(self, name):
        node = self
        for part in [p for p in name.split("/") if p]:
            links = node._load()
            if part not in links:
                raise KeyError(name)
            addr = links[part]
            path = node.name.rstrip("/") + "/" + part
            types = {m[0] for m in node._r.messages(addr)}
            node = Dataset(node._r, addr, path) if 0x08 in types else Group(node._r, addr, path)
        return node


class File(Group):
    """h5py.File(path, 'r') for the files described in the module docstring."""

    def __init__(self, path, mode="r"):
        if mode != "r":
            raise NotImplementedError("h5lite.File is read-only; write_datum_file() writes fixture files")
        r = _Reader(path)
        super().__init__(r, r.root_header, "/")
        self.filename = path

    def close(self):
        self._r.close()

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()


# ------------------------------------------------------------------------------------------
# writer (fixtures): superblock 0, symbol-table groups, v1 headers, contiguous data, vlen-string attributes
# ------------------------------------------------------------------------------------------
class _Writer:
    LEAF_K, INTERNAL_K = 4, 16

    def __init__(self):
        self.data = bytearray()

    def alloc(self, n, align=8):
        pad = (-len(self.data)) % align
        self.data += b"\0" * pad
        off = len(self.data)
        self.data += b"\0" * n
        return off

    def put(self, off, raw):
        self.data[off:off + len(raw)] = raw

    @staticmethod
    def message(mtype, body, flags=0):
        body = body + b"\0" * (_pad8(len(body)) - len(body))
        return struct.pack("<HHB3x", mtype, len(body), flags) + body

    def object_header(self, messages):
        raw = b"".join(messages)
        off = self.alloc(16 + len(raw))
        self.put(off, struct.pack("<BxHII4x", 1, len(messages), 1, len(raw)) + raw)
        return off

    @staticmethod
    def dataspace(shape):
        return struct.pack("<BBB5x", 1, len(shape), 0) + b"".join(struct.pack("<Q", d) for d in shape)

    @staticmethod
    def datatype(dt):
        dt = np.dtype(dt)
        if dt.kind in "iu":
            bits = (8 if dt.kind == "i" else 0) | (1 if dt.byteorder == ">" else 0)
            return struct.pack("<BBBBI", 0x10 | 0, bits, 0, 0, dt.itemsize) + struct.pack("<HH", 0, 8 * dt.itemsize)
        if dt.kind == "f" and dt.itemsize in (4, 8):
            # IEEE little-endian: sign position, exponent / mantissa location and size, exponent bias
            if dt.itemsize == 4:
                props = struct.pack("<HHBBBBI", 0, 32, 23, 8, 0, 23, 127)
                return struct.pack("<BBBBI", 0x10 | 1, 0x20, 31, 0, 4) + props
            props = struct.pack("<HHBBBBI", 0, 64, 52, 11, 0, 52, 1023)
            return struct.pack("<BBBBI", 0x10 | 1, 0x20, 63, 0, 8) + props
        raise NotImplementedError(str(dt))

    @staticmethod
    def vlen_string_type(utf8=True):
        base = struct.pack("<BBBBI", 0x10 | 3, 0x00, 0, 0, 1)               # 1-byte null-terminated ASCII character
        return struct.pack("<BBBBI", 0x10 | 9, 0x01, 1 if utf8 else 0, 0, 16) + base

    def global_heap(self, objects):
        """One collection holding `objects` (bytes each); returns (address, [index of each object])."""
        body = b""
        for i, o in enumerate(objects):
            body += struct.pack("<HH4xQ", i + 1, 1, len(o)) + o + b"\0" * (_pad8(len(o)) - len(o))
        size = max(4096, _pad8(16 + len(body) + 16))
        free = size - 16 - len(body)
        body += struct.pack("<HH4xQ", 0, 0, free - 16) + b"\0" * (free - 16)
        off = self.alloc(size)
        self.put(off, b"GCOL" + struct.pack("<B3xQ", 1, size) + body)
        return off, list(range(1, len(objects) + 1))

    def group(self, entries):
        """entries: {name: object header address}; returns the group's object header address."""
        names = sorted(entries, key=lambda s: s.encode("utf-8"))
        heap_data = bytearray(b"\0" * 8)                  # offset 0: the empty name (first B-tree key)
        name_off = {}
        for n in names:
            name_off[n] = len(heap_data)
            raw = n.encode("utf-8") + b"\0"
            heap_data += raw + b"\0" * (_pad8(len(raw)) - len(raw))
        free_off = len(heap_data)
        heap_data += b"\0" * max(16, 88 - len(heap_data) % 8)
        struct.pack_into("<QQ", heap_data, free_off, 1, len(heap_data) - free_off)     # one free block to the end
        heap_seg = self.alloc(len(heap_data))
        self.put(heap_seg, bytes(heap_data))
        heap = self.alloc(32)
        self.put(heap, b"HEAP" + struct.pack("<B3xQQQ", 0, len(heap_data), free_off, heap_seg))
        per = 2 * self.LEAF_K
        chunks = [names[i:i + per] for i in range(0, len(names), per)] or [[]]
        if len(chunks) > 2 * self.INTERNAL_K:
            raise NotImplementedError("write_datum_file holds at most %d entries per group" % (per * 2 * self.INTERNAL_K))
        snods = []
        for ch in chunks:
            off = self.alloc(8 + per * 40)
            raw = b"SNOD" + struct.pack("<BxH", 1, len(ch))
            for n in ch:
                raw += struct.pack("<QQII16x", name_off[n], entries[n], 0, 0)
            self.put(off, raw)
            snods.append(off)
        tree = self.alloc(24 + (2 * self.INTERNAL_K) * 16 + 8)
        raw = b"TREE" + struct.pack("<BBHQQ", 0, 0, len(snods), UNDEF, UNDEF) + struct.pack("<Q", 0)
        for ch, off in zip(chunks, snods):
            raw += struct.pack("<QQ", off, name_off[ch[-1]] if ch else 0)
        self.put(tree, raw)
        header = self.object_header([self.message(0x11, struct.pack("<QQ", tree, heap))])
        return header, tree, heap

    def dataset(self, array, attrs):
        array = np.ascontiguousarray(array)
        raw = array.tobytes()
        data_off = self.alloc(len(raw))
        self.put(data_off, raw)
        msgs = [self.message(0x01, self.dataspace(array.shape)),
                self.message(0x03, self.datatype(array.dtype), flags=1),
                self.message(0x05, struct.pack("<BBBB", 2, 2, 2, 0)),      # fill value: late allocation, write if set, undefined
                self.message(0x08, struct.pack("<BBQQ", 3, 1, data_off, len(raw)))]
        if attrs:
            strings = [v.encode("utf-8") for v in attrs.values()]
            gaddr, idx = self.global_heap(strings)
            for (k, _), s, i in zip(attrs.items(), strings, idx):
                name = k.encode("utf-8") + b"\0"
                dt, sp = self.vlen_string_type(), struct.pack("<BBB5x", 1, 0, 0)
                body = struct.pack("<BxHHH", 1, len(name), len(dt), len(sp))
                body += name + b"\0" * (_pad8(len(name)) - len(name))
                body += dt + b"\0" * (_pad8(len(dt)) - len(dt))
                body += sp + b"\0" * (_pad8(len(sp)) - len(sp))
                body += struct.pack("<IQI", len(s), gaddr, i)
                msgs.append(self.message(0x0C, body))
        return self.object_header(msgs)


def write_datum_file(path, samples, group="datum"):
    """samples: {key: (array (6,H,W) u8, meta dict or JSON string)} -> an HDF5 file shaped like the reference's
    training set: /<group>/<key> contiguous datasets with a variable-length string attribute 'meta'."""
    w = _Writer()
    w.alloc(96)                                           # superblock (version 0, 8-byte offsets and lengths)
    entries = {}
    for key, (arr, meta) in samples.items():
        meta = meta if isinstance(meta, str) else json.dumps(meta)
        entries[key] = w.dataset(arr, {"meta": meta})
    ghdr, _, _ = w.group(entries)
    rhdr, rtree, rheap = w.group({group: ghdr})
    eof = len(w.data)
    sb = SIGNATURE + struct.pack("<BBBxBBBxHHI", 0, 0, 0, 0, 8, 8, _Writer.LEAF_K, _Writer.INTERNAL_K, 0)
    sb += struct.pack("<QQQQ", 0, UNDEF, eof, UNDEF)
    sb += struct.pack("<QQII", 0, rhdr, 1, 0) + struct.pack("<QQ", rtree, rheap)      # root entry caches its B-tree and heap
    w.put(0, sb)
    with open(path, "wb") as fh:
        fh.write(bytes(w.data))
