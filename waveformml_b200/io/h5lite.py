"""Minimal HDF5 reader / writer in numpy + zlib -- just enough for the reference's pulse files.

The reference reads its events with h5py (src/utils/HDF5Utils.py, src/datasets/HDF5Dataset.py:377-403,
src/datasets/HDF5IO.py:40-79) and writes predictions with `create_dataset(..., compression="gzip", chunks=(1024,),
maxshape=(None,))` (src/datasets/HDF5IO.py:88-97).  Neither h5py nor libhdf5 exists in this image, so this module
implements the subset of the HDF5 file format (HDF5 File Format Specification version 2.0/3.0) those files use:

  reader  superblock 0/1 (and 2/3), user block / base address, version-1 and version-2 object headers with
          continuation blocks, old-style groups (symbol table: B-tree v1 + SNOD + local heap) and compact new-style
          groups (Link messages), dataspace v1/v2, datatypes: fixed-point, floating-point, string, array, compound
          (encodings 1-3), layouts: compact / contiguous / chunked through a version-1 B-tree (layout message v1-v3),
          filters: deflate, shuffle, fletcher32, attributes (message v1-v3).
  writer  superblock 0, old-style root group, datasets of simple or compound (with array members) element type,
          contiguous or chunked + gzip (+ shuffle), fixed shape, attributes -- what H5Output.create_table produces,
          minus resizing.

Pinned against a file written by the real HDF5 library: scipy ships a MATLAB v7.3 file (an HDF5 file with a 512-byte
user block) whose values are also available through scipy's own v5 reader -- tests/test_h5lite.py.  The compound /
chunked / gzip paths have no third-party-written file in this image: they are checked writer -> reader and against
hand-assembled bytes, and stay "parity unpinned" until a real pulse file is available.
"""
import struct
import zlib

import numpy as np

SIGNATURE = b"\x89HDF\r\n\x1a\n"
UNDEF = 0xFFFFFFFFFFFFFFFF


class H5Error(RuntimeError):
    pass


def _pad8(n):
    return (n + 7) & ~7


# ======================================================================================================= datatypes
def parse_datatype(buf, pos=0):
    """Datatype message -> (numpy dtype, bytes consumed).  buf: bytes / memoryview."""
    cv, b0, b1, b2, size = struct.unpack_from("<BBBBI", buf, pos)
    cls, ver = cv & 0x0F, cv >> 4
    p = pos + 8
    if cls == 0:  # fixed point: bit 0 byte order, bit 3 signed; properties: bit offset, precision
        order = ">" if (b0 & 1) else "<"
        signed = bool(b0 & 8)
        return np.dtype("%s%s%d" % (order, "i" if signed else "u", size)), p + 4 - pos
    if cls == 1:  # floating point (IEEE layouts only)
        order = ">" if (b0 & 1) else "<"
        if size not in (2, 4, 8):
            raise H5Error("floating-point type of %d bytes" % size)
        return np.dtype("%sf%d" % (order, size)), p + 12 - pos
    if cls == 3:  # fixed-length string
        return np.dtype("S%d" % size), p - pos
    if cls == 4:  # bit field
        return np.dtype("<u%d" % size), p + 4 - pos
    if cls == 6:  # compound
        nmemb = b0 | (b1 << 8)
        names, formats, offsets = [], [], []
        for _ in range(nmemb):
            end = bytes(buf[p:]).index(b"\0")
            name = bytes(buf[p:p + end]).decode()
            if ver < 3:
                p += _pad8(end + 1)
            else:
                p += end + 1
            if ver < 3:
                (off,) = struct.unpack_from("<I", buf, p)
                p += 4
            else:
                nb = 1 if size < 256 else (2 if size < 65536 else (3 if size < (1 << 24) else 4))
                off = int.from_bytes(bytes(buf[p:p + nb]), "little")
                p += nb
            dims = None
            if ver == 1:
                rank = buf[p]
                p += 4 + 4 + 4  # dimensionality + reserved(3), permutation, reserved
                d4 = struct.unpack_from("<4I", buf, p)
                p += 16
                if rank:
                    dims = tuple(int(x) for x in d4[:rank])
            mt, used = parse_datatype(buf, p)
            p += used
            if dims:
                mt = np.dtype((mt, dims))
            names.append(name)
            formats.append(mt)
            offsets.append(off)
        return np.dtype({"names": names, "formats": formats, "offsets": offsets, "itemsize": size}), p - pos
    if cls == 10:  # array
        rank = buf[p]
        p += 1 if ver >= 3 else 4
        dims = struct.unpack_from("<%dI" % rank, buf, p)
        p += 4 * rank
        if ver < 3:
            p += 4 * rank  # permutation indices
        base, used = parse_datatype(buf, p)
        p += used
        return np.dtype((base, tuple(int(d) for d in dims))), p - pos
    if cls == 8:  # enumeration: values of the base type
        base, used = parse_datatype(buf, p)
        nmemb = b0 | (b1 << 8)
        p += used
        for _ in range(nmemb):
            end = bytes(buf[p:]).index(b"\0")
            p += _pad8(end + 1) if ver < 3 else end + 1
        p += nmemb * base.itemsize
        return base, p - pos
    raise H5Error("datatype class %d is not supported" % cls)


def encode_datatype(dt):
    """numpy dtype -> datatype message bytes (fixed / float / string / compound v2 with array members v2)."""
    dt = np.dtype(dt)
    if dt.subdtype is not None:
        base, shape = dt.subdtype
        body = struct.pack("<B3x", len(shape)) + struct.pack("<%dI" % len(shape), *shape)
        body += struct.pack("<%dI" % len(shape), *range(len(shape)))  # permutation (unused)
        return struct.pack("<BBBBI", (2 << 4) | 10, 0, 0, 0, dt.itemsize) + body + encode_datatype(base)
    if dt.names is not None:
        out = struct.pack("<BBBBI", (2 << 4) | 6, len(dt.names) & 0xFF, len(dt.names) >> 8, 0, dt.itemsize)
        for name in dt.names:
            mt, off = dt.fields[name][0], dt.fields[name][1]
            nm = name.encode() + b"\0"
            out += nm + b"\0" * (_pad8(len(nm)) - len(nm)) + struct.pack("<I", off) + encode_datatype(mt)
        return out
    big = 1 if dt.byteorder == ">" else 0
    if dt.kind in "iu":
        b0 = big | (8 if dt.kind == "i" else 0)
        return struct.pack("<BBBBI", (1 << 4) | 0, b0, 0, 0, dt.itemsize) + struct.pack("<HH", 0, 8 * dt.itemsize)
    if dt.kind == "f":
        exp, man, bias = {2: (5, 10, 15), 4: (8, 23, 127), 8: (11, 52, 1023)}[dt.itemsize]
        # bits: byte order, padding 0, mantissa normalisation 2 (implied msb) in bits 4-5; sign location in byte 1
        return (struct.pack("<BBBBI", (1 << 4) | 1, big | 0x20, 8 * dt.itemsize - 1, 0, dt.itemsize) +
                struct.pack("<HHBBBBI", 0, 8 * dt.itemsize, man, exp, 0, man, bias))
    if dt.kind == "S":
        return struct.pack("<BBBBI", (1 << 4) | 3, 0, 0, 0, dt.itemsize)
    raise H5Error("cannot encode dtype %r" % dt)


# ======================================================================================================= reader
class Dataset:
    def __init__(self, f, name, dtype, shape, maxshape, layout, filters, attrs):
        self.file, self.name, self.dtype, self.shape, self.maxshape = f, name, dtype, shape, maxshape
        self._layout, self._filters, self.attrs = layout, filters, attrs
        self._chunk_index = None

    def __len__(self):
        return self.shape[0]

    len = __len__

    @property
    def chunks(self):
        return self._layout["chunk"][:-1] if self._layout["class"] == 2 else None

    def _unfilter(self, raw, mask):
        for i in range(len(self._filters) - 1, -1, -1):
            fid, cd = self._filters[i]
            if mask & (1 << i):
                continue
            if fid == 1:
                raw = zlib.decompress(raw)
            elif fid == 2:
                es = cd[0] if cd else self.dtype.itemsize
                a = np.frombuffer(raw, dtype=np.uint8)
                n = a.size // es
                raw = a[:n * es].reshape(es, n).T.tobytes() + a[n * es:].tobytes()
            elif fid == 3:
                raw = raw[:-4]
            else:
                raise H5Error("filter %d is not supported" % fid)
        return raw

    def _chunks_1d(self):
        """[(first row, address, stored bytes, filter mask)] sorted by row -- chunked layouts."""
        if self._chunk_index is None:
            out = []
            self.file._walk_chunk_btree(self._layout["btree"], len(self._layout["chunk"]), out)
            out.sort()
            self._chunk_index = out
        return self._chunk_index

    def read(self, start=0, stop=None):
        """rows [start, stop) along the first dimension as a numpy array of self.dtype."""
        n = self.shape[0] if self.shape else 1
        stop = n if stop is None else min(stop, n)
        start = max(0, min(start, stop))
        inner = tuple(self.shape[1:])
        per_row = int(np.prod(inner)) if inner else 1
        row_bytes = per_row * self.dtype.itemsize
        cls = self._layout["class"]
        if cls == 0:
            data = np.frombuffer(self._layout["data"], dtype=self.dtype, count=n * per_row)
            return data.reshape((n,) + inner)[start:stop].copy()
        if cls == 1:
            addr = self._layout["address"]
            if addr == UNDEF:
                return np.zeros((stop - start,) + inner, dtype=self.dtype)
            raw = self.file._read(addr + start * row_bytes, (stop - start) * row_bytes)
            return np.frombuffer(raw, dtype=self.dtype).reshape((stop - start,) + inner).copy()
        chunk = self._layout["chunk"]
        if len(chunk) - 1 != len(self.shape) or any(c != s for c, s in zip(chunk[1:-1], self.shape[1:])):
            raise H5Error("only datasets chunked along their first dimension are supported (chunk %r, shape %r)"
                          % (chunk, self.shape))
        crow = chunk[0]
        out = np.zeros((stop - start,) + inner, dtype=self.dtype)
        flat = out.reshape(stop - start, -1) if inner else out
        for row0, addr, nbytes, mask in self._chunks_1d():
            if row0 + crow <= start or row0 >= stop:
                continue
            raw = self._unfilter(self.file._read(addr, nbytes), mask)
            rows = np.frombuffer(raw, dtype=self.dtype, count=crow * per_row).reshape((crow,) + inner)
            lo, hi = max(start, row0), min(stop, row0 + crow)
            out[lo - start:hi - start] = rows[lo - row0:hi - row0]
        del flat
        return out

    def __getitem__(self, key):
        if isinstance(key, str):  # field of a compound table, h5py style: ds["coord"]
            return self.read()[key]
        if isinstance(key, tuple) and key == ():
            return self.read()
        if isinstance(key, slice):
            start, stop, step = key.indices(self.shape[0])
            a = self.read(start, stop)
            return a[::step] if step != 1 else a
        if isinstance(key, (int, np.integer)):
            k = int(key) + (self.shape[0] if key < 0 else 0)
            return self.read(k, k + 1)[0]
        raise TypeError("unsupported index %r" % (key,))


class Group:
    def __init__(self, f, name, links, attrs):
        self.file, self.name, self._links, self.attrs = f, name, links, attrs

    def keys(self):
        return list(self._links)

    def items(self):
        return [(k, self[k]) for k in self._links]

    def __contains__(self, k):
        return k in self._links

    def __getitem__(self, path):
        node = self
        for part in [p for p in path.split("/") if p]:
            if not isinstance(node, Group) or part not in node._links:
                raise KeyError(path)
            node = node.file._object(node._links[part], (node.name.rstrip("/") + "/" + part))
        return node


class File(Group):
    """h5py-like read-only view: f["WaveformPairs"], .attrs, .shape, .dtype, slicing, f.items()."""

    def __init__(self, path):
        self._fh = open(path, "rb")
        self.path = path
        self._cache = {}
        self._parse_superblock()
        root = self._object(self._root_addr, "/")
        Group.__init__(self, self, "/", root._links, root.attrs)

    def close(self):
        self._fh.close()

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()
        return False

    # ------------------------------------------------------------------ raw access
    def _read(self, addr, n):
        self._fh.seek(self._base + addr)
        b = self._fh.read(n)
        if len(b) != n:
            raise H5Error("short read at %d (+%d)" % (addr, n))
        return b

    def _parse_superblock(self):
        fh = self._fh
        off = 0
        while True:  # the superblock sits at 0, 512, 1024, 2048, ... (user block)
            fh.seek(off)
            sig = fh.read(8)
            if sig == SIGNATURE:
                break
            if len(sig) < 8:
                raise H5Error("%s: no HDF5 signature" % self.path)
            off = 512 if off == 0 else off * 2
        self._sb_off = off
        ver = fh.read(1)[0]
        if ver in (0, 1):
            hdr = fh.read(15 + (4 if ver == 1 else 0))
            self._so, self._sl = hdr[4], hdr[5]
            self._leaf_k, self._int_k = struct.unpack_from("<HH", hdr, 7)
            if self._so != 8 or self._sl != 8:
                raise H5Error("only 8-byte offsets / lengths are supported")
            base, _free, self._eof, _drv = struct.unpack("<4Q", fh.read(32))
            self._base = base
            ste = fh.read(40)
            self._root_addr = struct.unpack_from("<Q", ste, 8)[0]
        elif ver in (2, 3):
            so, sl, _flags = fh.read(3)
            if so != 8 or sl != 8:
                raise H5Error("only 8-byte offsets / lengths are supported")
            self._so, self._sl = so, sl
            base, _ext, self._eof, self._root_addr = struct.unpack("<4Q", fh.read(32))
            self._base = base
        else:
            raise H5Error("superblock version %d" % ver)

    # ------------------------------------------------------------------ object headers
    def _messages(self, addr):
        """[(type, flags, payload bytes)] of the object header at addr (v1 or v2, continuation blocks followed)."""
        head = self._read(addr, 16)
        msgs = []
        if head[:4] == b"OHDR":
            ver, flags = head[4], head[5]
            p = 6
            if flags & 0x20:
                p += 16  # four timestamps
            if flags & 0x10:
                p += 4   # max compact / min dense attributes
            szb = 1 << (flags & 3)
            head = self._read(addr, p + szb)
            chunk0 = int.from_bytes(head[p:p + szb], "little")
            p += szb
            blocks = [(addr + p, chunk0)]
            track = bool(flags & 0x04)
            while blocks:
                a, n = blocks.pop(0)
                buf = self._read(a, n)
                q = 0
                while q + 4 <= n:
                    mtype, msize, mflags = buf[q], struct.unpack_from("<H", buf, q + 1)[0], buf[q + 3]
                    q += 4 + (2 if track else 0)
                    body = buf[q:q + msize]
                    q += msize
                    if mtype == 0x10:
                        ca, cl = struct.unpack_from("<QQ", body, 0)
                        blocks.append((ca + 4, cl - 8))  # skip "OCHK", drop the checksum
                    elif mtype != 0:
                        msgs.append((mtype, mflags, body))
            return msgs
        ver, _r, nmsg, _ref, hsize = struct.unpack_from("<BBHII", head, 0)
        if ver != 1:
            raise H5Error("object header version %d at %d" % (ver, addr))
        blocks = [(addr + 16, hsize)]
        while blocks and len(msgs) < nmsg + 64:
            a, n = blocks.pop(0)
            buf = self._read(a, n)
            q = 0
            while q + 8 <= n:
                mtype, msize, mflags = struct.unpack_from("<HHB", buf, q)
                q += 8
                body = buf[q:q + msize]
                q += msize
                if mtype == 0x10:
                    ca, cl = struct.unpack_from("<QQ", body, 0)
                    blocks.append((ca, cl))
                elif mtype != 0:
                    msgs.append((mtype, mflags, body))
        return msgs

    def _parse_attribute(self, body):
        ver = body[0]
        nsz, tsz, ssz = struct.unpack_from("<HHH", body, 2)
        p = 8 + (1 if ver == 3 else 0)
        pad = _pad8 if ver == 1 else (lambda x: x)
        name = bytes(body[p:p + nsz]).split(b"\0")[0].decode()
        p += pad(nsz)
        dt, _ = parse_datatype(body, p)
        p += pad(tsz)
        shape, _ = self._parse_dataspace(body[p:p + ssz])
        p += pad(ssz)
        count = int(np.prod(shape)) if shape else 1
        val = np.frombuffer(bytes(body[p:p + count * dt.itemsize]), dtype=dt, count=count).reshape(shape).copy()
        return name, val

    @staticmethod
    def _parse_dataspace(body):
        ver, rank, flags = body[0], body[1], body[2]
        p = 8 if ver == 1 else 4
        dims = struct.unpack_from("<%dQ" % rank, body, p)
        p += 8 * rank
        maxd = dims
        if flags & 1:
            maxd = struct.unpack_from("<%dQ" % rank, body, p)
        return tuple(int(d) for d in dims), tuple(None if d == UNDEF else int(d) for d in maxd)

    def _object(self, addr, name):
        if addr in self._cache:
            return self._cache[addr]
        msgs = self._messages(addr)
        attrs, links = {}, None
        dtype = shape = maxshape = layout = None
        filters = []
        for mtype, _mflags, body in msgs:
            if mtype == 0x0C:
                k, v = self._parse_attribute(body)
                attrs[k] = v
            elif mtype == 0x11:  # symbol table: B-tree + local heap
                btree, heap = struct.unpack_from("<QQ", body, 0)
                links = {} if links is None else links
                self._walk_group_btree(btree, self._heap_data(heap), links)
            elif mtype == 0x06:  # link message (compact new-style group)
                links = {} if links is None else links
                k, a = self._parse_link(body)
                if a is not None:
                    links[k] = a
            elif mtype == 0x02:
                links = {} if links is None else links
                fheap = struct.unpack_from("<Q", body, 2 + (8 if body[1] & 1 else 0))[0]
                if fheap != UNDEF:
                    raise H5Error("%s: densely stored links (fractal heap) are not supported" % name)
            elif mtype == 0x01:
                shape, maxshape = self._parse_dataspace(body)
            elif mtype == 0x03:
                dtype, _ = parse_datatype(body, 0)
            elif mtype == 0x08:
                layout = self._parse_layout(body)
            elif mtype == 0x0B:
                filters = self._parse_filters(body)
        if links is not None and layout is None:
            obj = Group(self, name, links, attrs)
        elif layout is not None and dtype is not None and shape is not None:
            obj = Dataset(self, name, dtype, shape, maxshape, layout, filters, attrs)
        else:
            raise H5Error("%s: neither a group nor a dataset this reader understands" % name)
        self._cache[addr] = obj
        return obj

    @staticmethod
    def _parse_link(body):
        ver, flags = body[0], body[1]
        p = 2
        ltype = 0
        if flags & 0x08:
            ltype = body[p]
            p += 1
        if flags & 0x04:
            p += 8
        if flags & 0x10:
            p += 1
        nb = 1 << (flags & 3)
        ln = int.from_bytes(bytes(body[p:p + nb]), "little")
        p += nb
        name = bytes(body[p:p + ln]).decode()
        p += ln
        if ltype != 0:
            return name, None  # soft / external links are skipped
        return name, struct.unpack_from("<Q", body, p)[0]

    def _parse_layout(self, body):
        ver = body[0]
        if ver == 3:
            cls = body[1]
            if cls == 0:
                (n,) = struct.unpack_from("<H", body, 2)
                return {"class": 0, "data": bytes(body[4:4 + n])}
            if cls == 1:
                addr, size = struct.unpack_from("<QQ", body, 2)
                return {"class": 1, "address": addr, "size": size}
            if cls == 2:
                rank = body[2]
                (btree,) = struct.unpack_from("<Q", body, 3)
                dims = struct.unpack_from("<%dI" % rank, body, 11)
                return {"class": 2, "btree": btree, "chunk": tuple(int(d) for d in dims)}
            raise H5Error("layout class %d" % cls)
        if ver in (1, 2):
            rank, cls = body[1], body[2]
            p = 8
            addr = None
            if cls != 0:
                (addr,) = struct.unpack_from("<Q", body, p)
                p += 8
            dims = struct.unpack_from("<%dI" % rank, body, p)
            p += 4 * rank
            if cls == 0:
                (n,) = struct.unpack_from("<I", body, p)
                return {"class": 0, "data": bytes(body[p + 4:p + 4 + n])}
            if cls == 1:
                return {"class": 1, "address": addr, "size": None}
            return {"class": 2, "btree": addr, "chunk": tuple(int(d) for d in dims)}
        raise H5Error("data layout message version %d (chunk indexes other than the version-1 B-tree need libver "
                      "'latest' files, which the reference's writers do not produce)" % ver)

    @staticmethod
    def _parse_filters(body):
        ver, n = body[0], body[1]
        p = 8 if ver == 1 else 2
        out = []
        for _ in range(n):
            (fid,) = struct.unpack_from("<H", body, p)
            p += 2
            nlen = 0
            if ver == 1 or fid >= 256:
                (nlen,) = struct.unpack_from("<H", body, p)
                p += 2
            _flags, ncd = struct.unpack_from("<HH", body, p)
            p += 4
            p += _pad8(nlen) if ver == 1 else nlen
            cd = struct.unpack_from("<%dI" % ncd, body, p)
            p += 4 * ncd
            if ver == 1 and ncd % 2:
                p += 4
            out.append((fid, tuple(cd)))
        return out

    # ------------------------------------------------------------------ groups / chunk index
    def _heap_data(self, addr):
        h = self._read(addr, 32)
        if h[:4] != b"HEAP":
            raise H5Error("bad local heap at %d" % addr)
        size, _free, data = struct.unpack_from("<QQQ", h, 8)
        return self._read(data, size)

    def _walk_group_btree(self, addr, heap, links):
        h = self._read(addr, 24)
        if h[:4] == b"SNOD":
            nsym = struct.unpack_from("<H", h, 6)[0]
            buf = self._read(addr + 8, nsym * 40)
            for i in range(nsym):
                noff, oaddr = struct.unpack_from("<QQ", buf, i * 40)
                name = heap[noff:heap.index(b"\0", noff)].decode()
                links[name] = oaddr
            return
        if h[:4] != b"TREE":
            raise H5Error("bad group B-tree node at %d" % addr)
        _ntype, _level, used = struct.unpack_from("<BBH", h, 4)
        buf = self._read(addr + 24, (2 * used + 1) * 8)
        for i in range(used):
            self._walk_group_btree(struct.unpack_from("<Q", buf, (2 * i + 1) * 8)[0], heap, links)

    def _walk_chunk_btree(self, addr, ndims, out):
        if addr == UNDEF:
            return
        h = self._read(addr, 24)
        if h[:4] != b"TREE" or h[4] != 1:
            raise H5Error("bad chunk B-tree node at %d" % addr)
        level, used = h[5], struct.unpack_from("<H", h, 6)[0]
        ksz = 8 + 8 * ndims
        buf = self._read(addr + 24, used * (ksz + 8) + ksz)
        for i in range(used):
            q = i * (ksz + 8)
            nbytes, mask = struct.unpack_from("<II", buf, q)
            offs = struct.unpack_from("<%dQ" % ndims, buf, q + 8)
            (child,) = struct.unpack_from("<Q", buf, q + ksz)
            if level == 0:
                out.append((int(offs[0]), child, nbytes, mask))
            else:
                self._walk_chunk_btree(child, ndims, out)


# ======================================================================================================= writer
class Writer:
    """Writes a flat HDF5 file (datasets in the root group).  Usage:
        w = Writer(path); w.create_dataset("WaveformPairs", array, chunks=1024, gzip=9, attrs={"nevents": [n]}); w.close()
    """

    LEAF_K, INT_K = 32, 16  # symbol-table node holds up to 2*LEAF_K entries: one leaf is enough for a flat file

    def __init__(self, path):
        self._fh = open(path, "wb")
        self._fh.write(b"\0" * 96)  # superblock v0 (8 + 8 + 8 + 4 + 32 + 40 = 96 bytes, written at close)
        self._objects = []  # (name, object header address)

    def _tell(self):
        return self._fh.tell()

    def _align(self, n=8):
        pos = self._fh.tell()
        if pos % n:
            self._fh.write(b"\0" * (n - pos % n))
        return self._fh.tell()

    def _put(self, data):
        addr = self._align()
        self._fh.write(data)
        return addr

    @staticmethod
    def _msg(mtype, body, flags=0):
        body = body + b"\0" * (_pad8(len(body)) - len(body))
        return struct.pack("<HHB3x", mtype, len(body), flags) + body

    @staticmethod
    def _dataspace(shape, maxshape=None):
        rank = len(shape)
        body = struct.pack("<BBB5x", 1, rank, 1 if maxshape is not None else 0)
        body += struct.pack("<%dQ" % rank, *shape)
        if maxshape is not None:
            body += struct.pack("<%dQ" % rank, *[UNDEF if m is None else m for m in maxshape])
        return body

    def _attribute(self, name, value):
        v = np.ascontiguousarray(value)
        nm = name.encode() + b"\0"
        dt = encode_datatype(v.dtype)
        ds = self._dataspace(v.shape)
        body = struct.pack("<BxHHH", 1, len(nm), len(dt), len(ds))
        for part in (nm, dt, ds):
            body += part + b"\0" * (_pad8(len(part)) - len(part))
        return self._msg(0x0C, body + v.tobytes())

    def _object_header(self, msgs):
        blob = b"".join(msgs)
        return self._put(struct.pack("<BxHII4x", 1, len(msgs), 1, len(blob)) + blob)

    def create_dataset(self, name, data, chunks=None, gzip=None, shuffle=False, attrs=None, maxshape=None):
        data = np.ascontiguousarray(data)
        shape = data.shape
        inner = int(np.prod(shape[1:])) if len(shape) > 1 else 1
        row_bytes = inner * data.dtype.itemsize
        msgs = [self._msg(0x01, self._dataspace(shape, maxshape)), self._msg(0x03, encode_datatype(data.dtype), 1),
                self._msg(0x05, struct.pack("<BBBB", 2, 2, 2, 0))]  # fill value v2: allocate late, never written, undefined
        if chunks is None:
            assert gzip is None and not shuffle, "filters need a chunked layout"
            addr = self._put(data.tobytes()) if data.size else UNDEF
            msgs.append(self._msg(0x08, struct.pack("<BBQQ", 3, 1, addr, data.nbytes)))
        else:
            crow = int(chunks)
            filt = []
            if shuffle:
                filt.append((2, (data.dtype.itemsize,)))
            if gzip is not None:
                filt.append((1, (int(gzip),)))
            entries = []
            raw_all = data.reshape(shape[0], -1).view(np.uint8).reshape(shape[0], row_bytes) if data.size else None
            for r0 in range(0, shape[0], crow):
                block = np.zeros((crow, row_bytes), dtype=np.uint8)  # edge chunks are stored full size
                block[:min(crow, shape[0] - r0)] = raw_all[r0:r0 + crow]
                raw = block.tobytes()
                for fid, cd in filt:
                    if fid == 2:
                        es = cd[0]
                        raw = np.frombuffer(raw, dtype=np.uint8).reshape(-1, es).T.tobytes()
                    elif fid == 1:
                        raw = zlib.compress(raw, cd[0])
                entries.append((r0, self._put(raw), len(raw)))
            ndims = len(shape) + 1
            btree = self._chunk_btree(entries, ndims, shape, crow)
            cdims = (crow,) + tuple(shape[1:]) + (data.dtype.itemsize,)
            msgs.append(self._msg(0x08, struct.pack("<BBBQ", 3, 2, ndims, btree) + struct.pack("<%dI" % ndims, *cdims)))
            if filt:
                body = struct.pack("<BB6x", 1, len(filt))
                for fid, cd in filt:
                    body += struct.pack("<HHHH", fid, 0, 1, len(cd)) + struct.pack("<%dI" % len(cd), *cd)
                    if len(cd) % 2:
                        body += b"\0" * 4
                msgs.append(self._msg(0x0B, body))
        for k, v in (attrs or {}).items():
            msgs.append(self._attribute(k, v))
        self._objects.append((name, self._object_header(msgs)))

    def _chunk_btree(self, entries, ndims, shape, crow):
        """version-1 B-tree of node type 1; leaves of up to 2*INT_K chunks, one level of internal nodes above them as
        needed (recursively)."""
        if not entries:
            return UNDEF
        ksz = 8 + 8 * ndims
        cap = 2 * self.INT_K

        def key(nbytes, row0):
            return struct.pack("<II", nbytes, 0) + struct.pack("<%dQ" % ndims, row0, *([0] * (ndims - 1)))

        last_key_row = ((shape[0] + crow - 1) // crow) * crow
        level_nodes = []  # (first row, address, stored bytes of the first chunk)
        groups = [entries[i:i + cap] for i in range(0, len(entries), cap)]
        for gi, grp in enumerate(groups):
            body = b""
            for r0, addr, nb in grp:
                body += key(nb, r0) + struct.pack("<Q", addr)
            nxt = groups[gi + 1][0][0] if gi + 1 < len(groups) else last_key_row
            body += key(0, nxt)
            body += b"\0" * ((cap - len(grp)) * (ksz + 8))
            level_nodes.append([grp[0][0], None, grp[0][2], body, len(grp)])
        level = 0
        while True:
            # write this level; sibling pointers need the addresses, so reserve them first
            size = 24 + cap * (ksz + 8) + ksz
            addrs = []
            for _ in level_nodes:
                addrs.append(self._align())
                self._fh.write(b"\0" * size)
            end = self._tell()
            for i, (r0, _a, nb, body, used) in enumerate(level_nodes):
                left = addrs[i - 1] if i > 0 else UNDEF
                right = addrs[i + 1] if i + 1 < len(addrs) else UNDEF
                self._fh.seek(addrs[i])
                self._fh.write(b"TREE" + struct.pack("<BBHQQ", 1, level, used, left, right) + body)
                level_nodes[i][1] = addrs[i]
            self._fh.seek(end)
            if len(level_nodes) == 1:
                return level_nodes[0][1]
            upper = []
            groups = [level_nodes[i:i + cap] for i in range(0, len(level_nodes), cap)]
            for gi, grp in enumerate(groups):
                body = b""
                for r0, addr, nb, _b, _u in grp:
                    body += key(nb, r0) + struct.pack("<Q", addr)
                nxt = groups[gi + 1][0][0] if gi + 1 < len(groups) else last_key_row
                body += key(0, nxt)
                body += b"\0" * ((cap - len(grp)) * (ksz + 8))
                upper.append([grp[0][0], None, grp[0][2], body, len(grp)])
            level_nodes = upper
            level += 1

    def close(self):
        objs = sorted(self._objects)  # symbol-table entries are ordered by name
        if len(objs) > 2 * self.LEAF_K:
            raise H5Error("flat writer: at most %d datasets" % (2 * self.LEAF_K))
        heap = b"\0" * 8  # offset 0: the empty name (root / first B-tree key)
        offs = []
        for name, _ in objs:
            offs.append(len(heap))
            nm = name.encode() + b"\0"
            heap += nm + b"\0" * (_pad8(len(nm)) - len(nm))
        free_off = len(heap)
        heap += struct.pack("<QQ", 1, 16) if False else b""
        heap_size = _pad8(len(heap)) + 16
        heap_data = heap + b"\0" * (heap_size - len(heap) - 16) + struct.pack("<QQ", 1, 16)  # one free block at the end
        free_off = heap_size - 16
        heap_data_addr = self._put(heap_data)
        heap_addr = self._put(b"HEAP" + struct.pack("<B3xQQQ", 0, heap_size, free_off, heap_data_addr))
        snod = b"SNOD" + struct.pack("<BxH", 1, len(objs))
        for (name, oaddr), noff in zip(objs, offs):
            snod += struct.pack("<QQII16x", noff, oaddr, 0, 0)
        snod += b"\0" * ((2 * self.LEAF_K - len(objs)) * 40)
        snod_addr = self._put(snod)
        keys_children = struct.pack("<Q", 0) + struct.pack("<Q", snod_addr) + struct.pack("<Q", offs[-1] if offs else 0)
        keys_children += b"\0" * ((2 * self.INT_K - 1) * 16)
        btree_addr = self._put(b"TREE" + struct.pack("<BBHQQ", 0, 0, 1 if objs else 0, UNDEF, UNDEF) + keys_children)
        root_addr = self._object_header([self._msg(0x11, struct.pack("<QQ", btree_addr, heap_addr))])
        eof = self._align()
        sb = SIGNATURE + struct.pack("<BBBBBBBBHHI", 0, 0, 0, 0, 0, 8, 8, 0, self.LEAF_K, self.INT_K, 0)
        sb += struct.pack("<QQQQ", 0, UNDEF, eof, UNDEF)
        sb += struct.pack("<QQII", 0, root_addr, 1, 0) + struct.pack("<QQ", btree_addr, heap_addr)
        assert len(sb) == 96
        self._fh.seek(0)
        self._fh.write(sb)
        self._fh.close()
