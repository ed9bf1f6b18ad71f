"""Minimal pure-Python reader for the NetCDF-4/HDF5 ``composite.nc`` files FHMCAnalysis writes.

Why this exists: the reference loads ``composite.nc`` through ``netCDF4.Dataset``
(reference moments/histogram/one_dim/ntot/gc_hist.pyx:143-182; the file is produced by
``window.to_nc``, moments/win_patch/fhmc_patch.pyx:551-634).  Neither netCDF4, h5py nor libhdf5
exist in the build container or on the GPU box, so the loader falls back to this reader when
``netCDF4`` cannot be imported.

Supported subset (everything the shipped fixtures and example composites use):
superblock v0 and v2, version-2 object headers (``OHDR`` + ``OCHK`` continuation chunks), root
group links in dense storage (LinkInfo -> fractal heap ``FRHP`` -> ``FHIB``/``FHDB``) or compact
Link messages, contiguous or compact dataset layout, little/big-endian fixed-point and IEEE float
element types, fixed-length string / numeric scalar attributes.  Chunked or filtered datasets raise
``NotImplementedError`` (the reference never writes them for the variables the hot path reads).

``write_hdf5`` / ``write_composite`` (bottom of the file) write the same subset; their metadata checksums (lookup3) are
verified against the values stored in the reference's own fixtures.  No libhdf5 exists in this image, so acceptance of
the written files by libhdf5/netCDF-4 itself is untested here.
"""

import struct

import numpy as np

_SIG = b"\x89HDF\r\n\x1a\n"
_UNDEF = 0xFFFFFFFFFFFFFFFF


class HDF5FormatError(Exception):
    pass


class _Reader(object):
    def __init__(self, buf):
        self.b = buf
        self.O = 8  # size of offsets
        self.L = 8  # size of lengths

    def u(self, pos, n):
        return int.from_bytes(self.b[pos:pos + n], "little")


def _parse_datatype(b, pos):
    """Return (numpy dtype or ('S', n) / None, total element size)."""
    cls_ver = b[pos]
    cls = cls_ver & 0x0F
    bits0 = b[pos + 1]
    size = int.from_bytes(b[pos + 4:pos + 8], "little")
    order = ">" if (bits0 & 1) else "<"
    if cls == 0:  # fixed point
        signed = bool(bits0 & 0x08)
        return np.dtype("%s%s%d" % (order, "i" if signed else "u", size)), size
    if cls == 1:  # floating point
        return np.dtype("%sf%d" % (order, size)), size
    if cls == 3:  # fixed-length string
        return np.dtype("S%d" % size), size
    return None, size  # references, vlen, compound ...: not needed on the hot path


def _parse_dataspace(r, pos):
    b = r.b
    ver = b[pos]
    rank = b[pos + 1]
    flags = b[pos + 2]
    if ver == 1:
        p = pos + 8
    elif ver == 2:
        if b[pos + 3] == 2:  # null dataspace
            return None
        p = pos + 4
    else:
        raise HDF5FormatError("dataspace version %d" % ver)
    dims = tuple(r.u(p + i * r.L, r.L) for i in range(rank))
    return dims


class _Object(object):
    """Parsed object header: the messages we care about."""

    def __init__(self):
        self.shape = None
        self.dtype = None
        self.layout = None  # ('contiguous', addr, size) | ('compact', bytes)
        self.attrs = {}
        self.links = {}  # compact link messages
        self.linkinfo_heap = None


def _parse_link(r, pos):
    """Parse one Link message body at ``pos``; return (name, address or None, next_pos)."""
    b = r.b
    ver = b[pos]
    if ver != 1:
        return None, None, pos
    flags = b[pos + 1]
    p = pos + 2
    ltype = 0
    if flags & 0x08:
        ltype = b[p]
        p += 1
    if flags & 0x04:
        p += 8
    if flags & 0x10:
        p += 1
    nlen_size = 1 << (flags & 3)
    nlen = r.u(p, nlen_size)
    p += nlen_size
    name = bytes(b[p:p + nlen]).decode("utf-8", "replace")
    p += nlen
    addr = None
    if ltype == 0:
        addr = r.u(p, r.O)
        p += r.O
    elif ltype == 1:  # soft link
        n = r.u(p, 2)
        p += 2 + n
    else:  # external / user defined
        n = r.u(p, 2)
        p += 2 + n
    return name, addr, p


def _parse_attribute(r, pos, obj):
    b = r.b
    ver = b[pos]
    name_size = r.u(pos + 2, 2)
    dt_size = r.u(pos + 4, 2)
    ds_size = r.u(pos + 6, 2)
    if ver == 1:
        p = pos + 8

        def pad(n):
            return (n + 7) & ~7
    elif ver == 2:
        p = pos + 8

        def pad(n):
            return n
    elif ver == 3:
        p = pos + 9

        def pad(n):
            return n
    else:
        return
    name = bytes(b[p:p + name_size]).split(b"\x00")[0].decode("utf-8", "replace")
    p += pad(name_size)
    dtype, esize = _parse_datatype(b, p)
    p += pad(dt_size)
    dims = _parse_dataspace(r, p)
    p += pad(ds_size)
    if dtype is None or dims is None:
        return
    count = 1
    for d in dims:
        count *= d
    raw = bytes(b[p:p + count * esize])
    if dtype.kind == "S":
        val = raw.split(b"\x00")[0].decode("utf-8", "replace")
    else:
        arr = np.frombuffer(raw, dtype=dtype, count=count)
        val = arr[0].item() if count == 1 else arr.astype(dtype.newbyteorder("=")).reshape(dims)
    obj.attrs[name] = val


def _parse_messages(r, pos, end, hdr_flags, obj, todo):
    b = r.b
    extra = 2 if (hdr_flags & 0x04) else 0
    while pos + 4 + extra <= end:
        mtype = b[pos]
        msize = r.u(pos + 1, 2)
        body = pos + 4 + extra
        if body + msize > end:
            break
        if mtype == 0x01:
            obj.shape = _parse_dataspace(r, body)
        elif mtype == 0x03:
            obj.dtype, _ = _parse_datatype(b, body)
        elif mtype == 0x08:
            lver = b[body]
            lclass = b[body + 1]
            if lver not in (3, 4):
                raise HDF5FormatError("layout version %d unsupported" % lver)
            if lclass == 1:
                obj.layout = ("contiguous", r.u(body + 2, r.O), r.u(body + 2 + r.O, r.L))
            elif lclass == 0:
                n = r.u(body + 2, 2)
                obj.layout = ("compact", bytes(b[body + 4:body + 4 + n]))
            else:
                obj.layout = ("chunked",)
        elif mtype == 0x0C:
            _parse_attribute(r, body, obj)
        elif mtype == 0x06:
            name, addr, _ = _parse_link(r, body)
            if name is not None and addr is not None:
                obj.links[name] = addr
        elif mtype == 0x02:
            lflags = b[body + 1]
            p = body + 2 + (8 if (lflags & 1) else 0)
            heap = r.u(p, r.O)
            if heap != _UNDEF:
                obj.linkinfo_heap = heap
        elif mtype == 0x10:
            todo.append((r.u(body, r.O), r.u(body + r.O, r.L)))
        pos = body + msize


def _parse_object_header(r, addr):
    b = r.b
    if bytes(b[addr:addr + 4]) != b"OHDR":
        raise HDF5FormatError("only version-2 object headers are supported (no OHDR at %d)" % addr)
    if b[addr + 4] != 2:
        raise HDF5FormatError("object header version %d" % b[addr + 4])
    flags = b[addr + 5]
    p = addr + 6
    if flags & 0x20:
        p += 16
    if flags & 0x10:
        p += 4
    csize_bytes = 1 << (flags & 3)
    csize = r.u(p, csize_bytes)
    p += csize_bytes
    obj = _Object()
    todo = []
    _parse_messages(r, p, p + csize, flags, obj, todo)
    while todo:
        caddr, clen = todo.pop(0)
        if bytes(b[caddr:caddr + 4]) != b"OCHK":
            raise HDF5FormatError("bad continuation chunk at %d" % caddr)
        _parse_messages(r, caddr + 4, caddr + clen - 4, flags, obj, todo)
    return obj


def _fractal_heap_links(r, heap_addr):
    """Walk a fractal heap holding dense link storage; return {name: object header address}."""
    b = r.b
    if bytes(b[heap_addr:heap_addr + 4]) != b"FRHP":
        raise HDF5FormatError("no fractal heap at %d" % heap_addr)
    O, L = r.O, r.L
    p = heap_addr + 5
    p += 2  # heap id length
    filt_len = r.u(p, 2)
    p += 2
    hflags = b[p]
    p += 1
    p += 4  # max size of managed objects
    p += L + O + L + O  # next huge id, huge btree, free space, free-space manager
    p += 4 * L  # managed space, allocated, iterator offset, n managed objects
    p += 4 * L  # huge size, n huge, tiny size, n tiny
    width = r.u(p, 2)
    p += 2
    start_size = r.u(p, L)
    p += L
    max_direct = r.u(p, L)
    p += L
    max_heap_bits = r.u(p, 2)
    p += 2
    p += 2  # starting rows in root indirect block
    root = r.u(p, O)
    p += O
    cur_rows = r.u(p, 2)
    if filt_len:
        raise NotImplementedError("filtered fractal heaps are not supported")
    off_bytes = (max_heap_bits + 7) // 8
    has_cksum = bool(hflags & 0x02)
    links = {}

    def row_size(row):
        return start_size if row < 2 else start_size << (row - 1)

    def direct(addr, size):
        if bytes(b[addr:addr + 4]) != b"FHDB":
            raise HDF5FormatError("no direct block at %d" % addr)
        q = addr + 5 + O + off_bytes + (4 if has_cksum else 0)
        end = addr + size
        while q < end and b[q] == 1:
            name, a, nq = _parse_link(r, q)
            if name is None or nq <= q or nq > end:
                break
            if a is not None:
                links[name] = a
            q = nq

    def indirect(addr, nrows):
        if bytes(b[addr:addr + 4]) != b"FHIB":
            raise HDF5FormatError("no indirect block at %d" % addr)
        q = addr + 5 + O + off_bytes
        max_direct_rows = 2
        s = start_size
        while s < max_direct:
            s <<= 1
            max_direct_rows += 1
        for row in range(nrows):
            for _ in range(width):
                child = r.u(q, O)
                q += O
                if child == _UNDEF:
                    continue
                if row < max_direct_rows:
                    direct(child, row_size(row))
                else:
                    sub_rows = 1
                    t = row_size(row) // (start_size * width)
                    while t > 1:
                        t >>= 1
                        sub_rows += 1
                    indirect(child, sub_rows + 1)

    if root != _UNDEF:
        if cur_rows == 0:
            direct(root, start_size)
        else:
            indirect(root, cur_rows)
    return links


class Variable(object):
    """Lazy dataset handle; ``var[:]`` (or any numpy index) returns a native-endian ndarray."""

    def __init__(self, f, name, obj):
        self._f = f
        self.name = name
        self.shape = obj.shape if obj.shape is not None else ()
        self.dtype = obj.dtype
        self._layout = obj.layout
        self.attrs = obj.attrs

    def read(self):
        if self.dtype is None:
            raise NotImplementedError("dataset %r has an unsupported element type" % self.name)
        count = 1
        for d in self.shape:
            count *= d
        if self._layout is None:
            raise HDF5FormatError("dataset %r has no layout message" % self.name)
        kind = self._layout[0]
        if kind == "contiguous":
            addr = self._layout[1]
            if addr == _UNDEF:
                arr = np.zeros(count, dtype=self.dtype)
            else:
                addr += self._f._base
                arr = np.frombuffer(self._f._buf, dtype=self.dtype, count=count, offset=addr)
        elif kind == "compact":
            arr = np.frombuffer(self._layout[1], dtype=self.dtype, count=count)
        else:
            raise NotImplementedError("dataset %r uses chunked storage; only contiguous/compact "
                                      "layouts are supported by the built-in reader" % self.name)
        return np.array(arr.reshape(self.shape), dtype=self.dtype.newbyteorder("="))

    def __getitem__(self, idx):
        return self.read()[idx]

    def __len__(self):
        return self.shape[0]


class File(object):
    """Read-only view of the root group of an HDF5 file: ``.attrs`` and ``.variables``."""

    def __init__(self, fname):
        with open(fname, "rb") as fh:
            self._buf = fh.read()
        b = memoryview(self._buf)
        if bytes(b[:8]) != _SIG:
            raise HDF5FormatError("%s is not an HDF5 file" % fname)
        r = _Reader(b)
        ver = b[8]
        if ver in (0, 1):
            r.O = b[13]
            r.L = b[14]
            p = 24 + (4 if ver == 1 else 0)
            self._base = r.u(p, r.O)
            ste = p + 4 * r.O
            root_addr = r.u(ste + r.O, r.O)
        elif ver in (2, 3):
            r.O = b[9]
            r.L = b[10]
            self._base = r.u(12, r.O)
            root_addr = r.u(12 + 3 * r.O, r.O)
        else:
            raise HDF5FormatError("superblock version %d" % ver)
        if r.O != 8 or r.L != 8:
            raise HDF5FormatError("only 8-byte offsets/lengths are supported")
        root = _parse_object_header(r, root_addr + self._base)
        self.attrs = root.attrs
        links = dict(root.links)
        if root.linkinfo_heap is not None:
            links.update(_fractal_heap_links(r, root.linkinfo_heap + self._base))
        self.variables = {}
        for name, addr in links.items():
            try:
                obj = _parse_object_header(r, addr + self._base)
            except HDF5FormatError:
                continue
            if obj.layout is not None:
                self.variables[name] = Variable(self, name, obj)

    def close(self):
        self._buf = None


class Dataset(object):
    """Just enough of ``netCDF4.Dataset`` for ``histogram.reload`` (reference gc_hist.pyx:143-182):
    ``.variables[name][:]``, global attributes as Python attributes, ``.close()``."""

    def __init__(self, fname, mode="r", format="NETCDF4"):
        if mode != "r":
            raise NotImplementedError("built-in reader is read-only; use write_composite() to write")
        self._f = File(fname)
        self.variables = self._f.variables

    def __getattr__(self, name):
        f = self.__dict__.get("_f")
        if f is not None and name in f.attrs:
            return f.attrs[name]
        raise AttributeError(name)

    def ncattrs(self):
        return list(self._f.attrs.keys())

    def close(self):
        self._f.close()


# ---------------------------------------------------------------------------------------------------------------------
# Writer: the smallest HDF5 layout the reader above, libhdf5 >= 1.8 and netCDF-4 accept.
# Superblock v2, version-2 object headers, root links as compact Link messages, contiguous little-endian datasets,
# scalar/1-element attributes.  No dimension scales are written, so a netCDF-4 reader names the dimensions
# ``phony_dim_k``; ``histogram.reload`` (reference gc_hist.pyx:143-182) only indexes ``.variables[name][:]`` and reads
# the four global attributes, which is what this layout provides.
# ---------------------------------------------------------------------------------------------------------------------

def _rot(x, k):
    return ((x << k) | (x >> (32 - k))) & 0xFFFFFFFF


def lookup3(data, initval=0):
    """Bob Jenkins' lookup3 ``hashlittle`` — the metadata checksum of HDF5 v2 superblocks and object headers."""
    M = 0xFFFFFFFF
    n = len(data)
    a = b = c = (0xDEADBEEF + n + initval) & M
    p = 0
    while n > 12:
        a = (a + int.from_bytes(data[p:p + 4], "little")) & M
        b = (b + int.from_bytes(data[p + 4:p + 8], "little")) & M
        c = (c + int.from_bytes(data[p + 8:p + 12], "little")) & M
        a = (a - c) & M; a ^= _rot(c, 4); c = (c + b) & M
        b = (b - a) & M; b ^= _rot(a, 6); a = (a + c) & M
        c = (c - b) & M; c ^= _rot(b, 8); b = (b + a) & M
        a = (a - c) & M; a ^= _rot(c, 16); c = (c + b) & M
        b = (b - a) & M; b ^= _rot(a, 19); a = (a + c) & M
        c = (c - b) & M; c ^= _rot(b, 4); b = (b + a) & M
        p += 12
        n -= 12
    if n == 0:
        return c
    tail = bytes(data[p:p + n]) + b"\x00" * (12 - n)
    a = (a + int.from_bytes(tail[0:4], "little")) & M
    b = (b + int.from_bytes(tail[4:8], "little")) & M
    c = (c + int.from_bytes(tail[8:12], "little")) & M
    c ^= b; c = (c - _rot(b, 14)) & M
    a ^= c; a = (a - _rot(c, 11)) & M
    b ^= a; b = (b - _rot(a, 25)) & M
    c ^= b; c = (c - _rot(b, 16)) & M
    a ^= c; a = (a - _rot(c, 4)) & M
    b ^= a; b = (b - _rot(a, 14)) & M
    c ^= b; c = (c - _rot(b, 24)) & M
    return c


def _dtype_message(dt):
    dt = np.dtype(dt)
    if dt.kind == "f" and dt.itemsize == 8:
        return bytes([0x11, 0x20, 0x3F, 0x00]) + struct.pack("<I", 8) + struct.pack("<HHBBBBI", 0, 64, 52, 11, 0, 52, 1023)
    if dt.kind == "f" and dt.itemsize == 4:
        return bytes([0x11, 0x20, 0x1F, 0x00]) + struct.pack("<I", 4) + struct.pack("<HHBBBBI", 0, 32, 23, 8, 0, 23, 127)
    if dt.kind in "iu":
        return bytes([0x10, 0x08 if dt.kind == "i" else 0x00, 0, 0]) + struct.pack("<I", dt.itemsize) + \
            struct.pack("<HH", 0, 8 * dt.itemsize)
    if dt.kind == "S":
        return bytes([0x13, 0x00, 0x00, 0x00]) + struct.pack("<I", dt.itemsize)
    raise TypeError("cannot store dtype %s" % dt)


def _dataspace_message(shape):
    if shape == ():
        return bytes([2, 0, 0, 0])
    return bytes([2, len(shape), 0, 1]) + b"".join(struct.pack("<Q", int(d)) for d in shape)


def _message(mtype, body, flags=0):
    return bytes([mtype]) + struct.pack("<H", len(body)) + bytes([flags]) + body


def _attribute_message(name, value):
    if isinstance(value, (bytes, str)):
        raw = value.encode("utf-8") if isinstance(value, str) else value
        raw += b"\x00"
        dt, ds = _dtype_message(np.dtype("S%d" % len(raw))), _dataspace_message(())
    else:
        arr = np.atleast_1d(np.asarray(value))
        if arr.dtype.kind == "b":
            arr = arr.astype(np.int64)
        arr = np.ascontiguousarray(arr.astype(arr.dtype.newbyteorder("<")))
        raw, dt, ds = arr.tobytes(), _dtype_message(arr.dtype), _dataspace_message(arr.shape)
    nm = name.encode("utf-8") + b"\x00"
    body = bytes([3, 0]) + struct.pack("<HHH", len(nm), len(dt), len(ds)) + b"\x00" + nm + dt + ds + raw
    return _message(0x0C, body)


def _object_header(messages):
    body = b"".join(messages)
    head = b"OHDR" + bytes([2, 0x02]) + struct.pack("<I", len(body)) + body
    return head + struct.pack("<I", lookup3(head))


def write_hdf5(fname, variables, attrs=None, var_attrs=None):
    """Write ``variables`` (name -> ndarray) and global ``attrs`` (name -> str | number) to a new HDF5 file."""
    attrs = attrs or {}
    var_attrs = var_attrs or {}
    names = list(variables.keys())
    arrays = []
    for n in names:
        a = np.asarray(variables[n])
        if a.dtype.kind == "b":
            a = a.astype(np.int8)
        arrays.append(np.array(a.astype(a.dtype.newbyteorder("<")), order="C"))   # keeps 0-d arrays 0-d

    def dataset_header(a, addr, extra):
        msgs = [_message(0x01, _dataspace_message(a.shape)), _message(0x03, _dtype_message(a.dtype), 1),
                _message(0x05, bytes([3, 0x0A]), 1),
                _message(0x08, bytes([3, 1]) + struct.pack("<QQ", addr, a.nbytes), 0)]
        msgs += [_attribute_message(k, v) for k, v in extra.items()]
        return _object_header(msgs)

    def root_header(addrs):
        msgs = [_message(0x02, bytes([0, 0]) + struct.pack("<QQ", _UNDEF, _UNDEF)), _message(0x0A, bytes([0, 0]), 1)]
        for n, ad in zip(names, addrs):
            nm = n.encode("utf-8")
            if len(nm) > 255:
                raise ValueError("link name too long: %r" % n)
            msgs.append(_message(0x06, bytes([1, 0, len(nm)]) + nm + struct.pack("<Q", ad)))
        msgs += [_attribute_message(k, v) for k, v in attrs.items()]
        return _object_header(msgs)

    # header sizes do not depend on the addresses they hold: lay out with placeholders first
    root_len = len(root_header([0] * len(names)))
    hdr_len = [len(dataset_header(a, 0, var_attrs.get(n, {}))) for n, a in zip(names, arrays)]
    pos = 48 + root_len
    hdr_addr = []
    for ln in hdr_len:
        hdr_addr.append(pos)
        pos += ln
    data_addr = []
    for a in arrays:
        pos = (pos + 7) & ~7
        data_addr.append(pos if a.nbytes else _UNDEF)
        pos += a.nbytes
    eof = pos
    sb = _SIG + bytes([2, 8, 8, 0]) + struct.pack("<QQQQ", 0, _UNDEF, eof, 48)
    sb += struct.pack("<I", lookup3(sb))
    out = bytearray(eof)
    out[0:48] = sb
    rh = root_header(hdr_addr)
    out[48:48 + len(rh)] = rh
    for n, a, ha, da in zip(names, arrays, hdr_addr, data_addr):
        dh = dataset_header(a, da, var_attrs.get(n, {}))
        out[ha:ha + len(dh)] = dh
        if a.nbytes:
            out[da:da + a.nbytes] = a.tobytes()
    with open(fname, "wb") as fh:
        fh.write(out)


def write_composite(fname, lnpi, ntot, mom, volume, nspec, max_order, history="", histograms=None, op_name="N_{tot}"):
    """Write a ``composite.nc`` with the variable/attribute names of ``window.to_nc`` (reference
    moments/win_patch/fhmc_patch.pyx:551-634) so that ``histogram(fname, ...)`` — here or in the reference — loads it.
    ``histograms``: optional dict of the ``P_{N_i}(N_{tot})`` / ``P_{U}(N_{tot})`` families, written as given."""
    lnpi = np.asarray(lnpi, dtype=np.float64)
    ntot = np.asarray(ntot, dtype=np.int64)
    mom = np.asarray(mom, dtype=np.float64)
    n = len(lnpi)
    if ntot.shape != (n,) or mom.shape != (nspec, max_order + 1, nspec, max_order + 1, max_order + 1, n):
        raise ValueError("inconsistent composite shapes")
    v = {op_name: ntot, "ln(PI)": lnpi,
         "i": np.arange(1, nspec + 1, dtype=np.int64), "j": np.arange(max_order + 1, dtype=np.int64),
         "k": np.arange(1, nspec + 1, dtype=np.int64), "m": np.arange(max_order + 1, dtype=np.int64),
         "p": np.arange(max_order + 1, dtype=np.int64), "N_{i}^{j}*N_{k}^{m}*U^{p}": mom}
    for k, a in (histograms or {}).items():
        v[k] = np.asarray(a)
    write_hdf5(fname, v, {"history": history, "volume": float(volume), "nspec": int(nspec), "max_order": int(max_order)})
