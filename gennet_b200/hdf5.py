"""Minimal pure-Python HDF5 reader / writer for the Keras weight and model files of the reference
(``generator.h5``, ``signal_pe.h5``, ``best_*_weights.hdf5`` ...; bbhMahoGANy.py:1133-1142,1173,1373-1375).

h5py / libhdf5 are not available where this package runs, and Keras files use only a small, stable corner
of the format, which is what is implemented here (HDF5 File Format Specification, version 1.1 structures as
written by libhdf5 1.8/1.10 through h5py with default settings):

* superblock version 0, 8-byte offsets and lengths;
* "old style" groups: version-1 object header with a Symbol Table message -> version-1 B-tree ('TREE') of
  symbol-table nodes ('SNOD') whose link names live in a local heap ('HEAP');
* datasets with contiguous, compact or (unfiltered / deflate) chunked layout, fixed-point, IEEE float and
  fixed-length string element types;
* attributes (message 0x000C, versions 1-3) with scalar / simple dataspaces, including variable-length
  strings stored in global heap collections ('GCOL') -- h5py writes Python ``str`` attributes that way
  (``keras_version``, ``backend``, ``model_config``) and NumPy ``S`` arrays (``layer_names``,
  ``weight_names``) as fixed-length strings.

The writer emits the same structures (one contiguous dataset per array, fixed-length string attributes for
NumPy bytes arrays, variable-length strings for ``str``), so files written here are read by h5py/Keras and
vice versa.  Everything is little-endian.
"""
import struct
import zlib

import numpy as np

UNDEF = 0xFFFFFFFFFFFFFFFF
SIGNATURE = b'\x89HDF\r\n\x1a\n'


# ===================================================================================================== reader
class Dataset:
    def __init__(self, name, data, attrs):
        self.name, self._data, self.attrs = name, data, attrs

    @property
    def shape(self):
        return self._data.shape

    @property
    def dtype(self):
        return self._data.dtype

    def __getitem__(self, key):
        return self._data[key]

    def __array__(self, dtype=None, copy=None):
        return np.asarray(self._data, dtype=dtype)

    @property
    def value(self):
        return self._data


class Group:
    """dict-like view of a group: ``g['dense_1']['kernel:0']``, ``g.attrs['layer_names']``, ``g.keys()``."""

    def __init__(self, name, attrs=None):
        self.name = name
        self.attrs = attrs if attrs is not None else {}
        self._links = {}

    def keys(self):
        return list(self._links.keys())

    def items(self):
        return list(self._links.items())

    def __iter__(self):
        return iter(self._links)

    def __len__(self):
        return len(self._links)

    def __contains__(self, key):
        try:
            self[key]
            return True
        except KeyError:
            return False

    def __getitem__(self, path):
        node = self
        for part in [p for p in path.split('/') if p]:
            if not isinstance(node, Group) or part not in node._links:
                raise KeyError(path)
            node = node._links[part]
        return node

    def visit_datasets(self, prefix=''):
        for k, v in self._links.items():
            p = prefix + '/' + k if prefix else k
            if isinstance(v, Group):
                for item in v.visit_datasets(p):
                    yield item
            else:
                yield p, v


class _Reader:
    def __init__(self, buf):
        self.b = buf
        if buf[:8] != SIGNATURE:
            raise ValueError('not an HDF5 file (bad signature)')
        ver = buf[8]
        if ver not in (0, 1):
            raise ValueError('HDF5 superblock version %d is not supported (Keras/h5py files use version 0)' % ver)
        self.so, self.sl = buf[13], buf[14]
        if self.so != 8 or self.sl != 8:
            raise ValueError('only 8-byte offsets/lengths are supported')
        off = 24 if ver == 0 else 28
        self.base, _, self.eof, _ = struct.unpack_from('<4Q', buf, off)
        # root group symbol table entry
        self.root_entry = self._symbol_entry(off + 32)
        self._gcol = {}

    # ---- primitives
    def u(self, fmt, off):
        return struct.unpack_from('<' + fmt, self.b, off)

    def _symbol_entry(self, off):
        name_off, header, cache_type = self.u('QQI', off)
        scratch = self.b[off + 24:off + 40]
        return {'name_off': name_off, 'header': header, 'cache': cache_type, 'scratch': scratch}

    # ---- object headers
    def messages(self, addr):
        """(type, flags, body bytes) of a version-1 or version-2 object header, continuation blocks followed."""
        b = self.b
        out = []
        if b[addr:addr + 4] == b'OHDR':
            return self._messages_v2(addr)
        ver, _, nmsg, _refcnt, hsize = self.u('BBHII', addr)
        if ver != 1:
            raise ValueError('unsupported object header version %d at %d' % (ver, addr))
        blocks = [(addr + 16, hsize)]
        while blocks and len(out) < nmsg:
            pos, size = blocks.pop(0)
            end = pos + size
            while pos + 8 <= end and len(out) < nmsg:
                mtype, msize, mflags = self.u('HHB', pos)
                body = b[pos + 8:pos + 8 + msize]
                pos += 8 + msize
                if mtype == 0x0010:      # continuation
                    coff, clen = struct.unpack_from('<QQ', body, 0)
                    blocks.append((coff, clen))
                out.append((mtype, mflags, body))
        return out

    def _messages_v2(self, addr):
        b = self.b
        out = []
        flags = b[addr + 5]
        pos = addr + 6
        if flags & 0x20:
            pos += 16
        if flags & 0x10:
            pos += 4
        szbytes = 1 << (flags & 3)
        chunk0 = int.from_bytes(b[pos:pos + szbytes], 'little')
        pos += szbytes
        track = bool(flags & 0x04)
        blocks = [(pos, chunk0)]
        while blocks:
            pos, size = blocks.pop(0)
            end = pos + size
            while pos + 4 + (2 if track else 0) <= end:
                mtype = b[pos]
                msize, = self.u('H', pos + 1)
                mflags = b[pos + 3]
                pos += 4 + (2 if track else 0)
                body = b[pos:pos + msize]
                pos += msize
                if mtype == 0x10:
                    coff, clen = struct.unpack_from('<QQ', body, 0)
                    blocks.append((coff + 4, clen - 8))      # skip 'OCHK', drop checksum
                out.append((mtype, mflags, body))
        return out

    # ---- datatypes -> (numpy dtype | ('vlen_str',) | ('vlen', base), size)
    def datatype(self, body, off=0):
        cv, b0, b1, b2, size = struct.unpack_from('<BBBBI', body, off)
        cls, ver = cv & 0x0F, cv >> 4
        if cls == 0:       # fixed point
            signed = bool(b0 & 0x08)
            order = '>' if (b0 & 1) else '<'
            return np.dtype('%s%s%d' % (order, 'i' if signed else 'u', size)), size, 8 + 4
        if cls == 1:       # floating point
            order = '>' if (b0 & 1) else '<'
            return np.dtype('%sf%d' % (order, size)), size, 8 + 12
        if cls == 3:       # string
            return np.dtype('S%d' % size), size, 8
        if cls == 9:       # variable length
            vtype = b0 & 0x0F
            base, bsize, blen = self.datatype(body, off + 8)
            if vtype == 1:
                return ('vlen_str', (b1 & 0x0F)), size, 8 + blen
            return ('vlen', base), size, 8 + blen
        if cls == 6:       # compound (not used by Keras); report opaque bytes
            return np.dtype('V%d' % size), size, 8
        if cls == 8:       # enum (h5py bool): base type follows
            base, bsize, blen = self.datatype(body, off + 8)
            return base, size, 8 + blen
        raise ValueError('unsupported HDF5 datatype class %d' % cls)

    def dataspace(self, body):
        ver = body[0]
        rank = body[1]
        flags = body[2]
        if ver == 1:
            pos = 8
        elif ver == 2:
            if body[3] == 2:      # null dataspace
                return None
            pos = 4
        else:
            raise ValueError('unsupported dataspace version %d' % ver)
        dims = struct.unpack_from('<%dQ' % rank, body, pos) if rank else ()
        return tuple(int(d) for d in dims)

    def _global_heap_object(self, addr, index):
        if addr not in self._gcol:
            b = self.b
            if b[addr:addr + 4] != b'GCOL':
                raise ValueError('bad global heap collection at %d' % addr)
            csize, = self.u('Q', addr + 8)
            objs = {}
            pos = addr + 16
            end = addr + csize
            while pos + 16 <= end:
                idx, _ref, _res, osize = self.u('HHIQ', pos)
                if idx == 0:
                    break
                objs[idx] = b[pos + 16:pos + 16 + osize]
                pos += 16 + ((osize + 7) // 8) * 8
            self._gcol[addr] = objs
        return self._gcol[addr][index]

    def _decode(self, dt, shape, raw):
        n = 1
        for d in (shape or ()):
            n *= d
        if isinstance(dt, tuple) and dt[0] == 'vlen_str':
            vals = []
            for i in range(n):
                ln, addr, idx = struct.unpack_from('<IQI', raw, i * 16)
                s = self._global_heap_object(addr, idx)[:ln] if addr not in (0, UNDEF) and ln else b''
                vals.append(s.decode('utf-8', 'replace') if dt[1] == 1 else s.decode('latin1'))
            if shape == () or shape is None:
                return vals[0] if vals else ''
            return np.array(vals, dtype=object).reshape(shape)
        if isinstance(dt, tuple):
            raise ValueError('variable-length sequences are not supported')
        arr = np.frombuffer(raw, dtype=dt, count=n).reshape(shape or ())
        arr = arr.astype(dt.newbyteorder('=')) if dt.byteorder == '>' else arr.copy()
        if shape == ():
            return arr[()]
        return arr

    def attribute(self, body):
        ver = body[0]
        if ver == 1:
            nsz, dsz, ssz = struct.unpack_from('<HHH', body, 2)
            pos = 8
            pad = lambda x: (x + 7) // 8 * 8
        elif ver in (2, 3):
            nsz, dsz, ssz = struct.unpack_from('<HHH', body, 2)
            pos = 8 + (1 if ver == 3 else 0)
            pad = lambda x: x
        else:
            raise ValueError('unsupported attribute version %d' % ver)
        name = body[pos:pos + nsz].split(b'\x00')[0].decode('utf-8')
        pos += pad(nsz)
        dt, esize, _ = self.datatype(body, pos)
        pos += pad(dsz)
        shape = self.dataspace(body[pos:pos + ssz])
        pos += pad(ssz)
        if shape is None:
            return name, None
        n = 1
        for d in shape:
            n *= d
        raw = body[pos:pos + n * esize]
        return name, self._decode(dt, shape, raw)

    # ---- objects
    def read_object(self, addr, name):
        msgs = self.messages(self.base + addr)
        attrs = {}
        symtab = None
        dt = shape = layout = None
        filters = []
        for mtype, _fl, body in msgs:
            if mtype == 0x000C:
                k, v = self.attribute(body)
                attrs[k] = v
            elif mtype == 0x0011:
                symtab = struct.unpack_from('<QQ', body, 0)
            elif mtype == 0x0001:
                shape = self.dataspace(body)
            elif mtype == 0x0003:
                dt = self.datatype(body)
            elif mtype == 0x0008:
                layout = body
            elif mtype == 0x000B:
                filters = self._filters(body)
        if symtab is not None:
            g = Group(name, attrs)
            for lname, laddr in self._group_links(symtab[0], symtab[1]):
                g._links[lname] = self.read_object(laddr, (name.rstrip('/') + '/' + lname))
            return g
        if dt is None or layout is None:
            # e.g. a new-style group (link messages) -- not written by h5py defaults
            g = Group(name, attrs)
            for mtype, _fl, body in msgs:
                if mtype == 0x0006:
                    lname, laddr = self._link_message(body)
                    if laddr is not None:
                        g._links[lname] = self.read_object(laddr, (name.rstrip('/') + '/' + lname))
            return g
        return Dataset(name, self._read_data(dt, shape, layout, filters), attrs)

    def _link_message(self, body):
        ver, flags = body[0], body[1]
        pos = 2
        ltype = 0
        if flags & 0x08:
            ltype = body[pos]
            pos += 1
        if flags & 0x04:
            pos += 8
        if flags & 0x10:
            pos += 1
        lsz = 1 << (flags & 3)
        ln = int.from_bytes(body[pos:pos + lsz], 'little')
        pos += lsz
        lname = body[pos:pos + ln].decode('utf-8')
        pos += ln
        if ltype != 0:
            return lname, None
        addr, = struct.unpack_from('<Q', body, pos)
        return lname, addr

    def _filters(self, body):
        ver, n = body[0], body[1]
        pos = 8 if ver == 1 else 2
        out = []
        for _ in range(n):
            fid, = struct.unpack_from('<H', body, pos)
            if ver == 1 or fid >= 256:
                nlen, = struct.unpack_from('<H', body, pos + 2)
            else:
                nlen = 0
            fflags, ncv = struct.unpack_from('<HH', body, pos + (4 if (ver == 1 or fid >= 256) else 2))
            pos += 8 if (ver == 1 or fid >= 256) else 6
            pos += (nlen + 7) // 8 * 8 if ver == 1 else nlen
            pos += 4 * ncv
            if ver == 1 and ncv % 2:
                pos += 4
            out.append(fid)
        return out

    def _group_links(self, btree, heap):
        b = self.b
        heap += self.base
        if b[heap:heap + 4] != b'HEAP':
            raise ValueError('bad local heap at %d' % heap)
        data_addr, = self.u('Q', heap + 24)
        data_addr += self.base
        out = []

        def name_at(off):
            end = b.index(b'\x00', data_addr + off)
            return b[data_addr + off:end].decode('utf-8')

        def walk(addr):
            addr += self.base
            if b[addr:addr + 4] == b'TREE':
                ntype, level, used = self.u('BBH', addr + 4)
                pos = addr + 24
                # keys and children alternate: key0 child0 key1 child1 ... keyN
                for i in range(used):
                    child, = self.u('Q', pos + 8)
                    walk(child)
                    pos += 16
            elif b[addr:addr + 4] == b'SNOD':
                n, = self.u('H', addr + 6)
                for i in range(n):
                    e = self._symbol_entry(addr + 8 + 40 * i)
                    out.append((name_at(e['name_off']), e['header']))
            else:
                raise ValueError('bad group B-tree node at %d' % addr)
        if btree != UNDEF:
            walk(btree)
        return out

    def _read_data(self, dt, shape, layout, filters):
        dtype, esize, _ = dt
        if shape is None:
            return np.zeros((0,), dtype if not isinstance(dtype, tuple) else object)
        n = 1
        for d in shape:
            n *= d
        ver = layout[0]
        if ver == 3:
            cls = layout[1]
            if cls == 0:        # compact
                size, = struct.unpack_from('<H', layout, 2)
                raw = layout[4:4 + size]
            elif cls == 1:      # contiguous
                addr, size = struct.unpack_from('<QQ', layout, 2)
                raw = b'\x00' * (n * esize) if addr == UNDEF else self.b[self.base + addr:self.base + addr + n * esize]
            elif cls == 2:      # chunked
                rank = layout[2]
                addr, = struct.unpack_from('<Q', layout, 3)
                cdims = struct.unpack_from('<%dI' % rank, layout, 11)
                raw = self._read_chunked(addr, cdims[:-1], shape, esize, filters)
            else:
                raise ValueError('unsupported layout class %d' % cls)
        elif ver in (1, 2):
            rank, cls = layout[1], layout[2]
            pos = 8
            addr = UNDEF
            if cls != 0:
                addr, = struct.unpack_from('<Q', layout, pos)
                pos += 8
            dims = struct.unpack_from('<%dI' % rank, layout, pos)
            pos += 4 * rank
            if cls == 1:
                raw = b'\x00' * (n * esize) if addr == UNDEF else self.b[self.base + addr:self.base + addr + n * esize]
            elif cls == 2:
                raw = self._read_chunked(addr, dims[:-1], shape, esize, filters)
            else:
                size, = struct.unpack_from('<I', layout, pos)
                raw = layout[pos + 4:pos + 4 + size]
        else:
            raise ValueError('unsupported data layout version %d' % ver)
        return self._decode(dtype, shape, raw)

    def _read_chunked(self, btree, cdims, shape, esize, filters):
        if any(f not in (1, 2) for f in filters):      # deflate, shuffle
            raise ValueError('unsupported HDF5 filter pipeline %s' % (filters,))
        rank = len(shape)
        out = np.zeros(shape, dtype='V%d' % esize)
        b = self.b
        csize = esize
        for c in cdims:
            csize *= c

        def walk(addr):
            addr += self.base
            if b[addr:addr + 4] != b'TREE':
                raise ValueError('bad chunk B-tree node')
            ntype, level, used = self.u('BBH', addr + 4)
            pos = addr + 24
            ksz = 8 + 8 * (rank + 1)
            for i in range(used):
                nbytes, fmask = self.u('II', pos)
                offs = self.u('%dQ' % (rank + 1), pos + 8)[:rank]
                child, = self.u('Q', pos + ksz)
                if level > 0:
                    walk(child)
                else:
                    raw = b[self.base + child:self.base + child + nbytes]
                    for f in reversed(filters):
                        if f == 1:
                            raw = zlib.decompress(raw)
                        elif f == 2:
                            a = np.frombuffer(raw, np.uint8).reshape(esize, -1)
                            raw = a.T.tobytes()
                    chunk = np.frombuffer(raw[:csize], dtype='V%d' % esize).reshape(cdims)
                    sl = tuple(slice(o, min(o + c, s)) for o, c, s in zip(offs, cdims, shape))
                    out[sl] = chunk[tuple(slice(0, s.stop - s.start) for s in sl)]
                pos += ksz + 8
        if btree != UNDEF:
            walk(btree)
        return out.tobytes()


class File(Group):
    """``hdf5.File(path)`` (read) / ``hdf5.File(path, 'w')`` then ``create_group`` / ``create_dataset`` / ``attrs``."""

    def __init__(self, path, mode='r'):
        Group.__init__(self, '/')
        self.path, self.mode = path, mode
        if mode == 'r':
            with open(path, 'rb') as f:
                rd = _Reader(f.read())
            root = rd.read_object(rd.root_entry['header'], '/')
            self.attrs, self._links = root.attrs, root._links
        elif mode == 'w':
            self._w = WGroup()
        else:
            raise ValueError("mode must be 'r' or 'w'")

    # write API (delegates to the tree that is serialised at close())
    def create_group(self, name):
        return self._w.create_group(name)

    def create_dataset(self, name, data):
        return self._w.create_dataset(name, data)

    @property
    def wattrs(self):
        return self._w.attrs

    def close(self):
        if self.mode == 'w':
            with open(self.path, 'wb') as f:
                f.write(_Writer().serialise(self._w))

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()


# ===================================================================================================== writer
class WGroup:
    def __init__(self):
        self.attrs = {}
        self.links = {}      # insertion order kept; serialised sorted by name (B-tree key order)

    def create_group(self, name):
        node = self
        for part in [p for p in name.split('/') if p]:
            if part not in node.links:
                node.links[part] = WGroup()
            node = node.links[part]
        return node

    def create_dataset(self, name, data):
        parts = [p for p in name.split('/') if p]
        node = self.create_group('/'.join(parts[:-1])) if len(parts) > 1 else self
        arr = np.asarray(data)
        ds = WDataset(arr if arr.ndim == 0 or arr.flags['C_CONTIGUOUS'] else np.ascontiguousarray(arr))
        node.links[parts[-1]] = ds
        return ds


class WDataset:
    def __init__(self, data):
        self.data = data
        self.attrs = {}


def _pad8(b):
    return b + b'\x00' * (-len(b) % 8)


class _Writer:
    """Serialises a WGroup tree.  Layout: superblock (96 bytes), then objects appended as they are
    emitted; every structure is 8-byte aligned.  Group B-trees have a single leaf level: symbol-table nodes hold
    up to 2*leafK entries, and one B-tree node up to 2*internalK children, so a group may have up to
    2*LEAF_K * 2*INTERNAL_K = 64 * 64 = 4096 links with the constants below (Keras models have tens)."""
    LEAF_K = 32
    INTERNAL_K = 32

    def __init__(self):
        self.buf = bytearray()
        self.gheap = []          # pending variable-length strings of the collection being built

    def alloc(self, data):
        off = len(self.buf)
        self.buf += _pad8(bytes(data))
        return off

    # ---- datatype / dataspace messages
    @staticmethod
    def dtype_msg(dt):
        dt = np.dtype(dt)
        if dt.kind == 'f':
            size = dt.itemsize
            if size == 4:
                props = struct.pack('<HHBBBBI', 0, 32, 23, 8, 0, 23, 127)
                sign = 31
            elif size == 8:
                props = struct.pack('<HHBBBBI', 0, 64, 52, 11, 0, 52, 1023)
                sign = 63
            elif size == 2:
                props = struct.pack('<HHBBBBI', 0, 16, 10, 5, 0, 10, 15)
                sign = 15
            else:
                raise ValueError('unsupported float size %d' % size)
            # class 1, version 1; bit field: little-endian, mantissa normalisation = implied (2 << 4), sign location
            return struct.pack('<BBBBI', 0x11, 0x20, sign, 0, size) + props
        if dt.kind in 'iu':
            b0 = 0x08 if dt.kind == 'i' else 0
            return struct.pack('<BBBBI', 0x10, b0, 0, 0, dt.itemsize) + struct.pack('<HH', 0, dt.itemsize * 8)
        if dt.kind == 'b':
            return struct.pack('<BBBBI', 0x10, 0, 0, 0, 1) + struct.pack('<HH', 0, 8)
        if dt.kind == 'S':
            # class 3, version 1; null-padded (1), ASCII (0)
            return struct.pack('<BBBBI', 0x13, 0x01, 0, 0, max(dt.itemsize, 1))
        raise ValueError('unsupported dtype %s' % dt)

    @staticmethod
    def vlen_str_msg(utf8=True):
        # class 9 version 1: type = string (1), padding null-terminate (0), charset UTF-8 (1) / ASCII (0); size 16;
        # base type = 1-byte unsigned fixed point, exactly what libhdf5 records for H5T_C_S1 variable-length
        # strings (h5py writes `str` as UTF-8 and `bytes` as ASCII variable-length strings)
        base = struct.pack('<BBBBI', 0x10, 0x00, 0, 0, 1) + struct.pack('<HH', 0, 8)
        return struct.pack('<BBBBI', 0x19, 0x01, 0x01 if utf8 else 0x00, 0, 16) + base

    @staticmethod
    def space_msg(shape):
        if shape == ():
            return struct.pack('<BBBB4x', 1, 0, 0, 0)
        return struct.pack('<BBBB4x', 1, len(shape), 1, 0) + b''.join(struct.pack('<Q', d) for d in shape) + \
            b''.join(struct.pack('<Q', d) for d in shape)

    # ---- global heap for variable-length strings (one collection per string keeps the bookkeeping trivial)
    def gcol(self, payload):
        size = 16 + 16 + len(_pad8(payload)) + 16
        size = max(size, 4096)
        body = bytearray()
        body += b'GCOL' + struct.pack('<B3xQ', 1, size)
        body += struct.pack('<HHIQ', 1, 1, 0, len(payload)) + _pad8(payload)
        free = size - len(body)
        body += struct.pack('<HHIQ', 0, 0, 0, free) + b'\x00' * (free - 16)
        return self.alloc(body)

    def attr_msg(self, name, value):
        nm = name.encode('utf-8') + b'\x00'
        if isinstance(value, (str, bytes)):
            payload = value.encode('utf-8') if isinstance(value, str) else bytes(value)
            addr = self.gcol(payload)
            dt = self.vlen_str_msg(utf8=isinstance(value, str))
            sp = self.space_msg(())
            data = struct.pack('<IQI', len(payload), addr, 1)
        else:
            arr = np.asarray(value)
            if arr.dtype.kind == 'U':
                arr = np.char.encode(arr, 'utf-8')
            if arr.dtype.kind == 'O':
                arr = np.array([x if isinstance(x, bytes) else str(x).encode('utf-8') for x in arr.ravel()]).reshape(arr.shape)
            dt = self.dtype_msg(arr.dtype)
            sp = self.space_msg(arr.shape)
            data = np.ascontiguousarray(arr).tobytes()
        body = struct.pack('<BxHHH', 1, len(nm), len(dt), len(sp)) + _pad8(nm) + _pad8(dt) + _pad8(sp) + data
        return (0x000C, 0x04, body)      # "do not share", as libhdf5 flags attribute messages

    def object_header(self, msgs):
        body = bytearray()
        for mtype, flags, mb in msgs:
            mb = _pad8(mb)
            body += struct.pack('<HHB3x', mtype, len(mb), flags) + mb
        hdr = struct.pack('<BxHII4x', 1, len(msgs), 1, len(body))
        return self.alloc(hdr + body)

    # ---- objects
    def write_dataset(self, ds):
        arr = ds.data
        if arr.dtype.kind == 'U':
            arr = np.char.encode(arr, 'utf-8')
        raw = arr.tobytes()
        addr = self.alloc(raw) if len(raw) else UNDEF
        msgs = [(0x0001, 0, self.space_msg(arr.shape)),
                (0x0003, 1, self.dtype_msg(arr.dtype)),
                (0x0005, 1, struct.pack('<BBBBI', 2, 2, 2, 1, 0)),      # fill value v2: late alloc, write if set, default
                (0x0008, 1, struct.pack('<BBQQ', 3, 1, addr, len(raw)))]
        msgs += [self.attr_msg(k, v) for k, v in ds.attrs.items()]
        return self.object_header(msgs)

    def write_group(self, g):
        # children first
        entries = []
        for name in sorted(g.links):
            child = g.links[name]
            if isinstance(child, WGroup):
                addr, bt, hp = self.write_group(child)
                entries.append((name, addr, 1, struct.pack('<QQ', bt, hp)))
            else:
                entries.append((name, self.write_dataset(child), 0, b'\x00' * 16))
        # local heap: offset 0 holds the empty string
        heap_data = bytearray(b'\x00' * 8)
        name_offs = []
        for name, *_ in entries:
            name_offs.append(len(heap_data))
            heap_data += _pad8(name.encode('utf-8') + b'\x00')
        free_off = len(heap_data)
        heap_data += struct.pack('<QQ', 1, 16)          # free block: next = 1 (none), size 16
        data_addr = self.alloc(heap_data)
        heap_addr = self.alloc(b'HEAP' + struct.pack('<B3xQQQ', 0, len(heap_data), free_off, data_addr))
        # symbol table nodes
        cap = 2 * self.LEAF_K
        if len(entries) > cap * 2 * self.INTERNAL_K:
            raise ValueError('too many links in one group')
        snods = []
        for i in range(0, max(len(entries), 1), cap):
            part = entries[i:i + cap]
            node = bytearray(b'SNOD' + struct.pack('<BxH', 1, len(part)))
            for j, (name, addr, cache, scratch) in enumerate(part):
                node += struct.pack('<QQI4x', name_offs[i + j], addr, cache) + scratch
            node += b'\x00' * (40 * (cap - len(part)))
            last_name_off = name_offs[i + len(part) - 1] if part else 0
            snods.append((self.alloc(node), last_name_off))
        # one leaf-level B-tree node: key0 (=0: empty string) child0 key1 ... ; key i+1 = heap offset of the largest
        # name in child i
        bt = bytearray(b'TREE' + struct.pack('<BBHQQ', 0, 0, len(snods), UNDEF, UNDEF))
        bt += struct.pack('<Q', 0)
        for addr, last in snods:
            bt += struct.pack('<QQ', addr, last)
        bt += b'\x00' * (16 * (2 * self.INTERNAL_K - len(snods)))
        bt_addr = self.alloc(bt)
        msgs = [(0x0011, 0, struct.pack('<QQ', bt_addr, heap_addr))]
        msgs += [self.attr_msg(k, v) for k, v in g.attrs.items()]
        return self.object_header(msgs), bt_addr, heap_addr

    def serialise(self, root):
        self.buf = bytearray(b'\x00' * 96)
        root_addr, bt, hp = self.write_group(root)
        eof = len(self.buf)
        sb = bytearray()
        sb += SIGNATURE
        sb += struct.pack('<BBBBBBBB', 0, 0, 0, 0, 0, 8, 8, 0)
        sb += struct.pack('<HHI', self.LEAF_K, self.INTERNAL_K, 0)
        sb += struct.pack('<QQQQ', 0, UNDEF, eof, UNDEF)
        sb += struct.pack('<QQI4x', 0, root_addr, 1) + struct.pack('<QQ', bt, hp)
        assert len(sb) == 96
        self.buf[:96] = sb
        return bytes(self.buf)
