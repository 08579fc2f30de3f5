"""A minimal HDF5 reader / writer for the files on either side of the path, for machines without h5py:
    word_weights_*/weights.hdf5      class_weights, class_biases, ...          (vlmap_memft/export_word_weights.py:60-73)
    the feature bank                 image_features, spatial_features, normal_boxes, num_boxes, data_info/{...}
                                                                               (generator_bottomup_vqa_tf_record_memft.py:95-126)
    tf_record_dir/data_info.hdf5     data_info/num_answers                     (input_ops_vqa_tf_record_memft.py:13-15)

Scope = what h5py 2.x / HDF5 1.8 writes for such files with default settings (restated from the HDF5 File Format
Specification, version 2.0): superblock version 0 / 1, version-1 object headers (with continuation blocks), groups stored
as symbol tables (v1 B-tree + local heap + SNOD nodes), dataspace v1 / v2, fixed-point / floating-point / fixed-length
string datatypes, data layout v3 compact / contiguous / chunked (v1 chunk B-tree) with the deflate and shuffle filters.
Anything else (superblock 2 / 3 with 'OHDR' headers, variable-length types, compound types, other filters) raises
NotImplementedError naming what was met. No h5py / libhdf5 exists in the build image, so no file written by the real
library could be cross-checked: tests pin the structure signatures and field layouts byte by byte against the
specification and the writer / reader against each other.
"""
import struct
import zlib

import numpy as np

SIGNATURE = b"\x89HDF\r\n\x1a\n"
UNDEF = 0xFFFFFFFFFFFFFFFF


class Hdf5Error(ValueError):
    pass


# ---------------------------------------------------------------------------------------------------------------------
# reader
# ---------------------------------------------------------------------------------------------------------------------
class File:
    """Read-only view: f['name'], f['group/name'] -> ndarray; 'name' in f; f.keys(group='')."""

    def __init__(self, path):
        with open(path, "rb") as fh:
            self.buf = fh.read()
        b = self.buf
        if b[:8] != SIGNATURE:
            raise Hdf5Error(f"{path}: not an HDF5 file (or it starts with a user block)")
        ver = b[8]
        if ver > 1:
            raise NotImplementedError(f"{path}: superblock version {ver} (libver='latest' files use version-2 object headers)")
        if b[13] != 8 or b[14] != 8:
            raise NotImplementedError(f"{path}: offsets / lengths of {b[13]} / {b[14]} bytes (only 8 / 8 is handled)")
        pos = 24 if ver == 0 else 28            # v1 adds indexed-storage K (2) + reserved (2) before the flags
        self.base = self._u64(pos)
        ste = pos + 32                          # base, free-space, end-of-file, driver-info addresses
        self.root = self._u64(ste + 8)          # root group symbol table entry: link name offset, object header address
        self._groups = {}

    # -- primitives
    def _u16(self, p):
        return struct.unpack_from("<H", self.buf, p)[0]

    def _u32(self, p):
        return struct.unpack_from("<I", self.buf, p)[0]

    def _u64(self, p):
        return struct.unpack_from("<Q", self.buf, p)[0]

    # -- object headers (version 1)
    def _messages(self, addr):
        b = self.buf
        addr += self.base
        if b[addr] != 1:
            if b[addr:addr + 4] == b"OHDR":
                raise NotImplementedError("version-2 object header (file written with libver='latest')")
            raise Hdf5Error(f"object header version {b[addr]} at {addr}")
        nmsgs, size = self._u16(addr + 2), self._u32(addr + 8)
        blocks = [(addr + 16, size)]            # 12-byte prefix padded to 8 bytes
        out = []
        while blocks and len(out) < nmsgs:
            pos, left = blocks.pop(0)
            end = pos + left
            while pos + 8 <= end and len(out) < nmsgs:
                mtype, msize = self._u16(pos), self._u16(pos + 2)
                data = b[pos + 8:pos + 8 + msize]
                pos += 8 + msize
                if mtype == 0x10:               # continuation: offset, length
                    blocks.append((struct.unpack_from("<Q", data, 0)[0] + self.base, struct.unpack_from("<Q", data, 8)[0]))
                out.append((mtype, data))
        return out

    # -- groups (symbol tables)
    def _heap_name(self, heap_data_addr, off):
        p = heap_data_addr + off
        e = self.buf.index(b"\x00", p)
        return self.buf[p:e].decode("utf-8")

    def _walk_group_btree(self, addr, heap_data, out):
        b = self.buf
        addr += self.base
        if b[addr:addr + 4] != b"TREE":
            raise Hdf5Error("expected a v1 B-tree node")
        ntype, level, used = b[addr + 4], b[addr + 5], self._u16(addr + 6)
        if ntype != 0:
            raise Hdf5Error("group B-tree expected")
        p = addr + 24 + 8                       # signature, type, level, entries, two siblings; then key 0
        for _ in range(used):
            child = self._u64(p)
            p += 16                             # child address + the next key
            if level > 0:
                self._walk_group_btree(child, heap_data, out)
                continue
            s = child + self.base
            if b[s:s + 4] != b"SNOD":
                raise Hdf5Error("expected a symbol table node")
            n = self._u16(s + 6)
            for i in range(n):
                e = s + 8 + 40 * i
                out[self._heap_name(heap_data, self._u64(e))] = self._u64(e + 8)

    def _entries(self, header_addr):
        if header_addr in self._groups:
            return self._groups[header_addr]
        for mtype, data in self._messages(header_addr):
            if mtype == 0x11:                   # symbol table message: B-tree address, local heap address
                btree, heap = struct.unpack_from("<QQ", data, 0)
                h = heap + self.base
                if self.buf[h:h + 4] != b"HEAP":
                    raise Hdf5Error("expected a local heap")
                heap_data = self._u64(h + 24) + self.base
                out = {}
                self._walk_group_btree(btree, heap_data, out)
                self._groups[header_addr] = out
                return out
            if mtype in (0x02, 0x06):           # link info / link messages: the 1.8 "compact / dense" group storage
                raise NotImplementedError("groups stored as link messages (file written with libver='latest')")
        return None                             # not a group

    def _resolve(self, path):
        addr = self.root
        for part in [p for p in path.split("/") if p]:
            ent = self._entries(addr)
            if ent is None or part not in ent:
                return None
            addr = ent[part]
        return addr

    def __contains__(self, path):
        return self._resolve(path) is not None

    def keys(self, group=""):
        addr = self._resolve(group)
        ent = self._entries(addr) if addr is not None else None
        return sorted(ent) if ent else []

    # -- datasets
    @staticmethod
    def _dtype(data):
        cls, ver = data[0] & 0x0F, data[0] >> 4
        bits0 = data[1]
        size = struct.unpack_from("<I", data, 4)[0]
        if ver not in (1, 2, 3):
            raise Hdf5Error(f"datatype message version {ver}")
        order = ">" if bits0 & 1 else "<"
        if cls == 0:
            return np.dtype(f"{order}{'i' if bits0 & 0x08 else 'u'}{size}")
        if cls == 1:
            return np.dtype(f"{order}f{size}")
        if cls == 3:
            return np.dtype(f"S{size}")
        raise NotImplementedError(f"HDF5 datatype class {cls} (only integers, floats and fixed-length strings are handled)")

    @staticmethod
    def _shape(data):
        ver, rank = data[0], data[1]
        if ver == 1:
            off = 8
        elif ver == 2:
            off = 4
        else:
            raise Hdf5Error(f"dataspace message version {ver}")
        return tuple(struct.unpack_from(f"<{rank}Q", data, off)) if rank else ()

    @staticmethod
    def _filters(data):
        ver, n = data[0], data[1]
        pos = 8 if ver == 1 else 2
        out = []
        for _ in range(n):
            fid = struct.unpack_from("<H", data, pos)[0]
            if ver == 1 or fid >= 256:
                name_len = struct.unpack_from("<H", data, pos + 2)[0]
                pos += 4
            else:
                name_len = 0
                pos += 2
            nvals = struct.unpack_from("<H", data, pos + 2)[0]
            pos += 4
            pos += (name_len + 7) // 8 * 8 if ver == 1 else name_len
            vals = struct.unpack_from(f"<{nvals}I", data, pos) if nvals else ()
            pos += 4 * nvals
            if ver == 1 and nvals % 2:
                pos += 4
            out.append((fid, vals))
        return out

    def _chunks(self, addr, rank, out):
        b = self.buf
        addr += self.base
        if b[addr:addr + 4] != b"TREE" or b[addr + 4] != 1:
            raise Hdf5Error("expected a chunk B-tree node")
        level, used = b[addr + 5], self._u16(addr + 6)
        key = 8 + 8 * (rank + 1)                # chunk size, filter mask, rank + 1 offsets
        p = addr + 24
        for _ in range(used):
            nbytes, mask = self._u32(p), self._u32(p + 4)
            offs = struct.unpack_from(f"<{rank}Q", b, p + 8)
            child = self._u64(p + key)
            if level > 0:
                self._chunks(child, rank, out)
            else:
                out.append((offs, nbytes, mask, child))
            p += key + 8

    def __getitem__(self, path):
        addr = self._resolve(path)
        if addr is None:
            raise KeyError(path)
        msgs = self._messages(addr)
        shape = dtype = layout = None
        filters = []
        for mtype, data in msgs:
            if mtype == 0x01:
                shape = self._shape(data)
            elif mtype == 0x03:
                dtype = self._dtype(data)
            elif mtype == 0x08:
                layout = data
            elif mtype == 0x0B:
                filters = self._filters(data)
        if shape is None or dtype is None or layout is None:
            raise KeyError(f"{path} is not a dataset")
        if layout[0] != 3:
            raise NotImplementedError(f"data layout message version {layout[0]}")
        count = int(np.prod(shape, dtype=np.int64)) if shape else 1
        lclass = layout[1]
        if lclass == 0:                         # compact: the data sit in the message
            n = struct.unpack_from("<H", layout, 2)[0]
            return np.frombuffer(layout[4:4 + n], dtype=dtype, count=count).reshape(shape).copy()
        if lclass == 1:                         # contiguous
            daddr, dsize = struct.unpack_from("<QQ", layout, 2)
            if daddr == UNDEF:
                return np.zeros(shape, dtype)   # never written: the fill value (0)
            if dsize < count * dtype.itemsize:
                raise Hdf5Error(f"{path}: storage smaller than the dataspace")
            return np.frombuffer(self.buf, dtype=dtype, count=count, offset=daddr + self.base).reshape(shape).copy()
        if lclass == 2:                         # chunked
            rank = layout[2] - 1
            btree = struct.unpack_from("<Q", layout, 3)[0]
            cdims = struct.unpack_from(f"<{rank}I", layout, 11)
            out = np.zeros(shape, dtype)
            if btree == UNDEF:
                return out
            chunks = []
            self._chunks(btree, rank, chunks)
            for offs, nbytes, mask, caddr in chunks:
                raw = self.buf[caddr + self.base:caddr + self.base + nbytes]
                for i, (fid, vals) in reversed(list(enumerate(filters))):
                    if mask & (1 << i):
                        continue
                    if fid == 1:
                        raw = zlib.decompress(raw)
                    elif fid == 2:              # shuffle: bytes of equal significance were stored together
                        es = vals[0] if vals else dtype.itemsize
                        raw = np.frombuffer(raw, np.uint8).reshape(es, -1).T.tobytes()
                    else:
                        raise NotImplementedError(f"HDF5 filter {fid}")
                block = np.frombuffer(raw, dtype=dtype, count=int(np.prod(cdims))).reshape(cdims)
                sl = tuple(slice(o, min(o + c, s)) for o, c, s in zip(offs, cdims, shape))
                out[sl] = block[tuple(slice(0, s.stop - s.start) for s in sl)]
            return out
        raise NotImplementedError(f"data layout class {lclass}")

    def get(self, path, default=None):
        try:
            return self[path]
        except KeyError:
            return default

    def close(self):
        self.buf = b""

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()


# ---------------------------------------------------------------------------------------------------------------------
# writer (contiguous datasets, or chunked with deflate (+ shuffle); groups as symbol tables)
# ---------------------------------------------------------------------------------------------------------------------
def _pad8(b):
    return b + b"\x00" * (-len(b) % 8)


def _msg(mtype, data):
    data = _pad8(data)
    return struct.pack("<HHB3x", mtype, len(data), 0) + data


def _dtype_msg(dt):
    dt = np.dtype(dt)
    if dt.kind in "iu":
        bits = (0x08 if dt.kind == "i" else 0) | (1 if dt.byteorder == ">" else 0)
        return struct.pack("<BBBBI", 0x10 | 0, bits, 0, 0, dt.itemsize) + struct.pack("<HH", 0, 8 * dt.itemsize)
    if dt.kind == "f":
        props = {4: (0, 32, 23, 8, 0, 23, 127), 8: (0, 64, 52, 11, 0, 52, 1023)}[dt.itemsize]
        sign_loc = 8 * dt.itemsize - 1
        return struct.pack("<BBBBI", 0x10 | 1, 0x20 | (1 if dt.byteorder == ">" else 0), sign_loc, 0, dt.itemsize) + \
            struct.pack("<HHBBBBI", *props)
    if dt.kind == "S":
        return struct.pack("<BBBBI", 0x10 | 3, 0, 0, 0, dt.itemsize)
    raise NotImplementedError(f"dtype {dt}")


class _Writer:
    def __init__(self):
        self.out = bytearray(b"\x00" * 96)      # superblock (24 + 4 * 8 + 40) is filled in at the end

    def alloc(self, data):
        self.out += b"\x00" * (-len(self.out) % 8)
        addr = len(self.out)
        self.out += data
        return addr

    def object_header(self, msgs):
        body = b"".join(msgs)
        return self.alloc(struct.pack("<BBHII4x", 1, 0, len(msgs), 1, len(body)) + body)

    def dataset(self, arr, chunks=None, compress=False, shuffle=False):
        arr = np.asarray(arr)
        if arr.ndim and not arr.flags.c_contiguous:
            arr = np.ascontiguousarray(arr)
        dims = b"".join(struct.pack("<Q", d) for d in arr.shape)
        space = _msg(0x01, struct.pack("<BBB5x", 1, arr.ndim, 0) + dims)
        dtype = _msg(0x03, _dtype_msg(arr.dtype))
        msgs = [space, dtype]
        if chunks is None:
            daddr = self.alloc(arr.tobytes()) if arr.size else UNDEF
            msgs.append(_msg(0x08, struct.pack("<BBQQ", 3, 1, daddr, arr.nbytes)))
        else:
            filters = []
            if shuffle:
                filters.append((2, (arr.dtype.itemsize,)))
            if compress:
                filters.append((1, (4,)))
            rank = arr.ndim
            keys = []
            import itertools
            for idx in itertools.product(*[range(0, s, c) for s, c in zip(arr.shape, chunks)]):
                block = np.zeros(chunks, arr.dtype)
                sl = tuple(slice(o, min(o + c, s)) for o, c, s in zip(idx, chunks, arr.shape))
                block[tuple(slice(0, s.stop - s.start) for s in sl)] = arr[sl]
                raw = block.tobytes()
                if shuffle:
                    raw = np.frombuffer(raw, np.uint8).reshape(-1, arr.dtype.itemsize).T.tobytes()
                if compress:
                    raw = zlib.compress(raw, 4)
                keys.append((idx, len(raw), self.alloc(raw)))
            if len(keys) > 64:
                raise NotImplementedError("the minimal writer keeps all chunks in one B-tree node (<= 64 chunks)")
            node = b"TREE" + struct.pack("<BBHQQ", 1, 0, len(keys), UNDEF, UNDEF)
            for idx, nbytes, caddr in keys:
                node += struct.pack("<II", nbytes, 0) + b"".join(struct.pack("<Q", o) for o in idx) + struct.pack("<Q", 0)
                node += struct.pack("<Q", caddr)
            node += struct.pack("<II", 0, 0) + b"".join(struct.pack("<Q", s) for s in arr.shape) + struct.pack("<Q", 0)
            baddr = self.alloc(node)
            msgs.append(_msg(0x08, struct.pack("<BBB", 3, 2, rank + 1) + struct.pack("<Q", baddr) +
                             b"".join(struct.pack("<I", c) for c in chunks) + struct.pack("<I", arr.dtype.itemsize)))
            if filters:
                body = struct.pack("<BB6x", 1, len(filters))
                for fid, vals in filters:
                    body += struct.pack("<HHHH", fid, 0, 1, len(vals)) + b"".join(struct.pack("<I", v) for v in vals)
                    if len(vals) % 2:
                        body += b"\x00" * 4
                msgs.append(_msg(0x0B, body))
        return self.object_header(msgs)

    def group(self, entries):
        """entries: dict name -> object header address. Returns (object header address, B-tree address, heap address)."""
        names = sorted(entries)
        heap = bytearray(b"\x00" * 8)           # offset 0: the empty name
        offs = {}
        for n in names:
            offs[n] = len(heap)
            heap += _pad8(n.encode("utf-8") + b"\x00")
        heap_data = self.alloc(bytes(heap))
        heap_addr = self.alloc(b"HEAP" + struct.pack("<B3xQQQ", 0, len(heap), UNDEF, heap_data))
        snods = []
        for i in range(0, max(len(names), 1), 8):   # leaf K = 4: up to 8 symbols per node
            part = names[i:i + 8]
            node = b"SNOD" + struct.pack("<BBH", 1, 0, len(part))
            for n in part:
                node += struct.pack("<QQII16x", offs[n], entries[n], 0, 0)
            node += b"\x00" * (40 * (8 - len(part)))
            snods.append((self.alloc(node), offs[part[-1]] if part else 0))
        if len(snods) > 32:
            raise NotImplementedError("the minimal writer keeps a group in one B-tree node (<= 256 members)")
        tree = b"TREE" + struct.pack("<BBHQQ", 0, 0, len(snods), UNDEF, UNDEF) + struct.pack("<Q", 0)
        for addr, last in snods:
            tree += struct.pack("<QQ", addr, last)
        btree = self.alloc(tree)
        hdr = self.object_header([_msg(0x11, struct.pack("<QQ", btree, heap_addr))])
        return hdr, btree, heap_addr

    def finish(self, root):
        hdr, btree, heap = root
        sb = SIGNATURE + struct.pack("<BBBBBBBBHHI", 0, 0, 0, 0, 0, 8, 8, 0, 4, 16, 0)
        sb += struct.pack("<QQQQ", 0, UNDEF, len(self.out), UNDEF)
        sb += struct.pack("<QQII", 0, hdr, 1, 0) + struct.pack("<QQ", btree, heap)
        self.out[:len(sb)] = sb
        return bytes(self.out)


def write(path, tree, chunks=None, compress=False, shuffle=False):
    """tree: dict name -> array | dict (a group). chunks: optional dict dataset name -> chunk shape."""
    w = _Writer()

    def build(node, prefix):
        ent = {}
        for name, v in node.items():
            if isinstance(v, dict):
                ent[name] = build(v, prefix + name + "/")[0]
            else:
                ch = (chunks or {}).get(prefix + name)
                ent[name] = w.dataset(v, chunks=ch, compress=compress and ch is not None, shuffle=shuffle and ch is not None)
        return w.group(ent)

    data = w.finish(build(tree, ""))
    with open(path, "wb") as f:
        f.write(data)
