"""Self-contained minimal HDF5 writer/reader for the simfile layout (no h5py / libhdf5 in the image).

mdsea stores a run as ``simfiles/<simid>/data.h5`` holding five root datasets, float64,
contiguous, no attributes: ``r_coords``/``v_coords`` (STEPS, NDIM, N) and
``pot_energy``/``kin_energy``/``temp`` (STEPS,) -- mdsea/core.py:230-244.  This module writes
exactly that subset of the HDF5 1.x file format, in the "earliest" encodings that h5py 3.0's
``create_dataset(name, dtype, shape)`` produces: version-0 superblock, root group as a
version-1 object header with a symbol-table message (v1 B-tree "TREE" + symbol node "SNOD" +
local heap "HEAP"), datasets as version-1 object headers with dataspace (v1), datatype
(IEEE little-endian float, v1), fill value (v2, default) and contiguous layout (v3) messages.
Field encodings were cross-checked byte by byte against a libhdf5-written file that ships with
SciPy (scipy/io/matlab/tests/data/testhdf5_7.4_GLNX86.mat), which tests/test_hdf5min.py also
parses with the reader below.

The API is the small part of h5py the reference touches (core.py:218,238,242,321,330,349):
``File(path, mode)``, ``create_dataset(name, dtype=, shape=)``, ``f[name][i] = x``,
``f[name][:]``, ``f.id.valid``, ``f.close()``.  Dataset payloads are numpy memmaps, so a row
assignment is a memcpy into the page cache and the buffered k-step flush of the solver can
write straight into the file mapping.
"""
from __future__ import annotations

import os
import struct
from typing import Dict, Optional, Tuple

import numpy as np

SIGNATURE = b"\x89HDF\r\n\x1a\n"
UNDEF = 0xFFFFFFFFFFFFFFFF
LEAF_K = 4        # symbol-table node holds up to 2*LEAF_K entries
INTERNAL_K = 16
MAX_DATASETS = 2 * LEAF_K

# fixed metadata map of files written by this module (all inside the first 4 KiB)
_OFF_ROOT_OH = 96
_OFF_BTREE = 136
_OFF_HEAP = 680
_OFF_HEAP_DATA = 712
_HEAP_DATA_SIZE = 256
_OFF_SNOD = 968
_OFF_DSET_OH = 1296
_DSET_OH_STRIDE = 256
_DATA_START = 4096
_DATA_ALIGN = 4096


def _pad8(n: int) -> int:
    return (n + 7) & ~7


class _Id:
    def __init__(self):
        self.valid = True


class Dataset:
    """h5py.Dataset look-alike over a numpy memmap (contiguous layout only)."""

    def __init__(self, name: str, arr: np.ndarray):
        self.name = "/" + name
        self._a = arr
        self.shape = arr.shape
        self.dtype = arr.dtype

    def __setitem__(self, key, value):
        # h5py broadcasts ``dset[i] = [arr]`` (a leading length-1 axis) onto the row -- the reference's
        # update_files passes exactly that (mdsea/simulator.py:147-150)
        target = self._a[key]
        val = np.asarray(value, dtype=self._a.dtype)
        if isinstance(target, np.ndarray):
            if val.shape != target.shape and val.size == target.size:
                val = val.reshape(target.shape)
        elif val.size == 1:
            val = val.reshape(())[()]
        self._a[key] = val

    def __getitem__(self, key):
        out = self._a[key]
        return np.array(out) if isinstance(out, np.ndarray) else out

    def __len__(self):
        return self.shape[0]

    @property
    def raw(self) -> np.ndarray:
        """The live memmap view (zero-copy target for the solver's buffered flush)."""
        return self._a


def _dtype_message(dt: np.dtype) -> bytes:
    dt = np.dtype(dt)
    if dt == np.float64:
        return bytes([0x11, 0x20, 0x3F, 0x00]) + struct.pack("<I", 8) + struct.pack("<HHBBBBI", 0, 64, 52, 11, 0, 52, 1023)
    if dt == np.float32:
        return bytes([0x11, 0x20, 0x1F, 0x00]) + struct.pack("<I", 4) + struct.pack("<HHBBBBI", 0, 32, 23, 8, 0, 23, 127)
    if dt.kind in "iu" and dt.itemsize in (1, 2, 4, 8):
        signed = 0x08 if dt.kind == "i" else 0x00
        return bytes([0x10, signed, 0x00, 0x00]) + struct.pack("<I", dt.itemsize) + struct.pack("<HH", 0, 8 * dt.itemsize)
    raise TypeError(f"hdf5min: unsupported dtype {dt}")


def _parse_dtype(msg: bytes) -> np.dtype:
    cls = msg[0] & 0x0F
    bits0 = msg[1]
    size = struct.unpack_from("<I", msg, 4)[0]
    if bits0 & 1:
        raise TypeError("hdf5min: big-endian datatypes are not supported")
    if cls == 1:
        return np.dtype({4: "<f4", 8: "<f8"}[size])
    if cls == 0:
        return np.dtype(("<i" if bits0 & 0x08 else "<u") + str(size))
    raise TypeError(f"hdf5min: unsupported datatype class {cls}")


def _msg(mtype: int, body: bytes, flags: int = 0) -> bytes:
    body = body + b"\0" * (_pad8(len(body)) - len(body))
    return struct.pack("<HHB3x", mtype, len(body), flags) + body


def _object_header(messages) -> bytes:
    blob = b"".join(messages)
    return struct.pack("<BxHII4x", 1, len(messages), 1, len(blob)) + blob


class File:
    """h5py.File look-alike.  Modes: 'w' (create/truncate), 'a' (open or create), 'r' / 'r+'."""

    def __init__(self, path: str, mode: str = "a"):
        self.filename = path
        self.mode = mode
        self.id = _Id()
        self._base = 0
        self._order = []  # dataset names in creation order
        self._meta: Dict[str, Tuple[np.dtype, Tuple[int, ...], int]] = {}
        self._maps: Dict[str, Dataset] = {}
        self._foreign = False
        exists = os.path.exists(path) and os.path.getsize(path) > 0
        if mode == "r" and not exists:
            raise FileNotFoundError(path)
        if mode == "w" or not exists:
            if mode in ("r", "r+"):
                raise FileNotFoundError(path)
            with open(path, "wb") as f:
                f.truncate(_DATA_START)
            self._write_metadata()
        else:
            self._read_metadata()

    # -- h5py surface --------------------------------------------------------------------
    def create_dataset(self, name: str, shape=None, dtype=None, data=None) -> Dataset:
        self._check_open()
        if self.mode == "r":
            raise ValueError("hdf5min: file is open read-only")
        if self._foreign:
            raise NotImplementedError("hdf5min: can only add datasets to files written by hdf5min")
        if name in self._meta:
            raise ValueError(f"Unable to create dataset (name already exists): {name}")
        if len(self._meta) >= MAX_DATASETS:
            raise NotImplementedError(f"hdf5min: at most {MAX_DATASETS} root datasets")
        if data is not None:
            data = np.asarray(data)
            shape = data.shape if shape is None else shape
            dtype = data.dtype if dtype is None else dtype
        if isinstance(shape, int):
            shape = (shape,)
        dt = np.dtype(dtype if dtype is not None else np.float32)
        shape = tuple(int(s) for s in shape)
        if len(name.encode()) + 1 > 31:
            raise NotImplementedError("hdf5min: dataset names are limited to 30 bytes")
        nbytes = int(np.prod(shape, dtype=np.int64)) * dt.itemsize
        size = os.path.getsize(self.filename)
        addr = max(_DATA_START, (size + _DATA_ALIGN - 1) // _DATA_ALIGN * _DATA_ALIGN)
        with open(self.filename, "r+b") as f:
            f.truncate(addr + max(nbytes, 1))
        self._order.append(name)
        self._meta[name] = (dt, shape, addr)
        self._write_metadata()
        ds = self[name]
        if data is not None:
            ds[...] = data
        return ds

    def __getitem__(self, name: str) -> Dataset:
        self._check_open()
        name = name.lstrip("/")
        if name not in self._meta:
            raise KeyError(f"Unable to open object (object '{name}' doesn't exist)")
        if name not in self._maps:
            dt, shape, addr = self._meta[name]
            mode = "r" if self.mode == "r" else "r+"
            if int(np.prod(shape, dtype=np.int64)) == 0:
                arr = np.zeros(shape, dtype=dt)
            else:
                arr = np.memmap(self.filename, dtype=dt, mode=mode, offset=self._base + addr, shape=shape)
            self._maps[name] = Dataset(name, arr)
        return self._maps[name]

    def __contains__(self, name):
        return name.lstrip("/") in self._meta

    def keys(self):
        return sorted(self._meta)

    def flush(self):
        for ds in self._maps.values():
            if isinstance(ds._a, np.memmap):
                ds._a.flush()

    def close(self):
        if not self.id.valid:
            return
        self.flush()
        self._maps.clear()
        self.id.valid = False

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def _check_open(self):
        if not self.id.valid:
            # h5py raises ValueError on a closed file id; SysManager.update_ds relies on it (core.py:329-340)
            raise ValueError("Invalid location identifier (file is closed)")

    # -- writer --------------------------------------------------------------------------
    def _write_metadata(self):
        names = sorted(self._meta)  # symbol-table entries are ordered by name
        heap = bytearray(_HEAP_DATA_SIZE)
        offs = {}
        pos = 8  # offset 0 holds the empty string
        for nm in names:
            raw = nm.encode() + b"\0"
            heap[pos:pos + len(raw)] = raw
            offs[nm] = pos
            pos += _pad8(len(raw))
        if _HEAP_DATA_SIZE - pos >= 16:
            free_head = pos
            struct.pack_into("<QQ", heap, pos, 1, _HEAP_DATA_SIZE - pos)  # next = H5HL_FREE_NULL, size
        else:
            free_head = 1
        eof = max(os.path.getsize(self.filename), _DATA_START)

        sb = bytearray()
        sb += SIGNATURE
        sb += bytes([0, 0, 0, 0, 0, 8, 8, 0])
        sb += struct.pack("<HHI", LEAF_K, INTERNAL_K, 0)
        sb += struct.pack("<QQQQ", 0, UNDEF, eof, UNDEF)
        sb += struct.pack("<QQII", 0, _OFF_ROOT_OH, 1, 0) + struct.pack("<QQ", _OFF_BTREE, _OFF_HEAP)
        assert len(sb) == 96

        root_oh = _object_header([_msg(0x0011, struct.pack("<QQ", _OFF_BTREE, _OFF_HEAP))])
        assert len(root_oh) == _OFF_BTREE - _OFF_ROOT_OH

        nent = 1 if names else 0
        bt = bytearray(24 + (2 * INTERNAL_K + 1) * 8 + 2 * INTERNAL_K * 8)
        bt[0:4] = b"TREE"
        struct.pack_into("<BBHQQ", bt, 4, 0, 0, nent, UNDEF, UNDEF)
        struct.pack_into("<Q", bt, 24, 0)
        if names:
            struct.pack_into("<QQ", bt, 32, _OFF_SNOD, offs[names[-1]])
        assert len(bt) == _OFF_HEAP - _OFF_BTREE

        hp = b"HEAP" + bytes([0, 0, 0, 0]) + struct.pack("<QQQ", _HEAP_DATA_SIZE, free_head, _OFF_HEAP_DATA)

        sn = bytearray(8 + 2 * LEAF_K * 40)
        sn[0:4] = b"SNOD"
        struct.pack_into("<BBH", sn, 4, 1, 0, len(names))
        for k, nm in enumerate(names):
            slot = self._order.index(nm)
            struct.pack_into("<QQII16x", sn, 8 + 40 * k, offs[nm], _OFF_DSET_OH + slot * _DSET_OH_STRIDE, 0, 0)

        with open(self.filename, "r+b") as f:
            f.seek(0)
            f.write(sb)
            f.write(root_oh)
            f.write(bt)
            f.write(hp)
            f.write(heap)
            assert f.tell() == _OFF_SNOD
            f.write(sn)
            for nm in self._order:
                dt, shape, addr = self._meta[nm]
                slot = self._order.index(nm)
                nbytes = int(np.prod(shape, dtype=np.int64)) * dt.itemsize
                msgs = [
                    _msg(0x0001, struct.pack("<BBB5x", 1, len(shape), 0) + b"".join(struct.pack("<Q", s) for s in shape)),
                    _msg(0x0003, _dtype_message(dt), flags=1),
                    _msg(0x0005, struct.pack("<BBBBI", 2, 2, 2, 1, 0), flags=1),
                    _msg(0x0008, struct.pack("<BBQQ", 3, 1, addr, nbytes)),
                ]
                oh = _object_header(msgs)
                assert len(oh) <= _DSET_OH_STRIDE
                f.seek(_OFF_DSET_OH + slot * _DSET_OH_STRIDE)
                f.write(oh)

    # -- reader --------------------------------------------------------------------------
    def _read_metadata(self):
        with open(self.filename, "rb") as f:
            buf_head = f.read(1 << 20)
            base = -1
            off = 0
            while off < len(buf_head):  # superblock may sit after a user block at 0, 512, 1024, ...
                if buf_head[off:off + 8] == SIGNATURE:
                    base = off
                    break
                off = 512 if off == 0 else off * 2
            if base < 0:
                raise OSError(f"hdf5min: {self.filename} is not an HDF5 file")
            sb = buf_head[base:base + 96]
            if sb[8] != 0 or sb[13] != 8 or sb[14] != 8:
                raise NotImplementedError("hdf5min: only version-0 superblocks with 8-byte offsets/lengths")
            base_addr = struct.unpack_from("<Q", sb, 24)[0]
            self._base = base_addr
            btree, heap = struct.unpack_from("<QQ", sb, 80)
            size = os.path.getsize(self.filename)

            def rd(addr, n):
                f.seek(self._base + addr)
                return f.read(n)

            hh = rd(heap, 32)
            if hh[:4] != b"HEAP":
                raise OSError("hdf5min: bad local heap")
            hsize, _free, hdata = struct.unpack_from("<QQQ", hh, 8)
            names_blob = rd(hdata, hsize)

            entries = []

            def walk(addr):
                node = rd(addr, 24)
                if node[:4] != b"TREE":
                    raise OSError("hdf5min: bad B-tree node")
                level, nent = node[5], struct.unpack_from("<H", node, 6)[0]
                body = rd(addr + 24, 8 * (2 * nent + 1))
                for k in range(nent):
                    child = struct.unpack_from("<Q", body, 8 * (2 * k + 1))[0]
                    if level > 0:
                        walk(child)
                    else:
                        sn = rd(child, 8)
                        if sn[:4] != b"SNOD":
                            raise OSError("hdf5min: bad symbol node")
                        ns = struct.unpack_from("<H", sn, 6)[0]
                        eb = rd(child + 8, 40 * ns)
                        for j in range(ns):
                            noff, oh = struct.unpack_from("<QQ", eb, 40 * j)
                            nm = names_blob[noff:names_blob.index(b"\0", noff)].decode()
                            entries.append((nm, oh))

            if btree != UNDEF:
                walk(btree)

            for nm, oh in entries:
                info = self._parse_object_header(rd, oh)
                if info is not None:
                    self._order.append(nm)
                    self._meta[nm] = info
            self._foreign = not (self._base == 0 and btree == _OFF_BTREE and heap == _OFF_HEAP and all(
                _OFF_DSET_OH <= oh < _DATA_START for _, oh in entries))
            if not self._foreign:
                self._order.sort(key=lambda n: dict(entries)[n])
            del size

    @staticmethod
    def _parse_object_header(rd, addr) -> Optional[Tuple[np.dtype, Tuple[int, ...], int]]:
        pre = rd(addr, 16)
        if pre[0] != 1:
            return None  # version-2 headers ("OHDR") are not produced by the layouts handled here
        nmsg, _ref, hsz = struct.unpack_from("<HII", pre, 2)
        blocks = [(addr + 16, hsz)]
        shape = dt = data_addr = None
        seen = 0
        while blocks and seen < nmsg:
            baddr, blen = blocks.pop(0)
            blob = rd(baddr, blen)
            off = 0
            while off + 8 <= blen and seen < nmsg:
                mtype, msz = struct.unpack_from("<HH", blob, off)
                body = blob[off + 8:off + 8 + msz]
                off += 8 + msz
                seen += 1
                if mtype == 0x0001:
                    ver, rank, flags = body[0], body[1], body[2]
                    start = 8 if ver == 1 else 4
                    shape = tuple(struct.unpack_from("<Q", body, start + 8 * k)[0] for k in range(rank))
                elif mtype == 0x0003:
                    try:
                        dt = _parse_dtype(body)
                    except (TypeError, KeyError):
                        return None
                elif mtype == 0x0008:
                    ver = body[0]
                    if ver == 3 and body[1] == 1:
                        data_addr = struct.unpack_from("<Q", body, 2)[0]
                    elif ver in (1, 2) and body[2] == 1:
                        data_addr = struct.unpack_from("<Q", body, 8)[0]
                    else:
                        return None  # chunked / compact
                elif mtype == 0x0010:
                    caddr, clen = struct.unpack_from("<QQ", body, 0)
                    blocks.append((caddr, clen))
        if shape is None or dt is None or data_addr is None or data_addr == UNDEF:
            return None
        return dt, shape, data_addr
