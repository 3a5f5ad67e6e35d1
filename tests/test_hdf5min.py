"""CPU: the built-in HDF5 writer/reader (mdsea_b200/hdf5min.py) and the SysManager file contract."""
import os
import struct

import numpy as np
import pytest

from mdsea_b200 import hdf5min

SCIPY_H5 = None
try:
    import scipy.io.matlab as _m

    _p = os.path.join(os.path.dirname(_m.__file__), "tests", "data", "testhdf5_7.4_GLNX86.mat")
    if os.path.exists(_p):
        SCIPY_H5 = _p
except Exception:  # pragma: no cover
    pass


def test_roundtrip_simfile_layout(tmp_path):
    p = str(tmp_path / "data.h5")
    f = hdf5min.File(p, "a")
    steps, ndim, n = 7, 3, 11
    for name in ("r_coords", "v_coords"):
        f.create_dataset(name, dtype=np.float64, shape=(steps, ndim, n))
    for name in ("pot_energy", "kin_energy", "temp"):
        f.create_dataset(name, dtype=np.float64, shape=(steps,))
    rng = np.random.default_rng(0)
    r = rng.normal(size=(steps, ndim, n))
    for i in range(steps):
        f["r_coords"][i] = [r[i]]          # h5py-style broadcast of a leading length-1 axis (update_files)
        f["pot_energy"][i] = [float(i)]
    f.close()
    assert not f.id.valid
    with pytest.raises(ValueError):
        f["r_coords"]
    g = hdf5min.File(p, "r")
    assert g.keys() == ["kin_energy", "pot_energy", "r_coords", "temp", "v_coords"]
    np.testing.assert_array_equal(g["r_coords"][:], r)
    np.testing.assert_array_equal(g["pot_energy"][:], np.arange(steps, dtype=float))
    np.testing.assert_array_equal(g["v_coords"][:], 0.0)   # HDF5 default fill value
    assert g["r_coords"].shape == (steps, ndim, n) and g["temp"].dtype == np.float64
    g.close()
    # reopen in append mode and keep writing (SysManager.load -> update_ds)
    h = hdf5min.File(p, "a")
    h["temp"][3] = 2.5
    h.close()
    assert hdf5min.File(p, "r")["temp"][3] == 2.5


def test_file_structure_follows_hdf5_spec(tmp_path):
    """Byte-level checks of the structures libhdf5 will look for (superblock v0, TREE, HEAP, SNOD, v1 headers)."""
    p = str(tmp_path / "s.h5")
    f = hdf5min.File(p, "w")
    f.create_dataset("pot_energy", dtype=np.float64, shape=(4,))
    f.create_dataset("r_coords", dtype=np.float64, shape=(4, 2, 3))
    f.close()
    b = open(p, "rb").read()
    assert b[:8] == b"\x89HDF\r\n\x1a\n" and b[8] == 0 and b[13] == 8 and b[14] == 8
    leaf_k, int_k = struct.unpack_from("<HH", b, 16)
    assert (leaf_k, int_k) == (4, 16)
    base, fsinfo, eof, drv = struct.unpack_from("<QQQQ", b, 24)
    assert base == 0 and fsinfo == drv == 0xFFFFFFFFFFFFFFFF and eof == len(b)
    _, root_oh, cache, _, btree, heap = struct.unpack_from("<QQIIQQ", b, 56)
    assert cache == 1 and b[btree:btree + 4] == b"TREE" and b[heap:heap + 4] == b"HEAP"
    assert b[root_oh] == 1 and struct.unpack_from("<H", b, root_oh + 16)[0] == 0x0011   # symbol-table message
    nent = struct.unpack_from("<H", b, btree + 6)[0]
    k0, child, k1 = struct.unpack_from("<QQQ", b, btree + 24)
    hsize, hfree, hdata = struct.unpack_from("<QQQ", b, heap + 8)
    assert nent == 1 and k0 == 0 and b[child:child + 4] == b"SNOD"
    assert b[hdata + k1:hdata + k1 + 9] == b"r_coords\0"          # key = largest name in the node
    nxt, fsz = struct.unpack_from("<QQ", b, hdata + hfree)
    assert nxt == 1 and hfree + fsz == hsize                        # single free block to the end of the heap
    nsym = struct.unpack_from("<H", b, child + 6)[0]
    assert nsym == 2
    names = []
    for k in range(nsym):
        noff, oh = struct.unpack_from("<QQ", b, child + 8 + 40 * k)
        names.append(b[hdata + noff:b.index(b"\0", hdata + noff)].decode())
        ver, nmsg, refc, hsz = struct.unpack_from("<BxHII", b, oh)
        assert ver == 1 and nmsg == 4 and refc == 1 and hsz % 8 == 0
        types = []
        off = oh + 16
        for _ in range(nmsg):
            t, sz = struct.unpack_from("<HH", b, off)
            types.append(t)
            assert sz % 8 == 0
            off += 8 + sz
        assert types == [0x0001, 0x0003, 0x0005, 0x0008] and off == oh + 16 + hsz
    assert names == sorted(names)                                    # symbol entries ordered by name
    # the float64 datatype message is byte-identical to what libhdf5 writes
    assert bytes.fromhex("11203f000800000000004000340b0034ff030000") in b


@pytest.mark.skipif(SCIPY_H5 is None, reason="no libhdf5-written sample file available")
def test_reader_parses_a_libhdf5_written_file():
    f = hdf5min.File(SCIPY_H5, "r")
    assert "testdouble" in f
    d = f["testdouble"]
    assert d.dtype == np.float64 and d.shape == (9, 1)
    np.testing.assert_allclose(d[:].ravel(), np.arange(9) * np.pi / 4, rtol=1e-12)
    f.close()


def test_limits(tmp_path):
    f = hdf5min.File(str(tmp_path / "l.h5"), "w")
    for k in range(8):
        f.create_dataset(f"d{k}", dtype=np.float64, shape=(2,))
    with pytest.raises(NotImplementedError):
        f.create_dataset("one_too_many", dtype=np.float64, shape=(2,))
    with pytest.raises(ValueError):
        f.create_dataset("d0", dtype=np.float64, shape=(2,))
    with pytest.raises(KeyError):
        f["nope"]
    f.create_dataset  # noqa
    f.close()
    z = hdf5min.File(str(tmp_path / "z.h5"), "w")
    z.create_dataset("empty", dtype=np.float64, shape=(0, 3))
    z.create_dataset("i", data=np.arange(5, dtype=np.int32))
    z.close()
    z = hdf5min.File(str(tmp_path / "z.h5"), "r")
    assert z["empty"].shape == (0, 3)
    np.testing.assert_array_equal(z["i"][:], np.arange(5))
