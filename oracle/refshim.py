"""TEST INFRASTRUCTURE ONLY -- never imported by the product (mdsea_b200/).

Loader for the *unmodified* upstream mdsea sources at /root/reference so they can
be executed in this container as the fp64 ground truth.
The reference does not import as shipped here; the three shims below do not touch
its arithmetic:

  1. importlib.metadata.version("mdsea") -> "0.1.0"   (mdsea/_version.py:9; package not installed)
  2. numpy.core.umath_tests.inner1d                  (mdsea/simulator.py:14,244; module removed in NumPy 2)
     -> np.einsum("ij,ij->i", a, b)
  3. h5py.File                                        (mdsea/core.py:7,218,238,321,349; h5py absent)
     -> in-memory stand-in offering create_dataset / __getitem__ / id.valid / close

/root/reference does not exist on the GPU box: nothing under tests -m gpu, smoke()
or bench.py may call this module at run time.  It is used by oracle/gen_golden.py
(which writes tests/golden/*.npz) and by the CPU-only pinning tests when the
reference tree is present.
"""
import importlib
import importlib.metadata
import os
import sys
import types

import numpy as np

REFERENCE_ROOT = os.environ.get("MDSEA_REFERENCE_ROOT", "/root/reference")


def reference_available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "mdsea"))


class _FakeId:
    def __init__(self):
        self.valid = True


class _FakeDataset:
    """h5py.Dataset stand-in: row assignment broadcasts like h5py ([arr] -> arr)."""

    def __init__(self, shape, dtype):
        self.a = np.zeros(shape, dtype=dtype)
        self.shape = self.a.shape
        self.dtype = self.a.dtype

    def __setitem__(self, i, x):
        x = np.asarray(x, dtype=self.a.dtype)
        self.a[i] = x.reshape(self.a[i].shape)

    def __getitem__(self, i):
        return self.a[i]


class _FakeH5File:
    """In-memory h5py.File stand-in; one store per path so reopen sees the data."""

    _stores = {}

    def __init__(self, path, mode="a"):
        self._path = path
        self._d = _FakeH5File._stores.setdefault(path, {})
        self.id = _FakeId()

    def create_dataset(self, name, dtype=None, shape=None):
        self._d[name] = _FakeDataset(shape, dtype)
        return self._d[name]

    def __getitem__(self, name):
        if not self.id.valid:
            raise ValueError("Invalid location identifier (file closed)")
        return self._d[name]

    def close(self):
        self.id.valid = False


def install_shims():
    # 1. version metadata
    _orig_version = importlib.metadata.version

    def _version(name):
        if name == "mdsea":
            return "0.1.0"
        return _orig_version(name)

    importlib.metadata.version = _version

    # 2. inner1d
    if "numpy.core.umath_tests" not in sys.modules:
        m = types.ModuleType("numpy.core.umath_tests")
        m.inner1d = lambda a, b: np.einsum("ij,ij->i", a, b)
        sys.modules["numpy.core.umath_tests"] = m

    # 3. h5py
    if "h5py" not in sys.modules:
        h = types.ModuleType("h5py")
        h.File = _FakeH5File
        sys.modules["h5py"] = h


def load_reference():
    """Return the reference's modules as a namespace (simulator, core, potentials, ...)."""
    if not reference_available():
        raise RuntimeError(f"reference tree not found at {REFERENCE_ROOT}")
    install_shims()
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    import warnings

    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        mods = types.SimpleNamespace(
            simulator=importlib.import_module("mdsea.simulator"),
            core=importlib.import_module("mdsea.core"),
            potentials=importlib.import_module("mdsea.potentials"),
            gen=importlib.import_module("mdsea.gen"),
            config=importlib.import_module("mdsea.config"),
            helpers=importlib.import_module("mdsea.helpers"),
            quicker=importlib.import_module("mdsea.quicker"),
        )
    return mods
