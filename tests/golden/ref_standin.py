"""
Stand-ins that let the UNMODIFIED reference package (/root/reference/SOAP) be imported and its
halo-property classes be executed in an image without unyt / mpi4py / h5py / VirgoDC / astropy.

Used only by the golden-fixture generators of this directory (build container only; nothing here is
imported by the product or by the GPU-box tests).

The unyt stand-in carries NO dimensions: every ``.to()`` / unit multiplication is the identity, so the
reference's class bodies run with every conversion factor equal to exactly 1.0.  The fixtures are
therefore generated in ONE consistent unit system (lengths in kpc, a = 1: "physical" == "comoving",
``30 * unyt.kpc`` == 30 length units, ``20 * unyt.Mpc`` == 20 length units).  What this pins is the
class-level logic of the reference -- selections, ``<`` vs ``<=``, particle-type concatenation order,
dtypes of the intermediate sums, the radius ladder and its error protocol -- not unyt's coercions
(they need the real unyt).
"""

import sys
import types

import numpy as np


class _Units:
    """Unit object with no dimension bookkeeping."""

    registry = None
    base_value = 1.0
    __array_ufunc__ = None  # ndarray * units defers to __rmul__ below

    def __rmul__(self, other):
        return unyt_array(other)

    def __mul__(self, other):
        return self if isinstance(other, _Units) else unyt_array(other)

    def __pow__(self, p):
        return self

    def __truediv__(self, o):
        return self

    def __rtruediv__(self, o):
        return self if isinstance(o, _Units) else unyt_array(o)

    def __eq__(self, o):
        return isinstance(o, _Units)

    def __ne__(self, o):
        return not isinstance(o, _Units)

    def __hash__(self):
        return 0

    def __repr__(self):
        return "unit"

    def __str__(self):
        return "unit"

    @property
    def units(self):
        return self

    def get_conversion_factor(self, other, dtype=None):
        return 1.0, None

    def same_dimensions_as(self, o):
        return True

    @property
    def dimensions(self):
        return 1


class unyt_array(np.ndarray):
    def __new__(cls, input_array, units=None, dtype=None, registry=None, **kw):
        return np.asarray(input_array, dtype=dtype).view(cls)

    @property
    def units(self):
        return _Units()

    @property
    def value(self):
        return np.asarray(self)

    v = value
    d = value

    def to(self, units, *a, **k):
        return self

    def to_value(self, *a, **k):
        return np.asarray(self)

    in_units = to
    in_base = to
    to_physical = to
    to_comoving = to

    def convert_to_units(self, *a, **k):
        return None

    def copy(self, *a, **k):
        return np.ndarray.copy(self, *a, **k).view(unyt_array)

    def __getitem__(self, idx):
        out = np.ndarray.__getitem__(self, idx)
        if not isinstance(out, np.ndarray):
            out = np.asarray(out).view(unyt_array)
        return out

    def __array_function__(self, func, types_, args, kwargs):
        def strip(x):
            if isinstance(x, unyt_array):
                return np.asarray(x)
            if isinstance(x, (list, tuple)):
                return type(x)(strip(y) for y in x)
            if isinstance(x, dict):
                return {k: strip(v) for k, v in x.items()}
            return x

        def wrap(x):
            if isinstance(x, (np.ndarray, np.generic)) and not isinstance(x, np.bool_):
                return np.asarray(x).view(unyt_array)
            if isinstance(x, tuple):
                return tuple(wrap(y) for y in x)
            return x

        return wrap(func(*strip(args), **strip(kwargs)))

    def __array_wrap__(self, arr, context=None, return_scalar=False):
        return np.asarray(arr).view(unyt_array)


def unyt_quantity(value=0.0, units=None, dtype=None, registry=None, **kw):
    return unyt_array(value, dtype=dtype)


class UnitRegistry:
    lut = {}

    def __init__(self, *a, **k):
        pass

    def add(self, *a, **k):
        pass


class _Anything:
    """Placeholder for third-party objects the executed code paths never touch."""

    def __init__(self, *a, **k):
        pass

    def __call__(self, *a, **k):
        return _Anything()

    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        return _Anything()


def _module(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


class _Comm:
    rank, size = 0, 1

    def Get_rank(self):
        return 0

    def Get_size(self):
        return 1

    def allreduce(self, x, op=None):
        return x

    def Allreduce(self, src, dst, op=None):
        dst[...] = src

    def Reduce(self, src, dst, op=None, root=0):
        dst[...] = src

    def bcast(self, x, root=0):
        return x

    def barrier(self):
        pass

    Barrier = barrier


def parallel_sort(arr, comm=None, return_index=False):
    """single rank: a stable argsort (the within-cell order VirgoDC defines is not pinned by the reference)"""
    idx = np.argsort(arr, kind="stable")
    arr[:] = arr[idx]
    return idx if return_index else None


def install():
    """Register the stand-in modules and put /root/reference on sys.path."""
    if "unyt" in sys.modules and getattr(sys.modules["unyt"], "_soap_b200_standin", False):
        return
    u = _module("unyt", unyt_array=unyt_array, unyt_quantity=unyt_quantity, Unit=lambda *a, **k: _Units(),
                UnitRegistry=UnitRegistry, dimensionless=_Units(), _soap_b200_standin=True)
    u.__path__ = []  # a package: ``import unyt.dimensions`` resolves to the stand-in below

    def _unit_attr(name):  # unyt.kpc, unyt.Mpc, unyt.km, unyt.s, unyt.Msun, ...
        if name.startswith("__"):
            raise AttributeError(name)
        return _Units()

    u.__getattr__ = _unit_attr
    dims = _module("unyt.dimensions")
    dims.__getattr__ = lambda name: 1
    u.dimensions = dims
    # what SOAP/core/swift_units.py (unit_registry_from_snapshot) touches: registries and unit definitions are no-ops
    u.unit_registry = types.SimpleNamespace(UnitRegistry=UnitRegistry)
    u.define_unit = lambda *a, **k: None
    u.UnitSystem = lambda *a, **k: None
    _module("unyt.array", unyt_array=unyt_array, unyt_quantity=unyt_quantity)
    _module("unyt.physical_constants").__getattr__ = lambda name: unyt_array(1.0)
    mpi = types.SimpleNamespace(MIN="min", MAX="max", SUM="sum", COMM_WORLD=_Comm(), COMM_TYPE_SHARED=0,
                                Comm=_Comm, Win=_Anything(), DOUBLE=None, Wtime=lambda: 0.0, IN_PLACE=None)
    _module("mpi4py", MPI=mpi)
    sys.modules["mpi4py.MPI"] = mpi
    _module("h5py", File=_Anything, Dataset=_Anything, Group=_Anything)
    _module("virgo")
    _module("virgo.mpi")
    _module("virgo.mpi.parallel_sort", parallel_sort=parallel_sort)
    _module("virgo.mpi.parallel_hdf5", MultiFile=_Anything, collective_read=_Anything())
    _module("virgo.mpi.util")
    _module("virgo.mpi.gather_array", gather_array=_Anything())
    _module("virgo.util")
    _module("virgo.util.partial_formatter", PartialFormatter=_Anything)
    _module("virgo.util.match", match=_Anything())
    _module("astropy")
    _module("astropy.cosmology", w0waCDM=_Anything, Cosmology=_Anything, z_at_value=_Anything())
    _module("astropy.constants").__getattr__ = lambda name: _Anything()
    _module("astropy.units").__getattr__ = lambda name: _Anything()
    if "psutil" not in sys.modules:
        try:
            import psutil  # noqa: F401
        except ImportError:
            _module("psutil", virtual_memory=lambda: types.SimpleNamespace(available=1 << 40, total=1 << 40))
    if "/root/reference" not in sys.path:
        sys.path.insert(0, "/root/reference")


class SharedArrayStandin:
    """SOAP.core.shared_array.SharedArray on one rank: ``.full`` and ``.local`` are the same array."""

    def __init__(self, local_shape, dtype, comm=None, units=None):
        self.full = unyt_array(np.zeros(local_shape, dtype=dtype))
        self.local = self.full
        self.comm = comm

    def sync(self):
        pass

    def free(self):
        pass


def shared(arr):
    s = SharedArrayStandin.__new__(SharedArrayStandin)
    s.full = unyt_array(arr)
    s.local = s.full
    s.comm = _Comm()
    return s
