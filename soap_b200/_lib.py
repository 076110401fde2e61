"""
ctypes loader for libsoap_b200.so (the C ABI declared in include/soap_b200.h).

There is no CPU fallback: importing works anywhere (so the host logic can be
unit-tested), but every compute entry point raises if the shared library is
missing or no B200 is visible.
"""

import ctypes as C
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libsoap_b200.so")
CSRC = os.path.join(_HERE, "csrc")

SOAP_MAX_SO = 8
SOAP_MAX_APERTURES = 16
SOAP_MAX_PTYPES = 8
SOAP_MAX_FILTERS = 8
ABI_VERSION = 2

# per-halo status codes (include/soap_b200.h)
HALO_OK, HALO_RADIUS_TOO_SMALL, HALO_COUNT_MISMATCH, HALO_SO_NOT_FOUND, HALO_ROOT_FAILED = range(5)


class SoapError(RuntimeError):
    pass


class PtypeArrays(C.Structure):
    _fields_ = [
        ("ptype", C.c_int),
        ("ids_are_int64", C.c_int),
        ("n", C.c_int64),
        ("pos", C.c_void_p),
        ("mass", C.c_void_p),
        ("vel", C.c_void_p),
        ("grnr", C.c_void_p),
        ("fof", C.c_void_p),
    ]


class HaloConfig(C.Structure):
    _fields_ = [
        ("boxsize", C.c_double),
        ("G", C.c_double),
        ("H", C.c_double),
        ("kpc_per_length", C.c_double),
        ("r_20mpc", C.c_double),
        ("nu_density", C.c_double),
        ("phys_mpc_to_coord", C.c_double),
        ("softening", C.c_double * SOAP_MAX_PTYPES),
        ("target_density", C.c_double),
        ("do_subhalo", C.c_int),
        ("n_so", C.c_int),
        ("so_reference_density", C.c_double * SOAP_MAX_SO),
        ("so_virial", C.c_int * SOAP_MAX_SO),
        ("n_apertures", C.c_int),
        ("ap_radius", C.c_double * SOAP_MAX_APERTURES),
        ("ap_physical_mpc", C.c_double * SOAP_MAX_APERTURES),
        ("ap_inclusive", C.c_int * SOAP_MAX_APERTURES),
        ("n_projected", C.c_int),
        ("proj_radius", C.c_double * SOAP_MAX_APERTURES),
        ("proj_physical_mpc", C.c_double * SOAP_MAX_APERTURES),
        ("property_flags", C.c_uint32),
        ("dmo", C.c_int),
        ("n_filters", C.c_int),
        ("filter_limit", C.c_int64 * SOAP_MAX_FILTERS),
        ("filter_types", C.c_uint32 * SOAP_MAX_FILTERS),
        ("so_filter", C.c_int * SOAP_MAX_SO),
        ("ap_filter", C.c_int * SOAP_MAX_APERTURES),
        ("proj_filter", C.c_int * SOAP_MAX_APERTURES),
        ("ap_prev_radius", C.c_double * SOAP_MAX_APERTURES),
        ("proj_prev_radius", C.c_double * SOAP_MAX_APERTURES),
        ("debug_flags", C.c_uint32),
    ]


def build(verbose=False):
    """Compile libsoap_b200.so in-tree with nvcc for sm_100a (cross-compiles
    without a GPU)."""
    r = subprocess.run(["make", "-C", CSRC, "-j8"], capture_output=True, text=True)
    if r.returncode != 0:
        raise SoapError("building libsoap_b200.so failed:\n" + r.stdout + r.stderr)
    if verbose:
        print(r.stdout)
    return LIB_PATH


_lib = None

# name -> (restype, argtypes); every symbol include/soap_b200.h declares
SYMBOLS = {
    "soap_abi_version": (C.c_int, []),
    "soap_last_error": (C.c_char_p, []),
    "soap_create": (C.c_int, [C.c_int, C.POINTER(C.c_void_p)]),
    "soap_destroy": (C.c_int, [C.c_void_p]),
    "soap_launch_count": (C.c_int64, [C.c_void_p]),
    "soap_kernel_timing": (C.c_int, [C.c_void_p, C.c_int]),
    "soap_kernel_timings": (C.c_int64, [C.c_void_p, C.c_char_p, C.c_int64]),
    "soap_box_wrap": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.POINTER(C.c_double), C.c_double, C.c_void_p]),
    "soap_mesh_build": (
        C.c_int,
        [C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_double),
         C.POINTER(C.c_double), C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p],
    ),
    "soap_sphere_query": (
        C.c_int,
        [C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_double),
         C.POINTER(C.c_double), C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64,
         C.c_double, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p],
    ),
    "soap_chunk_create": (
        C.c_int,
        [C.c_void_p, C.POINTER(PtypeArrays), C.c_int, C.c_double, C.c_int, C.POINTER(C.c_void_p), C.c_void_p],
    ),
    "soap_chunk_destroy": (C.c_int, [C.c_void_p]),
    "soap_chunk_num_particles": (C.c_int64, [C.c_void_p]),
    "soap_result_layout": (C.c_int64, [C.POINTER(HaloConfig), C.c_char_p, C.c_int64]),
    "soap_process_halos": (
        C.c_int,
        [C.c_void_p, C.POINTER(HaloConfig), C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
         C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p],
    ),
    "soap_chunk_last_pairs": (C.c_int64, [C.c_void_p]),
    "soap_chunk_timings": (C.c_int64, [C.c_void_p, C.c_char_p, C.c_int64]),
}


def lib():
    """The loaded shared library; raises SoapError if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise SoapError(
            f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
            "(there is no CPU fallback for the soap_b200 hot path)"
        )
    L = C.CDLL(LIB_PATH)
    for name, (res, args) in SYMBOLS.items():
        fn = getattr(L, name)  # AttributeError if the ABI is incomplete
        fn.restype = res
        fn.argtypes = args
    if L.soap_abi_version() != ABI_VERSION:
        raise SoapError("libsoap_b200.so ABI version mismatch")
    _lib = L
    return L


def check(rc):
    if rc != 0:
        raise SoapError(lib().soap_last_error().decode(errors="replace"))


class Handle:
    """Per-device context (scratch workspace).  One per GPU / Python thread."""

    def __init__(self, device=0):
        import torch

        if not torch.cuda.is_available():
            raise SoapError("soap_b200 needs a CUDA device (B200, sm_100a); none is visible")
        self.device = int(device)
        self.ptr = C.c_void_p()
        check(lib().soap_create(self.device, C.byref(self.ptr)))

    def launches(self):
        return int(lib().soap_launch_count(self.ptr))

    def kernel_timing(self, on=True):
        check(lib().soap_kernel_timing(self.ptr, int(bool(on))))

    def kernel_timings(self):
        """{kernel name: (launches, total ms)} since the last call (synchronises the device)"""
        buf = C.create_string_buffer(1 << 16)
        lib().soap_kernel_timings(self.ptr, buf, len(buf))
        out = {}
        for line in buf.value.decode().strip().split("\n"):
            if line:
                name, n, ms = line.split("\t")
                out[name] = (int(n), float(ms))
        return out

    def close(self):
        if self.ptr:
            lib().soap_destroy(self.ptr)
            self.ptr = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


_handles = {}


def default_handle(device=0):
    h = _handles.get(device)
    if h is None:
        h = _handles[device] = Handle(device)
    return h


def cur_stream_ptr(device=None):
    """torch's current stream ON ``device`` (a chunk / handle on cuda:1 must not be handed the stream of cuda:0)"""
    import torch

    return C.c_void_p(torch.cuda.current_stream(device).cuda_stream)
