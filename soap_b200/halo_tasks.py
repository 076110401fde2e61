"""
Device-side halo batching: the replacement for SOAP's per-core halo loop
``process_halos`` / ``process_single_halo`` (SOAP/core/halo_tasks.py:23-430).

``DeviceChunk`` holds a chunk's particles on the GPU (what the reference keeps
in MPI shared memory after SOAP/core/chunk_tasks.py:256-288) and
``process_halos`` runs the whole radius ladder + property reductions for every
halo of the chunk in a handful of kernel launches per ladder rung, through the
C ABI (``soap_chunk_create`` / ``soap_process_halos``).

The host adapter owns everything unyt does in the reference: every threshold
(reference densities, aperture radii, softenings, G, H) is handed over already
converted to coordinate units (SURVEY.md 8(c) detail 11).
"""

import ctypes as C
from dataclasses import dataclass, field
from typing import Dict, List, Optional

import numpy as np

from . import _lib

PF_KIN, PF_KAPPA, PF_TENS, PF_HMR, PF_ITER = 1, 2, 4, 8, 16

SEARCH_RADIUS_FACTOR = 1.2  # halo_tasks.py:14
READ_RADIUS_FACTOR = 1.5  # halo_tasks.py:17


@dataclass
class HaloPropConfig:
    """The numerical content of ``halo_prop_list`` (compute_halo_properties.py:232-505)
    plus the cellgrid scalars, in coordinate units."""

    boxsize: float
    G: float
    critical_density: float
    mean_density: float
    softening: Dict[int, float] = field(default_factory=dict)  # by ptype
    H: float = 0.0
    kpc_per_length: float = 1000.0
    r_20mpc: float = 20.0
    nu_density: float = 0.0
    phys_mpc_to_coord: float = 1.0
    do_subhalo: bool = True
    # each SO: (type in {"crit","mean","BN98"}, value)
    so: List[tuple] = field(default_factory=list)
    # each aperture: (radius in coordinate units, physical radius in Mpc, inclusive)
    apertures: List[tuple] = field(default_factory=list)
    # each projected aperture: (radius in coordinate units, physical radius in Mpc)
    projected: List[tuple] = field(default_factory=list)
    property_flags: int = 0
    dmo: bool = False
    # cross-check switch (bit 0: every halo through the general kernel-sequence path); 0 in production
    debug_flags: int = 0
    # CategoryFilter (category_filter.py:69-110): name -> (limit, ptypes whose BoundSubhalo counts are summed);
    # "basic" is implicit.  so_filter / ap_filter / proj_filter give the halo_filter name of each variation,
    # aligned with ``so`` / ``apertures`` / ``projected`` as given (default "basic").
    filters: Dict[str, tuple] = field(default_factory=dict)
    so_filter: List[str] = field(default_factory=list)
    ap_filter: List[str] = field(default_factory=list)
    proj_filter: List[str] = field(default_factory=list)
    # aperture kinds ("exclusive", "inclusive") whose variations know the radii of their siblings
    # (all_radii_kpc; compute_halo_properties.py:345-395 always passes them to exclusive spheres, to inclusive
    # ones with skip_gt_enclose_radius) and therefore use the EncloseRadius shortcut (needs
    # BoundSubhalo/EncloseRadius to be enabled in the parameter file)
    skip_gt: tuple = ()
    # exact values read from the reference's property objects (dropin.py), used instead of the ones derived
    # from ``so`` when given: SOProperties.reference_density / .virial_definition and the target density of
    # halo_tasks.py:306-317
    so_rho: Optional[List[float]] = None
    so_virial_flags: Optional[List[bool]] = None
    target_density_value: Optional[float] = None

    def so_reference_density(self, i):
        """SO_properties.py:3494-3512."""
        t, val = self.so[i]
        if t == "mean":
            return val * self.mean_density
        if t in ("crit", "BN98"):
            return val * self.critical_density
        raise AttributeError(f"Unknown SO type: {t}!")

    def so_virial(self, i):
        """SO_properties.py:3463-3476."""
        t, val = self.so[i]
        return t == "BN98" or (t in ("crit", "mean") and val == 200)

    def target_density(self):
        """halo_tasks.py:306-317 with the SOProperties defaults of 1000
        (SO_properties.py:3459-3462)."""
        target = None
        for t, val in self.so:
            mean_mult = val if t == "mean" else 1000.0
            crit_mult = val if t in ("crit", "BN98") else 1000.0
            for d in (mean_mult * self.mean_density, crit_mult * self.critical_density):
                if target is None or d < target:
                    target = d
        return target

    def to_c(self):
        c = _lib.HaloConfig()
        c.boxsize = self.boxsize
        c.G = self.G
        c.H = self.H
        c.kpc_per_length = self.kpc_per_length
        c.r_20mpc = self.r_20mpc
        c.nu_density = self.nu_density
        c.phys_mpc_to_coord = self.phys_mpc_to_coord
        for pt in range(_lib.SOAP_MAX_PTYPES):
            c.softening[pt] = float(self.softening.get(pt, 0.0))
        td = self.target_density() if self.target_density_value is None else self.target_density_value
        c.target_density = -1.0 if td is None else float(td)
        c.do_subhalo = int(self.do_subhalo)
        if len(self.so) > _lib.SOAP_MAX_SO or len(self.apertures) > _lib.SOAP_MAX_APERTURES:
            raise ValueError("too many SO / aperture variations")
        c.n_so = len(self.so)
        for i in range(len(self.so)):
            c.so_reference_density[i] = float(self.so_reference_density(i) if self.so_rho is None else self.so_rho[i])
            c.so_virial[i] = int(self.so_virial(i) if self.so_virial_flags is None else self.so_virial_flags[i])
        # filters: index 0 is "basic"
        fnames = ["basic"] + [n for n in self.filters if n != "basic"]
        if len(fnames) > _lib.SOAP_MAX_FILTERS:
            raise ValueError("too many category filters")
        c.n_filters = len(fnames)
        code = {0: 0, 1: 1, 4: 2, 5: 3}
        for f, name in enumerate(fnames[1:], start=1):
            limit, ptypes = self.filters[name]
            c.filter_limit[f] = int(limit)
            c.filter_types[f] = sum(1 << code[int(str(t)[-1])] for t in ptypes)

        def fidx(lst, i):
            name = lst[i] if i < len(lst) else "basic"
            if name not in fnames:
                raise KeyError(f'filter "{name}" is not defined')
            return fnames.index(name)

        for i in range(len(self.so)):
            c.so_filter[i] = fidx(self.so_filter, i)
        order = sorted(range(len(self.apertures)), key=lambda i: (self.apertures[i][0], self.apertures[i][2]))
        aps = [self.apertures[i] for i in order]
        self._sorted_apertures = aps
        c.n_apertures = len(aps)
        for i in range(_lib.SOAP_MAX_APERTURES):
            c.ap_prev_radius[i] = -1.0
            c.proj_prev_radius[i] = -1.0
        prev = {}
        for i, (r, mpc, incl) in enumerate(aps):
            c.ap_radius[i] = float(r)
            c.ap_physical_mpc[i] = float(mpc)
            c.ap_inclusive[i] = int(bool(incl))
            c.ap_filter[i] = fidx(self.ap_filter, order[i])
            kind = "inclusive" if incl else "exclusive"
            if kind in self.skip_gt and kind in prev:
                c.ap_prev_radius[i] = float(prev[kind])
            prev[kind] = r
        porder = sorted(range(len(self.projected)), key=lambda i: self.projected[i][0])
        pj = [self.projected[i] for i in porder]
        if len(pj) > _lib.SOAP_MAX_APERTURES:
            raise ValueError("too many projected aperture variations")
        self._sorted_projected = pj
        c.n_projected = len(pj)
        for i, (r, mpc) in enumerate(pj):
            c.proj_radius[i] = float(r)
            c.proj_physical_mpc[i] = float(mpc)
            c.proj_filter[i] = fidx(self.proj_filter, porder[i])
        c.property_flags = int(self.property_flags)
        c.dmo = int(self.dmo)
        c.debug_flags = int(self.debug_flags)
        return c


def result_layout(cfg_c):
    """name -> (column offset, width) of the result table (soap_result_layout)."""
    buf = C.create_string_buffer(1 << 16)
    ncol = _lib.lib().soap_result_layout(C.byref(cfg_c), buf, len(buf))
    if ncol < 0:
        raise _lib.SoapError(_lib.lib().soap_last_error().decode())
    cols = {}
    off = 0
    for line in buf.value.decode().strip().split("\n"):
        name, w = line.rsplit(":", 1)
        cols[name] = (off, int(w))
        off += int(w)
    assert off == ncol
    return int(ncol), cols


def _dev(x, dtype, device):
    import torch

    if hasattr(x, "full"):
        x = x.full
    if hasattr(x, "ndarray_view"):
        x = x.ndarray_view()
    if isinstance(x, torch.Tensor):
        return x.to(device=device, dtype=dtype, non_blocking=True).contiguous()
    return torch.as_tensor(np.ascontiguousarray(x), device=device).to(dtype).contiguous()


class DeviceChunk:
    """A chunk's particles resident in HBM, binned and reordered for the halo
    kernels.  ``data[ptype]`` needs Coordinates (f64 [N,3], already box-wrapped),
    Masses (the mass_dataset of the ptype), Velocities, GroupNr_bound,
    FOFGroupIDs; ptype is 0/1/4/5 or "PartType0"..."""

    def __init__(self, data, boxsize, device=0, fine_ppc=0, handle=None):
        import torch

        self.device = torch.device("cuda", device)
        self.handle = handle or _lib.default_handle(device)
        self.boxsize = float(boxsize)
        arr = (_lib.PtypeArrays * 4)()
        self._keep = []
        n_types = 0
        for key in sorted(data, key=lambda k: int(str(k)[-1])):
            d = data[key]
            pt = int(str(key)[-1])
            if pt not in (0, 1, 4, 5):
                continue
            mass_name = "DynamicalMasses" if (pt == 5 and "DynamicalMasses" in d) else "Masses"
            pos = _dev(d["Coordinates"], torch.float64, self.device)
            n = int(pos.shape[0])
            if n == 0:
                continue
            mass = _dev(d[mass_name], torch.float32, self.device)
            vel = _dev(d["Velocities"], torch.float32, self.device)
            g = d["GroupNr_bound"]
            gdt = torch.int64 if "64" in str(getattr(g, "dtype", "int32")) else torch.int32
            grnr = _dev(g, gdt, self.device)
            fof = _dev(d["FOFGroupIDs"], gdt, self.device)
            self._keep += [pos, mass, vel, grnr, fof]
            a = arr[n_types]
            a.ptype = pt
            a.ids_are_int64 = int(gdt == torch.int64)
            a.n = n
            a.pos = pos.data_ptr()
            a.mass = mass.data_ptr()
            a.vel = vel.data_ptr()
            a.grnr = grnr.data_ptr()
            a.fof = fof.data_ptr()
            n_types += 1
        if n_types == 0:
            raise ValueError("DeviceChunk: no particles")
        self.ptr = C.c_void_p()
        _lib.check(
            _lib.lib().soap_chunk_create(
                self.handle.ptr, arr, n_types, self.boxsize, int(fine_ppc), C.byref(self.ptr), _lib.cur_stream_ptr(self.device)
            )
        )
        # the SoA copy lives in the chunk; the staging tensors can go
        self._keep = []
        self.n = int(_lib.lib().soap_chunk_num_particles(self.ptr))

    def timings(self):
        buf = C.create_string_buffer(1 << 14)
        _lib.lib().soap_chunk_timings(self.ptr, buf, len(buf))
        out = {}
        for line in buf.value.decode().strip().split("\n"):
            if ":" in line:
                k, v = line.rsplit(":", 1)
                out[k] = float(v)
        return out

    def last_pairs(self):
        return int(_lib.lib().soap_chunk_last_pairs(self.ptr))

    def free(self):
        if getattr(self, "ptr", None):
            _lib.lib().soap_chunk_destroy(self.ptr)
            self.ptr = C.c_void_p()

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class ChunkFeed:
    """Asynchronous chunk feed: uploads the next chunk's particle and halo arrays
    from pinned host memory on a copy stream while the current chunk is being
    processed (the reference reads chunk k+1's cells while chunk k is in
    ``process_halos`` only across nodes; within a node this is SURVEY.md 8(f) rank 3:
    SOAP/core/chunk_tasks.py:194-288).  Host tensors must already have the device
    dtypes (float64 positions, float32 masses / velocities, int32 or int64 ids)."""

    def __init__(self, device=0):
        import torch

        self.device = torch.device("cuda", device)
        self.stream = torch.cuda.Stream(self.device)

    @staticmethod
    def _host_buffers(arrays):
        out = []
        for v in arrays.values() if isinstance(arrays, dict) else arrays:
            if isinstance(v, dict):
                out += ChunkFeed._host_buffers(v)
            elif hasattr(v, "full"):  # SharedArray
                out.append(v.full)
            else:
                out.append(v)
        return out

    def register(self, arrays):
        """Page-lock host buffers in place (cudaHostRegister) so that ``upload`` runs at the PCIe rate.  SOAP's
        particle arrays live in MPI shared windows (SOAP/core/shared_array.py:33), i.e. pageable memory that
        ``tensor.pin_memory()`` would have to copy; registration pins the window itself, once per buffer.
        ``arrays``: (nested) dict or list of numpy arrays / CPU tensors / SharedArrays.  Returns a token for
        ``unregister``."""
        import numpy as np
        import torch

        rt = torch.cuda.cudart()
        token = []
        for a in self._host_buffers(arrays):
            if isinstance(a, torch.Tensor):
                if a.is_cuda or a.is_pinned() or a.numel() == 0:
                    continue
                ptr, nbytes = a.data_ptr(), a.numel() * a.element_size()
            else:
                a = np.asarray(a)
                if a.size == 0 or not a.flags["C_CONTIGUOUS"]:
                    continue
                ptr, nbytes = a.ctypes.data, a.nbytes
            err = rt.cudaHostRegister(ptr, nbytes, 0)
            if int(err) != 0:
                for p in token:
                    rt.cudaHostUnregister(p)
                raise _lib.SoapError(f"cudaHostRegister of {nbytes} bytes failed with error {int(err)}")
            token.append(ptr)
        return token

    def unregister(self, token):
        import torch

        rt = torch.cuda.cudart()
        for p in token:
            rt.cudaHostUnregister(p)
        token.clear()

    def upload(self, data_host, halos_host):
        """Start the upload; returns a ticket for ``wait``."""
        import torch

        with torch.cuda.stream(self.stream):
            data = {pt: {k: v.to(self.device, non_blocking=True) for k, v in d.items()} for pt, d in data_host.items()}
            halos = {k: v.to(self.device, non_blocking=True) for k, v in halos_host.items()}
            ev = torch.cuda.Event()
            ev.record(self.stream)
        return data, halos, ev

    def wait(self, ticket):
        """Make the current stream wait for the upload; returns (data, halos) on the device."""
        import torch

        data, halos, ev = ticket
        cur = torch.cuda.current_stream(self.device)
        cur.wait_event(ev)
        for d in list(data.values()) + [halos]:
            for t in d.values():
                t.record_stream(cur)
        return data, halos


class HaloResults:
    """Result table of ``process_halos``: one float64 row per halo."""

    def __init__(self, table, status, cols, cfg):
        self.table = table  # torch [H, ncol] on device
        self.status = status  # torch int32 [H]
        self.cols = cols
        self.cfg = cfg
        self._host = None

    def host(self):
        if self._host is None:
            self._host = self.table.cpu().numpy()
        return self._host

    def get(self, name):
        off, w = self.cols[name]
        t = self.host()
        return t[:, off] if w == 1 else t[:, off : off + w]

    def names(self):
        return list(self.cols)


def process_halos(chunk: DeviceChunk, cfg: HaloPropConfig, halo_arrays, out=None):
    """Batched halo_tasks.process_halos.  ``halo_arrays`` is the reference's dict
    of per-halo arrays: cofp [H,3], search_radius, read_radius, index,
    is_central, nr_bound_part (numpy or torch).  Returns HaloResults; halos with
    status 1 carry the updated search/read radii in InputHalos/search_radius
    and InputHalos/read_radius (halo_tasks.py:390-402)."""
    import torch

    dev = chunk.device
    cfg_c = cfg.to_c()
    ncol, cols = result_layout(cfg_c)
    cofp = _dev(halo_arrays["cofp"], torch.float64, dev)
    H = int(cofp.shape[0])
    sr = _dev(halo_arrays["search_radius"], torch.float64, dev)
    rr = _dev(halo_arrays["read_radius"], torch.float64, dev)
    index = _dev(halo_arrays["index"], torch.int64, dev)
    cen = _dev(halo_arrays["is_central"], torch.int32, dev)
    nb = _dev(halo_arrays["nr_bound_part"], torch.int64, dev)
    table = out if out is not None else torch.empty((H, ncol), dtype=torch.float64, device=dev)
    status = torch.empty(H, dtype=torch.int32, device=dev)
    _lib.check(
        _lib.lib().soap_process_halos(
            chunk.ptr,
            C.byref(cfg_c),
            H,
            C.c_void_p(cofp.data_ptr()),
            C.c_void_p(sr.data_ptr()),
            C.c_void_p(rr.data_ptr()),
            C.c_void_p(index.data_ptr()),
            C.c_void_p(cen.data_ptr()),
            C.c_void_p(nb.data_ptr()),
            C.c_void_p(table.data_ptr()),
            ncol,
            C.c_void_p(status.data_ptr()),
            _lib.cur_stream_ptr(dev),
        )
    )
    return HaloResults(table, status, cols, cfg)
