"""
Device-side drop-in for SOAP's SharedMesh (SOAP/core/shared_mesh.py:10-200).

Same constructor arguments, attributes (``empty, pos_min, pos_max, resolution,
cell_size, cell_count, cell_offset, sort_idx``) and ``query_radius_periodic``
signature as the reference class; the cell-list build and the sphere query run
as CUDA kernels behind the C ABI (``soap_mesh_build`` / ``soap_sphere_query``).
``comm`` is accepted and ignored: one process drives one GPU, so the
collective build (Allreduce / Reduce / parallel_sort, shared_mesh.py:50-114)
collapses to a single-rank build.

``pos`` may be a numpy array, a torch tensor, or any object with a ``.full``
attribute holding one of those (the reference's SharedArray).  unyt arrays
are accepted through ``.value`` / ``.ndarray_view()``.
"""

import ctypes as C

import numpy as np

from . import _lib


def box_wrap(pos, ref_pos, boxsize, handle=None):
    """SOAP/core/chunk_tasks.py:48-50 on the device, in place on a CUDA tensor."""
    import torch

    assert pos.is_cuda and pos.dtype == torch.float64 and pos.is_contiguous()
    h = handle or _lib.default_handle(pos.device.index or 0)
    ref = (C.c_double * 3)(*[float(x) for x in ref_pos])
    _lib.check(
        _lib.lib().soap_box_wrap(
            h.ptr, C.c_void_p(pos.data_ptr()), pos.shape[0], ref, float(boxsize), _lib.cur_stream_ptr(pos.device)
        )
    )
    return pos


def mesh_resolution(nr_parts):
    """SOAP/core/chunk_tasks.py:299-302."""
    return int(max(1, min(256, int((nr_parts / 1000.0) ** (1.0 / 3.0)))))


def _strip(x):
    """unyt / SharedArray -> plain ndarray or tensor."""
    if hasattr(x, "full"):
        x = x.full
    if hasattr(x, "ndarray_view"):
        x = x.ndarray_view()
    return x


def _to_device_f64(x, device):
    import torch

    x = _strip(x)
    if isinstance(x, torch.Tensor):
        return x.to(device=device, dtype=torch.float64).contiguous()
    return torch.as_tensor(np.ascontiguousarray(x, dtype=np.float64), device=device)


class SharedMesh:
    def __init__(self, comm, pos, resolution, device=0, stable=True, handle=None):
        import torch

        self.device = torch.device("cuda", device)
        self.handle = handle or _lib.default_handle(device)
        p = _strip(pos)
        n = int(p.shape[0])
        # shared_mesh.py:25-29
        if n == 0:
            self.empty = True
            return
        self.empty = False
        self._pos = _to_device_f64(p, self.device)
        self.resolution = int(resolution)
        nr_cells = self.resolution**3
        pmin = (C.c_double * 3)()
        pmax = (C.c_double * 3)()
        cs = (C.c_double * 3)()
        self.cell_idx = torch.empty(n, dtype=torch.int32, device=self.device)
        self.cell_count = torch.empty(nr_cells, dtype=torch.int64, device=self.device)
        self.cell_offset = torch.empty(nr_cells, dtype=torch.int64, device=self.device)
        self.sort_idx = torch.empty(n, dtype=torch.int64, device=self.device)
        _lib.check(
            _lib.lib().soap_mesh_build(
                self.handle.ptr,
                C.c_void_p(self._pos.data_ptr()),
                n,
                self.resolution,
                pmin,
                pmax,
                cs,
                C.c_void_p(self.cell_idx.data_ptr()),
                C.c_void_p(self.cell_count.data_ptr()),
                C.c_void_p(self.cell_offset.data_ptr()),
                C.c_void_p(self.sort_idx.data_ptr()),
                1 if stable else 0,
                _lib.cur_stream_ptr(self.device),
            )
        )
        self.pos_min = np.array(pmin[:], dtype=np.float64)
        self.pos_max = np.array(pmax[:], dtype=np.float64)
        self.cell_size = np.array(cs[:], dtype=np.float64)
        self._c = (pmin, pmax, cs)

    def free(self):
        """shared_mesh.py:116-120: drop the device arrays."""
        if not self.empty:
            self.cell_count = self.cell_offset = self.sort_idx = self.cell_idx = self._pos = None

    def query_many(self, centres, radii, boxsize, mass=None):
        """Batched query: returns (counts[int64 n_q], offsets[int64 n_q+1],
        idx[int64 total], enclosed_mass or None) as CUDA tensors."""
        import torch

        centres = _to_device_f64(centres, self.device).reshape(-1, 3)
        radii = _to_device_f64(radii, self.device).reshape(-1)
        nq = int(radii.shape[0])
        counts = torch.zeros(nq, dtype=torch.int64, device=self.device)
        if self.empty or nq == 0:
            z = torch.zeros(nq + 1, dtype=torch.int64, device=self.device)
            return counts, z, torch.zeros(0, dtype=torch.int64, device=self.device), None
        enclosed = None
        mptr = C.c_void_p(0)
        eptr = C.c_void_p(0)
        if mass is not None:
            m = _strip(mass)
            m = m if isinstance(m, torch.Tensor) else torch.as_tensor(np.ascontiguousarray(m, dtype=np.float32))
            self._mass = m.to(device=self.device, dtype=torch.float32).contiguous()
            enclosed = torch.zeros(nq, dtype=torch.float64, device=self.device)
            mptr = C.c_void_p(self._mass.data_ptr())
            eptr = C.c_void_p(enclosed.data_ptr())
        pmin, pmax, cs = self._c

        def call(offsets, idx):
            _lib.check(
                _lib.lib().soap_sphere_query(
                    self.handle.ptr,
                    C.c_void_p(self._pos.data_ptr()),
                    self._pos.shape[0],
                    self.resolution,
                    pmin,
                    pmax,
                    cs,
                    C.c_void_p(self.cell_count.data_ptr()),
                    C.c_void_p(self.cell_offset.data_ptr()),
                    C.c_void_p(self.sort_idx.data_ptr()),
                    C.c_void_p(centres.data_ptr()),
                    C.c_void_p(radii.data_ptr()),
                    nq,
                    float(boxsize),
                    C.c_void_p(counts.data_ptr()),
                    C.c_void_p(offsets.data_ptr() if offsets is not None else 0),
                    C.c_void_p(idx.data_ptr() if idx is not None else 0),
                    mptr,
                    eptr,
                    _lib.cur_stream_ptr(self.device),
                )
            )

        call(None, None)
        offsets = torch.zeros(nq + 1, dtype=torch.int64, device=self.device)
        torch.cumsum(counts, 0, out=offsets[1:])
        total = int(offsets[-1].item())
        idx = torch.empty(max(total, 1), dtype=torch.int64, device=self.device)
        if total > 0:
            call(offsets, idx)
        return counts, offsets, idx[:total], enclosed

    def query_radius_periodic(self, centre, radius, pos, boxsize):
        """shared_mesh.py:122-200: indices (numpy int64) of particles within
        ``radius`` of ``centre`` under the periodic minimum image."""
        if self.empty:
            return np.ndarray(0, dtype=int)
        c = np.asarray(_strip(centre), dtype=np.float64).reshape(1, 3)
        r = np.asarray([float(_strip(radius))], dtype=np.float64)
        _, _, idx, _ = self.query_many(c, r, float(_strip(boxsize)))
        return idx.cpu().numpy()
