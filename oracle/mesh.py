"""
Oracle for stage A (cell-list build) and stage B1 (periodic sphere query).

Test infrastructure only -- see oracle/__init__.py.

Restates, line by line, without MPI and unyt:
  * box_wrap                 SOAP/core/chunk_tasks.py:48-50
  * mesh resolution rule     SOAP/core/chunk_tasks.py:299-302
  * SharedMesh.__init__      SOAP/core/shared_mesh.py:11-114
  * query_radius_periodic    SOAP/core/shared_mesh.py:122-200

The within-cell order of ``sort_idx`` is defined in the reference by VirgoDC's
``parallel_sort(return_index=True)`` which is not in /root/reference
(virgodc>=1.0.3, requirements.txt:9): parity on that order is unpinned; the
oracle uses a stable sort (ascending particle index within a cell).
"""

import numpy as np


def box_wrap(pos, ref_pos, boxsize):
    """SOAP/core/chunk_tasks.py:48-50 (numpy floored modulo)."""
    shift = ref_pos[None, :] - 0.5 * boxsize
    return (pos - shift) % boxsize + shift


def mesh_resolution(nr_parts):
    """SOAP/core/chunk_tasks.py:299-302: ~1000 particles per cell, 1..256."""
    return int(max(1, min(256, int((nr_parts / 1000.0) ** (1.0 / 3.0)))))


class MeshOracle:
    """SOAP/core/shared_mesh.py:8-200 on a single rank, raw float64 arrays."""

    def __init__(self, pos, resolution):
        pos = np.ascontiguousarray(pos, dtype=np.float64)
        self.pos = pos
        # shared_mesh.py:25-29
        if pos.shape[0] == 0:
            self.empty = True
            return
        self.empty = False
        # shared_mesh.py:35-52 (single rank: the Allreduce is the identity)
        self.pos_min = np.amin(pos, axis=0) * 1
        self.pos_max = np.amax(pos, axis=0) * 1
        # shared_mesh.py:56-58
        for i in range(3):
            if self.pos_min[i] == self.pos_max[i]:
                self.pos_max[i] = self.pos_min[i] + 1.0
        # shared_mesh.py:64-66
        self.resolution = int(resolution)
        nr_cells = self.resolution**3
        self.cell_size = (self.pos_max - self.pos_min) / self.resolution
        # shared_mesh.py:69-77
        cell_idx = np.floor(
            (pos - self.pos_min[None, :]) / self.cell_size[None, :]
        ).astype(np.int32)
        cell_idx = np.clip(cell_idx, 0, self.resolution - 1)
        cell_idx = (
            cell_idx[:, 0]
            + self.resolution * cell_idx[:, 1]
            + (self.resolution**2) * cell_idx[:, 2]
        )
        self.cell_idx = cell_idx
        # shared_mesh.py:80-93
        self.cell_count = np.bincount(cell_idx, minlength=nr_cells)
        # shared_mesh.py:96-102
        self.cell_offset = np.zeros_like(self.cell_count)
        if nr_cells > 1:
            self.cell_offset[1:] = np.cumsum(self.cell_count[:-1])
        # shared_mesh.py:105-114 (stable order chosen, see module docstring)
        self.sort_idx = np.argsort(cell_idx, kind="stable")

    def query_radius_periodic(self, centre, radius, pos, boxsize):
        """SOAP/core/shared_mesh.py:122-200."""
        if self.empty:
            return np.ndarray(0, dtype=int)

        def periodic_distance_squared(pos, centre):
            # shared_mesh.py:138-142
            dr = pos - centre[None, :]
            dr[dr > 0.5 * boxsize] -= boxsize
            dr[dr < -0.5 * boxsize] += boxsize
            return np.sum(dr**2, axis=1)

        # shared_mesh.py:146-179.  The reference holds the cell coordinates in
        # python sets; their iteration order is unspecified, we use ascending.
        cell_coords = [set() for _ in range(3)]
        for dim in (0, 1, 2):
            min_copy_nr = 0
            while (
                centre[dim] + (min_copy_nr - 1) * boxsize + radius
                >= self.pos_min[dim]
            ):
                min_copy_nr -= 1
            max_copy_nr = 0
            while (
                centre[dim] + (max_copy_nr + 1) * boxsize - radius
                <= self.pos_max[dim]
            ):
                max_copy_nr += 1
            for copy_nr in range(min_copy_nr, max_copy_nr + 1, 1):
                min_coord = max(
                    self.pos_min[dim], centre[dim] + copy_nr * boxsize - radius
                )
                min_idx = int(
                    np.floor((min_coord - self.pos_min[dim]) / self.cell_size[dim])
                )
                max_coord = min(
                    self.pos_max[dim], centre[dim] + copy_nr * boxsize + radius
                )
                max_idx = int(
                    np.floor((max_coord - self.pos_min[dim]) / self.cell_size[dim])
                )
                for cell_nr in range(min_idx, max_idx + 1):
                    if cell_nr >= 0 and cell_nr < self.resolution:
                        cell_coords[dim].add(cell_nr)

        # shared_mesh.py:181-194
        idx = []
        for k in sorted(cell_coords[2]):
            for j in sorted(cell_coords[1]):
                for i in sorted(cell_coords[0]):
                    cell_nr = i + self.resolution * j + (self.resolution**2) * k
                    start = self.cell_offset[cell_nr]
                    count = self.cell_count[cell_nr]
                    if count > 0:
                        idx_in_cell = self.sort_idx[start : start + count]
                        r2 = periodic_distance_squared(pos[idx_in_cell, :], centre)
                        keep = r2 <= radius * radius
                        if np.sum(keep) > 0:
                            idx.append(idx_in_cell[keep])
        # shared_mesh.py:197-200
        if len(idx) > 0:
            return np.concatenate(idx)
        else:
            return np.ndarray(0, dtype=int)


def brute_force_query(pos, centre, radius, boxsize):
    """tests/test_shared_mesh.py:76-80 style brute force (same arithmetic)."""
    dr = pos - centre[None, :]
    dr[dr > 0.5 * boxsize] -= boxsize
    dr[dr < -0.5 * boxsize] += boxsize
    r2 = np.sum(dr**2, axis=1)
    return np.nonzero(r2 <= radius * radius)[0]
