"""chunk_tasks.cell_cover: the read mask of a chunk (what SOAP/core/mask_cells.py:6-38 marks halo by halo),
rasterised for all halos at once, against the halo-by-halo loop."""

import numpy as np

from soap_b200 import chunk_tasks as ct


def _loop_cover(cofp, rr, L, n):
    cs = L / n
    out = np.zeros((n, n, n), dtype=bool)
    for c, r in zip(cofp, rr):
        lo = np.floor((c - r) / cs).astype(int)
        hi = np.floor((c + r) / cs).astype(int)
        idx = [np.arange(lo[d], hi[d] + 1) % n if hi[d] - lo[d] + 1 < n else np.arange(n) for d in range(3)]
        out[np.ix_(*idx)] = True
    return out


def test_cell_cover_matches_the_per_halo_loop():
    rng = np.random.default_rng(3)
    L, n = 50.0, 16
    for n_halo in (0, 1, 40):
        cofp = rng.uniform(0.0, L, size=(n_halo, 3))
        rr = rng.uniform(0.1, 6.0, size=n_halo)
        if n_halo == 40:
            cofp[0] = [0.2, L - 0.3, 25.0]  # cubes that wrap around the box edge
            rr[0] = 4.0
            rr[1] = 30.0  # spans whole axes
        got = ct.cell_cover(cofp, rr, L, n)
        assert got.dtype == bool and got.shape == (n, n, n)
        assert np.array_equal(got, _loop_cover(cofp, rr, L, n))


def test_ghost_mask_keeps_every_particle_of_every_read_sphere():
    rng = np.random.default_rng(4)
    L = 40.0
    pos = rng.uniform(0.0, L, size=(20000, 3))
    cofp = rng.uniform(0.0, L, size=(25, 3))
    rr = rng.uniform(0.5, 5.0, size=25)
    keep = ct.ghost_mask(pos, cofp, rr, L, cells_per_dim=32)
    d = np.abs(pos[None, :, :] - cofp[:, None, :])
    d = np.minimum(d, L - d)
    inside = ((d**2).sum(axis=2) <= (rr**2)[:, None]).any(axis=0)
    assert np.all(keep[inside]) and keep.sum() < len(pos)
    # positions outside [0, L) (chunks are box-wrapped around their reference position) hit the same cells
    keep2 = ct.ghost_mask(pos - L * (pos[:, :1] > 30.0), cofp, rr, L, cells_per_dim=32)
    assert np.array_equal(keep, keep2)
