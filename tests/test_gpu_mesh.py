"""GPU parity of stage A/B against the oracle (ports of the reference's
tests/test_shared_mesh.py:95-221 scenario families + bit-exact mesh arrays)."""

import numpy as np
import pytest

from oracle import mesh as om

pytestmark = pytest.mark.gpu


def _check_mesh(pos, res):
    import torch
    from soap_b200.shared_mesh import SharedMesh

    o = om.MeshOracle(pos, res)
    m = SharedMesh(None, pos, res)
    assert np.array_equal(m.pos_min, o.pos_min) and np.array_equal(m.pos_max, o.pos_max)
    assert np.array_equal(m.cell_size, o.cell_size)
    assert np.array_equal(m.cell_idx.cpu().numpy(), o.cell_idx)
    assert np.array_equal(m.cell_count.cpu().numpy(), o.cell_count)
    assert np.array_equal(m.cell_offset.cpu().numpy(), o.cell_offset)
    assert np.array_equal(m.sort_idx.cpu().numpy(), o.sort_idx)  # stable order
    return m, o


@pytest.mark.parametrize("res", [1, 2, 4, 8, 16, 32])
@pytest.mark.parametrize("n", [1000, 1, 20000])
def test_mesh_build_bit_exact(res, n):
    rng = np.random.default_rng(res * 100 + n)
    pos = rng.random((n, 3))
    _check_mesh(pos, res)


def test_mesh_box_wrap_bit_exact():
    import torch
    from soap_b200.shared_mesh import box_wrap

    rng = np.random.default_rng(5)
    L = 7.3
    pos = (rng.random((5000, 3)) - 0.5) * 40.0
    pos[0] = [-1e-20, 1e-20, L]  # floored-mod edge: tiny negative wraps to L
    ref = np.array([1.3, 6.9, 3.3])
    want = om.box_wrap(pos, ref, L)
    got = box_wrap(torch.as_tensor(pos, device="cuda").clone(), ref, L).cpu().numpy()
    assert np.array_equal(got, want)


def _query_all(m, o, pos, centres, radii, L):
    for c, r in zip(centres, radii):
        got = m.query_radius_periodic(c, r, None, L)
        want = o.query_radius_periodic(c, r, pos, L)
        brute = om.brute_force_query(pos, c, r, L)
        assert len(np.unique(got)) == len(got)
        assert np.array_equal(np.sort(got), np.sort(want))
        assert np.array_equal(np.sort(got), brute)
        # same order as the oracle (cells ascending k, j, i; stable within cell)
        assert np.array_equal(got, want)


@pytest.mark.parametrize("res", [1, 2, 4, 8, 16, 32])
def test_query_box_filling(res):
    rng = np.random.default_rng(res)
    L = 1.0
    pos = rng.random((1000, 3))
    m, o = _check_mesh(pos, res)
    centres = rng.random((25, 3)) * 1.4 - 0.2
    radii = rng.random(25) * 0.6 * rng.choice([0.01, 0.1, 1.0], 25)
    _query_all(m, o, pos, centres, radii, L)


@pytest.mark.parametrize("res", [1, 4, 16])
def test_query_subregion_wrapped(res):
    rng = np.random.default_rng(100 + res)
    L = 1.0
    for _ in range(5):
        corner = rng.random(3)
        size = 0.1 + 0.4 * rng.random(3)
        pos = (corner + size * rng.random((1000, 3))) % L  # may straddle the box edge
        ref = (corner + 0.5 * size) % L
        pos = om.box_wrap(pos, ref, L)
        m, o = _check_mesh(pos, res)
        centres = (corner + size * rng.random((10, 3))) % L
        radii = 0.3 * rng.random(10)
        _query_all(m, o, pos, centres, radii, L)


def test_query_single_particle_and_batch():
    import torch
    from soap_b200.shared_mesh import SharedMesh

    pos = np.array([[0.5, 0.5, 0.5]])
    m, o = _check_mesh(pos, 4)
    _query_all(m, o, pos, np.array([[0.5, 0.5, 0.5], [0.9, 0.5, 0.5]]), np.array([0.1, 0.1]), 1.0)
    assert SharedMesh(None, np.zeros((0, 3)), 4).empty
    # batched API with enclosed mass
    rng = np.random.default_rng(9)
    pos = rng.random((5000, 3))
    mass = rng.random(5000).astype(np.float32)
    m = SharedMesh(None, pos, 8)
    c = rng.random((40, 3))
    r = 0.2 * rng.random(40)
    counts, offsets, idx, enc = m.query_many(c, r, 1.0, mass=mass)
    counts, offsets, idx, enc = counts.cpu().numpy(), offsets.cpu().numpy(), idx.cpu().numpy(), enc.cpu().numpy()
    for q in range(40):
        b = om.brute_force_query(pos, c[q], r[q], 1.0)
        assert counts[q] == len(b)
        assert np.array_equal(np.sort(idx[offsets[q]:offsets[q + 1]]), b)
        assert abs(enc[q] - mass[b].astype(np.float64).sum()) <= 1e-12 * max(1.0, enc[q])


def test_mesh_and_queries_match_reference_golden():
    """CUDA stage A/B against fixtures produced by executing the reference's own
    SharedMesh source (tests/golden/make_golden.py): cell arrays bit-exact, query
    results equal as index sets (the within-cell order is unpinned upstream)."""
    import os

    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "shared_mesh.npz"))
    from soap_b200.shared_mesh import SharedMesh

    L = float(g["mesh_L"])
    for i in range(int(g["mesh_n"])):
        pos, res = g[f"mesh{i}_pos"], int(g[f"mesh{i}_res"])
        m = SharedMesh(None, pos, res)
        assert np.array_equal(m.pos_min, g[f"mesh{i}_pos_min"])
        assert np.array_equal(m.pos_max, g[f"mesh{i}_pos_max"])
        assert np.array_equal(m.cell_size, g[f"mesh{i}_cell_size"])
        assert np.array_equal(m.cell_count.cpu().numpy(), g[f"mesh{i}_cell_count"])
        assert np.array_equal(m.cell_offset.cpu().numpy(), g[f"mesh{i}_cell_offset"])
        assert np.array_equal(m.sort_idx.cpu().numpy(), g[f"mesh{i}_sort_idx"])  # stable == the generator's stand-in sort
        for q in range(len(g[f"mesh{i}_radii"])):
            got = m.query_radius_periodic(g[f"mesh{i}_centres"][q], g[f"mesh{i}_radii"][q], None, L)
            assert np.array_equal(np.sort(got), g[f"mesh{i}_q{q}"]), f"mesh {i} query {q}"
