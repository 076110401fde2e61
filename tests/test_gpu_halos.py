"""GPU parity of the batched halo path (stage B+C) against the oracle."""

import numpy as np
import pytest

from soap_b200 import synth
from tests import _compare as cmp

pytestmark = pytest.mark.gpu

SO4 = [("crit", 200.0), ("mean", 200.0), ("crit", 500.0), ("BN98", float(synth.virBN98()))]


def _run(data, H, cp, so, apertures, flags, dmo, fine_ppc=0, halos=None, projected=()):
    from soap_b200.halo_tasks import DeviceChunk, process_halos

    cfg = cmp.device_config(cp, so=so, apertures=apertures, flags=flags, dmo=dmo, projected=projected)
    chunk = DeviceChunk(data, cp["boxsize"], fine_ppc=fine_ppc)
    res = process_halos(chunk, cfg, H)
    # against the float64 oracle (all sums in float64) and against the reference-faithful one (float32 where
    # the reference sums in float32); the observed maxima are printed
    rep = None
    for faithful in (False, True):
        oracle_out, props = cmp.run_oracle(data, H, cp, so, apertures, faithful=faithful, halos=halos,
                                           projected=projected, iterative=bool(flags & 16))
        r = cmp.compare(res, oracle_out, props, cp, halos=halos, flags=flags, faithful=faithful)
        print("max errors vs %s oracle:" % ("faithful" if faithful else "float64"),
              {k: float(f"{v:.3g}") for k, v in sorted(r.maxerr.items())})
        r.assert_ok()
        rep = rep or r
    print("timings:", chunk.timings())
    return res, rep


def test_dmo_nfw_chunk_so_and_subhalo():
    L = 40.0
    cp = synth.coordinate_unit_params(L)
    data, H = synth.to_numpy(*synth.nfw_chunk(400000, 400, L, seed=1, max_np=20000))
    _run(data, H, cp, SO4, [], flags=cmp_flags(hmr=True), dmo=True)


def test_dmo_nfw_chunk_fine_mesh_independent():
    """membership and results must not depend on the internal mesh resolution"""
    L = 30.0
    cp = synth.coordinate_unit_params(L)
    data, H = synth.to_numpy(*synth.nfw_chunk(150000, 150, L, seed=3, max_np=30000))
    for ppc in (1, 64, 100000):
        _run(data, H, cp, SO4[:2], [], flags=0, dmo=True, fine_ppc=ppc)


def test_dummy_chunk_all_types_apertures():
    """reference-fixture-like halos (all particle types, satellites, unbound
    particles, halos across the periodic edge) with SO + exclusive/inclusive
    apertures + kinematics + tensors + half-mass radii"""
    L = 20.0
    cp = synth.coordinate_unit_params(L)
    data, H = synth.dummy_chunk(4251, 40, boxsize=L, n_background=200000,
                                npart_choices=(1, 10, 100, 1000, 10000))
    aps = []
    for kpc in (30.0, 100.0):
        for incl in (0, 1):
            aps.append((kpc * 1e-3 * cp["phys_mpc_to_coord"], kpc * 1e-3, incl))
    _run(data, H, cp, SO4[:3], aps, flags=1 | 4 | 8, dmo=False)


def test_dummy_chunk_kappa_corot_and_disc_fractions():
    """kappa_corot_{gas,star,baryons} and DtoT{gas,star} (kinematic_properties.py:266-425)
    for BoundSubhalo and exclusive / inclusive apertures"""
    L = 20.0
    cp = synth.coordinate_unit_params(L)
    data, H = synth.dummy_chunk(3256, 30, boxsize=L, n_background=50000,
                                npart_choices=(1, 10, 100, 1000, 5000))
    aps = [(kpc * 1e-3 * cp["phys_mpc_to_coord"], kpc * 1e-3, incl) for kpc in (30.0, 100.0) for incl in (0, 1)]
    _run(data, H, cp, SO4[:1], aps, flags=1 | 2, dmo=False)


def test_iterative_inertia_tensors_hydro():
    """Total/StellarInertiaTensor[Reduced] with the reference's default 20 re-selection passes
    (inertia_tensors.py:19-132) for BoundSubhalo, SO (in-sphere + surrounding particles) and
    exclusive / inclusive apertures"""
    L = 20.0
    cp = synth.coordinate_unit_params(L)
    data, H = synth.dummy_chunk(977, 30, boxsize=L, n_background=50000,
                                npart_choices=(10, 100, 1000, 5000))
    aps = [(kpc * 1e-3 * cp["phys_mpc_to_coord"], kpc * 1e-3, incl) for kpc in (30.0, 100.0) for incl in (0, 1)]
    res, rep = _run(data, H, cp, SO4[:2], aps, flags=1 | 4 | 8 | 16, dmo=False)
    t = res.get("BoundSubhalo/TotalInertiaTensor")
    assert (np.abs(t).sum(axis=1) > 0).sum() >= 10  # the iterative tensors were really computed


def test_iterative_projected_inertia_tensors():
    """ProjectedTotalInertiaTensor[Reduced] (inertia_tensors.py:226-343, 20 passes) per axis"""
    L = 20.0
    cp = synth.coordinate_unit_params(L)
    data, H = synth.dummy_chunk(431, 30, boxsize=L, n_background=50000,
                                npart_choices=(10, 100, 1000, 5000))
    pj = [(kpc * 1e-3 * cp["phys_mpc_to_coord"], kpc * 1e-3) for kpc in (30.0, 100.0)]
    aps = [(0.05 * cp["phys_mpc_to_coord"], 0.05, 0)]
    res, rep = _run(data, H, cp, SO4[:1], aps, flags=4 | 8 | 16, dmo=False, projected=pj)
    t = res.get("ProjectedAperture/1/projz/ProjectedTotalInertiaTensor")
    assert (np.abs(t).sum(axis=1) > 0).sum() >= 10


def test_iterative_inertia_tensors_dmo():
    L = 40.0
    cp = synth.coordinate_unit_params(L)
    data, H = synth.to_numpy(*synth.nfw_chunk(300000, 200, L, seed=5, max_np=20000))
    res, rep = _run(data, H, cp, SO4[:2], [], flags=4 | 16, dmo=True)
    t = res.get("SO/0/TotalInertiaTensorReduced")
    assert (np.abs(t).sum(axis=1) > 0).sum() >= 50


def test_dummy_chunk_projected_apertures():
    """ProjectedAperture/{R}/proj{x,y,z}: masses, counts, com, vcom, 1-D velocity
    dispersions and projected non-iterative inertia tensors, several radii at once"""
    L = 20.0
    cp = synth.coordinate_unit_params(L)
    data, H = synth.dummy_chunk(127, 30, boxsize=L, n_background=50000,
                                npart_choices=(1, 10, 100, 1000, 5000))
    pj = [(kpc * 1e-3 * cp["phys_mpc_to_coord"], kpc * 1e-3) for kpc in (10.0, 30.0, 50.0, 100.0)]
    _run(data, H, cp, SO4[:1], [], flags=8, dmo=False, projected=pj)


def test_read_radius_too_small_status():
    """halos that cannot reach the target density inside read_radius come back
    with status 1 and the reference's updated radii (halo_tasks.py:166-181)"""
    L = 20.0
    cp = synth.coordinate_unit_params(L)
    data, H = synth.dummy_chunk(77, 12, boxsize=L, n_background=0, npart_choices=(100, 1000))
    H["read_radius"][:] = np.maximum(H["search_radius"], 0.3)
    _run(data, H, cp, SO4[:1], [], flags=0, dmo=False)


def test_reread_loop_finishes_every_halo_like_one_generous_read():
    """run_chunks(reread=True): halos flagged status 1 are re-read with the returned radii
    (chunk_tasks.py:188-367) and end with the rows a generous read radius gives directly"""
    from soap_b200 import chunk_tasks as ct
    from soap_b200.halo_tasks import DeviceChunk, process_halos

    L = 20.0
    cp = synth.coordinate_unit_params(L)
    data, H = synth.dummy_chunk(77, 12, boxsize=L, n_background=20000, npart_choices=(100, 1000))
    cfg = cmp.device_config(cp, so=SO4[:2], flags=8, dmo=False)
    passes = []

    def compute(cd, hc):
        passes.append(len(hc["index"]))
        return process_halos(DeviceChunk(cd, L), cfg, hc).table

    big = dict(H)
    big["read_radius"] = np.full(len(H["index"]), 8.0)
    ref, iref = ct.run_chunks(data, big, L, 2, compute)
    passes.clear()
    small = dict(H)
    small["read_radius"] = np.maximum(H["search_radius"], 0.3)
    got, igot = ct.run_chunks(data, small, L, 2, compute, reread=True)
    assert np.array_equal(iref.cpu().numpy(), igot.cpu().numpy())
    ref, got = ref.cpu().numpy(), got.cpu().numpy()
    assert (ref[:, 0] == 0).all() and (got[:, 0] == 0).all()
    assert len(passes) > 2 and passes[-1] < passes[0]  # later passes only repeat the unfinished halos
    # same property columns; the InputHalos columns (rungs walked, accepted radius, pairs) depend on
    # where the ladder was clamped by read_radius, in the reference too (halo_tasks.py:166-181)
    cols = list(range(6, ref.shape[1]))
    np.testing.assert_allclose(got[:, cols], ref[:, cols], rtol=1e-12, atol=1e-12)


def cmp_flags(kin=False, tens=False, hmr=False):
    return (1 if kin else 0) | (4 if tens else 0) | (8 if hmr else 0)
