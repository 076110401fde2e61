"""GPU parity on the edge cases the reference's tests and code paths name
(SURVEY.md Appendix B): missing particle types, int64 membership ids, duplicate
radii, particles exactly at the centre, one-particle halos, halos with every
particle below / above the SO threshold, halos across the periodic corner,
empty inputs."""

import numpy as np
import pytest

from soap_b200 import synth
from tests import _compare as cmp

pytestmark = pytest.mark.gpu

SO4 = [("crit", 200.0), ("mean", 200.0), ("crit", 500.0), ("BN98", float(synth.virBN98()))]


def _run(data, H, cp, so, apertures, flags, dmo, projected=()):
    from soap_b200.halo_tasks import DeviceChunk, process_halos

    cfg = cmp.device_config(cp, so=so, apertures=apertures, flags=flags, dmo=dmo, projected=projected)
    chunk = DeviceChunk(data, cp["boxsize"])
    res = process_halos(chunk, cfg, H)
    oracle_out, props = cmp.run_oracle(data, H, cp, so, apertures, faithful=False, projected=projected)
    rep = cmp.compare(res, oracle_out, props, cp, flags=flags)
    print("max errors:", {k: float(f"{v:.3g}") for k, v in sorted(rep.maxerr.items())})
    rep.assert_ok()
    return res, oracle_out


def test_missing_particle_types_and_int64_ids():
    """no stars, no black holes (mesh.empty for those types: shared_mesh.py:25-29) and
    int64 GroupNr_bound / FOFGroupIDs as real membership files have them"""
    L = 20.0
    cp = synth.coordinate_unit_params(L)
    data, H = synth.dummy_chunk(911, 25, boxsize=L, n_background=60000, npart_choices=(1, 10, 100, 1000))
    data = {t: d for t, d in data.items() if t in (0, 1)}
    for d in data.values():
        d["GroupNr_bound"] = d["GroupNr_bound"].astype(np.int64)
        d["FOFGroupIDs"] = d["FOFGroupIDs"].astype(np.int64)
    # halos lost their stars / black holes: recount the bound particles
    for i, idx in enumerate(H["index"]):
        H["nr_bound_part"][i] = sum(int((d["GroupNr_bound"] == idx).sum()) for d in data.values())
    aps = [(kpc * 1e-3 * cp["phys_mpc_to_coord"], kpc * 1e-3, incl) for kpc in (50.0,) for incl in (0, 1)]
    _run(data, H, cp, SO4[:2], aps, flags=1 | 2 | 4 | 8, dmo=False)


def test_int64_ids_outside_int32_are_an_error():
    """membership is an integer-exact contract (halo_tasks.py:121-123): a 64-bit group id that does not fit the
    int32 the device keeps must be refused, not wrapped onto another halo's id"""
    from soap_b200 import _lib
    from soap_b200.halo_tasks import DeviceChunk

    L = 20.0
    data, H = synth.dummy_chunk(912, 5, boxsize=L, n_background=2000, npart_choices=(10, 100))
    data = {t: d for t, d in data.items() if t in (0, 1)}
    for d in data.values():
        d["GroupNr_bound"] = d["GroupNr_bound"].astype(np.int64)
        d["FOFGroupIDs"] = d["FOFGroupIDs"].astype(np.int64)
    data[1]["GroupNr_bound"][3] = (1 << 32) + 7  # would alias halo 7 after narrowing
    with pytest.raises(_lib.SoapError, match="32 bits"):
        DeviceChunk(data, L)


def _handmade_chunk(L):
    """A few dark matter halos built by hand in an otherwise thin uniform background."""
    rng = np.random.default_rng(42)
    pos, mass, grnr, fof, halos = [], [], [], [], []

    def add(centre, offsets, m, hid, central=1, search=None):
        offsets = np.asarray(offsets, dtype=np.float64)
        p = (np.asarray(centre) + offsets) % L
        pos.append(p)
        mass.append(np.asarray(m, dtype=np.float32) * np.ones(len(p), dtype=np.float32))
        grnr.append(np.full(len(p), hid, dtype=np.int32))
        fof.append(np.full(len(p), hid, dtype=np.int32))
        rmax = np.sqrt((offsets**2).sum(axis=1)).max()
        halos.append((np.asarray(centre, dtype=np.float64), search or max(1.01 * rmax, 0.01), central, len(p), hid))

    # 1: duplicate radii (SO_properties.py:186): six particles on the axes at the same distance + a core
    r = 0.05
    add([5.0, 5.0, 5.0], [[0, 0, 0], [r, 0, 0], [-r, 0, 0], [0, r, 0], [0, -r, 0], [0, 0, r], [0, 0, -r],
                          [2 * r, 0, 0], [0, 2 * r, 0]], [3.0, 1, 1, 1, 1, 1, 1, 0.5, 0.5], 11)
    # 2: two particles exactly at the centre (nskip = argmax(r > 0): SO_properties.py:416)
    add([9.0, 3.0, 7.0], np.vstack([np.zeros((2, 3)), rng.normal(size=(40, 3)) * 0.03]), 1.0, 12)
    # 3: one-particle halo (tests/test_aperture_properties.py:101)
    add([2.0, 8.0, 1.0], [[0.0, 0.0, 0.0]], 1.0, 13)
    # 4: diffuse halo, every particle below the thresholds (SO_properties.py:157-177)
    add([14.0, 14.0, 4.0], rng.normal(size=(30, 3)) * 0.9, 0.002, 14, search=3.0)
    # 5: compact massive halo across the periodic corner
    add([L - 1e-3, 1e-3, L - 2e-3], rng.normal(size=(3000, 3)) * 0.04, 0.5, 15)
    # 6: satellite (no SO properties: SO_properties.py:3627)
    add([11.0, 6.0, 12.0], rng.normal(size=(200, 3)) * 0.02, 0.3, 16, central=0)
    nb = 40000
    pos.append(rng.random((nb, 3)) * L)
    mass.append(np.full(nb, 0.02, dtype=np.float32))
    grnr.append(np.full(nb, -1, dtype=np.int32))
    fof.append(np.full(nb, -1, dtype=np.int32))
    n = sum(len(p) for p in pos)
    data = {1: dict(Coordinates=np.concatenate(pos), Masses=np.concatenate(mass),
                    Velocities=(1000.0 * (rng.random((n, 3)) - 0.5)).astype(np.float32),
                    GroupNr_bound=np.concatenate(grnr), FOFGroupIDs=np.concatenate(fof))}
    H = {
        "cofp": np.array([h[0] for h in halos]),
        "search_radius": np.array([h[1] for h in halos]),
        "read_radius": np.array([max(h[1], 5.0) for h in halos]),
        "is_central": np.array([h[2] for h in halos], dtype=np.int32),
        "nr_bound_part": np.array([h[3] for h in halos], dtype=np.int64),
        "index": np.array([h[4] for h in halos], dtype=np.int64),
    }
    return data, H


def test_handmade_profiles():
    L = 16.0
    cp = synth.coordinate_unit_params(L)
    data, H = _handmade_chunk(L)
    res, oracle_out = _run(data, H, cp, SO4, [], flags=8, dmo=True)
    st = res.status.cpu().numpy()
    assert (st == 0).sum() >= 5
    # the satellite has no SO, the one-particle halo no Vmax radius beyond its single point
    sat = list(H["index"]).index(16)
    assert res.get("SO/0/r")[sat] == 0.0 and res.get("SO/0/Ndm")[sat] == 0.0


def test_handmade_profiles_with_projected_and_tensors():
    L = 16.0
    cp = synth.coordinate_unit_params(L)
    data, H = _handmade_chunk(L)
    pj = [(kpc * 1e-3 * cp["phys_mpc_to_coord"], kpc * 1e-3) for kpc in (30.0, 100.0)]
    aps = [(0.1 * cp["phys_mpc_to_coord"], 0.1, 1)]
    _run(data, H, cp, SO4[:1], aps, flags=1 | 4 | 8, dmo=True, projected=pj)


def test_empty_inputs():
    import torch

    from soap_b200 import _lib
    from soap_b200.halo_tasks import DeviceChunk, process_halos

    L = 10.0
    cp = synth.coordinate_unit_params(L)
    with pytest.raises(ValueError):
        DeviceChunk({1: dict(Coordinates=np.zeros((0, 3)), Masses=np.zeros(0, np.float32),
                             Velocities=np.zeros((0, 3), np.float32), GroupNr_bound=np.zeros(0, np.int32),
                             FOFGroupIDs=np.zeros(0, np.int32))}, L)
    rng = np.random.default_rng(1)
    data = {1: dict(Coordinates=rng.random((1000, 3)) * L, Masses=np.ones(1000, np.float32),
                    Velocities=np.zeros((1000, 3), np.float32), GroupNr_bound=np.full(1000, -1, np.int32),
                    FOFGroupIDs=np.full(1000, -1, np.int32))}
    chunk = DeviceChunk(data, L)
    H = {"cofp": np.zeros((0, 3)), "search_radius": np.zeros(0), "read_radius": np.zeros(0),
         "index": np.zeros(0, np.int64), "is_central": np.zeros(0, np.int32), "nr_bound_part": np.zeros(0, np.int64)}
    res = process_halos(chunk, cmp.device_config(cp, so=SO4[:1], dmo=True), H)
    assert res.table.shape[0] == 0 and res.status.shape[0] == 0
    # a halo in an empty corner of the box: no particle in its sphere, status comes from the reference's rules
    H1 = {"cofp": np.array([[5.0, 5.0, 5.0]]), "search_radius": np.array([1e-6]), "read_radius": np.array([1e-6]),
          "index": np.array([7], np.int64), "is_central": np.array([1], np.int32), "nr_bound_part": np.array([0], np.int64)}
    res = process_halos(chunk, cmp.device_config(cp, so=[], dmo=True), H1)
    assert int(res.status.cpu()[0]) == _lib.HALO_OK and res.get("BoundSubhalo/Ndm")[0] == 0.0
    # bad configuration: error string through the ABI, no crash
    with pytest.raises(_lib.SoapError):
        process_halos(chunk, cmp.device_config(cp, so=SO4[:1], projected=[(0.1, 0.1)], do_subhalo=False, dmo=True), H1)
    assert torch.cuda.is_available()
