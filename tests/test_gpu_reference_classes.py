"""GPU path against the reference's own classes: the chunk and configuration of the class-level golden
fixture (tests/golden/halo_classes.npz, produced by running the UNMODIFIED SubhaloProperties / SOProperties /
Exclusive- and InclusiveSphereProperties / ProjectedApertureProperties through the reference's
process_single_halo): four SO variations, four spheres, two projected apertures, a "general" category filter
(limit 100) on two variations, the EncloseRadius shortcut, satellites, halos across the periodic edge, two
halos whose read radius is too small."""

import numpy as np
import pytest

from tests import _compare as cmp
from tests.test_oracle_golden import _class_fixture

pytestmark = pytest.mark.gpu


def _setup():
    gen, g, data, H = _class_fixture()
    cp = gen.cosmology_params()
    so_cfg = [s.split(":") for s in g["config/so"]]
    ap_cfg = [s.split(":") for s in g["config/ap"]]
    pj_cfg = [s.split(":") for s in g["config/proj"]]
    kw = dict(
        so=[(t, 177.65 if t == "BN98" else float(v)) for t, v, _ in so_cfg],
        apertures=[(float(k), float(k) * 1e-3, int(i)) for k, i, _ in ap_cfg],
        projected=[(float(k), float(k) * 1e-3) for k, _ in pj_cfg],
        filters={"general": (int(g["config/filter_general_limit"]), (0, 1, 4, 5))},
        so_filters=[f for _, _, f in so_cfg], ap_filters=[f for _, _, f in ap_cfg], proj_filters=[f for _, f in pj_cfg],
        skip_gt=("exclusive", "inclusive"),
    )
    return g, data, H, cp, kw


def test_gpu_matches_reference_class_fixture():
    from soap_b200.halo_tasks import DeviceChunk, process_halos

    g, data, H, cp, kw = _setup()
    flags = 1 | 2 | 4 | 8
    cfg = cmp.device_config(cp, flags=flags, dmo=False, **kw)
    chunk = DeviceChunk(data, cp["boxsize"])
    res = process_halos(chunk, cfg, H)
    st = res.status.cpu().numpy()
    done = g["done"]
    # the two halos with a too small read radius come back with status 1 and the reference's new search radius
    for i in np.flatnonzero(done == 0):
        assert st[i] == 1
        assert res.get("InputHalos/search_radius")[i] == float(g["search_radius_out"][i])
    ok = done == 1
    assert np.all(st[ok] == 0)
    assert np.array_equal(res.get("InputHalos/n_loop")[ok], g["n_loop"][ok])  # the ladder, rung for rung
    # straight against the reference's numbers: counts exact, float32 outputs to a few ulp (masses, radii) ...
    so_names = ["SO/200_crit", "SO/500_crit", "SO/200_mean", "SO/BN98"]
    checks = 0
    for q, gname in enumerate(so_names):
        for ref_key, dev_key, tol in (("Ndm", "Ndm", 0), ("Ngas", "Ngas", 0), ("Nstar", "Nstar", 0), ("r", "r", 1e-6),
                                      ("Mtot", "Mso", 1e-6), ("Mdm", "Mdm", 1e-6), ("Mstar", "Mstar", 1e-6)):
            ref = g[f"val/{gname}/{ref_key}"][:, 0][ok].astype(np.float64)
            got = res.get(f"SO/{q}/{dev_key}")[ok]
            if tol == 0:
                assert np.array_equal(got, ref), (gname, ref_key)
            else:
                assert np.all(np.abs(got - ref) <= tol * np.maximum(np.abs(ref), 1e-30) + 0.0), (gname, ref_key, got, ref)
            checks += ref.size
    # filtered variations are exact zeros for halos below the limit, like in the reference
    nb = sum(res.get(f"BoundSubhalo/{k}") for k in ("Ngas", "Ndm", "Nstar", "Nbh"))
    small = ok & (nb < 100)
    assert small.sum() > 3
    assert np.all(res.get("SO/2/Mso")[small] == 0.0) and np.all(g["val/SO/200_mean/Mtot"][:, 0][small] == 0.0)
    # ... and everything else through the oracle, which reproduces the fixture (tests/test_oracle_golden.py)
    for faithful in (False, True):
        oracle_out, props = cmp.run_oracle(data, H, cp, kw["so"], kw["apertures"], faithful=faithful,
                                           projected=kw["projected"], filters=kw["filters"], so_filters=kw["so_filters"],
                                           ap_filters=kw["ap_filters"], proj_filters=kw["proj_filters"],
                                           skip_gt=kw["skip_gt"], halos=[i for i in range(len(done)) if done[i] >= 0])
        rep = cmp.compare(res, oracle_out, props, cp, halos=[i for i in range(len(done)) if done[i] >= 0], flags=flags,
                          faithful=faithful)
        print("max errors vs %s oracle:" % ("faithful" if faithful else "float64"),
              {k: float(f"{v:.3g}") for k, v in sorted(rep.maxerr.items())})
        rep.assert_ok()
    assert checks > 400
    chunk.free()
