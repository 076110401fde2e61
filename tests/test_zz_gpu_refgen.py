"""GPU path on the reference's OWN test halos (BASELINE config 1): tests/golden/halo_refgen.npz holds halos drawn by
the reference's unmodified DummyHaloGenerator and what its process_single_halo + property classes returned for them
(tests/golden/make_golden_refgen.py).  Every halo is its own chunk, as in the reference's tests."""

import importlib.util
import os
import sys

import numpy as np
import pytest

from tests import _compare as cmp

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def test_gpu_matches_reference_on_its_own_fixture_halos():
    from soap_b200.halo_tasks import DeviceChunk, process_halos
    from tests.test_oracle_golden import _fixture_config

    sys.path.insert(0, GOLD)
    spec = importlib.util.spec_from_file_location("make_golden_refgen", os.path.join(GOLD, "make_golden_refgen.py"))
    gen = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(gen)
    g = np.load(os.path.join(GOLD, "halo_refgen.npz"))
    cp = gen.cosmology_params(g)
    so_cfg, ap_cfg, pj_cfg, so, aps, proj, filters = _fixture_config(g)
    kw = dict(filters=filters, so_filters=[f for _, _, f in so_cfg], ap_filters=[f for _, _, f in ap_cfg],
              proj_filters=[f for _, f in pj_cfg], skip_gt=("exclusive", "inclusive"))
    flags = 1 | 2 | 4 | 8
    cfg = cmp.device_config(cp, so=so, apertures=aps, projected=proj, flags=flags, dmo=False, **kw)
    done = g["done"]
    rep = cmp.Report()
    n_ok = n_small = 0
    for i in range(len(done)):
        if done[i] == -1:
            continue  # the reference itself aborts on this halo (SO_properties.py:457)
        data, H = gen.fixture_halo(g, i)
        chunk = DeviceChunk(data, cp["boxsize"])
        res = process_halos(chunk, cfg, H)
        st = int(res.status.cpu().numpy()[0])
        if done[i] == 0:  # read radius too small: status 1 and the reference's new search radius
            assert st == 1, (i, st)
            assert res.get("InputHalos/search_radius")[0] == float(g["search_radius_out"][i]), i
            n_small += 1
        else:
            assert st == 0, (i, st)
            assert int(res.get("InputHalos/n_loop")[0]) == int(g["n_loop"][i]), i
            # straight against the reference's numbers where names coincide
            for q, (t, v, _) in enumerate(so_cfg):
                ref_r = float(g[f"val/SO/{float(v):.0f}_{t}/r"][i, 0])
                got_r = float(res.get(f"SO/{q}/r")[0])
                assert abs(got_r - ref_r) <= 1e-6 * max(abs(ref_r), 1e-30), (i, q, got_r, ref_r)
            n_ok += 1
        # ... and everything else through the oracle, which reproduces this fixture (tests/test_oracle_golden.py)
        oracle_out, props = cmp.run_oracle(data, H, cp, so, aps, projected=proj, mesh_resolution=4, **kw)
        cmp.compare(res, oracle_out, props, cp, halos=[0], flags=flags, rep=rep)
        chunk.free()
    print("reference fixture halos on the GPU: done", n_ok, "too small", n_small,
          "max errors", {k: float(f"{v:.3g}") for k, v in sorted(rep.maxerr.items())})
    rep.assert_ok()
    assert n_ok >= 10 and n_small >= 1
