#!/usr/bin/env python
"""
Class-level golden fixture on the reference's OWN test halos (BASELINE config 1): the halos are drawn by the
unmodified ``tests/dummy_halo_generator.py: DummyHaloGenerator`` of the reference (its numpy RNG stream, seed 4251 as
in tests/test_SO_properties.py:26), each is processed by the reference's ``process_single_halo`` + ``SharedMesh`` with
the unmodified SubhaloProperties / SOProperties / Exclusive- and InclusiveSphereProperties /
ProjectedApertureProperties, and inputs + outputs are stored in ``halo_refgen.npz``.

    python tests/golden/make_golden_refgen.py          (build container only: needs /root/reference)

Like make_golden_classes.py this runs under the dimension-blind stand-ins of ref_standin.py, so one consistent unit
system is needed: the generator's lengths (box 100, halo radii ~0.1) are multiplied by LSCALE = 40 so that halo radii
are a few units (apertures of 1 and 3 "kpc" cut through them, everything stays below the 20-"Mpc" guard of
SO_properties.py:150); masses and velocities are the generator's.  The generator's halos have no particles beyond
their own extent, so spherical-overdensity radii exist only for thresholds well above 200 x critical: 20 000 / 50 000.
The inputs cannot be regenerated without the reference, so they are stored with the outputs.
"""

import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, HERE)
sys.path.insert(0, ROOT)

import make_golden_classes as mgc  # noqa: E402
import ref_standin as rs  # noqa: E402

LSCALE = 40.0
N_HALOS = 22
NPART = [1, 10, 100, 1000, 10000]  # tests/test_SO_properties.py draws from the same list
MAX_BIG = 2  # halos of 10 000 particles kept (450 KB of input each); further ones are drawn and skipped
SO_LIST = [(20000.0, "crit", "basic"), (50000.0, "crit", "basic"), (50000.0, "mean", "general")]
AP_LIST = [(1.0, False, "basic"), (1.0, True, "basic"), (3.0, False, "general"), (3.0, True, "basic")]
PROJ_LIST = [(1.0, "basic"), (3.0, "basic")]
FILTER_LIMIT = 100
PT = {"PartType0": 0, "PartType1": 1, "PartType4": 4, "PartType5": 5}


def cosmology_params(g):
    """oracle / device parameters for the fixture's unit system (stored in the fixture)"""
    return dict(boxsize=float(g["cosmo/BOX"]), critical_density=float(g["cosmo/CRIT"]),
                mean_density=float(g["cosmo/CRIT"]) * float(g["cosmo/OMEGA_M"]), softening=float(g["cosmo/SOFT"]), G=1.0,
                H=float(g["cosmo/H_INT"]), kpc_per_length=1.0, r_20mpc=20.0, phys_mpc_to_coord=1.0, nu_density=0.0)


def fixture_halo(g, i):
    """(data, H) of fixture halo i in the layout of soap_b200.synth chunks"""
    data = {}
    for t in (0, 1, 4, 5):
        if f"in/{i}/{t}/Coordinates" in g.files:
            data[t] = {k: g[f"in/{i}/{t}/{k}"] for k in ("Coordinates", "Masses", "Velocities", "GroupNr_bound", "FOFGroupIDs")}
    H = {k: g[f"halo/{k}"][i : i + 1] for k in ("cofp", "index", "is_central", "nr_bound_part", "search_radius", "read_radius")}
    return data, H


def main():
    if not os.path.isdir("/root/reference"):
        sys.exit("make_golden_refgen.py needs /root/reference (build container only)")
    rs.install()
    sys.path.insert(0, "/root/reference/tests")
    import dummy_halo_generator as dhg
    from oracle import halo as oh
    from tests import _compare as cmp

    gen = dhg.DummyHaloGenerator(4251)
    cg = gen.get_cell_grid()
    cosmo = dict(BOX=float(np.asarray(cg.boxsize).ravel()[0]) * LSCALE, CRIT=float(cg.critical_density) / LSCALE**3,
                 OMEGA_M=float(cg.mean_density) / float(cg.critical_density), SOFT=float(cg.dark_matter_softening) * LSCALE,
                 H_INT=float(cg.cosmology["H [internal units]"]) / LSCALE)
    out = {f"cosmo/{k}": v for k, v in cosmo.items()}
    cp = cosmology_params({k: np.asarray(v) for k, v in out.items()})
    halos = []
    n_big = 0
    while len(halos) < N_HALOS:
        ih, data, rmax, Mtot, Npart, pn = gen.get_random_halo(NPART)
        d = {}
        for pt, t in PT.items():
            if pt not in data or len(data[pt]["Coordinates"]) == 0:
                continue
            mname = "DynamicalMasses" if pt == "PartType5" else "Masses"
            d[t] = dict(Coordinates=np.ascontiguousarray(np.asarray(data[pt]["Coordinates"], dtype=np.float64) * LSCALE),
                        Masses=np.asarray(data[pt][mname], dtype=np.float32).copy(),
                        Velocities=np.ascontiguousarray(np.asarray(data[pt]["Velocities"], dtype=np.float32)),
                        GroupNr_bound=np.asarray(data[pt]["GroupNr_bound"], dtype=np.int32).copy(),
                        FOFGroupIDs=np.asarray(data[pt]["FOFGroupIDs"], dtype=np.int32).copy())
        if not d:
            continue
        if sum(len(x["Masses"]) for x in d.values()) > 5000:
            if n_big >= MAX_BIG:
                continue
            n_big += 1
        idx = int(ih["index"])
        nb = sum(int((x["GroupNr_bound"] == idx).sum()) for x in d.values())
        r = float(rmax) * LSCALE
        halos.append((d, dict(cofp=np.asarray(ih["cofp"], dtype=np.float64) * LSCALE, index=idx, is_central=int(ih["is_central"]),
                              nr_bound_part=nb, search_radius=max(0.6 * r, 0.05), read_radius=max(2.5 * r, 1.0))))
    H = {k: np.array([h[k] for _, h in halos]) for k in halos[0][1]}
    H["index"] = H["index"].astype(np.int64)
    H["is_central"] = H["is_central"].astype(np.int32)
    H["nr_bound_part"] = H["nr_bound_part"].astype(np.int64)
    for k, v in H.items():
        out[f"halo/{k}"] = v
    for i, (d, _) in enumerate(halos):
        for t, x in d.items():
            for k, v in x.items():
                out[f"in/{i}/{t}/{k}"] = v

    # which internal keys the oracle restates per class (dry run on the largest halo)
    so_o = [(t, v) for v, t, _ in SO_LIST]
    aps_o = [(kpc, kpc * 1e-3, incl) for kpc, incl, _ in AP_LIST]
    proj_o = [(kpc, kpc * 1e-3) for kpc, _ in PROJ_LIST]
    big = int(np.argmax(H["nr_bound_part"]))
    Hb = {k: v[big : big + 1] for k, v in H.items()}
    Hb["read_radius"] = Hb["read_radius"] * 4.0
    out_o, props_o = cmp.run_oracle(halos[big][0], Hb, cp, so_o, aps_o, faithful=True, projected=proj_o, mesh_resolution=4)
    assert out_o[0][0] is not None, "the dry-run halo must finish"
    keys = {"sub": set(), "so": set(), "ap": set(), "proj": set()}
    for p in props_o:
        kind = {oh.SubhaloOracle: "sub", oh.SOOracle: "so", oh.ApertureOracle: "ap", oh.ProjectedApertureOracle: "proj"}[type(p)]
        for gname, blk in out_o[0][0].items():
            if (kind == "proj" and gname.startswith(p.group_name)) or gname == p.group_name:
                keys[kind] |= set(blk)
    want = lambda names: {n: "basic" for n in names}  # noqa: E731
    filters_cfg = {"general": {"limit": FILTER_LIMIT, "combine_properties": "sum",
                               "properties": ["BoundSubhalo/NumberOfGasParticles", "BoundSubhalo/NumberOfDarkMatterParticles",
                                              "BoundSubhalo/NumberOfStarParticles", "BoundSubhalo/NumberOfBlackHoleParticles"]}}
    cellgrid, props = mgc.build_reference(filters_cfg, SO_LIST, AP_LIST, PROJ_LIST, want(keys["sub"]), want(keys["so"]),
                                          want(keys["ap"]), want(keys["proj"]), cosmo=cosmo)
    from SOAP.core import halo_tasks
    from SOAP.core.shared_mesh import SharedMesh

    ua = rs.unyt_array
    crit, mean, box = cosmo["CRIT"], cosmo["CRIT"] * cosmo["OMEGA_M"], cosmo["BOX"]
    target = None
    for hp in props:
        for mult, dens in ((hp.mean_density_multiple, mean), (hp.critical_density_multiple, crit)):
            if mult is not None and (target is None or mult * dens < target):
                target = mult * dens
    out["target_density"] = target
    out["config/so"] = np.array([f"{t}:{v}:{f}" for v, t, f in SO_LIST])
    out["config/ap"] = np.array([f"{k}:{int(i)}:{f}" for k, i, f in AP_LIST])
    out["config/proj"] = np.array([f"{k}:{f}" for k, f in PROJ_LIST])
    out["config/filter_general_limit"] = FILTER_LIMIT
    n_h = len(halos)
    done = np.zeros(n_h, dtype=np.int32)
    sr_out = np.zeros(n_h)
    n_loop = np.zeros(n_h, dtype=np.int64)
    vals = {}
    for i, (d, _) in enumerate(halos):
        ref_data = {}
        for t, x in d.items():
            ref_data[f"PartType{t}"] = {k: rs.shared(v.copy()) for k, v in x.items()}
            if t == 5:
                ref_data["PartType5"]["DynamicalMasses"] = ref_data["PartType5"]["Masses"]
        mesh = {pt: SharedMesh(rs._Comm(), ref_data[pt]["Coordinates"], 4) for pt in ref_data}
        ih = {"cofp": ua(H["cofp"][i].copy()), "index": ua(H["index"][i]), "is_central": ua(H["is_central"][i]),
              "nr_bound_part": ua(H["nr_bound_part"][i]), "search_radius": ua(H["search_radius"][i]),
              "read_radius": ua(H["read_radius"][i]), "n_loop": ua(0)}
        td = ua(target) if ih["is_central"] == 1 else None
        try:
            res = halo_tasks.process_single_halo(mesh, None, ref_data, props, ua(crit), ua(mean), ua(box), ih, td)
        except AttributeError as e:  # SO_properties.py:457 (see make_golden_classes.py)
            assert "SO_r" in str(e), e
            done[i] = -1
            continue
        if isinstance(res, tuple):
            res = res[0]
        done[i] = int(res is not None)
        sr_out[i] = float(ih["search_radius"])
        if res is None:
            continue
        n_loop[i] = int(np.asarray(res["InputHalos/n_loop"][0])) if "InputHalos/n_loop" in res else 0
        for hp in props:
            for name, prop in hp.property_list.items():
                gnames = [hp.group_name]
                if hp.__class__.__name__ == "ProjectedApertureProperties":
                    gnames = [f"{hp.group_name}/proj{ax}" for ax in "xyz"]
                for gname in gnames:
                    key = f"{gname}/{prop.name}"
                    if key in res:
                        v = np.atleast_1d(np.asarray(res[key][0]))
                        arr = vals.setdefault(f"{gname}/{name}", np.zeros((n_h,) + v.shape, dtype=v.dtype))
                        arr[i] = v
    out["done"] = done
    out["search_radius_out"] = sr_out
    out["n_loop"] = n_loop
    for k, v in vals.items():
        out[f"val/{k}"] = v
    path = os.path.join(HERE, "halo_refgen.npz")
    np.savez_compressed(path, **out)
    print(f"{path}: {len(out)} arrays, {os.path.getsize(path) / 1024:.1f} KiB; done {int((done == 1).sum())}, "
          f"need a larger read radius {int((done == 0).sum())}, abort in the reference {int((done == -1).sum())}; "
          f"bound particles {H['nr_bound_part'].tolist()}")


if __name__ == "__main__":
    main()
