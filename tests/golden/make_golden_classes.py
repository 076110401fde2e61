#!/usr/bin/env python
"""
Class-level golden fixtures: run the UNMODIFIED reference classes -- SubhaloProperties,
SOProperties, ExclusiveSphereProperties / InclusiveSphereProperties, ProjectedApertureProperties --
through the reference's own ``process_single_halo`` (SOAP/core/halo_tasks.py:23-273) and SharedMesh
on synthetic halos, and store inputs + every output the oracle restates.

    python tests/golden/make_golden_classes.py          (build container only: needs /root/reference)

The reference package is imported as it is from /root/reference; the third-party modules it needs and
that are absent here (unyt, mpi4py, h5py, VirgoDC, astropy) are replaced by the stand-ins of
``ref_standin.py``.  The unyt stand-in is dimension-blind (every conversion is x 1.0), so the inputs
are built in one consistent unit system: lengths in kpc, a = 1.  ``tests/test_oracle_golden.py``
then requires ``oracle/halo.py`` (run with the same numbers and every unit factor = 1) to reproduce
these outputs, which pins the oracle's class-level restatement -- selections, ``<`` / ``<=``, type
order, dtypes of the sums, the search-radius ladder with its SearchRadiusTooSmallError protocol,
filtered halos -> exact zeros -- against the reference's code.
"""

import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, HERE)
sys.path.insert(0, ROOT)

import ref_standin as rs  # noqa: E402

# one consistent unit system (kpc, 1e10 Msun, km/s), a = 1
BOX = 400.0
CRIT = 0.005  # critical density, mass / kpc^3
OMEGA_M = 0.3
SOFT = 1.2
H_INT = 0.07  # km/s/kpc
G_INT = 1.0  # the stand-in physical constants are 1.0 (newton_G)


def cosmology_params():
    """what tests/_compare.oracle_params / device_config need to run the same numbers"""
    return dict(boxsize=BOX, critical_density=CRIT, mean_density=CRIT * OMEGA_M, softening=SOFT, G=G_INT, H=H_INT,
                kpc_per_length=1.0, r_20mpc=20.0, phys_mpc_to_coord=1.0, nu_density=0.0)


def make_chunk(seed, n_halos, n_background):
    """Fixture-like halos (soap_b200.synth.dummy_chunk: exponential radii, four particle types, bound /
    unbound / foreign members, satellites, halos across the periodic edge) scaled to kpc, on a uniform
    background so that SO radii exist."""
    from soap_b200 import synth

    data, H = synth.dummy_chunk(seed, n_halos, npart_choices=(1, 10, 100, 1000, 3000), boxsize=100.0,
                                n_background=n_background, background_mass=0.02)
    s = BOX / 100.0
    for d in data.values():
        d["Coordinates"] = d["Coordinates"] * s
    H["cofp"] = H["cofp"] * s
    H["search_radius"] = H["search_radius"] * s
    H["read_radius"] = np.maximum(H["search_radius"], 120.0)
    # two halos whose read radius is too small: process_single_halo returns None and updates search_radius
    H["read_radius"][[3, 7]] = H["search_radius"][[3, 7]]
    return data, H


def build_reference(filters_cfg, so_list, ap_list, proj_list, subhalo_props, so_props, ap_props, proj_props, cosmo=None):
    """cosmo: optional overrides of the module's unit system {BOX, CRIT, OMEGA_M, SOFT, H_INT}"""
    c_ = dict(BOX=BOX, CRIT=CRIT, OMEGA_M=OMEGA_M, SOFT=SOFT, H_INT=H_INT)
    c_.update(cosmo or {})
    rs.install()
    import SOAP.core.shared_array as sa

    sa.SharedArray = rs.SharedArrayStandin
    from SOAP.core.category_filter import CategoryFilter
    from SOAP.core.parameter_file import ParameterFile
    from SOAP.particle_selection.aperture_properties import ExclusiveSphereProperties, InclusiveSphereProperties
    from SOAP.particle_selection.projected_aperture_properties import ProjectedApertureProperties
    from SOAP.particle_selection.SO_properties import SOProperties
    from SOAP.particle_selection.subhalo_properties import SubhaloProperties

    ua = rs.unyt_array

    class Datasets:
        def get_dataset(self, name, data):
            ptype, dset = name.split("/")
            return data[ptype][dset]

        def get_column_index(self, name, column):
            raise KeyError(name)

        def get_defined_constant(self, name):
            raise KeyError(name)

    cellgrid = types.SimpleNamespace(
        snap_unit_registry=None, critical_density=ua(c_["CRIT"]), mean_density=ua(c_["CRIT"] * c_["OMEGA_M"]), a=1.0,
        a_unit=rs._Units(), z=0.0, boxsize=ua(c_["BOX"]), baryon_softening=ua(c_["SOFT"]),
        dark_matter_softening=ua(c_["SOFT"]), nu_softening=ua(c_["SOFT"]), observer_position=ua([0.5 * c_["BOX"]] * 3),
        snapshot_datasets=Datasets(),
        cosmology={"Omega_nu_0": 0.0, "H0 [internal units]": c_["H_INT"], "H [internal units]": c_["H_INT"], "Omega_g": 0.0,
                   "Omega_m": c_["OMEGA_M"]}, virBN98=177.65, get_unit=lambda name: rs._Units(),
    )

    def enabled(cls, wanted):
        out = {}
        for name, prop in cls.property_list.items():
            if name in wanted or prop.name in wanted:
                out[prop.name] = wanted.get(name, wanted.get(prop.name))
        return out

    pdict = {
        "calculations": {"calculate_missing_properties": False, "strict_halo_copy": False},
        "filters": filters_cfg,
        "SubhaloProperties": {"properties": enabled(SubhaloProperties, subhalo_props)},
        "SOProperties": {"properties": enabled(SOProperties, so_props)},
        "ApertureProperties": {"properties": enabled(ExclusiveSphereProperties, ap_props)},
        "ProjectedApertureProperties": {"properties": enabled(ProjectedApertureProperties, proj_props)},
    }
    parameters = ParameterFile(parameter_dictionary=pdict)
    cat = CategoryFilter(filters_cfg, dmo=False)
    cat.get_filter_metadata = lambda name: {"Masked": name != "basic"}
    gas_filter = rs._Anything()
    props = [SubhaloProperties(cellgrid, parameters, gas_filter, rs._Anything(), cat)]
    for val, typ, flt in so_list:
        props.append(SOProperties(cellgrid, parameters, gas_filter, cat, flt, val, typ))
    for kpc, incl, flt in ap_list:
        cls = InclusiveSphereProperties if incl else ExclusiveSphereProperties
        props.append(cls(cellgrid, parameters, kpc, None, gas_filter, rs._Anything(), rs._Anything(), cat, flt, sorted({k for k, _, _ in ap_list})))
    for kpc, flt in proj_list:
        props.append(ProjectedApertureProperties(cellgrid, parameters, kpc, None, cat, flt, sorted({k for k, _ in proj_list})))
    return cellgrid, props


def main():
    if not os.path.isdir("/root/reference"):
        sys.exit("make_golden_classes.py needs /root/reference (build container only)")
    rs.install()
    from oracle import halo as oh

    # the internal names the oracle restates (collected from a dry run below), per class
    want = lambda names, flt="basic": {n: flt for n in names}  # noqa: E731
    from tests import _compare as cmp

    cp = cosmology_params()
    params = cmp.oracle_params(cp, faithful=True)
    so_list = [(200.0, "crit", "basic"), (500.0, "crit", "basic"), (200.0, "mean", "general"), (0.0, "BN98", "basic")]
    ap_list = [(3.0, False, "basic"), (3.0, True, "basic"), (10.0, False, "general"), (10.0, True, "basic")]
    proj_list = [(3.0, "basic"), (10.0, "basic")]
    filters_cfg = {
        "general": {"limit": 100, "properties": ["BoundSubhalo/NumberOfGasParticles",
                                                 "BoundSubhalo/NumberOfDarkMatterParticles",
                                                 "BoundSubhalo/NumberOfStarParticles",
                                                 "BoundSubhalo/NumberOfBlackHoleParticles"],
                    "combine_properties": "sum"},
    }
    data, H = make_chunk(4251, 24, 60000)
    # oracle dry run on one big halo: which internal keys exist per class
    so_o = [(t, v) for v, t, _ in so_list]
    so_o = [(t, 177.65 if t == "BN98" else v) for t, v in so_o]
    aps_o = [(kpc, kpc * 1e-3, incl) for kpc, incl, _ in ap_list]
    proj_o = [(kpc, kpc * 1e-3) for kpc, _ in proj_list]
    out_o, props_o = cmp.run_oracle(data, H, cp, so_o, aps_o, faithful=True, projected=proj_o, halos=[0])
    keys = {"sub": set(), "so": set(), "ap": set(), "proj": set()}
    for p in props_o:
        kind = {oh.SubhaloOracle: "sub", oh.SOOracle: "so", oh.ApertureOracle: "ap", oh.ProjectedApertureOracle: "proj"}[type(p)]
        res = out_o[0][0]
        for g, blk in res.items():
            if kind == "proj":
                if g.startswith(p.group_name):
                    keys[kind] |= set(blk)
            elif g == p.group_name:
                keys[kind] |= set(blk)
    # properties that need datasets outside the path's 48 B / particle stay off
    cellgrid, props = build_reference(filters_cfg, so_list, ap_list, proj_list, want(keys["sub"]), want(keys["so"]),
                                      want(keys["ap"]), want(keys["proj"]))
    from SOAP.core import halo_tasks
    from SOAP.core.shared_mesh import SharedMesh

    ua = rs.unyt_array
    ref_data = {}
    for t, d in data.items():
        ref_data[f"PartType{t}"] = {k: rs.shared(v.copy()) for k, v in d.items()}
        if t == 5:
            ref_data["PartType5"]["DynamicalMasses"] = ref_data["PartType5"]["Masses"]
    mesh = {pt: SharedMesh(rs._Comm(), ref_data[pt]["Coordinates"], 8) for pt in ref_data}
    # target density as process_halos computes it (halo_tasks.py:306-317)
    target = None
    for hp in props:
        for mult, dens in ((hp.mean_density_multiple, CRIT * OMEGA_M), (hp.critical_density_multiple, CRIT)):
            if mult is not None and (target is None or mult * dens < target):
                target = mult * dens
    # inputs are regenerated by the test from the same seeded recipe (make_chunk); only a checksum is kept
    out = {"n_halos": len(H["index"]), "target_density": target, "boxsize": BOX,
           "input_checksum": float(sum(float(np.sum(d["Coordinates"])) + float(np.sum(d["Masses"], dtype=np.float64))
                                       for d in data.values()))}
    n_h = len(H["index"])
    done = np.zeros(n_h, dtype=np.int32)
    sr_out = np.zeros(n_h)
    vals = {}
    out["config/so"] = np.array([f"{t}:{v}:{f}" for v, t, f in so_list])
    out["config/ap"] = np.array([f"{k}:{int(i)}:{f}" for k, i, f in ap_list])
    out["config/proj"] = np.array([f"{k}:{f}" for k, f in proj_list])
    out["config/filter_general_limit"] = 100
    groups = []
    n_done = n_fail = n_crash = 0
    for i in range(len(H["index"])):
        ih = {"cofp": ua(H["cofp"][i].copy()), "index": ua(H["index"][i]), "is_central": ua(H["is_central"][i]),
              "nr_bound_part": ua(H["nr_bound_part"][i]), "search_radius": ua(H["search_radius"][i]),
              "read_radius": ua(H["read_radius"][i]), "n_loop": ua(0)}
        td = ua(target) if ih["is_central"] == 1 else None  # halo_tasks.py:381
        try:
            res = halo_tasks.process_single_halo(mesh, None, ref_data, props, ua(CRIT), ua(CRIT * OMEGA_M), ua(BOX), ih, td)
        except AttributeError as e:
            # SO_properties.py:457 reads self.SO_r, which :424-433 never set when no particle is left after
            # the innermost one is skipped (a central with a single particle in its sphere): the reference
            # itself aborts on such a halo, so there is nothing to pin
            assert "SO_r" in str(e), e
            done[i] = -1
            n_crash += 1
            continue
        # returns the halo_result dict, or None when the read radius was too small
        if isinstance(res, tuple):
            res = res[0]
        done[i] = int(res is not None)
        sr_out[i] = float(ih["search_radius"])
        if res is None:
            n_fail += 1
            continue
        n_done += 1
        for hp in props:
            for name, prop in hp.property_list.items():
                gnames = [hp.group_name]
                if hp.__class__.__name__ == "ProjectedApertureProperties":
                    gnames = [f"{hp.group_name}/proj{ax}" for ax in "xyz"]
                for g in gnames:
                    key = f"{g}/{prop.name}"
                    if key in res:
                        v = np.atleast_1d(np.asarray(res[key][0]))
                        arr = vals.setdefault(f"{g}/{name}", np.zeros((n_h,) + v.shape, dtype=v.dtype))
                        arr[i] = v
                        if g not in groups:
                            groups.append(g)
        out.setdefault("n_loop", np.zeros(n_h, dtype=np.int64))[i] = int(np.asarray(res["InputHalos/n_loop"][0])) if "InputHalos/n_loop" in res else 0
    # what the drop-in adapter reads from the reference's property objects, for duck-typed stand-ins on hosts
    # without /root/reference (tests/test_dropin.py)
    import json

    meta = []
    for hp in props:
        ent = {"class": hp.__class__.__name__, "base_halo_type": hp.base_halo_type, "group_name": hp.group_name,
               "halo_filter": hp.halo_filter, "physical_radius_mpc": float(hp.physical_radius_mpc),
               "mean_density_multiple": None if hp.mean_density_multiple is None else float(hp.mean_density_multiple),
               "critical_density_multiple": None if hp.critical_density_multiple is None else float(hp.critical_density_multiple)}
        for k in ("type", "virial_definition", "inclusive", "all_radii_kpc", "aperture_physical_radius_kpc", "label"):
            if hasattr(hp, k):
                v = getattr(hp, k)
                ent[k] = v if isinstance(v, (str, bool, list, type(None))) else float(v)
        if hasattr(hp, "reference_density"):
            ent["reference_density"] = float(hp.reference_density)
        ent["properties"] = [[name, prop.name, int(prop.shape), np.dtype(prop.dtype).name, str(prop.unit), str(prop.description),
                              bool(prop.output_physical), prop.a_scale_exponent, bool(prop.dmo_property),
                              hp.property_filters[prop.name]] for name, prop in hp.property_list.items()
                             if hp.property_filters[prop.name]]
        meta.append(ent)
    out["meta_json"] = np.array(json.dumps(meta))
    out["groups"] = np.array(groups)
    out["done"] = done
    out["search_radius_out"] = sr_out
    for k, v in vals.items():
        out[f"val/{k}"] = v
    path = os.path.join(HERE, "halo_classes.npz")
    np.savez_compressed(path, **out)
    print(f"{path}: {len(out)} arrays, {os.path.getsize(path) / 1024:.1f} KiB; {n_done} halos done, {n_fail} need a larger read radius, {n_crash} abort in the reference")


if __name__ == "__main__":
    main()
