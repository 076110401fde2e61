"""Shared helpers of the parity tests: run the oracle and the CUDA path on the
same chunk and compare property by property with the tolerances of SURVEY.md
Appendix B (stated per class below)."""

import numpy as np

from oracle import halo as oh
from oracle import mesh as om
from soap_b200 import synth

# tolerance classes (relative to the stated scale)
TOL_MASS_RADIUS = 1e-6   # float64-accumulated masses and radii (north_star)
TOL_FIRST_MOMENT = 1e-5  # vcom-like first moments with cancellation, of rms speed
TOL_SECOND = 1e-4        # second-moment tensors, L, spin, Ekin (north_star)


def oracle_params(cp, faithful=False):
    return oh.Params(
        boxsize=cp["boxsize"], G=cp["G"], softening={t: cp["softening"] for t in (0, 1, 4, 5)},
        H=cp["H"], kpc_per_length=cp["kpc_per_length"], r_20mpc=cp["r_20mpc"],
        critical_density=cp["critical_density"], mean_density=cp["mean_density"],
        phys_mpc_to_coord=cp["phys_mpc_to_coord"], nu_density=cp["nu_density"], faithful=faithful,
    )


def device_config(cp, so=(), apertures=(), flags=0, dmo=False, do_subhalo=True, projected=(), filters=None,
                  so_filters=None, ap_filters=None, proj_filters=None, skip_gt=()):
    from soap_b200.halo_tasks import HaloPropConfig

    return HaloPropConfig(
        filters=dict(filters or {}), so_filter=list(so_filters or []), ap_filter=list(ap_filters or []),
        proj_filter=list(proj_filters or []), skip_gt=tuple(skip_gt),
        boxsize=cp["boxsize"], G=cp["G"], critical_density=cp["critical_density"],
        mean_density=cp["mean_density"], softening={t: cp["softening"] for t in (0, 1, 4, 5)},
        H=cp["H"], kpc_per_length=cp["kpc_per_length"], r_20mpc=cp["r_20mpc"],
        nu_density=cp["nu_density"], phys_mpc_to_coord=cp["phys_mpc_to_coord"],
        do_subhalo=do_subhalo, so=list(so), apertures=list(apertures), projected=list(projected),
        property_flags=flags, dmo=dmo,
    )


def oracle_prop_list(params, cp, so, apertures, do_subhalo=True, projected=(), so_filters=None, ap_filters=None,
                     proj_filters=None, skip_gt=()):
    """halo_prop_list of the oracle; *_filters give the halo_filter category of each variation (in the order
    of ``so`` / of the sorted apertures), default "basic".  skip_gt names the aperture kinds ("exclusive",
    "inclusive", "projected") that know the radii of their siblings (all_radii_kpc, compute_halo_properties.py:
    345-395; the reference always passes them to exclusive spheres) and so use the EncloseRadius shortcut."""
    props = []
    if do_subhalo:
        props.append(oh.SubhaloOracle(params))
    for k, (t, val) in enumerate(so):
        props.append(oh.SOOracle(params, val, t, halo_filter=so_filters[k] if so_filters else "basic"))
    order = sorted(range(len(apertures)), key=lambda i: (apertures[i][0], apertures[i][2]))
    prev = {}  # kind -> (radius, group name) of the previous aperture of that kind
    for i, j in enumerate(order):
        r, mpc, incl = apertures[j]
        kind = "inclusive" if incl else "exclusive"
        pr, pg = prev.get(kind, (None, None)) if kind in skip_gt else (None, None)
        props.append(oh.ApertureOracle(params, r, mpc, bool(incl), f"{i}",
                                       halo_filter=ap_filters[j] if ap_filters else "basic", prev_radius=pr,
                                       prev_group=pg))
        prev[kind] = (r, props[-1].group_name)
    porder = sorted(range(len(projected)), key=lambda i: projected[i][0])
    for i, j in enumerate(porder):
        r, mpc = projected[j]
        pr, pg = prev.get("projected", (None, None)) if "projected" in skip_gt else (None, None)
        props.append(oh.ProjectedApertureOracle(params, r, mpc, f"{i}",
                                                halo_filter=proj_filters[j] if proj_filters else "basic",
                                                prev_radius=pr, prev_group=pg))
        prev["projected"] = (r, props[-1].group_name)
    return props


def run_oracle(data, H, cp, so, apertures, faithful=False, halos=None, do_subhalo=True, projected=(), iterative=False,
               filters=None, so_filters=None, ap_filters=None, proj_filters=None, mesh_resolution=None, skip_gt=()):
    """Returns list (per halo) of (halo_result or None, info, input_halo)."""
    params = oracle_params(cp, faithful)
    params.iterative_tensors = bool(iterative)
    params.filters = dict(filters or {})
    meshes = {t: om.MeshOracle(d["Coordinates"], mesh_resolution or om.mesh_resolution(len(d["Masses"])))
              for t, d in data.items()}
    props = oracle_prop_list(params, cp, so, apertures, do_subhalo, projected, so_filters, ap_filters, proj_filters, skip_gt)
    td = oh.target_density_of(props, params)
    out = []
    idxs = range(len(H["index"])) if halos is None else halos
    for i in idxs:
        ih = {k: (v[i].copy() if v.ndim > 1 else v[i]) for k, v in H.items()}
        try:
            res, info = oh.process_single_halo(meshes, data, props, params, ih, td if ih["is_central"] == 1 else None)
            err = None
        except RuntimeError as e:  # the reference would abort the run here
            res, info, err = None, {"n_loop": -1}, str(e)
        out.append((res, info, ih, err))
    return out, props


GEN_KEYS = ["Ngas", "Ndm", "Nstar", "Nbh", "Mgas", "Mdm", "Mstar", "Mbh", "com", "vcom"]


def _group_names(props):
    """device block prefix for each oracle prop, in halo_prop_list order"""
    names = []
    k = a = 0
    for p in props:
        if isinstance(p, oh.SubhaloOracle):
            names.append(("BoundSubhalo/", p.group_name, "sub"))
        elif isinstance(p, oh.SOOracle):
            names.append((f"SO/{k}/", p.group_name, "so"))
            k += 1
        elif isinstance(p, oh.ApertureOracle):
            names.append((f"Aperture/{a}/", p.group_name, "ap"))
            a += 1
        elif isinstance(p, oh.ProjectedApertureOracle):
            j = sum(1 for n in names if n[2] == "proj") // 3
            for ax in "xyz":
                names.append((f"ProjectedAperture/{j}/proj{ax}/", f"{p.group_name}/proj{ax}", "proj"))
    return names


class Report:
    def __init__(self):
        self.maxerr = {}
        self.bad = []

    def check(self, name, halo, got, ref, tol, scale=None, exact=False):
        got = np.asarray(got, dtype=np.float64)
        ref = np.asarray(ref, dtype=np.float64)
        if exact:
            err = float(np.max(np.abs(got - ref))) if got.size else 0.0
            ok = err == 0.0
        else:
            sc = float(np.max(np.abs(ref))) if scale is None else float(scale)
            if sc == 0.0:
                err = float(np.max(np.abs(got - ref)))
                ok = err <= 1e-300 or err <= tol
            else:
                err = float(np.max(np.abs(got - ref))) / sc
                ok = err <= tol
        key = name.split("/")[-1]
        self.maxerr[key] = max(self.maxerr.get(key, 0.0), err)
        if not ok:
            self.bad.append((name, halo, got.tolist(), ref.tolist(), err))

    def assert_ok(self):
        assert not self.bad, "parity failures (first 10): " + "\n".join(str(b) for b in self.bad[:10])


def compare(res, oracle_out, props, cp, halos=None, flags=0, rep=None, faithful=False):
    """res: HaloResults of the device path; oracle_out from run_oracle.  faithful = the oracle ran in its
    dtype-for-dtype mode (float32 sums where the reference has them): SURVEY.md Appendix B allows 4e-6 there
    for the 1e-6 class (numpy's pairwise float32 sums carry that much noise themselves)."""
    rep = rep or Report()
    TOL_MASS_RADIUS = 4e-6 if faithful else 1e-6
    TOL_FIRST_MOMENT = 4e-5 if faithful else 1e-5
    L = cp["boxsize"]
    names = _group_names(props)
    status = res.status.cpu().numpy()
    hsel = range(len(oracle_out)) if halos is None else halos
    tab = {n: res.get(n) for n in res.names()}
    for j, h in enumerate(hsel):
        ores, info, ih, err = oracle_out[j]
        if err is not None:
            assert status[h] >= 2, (h, err, status[h])
            continue
        if ores is None:
            assert status[h] == 1, (h, status[h])
            rep.check("InputHalos/search_radius", h, tab["InputHalos/search_radius"][h], ih["search_radius"], 1e-14)
            continue
        assert status[h] == 0, (h, status[h], info)
        rep.check("InputHalos/n_loop", h, tab["InputHalos/n_loop"][h], info["n_loop"], 0, exact=True)
        rep.check("InputHalos/radius", h, tab["InputHalos/radius"][h], info["radius"], 1e-15)
        npairs = sum(len(v) for v in info["idx"].values())
        rep.check("InputHalos/n_pairs", h, tab["InputHalos/n_pairs"][h], npairs, 0, exact=True)
        for pre, gname, kind in names:
            o = ores.get(gname, {})
            g = lambda k: tab[pre + k][h]
            if kind == "proj":
                for k in ("Ngas", "Ndm", "Nstar", "Nbh"):
                    rep.check(pre + k, h, g(k), o.get(k, 0), 0, exact=True)
                for k in ("Mgas", "Mdm", "Mstar", "Mbh", "Mtot"):
                    rep.check(pre + k, h, g(k), o.get(k, 0.0), TOL_MASS_RADIUS)
                if "com" in o:
                    d = (g("com") - np.asarray(o["com"]) + 0.5 * L) % L - 0.5 * L
                    rep.check(pre + "com", h, d, np.zeros(3), TOL_MASS_RADIUS, scale=info["radius"])
                    rep.check(pre + "vcom", h, g("vcom"), o["vcom"], TOL_FIRST_MOMENT, scale=300.0)
                for nm in ("gas", "dm", "star"):
                    # 1-D dispersion about the type's mean: cancels to 0 for a single particle
                    rep.check(pre + f"proj_veldisp_{nm}", h, g(f"proj_veldisp_{nm}"), o.get(f"proj_veldisp_{nm}", 0.0),
                              1e-2 * TOL_SECOND ** 0.5 if o.get(f"proj_veldisp_{nm}", 0.0) == 0.0 else TOL_SECOND, scale=300.0)
                    if flags & 8:
                        k = f"HalfMassRadius{nm.capitalize()}"
                        rep.check(pre + k, h, g(k), o.get(k, 0.0), TOL_MASS_RADIUS)
                for suffix in ("Noniterative", "ReducedNoniterative") + (("", "Reduced") if flags & 16 else ()):
                    k = "ProjectedTotalInertiaTensor" + suffix
                    ref = np.asarray(o.get(k, np.zeros(3)), dtype=np.float64)
                    rep.check(pre + k, h, g(k), ref, TOL_SECOND, scale=(np.sqrt((ref[:2] ** 2).sum() + 2 * ref[2] ** 2) or None))
                continue
            # counts: bit exact
            for k in ("Ngas", "Ndm", "Nstar", "Nbh"):
                rep.check(pre + k, h, g(k), o.get(k, 0), 0, exact=True)
            for k in ("Mgas", "Mdm", "Mstar", "Mbh"):
                rep.check(pre + k, h, g(k), o.get(k, 0.0), TOL_MASS_RADIUS)
            mt_key = "Mtotpart" if kind == "so" else "Mtot"
            rep.check(pre + "Mtot", h, g("Mtot"), o.get(mt_key, 0.0), TOL_MASS_RADIUS)
            # scale for positions: the selection radius
            rscale = info["radius"]
            if "com" in o:
                d = (g("com") - np.asarray(o["com"]) + 0.5 * L) % L - 0.5 * L
                rep.check(pre + "com", h, d, np.zeros(3), TOL_MASS_RADIUS, scale=rscale)
                rep.check(pre + "vcom", h, g("vcom"), o["vcom"], TOL_FIRST_MOMENT, scale=300.0)
            else:
                rep.check(pre + "com", h, g("com"), np.zeros(3), 0, exact=True)
            # get_vmax sums the masses with a float32 cumsum (kinematic_properties.py:583): the reference's own
            # Vmax carries up to n * 2^-24 of rounding, which the faithful oracle reproduces and the float64
            # device sums do not (the float64 oracle comparison keeps the plain tolerance)
            nsel = sum(float(o.get(k, 0)) for k in ("Ngas", "Ndm", "Nstar", "Nbh"))
            tol_vmax = TOL_MASS_RADIUS + (nsel * 2.0**-24 if faithful else 0.0)
            if kind in ("so", "sub"):
                rep.check(pre + "Vmax_soft", h, g("Vmax_soft"), o.get("Vmax_soft", 0.0), tol_vmax)
                rep.check(pre + "R_vmax_soft", h, g("R_vmax_soft"), o.get("R_vmax_soft", 0.0), TOL_MASS_RADIUS)
                rep.check(pre + "spin_parameter", h, g("spin_parameter"), o.get("spin_parameter", 0.0), TOL_SECOND + tol_vmax)
            if kind == "ap":  # aperture_properties.py:3553-3577
                rep.check(pre + "Vmax_soft", h, g("Vmax_soft"), o.get("Vmax_soft", 0.0), tol_vmax)
                rep.check(pre + "R_vmax_soft", h, g("R_vmax_soft"), o.get("R_vmax_soft", 0.0), TOL_MASS_RADIUS)
            if kind == "sub":
                for k in ("EncloseRadius", "R_vmax_unsoft", "HalfMassRadiusTot"):
                    rep.check(pre + k, h, g(k), o.get(k, 0.0), TOL_MASS_RADIUS)
                rep.check(pre + "Vmax_unsoft", h, g("Vmax_unsoft"), o.get("Vmax_unsoft", 0.0), tol_vmax)
            if kind == "so":
                rep.check(pre + "r", h, g("r"), o.get("r", 0.0), TOL_MASS_RADIUS)
                rep.check(pre + "Mso", h, g("Mso"), o.get("Mtot", 0.0), TOL_MASS_RADIUS)
                for k in ("Mfrac_satellites", "Mfrac_external"):
                    rep.check(pre + k, h, g(k), o.get(k, 0.0), TOL_MASS_RADIUS, scale=1.0)
                for k in ("concentration_unsoft", "concentration_soft", "concentration_dmo_unsoft", "concentration_dmo_soft"):
                    rep.check(pre + k, h, g(k), o.get(k, 0.0), 1e-5)
            if flags & 8 and kind in ("sub", "ap"):
                for k in ("HalfMassRadiusGas", "HalfMassRadiusDM", "HalfMassRadiusStar", "HalfMassRadiusBaryon"):
                    rep.check(pre + k, h, g(k), o.get(k, 0.0), TOL_MASS_RADIUS)
            if flags & 1:
                for grp in ("gas", "dm", "star"):
                    if f"com_{grp}" in o:
                        d = (g(f"com_{grp}") - np.asarray(o[f"com_{grp}"]) + 0.5 * L) % L - 0.5 * L
                        rep.check(pre + f"com_{grp}", h, d, np.zeros(3), TOL_MASS_RADIUS, scale=rscale)
                        rep.check(pre + f"vcom_{grp}", h, g(f"vcom_{grp}"), o[f"vcom_{grp}"], TOL_FIRST_MOMENT, scale=300.0)
                        vd = np.asarray(o[f"veldisp_matrix_{grp}"], dtype=np.float64)
                        rep.check(pre + f"veldisp_matrix_{grp}", h, g(f"veldisp_matrix_{grp}"), vd, TOL_SECOND,
                                  scale=np.sqrt((vd[:3] ** 2).sum() + 2 * (vd[3:] ** 2).sum()))
                        Lr = np.asarray(o[f"L{grp}"], dtype=np.float64)
                        # |L| can cancel to ~0: scale with M * r * v
                        Lscale = max(np.linalg.norm(Lr), 1e-3 * o[{"gas": "Mgas", "dm": "Mdm", "star": "Mstar"}[grp]] * rscale * 300.0)
                        rep.check(pre + f"L{grp}", h, g(f"L{grp}"), Lr, TOL_SECOND, scale=Lscale)
                if "Lbaryons" in o:
                    Lr = np.asarray(o["Lbaryons"], dtype=np.float64)
                    Lscale = max(np.linalg.norm(Lr), 1e-3 * (o["Mgas"] + o["Mstar"]) * rscale * 300.0)
                    rep.check(pre + "Lbaryons", h, g("Lbaryons"), Lr, TOL_SECOND, scale=Lscale)
                if kind == "sub" and "KineticEnergyTotal" in o:
                    # central second moment: can cancel to ~0 (a one-particle halo has none): scale with M v^2
                    ek = float(o["KineticEnergyTotal"])
                    rep.check(pre + "Ekin_tot", h, g("Ekin_tot"), ek, TOL_SECOND,
                              scale=max(abs(ek), 1e-3 * o["Mtot"] * 300.0**2))
            if flags & 2 and kind == "so":
                # SOProperties has the disc fractions only
                for k in ("DtoTgas", "DtoTstar"):
                    rep.check(pre + k, h, g(k), o.get(k, 0.0), TOL_SECOND, scale=1.0)
            if flags & 2 and kind in ("sub", "ap"):
                # kappa_corot / DtoT are ratios of second moments: absolute tolerance
                for k in ("kappa_corot_gas", "kappa_corot_star", "kappa_corot_baryons", "DtoTgas", "DtoTstar"):
                    rep.check(pre + k, h, g(k), o.get(k, 0.0), TOL_SECOND, scale=1.0)
                # stellar rotation / cylindrical dispersions: velocities, scale = typical particle speed
                for k in ("StellarRotationalVelocity", "StellarCylindricalVelocityDispersion",
                          "StellarCylindricalVelocityDispersionVertical", "StellarCylindricalVelocityDispersionDiscPlane"):
                    rep.check(pre + k, h, g(k), o.get(k, 0.0), TOL_SECOND, scale=300.0)
            if flags & 4:
                tn = "StellarInertiaTensor" if kind == "ap" else "TotalInertiaTensor"
                # iterative variants (flags bit 4): same tolerance; a particle exactly on the
                # ellipsoid surface could flip a pass, which none of the seeded cases hits
                for suffix in ("Noniterative", "ReducedNoniterative") + (("", "Reduced") if flags & 16 else ()):
                    k = tn + suffix
                    ref = np.asarray(o.get(k, np.zeros(6)), dtype=np.float64)
                    rep.check(pre + k, h, g(k), ref, TOL_SECOND,
                              scale=(np.sqrt((ref[:3] ** 2).sum() + 2 * (ref[3:] ** 2).sum()) or None))
    return rep
