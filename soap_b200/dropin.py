"""
Drop-in for the reference's halo loop: ``process_halos`` with the reference's own signature and side
effects (SOAP/core/halo_tasks.py:276-430), fed by the reference's own ``halo_prop_list`` objects
(SubhaloProperties / SOProperties / Exclusive- and InclusiveSphereProperties /
ProjectedApertureProperties, SURVEY.md 8(b) boundaries 2 and 3).

    from soap_b200.dropin import process_halos
    total_time, task_time, nr_left, nr_done, min_free_mem_gb = process_halos(
        comm, unit_registry, data, mesh, halo_prop_list, critical_density, mean_density, boxsize,
        halo_arrays, results)

What it does instead of the per-core task loop:

* ``HaloPropConfig.from_halo_prop_list`` reads the numerical content of the property objects -- SO type and
  ``reference_density``, aperture radii, ``inclusive``, ``halo_filter`` and the CategoryFilter limits,
  ``all_radii_kpc`` (the EncloseRadius shortcut), softenings, cosmology scalars -- and hands every threshold
  to the device in coordinate units (comoving ``snap_length``, ``snap_mass``, the snapshot's velocity unit),
  i.e. the value unyt produces at the comparison it appears in;
* the chunk's particles go to the GPU once (``DeviceChunk``; ``mesh`` -- the per-ptype SharedMesh dict of the
  reference -- is not needed: the device builds its own cell list) and ``soap_process_halos`` runs the radius
  ladder and every reduction for all halos that are not done;
* ``rows_to_halo_result`` turns a finished halo's table row back into the reference's
  ``halo_result`` dict: ``f"{group_name}/{prop.name}" -> (array[dtype of PropertyTable], description, physical,
  a_exponent)`` for every property enabled in the parameter file, per-property category filters applied
  (a property whose filter is not satisfied stays zero, e.g. SO_properties.py:3675);
* side effects on ``halo_arrays`` / ``results`` are the reference's: ``results.append``, ``done = 1``, and for
  halos that need a larger region ``read_radius`` (x1.5 or the required radius) and ``search_radius``.

Properties outside the device path (gas thermodynamics, metals, luminosities, ...: SURVEY.md 8(f) rank 1)
are refused with a NotImplementedError naming them, never silently dropped or computed on the CPU.

Units: with a real ``unyt`` the thresholds are converted with ``.to()`` and the results are wrapped in
``unyt_array``s with the property's output unit; the numbers are then in snapshot units by construction of
the device configuration.  That conversion path cannot be exercised in an image without unyt (SURVEY.md
Appendix C); the tests drive the adapter with dimension-blind stand-ins and with plain floats.
"""

import time

import numpy as np

from . import _lib
from .halo_tasks import PF_HMR, PF_ITER, PF_KAPPA, PF_KIN, PF_TENS, DeviceChunk, HaloPropConfig
from .halo_tasks import process_halos as _device_process_halos

READ_RADIUS_FACTOR = 1.5  # halo_tasks.py:17

# ------------------------------------------------------------------ property name -> device column
# internal property name of the reference class -> device column of the block, where they differ
_ALIAS = {
    ("so", "Mtot"): "Mso",  # SO TotalMass is the SO mass; the device's Mtot column is the particle mass inside
    ("*", "Mbh_dynamical"): "Mbh",
    ("*", "KineticEnergyTotal"): "Ekin_tot",
    ("*", "KineticEnergyGas"): "Ekin_gas",
    ("*", "KineticEnergyStars"): "Ekin_star",
    ("proj", "HalfMassRadiusDM"): "HalfMassRadiusDm",
}
# property group each device column belongs to (property_flags bit that must be on)
_GROUP_OF = {}
for _n in ("com_gas", "com_dm", "com_star", "vcom_gas", "vcom_dm", "vcom_star", "Lgas", "Ldm", "Lstar", "Lbaryons",
           "veldisp_matrix_gas", "veldisp_matrix_dm", "veldisp_matrix_star", "Ekin_tot", "Ekin_gas", "Ekin_star"):
    _GROUP_OF[_n] = PF_KIN
for _n in ("kappa_corot_gas", "kappa_corot_star", "kappa_corot_baryons", "DtoTgas", "DtoTstar", "StellarRotationalVelocity",
           "StellarCylindricalVelocityDispersion", "StellarCylindricalVelocityDispersionVertical",
           "StellarCylindricalVelocityDispersionDiscPlane"):
    _GROUP_OF[_n] = PF_KIN | PF_KAPPA
for _n in ("TotalInertiaTensorNoniterative", "TotalInertiaTensorReducedNoniterative", "StellarInertiaTensorNoniterative",
           "StellarInertiaTensorReducedNoniterative"):
    _GROUP_OF[_n] = PF_TENS
for _n in ("TotalInertiaTensor", "TotalInertiaTensorReduced", "StellarInertiaTensor", "StellarInertiaTensorReduced",
           "ProjectedTotalInertiaTensor", "ProjectedTotalInertiaTensorReduced"):
    _GROUP_OF[_n] = PF_TENS | PF_ITER
for _n in ("HalfMassRadiusGas", "HalfMassRadiusDM", "HalfMassRadiusDm", "HalfMassRadiusStar", "HalfMassRadiusBaryon"):
    _GROUP_OF[_n] = PF_HMR
_ALL_FLAGS = PF_KIN | PF_KAPPA | PF_TENS | PF_HMR | PF_ITER


def _kind(halo_prop):
    t = getattr(halo_prop, "base_halo_type", None)
    if t is None:
        raise TypeError(f"{type(halo_prop).__name__} has no base_halo_type: not a SOAP HaloProperty")
    return {"SubhaloProperties": "sub", "SOProperties": "so", "ApertureProperties": "ap",
            "ProjectedApertureProperties": "proj"}[t]


def _val(q):
    """number(s) of a unyt quantity / array, a SharedArray, or a plain number"""
    if hasattr(q, "full"):
        q = q.full
    if hasattr(q, "value"):
        q = q.value
    return np.asarray(q)


def _in_units(q, unit):
    """q expressed in ``unit`` (a unyt.Unit, or None when no unit system is in play)"""
    if unit is not None and hasattr(q, "to"):
        q = q.to(unit)
    return _val(q)


class _Units:
    """The coordinate unit system of the device path, as unyt units (None without unyt / registry)."""

    def __init__(self, unit_registry):
        self.length = self.mass = self.velocity = self.density = self.mpc = self.kpc = self.G = self.H = None
        self.make = None
        if unit_registry is None:
            return
        try:
            import unyt
        except ImportError:
            return
        u = lambda s: unyt.Unit(s, registry=unit_registry)  # noqa: E731
        self.length = u("snap_length") * u("a")
        self.mass = u("snap_mass")
        self.velocity = u("snap_length") / u("snap_time")
        self.density = self.mass / self.length**3
        self.mpc, self.kpc = u("Mpc"), u("kpc")
        self.G = self.velocity**2 * self.length / self.mass
        self.H = self.velocity / self.length
        self.make = lambda val, unit_str, a_exp, dtype: unyt.unyt_array(  # noqa: E731
            val, dtype=dtype, units=u(unit_str) * (u("a") ** a_exp if a_exp else 1), registry=unit_registry)


def _enabled_properties(halo_prop):
    """(internal name, Property) of every property the parameter file enables for this variation"""
    dmo = bool(getattr(getattr(halo_prop, "category_filter", None), "dmo", False))
    out = []
    for name, prop in halo_prop.property_list.items():
        flt = halo_prop.property_filters[prop.name]
        if not flt:
            continue
        if dmo and not prop.dmo_property:
            continue
        out.append((name, prop, flt))
    return out


def config_from_halo_prop_list(halo_prop_list, boxsize, critical_density, mean_density, unit_registry=None, dmo=None):
    """HaloPropConfig + the (kind, device prefix) of every entry of ``halo_prop_list``.

    Order matters in the reference (BoundSubhalo first, SOs, spheres ascending, projected apertures): the
    device commits properties in that order, so another order is refused."""
    U = _Units(unit_registry)
    kinds = [_kind(hp) for hp in halo_prop_list]
    order = {"sub": 0, "so": 1, "ap": 2, "proj": 3}
    if [order[k] for k in kinds] != sorted(order[k] for k in kinds):
        raise ValueError("halo_prop_list must be ordered BoundSubhalo, SO..., spheres..., projected apertures...")
    if kinds.count("sub") > 1:
        raise ValueError("more than one SubhaloProperties entry")
    first = halo_prop_list[0]
    cat = getattr(first, "category_filter", None)
    if dmo is None:
        dmo = bool(getattr(cat, "dmo", False))
    a = float(getattr(first, "a", 1.0))
    soft = {}
    for pt, s in getattr(first, "softening_of_parttype", {}).items():
        soft[int(str(pt)[-1])] = float(_in_units(s, U.length))
    G = 1.0
    try:
        import unyt

        G = float(_in_units(unyt.physical_constants.newton_G, U.G))
    except Exception:
        G = float(getattr(first, "newton_G", 1.0))
    cosmo = {}
    for hp in halo_prop_list:
        cosmo.update(getattr(hp, "cosmology", {}) or {})
    H = float(_in_units(cosmo["H"], U.H)) if "H" in cosmo else 0.0
    nu = float(_in_units(cosmo["nu_density"], U.density)) if "nu_density" in cosmo else 0.0
    one = lambda x, unit: float(_in_units(x, unit))  # noqa: E731
    try:
        import unyt

        kpc_to_coord = one(1.0 * unyt.Unit("kpc", registry=unit_registry), U.length) if U.length is not None else 1.0
        mpc_to_coord = one(1.0 * unyt.Unit("Mpc", registry=unit_registry), U.length) if U.length is not None else 1.0
    except Exception:
        kpc_to_coord, mpc_to_coord = 1.0, 1.0
    kpc_per_length = 1.0 / kpc_to_coord
    filters = {}
    for name, info in (getattr(cat, "filters", None) or {}).items():
        ptypes = []
        for p in info["properties"]:
            key = p.split("/")[-1]
            ptypes.append({"NumberOfGasParticles": 0, "NumberOfDarkMatterParticles": 1, "NumberOfStarParticles": 4,
                           "NumberOfBlackHoleParticles": 5}[key])
            if not p.startswith("BoundSubhalo/"):
                raise NotImplementedError(f"filter {name} reads {p}: only BoundSubhalo particle counts are supported")
        if len(info["properties"]) > 1 and info.get("combine_properties") != "sum":
            raise NotImplementedError(f"Invalid combine_properties function for filter {name}")
        filters[name] = (int(info["limit"]), tuple(ptypes))
    so, so_filter, aps, ap_filter, proj, proj_filter, so_rho, so_vir = [], [], [], [], [], [], [], []
    # halo_tasks.py:306-317
    target = None
    for hp in halo_prop_list:
        for mult, dens in ((getattr(hp, "mean_density_multiple", None), mean_density),
                           (getattr(hp, "critical_density_multiple", None), critical_density)):
            if mult is not None:
                d = float(_in_units(mult * dens, U.density))
                if target is None or d < target:
                    target = d
    skip_gt = set()
    flags = 0
    enclose_on = False
    for hp, kind in zip(halo_prop_list, kinds):
        enabled = _enabled_properties(hp)
        for name, prop, _ in enabled:
            col = _ALIAS.get((kind, name), _ALIAS.get(("*", name), name))
            flags |= _GROUP_OF.get(col, 0)
        if kind == "sub":
            enclose_on = any(name == "EncloseRadius" for name, _, _ in enabled)
        elif kind == "so":
            if hp.type == "physical":
                raise NotImplementedError("SO variations with a fixed physical radius are not on the device path")
            if getattr(hp, "core_excision_fraction", None) is not None:
                raise NotImplementedError("core-excised SO variations are not on the device path")
            # the device takes the reference density itself (what compute_SO_radius_and_mass is called with)
            rho = float(_in_units(hp.reference_density, U.density))
            mult = rho / float(_in_units(critical_density if hp.type in ("crit", "BN98") else mean_density, U.density))
            so.append((hp.type, mult))
            so_rho.append(rho)
            so_vir.append(bool(hp.virial_definition))
            so_filter.append(hp.halo_filter)
        elif kind == "ap":
            if getattr(hp, "aperture_physical_radius_kpc", None) is None:
                raise NotImplementedError("apertures defined by another property are not on the device path")
            # the radius the mask uses is ``aperture_physical_radius_kpc * unyt.kpc`` (aperture_properties.py:4129);
            # physical_radius_mpc is what a too small search radius asks for (halo_tasks.py:168)
            mpc = float(hp.physical_radius_mpc)
            aps.append((float(hp.aperture_physical_radius_kpc) * kpc_to_coord, mpc, bool(hp.inclusive)))
            ap_filter.append(hp.halo_filter)
            if getattr(hp, "all_radii_kpc", None) is not None and len(hp.all_radii_kpc) > 1:
                skip_gt.add("inclusive" if hp.inclusive else "exclusive")
        else:
            if getattr(hp, "aperture_physical_radius_kpc", None) is None:
                raise NotImplementedError("projected apertures defined by another property are not on the device path")
            mpc = float(hp.physical_radius_mpc)
            proj.append((float(hp.aperture_physical_radius_kpc) * kpc_to_coord, mpc))
            proj_filter.append(hp.halo_filter)
    if not enclose_on:
        skip_gt = set()  # the shortcut needs BoundSubhalo/EncloseRadius in halo_result (aperture_properties.py:4090)
    cfg = HaloPropConfig(
        boxsize=float(_in_units(boxsize, U.length)), G=G, critical_density=float(_in_units(critical_density, U.density)),
        mean_density=float(_in_units(mean_density, U.density)), softening=soft, H=H, kpc_per_length=kpc_per_length * 1.0,
        r_20mpc=20.0 * mpc_to_coord, nu_density=nu, phys_mpc_to_coord=mpc_to_coord, do_subhalo="sub" in kinds, so=so,
        apertures=aps, projected=proj, property_flags=flags, dmo=dmo, filters=filters, so_filter=so_filter,
        ap_filter=ap_filter, proj_filter=proj_filter, skip_gt=tuple(sorted(skip_gt)), so_rho=so_rho,
        so_virial_flags=so_vir, target_density_value=-1.0 if target is None else target,
    )
    cfg._a = a
    return cfg


def _prefixes(cfg, halo_prop_list):
    """device block prefix(es) of every halo_prop, in halo_prop_list order: [(kind, [(prefix, group_name)])]"""
    cfg.to_c()  # fixes the sorted aperture order
    out = []
    k = 0
    ap_seen = {}
    pj_seen = {}
    for hp in halo_prop_list:
        kind = _kind(hp)
        if kind == "sub":
            out.append((kind, [("BoundSubhalo/", hp.group_name)]))
        elif kind == "so":
            out.append((kind, [(f"SO/{k}/", hp.group_name)]))
            k += 1
        elif kind == "ap":
            key = (float(hp.physical_radius_mpc), bool(hp.inclusive))
            idx = [i for i, (r, mpc, incl) in enumerate(cfg._sorted_apertures) if (float(mpc), bool(incl)) == key]
            i = idx[ap_seen.get(key, 0)]
            ap_seen[key] = ap_seen.get(key, 0) + 1
            out.append((kind, [(f"Aperture/{i}/", hp.group_name)]))
        else:
            key = float(hp.physical_radius_mpc)
            idx = [i for i, (r, mpc) in enumerate(cfg._sorted_projected) if float(mpc) == key]
            i = idx[pj_seen.get(key, 0)]
            pj_seen[key] = pj_seen.get(key, 0) + 1
            out.append((kind, [(f"ProjectedAperture/{i}/proj{ax}/", f"{hp.group_name}/proj{ax}") for ax in "xyz"]))
    return out


class ResultPacker:
    """Table rows -> the reference's halo_result dicts for one halo_prop_list / configuration."""

    def __init__(self, halo_prop_list, cfg, cols, unit_registry=None):
        self.U = _Units(unit_registry)
        self.cfg = cfg
        self.cols = cols
        self.plan = []  # (key, column offset, width, dtype, description, physical, a_exponent, unit, filter name)
        missing = []
        for hp, (kind, blocks) in zip(halo_prop_list, _prefixes(cfg, halo_prop_list)):
            for name, prop, flt in _enabled_properties(hp):
                col = _ALIAS.get((kind, name), _ALIAS.get(("*", name), name))
                for prefix, group in blocks:
                    if prefix + col not in cols:
                        missing.append(f"{group}/{prop.name}")
                        continue
                    off, w = cols[prefix + col]
                    if w != int(prop.shape):
                        raise ValueError(f"{group}/{prop.name}: device width {w} != PropertyTable shape {prop.shape}")
                    desc = prop.description
                    try:
                        desc = desc.format(label=getattr(hp, "label", ""), core_excision=getattr(hp, "core_excision_string", None))
                    except (KeyError, IndexError):
                        pass
                    self.plan.append((f"{group}/{prop.name}", off, w, prop.dtype, desc, prop.output_physical,
                                      prop.a_scale_exponent, prop.unit, flt))
        if missing:
            raise NotImplementedError(
                "properties outside the soap_b200 device path are enabled in the parameter file (SURVEY.md 8(f) rank 1); "
                "disable them or compute them with the reference: " + ", ".join(sorted(set(missing))[:20]))
        fl = cfg.filters
        self._filter_types = {n: tuple({0: "Ngas", 1: "Ndm", 4: "Nstar", 5: "Nbh"}[t] for t in types) for n, (_, types) in fl.items()}
        self._filter_limit = {n: lim for n, (lim, _) in fl.items()}

    def do_calculation(self, row):
        """CategoryFilter.get_do_calculation (category_filter.py:69-110) from the BoundSubhalo counts of the row"""
        out = {"basic": True}
        for n, keys in self._filter_types.items():
            v = sum(int(row[self.cols["BoundSubhalo/" + k][0]]) for k in keys)
            out[n] = v >= self._filter_limit[n]
        return out

    def halo_result(self, row):
        do = self.do_calculation(row)
        res = {}
        for key, off, w, dtype, desc, physical, a_exp, unit, flt in self.plan:
            val = row[off] if w == 1 else row[off:off + w]
            if not do[flt]:
                val = np.zeros_like(val)
            arr = np.asarray(val).astype(dtype)
            if self.U.make is not None:
                arr = self.U.make(arr, unit, None if physical else a_exp, dtype)
            res[key] = (arr, desc, physical, a_exp)
        return res


def rows_to_halo_result(table, halo_prop_list, cfg, cols, unit_registry=None):
    """list of halo_result dicts, one per row of the device table (halo_tasks.py:196-271, each ``calculate()`` tail)"""
    pk = ResultPacker(halo_prop_list, cfg, cols, unit_registry)
    return [pk.halo_result(r) for r in np.asarray(table)]


_chunk_cache = {}


def _device_chunk(data, boxsize_coord, device):
    """one DeviceChunk per ``data`` dict (the reference passes the same dict for every pass over a chunk)"""
    key = id(data)
    hit = _chunk_cache.get(key)
    if hit is not None and hit[0] is data:
        return hit[1]
    for _, (_, ch) in list(_chunk_cache.items()):
        ch.free()
    _chunk_cache.clear()
    ch = DeviceChunk({pt: {k: _val(v) for k, v in d.items()} for pt, d in data.items()}, boxsize_coord, device=device)
    _chunk_cache[key] = (data, ch)
    return ch


def process_halos(comm, unit_registry, data, mesh, halo_prop_list, critical_density, mean_density, boxsize, halo_arrays,
                  results, device=0):
    """Same call, same side effects, same 5-tuple as SOAP.core.halo_tasks.process_halos (:276-430)."""
    t0_all = time.time()
    if comm is not None:
        comm.barrier()
    cfg = config_from_halo_prop_list(halo_prop_list, boxsize, critical_density, mean_density, unit_registry)
    cfg_c = cfg.to_c()
    from .halo_tasks import result_layout

    ncol, cols = result_layout(cfg_c)
    packer = ResultPacker(halo_prop_list, cfg, cols, unit_registry)
    U = packer.U
    done = _val(halo_arrays["done"])
    todo = np.flatnonzero(done == 0)
    nr_done = 0
    t0_task = time.time()
    if len(todo):
        chunk = _device_chunk(data, cfg.boxsize, device)
        H = {
            "cofp": _in_units(halo_arrays["cofp"].full, U.length)[todo],
            "search_radius": _in_units(halo_arrays["search_radius"].full, U.length)[todo],
            "read_radius": _in_units(halo_arrays["read_radius"].full, U.length)[todo],
            "index": _val(halo_arrays["index"])[todo],
            "is_central": _val(halo_arrays["is_central"])[todo],
            "nr_bound_part": _val(halo_arrays["nr_bound_part"])[todo],
        }
        res = _device_process_halos(chunk, cfg, H)
        table = res.host()
        status = res.status.cpu().numpy()
        bad = np.flatnonzero(status >= 2)
        if len(bad):
            i = int(todo[bad[0]])
            what = {_lib.HALO_COUNT_MISMATCH: "Ntot > nr_bound_part (subhalo_properties.py:2642-2646)",
                    _lib.HALO_SO_NOT_FOUND: "SO radius not found within 20 Mpc (SO_properties.py:150-153)",
                    _lib.HALO_ROOT_FAILED: "brentq bracket has no sign change (SO_properties.py:208)"}[int(status[bad[0]])]
            raise RuntimeError(f"halo index={int(_val(halo_arrays['index'])[i])}: {what}")
        c_sr, c_rr, c_nl = cols["InputHalos/search_radius"][0], cols["InputHalos/read_radius"][0], cols["InputHalos/n_loop"][0]
        for j, i in enumerate(todo):
            row = table[j]
            if status[j] == 0:
                halo_result = packer.halo_result(row)
                # the halo finder's own columns (halo_tasks.py:196-271): everything in halo_arrays that is not
                # bookkeeping goes to InputHalos/ as it came in
                for name in halo_arrays:
                    if name in ("done", "task_id", "read_radius", "search_radius"):
                        continue
                    arr = halo_arrays[name].full[i, ...]
                    if name in ("n_loop",):
                        arr = np.asarray(int(row[c_nl]))
                    halo_result[f"InputHalos/{name}"] = (arr, "No description available", True, None)
                results.append(halo_result)
                halo_arrays["done"].full[i] = 1
                nr_done += 1
            else:
                # read radius too small: larger region next time, start from the radius reached (halo_tasks.py:390-402)
                halo_arrays["read_radius"].full[i] = row[c_rr]
                halo_arrays["search_radius"].full[i] = row[c_sr]
    task_time = time.time() - t0_task
    nr_left = int(np.sum(_val(halo_arrays["done"]) == 0))
    if comm is not None:
        comm.barrier()
        nr_left = comm.allreduce(nr_left)
        nr_done = comm.allreduce(nr_done)
    return time.time() - t0_all, task_time, nr_left, nr_done, float("inf")
