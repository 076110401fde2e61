"""
Oracle for the per-halo loop (stage B2/B3) and the four HaloProperty classes
(stage C), restricted to the north-star property set.

Test infrastructure only -- see oracle/__init__.py.

Restates (unyt stripped; every threshold pre-converted to coordinate units):
  * process_single_halo      SOAP/core/halo_tasks.py:23-273 (radius ladder,
                             density gate, gather, halo-centred re-wrap)
  * target density           SOAP/core/halo_tasks.py:306-317
  * SOParticleData           SOAP/particle_selection/SO_properties.py:313-692,
                             2724-2790 (+ per-type moments :694-1290)
  * SubhaloParticleData      SOAP/particle_selection/subhalo_properties.py:128-376,
                             798-1127, 2265-2343, 2632-2646
  * ApertureParticleData     SOAP/particle_selection/aperture_properties.py:270-539,
                             1098-1706, 3468-3699, 4140-4143
  * ProjectedApertureParticleData
                             SOAP/particle_selection/projected_aperture_properties.py:98-198

Particle data layout: ``data[ptype]`` with ptype in (0, 1, 4, 5) ("PartType0"
...), each a dict with "Coordinates" f64[N,3], "Masses" f32[N] (the
mass_dataset: DynamicalMasses for ptype 5, SOAP/core/dataset_names.py:7),
"Velocities" f32[N,3], "GroupNr_bound" int[N], "FOFGroupIDs" int[N].

``faithful=True`` keeps the reference's dtypes (float32 masses/velocities, so
float32 pairwise sums); ``faithful=False`` promotes masses and velocities to
float64 first (same selections; the SO cumulative mass keeps its float32
rounding step, SO_properties.py:400-402, in both modes).
"""

from dataclasses import dataclass, field

import numpy as np

from .calc import (
    SearchRadiusTooSmallError,
    find_SO_radius_and_mass,
    get_half_weight_radius,
    get_velocity_dispersion_matrix,
    get_angular_momentum,
    get_angular_momentum_and_kappa_corot_mass_weighted,
    calculate_cylindrical_velocities,
    get_rotation_velocity_mass_weighted,
    get_cylindrical_velocity_dispersion_vector_mass_weighted,
    get_vmax,
    get_weighted_inertia_tensor,
    get_weighted_projected_inertia_tensor,
)

SEARCH_RADIUS_FACTOR = 1.2  # halo_tasks.py:14
READ_RADIUS_FACTOR = 1.5  # halo_tasks.py:17
PTYPES = (0, 1, 4, 5)  # SO_properties.py:3520-3550 dict order
PTYPES_FOR_SO_MASSES = (0, 1, 4, 5)  # SOAP/core/dataset_names.py:4


@dataclass
class Params:
    """Scalars the reference takes from cellgrid / unyt, in coordinate units."""

    boxsize: float
    G: float = 43.009  # newton_G such that vmax = sqrt(G*M/r) in coordinate units
    softening: dict = field(default_factory=lambda: {0: 0.0, 1: 0.0, 4: 0.0, 5: 0.0})
    nu_density: float = 0.0
    H: float = 0.0  # Hubble flow term of KineticEnergy*, velocity / coordinate length
    kpc_per_length: float = 1000.0  # inertia_tensors.py:77-78 .to("kpc")
    r_20mpc: float = 20.0
    phys_mpc_to_coord: float = 1.0  # halo_tasks.py:168 swift_mpc -> search_radius units
    critical_density: float = 1.0
    mean_density: float = 1.0
    faithful: bool = True
    # also the iterative tensors (inertia_tensors.py default max_iterations = 20)
    iterative_tensors: bool = False
    # CategoryFilter (category_filter.py:69-110): name -> (limit, particle types whose BoundSubhalo counts are summed)
    filters: dict = field(default_factory=dict)


N_KEY = {0: "Ngas", 1: "Ndm", 4: "Nstar", 5: "Nbh"}


def do_calculation(params, halo_result, name):
    """CategoryFilter.get_do_calculation (category_filter.py:69-110) for one category: "basic" is always
    computed, the others need sum(BoundSubhalo/NumberOf*Particles) >= limit."""
    if name == "basic":
        return True
    limit, types = params.filters[name]
    sub = halo_result["BoundSubhalo"]  # KeyError like the reference if BoundSubhalo was not computed first
    return sum(int(sub.get(N_KEY[t], 0)) for t in types) >= limit


def _cast(arr, faithful):
    return arr if faithful else arr.astype(np.float64)


def _wrap_mod(x, L):
    return x % L


# ------------------------------------------------------------------ shared sets


def _type_block(mass, pos, vel, radius, types, params, prefix_total=True):
    """Per-type masses, counts, first and second moments as every class defines
    them (e.g. aperture_properties.py:374-539,1098-1706).  ``pos`` is relative
    to the halo centre."""
    out = {}
    names = {0: "gas", 1: "dm", 4: "star", 5: "bh"}
    for t, nm in names.items():
        sel = types == t
        out[f"N{nm}"] = int(sel.sum())
        out[f"M{nm}"] = mass[sel].sum()
    return out


def _kin_block(mass, pos, vel, centre, params, with_kappa):
    """com, vcom, L (about vcom), veldisp matrix, kappa/Mcountrot for one
    particle group.  aperture_properties.py:1098-1270."""
    M = mass.sum()
    if M == 0:
        return None
    mf = mass / M
    out = {"M": M}
    out["com"] = ((mf[:, None] * pos).sum(axis=0) + centre) % params.boxsize
    out["vcom"] = (mf[:, None] * vel).sum(axis=0)
    out["veldisp"] = get_velocity_dispersion_matrix(mf, vel, out["vcom"])
    if with_kappa:
        L, kappa, Mcr = get_angular_momentum_and_kappa_corot_mass_weighted(
            mass, pos, vel, reference_velocity=out["vcom"], do_counterrot_mass=True
        )
        out["L"] = L
        out["kappa"] = kappa
        out["DtoT"] = 1.0 - 2.0 * Mcr / M
    else:
        out["L"] = get_angular_momentum(mass, pos, vel, ref_velocity=out["vcom"])
    return out


def _star_cylindrical(res, mass, pos, vel):
    """StellarRotationalVelocity and the cylindrical dispersions
    (aperture_properties.py:1477-1536, subhalo_properties.py:1410-1470): stars in
    the frame of their own angular momentum, velocities about vcom_star."""
    if len(mass) < 2 or "Lstar" not in res or np.sum(res["Lstar"]) == 0:
        return
    cyl = calculate_cylindrical_velocities(pos, vel, np.asarray(res["Lstar"], dtype=np.float64),
                                           reference_velocity=res["vcom_star"])
    res["StellarRotationalVelocity"] = get_rotation_velocity_mass_weighted(mass, cyl[:, 1])
    sig = get_cylindrical_velocity_dispersion_vector_mass_weighted(mass, cyl)
    res["StellarCylindricalVelocityDispersion"] = np.sqrt((sig**2).sum() / 3)
    res["StellarCylindricalVelocityDispersionVertical"] = sig[2]
    res["StellarCylindricalVelocityDispersionDiscPlane"] = np.sqrt((sig[:2] ** 2).sum())


def _store_group(res, prefix, blk, with_kappa):
    if blk is None:
        return
    res[f"com_{prefix}"] = blk["com"]
    res[f"vcom_{prefix}"] = blk["vcom"]
    res[f"veldisp_matrix_{prefix}"] = blk["veldisp"]
    res[f"L{prefix}"] = blk["L"]
    if with_kappa:
        res[f"kappa_corot_{prefix}"] = blk["kappa"]
        res[f"DtoT{prefix}"] = blk["DtoT"]


def _tensor(mass, pos, R, params, reduced, search_radius=None, max_iterations=1):
    t = get_weighted_inertia_tensor(
        mass,
        pos,
        R,
        search_radius=search_radius,
        reduced=reduced,
        max_iterations=max_iterations,
        kpc_per_length=params.kpc_per_length,
    )
    return t


# --------------------------------------------------------------------------- SO


class SOOracle:
    """SOProperties (SO_properties.py:3215-3742) for the north-star keys."""

    def __init__(self, params, SOval=200.0, type="crit", name=None, halo_filter="basic"):
        self.params = params
        self.halo_filter = halo_filter
        self.type = type
        self.mean_density_multiple = 1000.0  # SO_properties.py:3459-3462
        self.critical_density_multiple = 1000.0
        self.physical_radius_mpc = 0.0
        self.virial_definition = False
        if type == "mean":
            self.mean_density_multiple = SOval
            self.virial_definition = SOval == 200
            self.reference_density = SOval * params.mean_density
            self.name = f"SO_{SOval:.0f}_mean"
            self.group_name = f"SO/{SOval:.0f}_mean"
        elif type == "crit":
            self.critical_density_multiple = SOval
            self.virial_definition = SOval == 200
            self.reference_density = SOval * params.critical_density
            self.name = f"SO_{SOval:.0f}_crit"
            self.group_name = f"SO/{SOval:.0f}_crit"
        elif type == "BN98":
            self.critical_density_multiple = SOval  # caller passes cellgrid.virBN98
            self.virial_definition = True
            self.reference_density = SOval * params.critical_density
            self.name = "SO_BN98"
            self.group_name = "SO/BN98"
        else:
            raise AttributeError(f"Unknown SO type: {type}!")

    def calculate(self, input_halo, search_radius, data, halo_result):
        p = self.params
        res = {}
        # SO_properties.py:3627: centrals only, and only if the halo passes this variation's filter
        if not input_halo["is_central"] or not do_calculation(p, halo_result, self.halo_filter):
            halo_result[self.group_name] = res
            return
        centre = input_halo["cofp"]
        index = input_halo["index"]
        types_present = [t for t in PTYPES if t in data]
        # compute_basics, SO_properties.py:313-354
        mass, radius, position, velocity, types, groupnr, fofid, softening = (
            [] for _ in range(8)
        )
        for t in types_present:
            d = data[t]
            mass.append(_cast(d["Masses"], p.faithful))
            pos = d["Coordinates"] - centre[None, :]
            position.append(pos)
            r = np.sqrt(np.sum(pos**2, axis=1))
            radius.append(r)
            velocity.append(_cast(d["Velocities"], p.faithful))
            types.append(t * np.ones(r.shape, dtype=np.int32))
            groupnr.append(d["GroupNr_bound"])
            fofid.append(d["FOFGroupIDs"])
            softening.append(np.ones(r.shape, dtype=np.float64) * p.softening[t])
        mass = np.concatenate(mass)
        radius = np.concatenate(radius)
        position = np.concatenate(position)
        velocity = np.concatenate(velocity)
        types = np.concatenate(types)
        groupnr = np.concatenate(groupnr)
        fofid = np.concatenate(fofid)
        softening = np.concatenate(softening)

        # compute_SO_radius_and_mass, SO_properties.py:356-513
        order = np.argsort(radius)
        ordered_radius = radius[order]
        cumulative_mass = np.cumsum(mass[order], dtype=np.float64).astype(np.float32)
        cumulative_mass += p.nu_density * 4.0 / 3.0 * np.pi * ordered_radius**3
        SO_r = 0.0
        SO_mass = 0.0
        if len(order) > 0:
            cen_fofid = fofid[order[0]]
            nskip = max(1, int(np.argmax(ordered_radius > 0.0)))
            ordered_radius = ordered_radius[nskip:]
            cumulative_mass = cumulative_mass[nskip:]
        else:
            cen_fofid = -1
        nr_parts = len(ordered_radius)
        if nr_parts > 0:
            density = cumulative_mass / (4.0 / 3.0 * np.pi * ordered_radius**3)
            try:
                SO_r, SO_mass, _ = find_SO_radius_and_mass(
                    ordered_radius,
                    density,
                    cumulative_mass,
                    self.reference_density,
                    r_20mpc=p.r_20mpc,
                )
            except SearchRadiusTooSmallError:
                raise SearchRadiusTooSmallError("SO radius multiple was too small!")
        SO_exists = SO_r > 0 and SO_mass > 0
        halo_result[self.group_name] = res
        if not SO_exists:
            return
        is_sat = (groupnr >= 0) & (groupnr != index) & (fofid == cen_fofid)
        is_ext = (groupnr >= 0) & (groupnr != index) & (fofid != cen_fofid)
        # dm_missed_mass, SO_properties.py:471-482
        dm_r = radius[types == 1]
        dm_m = mass[types == 1]
        outside = dm_r > SO_r
        dm_missed_mass = 0.0
        if np.any(outside):
            inside = np.logical_not(outside)
            if np.any(inside):
                r1 = np.max(dm_r[inside])
                i = np.argmin(dm_r[outside])
                r2 = dm_r[outside][i]
                m2 = dm_m[outside][i]
                dm_missed_mass = m2 * (SO_r - r1) / (r2 - r1)
        sel = radius < SO_r  # SO_properties.py:485
        # in-sphere + "surrounding" particles (SO_properties.py:490-496,629-630)
        mass_all = np.concatenate([mass[sel], mass[~sel]])
        position_all = np.concatenate([position[sel], position[~sel]])
        mass = mass[sel]
        radius = radius[sel]
        position = position[sel]
        velocity = velocity[sel]
        types = types[sel]
        is_sat = is_sat[sel]
        is_ext = is_ext[sel]
        softening = softening[sel]

        res["r"] = SO_r
        res["Mtot"] = SO_mass
        res.update(_type_block(mass, position, velocity, radius, types, p))
        Mtotpart = mass.sum()  # SO_properties.py:531-538
        res["Mtotpart"] = Mtotpart
        if Mtotpart > 0:
            mf = mass / Mtotpart
            res["com"] = ((mf[:, None] * position).sum(axis=0) + centre) % p.boxsize
            res["vcom"] = (mf[:, None] * velocity).sum(axis=0)
            soft_r = np.maximum(softening, radius)
            r_vmax, vmax = get_vmax(mass, soft_r, p.G)
            res["R_vmax_soft"] = r_vmax
            res["Vmax_soft"] = vmax
            if vmax > 0:  # SO_properties.py:602-618
                vrel = velocity - res["vcom"][None, :]
                Ltot = np.linalg.norm(
                    (mass[:, None] * np.cross(position, vrel)).sum(axis=0)
                )
                res["spin_parameter"] = Ltot / (np.sqrt(2.0) * Mtotpart * SO_r * vmax)
            for reduced in (False, True):
                t = _tensor(mass, position, SO_r, p, reduced)
                if t is not None:
                    nm = "TotalInertiaTensor" + ("Reduced" if reduced else "")
                    res[nm + "Noniterative"] = t
            if p.iterative_tensors:  # SO_properties.py:621-648
                for reduced in (False, True):
                    t = _tensor(mass_all, position_all, SO_r, p, reduced, search_radius=search_radius,
                                max_iterations=20)
                    if t is not None:
                        res["TotalInertiaTensor" + ("Reduced" if reduced else "")] = t
        res["Mfrac_satellites"] = mass[is_sat].sum() / SO_mass
        res["Mfrac_external"] = mass[is_ext].sum() / SO_mass
        for t, nm, kap in ((0, "gas", True), (1, "dm", False), (4, "star", True)):
            s = types == t
            blk = _kin_block(mass[s], position[s], velocity[s], centre, p, kap)
            _store_group(res, nm, blk, kap)
        s = (types == 0) | (types == 4)
        blk = _kin_block(mass[s], position[s], velocity[s], centre, p, False)
        if blk is not None:
            res["Lbaryons"] = blk["L"]
        # concentration, SO_properties.py:2724-2790
        if self.virial_definition:
            def conc_from_R1(R1):
                polynomial = [-79.71, -222.46, -250.14, -140.17, -43.59, -5.07]
                c = 0
                for i, b in enumerate(polynomial[::-1]):
                    c += b * np.log10(R1) ** i
                c = max(min(c, 3), 0)
                return np.float32(10**c)

            def calc_conc(r):
                if r.shape[0] < 10:
                    return None
                R1 = np.sum(mass * r)
                missed_mass = SO_mass - np.sum(mass)
                R1 += np.pi * p.nu_density * SO_r**4
                missed_mass -= p.nu_density * 4.0 / 3.0 * np.pi * SO_r**3
                R1 += missed_mass * SO_r
                R1 /= SO_r * SO_mass
                return conc_from_R1(R1)

            def calc_conc_dmo(r):
                if r.shape[0] < 10:
                    return None
                Mdm = res["Mdm"]
                R1 = np.sum(mass[types == 1] * r)
                R1 += dm_missed_mass * SO_r
                R1 /= SO_r * (Mdm + dm_missed_mass)
                return conc_from_R1(R1)

            soft_r = np.maximum(softening, radius)
            for nm, val in (
                ("concentration_unsoft", calc_conc(radius)),
                ("concentration_soft", calc_conc(soft_r)),
                ("concentration_dmo_unsoft", calc_conc_dmo(radius[types == 1])),
                ("concentration_dmo_soft", calc_conc_dmo(soft_r[types == 1])),
            ):
                if val is not None:
                    res[nm] = val


# ---------------------------------------------------------------------- subhalo


class SubhaloOracle:
    """SubhaloProperties (subhalo_properties.py:2346-2740), north-star keys."""

    name = "BoundSubhalo"
    group_name = "BoundSubhalo"
    mean_density_multiple = None
    critical_density_multiple = None
    physical_radius_mpc = 0.0

    def __init__(self, params):
        self.params = params

    def calculate(self, input_halo, search_radius, data, halo_result):
        p = self.params
        centre = input_halo["cofp"]
        index = input_halo["index"]
        mass, position, radius, velocity, types, softening = ([] for _ in range(6))
        for t in [t for t in PTYPES if t in data]:
            d = data[t]
            in_halo = d["GroupNr_bound"] == index
            mass.append(_cast(d["Masses"], p.faithful)[in_halo])
            pos = d["Coordinates"][in_halo, :] - centre[None, :]
            position.append(pos)
            r = np.sqrt(pos[:, 0] ** 2 + pos[:, 1] ** 2 + pos[:, 2] ** 2)
            radius.append(r)
            velocity.append(_cast(d["Velocities"], p.faithful)[in_halo, :])
            types.append(t * np.ones(r.shape, dtype=np.int32))
            softening.append(np.ones(r.shape, dtype=np.float64) * p.softening[t])
        mass = np.concatenate(mass)
        position = np.concatenate(position)
        radius = np.concatenate(radius)
        velocity = np.concatenate(velocity)
        types = np.concatenate(types)
        softening = np.concatenate(softening)

        res = {}
        res.update(_type_block(mass, position, velocity, radius, types, p))
        # subhalo_properties.py:2632-2646
        Ntot = res["Ngas"] + res["Ndm"] + res["Nstar"] + res["Nbh"]
        Nexpected = input_halo["nr_bound_part"]
        if Ntot < Nexpected:
            raise SearchRadiusTooSmallError(
                "Search radius does not contain expected number of particles!"
            )
        elif Ntot > Nexpected:
            raise RuntimeError(
                f'Found more particles than expected for halo {input_halo["index"]}'
            )
        halo_result[self.group_name] = res
        Mtot = mass.sum()
        res["Mtot"] = Mtot
        if Mtot == 0:
            return
        mf = mass / Mtot
        res["com"] = ((mf[:, None] * position).sum(axis=0) + centre) % p.boxsize
        res["vcom"] = (mf[:, None] * velocity).sum(axis=0)
        # KineticEnergyTotal, subhalo_properties.py:848-858
        v_tot = velocity - res["vcom"][None, :]
        v_tot = v_tot + position * p.H
        res["KineticEnergyTotal"] = 0.5 * (mass * (v_tot**2).sum(axis=1)).sum()
        res["EncloseRadius"] = np.max(radius)
        r_vu, v_u = get_vmax(mass, radius, p.G, nskip=1)
        res["R_vmax_unsoft"], res["Vmax_unsoft"] = r_vu, v_u
        soft_r = np.maximum(softening, radius)
        r_vs, v_s = get_vmax(mass, soft_r, p.G)
        res["R_vmax_soft"], res["Vmax_soft"] = r_vs, v_s
        if r_vs > 0 and v_s > 0:  # subhalo_properties.py:1049-1073
            m = radius <= r_vs
            vrel = velocity[m, :] - res["vcom"][None, :]
            Ltot = np.linalg.norm(
                (mass[m, None] * np.cross(position[m, :], vrel)).sum(axis=0)
            )
            M_r_vmax = mass[m].sum()
            if M_r_vmax > 0:
                res["spin_parameter"] = Ltot / (np.sqrt(2.0) * M_r_vmax * v_s * r_vs)
        gas, dm, star = types == 0, types == 1, types == 4
        res["HalfMassRadiusTot"] = get_half_weight_radius(radius, mass, Mtot)
        res["HalfMassRadiusGas"] = get_half_weight_radius(
            radius[gas], mass[gas], res["Mgas"]
        )
        res["HalfMassRadiusDM"] = get_half_weight_radius(radius[dm], mass[dm], res["Mdm"])
        res["HalfMassRadiusStar"] = get_half_weight_radius(
            radius[star], mass[star], res["Mstar"]
        )
        res["HalfMassRadiusBaryon"] = get_half_weight_radius(
            radius[gas | star], mass[gas | star], res["Mgas"] + res["Mstar"]
        )
        for t, nm, kap in ((0, "gas", True), (1, "dm", False), (4, "star", True)):
            s = types == t
            blk = _kin_block(mass[s], position[s], velocity[s], centre, p, kap)
            _store_group(res, nm, blk, kap)
        _star_cylindrical(res, mass[star], position[star], velocity[star])
        s = gas | star
        blk = _kin_block(mass[s], position[s], velocity[s], centre, p, True)
        if blk is not None:
            res["Lbaryons"] = blk["L"]
            res["kappa_corot_baryons"] = blk["kappa"]
        for reduced in (False, True):
            t = _tensor(mass, position, 10 * res["HalfMassRadiusTot"], p, reduced)
            if t is not None:
                res["TotalInertiaTensor" + ("Reduced" if reduced else "") + "Noniterative"] = t
            if p.iterative_tensors:  # subhalo_properties.py:1075-1100
                t = _tensor(mass, position, 10 * res["HalfMassRadiusTot"], p, reduced, max_iterations=20)
                if t is not None:
                    res["TotalInertiaTensor" + ("Reduced" if reduced else "")] = t


# -------------------------------------------------------------------- apertures


class ApertureOracle:
    """ExclusiveSphere/InclusiveSphere (aperture_properties.py:3702-4398)."""

    mean_density_multiple = None
    critical_density_multiple = None

    def __init__(self, params, aperture_radius, physical_radius_mpc, inclusive, label, halo_filter="basic",
                 prev_radius=None, prev_group=None):
        self.params = params
        self.halo_filter = halo_filter
        # all_radii_kpc[i_radius - 1] in coordinate units and the group of that aperture, or None when this is
        # the first radius / the shortcut is off (compute_halo_properties.py:345-395)
        self.prev_radius = prev_radius
        self.prev_group = prev_group
        self.aperture_radius = aperture_radius  # coordinate units (unyt-converted)
        self.physical_radius_mpc = physical_radius_mpc
        self.inclusive = inclusive
        self.name = ("inclusive" if inclusive else "exclusive") + f"_sphere_{label}"
        self.group_name = ("InclusiveSphere/" if inclusive else "ExclusiveSphere/") + label

    def calculate(self, input_halo, search_radius, data, halo_result):
        p = self.params
        # aperture_properties.py:4082-4123: the previous aperture already held every bound particle ->
        # inclusive: skipped (zeros); exclusive: the previous exclusive values are copied
        do_calc = do_calculation(p, halo_result, self.halo_filter)
        if self.prev_radius is not None and "EncloseRadius" in halo_result.get("BoundSubhalo", {}):
            r_enclose = np.float32(halo_result["BoundSubhalo"]["EncloseRadius"])  # the stored output is float32
            if self.prev_radius > r_enclose:
                res = {}
                if not self.inclusive and do_calc:
                    res = dict(halo_result[self.prev_group])
                halo_result[self.group_name] = res
                return
        # aperture_properties.py:4127: a halo that fails this variation's filter keeps its zeros and
        # never asks for a larger radius
        if not do_calc:
            halo_result[self.group_name] = {}
            return
        # aperture_properties.py:4140-4143
        if search_radius < self.aperture_radius:
            raise SearchRadiusTooSmallError("Search radius is smaller than aperture")
        centre = input_halo["cofp"]
        index = input_halo["index"]
        mass, position, radius, velocity, types, softening = ([] for _ in range(6))
        star_mass_all = np.zeros(0)
        star_pos_all = np.zeros((0, 3))
        for t in [t for t in PTYPES if t in data]:
            d = data[t]
            grnr = d["GroupNr_bound"]
            in_halo = np.ones(grnr.shape, dtype=bool) if self.inclusive else grnr == index
            mass.append(_cast(d["Masses"], p.faithful)[in_halo])
            pos = d["Coordinates"][in_halo, :] - centre[None, :]
            position.append(pos)
            r = np.sqrt(pos[:, 0] ** 2 + pos[:, 1] ** 2 + pos[:, 2] ** 2)
            radius.append(r)
            velocity.append(_cast(d["Velocities"], p.faithful)[in_halo, :])
            types.append(t * np.ones(r.shape, dtype=np.int32))
            softening.append(np.ones(r.shape, dtype=np.float64) * p.softening[t])
            if t == 4:
                star_mass_all = mass[-1]
                star_pos_all = pos
        mass = np.concatenate(mass)
        position = np.concatenate(position)
        radius = np.concatenate(radius)
        velocity = np.concatenate(velocity)
        types = np.concatenate(types)
        softening = np.concatenate(softening)
        mask = radius <= self.aperture_radius  # aperture_properties.py:310
        mass, position, velocity = mass[mask], position[mask], velocity[mask]
        radius, types, softening = radius[mask], types[mask], softening[mask]

        res = {}
        halo_result[self.group_name] = res
        res.update(_type_block(mass, position, velocity, radius, types, p))
        Mtot = mass.sum()
        res["Mtot"] = Mtot
        if Mtot > 0:
            mf = mass / Mtot
            res["com"] = ((mf[:, None] * position).sum(axis=0) + centre) % p.boxsize
            res["vcom"] = (mf[:, None] * velocity).sum(axis=0)
            soft_r = np.maximum(softening, radius)
            res["R_vmax_soft"], res["Vmax_soft"] = get_vmax(mass, soft_r, p.G)
        gas, dm, star = types == 0, types == 1, types == 4
        res["HalfMassRadiusGas"] = get_half_weight_radius(radius[gas], mass[gas], res["Mgas"])
        res["HalfMassRadiusDM"] = get_half_weight_radius(radius[dm], mass[dm], res["Mdm"])
        res["HalfMassRadiusStar"] = get_half_weight_radius(
            radius[star], mass[star], res["Mstar"]
        )
        res["HalfMassRadiusBaryon"] = get_half_weight_radius(
            radius[gas | star], mass[gas | star], res["Mgas"] + res["Mstar"]
        )
        for t, nm, kap in ((0, "gas", True), (1, "dm", False), (4, "star", True)):
            s = types == t
            blk = _kin_block(mass[s], position[s], velocity[s], centre, p, kap)
            _store_group(res, nm, blk, kap)
        _star_cylindrical(res, mass[star], position[star], velocity[star])
        s = gas | star
        blk = _kin_block(mass[s], position[s], velocity[s], centre, p, True)
        if blk is not None:
            res["Lbaryons"] = blk["L"]
            res["kappa_corot_baryons"] = blk["kappa"]
        # stellar inertia tensors over ALL stars in the halo mask,
        # aperture_properties.py:3579-3594
        for reduced in (False, True):
            if res["Mstar"] == 0:
                continue
            t = _tensor(star_mass_all, star_pos_all, self.aperture_radius, p, reduced)
            if t is not None:
                res["StellarInertiaTensor" + ("Reduced" if reduced else "") + "Noniterative"] = t
            if p.iterative_tensors:  # aperture_properties.py:3613-3636
                t = _tensor(star_mass_all, star_pos_all, self.aperture_radius, p, reduced, max_iterations=20)
                if t is not None:
                    res["StellarInertiaTensor" + ("Reduced" if reduced else "")] = t


class ProjectedApertureOracle:
    """ProjectedApertureProperties (projected_aperture_properties.py:1580-2002)."""

    mean_density_multiple = None
    critical_density_multiple = None

    def __init__(self, params, aperture_radius, physical_radius_mpc, label, halo_filter="basic", prev_radius=None,
                 prev_group=None):
        self.params = params
        self.halo_filter = halo_filter
        self.prev_radius = prev_radius
        self.prev_group = prev_group
        self.aperture_radius = aperture_radius
        self.physical_radius_mpc = physical_radius_mpc
        self.name = f"projected_aperture_{label}"
        self.group_name = f"ProjectedAperture/{label}"

    def calculate(self, input_halo, search_radius, data, halo_result):
        p = self.params
        # projected_aperture_properties.py:1826-1885: previous projected aperture held every bound particle ->
        # its values are copied (whatever the filter says)
        if self.prev_radius is not None and "EncloseRadius" in halo_result.get("BoundSubhalo", {}):
            if self.prev_radius > np.float32(halo_result["BoundSubhalo"]["EncloseRadius"]):
                for ax in "xyz":
                    halo_result[f"{self.group_name}/proj{ax}"] = dict(halo_result[f"{self.prev_group}/proj{ax}"])
                return
        # projected_aperture_properties.py:1888
        if not do_calculation(p, halo_result, self.halo_filter):
            for ax in "xyz":
                halo_result[f"{self.group_name}/proj{ax}"] = {}
            return
        centre = input_halo["cofp"]
        index = input_halo["index"]
        mass, position, velocity, types = ([] for _ in range(4))
        for t in [t for t in PTYPES if t in data]:
            d = data[t]
            in_halo = d["GroupNr_bound"] == index
            mass.append(_cast(d["Masses"], p.faithful)[in_halo])
            position.append(d["Coordinates"][in_halo, :] - centre[None, :])
            velocity.append(_cast(d["Velocities"], p.faithful)[in_halo, :])
            types.append(t * np.ones(mass[-1].shape, dtype=np.int32))
        mass = np.concatenate(mass)
        position = np.concatenate(position)
        velocity = np.concatenate(velocity)
        types = np.concatenate(types)
        rproj = [
            np.sqrt(position[:, 1] ** 2 + position[:, 2] ** 2),
            np.sqrt(position[:, 0] ** 2 + position[:, 2] ** 2),
            np.sqrt(position[:, 0] ** 2 + position[:, 1] ** 2),
        ]
        for iproj, projname in enumerate(("projx", "projy", "projz")):
            res = {}
            halo_result[f"{self.group_name}/{projname}"] = res
            m = rproj[iproj] <= self.aperture_radius
            pm, pp, pv, pr, pt = mass[m], position[m], velocity[m], rproj[iproj][m], types[m]
            res.update(_type_block(pm, pp, pv, pr, pt, p))
            Mtot = pm.sum()
            res["Mtot"] = Mtot
            if Mtot > 0:
                mf = pm / Mtot
                res["com"] = ((mf[:, None] * pp).sum(axis=0) + centre) % p.boxsize
                res["vcom"] = (mf[:, None] * pv).sum(axis=0)
            for t, nm in ((0, "gas"), (1, "dm"), (4, "star")):
                s = pt == t
                Mt = pm[s].sum()
                if Mt > 0:
                    # projected_aperture_properties.py:865-875
                    mft = pm[s] / Mt
                    vt = pv[s, iproj]
                    vcom = (mft * vt).sum()
                    res[f"proj_veldisp_{nm}"] = np.sqrt((mft * (vt - vcom) ** 2).sum())
                res[f"HalfMassRadius{nm.capitalize()}"] = get_half_weight_radius(
                    pr[s], pm[s], Mt
                )
            for reduced in (False, True):
                if Mtot == 0:
                    continue
                # all bound particles, projected_aperture_properties.py:789-852
                t = get_weighted_projected_inertia_tensor(
                    mass, position, iproj, self.aperture_radius, reduced=reduced,
                    max_iterations=1, kpc_per_length=p.kpc_per_length,
                )
                if t is not None:
                    res["ProjectedTotalInertiaTensor" + ("Reduced" if reduced else "") + "Noniterative"] = t
                if p.iterative_tensors:  # projected_aperture_properties.py:789-852
                    t = get_weighted_projected_inertia_tensor(
                        mass, position, iproj, self.aperture_radius, reduced=reduced,
                        max_iterations=20, kpc_per_length=p.kpc_per_length,
                    )
                    if t is not None:
                        res["ProjectedTotalInertiaTensor" + ("Reduced" if reduced else "")] = t


# -------------------------------------------------------------------- halo loop


def target_density_of(halo_prop_list, params):
    """halo_tasks.py:306-317."""
    target_density = None
    for halo_prop in halo_prop_list:
        if halo_prop.mean_density_multiple is not None:
            density = halo_prop.mean_density_multiple * params.mean_density
            if target_density is None or density < target_density:
                target_density = density
        if halo_prop.critical_density_multiple is not None:
            density = halo_prop.critical_density_multiple * params.critical_density
            if target_density is None or density < target_density:
                target_density = density
    return target_density


def process_single_halo(mesh, data, halo_prop_list, params, input_halo, target_density):
    """halo_tasks.py:23-273 without timing and packaging.

    ``mesh`` maps ptype -> MeshOracle.  Returns (halo_result or None, info) where
    info holds n_loop, the accepted radius and the per-ptype index arrays."""
    boxsize = params.boxsize
    halo_prop_done = np.zeros(len(halo_prop_list), dtype=bool)
    halo_result = {}
    info = {"n_loop": 0}
    current_radius = input_halo["search_radius"]
    while True:
        info["n_loop"] += 1
        assert current_radius <= input_halo["read_radius"]
        mass_total = 0.0
        idx = {}
        for ptype in data:
            pos = data[ptype]["Coordinates"]
            idx[ptype] = mesh[ptype].query_radius_periodic(
                input_halo["cofp"], current_radius, pos, boxsize
            )
            if ptype in PTYPES_FOR_SO_MASSES:
                mass_total += np.sum(data[ptype]["Masses"][idx[ptype]], dtype=float)
        density = mass_total / (4.0 / 3.0 * np.pi * current_radius**3)
        max_physical_radius_mpc = 0.0
        if target_density is None or density <= target_density:
            particle_data = {}
            for ptype in data:
                particle_data[ptype] = {}
                for name in data[ptype]:
                    particle_data[ptype][name] = data[ptype][name][idx[ptype], ...]
            for ptype in particle_data:
                pos = particle_data[ptype]["Coordinates"]
                offset = input_halo["cofp"] - 0.5 * boxsize
                pos[:, :] = ((pos - offset) % boxsize) + offset
            for prop_nr, halo_prop in enumerate(halo_prop_list):
                if halo_prop_done[prop_nr]:
                    continue
                try:
                    halo_prop.calculate(
                        input_halo, current_radius, particle_data, halo_result
                    )
                except SearchRadiusTooSmallError:
                    max_physical_radius_mpc = max(
                        max_physical_radius_mpc, halo_prop.physical_radius_mpc
                    )
                    break
                else:
                    halo_prop_done[prop_nr] = True
            if np.all(halo_prop_done):
                info["radius"] = current_radius
                info["idx"] = idx
                break
        search_radius = input_halo["search_radius"]
        required_radius = max_physical_radius_mpc * params.phys_mpc_to_coord  # halo_tasks.py:168
        if required_radius > input_halo["read_radius"]:
            input_halo["search_radius"] = max(search_radius, required_radius)
            return None, info
        elif current_radius >= input_halo["read_radius"]:
            input_halo["search_radius"] = max(search_radius, current_radius)
            return None, info
        else:
            current_radius = min(
                current_radius * SEARCH_RADIUS_FACTOR, input_halo["read_radius"]
            )
            current_radius = max(current_radius, required_radius)
    return halo_result, info
