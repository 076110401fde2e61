"""
Oracle for the pure property_calculation functions (operator API).

Test infrastructure only -- see oracle/__init__.py.

Restates (unyt stripped, thresholds pre-converted):
  * cumulative_mass_intersection / find_SO_radius_and_mass
                           SOAP/particle_selection/SO_properties.py:50-217
  * get_half_weight_radius SOAP/property_calculation/half_mass_radius.py:16-97
  * get_velocity_dispersion_matrix
                           SOAP/property_calculation/kinematic_properties.py:91-127
  * get_angular_momentum   kinematic_properties.py:222-263
  * get_angular_momentum_and_kappa_corot_weighted (mass weighted branch)
                           kinematic_properties.py:266-425
  * get_vmax               kinematic_properties.py:555-593
  * get_weighted_inertia_tensor
                           SOAP/property_calculation/inertia_tensors.py:19-132
  * get_weighted_projected_inertia_tensor
                           inertia_tensors.py:226-343
  * build_rotation_matrix / calculate_cylindrical_velocities
                           SOAP/property_calculation/cylindrical_coordinates.py:13-93
  * get_rotation_velocity_mass_weighted,
    get_cylindrical_velocity_dispersion_vector_mass_weighted
                           kinematic_properties.py:17-51,130-178
"""

import numpy as np
from scipy.optimize import brentq


class SearchRadiusTooSmallError(Exception):
    """SOAP/particle_selection/halo_properties.py:4."""


# --------------------------------------------------------------------------- SO


def cumulative_mass_intersection(r, rho_dim, slope_dim):
    """SO_properties.py:50-77."""
    return 4.0 * np.pi / 3.0 * rho_dim * r**3 - slope_dim * r + slope_dim - 1.0


def find_SO_radius_and_mass(
    ordered_radius, density, cumulative_mass, reference_density, r_20mpc=20.0
):
    """SO_properties.py:80-217.  ``r_20mpc`` is 20 Mpc in coordinate units."""
    above_mask = density > reference_density
    if above_mask[0]:
        below_mask = ~above_mask
        i = int(np.argmax(below_mask))
        if i == 0:
            if ordered_radius[-1] > r_20mpc:
                raise RuntimeError(
                    "Cannot find SO radius, but search radius is already larger than 20 Mpc!"
                )
            raise SearchRadiusTooSmallError("SO radius multiple estimate was too small!")
    else:
        ipos = 0
        while ipos < len(cumulative_mass) and cumulative_mass[ipos] < 0.0:
            ipos += 1
        if ipos == len(cumulative_mass):
            raise RuntimeError("Should never happen!")
        SO_r = np.sqrt(
            0.75
            * cumulative_mass[ipos]
            / (np.pi * ordered_radius[ipos] * reference_density)
        )
        SO_mass = cumulative_mass[ipos] * SO_r / ordered_radius[ipos]
        return SO_r, SO_mass, 4.0 * np.pi / 3.0 * SO_r**3

    r1 = ordered_radius[i - 1]
    r2 = ordered_radius[i]
    M1 = cumulative_mass[i - 1]
    M2 = cumulative_mass[i]
    while r1 == r2 or (above_mask[i - 1] == above_mask[i]):
        i += 1
        if i >= len(density):
            if ordered_radius[-1] > r_20mpc:
                raise RuntimeError(
                    "Cannot find SO radius, but search radius is already larger than 20 Mpc!"
                )
            raise SearchRadiusTooSmallError("SO radius multiple estimate was too small!")
        r1 = r2
        r2 = ordered_radius[i]
        M1 = M2
        M2 = cumulative_mass[i]

    rho_dim = reference_density * r1**3 / M1
    slope_dim = (M2 - M1) / (r2 - r1) * (r1 / M1)
    SO_r = r1 * brentq(
        cumulative_mass_intersection, 1.0, r2 / r1, args=(rho_dim, slope_dim)
    )
    SO_volume = 4.0 / 3.0 * np.pi * SO_r**3
    SO_mass = SO_volume * reference_density
    return SO_r, SO_mass, SO_volume


# -------------------------------------------------------------- half-mass radius


def get_half_weight_radius(radius, weights, total_weight):
    """half_mass_radius.py:16-97."""
    if total_weight == 0.0 or len(weights) < 1:
        return 0.0
    target_weight = 0.5 * total_weight
    isort = np.argsort(radius)
    sorted_radius = radius[isort]
    cumulative_weights = weights[isort].cumsum(dtype=np.float64)
    if cumulative_weights[-1] < 0.998 * total_weight:
        raise RuntimeError("Weights sum up to less than the given total weight value")
    ihalf = int(np.argmax(cumulative_weights >= target_weight))
    if ihalf == 0:
        rmin = 0.0
        WeightMin = 0.0
    else:
        rmin = sorted_radius[ihalf - 1]
        WeightMin = cumulative_weights[ihalf - 1]
    rmax = sorted_radius[ihalf]
    WeightMax = cumulative_weights[ihalf]
    if WeightMin == WeightMax:
        half_weight_radius = 0.5 * (rmin + rmax)
    else:
        half_weight_radius = rmin + (target_weight - WeightMin) / (
            WeightMax - WeightMin
        ) * (rmax - rmin)
    if half_weight_radius > sorted_radius[-1]:
        raise RuntimeError("Half weight radius larger than input radii")
    return half_weight_radius


# -------------------------------------------------------------------- kinematics


def get_velocity_dispersion_matrix(mass_fraction, velocity, ref_velocity):
    """kinematic_properties.py:91-127 (float32 result, as the reference)."""
    result = np.zeros(6, dtype=np.float32)
    vrel = velocity - ref_velocity[None, :]
    result[0] += (mass_fraction * vrel[:, 0] * vrel[:, 0]).sum()
    result[1] += (mass_fraction * vrel[:, 1] * vrel[:, 1]).sum()
    result[2] += (mass_fraction * vrel[:, 2] * vrel[:, 2]).sum()
    result[3] += (mass_fraction * vrel[:, 0] * vrel[:, 1]).sum()
    result[4] += (mass_fraction * vrel[:, 0] * vrel[:, 2]).sum()
    result[5] += (mass_fraction * vrel[:, 1] * vrel[:, 2]).sum()
    return result


def get_angular_momentum(mass, position, velocity, ref_position=None, ref_velocity=None):
    """kinematic_properties.py:222-263."""
    prel = position if ref_position is None else position - ref_position[None, :]
    vrel = velocity if ref_velocity is None else velocity - ref_velocity[None, :]
    return (mass[:, None] * np.cross(prel, vrel)).sum(axis=0)


def get_angular_momentum_and_kappa_corot_mass_weighted(
    particle_masses,
    particle_positions,
    particle_velocities,
    reference_position=None,
    reference_velocity=None,
    do_counterrot_mass=False,
):
    """kinematic_properties.py:266-456, mass-weighted branch (weights None)."""
    prel = (
        particle_positions
        if reference_position is None
        else particle_positions - reference_position[None, :]
    )
    vrel = (
        particle_velocities
        if reference_velocity is None
        else particle_velocities - reference_velocity[None, :]
    )
    Lpart = particle_masses[:, None] * np.cross(prel, vrel)
    weights = np.ones(1)
    Ltot = 1 * (weights[:, None] * Lpart).sum(axis=0)
    Lnrm = np.linalg.norm(Ltot)
    kappa_corot = np.float32(0.0)
    M_counterrot = np.float32(0.0)
    if Lnrm > 0.0:
        K = 0.5 * (particle_masses[:, None] * vrel**2).sum()
        if K > 0.0 or do_counterrot_mass:
            Ldir = Ltot / Lnrm
            Li = (Lpart * Ldir[None, :]).sum(axis=1)
        if K > 0.0:
            r2 = prel[:, 0] ** 2 + prel[:, 1] ** 2 + prel[:, 2] ** 2
            rdotL = (prel * Ldir[None, :]).sum(axis=1)
            Ri2 = r2 - rdotL**2
            mask = Ri2 == 0.0
            Ri2[mask] = 1.0
            Krot = 0.5 * (Li**2 / (particle_masses * Ri2))
            Kcorot = Krot[(~mask) & (Li > 0.0)].sum()
            kappa_corot = np.float32(kappa_corot + Kcorot / K)
        if do_counterrot_mass:
            M_counterrot = np.float32(
                M_counterrot + particle_masses[Li < 0.0].sum()
            )
    if do_counterrot_mass:
        return Ltot, kappa_corot, M_counterrot
    return Ltot, kappa_corot


def get_vmax(mass, radius, G, nskip=0):
    """kinematic_properties.py:555-593.  Returns (r_vmax, vmax)."""
    isort = np.argsort(radius)
    ordered_radius = radius[isort]
    cumulative_mass = mass[isort].cumsum()
    nskip = max(nskip, int(np.argmin(np.isclose(ordered_radius, 0.0))))
    ordered_radius = ordered_radius[nskip:]
    if len(ordered_radius) == 0 or ordered_radius[0] == 0:
        return 0.0, 0.0
    cumulative_mass = cumulative_mass[nskip:]
    v_over_G = cumulative_mass / ordered_radius
    imax = int(np.argmax(v_over_G))
    return ordered_radius[imax], np.sqrt(v_over_G[imax] * G)


# --------------------------------------------------------------- inertia tensors


def get_weighted_inertia_tensor(
    particle_weights,
    particle_positions,
    sphere_radius,
    search_radius=None,
    reduced=False,
    max_iterations=20,
    min_particles=20,
    kpc_per_length=1.0,
):
    """inertia_tensors.py:19-132.  ``kpc_per_length`` is the factor unyt's
    ``.to("kpc")`` applies to positions and to ``sphere_radius`` (:77-78)."""
    if particle_weights.shape[0] < min_particles:
        return None
    if reduced:
        norm = np.linalg.norm(particle_positions, axis=1) ** 2
        mask = np.logical_not(np.isclose(norm, 0))
        norm = norm[mask]
        particle_weights = particle_weights[mask]
        particle_positions = particle_positions[mask]
    tol = 0.0001
    q = 1000
    R = sphere_radius * kpc_per_length
    particle_positions = particle_positions * kpc_per_length
    if reduced:
        # the reference keeps ``norm`` in the original units squared while the
        # positions are now in kpc (inertia_tensors.py:62-78,117-118); unyt
        # carries the unit ratio to the final dimensionless conversion, which
        # numerically equals dividing by the norm expressed in kpc^2.
        norm = norm * kpc_per_length**2
    eig_val = [1, 1, 1]
    eig_vec = np.array([[1, 0, 0], [0, 1, 0], [0, 0, 1]])
    tensor = None
    for i_iter in range(max_iterations):
        old_q = q
        q = np.sqrt(eig_val[1] / eig_val[2])
        s = np.sqrt(eig_val[0] / eig_val[2])
        p = np.sqrt(eig_val[0] / eig_val[1])
        if abs((old_q - q) / q) < tol:
            break
        axis = R * np.array([1 * np.cbrt(s * p), 1 * np.cbrt(q / p), 1 / np.cbrt(q * s)])
        p = np.dot(particle_positions, eig_vec) / axis
        r = np.linalg.norm(p, axis=1)
        if (i_iter == 0) and (np.sum(r <= 1) < min_particles):
            return None
        weight = particle_weights / np.sum(particle_weights[r <= 1])
        weight[r > 1] = 0
        if (search_radius is not None) and (
            np.max(R) > search_radius * kpc_per_length
        ):
            raise SearchRadiusTooSmallError("Inertia tensor required more particles")
        tensor = (
            weight[:, None, None]
            * particle_positions[:, :, None]
            * particle_positions[:, None, :]
        )
        if reduced:
            tensor /= norm[:, None, None]
        tensor = tensor.sum(axis=0)
        eig_val, eig_vec = np.linalg.eigh(tensor)
        eig_val = np.abs(eig_val)
        if q == 0:
            tensor.fill(0)
            break
    return np.concatenate([np.diag(tensor), tensor[np.triu_indices(3, 1)]])


def get_weighted_projected_inertia_tensor(
    particle_weights,
    particle_positions,
    axis,
    radius,
    reduced=False,
    max_iterations=20,
    min_particles=20,
    kpc_per_length=1.0,
):
    """inertia_tensors.py:226-343."""
    if particle_weights.shape[0] < min_particles:
        return None
    projected_position = np.zeros(
        (particle_positions.shape[0], 2), dtype=particle_positions.dtype
    )
    if axis == 0:
        projected_position[:, 0] = particle_positions[:, 1]
        projected_position[:, 1] = particle_positions[:, 2]
    elif axis == 1:
        projected_position[:, 0] = particle_positions[:, 2]
        projected_position[:, 1] = particle_positions[:, 0]
    elif axis == 2:
        projected_position[:, 0] = particle_positions[:, 0]
        projected_position[:, 1] = particle_positions[:, 1]
    else:
        raise AttributeError(f"Invalid axis: {axis}!")
    if reduced:
        norm = np.linalg.norm(projected_position, axis=1) ** 2
        mask = np.logical_not(np.isclose(norm, 0))
        norm = norm[mask]
        particle_weights = particle_weights[mask]
        projected_position = projected_position[mask]
    tol = 0.0001
    q = 1000
    R = radius * kpc_per_length
    projected_position = projected_position * kpc_per_length
    if reduced:
        norm = norm * kpc_per_length**2  # see get_weighted_inertia_tensor
    eig_val = [1, 1]
    eig_vec = np.array([[1, 0], [0, 1]])
    tensor = None
    for i_iter in range(max_iterations):
        old_q = q
        q = np.sqrt(eig_val[0] / eig_val[1])
        if abs((old_q - q) / q) < tol:
            break
        ax = R * np.array([1 * np.sqrt(q), 1 / np.sqrt(q)])
        p = np.dot(projected_position, eig_vec) / ax
        r = np.linalg.norm(p, axis=1)
        if (i_iter == 0) and (np.sum(r <= 1) < min_particles):
            return None
        weight = particle_weights / np.sum(particle_weights[r <= 1])
        weight[r > 1] = 0
        tensor = (
            weight[:, None, None]
            * projected_position[:, :, None]
            * projected_position[:, None, :]
        )
        if reduced:
            tensor /= norm[:, None, None]
        tensor = tensor.sum(axis=0)
        eig_val, eig_vec = np.linalg.eigh(tensor)
        if q == 0:
            tensor.fill(0)
            break
    return np.concatenate([np.diag(tensor), [tensor[(0, 1)]]])


# ------------------------------------------------------- cylindrical velocities


def build_rotation_matrix(z_target):
    """cylindrical_coordinates.py:13-42."""
    z_axis = z_target / np.linalg.norm(z_target)
    helper = np.array([1, 0, 0])
    if np.allclose(z_axis, helper / np.linalg.norm(helper), rtol=0.1):
        helper = np.array([0, 1, 0])
    x_axis = np.cross(helper, z_axis)
    x_axis /= np.linalg.norm(x_axis)
    y_axis = np.cross(z_axis, x_axis)
    return np.vstack([x_axis, y_axis, z_axis])


def calculate_cylindrical_velocities(
    positions, velocities, z_target, reference_position=None, reference_velocity=None
):
    """cylindrical_coordinates.py:45-93: (v_r, v_phi, v_z) per particle."""
    prel = positions if reference_position is None else positions - reference_position
    vrel = velocities if reference_velocity is None else velocities - reference_velocity
    R = build_rotation_matrix(z_target)
    positions_rot = prel @ R.T
    velocities_rot = vrel @ R.T
    x, y = positions_rot[:, 0], positions_rot[:, 1]
    vx, vy, vz = velocities_rot[:, 0], velocities_rot[:, 1], velocities_rot[:, 2]
    phi = np.arctan2(y, x)
    v_r = vx * np.cos(phi) + vy * np.sin(phi)
    v_phi = -vx * np.sin(phi) + vy * np.cos(phi)
    return np.stack([v_r, v_phi, vz], axis=1)


def get_rotation_velocity_mass_weighted(particle_masses, particle_azimuthal_velocities):
    """kinematic_properties.py:17-51."""
    mass_weights = particle_masses / particle_masses.sum()
    return (mass_weights * particle_azimuthal_velocities).sum()


def get_cylindrical_velocity_dispersion_vector_mass_weighted(
    particle_masses, particle_cylindrical_velocities
):
    """kinematic_properties.py:130-178: [sigma_r, sigma_phi, sigma_z]."""
    w = particle_masses / particle_masses.sum()
    mean_velocity = (w[:, None] * particle_cylindrical_velocities).sum(axis=0)
    squared = (w[:, None] * (particle_cylindrical_velocities - mean_velocity) ** 2).sum(axis=0)
    return np.sqrt(squared)
