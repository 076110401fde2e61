#!/usr/bin/env python
"""
Generate the golden fixtures of tests/golden/ by EXECUTING THE REFERENCE'S OWN
SOURCE for the operator-level functions of the hot path.

Run in the build container only (it reads /root/reference, which does not exist
on the GPU box):

    python tests/golden/make_golden.py

How the reference code is run here.  The reference cannot be imported as a
package in this image (unyt, mpi4py, h5py and VirgoDC are not installed), so
this script

  * parses the reference files with ``ast`` and compiles the *unmodified*
    function / class definitions it needs (never copies them into the repo);
  * supplies a minimal stand-in for the three third-party names those bodies
    touch: ``unyt`` (arrays that carry no conversion: every input below is in
    ONE consistent unit system, so every unyt conversion the bodies perform is
    a multiplication by exactly 1.0 -- positions handed to the inertia-tensor
    functions are already in kpc), ``MPI``/``comm`` (single rank: reductions are
    copies) and ``virgo.mpi.parallel_sort.parallel_sort`` (single rank: a stable
    argsort; the within-cell order it defines is NOT pinned by the reference's
    tests -- SURVEY.md 8(c) -- so fixtures keep index *sets* per cell/query).

What is pinned: find_SO_radius_and_mass (with scipy's brentq), half-weight
radius, Vmax, velocity-dispersion matrix, angular momentum, kappa_corot,
3-D and projected inertia tensors (iterative and not), SharedMesh cell
arrays and query_radius_periodic index sets.  What is not: the unit coercions
inside the four HaloProperty classes (they need a real unyt).
"""

import ast
import os
import sys
import types

import numpy as np
from scipy.optimize import brentq

REF = "/root/reference"
OUT = os.path.dirname(os.path.abspath(__file__))


# ----------------------------------------------------------------- unyt stand-in
class _Units:
    """Unit object with no dimension bookkeeping: x * units -> stand-in array."""

    registry = None
    __array_ufunc__ = None  # ndarray * units defers to __rmul__ below

    def __rmul__(self, other):
        return unyt_array(other)

    def __mul__(self, other):
        return self if isinstance(other, _Units) else unyt_array(other)

    def __pow__(self, p):
        return self

    def __truediv__(self, o):
        return self

    def __rtruediv__(self, o):
        return self if isinstance(o, _Units) else unyt_array(o)

    def __eq__(self, o):
        return isinstance(o, _Units)

    def __hash__(self):
        return 0


class unyt_array(np.ndarray):
    def __new__(cls, input_array, units=None, dtype=None, registry=None, **kw):
        return np.asarray(input_array, dtype=dtype).view(cls)

    @property
    def units(self):
        return _Units()

    @property
    def value(self):
        return np.asarray(self)

    v = value

    def to(self, units, *a, **k):
        return self

    def to_value(self, *a, **k):
        return np.asarray(self)

    in_units = to

    def __getitem__(self, idx):
        out = np.ndarray.__getitem__(self, idx)
        if not isinstance(out, np.ndarray):
            out = np.asarray(out).view(unyt_array)
        return out

    def __array_function__(self, func, types, args, kwargs):
        # numpy functions (np.linalg.norm, np.sum, ...) return stand-in arrays too, like unyt's
        def strip(x):
            if isinstance(x, unyt_array):
                return np.asarray(x)
            if isinstance(x, (list, tuple)):
                return type(x)(strip(y) for y in x)
            if isinstance(x, dict):
                return {k: strip(v) for k, v in x.items()}
            return x

        def wrap(x):
            if isinstance(x, (np.ndarray, np.generic)) and not isinstance(x, np.bool_):
                return np.asarray(x).view(unyt_array)
            if isinstance(x, tuple):
                return tuple(wrap(y) for y in x)
            return x

        return wrap(func(*strip(args), **strip(kwargs)))

    def __array_wrap__(self, arr, context=None, return_scalar=False):
        # keep 0-d results as stand-in arrays (so ``.units`` exists on scalars)
        return np.asarray(arr).view(unyt_array)


def unyt_quantity(value, units=None, dtype=None, registry=None):
    return unyt_array(value, dtype=dtype)


unyt = types.ModuleType("unyt")
unyt.unyt_array = unyt_array
unyt.unyt_quantity = unyt_quantity
unyt.dimensionless = _Units()
unyt.Unit = lambda *a, **k: _Units()
unyt.Mpc = unyt.kpc = unyt.km = unyt.s = _Units()  # radii below are in Mpc (tensors: kpc)


# ------------------------------------------------------------- MPI stand-ins
class _Comm:
    def Get_rank(self):
        return 0

    def Get_size(self):
        return 1

    def allreduce(self, x, op=None):
        return x

    def Allreduce(self, src, dst, op=None):
        dst[...] = src

    def Reduce(self, src, dst, op=None, root=0):
        dst[...] = src

    def barrier(self):
        pass


class _SharedArray:
    def __init__(self, local_shape, dtype, comm, units=None):
        self.full = unyt_array(np.zeros(local_shape, dtype=dtype))
        self.local = self.full

    def sync(self):
        pass

    def free(self):
        pass


def _parallel_sort(arr, comm=None, return_index=False):
    idx = np.argsort(arr, kind="stable")
    arr[:] = arr[idx]
    return idx if return_index else None


MPI = types.SimpleNamespace(MIN="min", MAX="max", SUM="sum")
shared_array = types.SimpleNamespace(SharedArray=_SharedArray)
ps = types.SimpleNamespace(parallel_sort=_parallel_sort)


class SearchRadiusTooSmallError(Exception):
    pass


# ------------------------------------------------- compile reference definitions
def load(relpath, names, extra=None):
    """Compile the named top-level defs of a reference file, unmodified."""
    path = os.path.join(REF, relpath)
    tree = ast.parse(open(path).read(), filename=path)
    ns = {
        "np": np, "unyt": unyt, "brentq": brentq, "MPI": MPI, "ps": ps,
        "shared_array": shared_array, "SearchRadiusTooSmallError": SearchRadiusTooSmallError,
        "Union": None, "Tuple": None, "NDArray": None, "Dict": None, "List": None,
    }
    import typing

    ns.update({k: getattr(typing, k) for k in ("Union", "Tuple", "Dict", "List")})
    from numpy.typing import NDArray

    ns["NDArray"] = NDArray
    if extra:
        ns.update(extra)
    body = [n for n in tree.body if isinstance(n, (ast.FunctionDef, ast.ClassDef)) and n.name in names]
    missing = set(names) - {n.name for n in body}
    assert not missing, f"{relpath}: {missing} not found"
    mod = ast.Module(body=body, type_ignores=[])
    exec(compile(mod, path, "exec"), ns)
    return ns


def ua(x, dtype=None):
    return unyt_array(np.array(x, dtype=dtype))


# ----------------------------------------------------------------------- cases
def nfw_profile(rng, n, conc=7.0, rmax=1.0):
    """radii of an NFW halo truncated at rmax (inverse-CDF by bisection table)"""
    x = np.linspace(1e-4, conc, 20000)
    m = np.log(1 + x) - x / (1 + x)
    u = rng.random(n) * m[-1]
    return np.interp(u, m, x) / conc * rmax


def gen_so(out):
    ns = load("SOAP/particle_selection/SO_properties.py", ["cumulative_mass_intersection", "find_SO_radius_and_mass"])
    f = ns["find_SO_radius_and_mass"]
    rng = np.random.default_rng(20261018)
    cases = []
    for ic, (n, rho_ref, neg) in enumerate(
        [(50, 200.0, False), (400, 200.0, False), (400, 2000.0, False), (3000, 60.0, False),
         (3000, 500.0, True), (30, 1e-3, False), (200, 1e9, False), (1000, 200.0, False)]
    ):
        r = np.sort(nfw_profile(rng, n))
        if ic == 7:
            r[10:14] = r[10]  # duplicate radii (SO_properties.py:186)
        mass = rng.uniform(0.5, 1.5, n).astype(np.float32) * np.float32(5e3 / n)
        if neg:
            mass[rng.random(n) < 0.05] *= np.float32(-1.0)  # neutrino-like negative weights
        cum = np.cumsum(mass, dtype=np.float64).astype(np.float32)  # SO_properties.py:400-402
        nskip = max(1, int(np.argmax(r > 0)))
        orad, cm = r[nskip:], cum[nskip:]
        dens = cm / (4.0 / 3.0 * np.pi * orad**3)
        try:
            SO_r, SO_mass, SO_vol = f(ua(orad), ua(dens), ua(cm), ua(rho_ref))
            res = np.array([float(SO_r), float(SO_mass), float(SO_vol)])
            err = 0
        except SearchRadiusTooSmallError:
            res, err = np.zeros(3), 1
        except RuntimeError:
            res, err = np.zeros(3), 2
        cases.append((orad, dens, cm, rho_ref, res, err))
    out["so_n"] = len(cases)
    for i, (orad, dens, cm, rho, res, err) in enumerate(cases):
        out[f"so{i}_r"], out[f"so{i}_dens"], out[f"so{i}_cm"] = orad, dens, cm
        out[f"so{i}_rho"], out[f"so{i}_res"], out[f"so{i}_err"] = rho, res, err


def gen_hmr(out):
    ns = load("SOAP/property_calculation/half_mass_radius.py", ["get_half_weight_radius"])
    f = ns["get_half_weight_radius"]
    rng = np.random.default_rng(7)
    k = 0
    for n in (1, 2, 3, 10, 257, 5000):
        r = rng.random(n) ** 2
        w = rng.uniform(0.1, 2.0, n).astype(np.float32)
        if n == 10:
            r[3] = r[4]
        tot = w.sum()
        out[f"hmr{k}_r"], out[f"hmr{k}_w"], out[f"hmr{k}_tot"] = r, w, tot
        out[f"hmr{k}_res"] = float(f(ua(r), ua(w), ua(tot)))
        k += 1
    # zero total weight, empty
    out[f"hmr{k}_r"], out[f"hmr{k}_w"], out[f"hmr{k}_tot"] = np.array([0.1, 0.2]), np.zeros(2, np.float32), np.float32(0)
    out[f"hmr{k}_res"] = float(f(ua([0.1, 0.2]), ua(np.zeros(2, np.float32)), ua(np.float32(0))))
    out["hmr_n"] = k + 1


def gen_kin(out):
    ns = load(
        "SOAP/property_calculation/kinematic_properties.py",
        ["get_velocity_dispersion_matrix", "get_angular_momentum",
         "get_angular_momentum_and_kappa_corot_weighted",
         "get_angular_momentum_and_kappa_corot_mass_weighted", "get_vmax"],
    )
    rng = np.random.default_rng(11)
    k = 0
    for n in (1, 5, 300, 4000):
        m = rng.uniform(0.5, 2.0, n).astype(np.float32)
        pos = rng.normal(size=(n, 3)) * 0.1
        vel = (rng.normal(size=(n, 3)) * 100).astype(np.float32)
        vel[:, 0] += (-pos[:, 1] * 800).astype(np.float32)  # net rotation about z
        vel[:, 1] += (pos[:, 0] * 800).astype(np.float32)
        mf = m / m.sum()
        vcom = (mf[:, None] * vel).sum(axis=0)
        out[f"kin{k}_m"], out[f"kin{k}_pos"], out[f"kin{k}_vel"] = m, pos, vel
        out[f"kin{k}_veldisp"] = np.asarray(ns["get_velocity_dispersion_matrix"](ua(mf), ua(vel), ua(vcom)))
        out[f"kin{k}_L"] = np.asarray(ns["get_angular_momentum"](ua(m), ua(pos), ua(vel), ref_velocity=ua(vcom)))
        L, kappa, Mcr = ns["get_angular_momentum_and_kappa_corot_mass_weighted"](
            ua(m), ua(pos), ua(vel), reference_velocity=ua(vcom), do_counterrot_mass=True
        )
        out[f"kin{k}_L2"], out[f"kin{k}_kappa"], out[f"kin{k}_Mcr"] = np.asarray(L), float(kappa), float(Mcr)
        r = np.sqrt((pos**2).sum(axis=1))
        if n == 300:
            r[:3] = 0.0  # particles at the centre are skipped (kinematic_properties.py:584-586)
        rv, vmax = ns["get_vmax"](ua(m), ua(r))
        out[f"kin{k}_r"] = r
        out[f"kin{k}_vmax"] = np.array([float(rv), float(vmax)])  # vmax = sqrt(max(cum/r)) with G = 1 here
        k += 1
    out["kin_n"] = k


def gen_tensors(out):
    ns = load(
        "SOAP/property_calculation/inertia_tensors.py",
        ["get_weighted_inertia_tensor", "get_weighted_projected_inertia_tensor"],
    )
    f3, f2 = ns["get_weighted_inertia_tensor"], ns["get_weighted_projected_inertia_tensor"]
    rng = np.random.default_rng(13)
    k = 0
    for n in (10, 19, 20, 500, 6000):
        pos = rng.normal(size=(n, 3)) * np.array([30.0, 18.0, 9.0])  # kpc, triaxial
        pos[0] = 0.0  # a particle exactly at the centre (reduced tensors drop it)
        w = rng.uniform(0.5, 2.0, n).astype(np.float32)
        out[f"ten{k}_pos"], out[f"ten{k}_w"] = pos, w
        for reduced in (False, True):
            for iters in (1, 20):
                R = 40.0
                t = f3(ua(w), ua(pos), ua(R), search_radius=ua(1e4), reduced=reduced, max_iterations=iters)
                out[f"ten{k}_3d_r{int(reduced)}_i{iters}"] = np.zeros(6) if t is None else np.asarray(t, dtype=np.float64)
                for axis in (0, 1, 2):
                    t = f2(ua(w), ua(pos), axis, ua(R), reduced=reduced, max_iterations=iters)
                    out[f"ten{k}_2d_a{axis}_r{int(reduced)}_i{iters}"] = (
                        np.zeros(3) if t is None else np.asarray(t, dtype=np.float64)
                    )
        k += 1
    out["ten_n"] = k


def gen_cyl(out):
    nsc = load("SOAP/property_calculation/cylindrical_coordinates.py",
               ["build_rotation_matrix", "calculate_cylindrical_velocities"])
    nsk = load("SOAP/property_calculation/kinematic_properties.py",
               ["get_weighted_rotation_velocity", "get_rotation_velocity_mass_weighted",
                "get_weighted_cylindrical_velocity_dispersion_vector",
                "get_cylindrical_velocity_dispersion_vector_mass_weighted"])
    rng = np.random.default_rng(19)
    k = 0
    for n, zt in ((2, [0.0, 0.0, 2.0]), (50, [1.0, 0.02, 0.0]), (400, [0.3, -0.5, 0.8]), (3000, [-1.0, 0.3, 0.1])):
        m = rng.uniform(0.5, 2.0, n).astype(np.float32)
        pos = rng.normal(size=(n, 3)) * 0.01
        pos[0] = 0.0  # a particle on the axis: phi = arctan2(0, 0)
        vel = (rng.normal(size=(n, 3)) * 80).astype(np.float32)
        zt = np.array(zt)
        vref = (m[:, None] * vel).sum(axis=0) / m.sum()
        cyl = nsc["calculate_cylindrical_velocities"](ua(pos), ua(vel), ua(zt), reference_velocity=ua(vref))
        out[f"cyl{k}_m"], out[f"cyl{k}_pos"], out[f"cyl{k}_vel"], out[f"cyl{k}_z"] = m, pos, vel, zt
        out[f"cyl{k}_vref"] = np.asarray(vref)
        out[f"cyl{k}_cyl"] = np.asarray(cyl)
        out[f"cyl{k}_R"] = np.asarray(nsc["build_rotation_matrix"](ua(zt)))
        out[f"cyl{k}_vrot"] = float(nsk["get_rotation_velocity_mass_weighted"](ua(m), ua(np.asarray(cyl)[:, 1])))
        out[f"cyl{k}_sig"] = np.asarray(nsk["get_cylindrical_velocity_dispersion_vector_mass_weighted"](ua(m), ua(np.asarray(cyl))))
        k += 1
    out["cyl_n"] = k


def gen_mesh(out):
    ns = load("SOAP/core/shared_mesh.py", ["SharedMesh"])
    SharedMesh = ns["SharedMesh"]
    rng = np.random.default_rng(17)
    k = 0
    L = 10.0
    for n, res in ((1, 4), (1000, 7), (20000, 13)):
        pos = rng.random((n, 3)) * L
        if n == 20000:
            pos[:5000] = 5.0 + rng.normal(size=(5000, 3)) * 0.15  # a clump
            pos[5000:5200, 0] = rng.random(200) * 0.05  # near the periodic edge
            pos %= L
        sa = types.SimpleNamespace(full=ua(pos), local=ua(pos))
        mesh = SharedMesh(_Comm(), sa, res)
        out[f"mesh{k}_pos"], out[f"mesh{k}_res"] = pos, res
        out[f"mesh{k}_pos_min"], out[f"mesh{k}_pos_max"] = np.asarray(mesh.pos_min), np.asarray(mesh.pos_max)
        out[f"mesh{k}_cell_size"] = np.asarray(mesh.cell_size)
        out[f"mesh{k}_cell_count"] = np.asarray(mesh.cell_count.full)
        out[f"mesh{k}_cell_offset"] = np.asarray(mesh.cell_offset.full)
        out[f"mesh{k}_sort_idx"] = np.asarray(mesh.sort_idx.full)
        # queries: interior, straddling each face, bigger than the box, empty
        centres = np.array([[5.0, 5.0, 5.0], [0.02, 5.0, 5.0], [9.99, 9.98, 0.01], [2.0, 8.0, 3.0], [5.0, 5.0, 5.0],
                            [7.3, 0.4, 9.1]])
        radii = np.array([0.5, 0.7, 1.3, 1e-6, 12.0, 3.3])
        out[f"mesh{k}_centres"], out[f"mesh{k}_radii"] = centres, radii
        for q in range(len(radii)):
            idx = mesh.query_radius_periodic(ua(centres[q]), ua(radii[q]), sa, ua(L))
            out[f"mesh{k}_q{q}"] = np.sort(np.asarray(idx, dtype=np.int64))
        k += 1
    out["mesh_n"], out["mesh_L"] = k, L


def main():
    if not os.path.isdir(REF):
        sys.exit("make_golden.py needs /root/reference (build container only)")
    for name, gen in (("so_radius", gen_so), ("half_mass_radius", gen_hmr), ("kinematics", gen_kin),
                      ("inertia_tensors", gen_tensors), ("cylindrical", gen_cyl), ("shared_mesh", gen_mesh)):
        out = {}
        gen(out)
        path = os.path.join(OUT, f"{name}.npz")
        np.savez_compressed(path, **out)
        print(f"{path}: {len(out)} arrays, {os.path.getsize(path) / 1024:.1f} KiB")


if __name__ == "__main__":
    main()
