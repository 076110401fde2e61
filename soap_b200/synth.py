"""
Deterministic synthetic particle/halo chunks for tests and bench.py
(SURVEY.md 8(d) "Concrete synthetic inputs").

Two families:

* ``dummy_chunk``   -- halos drawn like the reference's test fixture
  ``DummyHaloGenerator.get_random_halo`` (tests/dummy_halo_generator.py:866-945:
  exponential radii of scale 1/60, first particle at r=0, types
  p=[.2,.4,.39,.01], 60 % bound to the halo, 10 % unbound, coordinates f64,
  masses/velocities f32, membership int32), concatenated into one chunk.
  The reference's legacy ``np.random`` call sequence is NOT replayed (its hydro
  field draws are outside this path); each halo uses ``default_rng(seed+i)``.
* ``nfw_chunk``     -- BASELINE config 2 recipe (and its hydro variant for
  configs 3/4): NFW halos (inverse-CDF of mu(x)=ln(1+x)-x/(1+x), the function
  ``DummyHaloGenerator.rnfw`` inverts with Lambert W,
  tests/dummy_halo_generator.py:636-643) from a power-law mass function on a
  uniform background, written with torch so the 512^3 case can be generated
  on the device.

Constants follow ``DummySnapshot`` (tests/dummy_halo_generator.py:40-126).
"""

import math

import numpy as np

# DummySnapshot / DummyCellGrid constants (tests/dummy_halo_generator.py:40-126)
SCALE_FACTOR = 0.76923077
CRITICAL_DENSITY = 17.58736923  # internal (physical) units: 1e10 Msun / Mpc^3
OMEGA_M = 0.304611
OMEGA_K = 2.5212783e-09
OMEGA_LAMBDA = 0.693922
H_INTERNAL = 79.60499176
SOFTENING_PHYS = min(0.0446 * SCALE_FACTOR, 0.0114)
# newton_G in (Mpc, 1e10 Msun, km/s): 6.6743e-8 cgs * U_M / U_L / (U_L/U_t)^2
NEWTON_G = 6.6743e-08 * 1.98841e43 / 3.08567758e24 / (3.08567758e24 / 3.08567758e19) ** 2


def virBN98(a=SCALE_FACTOR):
    """DummyCellGrid.__init__ (tests/dummy_halo_generator.py:463-470)."""
    bnx = -(OMEGA_K / a**2 + OMEGA_LAMBDA) / (
        OMEGA_K / a**2 + OMEGA_M / a**3 + OMEGA_LAMBDA
    )
    return 18.0 * np.pi**2 + 82.0 * bnx - 39.0 * bnx**2


def coordinate_unit_params(boxsize, a=SCALE_FACTOR):
    """The scalar thresholds the host adapter hands to the kernels, i.e. what
    unyt would produce at each mixed-unit operation (SURVEY 8(c) detail 11),
    for coordinates in comoving snap_length and snap_length == Mpc."""
    return dict(
        boxsize=float(boxsize),
        # densities are compared against mass / comoving volume
        critical_density=CRITICAL_DENSITY * a**3,
        mean_density=CRITICAL_DENSITY * OMEGA_M * a**3,
        # np.maximum(softening[phys], radius[comoving]) -> in comoving units
        softening=SOFTENING_PHYS / a,
        # vmax = sqrt(G * M / r_phys) with r_phys = a * r_comoving
        G=NEWTON_G / a,
        H=H_INTERNAL * a,  # v += r_phys * H
        kpc_per_length=1000.0 * a,
        r_20mpc=20.0 / a,
        phys_mpc_to_coord=1.0 / a,
        nu_density=0.0,
    )


# --------------------------------------------------------------- dummy halos


def dummy_halo(rng, npart, centre, own_id, other_ids=(2, 3)):
    """One halo like DummyHaloGenerator.get_random_halo (no neutrinos)."""
    radius = rng.exponential(1.0 / 60.0, npart)
    radius[0] = 0.0
    phi = 2.0 * np.pi * rng.random(npart)
    sintheta = 2.0 * rng.random(npart) - 1.0
    costheta = np.sqrt((1.0 - sintheta) * (1.0 + sintheta))
    coords = np.zeros((npart, 3))
    coords[:, 0] = radius * np.cos(phi) * sintheta
    coords[:, 1] = radius * np.sin(phi) * sintheta
    coords[:, 2] = radius * costheta
    rmax = np.sqrt((coords**2).sum(axis=1)).max()
    coords += centre
    mass = (0.1 + 0.4 * rng.random(npart)).astype(np.float32)
    vs = (1000.0 * (rng.random((npart, 3)) - 0.5)).astype(np.float32)
    types = rng.choice([0, 1, 4, 5], size=npart, p=[0.2, 0.4, 0.39, 0.01])
    groupnr_all = rng.choice(
        [own_id, other_ids[0], other_ids[1]], size=npart, p=[0.6, 0.2, 0.2]
    ).astype(np.int32)
    unbound = rng.choice(npart, npart // 10, replace=False)
    groupnr_bound = groupnr_all.copy()
    groupnr_bound[unbound] = -1
    return dict(
        coords=coords,
        mass=mass,
        vel=vs,
        types=types,
        grnr=groupnr_bound,
        fof=groupnr_all.copy(),
        rmax=rmax,
    )


def dummy_chunk(seed, n_halos, npart_choices=(1, 10, 100, 1000, 10000), boxsize=100.0,
                periodic_edge=True, n_background=0, background_mass=0.02):
    """Concatenate ``n_halos`` dummy halos into one chunk.

    ``n_background`` uniform unbound DM particles (GroupNr_bound = FOFGroupIDs
    = -1) can be added so that spherical-overdensity radii exist (the bare
    fixture halos never fall below the density threshold).

    Returns (data, halos): ``data[ptype]`` dict of arrays, ``halos`` dict of
    arrays (cofp, search_radius, read_radius, is_central, nr_bound_part, index)
    sorted by nr_bound_part descending (SOAP/core/chunk_tasks.py:118-120)."""
    parts = []
    halos = dict(cofp=[], search_radius=[], read_radius=[], is_central=[],
                 nr_bound_part=[], index=[])
    for i in range(n_halos):
        rng = np.random.default_rng(seed + i)
        npart = int(rng.choice(npart_choices))
        centre = boxsize * rng.random(3)
        if periodic_edge and i % 7 == 0:
            # straddle the periodic boundary (tests/test_shared_mesh.py:170-188)
            centre[i % 3] = boxsize * (1.0 - 1.0e-4 * rng.random()) if i % 2 else 1.0e-4 * rng.random() * boxsize
        own = 10 * i + 1
        h = dummy_halo(rng, npart, centre, own, (10 * i + 2, 10 * i + 3))
        h["coords"] = h["coords"] % boxsize
        parts.append(h)
        halos["cofp"].append(centre)
        halos["search_radius"].append(max(1.01 * h["rmax"], 0.01))
        halos["read_radius"].append(max(halos["search_radius"][-1], 5.0))
        halos["is_central"].append(int(rng.choice([1, 0], p=[0.9, 0.1])))
        halos["nr_bound_part"].append(int((h["grnr"] == own).sum()))
        halos["index"].append(own)
    if n_background > 0:
        rng = np.random.default_rng(seed + 1000003)
        parts.append(dict(
            coords=boxsize * rng.random((n_background, 3)),
            mass=np.full(n_background, background_mass, dtype=np.float32),
            vel=(1000.0 * (rng.random((n_background, 3)) - 0.5)).astype(np.float32),
            types=np.ones(n_background, dtype=np.int64),
            grnr=np.full(n_background, -1, dtype=np.int32),
            fof=np.full(n_background, -1, dtype=np.int32),
        ))
    data = {}
    for t in (0, 1, 4, 5):
        sel = [p["types"] == t for p in parts]
        n_t = sum(int(s.sum()) for s in sel)
        if n_t == 0:
            continue
        data[t] = dict(
            Coordinates=np.concatenate([p["coords"][s] for p, s in zip(parts, sel)]),
            Masses=np.concatenate([p["mass"][s] for p, s in zip(parts, sel)]),
            Velocities=np.concatenate([p["vel"][s] for p, s in zip(parts, sel)]),
            GroupNr_bound=np.concatenate([p["grnr"][s] for p, s in zip(parts, sel)]),
            FOFGroupIDs=np.concatenate([p["fof"][s] for p, s in zip(parts, sel)]),
        )
    H = {
        "cofp": np.array(halos["cofp"], dtype=np.float64),
        "search_radius": np.array(halos["search_radius"], dtype=np.float64),
        "read_radius": np.array(halos["read_radius"], dtype=np.float64),
        "is_central": np.array(halos["is_central"], dtype=np.int32),
        "nr_bound_part": np.array(halos["nr_bound_part"], dtype=np.int64),
        "index": np.array(halos["index"], dtype=np.int64),
    }
    order = np.argsort(-H["nr_bound_part"], kind="stable")
    H = {k: v[order] for k, v in H.items()}
    return data, H


# ------------------------------------------------------------------ NFW chunks


def _mu(x):
    return math.log(1.0 + x) - x / (1.0 + x)


def nfw_chunk(
    n_part,
    n_halos,
    boxsize,
    seed=20261018,
    device="cpu",
    m_part=0.0843,
    min_np=20,
    max_np=2.0e6,
    slope=-1.9,
    halo_fraction=0.85,
    outer_factor=2.5,
    frac_central=0.95,
    min_read_radius=5.0,
    type_fractions=None,
    star_scale=0.1,
    a=SCALE_FACTOR,
    sort_cells=32,
):
    """BASELINE config-2 recipe as torch tensors on ``device``.

    ``n_halos`` NFW halos with bound particle numbers from dn/dN ~ N^slope on
    [min_np, max_np], concentration 7 (N/100)^-0.1, R_200c from N*m_part and the
    comoving critical density; bound members (GroupNr_bound = halo id) inside
    R_200c, plus unbound NFW outskirts out to ``outer_factor`` R_200c
    (GroupNr_bound = -1, FOFGroupIDs = halo id); one particle forced at r=0;
    remaining particles uniform background (GroupNr_bound = FOFGroupIDs = -1);
    velocities 1000 (U - 1/2) float32.  ``type_fractions`` = dict ptype ->
    fraction inside halos (hydro variant: stars use ``star_scale`` x radii).
    Particles are ordered by a coarse ``sort_cells``^3 grid (SWIFT top-level
    cell order), random within a cell.

    Returns (data, halos) with torch tensors; data keyed by ptype.
    """
    import torch

    g = torch.Generator(device=device)
    g.manual_seed(seed)
    f64 = dict(dtype=torch.float64, device=device)

    def rand(*shape):
        return torch.rand(*shape, generator=g, **f64)

    rho200 = 200.0 * CRITICAL_DENSITY * a**3
    # halo particle numbers from the power law (inverse CDF)
    e = slope + 1.0
    u = rand(n_halos)
    nh = (min_np**e + u * (max_np**e - min_np**e)) ** (1.0 / e)
    nh = torch.clamp(nh.floor(), min=min_np).to(torch.int64)
    nh, _ = torch.sort(nh, descending=True)
    conc = 7.0 * (nh.to(torch.float64) / 100.0) ** (-0.1)
    mu_c = torch.log1p(conc) - conc / (1.0 + conc)
    mu_o = torch.log1p(outer_factor * conc) - outer_factor * conc / (1.0 + outer_factor * conc)
    n_out = ((mu_o / mu_c - 1.0) * nh.to(torch.float64)).floor().to(torch.int64)
    # fit the particle budget by dropping the most massive halos' excess
    budget = int(halo_fraction * n_part)
    tot = nh + n_out
    csum = torch.cumsum(tot.flip(0), 0).flip(0)  # particles in halos i..end
    keep_from = int((csum > budget).sum().item())
    if keep_from > 0:
        # re-draw the too-massive halos as copies of the largest that fits
        nh[:keep_from] = nh[keep_from]
        conc = 7.0 * (nh.to(torch.float64) / 100.0) ** (-0.1)
        mu_c = torch.log1p(conc) - conc / (1.0 + conc)
        mu_o = torch.log1p(outer_factor * conc) - outer_factor * conc / (1.0 + outer_factor * conc)
        n_out = ((mu_o / mu_c - 1.0) * nh.to(torch.float64)).floor().to(torch.int64)
        tot = nh + n_out
        while int(tot.sum().item()) > n_part:
            nh = torch.clamp(nh // 2, min=min_np)
            n_out = n_out // 2
            tot = nh + n_out
    n_in_halos = int(tot.sum().item())
    n_bg = n_part - n_in_halos
    r200 = (nh.to(torch.float64) * m_part / (rho200 * 4.0 / 3.0 * math.pi)) ** (1.0 / 3.0)
    centres = boxsize * rand(n_halos, 3)

    # inverse of mu(x) by table interpolation
    xmax = float(outer_factor * conc.max().item()) * 1.001
    xt = torch.logspace(-5, math.log10(xmax), 8192, **f64)
    mt = torch.log1p(xt) - xt / (1.0 + xt)

    def inv_mu(m):
        j = torch.clamp(torch.searchsorted(mt, m), 1, mt.numel() - 1)
        m0, m1, x0, x1 = mt[j - 1], mt[j], xt[j - 1], xt[j]
        return x0 + (m - m0) / (m1 - m0) * (x1 - x0)

    hid = torch.repeat_interleave(torch.arange(n_halos, device=device), tot)
    start = torch.cumsum(tot, 0) - tot
    local = torch.arange(n_in_halos, device=device) - start[hid]
    bound = local < nh[hid]
    uu = rand(n_in_halos)
    mu_lo = torch.where(bound, torch.zeros_like(uu), mu_c[hid])
    mu_hi = torch.where(bound, mu_c[hid], mu_o[hid])
    x = inv_mu(mu_lo + uu * (mu_hi - mu_lo))
    x = torch.where(bound, torch.minimum(x, conc[hid] * (1.0 - 1e-9)), x)
    r = r200[hid] * x / conc[hid]
    r = torch.where(local == 0, torch.zeros_like(r), r)

    # particle types (hydro variant)
    if type_fractions is None:
        ptype = torch.ones(n_part, dtype=torch.int32, device=device)
    else:
        keys = sorted(type_fractions)
        p = torch.tensor([type_fractions[k] for k in keys], **f64)
        p = p / p.sum()
        draw = torch.multinomial(p, n_part, replacement=True, generator=g)
        ptype = torch.tensor(keys, dtype=torch.int32, device=device)[draw]
        is_star = ptype[:n_in_halos] == 4
        r = torch.where(is_star & (local != 0), r * star_scale, r)
        # the r=0 particle of each halo is dark matter (centre of potential)
        ptype[:n_in_halos] = torch.where(local == 0, torch.ones_like(ptype[:n_in_halos]), ptype[:n_in_halos])

    phi = 2.0 * math.pi * rand(n_in_halos)
    st = 2.0 * rand(n_in_halos) - 1.0
    ct = torch.sqrt((1.0 - st) * (1.0 + st))
    pos_h = torch.stack(
        [r * torch.cos(phi) * st, r * torch.sin(phi) * st, r * ct], dim=1
    ) + centres[hid]
    pos_h = torch.remainder(pos_h, boxsize)
    pos = torch.cat([pos_h, boxsize * rand(n_bg, 3)], dim=0)
    del pos_h
    grnr = torch.cat(
        [
            torch.where(bound, hid, torch.full_like(hid, -1)).to(torch.int32),
            torch.full((n_bg,), -1, dtype=torch.int32, device=device),
        ]
    )
    fof = torch.cat(
        [hid.to(torch.int32), torch.full((n_bg,), -1, dtype=torch.int32, device=device)]
    )
    vel = (1000.0 * (torch.rand(n_part, 3, generator=g, dtype=torch.float32, device=device) - 0.5))
    mass = torch.full((n_part,), m_part, dtype=torch.float32, device=device)
    if type_fractions is not None:
        # baryon particles are lighter (Omega_b / Omega_cdm ~ 0.19)
        mass = torch.where(ptype == 1, mass, mass * 0.19)

    # SWIFT-like coarse cell order, random within a cell
    ci = torch.clamp((pos / (boxsize / sort_cells)).floor().to(torch.int64), 0, sort_cells - 1)
    key = (ci[:, 0] * sort_cells + ci[:, 1]) * sort_cells + ci[:, 2]
    key = key * (1 << 20) + torch.randint(0, 1 << 20, (n_part,), generator=g, device=device)
    order = torch.argsort(key)
    del key, ci
    pos, vel, mass, grnr, fof, ptype = (
        pos[order], vel[order], mass[order], grnr[order], fof[order], ptype[order]
    )

    # bound radius per halo -> search radius
    rb = torch.zeros(n_halos, **f64)
    rb.scatter_reduce_(0, hid[bound], r[bound], reduce="amax", include_self=True)
    search_radius = torch.clamp(1.01 * rb, min=1.0e-3)
    read_radius = torch.clamp(search_radius, min=min_read_radius)
    ncen = int(round(frac_central * n_halos))
    is_central = torch.zeros(n_halos, dtype=torch.int32, device=device)
    is_central[torch.randperm(n_halos, generator=g, device=device)[:ncen]] = 1
    # nr_bound_part counts members actually present per type set
    nr_bound = torch.zeros(n_halos, dtype=torch.int64, device=device)
    nr_bound.scatter_add_(0, hid[bound], torch.ones_like(hid[bound]))

    halos = dict(
        cofp=centres,
        search_radius=search_radius,
        read_radius=read_radius,
        is_central=is_central,
        nr_bound_part=nr_bound,
        index=torch.arange(n_halos, dtype=torch.int64, device=device),
    )
    data = {}
    for t in torch.unique(ptype).tolist():
        s = ptype == t
        data[int(t)] = dict(
            Coordinates=pos[s].contiguous(),
            Masses=mass[s].contiguous(),
            Velocities=vel[s].contiguous(),
            GroupNr_bound=grnr[s].contiguous(),
            FOFGroupIDs=fof[s].contiguous(),
        )
    return data, halos


def to_numpy(data, halos):
    """torch -> numpy view of an ``nfw_chunk`` result."""
    d = {t: {k: v.cpu().numpy() for k, v in dd.items()} for t, dd in data.items()}
    h = {k: v.cpu().numpy() for k, v in halos.items()}
    return d, h


# ------------------------------------------------------------ one volume, many chunks
# BASELINE config 5 (and the N > 1 runs of bench.py): ONE periodic volume of the config-2 recipe whose
# particles are never materialised as a whole.  A global halo catalogue is drawn once (numpy, a few bytes per
# halo); the particles of a chunk -- the members of every halo that reaches into the chunk's region, ghost shell
# included, plus the background of the region's cells -- are generated where they are needed, on the device,
# from a counter-based hash of (global particle id, stream): a particle that lies in the ghost shells of two
# chunks comes out bit-identical in both, like a particle read twice from the same snapshot.

_M64 = (1 << 64) - 1


def _i64(x):
    """python int (mod 2^64) -> the int64 with the same bits"""
    x &= _M64
    return x - (1 << 64) if x >= (1 << 63) else x


def _mix64(x):
    """splitmix64 finaliser on an int64 tensor (wrapping arithmetic, logical shifts emulated)"""
    x = (x ^ ((x >> 30) & ((1 << 34) - 1))) * _i64(0xBF58476D1CE4E5B9)
    x = (x ^ ((x >> 27) & ((1 << 37) - 1))) * _i64(0x94D049BB133111EB)
    return x ^ ((x >> 31) & ((1 << 33) - 1))


def hash_uniform(ids, stream, seed):
    """uniform float64 in [0, 1) from (id, stream, seed); ids int64 tensor"""
    import torch

    x = _mix64(ids * _i64(0x9E3779B97F4A7C15) + _i64((stream + 1) * 0xD1B54A32D192ED03 + seed * 0x2545F4914F6CDD1D))
    return ((x >> 11) & ((1 << 53) - 1)).to(torch.float64) * (1.0 / (1 << 53))


def volume_catalogue(n_part, n_halos, boxsize, seed=20261018, m_part=0.0843, min_np=20, max_np=2.0e6, slope=-1.9,
                     halo_fraction=0.85, outer_factor=2.5, frac_central=0.95, bg_cells=64, a=SCALE_FACTOR):
    """The global halo catalogue of a volume (numpy): bound / outskirt particle numbers, concentration, R_200c,
    centres, first global particle id of every halo, and the background particle count of every cell of a
    bg_cells^3 grid.  Same distributions as nfw_chunk."""
    rng = np.random.default_rng(seed)
    rho200 = 200.0 * CRITICAL_DENSITY * a**3
    e = slope + 1.0
    nh = np.floor((min_np**e + rng.random(n_halos) * (max_np**e - min_np**e)) ** (1.0 / e)).astype(np.int64)
    nh = np.sort(np.maximum(nh, min_np))[::-1].copy()

    def outskirts(nh):
        conc = 7.0 * (nh / 100.0) ** (-0.1)
        mu_c = np.log1p(conc) - conc / (1.0 + conc)
        xo = outer_factor * conc
        mu_o = np.log1p(xo) - xo / (1.0 + xo)
        return conc, mu_c, mu_o, np.floor((mu_o / mu_c - 1.0) * nh).astype(np.int64)

    conc, mu_c, mu_o, n_out = outskirts(nh)
    # fit the particle budget like nfw_chunk: the halos that do not fit become copies of the largest that does
    budget = int(halo_fraction * n_part)
    csum = np.cumsum((nh + n_out)[::-1])[::-1]  # particles in halos i..end
    keep_from = int((csum > budget).sum())
    if keep_from > 0:
        nh[:keep_from] = nh[min(keep_from, n_halos - 1)]
        conc, mu_c, mu_o, n_out = outskirts(nh)
        while int((nh + n_out).sum()) > n_part and nh.max() > min_np:
            nh = np.maximum(nh // 2, min_np)
            conc, mu_c, mu_o, n_out = outskirts(nh)
    tot = nh + n_out
    start = np.cumsum(tot) - tot
    n_in_halos = int(tot.sum())
    n_bg = n_part - n_in_halos
    ncell = bg_cells**3
    bg_count = np.full(ncell, n_bg // ncell, dtype=np.int64)
    bg_count[: n_bg % ncell] += 1
    r200 = (nh * m_part / (rho200 * 4.0 / 3.0 * math.pi)) ** (1.0 / 3.0)
    is_central = (rng.random(n_halos) < frac_central).astype(np.int32)
    return dict(nh=nh, n_out=n_out, tot=tot, start=start, conc=conc, mu_c=mu_c, mu_o=mu_o, r200=r200,
                cofp=boxsize * rng.random((n_halos, 3)), is_central=is_central, n_in_halos=n_in_halos,
                bg_count=bg_count, bg_start=n_in_halos + np.cumsum(bg_count) - bg_count, bg_cells=bg_cells,
                boxsize=float(boxsize), m_part=float(m_part), seed=int(seed), outer_factor=float(outer_factor),
                index=np.arange(n_halos, dtype=np.int64), nr_bound_part=nh.copy(),
                # nothing of a halo lies beyond outer_factor R_200c: upper bounds for the decomposition's ghost shells
                search_radius=np.maximum(1.01 * r200, 1.0e-3), read_radius=np.maximum(1.01 * r200, 5.0))


def volume_chunk(cat, halo_sel, device="cpu", cells_per_dim=64, sort_cells=32, full_cover=False):
    """Particles and halo arrays of one chunk of the volume: ``halo_sel`` = catalogue rows of the chunk's halos.
    Returns (data, halos) like nfw_chunk (torch tensors on ``device``), the particles being every particle of the
    volume that lies in the chunk's cell cover (chunk_tasks.cell_cover of the chunk's read regions)."""
    import torch

    from .chunk_tasks import cell_cover

    L, seed = cat["boxsize"], cat["seed"]
    halo_sel = np.asarray(halo_sel)
    cover3 = cell_cover(cat["cofp"][halo_sel], cat["read_radius"][halo_sel], L, cells_per_dim)
    if full_cover:  # every particle of the volume, also those no halo of the selection can reach
        cover3[:] = True
    cover = [cover3.any(axis=(1, 2)), cover3.any(axis=(0, 2)), cover3.any(axis=(0, 1))]  # per-axis projections
    cs = L / cells_per_dim
    # halos that may reach into the covered region: some covered slab within [c - R, c + R] in every dimension
    # (a superset; the per-particle cut below is exact)
    R = cat["outer_factor"] * cat["r200"]
    touch = np.ones(len(R), dtype=bool)
    for d in range(3):
        if cover[d].all():
            continue
        csum = np.concatenate([[0], np.cumsum(np.tile(cover[d], 3))])  # three periods: ranges may wrap
        lo = np.floor((cat["cofp"][:, d] - R) / cs).astype(np.int64) + cells_per_dim
        hi = np.floor((cat["cofp"][:, d] + R) / cs).astype(np.int64) + cells_per_dim
        touch &= (csum[np.minimum(hi + 1, 3 * cells_per_dim)] - csum[np.maximum(lo, 0)]) > 0
    hs = np.flatnonzero(touch)
    t = lambda x, dt=torch.float64: torch.as_tensor(np.ascontiguousarray(x), device=device).to(dt)  # noqa: E731
    tot = t(cat["tot"][hs], torch.int64)
    n = int(tot.sum().item())
    hid = torch.repeat_interleave(torch.arange(len(hs), device=device), tot)
    local = torch.arange(n, device=device) - (torch.cumsum(tot, 0) - tot)[hid]
    gid = t(cat["start"][hs], torch.int64)[hid] + local  # global particle id
    nh, conc, mu_c, mu_o, r200 = (t(cat[k][hs]) for k in ("nh", "conc", "mu_c", "mu_o", "r200"))
    bound = local < nh[hid].to(torch.int64)
    xmax = float(cat["outer_factor"] * cat["conc"].max()) * 1.001
    xt = torch.logspace(-5, math.log10(xmax), 8192, dtype=torch.float64, device=device)
    mt = torch.log1p(xt) - xt / (1.0 + xt)
    uu = hash_uniform(gid, 0, seed)
    mu_lo = torch.where(bound, torch.zeros_like(uu), mu_c[hid])
    mu_hi = torch.where(bound, mu_c[hid], mu_o[hid])
    m = mu_lo + uu * (mu_hi - mu_lo)
    j = torch.clamp(torch.searchsorted(mt, m), 1, mt.numel() - 1)
    x = xt[j - 1] + (m - mt[j - 1]) / (mt[j] - mt[j - 1]) * (xt[j] - xt[j - 1])
    x = torch.where(bound, torch.minimum(x, conc[hid] * (1.0 - 1e-9)), x)
    r = r200[hid] * x / conc[hid]
    r = torch.where(local == 0, torch.zeros_like(r), r)
    del uu, mu_lo, mu_hi, m, j, x
    phi = 2.0 * math.pi * hash_uniform(gid, 1, seed)
    st = 2.0 * hash_uniform(gid, 2, seed) - 1.0
    ct_ = torch.sqrt((1.0 - st) * (1.0 + st))
    pos = torch.stack([r * torch.cos(phi) * st, r * torch.sin(phi) * st, r * ct_], dim=1) + t(cat["cofp"][hs])[hid]
    pos = torch.remainder(pos, L)
    del phi, st, ct_
    ghid = t(cat["index"][hs], torch.int64)[hid]
    # bound radius of every halo -> search radius, before the cut (a halo of the chunk is complete by construction)
    rb = torch.zeros(len(hs), dtype=torch.float64, device=device)
    rb.scatter_reduce_(0, hid[bound], r[bound], reduce="amax", include_self=True)
    # keep what lies in the region
    cov = torch.as_tensor(cover3, device=device)
    cell = torch.clamp(torch.floor(pos / cs).to(torch.int64), 0, cells_per_dim - 1)
    keep = cov[cell[:, 0], cell[:, 1], cell[:, 2]]
    pos, gid_k = pos[keep], gid[keep]
    grnr = torch.where(bound, ghid, torch.full_like(ghid, -1))[keep].to(torch.int32)
    fof = ghid[keep].to(torch.int32)
    del cell, keep, hid, local, bound, r, ghid, gid
    # background of the covered cells
    idx = np.flatnonzero(cover3.ravel())
    assert cat["bg_cells"] == cells_per_dim
    cnt = t(cat["bg_count"][idx], torch.int64)
    nb = int(cnt.sum().item())
    cid = torch.repeat_interleave(torch.arange(len(idx), device=device), cnt)
    bgid = t(cat["bg_start"][idx], torch.int64)[cid] + (torch.arange(nb, device=device) - (torch.cumsum(cnt, 0) - cnt)[cid])
    cflat = t(idx, torch.int64)[cid]
    org = torch.stack([cflat // (cells_per_dim * cells_per_dim), (cflat // cells_per_dim) % cells_per_dim,
                       cflat % cells_per_dim], dim=1).to(torch.float64) * cs
    bpos = org + cs * torch.stack([hash_uniform(bgid, s, seed) for s in (0, 1, 2)], dim=1)
    pos = torch.cat([pos, torch.clamp(bpos, max=L * (1.0 - 1e-16))])
    gid_all = torch.cat([gid_k, bgid])
    grnr = torch.cat([grnr, torch.full((nb,), -1, dtype=torch.int32, device=device)])
    fof = torch.cat([fof, torch.full((nb,), -1, dtype=torch.int32, device=device)])
    del org, bpos, cid, cflat, bgid, gid_k
    vel = (1000.0 * (torch.stack([hash_uniform(gid_all, s, seed) for s in (3, 4, 5)], dim=1) - 0.5)).to(torch.float32)
    npart = pos.shape[0]
    # SWIFT-like coarse cell order, hashed within a cell
    ci = torch.clamp((pos / (L / sort_cells)).floor().to(torch.int64), 0, sort_cells - 1)
    key = ((ci[:, 0] * sort_cells + ci[:, 1]) * sort_cells + ci[:, 2]) * (1 << 20) + (_mix64(gid_all) & ((1 << 20) - 1))
    order = torch.argsort(key)
    del key, ci, gid_all
    data = {1: dict(Coordinates=pos[order].contiguous(), Masses=torch.full((npart,), cat["m_part"], dtype=torch.float32, device=device),
                    Velocities=vel[order].contiguous(), GroupNr_bound=grnr[order].contiguous(), FOFGroupIDs=fof[order].contiguous())}
    # halo arrays of the chunk
    pos_of = {int(h): i for i, h in enumerate(hs)}
    rows = np.array([pos_of[int(h)] for h in halo_sel], dtype=np.int64)
    rb_c = rb[torch.as_tensor(rows, device=device)]
    sr = torch.clamp(1.01 * rb_c, min=1.0e-3)
    halos = dict(cofp=t(cat["cofp"][halo_sel]), search_radius=sr, read_radius=torch.clamp(sr, min=5.0),
                 is_central=t(cat["is_central"][halo_sel], torch.int32), nr_bound_part=t(cat["nh"][halo_sel], torch.int64),
                 index=t(cat["index"][halo_sel], torch.int64))
    return data, halos
