#!/usr/bin/env python
"""
bench.py -- throughput of the SOAP per-halo aggregation hot path on B200.

  python bench.py --gpus N --steps K --warmup W            (our arm)
  python bench.py --impl reference --gpus N --steps K --warmup W

A *step* is one pass of the hot path over one chunk: cell-list build + reorder
(stage A), then the batched radius ladder / sphere gather / radial sort /
scans / moment reductions (stage B+C) for every halo of the chunk.  Workload at
N=1 is BASELINE.json configs[1]: synthetic DMO chunk, 512^3 particles,
2e5 halos, L=284.4, SO {200_crit, 200_mean, 500_crit, BN98} + BoundSubhalo
(SURVEY.md 8(d).2).  For N>1 the ranks share ONE periodic volume of N times that recipe:
a global halo catalogue, Peano-Hilbert chunks with ghost shells (one per rank), particles
generated per chunk on the device and box-wrapped around the chunk's reference position
(BASELINE config 5's shape).  No data-path collective; a float32 copy of
the per-rank result tables is gathered to rank 0 asynchronously (weak scaling).

`value`   halos/s with the chunk's raw arrays already resident in HBM.
`e2e`     the same metric through the host-buffer API: pinned host arrays ->
          H2D -> kernels -> D2H of the result table, all inside the timed region.
`roofline` / `kernels`  per kernel, from an instrumented pass after the timed region:
          CUDA events around every launch with every kernel on one stream
          (soap_halo_config.debug_flags bit 1); `roofline` is the kernel with the most time.
`cpu_baseline` the numpy oracle port of the reference path on the host cores, on a fixed
          sub-cube of the same chunk; `parity` compares the GPU table of the timed step with it.
"""

import argparse
import json
import multiprocessing as mp
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (n_part, n_halos, boxsize, max_np)
    "config2": (512**3, 200000, 284.4, 2.0e6),
    "config2_small": (128**3, 3125, 71.1, 2.0e5),  # same number densities, 1/64 of the volume
    # hydro variants (BASELINE config 3: gas/DM/stars/BH, exclusive+inclusive 30/50/100 kpc apertures + SO;
    # "_kappa" adds kappa_corot, DtoT and the stellar rotation properties)
    "config3": (2 * 256**3, 50000, 142.2, 5.0e5),
    "config3_kappa": (2 * 256**3, 50000, 142.2, 5.0e5),
    "config3_iter": (2 * 256**3, 50000, 142.2, 5.0e5),  # + the iterative inertia tensors (20 passes)
    # BASELINE config 4: COLIBRE_THERMAL-like hydro chunk, projected apertures {1,3,10,30,50,100} kpc with
    # projected dispersions / half-mass radii / tensors + ExclusiveSphere 3-D kinematics, tensors, half-mass radii
    "config4": (2 * 376**3, 50000, 25.0, 2.0e5),
}
# particle mass (1e10 Msun) where it differs from the config-2 recipe: config 4 keeps config 2's mean density
M_PART = {"config4": 0.0843 * (25.0 / 284.4) ** 3 * 512**3 / (2 * 376**3)}
HYDRO_TYPES = {0: 0.45, 1: 0.50, 4: 0.049, 5: 0.001}
SEED = 20261018
SO_LIST = None  # filled in main (needs synth)
VOLUME_SLABS = 256  # cells per dimension of the ghost-shell cover (chunk_tasks.cell_cover) of the one-volume runs (N > 1)
CPU_SAMPLE_FRACTION = 0.05  # volume fraction of the fixed central sub-cube the CPU arm processes

# Algorithmic bytes per unit of every kernel.  unit = which counter the kernel's work is
# proportional to: "part" particles of the chunk, "tier0"/"tier1" final pairs of the staged tiers, "tier" both,
# "count" in-sphere particles of the count sweeps, "rec" records gathered on the general path, "mom" pairs of the
# general moment sweeps, "halo" halos.  Re-reads, out-of-sphere candidates and retried rungs are not credited.
KERNEL_MODEL = {
    "k_bounds_partial": (24, "part"),      # positions in
    "k_cell_keys": (32, "part"),           # 24 position in, 4 key + 4 index out
    "k_rs_hist": (4, "part"),              # per pass: keys in
    "k_rs_scatter": (16, "part"),          # per pass: 8 in, 8 out
    "k_cell_offsets": (4, "part"),
    "k_gather": (102, "part"),             # 4 permutation + 49 payload in, 49 out
    "k_tier_front<256>": (57, "tier0"),    # 24 position + 4 mass + 4 grnr + 4 fof + 1 type in, 16 record + 4 slot out
    "k_tier_front<1024>": (57, "tier1"),
    "k_solve_seq": (16, "seq"),            # one read of the sorted records (tier halos + the smallest general-path ones)
    "k_tier_moments": (52, "tier"),        # 4 slot + 48 payload
    "k_rows": (None, "halo"),              # ncol * 8 out, filled in at run time
    "k_count": (28, "count"),              # position 24 + mass 4
    "k_fine_hist_halo": (24, "rec"),
    "k_collect": (49, "rec"),              # 33 in, 16 record out
    "k_sort_bins": (32, "rec"),            # 16 in, 16 out
    "k_scan_solve<1>": (16, "rec_cta"),        # CTA per halo; <8> / <16>: clusters of 8 / 16 CTAs per halo, each on its own
    "k_scan_solve<8>": (16, "rec_c8"),        # size class of halos (concurrent on side streams in the timed region, one after
    "k_scan_solve<16>": (16, "rec_c16"),       # the other in the instrumented pass the table is taken from)
    "k_moments": (48, "mom"),              # position 24, mass 4, velocity 12, grnr 4, fof 4
    "k_projected": (48, "mom"),
    "k_kappa": (48, "mom"),
    "k_it_accum": (41, "mom"),
}


def kernel_key(name):
    """normalise a launch-site name ("(k_moments<V_FULL, 4>)") to its KERNEL_MODEL key"""
    n = name.strip("() ")
    if n in KERNEL_MODEL:
        return n
    if n.startswith("k_scan_solve<"):  # k_scan_solve<NCH, CS>: one entry per cluster size
        return "k_scan_solve<%s>" % n.split(",")[-1].strip(" >").replace("SCAN_CS_HUGE", "16").replace("SCAN_CS", "8")
    return n.split("<")[0]


def log(*a):
    print(*a, file=sys.stderr, flush=True)


def peak_hbm():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured"
        except Exception:
            pass
    return 6650.0, "fallback"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.idx = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                 "-i", str(self.idx)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {
            "sm_mhz": float(np.median(sm)) if sm else None,
            "sm_max_mhz": float(max(mx)) if mx else None,
            "reasons": sorted(reasons),
            "samples": len(sm),
        }


# ------------------------------------------------------------------ CPU arm

_G = {}


def _cpu_worker(args):
    """One host core pulling halos largest-first off a shared counter
    (mirrors the fetch-and-add loop of SOAP/core/halo_tasks.py:342-357)."""
    from oracle import halo as oh

    counter, n_h, keep = args
    data, H, meshes, props, params, td = (_G[k] for k in ("data", "H", "meshes", "props", "params", "td"))
    done = pairs = 0
    rows = []
    while True:
        with counter.get_lock():
            i = counter.value
            counter.value += 1
        if i >= n_h:
            break
        ih = {k: (v[i].copy() if v.ndim > 1 else v[i]) for k, v in H.items()}
        try:
            res, info = oh.process_single_halo(meshes, data, props, params, ih, td if ih["is_central"] == 1 else None)
            err = None
        except RuntimeError as e:
            res, info, err = None, {"n_loop": -1}, str(e)
        done += 1
        if res is not None:
            pairs += sum(len(v) for v in info["idx"].values())
        if keep:
            if res is not None:
                info = dict(info, idx={t: np.empty(len(v), dtype=np.int8) for t, v in info["idx"].items()})  # lengths only
            rows.append((i, res, info, ih, err))
    return done, pairs, rows


def cpu_reference_pass(sample, cp, so_list, cores, keep=False, faithful=True):
    """Stage A (mesh build) + stage B/C (halo loop) of the oracle on the host.  keep: also return the per-halo
    oracle results (for the parity summary of the bench line)."""
    from oracle import halo as oh
    from oracle import mesh as om
    from tests import _compare as cmp

    data, H = sample
    t0 = time.time()
    meshes = {t: om.MeshOracle(d["Coordinates"], om.mesh_resolution(len(d["Masses"]))) for t, d in data.items()}
    t_mesh = time.time() - t0
    params = cmp.oracle_params(cp, faithful=faithful)
    props = cmp.oracle_prop_list(params, cp, so_list, [])
    td = oh.target_density_of(props, params)
    _G.update(data=data, H=H, meshes=meshes, props=props, params=params, td=td)
    n_h = len(H["index"])
    ctx = mp.get_context("fork")
    counter = ctx.Value("l", 0)
    q = ctx.Queue()

    def run(counter, n_h, q):
        q.put(_cpu_worker((counter, n_h, keep)))

    procs = [ctx.Process(target=run, args=(counter, n_h, q)) for _ in range(cores)]
    t0 = time.time()
    for p in procs:
        p.start()
    res = [q.get() for _ in procs]
    for p in procs:
        p.join()
    t_halo = time.time() - t0
    done = sum(r[0] for r in res)
    pairs = sum(r[1] for r in res)
    rows = sorted((x for r in res for x in r[2]), key=lambda x: x[0])
    return dict(t_mesh=t_mesh, t_halo=t_halo, halos=done, pairs=pairs, rows=rows, props=props)


def make_sample(data_np, H_np, L, frac=CPU_SAMPLE_FRACTION, margin=6.0):
    """The CPU arm's bounded sample, fixed by the workload alone (no timing pilot): every halo whose centre lies
    in the central sub-cube holding ``frac`` of the volume and whose read radius fits in the ghost shell, with
    every particle within that cube + a shell of ``margin`` (how SOAP itself cuts chunks: SURVEY.md 8(e)).
    Returns the sub-chunk, the indices of its halos in the full list, and how many halos of the cube were left
    out because their read radius exceeds the shell."""
    side = L * frac ** (1.0 / 3.0)
    lo, hi = 0.5 * L - 0.5 * side, 0.5 * L + 0.5 * side
    c = H_np["cofp"]
    in_cube = np.all((c >= lo) & (c < hi), axis=1)
    hsel = in_cube & (H_np["read_radius"] <= margin)
    Hs = {k: v[hsel] for k, v in H_np.items()}
    ds = {}
    for t, d in data_np.items():
        p = np.mod(d["Coordinates"], L)  # chunks of a decomposed volume are box-wrapped around their reference position
        psel = np.all((p >= lo - margin) & (p < hi + margin), axis=1)
        ds[t] = {k: np.ascontiguousarray(v[psel]) for k, v in d.items()}
    return ds, Hs, side, margin, np.flatnonzero(hsel), int(in_cube.sum() - hsel.sum())


def sample_description(ds, Hs, side, margin, dropped, cores):
    return (f"fixed central sub-cube holding {100 * CPU_SAMPLE_FRACTION:.0f} % of the volume (side {side:.1f} Mpc) + "
            f"{margin} Mpc ghost shell: {len(Hs['index'])} halos ({dropped} more left out: read radius > shell), "
            f"{sum(len(d['Masses']) for d in ds.values())} particles; mesh build + halo loop, numpy oracle port, "
            f"{cores} processes pulling halos largest-first")


def parity_summary(res, r, hidx, cp, faithful):
    """GPU table of the timed step vs oracle rows of the same halos of the same chunk."""
    from tests import _compare as cmp

    rows = r["rows"]
    oracle_out = [(x[1], x[2], x[3], x[4]) for x in rows]
    halos = [int(hidx[x[0]]) for x in rows]
    rep = cmp.compare(res, oracle_out, r["props"], cp, halos=halos, flags=8, faithful=faithful)
    m = rep.maxerr
    for b in rep.bad[:10]:
        log("[bench] parity: out of tolerance:", b)
    int_keys = ("Ngas", "Ndm", "Nstar", "Nbh", "n_pairs")
    mr_keys = ("r", "Mso", "Mtot", "Mdm", "HalfMassRadiusTot", "HalfMassRadiusDM", "EncloseRadius", "R_vmax_soft", "R_vmax_unsoft",
               "Mfrac_satellites", "Mfrac_external", "Vmax_soft", "Vmax_unsoft")
    return {
        "halos": len(halos), "counts_exact": all(m.get(k, 0.0) == 0.0 for k in int_keys),
        "n_loop_exact": m.get("n_loop", 0.0) == 0.0, "radius_exact": m.get("radius", 0.0) <= 1e-15,
        "max_rel_mass_radius": max(m.get(k, 0.0) for k in mr_keys),
        "max_rel_second_moment": max(m.get(k, 0.0) for k in ("spin_parameter", "concentration_soft", "concentration_unsoft",
                                                             "concentration_dmo_soft", "concentration_dmo_unsoft")),
        "out_of_tolerance": len(rep.bad),
        "worst_keys": {k: float(f"{v:.3g}") for k, v in sorted(m.items(), key=lambda kv: -kv[1])[:5]},
    }


def parity_block(res, sample, hidx, cp, so_list, cores, r_faithful):
    """The bench line's parity field.  ``float64``: the oracle with every sum in float64 on >= 2000 halos of the
    sample including its 50 largest -- north_star's tolerances (1e-6 masses / radii, 1e-4 second moments).
    ``faithful``: the rows of the timed CPU run (reference dtypes: float32 sums and a float32 cumsum in get_vmax,
    kinematic_properties.py:583, whose argmax can land on another particle), reported for reference."""
    ds, Hs = sample
    n = len(Hs["index"])
    order = np.argsort(-Hs["nr_bound_part"], kind="stable")
    pick = np.unique(np.concatenate([order[:50], np.arange(0, n, max(1, n // 2000))]))
    Hp = {k: v[pick] for k, v in Hs.items()}
    r64 = cpu_reference_pass((ds, Hp), cp, so_list, cores, keep=True, faithful=False)
    out = {"against": "numpy oracle on the CPU sample of the chunk of the timed step",
           "float64": parity_summary(res, r64, hidx[pick], cp, faithful=False)}
    if r_faithful is not None:
        out["faithful"] = parity_summary(res, r_faithful, hidx, cp, faithful=True)
    return out


# ------------------------------------------------------------------ GPU arm


def build_config(cp, so_list, workload="config2"):
    from soap_b200.halo_tasks import HaloPropConfig, PF_HMR, PF_ITER, PF_KAPPA, PF_KIN, PF_TENS

    common = dict(
        boxsize=cp["boxsize"], G=cp["G"], critical_density=cp["critical_density"],
        mean_density=cp["mean_density"], softening={t: cp["softening"] for t in (0, 1, 4, 5)},
        H=cp["H"], kpc_per_length=cp["kpc_per_length"], r_20mpc=cp["r_20mpc"],
        nu_density=cp["nu_density"], phys_mpc_to_coord=cp["phys_mpc_to_coord"], do_subhalo=True)
    if workload.startswith("config3"):
        aps = [(kpc * 1e-3 * cp["phys_mpc_to_coord"], kpc * 1e-3, incl) for kpc in (30.0, 50.0, 100.0) for incl in (0, 1)]
        flags = PF_KIN | PF_TENS | PF_HMR | (PF_KAPPA if workload.endswith("kappa") else 0) | \
            (PF_ITER if workload.endswith("iter") else 0)
        return HaloPropConfig(so=list(so_list), apertures=aps, property_flags=flags, dmo=False, **common)
    if workload == "config4":
        # parameter_files/COLIBRE_THERMAL.yml: ProjectedApertureProperties variations 1..100 kpc, ExclusiveSphere
        # kinematics / tensors / half-mass radii
        radii = (1.0, 3.0, 10.0, 30.0, 50.0, 100.0)
        proj = [(kpc * 1e-3 * cp["phys_mpc_to_coord"], kpc * 1e-3) for kpc in radii]
        aps = [(kpc * 1e-3 * cp["phys_mpc_to_coord"], kpc * 1e-3, 0) for kpc in radii]
        return HaloPropConfig(so=[], apertures=aps, projected=proj, property_flags=PF_KIN | PF_TENS | PF_HMR, dmo=False,
                              skip_gt=("exclusive",), **common)
    return HaloPropConfig(so=list(so_list), apertures=[], property_flags=PF_HMR, dmo=True, **common)


def bind_to_gpu_numa_node(local_rank):
    """Best effort: run this rank (and first-touch its pinned host buffers) on the CPU cores of the
    NUMA node its GPU hangs off, so that N concurrent host->device uploads do not cross sockets."""
    try:
        out = subprocess.run(["nvidia-smi", "--query-gpu=pci.bus_id", "--format=csv,noheader", "-i", str(local_rank)],
                             capture_output=True, text=True, timeout=20).stdout.strip().lower()
        bdf = out[-12:] if len(out) >= 12 else out  # 0000:xx:yy.z
        node = int(open(f"/sys/bus/pci/devices/{bdf}/numa_node").read().strip())
        if node < 0:
            return None
        cpus = set()
        for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
            a, _, b = part.partition("-")
            cpus.update(range(int(a), int(b or a) + 1))
        if cpus:
            os.sched_setaffinity(0, cpus)
        return node
    except Exception:
        return None


_REAL_STDOUT = None


def protect_stdout():
    """Libraries underneath (NCCL prints its version banner to stdout) must not add lines to
    the one-JSON-line contract: send file descriptor 1 to stderr and keep the real one aside."""
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.dup(1)
        os.dup2(2, 1)


def emit(obj):
    line = (json.dumps(obj) + "\n").encode()
    if _REAL_STDOUT is None:
        sys.stdout.write(line.decode())
        sys.stdout.flush()
    else:
        os.write(_REAL_STDOUT, line)


def main():
    protect_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="config2", choices=sorted(WORKLOADS))
    ap.add_argument("--fine-ppc", type=int, default=0)
    ap.add_argument("--debug-flags", type=int, default=0, help="soap_halo_config.debug_flags (cross-check switches)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--cpu-budget", type=float, default=15.0)
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference" and rank != 0:
        return 0  # rank 0 alone runs the reference arm

    import torch

    from soap_b200 import synth

    so_list = [("crit", 200.0), ("mean", 200.0), ("crit", 500.0), ("BN98", float(synth.virBN98()))]
    n_part, n_halos, L, max_np = WORKLOADS[args.workload]
    cp = synth.coordinate_unit_params(L)
    if args.workload == "config4":
        wl_name = (f"{args.workload}: synthetic COLIBRE_THERMAL-like hydro chunk (gas/DM/stars/BH), {n_part} particles, "
                   f"{n_halos} halos, L={L} Mpc, projected apertures 1/3/10/30/50/100 kpc (3 axes: masses, dispersions, "
                   "half-mass radii, tensors) + ExclusiveSphere kinematics, tensors, half-mass radii + BoundSubhalo")
    elif args.workload.startswith("config3"):
        wl_name = (f"{args.workload}: synthetic hydro chunk (gas/DM/stars/BH), {n_part} particles, {n_halos} halos, "
                   f"L={L} Mpc, exclusive+inclusive 30/50/100 kpc apertures + SO x4 + BoundSubhalo, kinematics, "
                   "tensors, half-mass radii")
    else:
        wl_name = (f"{args.workload}: synthetic DMO chunk, {n_part} particles, {n_halos} halos, L={L} Mpc, "
                   "SO 200_crit/200_mean/500_crit/BN98 + BoundSubhalo (MINIMAL_FLAMINGO)")
    have_cuda = torch.cuda.is_available()
    gen_dev = f"cuda:{local_rank}" if have_cuda else "cpu"
    if have_cuda:
        torch.cuda.set_device(local_rank)
    cores = os.cpu_count() or 1
    numa = bind_to_gpu_numa_node(local_rank) if (have_cuda and world > 1 and args.impl == "ours") else None

    t0 = time.time()
    hydro = args.workload.startswith("config3") or args.workload == "config4"
    # N > 1 (DMO recipe): ONE periodic volume of N times the single-GPU workload (weak scaling), cut into N Peano-Hilbert
    # chunks that carry their own ghost shells (BASELINE config 5); rank r generates the particles of
    # chunk r on its device from the shared halo catalogue -- the volume is never materialised as a whole.
    volume = world > 1 and not hydro and args.impl == "ours"
    n_halos_total = n_halos * world
    if volume:
        from soap_b200 import chunk_tasks as ct

        Lv = L * world ** (1.0 / 3.0)
        cat = synth.volume_catalogue(n_part * world, n_halos * world, Lv, seed=SEED, max_np=max_np, bg_cells=VOLUME_SLABS)
        H_all = {k: cat[k] for k in ("cofp", "index", "search_radius", "read_radius", "nr_bound_part", "is_central")}
        hs, csz = ct.peano_decomposition(Lv, H_all, world)
        mine = ct.chunk_halos(hs, csz, ct.assign_chunks(len(csz), world)[rank][0])
        data, halos = synth.volume_chunk(cat, mine["index"], device=gen_dev, cells_per_dim=VOLUME_SLABS)
        n_own = int(sum(len(d["Masses"]) for d in data.values()))
        if have_cuda:
            # stage A1 of the path, as ChunkTask does after reading a chunk (chunk_tasks.py:210-215,284-288): particles
            # are wrapped to the periodic copies nearest the chunk's reference position, so that the ghost shell that
            # reaches across the box edge does not stretch the mesh over the whole box
            from soap_b200.shared_mesh import box_wrap

            c_np = mine["cofp"]
            ref_pos = 0.5 * (c_np.min(axis=0) + c_np.max(axis=0))
            for d in data.values():
                box_wrap(d["Coordinates"], ref_pos, Lv)
        wl_name = (f"{args.workload} x {world}: ONE synthetic DMO volume, L={Lv:.1f} Mpc, {n_part * world} particles, "
                   f"{n_halos_total} halos, cut into {world} Peano-Hilbert chunks with ghost shells (cell cover of the read "
                   f"spheres), one per GPU, particles generated per chunk on the device; SO 200_crit/200_mean/500_crit/BN98 "
                   f"+ BoundSubhalo")
        L = Lv
        cp = synth.coordinate_unit_params(L)
        n_part = n_own  # this rank's particles, ghost shell included
        del cat, H_all, hs
    else:
        data, halos = synth.nfw_chunk(n_part, n_halos, L, seed=SEED + rank + (1 if hydro else 0) + (1 if args.workload == "config4" else 0),
                                      device=gen_dev, max_np=max_np, type_fractions=HYDRO_TYPES if hydro else None,
                                      **({"m_part": M_PART[args.workload]} if args.workload in M_PART else {}))
    if have_cuda:
        torch.cuda.synchronize()
    log(f"[bench] rank {rank}: generated {args.workload} ({n_part} particles, {int(halos['cofp'].shape[0])} halos) on {gen_dev} "
        f"in {time.time() - t0:.1f}s")

    # ------------------------------------------------------------ reference arm
    if args.impl == "reference":
        data_np, H_np = synth.to_numpy(data, halos)
        del data, halos
        ds, Hs, side, margin, hidx, dropped = make_sample(data_np, H_np, L)
        n_s = len(Hs["index"])
        times, pairs = [], 0
        for it in range(args.warmup + args.steps):
            r = cpu_reference_pass((ds, Hs), cp, so_list, cores)
            if it >= args.warmup:
                times.append(r["t_mesh"] + r["t_halo"])
                pairs = r["pairs"]
        t = float(np.mean(times)) if times else float("nan")
        val = n_s / t
        sample_desc = sample_description(ds, Hs, side, margin, dropped, cores)
        out = {
            "impl": "reference", "metric": "halos_per_s", "value": val, "unit": "halos/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * t, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": wl_name}, "pairs_per_s": pairs / t,
            "cpu_baseline": {"value": val, "unit": "halos/s", "cores": cores, "kind": "port", "sample": sample_desc},
            "e2e": {"value": val, "unit": "halos/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0,
            "note": "reference = numpy oracle port (the reference needs unyt/mpi4py/h5py/virgo, absent here)",
        }
        emit(out)
        return 0

    # ------------------------------------------------------------------ our arm
    if not have_cuda:
        raise SystemExit("bench.py: no CUDA device; the soap_b200 path has no CPU fallback")
    import torch.distributed as dist

    from soap_b200 import _lib
    from soap_b200.halo_tasks import DeviceChunk, process_halos, result_layout

    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local_rank}"))
    dev = torch.device(f"cuda:{local_rank}")
    handle = _lib.default_handle(local_rank)
    cfg = build_config(cp, so_list, args.workload)
    cfg.debug_flags = args.debug_flags
    ncol, cols = result_layout(cfg.to_c())
    H = int(halos["cofp"].shape[0])
    table = torch.empty((H, ncol), dtype=torch.float64, device=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # Results stay per rank ("each rank appends to its own ResultSet"); what NCCL moves is a float32
    # copy of the table (PropertyTable's output precision) gathered to rank 0 asynchronously, overlapped with the
    # next chunk's kernels and waited for before the timed region ends.
    Hmax = H
    if world > 1:
        hm = torch.tensor([H], dtype=torch.int64, device=dev)
        dist.all_reduce(hm, op=dist.ReduceOp.MAX)
        Hmax = int(hm.item())
    gather32 = [torch.empty((Hmax, ncol), dtype=torch.float32, device=dev) for _ in range(world)] if (world > 1 and rank == 0) else None
    send32 = torch.zeros((Hmax, ncol), dtype=torch.float32, device=dev) if world > 1 else None
    pending = []

    def step_device():
        chunk = DeviceChunk(data, L, device=local_rank, fine_ppc=args.fine_ppc, handle=handle)
        res = process_halos(chunk, cfg, halos, out=table)
        if world > 1:
            for w in pending:
                w.wait()
            pending.clear()
            send32[:H].copy_(table)
            pending.append(dist.gather(send32, gather32, dst=0, async_op=True))
        return chunk, res

    # warm-up
    for _ in range(max(args.warmup, 0)):
        chunk, res = step_device()
        chunk.free()
    for w in pending:
        w.wait()
    pending.clear()
    barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    stats = {}
    l0 = handle.launches()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    ev0.record()
    for _ in range(args.steps):
        chunk, res = step_device()
        for k, v in chunk.timings().items():
            if k.startswith("stat/"):
                stats[k[5:]] = v
        chunk.free()
    for w in pending:
        w.wait()
    pending.clear()
    ev1.record()
    barrier()
    clocks = sampler.stop()
    launches = handle.launches() - l0
    # Per-kernel and per-phase device times come from a second, instrumented pass over the same chunk: CUDA events
    # around every launch (on the stream it is launched on) with the library told to launch every kernel on one stream
    # (soap_halo_config.debug_flags bit 1).  In the timed region the two tiers, the four scan variants and the general
    # path's first round run on seven streams at once, where an event pair measures how long a kernel was resident
    # next to the others, not how long its work takes.
    import dataclasses

    cfg_serial = dataclasses.replace(cfg, debug_flags=int(cfg.debug_flags) | 2)
    prof_steps = max(1, min(3, args.steps))
    phases = {}
    handle.kernel_timings()  # drop what came before
    handle.kernel_timing(True)
    evs0, evs1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    evs0.record()
    for _ in range(prof_steps):
        chunk = DeviceChunk(data, L, device=local_rank, fine_ppc=args.fine_ppc, handle=handle)
        process_halos(chunk, cfg_serial, halos, out=table)
        for k, v in chunk.timings().items():
            if not k.startswith("stat/"):
                phases[k] = phases.get(k, 0.0) + v
        chunk.free()
    evs1.record()
    torch.cuda.synchronize()
    serial_ms = evs0.elapsed_time(evs1) / prof_steps
    handle.kernel_timing(False)
    ktimes = handle.kernel_timings()
    barrier()
    ms = ev0.elapsed_time(ev1) / max(args.steps, 1)
    tmax = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
    ms_step = float(tmax.item())
    status = res.status.cpu().numpy()
    n_ok = int((status == 0).sum())
    pairs = stats.get("pairs", 0.0)
    pairs_all, n_ok_all, n_part_all = pairs, n_ok, n_part
    if world > 1:
        tot = torch.tensor([pairs, n_ok, n_part], dtype=torch.float64, device=dev)
        dist.all_reduce(tot)
        pairs_all, n_ok_all, n_part_all = float(tot[0].item()), int(tot[1].item()), int(tot[2].item())

    # ----------------------------------------------------------------- e2e
    e2e = None
    if not args.no_e2e:
        host = {t: {k: v.cpu().pin_memory() for k, v in d.items()} for t, d in data.items()}
        host_h = {k: v.cpu().pin_memory() for k, v in halos.items()}
        h2d = sum(v.numel() * v.element_size() for d in host.values() for v in d.values())
        h2d += sum(v.numel() * v.element_size() for v in host_h.values())
        out_host = torch.empty((H, ncol), dtype=torch.float64).pin_memory()
        d2h = out_host.numel() * 8 + H * 4
        del data
        torch.cuda.empty_cache()

        from soap_b200.halo_tasks import ChunkFeed

        feed = ChunkFeed(local_rank)

        def run_e2e(n):
            """n chunks through the public API; chunk i+1's host->device copy is in flight
            while chunk i is processed (every chunk is copied from pinned host memory and
            its result table is read back inside the timed region)."""
            if n <= 0:
                return None
            nxt = feed.upload(host, host_h)
            st = None
            for i in range(n):
                cur = nxt
                if i + 1 < n:
                    nxt = feed.upload(host, host_h)
                d_dev, h_dev = feed.wait(cur)
                ch = DeviceChunk(d_dev, L, device=local_rank, fine_ppc=args.fine_ppc, handle=handle)
                del d_dev
                r = process_halos(ch, cfg, h_dev, out=table)
                out_host.copy_(r.table, non_blocking=True)
                st = r.status.cpu()
                torch.cuda.synchronize()
                ch.free()
                del cur
            return st

        run_e2e(min(args.warmup, 2))
        barrier()
        # the host -> device copy of one chunk alone, all ranks at once: the ceiling the pipeline can reach when the
        # copy is the longer of the two overlapped legs (N concurrent uploads share the host's memory system)
        t0 = time.perf_counter()
        tk = feed.upload(host, host_h)
        tk[2].synchronize()
        barrier()
        copy_alone = time.perf_counter() - t0
        del tk
        tc = torch.tensor([copy_alone], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(tc, op=dist.ReduceOp.MAX)
        copy_alone = float(tc.item())
        t0 = time.perf_counter()
        n_e2e = max(1, args.steps)
        run_e2e(n_e2e)
        barrier()
        dt = (time.perf_counter() - t0) / n_e2e
        tt = torch.tensor([dt], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        dt = float(tt.item())
        e2e = {"value": n_halos_total / dt, "unit": "halos/s", "h2d_bytes_per_step": int(h2d),
               "d2h_bytes_per_step": int(d2h), "ms_per_step": 1e3 * dt,
               "pipeline": "ChunkFeed: upload of chunk i+1 overlaps processing of chunk i", "numa_node": numa,
               "h2d_copy_alone_ms": round(1e3 * copy_alone, 2),
               "h2d_copy_alone_gbs_per_gpu": round(h2d / copy_alone / 1e9, 2),
               "note": "host arrays are pinned once, outside the timed region, like SharedArray buffers registered with "
                       "cudaHostRegister at allocation would be; the step is bound by the PCIe copy whenever "
                       "h2d_copy_alone_ms exceeds the device-resident ms_per_step"}

    # -------------------------------------------------------- roofline (rank 0)
    # per kernel: launches and device time from the CUDA events the library records around every launch of the
    # timed region; achieved = algorithmic bytes of those launches / their time (KERNEL_MODEL)
    peak, peak_kind = peak_hbm()
    steps = prof_steps  # the instrumented pass
    ph = {k: v / steps for k, v in phases.items()}  # ms per step
    unit_count = {
        "part": float(n_part), "tier0": stats.get("small_pairs_0", 0.0), "tier1": stats.get("small_pairs_1", 0.0),
        "tier": stats.get("small_pairs", 0.0), "count": stats.get("count_pairs", 0.0), "rec": stats.get("try_pairs", 0.0),
        "mom": stats.get("moment_pairs", 0.0), "halo": float(H),
        "seq": stats.get("small_pairs", 0.0) + stats.get("rec_seq", 0.0), "rec_cta": stats.get("rec_cta", 0.0),
        "rec_c8": stats.get("rec_cluster8", 0.0), "rec_c16": stats.get("rec_cluster16", 0.0),
    }
    kernels = {}
    for name, (n_l, t_ms) in ktimes.items():
        key = kernel_key(name)
        e = kernels.setdefault(key, {"launches_per_step": 0.0, "ms": 0.0})
        e["launches_per_step"] += n_l / steps
        e["ms"] += t_ms / steps
    for key, e in kernels.items():
        bpu, unit = KERNEL_MODEL.get(key, (None, None))
        if key == "k_rows":
            bpu = 8.0 * ncol
        e["ms"] = round(e["ms"], 4)
        e["launches_per_step"] = round(e["launches_per_step"], 2)
        if bpu is None:
            continue
        n_units = unit_count[unit]
        # kernels launched once per pass over the data (radix passes) process the units once per launch
        per_launch = key in ("k_rs_hist", "k_rs_scatter", "k_bounds_partial")
        total_bytes = bpu * n_units * (e["launches_per_step"] if per_launch else 1.0)
        gbs = total_bytes / (e["ms"] * 1e-3) / 1e9 if e["ms"] > 0 else 0.0
        e.update(alg_bytes_per_unit=bpu, unit=unit, units_per_step=int(n_units), achieved_gbs=round(gbs, 2),
                 frac=round(gbs / peak, 4))
    modelled = {k: v for k, v in kernels.items() if "frac" in v}
    dom = max(modelled, key=lambda k: modelled[k]["ms"])
    traffic = None
    try:  # dram__bytes_read.sum + dram__bytes_write.sum per launch, from the committed ncu --set full capture
        tr = json.load(open(os.path.join(ROOT, "profiles", "r02_ncu_dram_traffic.json")))
        captured = tr.get(args.workload, {}).get(dom)
        n_cap = tr.get(args.workload + "_launches", {}).get(dom)
        if captured is not None and n_cap:  # bytes per captured launch
            traffic = round(captured / n_cap)
    except Exception:
        pass
    roofline = {"bound": "hbm", "kernel": dom, "achieved": modelled[dom]["achieved_gbs"], "peak": peak,
                "peak_kind": peak_kind, "unit": "GB/s", "frac": modelled[dom]["frac"], "traffic": traffic,
                "alg_bytes_per_launch": round(modelled[dom]["alg_bytes_per_unit"] * modelled[dom]["units_per_step"] /
                                              max(modelled[dom]["launches_per_step"], 1.0)),
                "launches_per_step": modelled[dom]["launches_per_step"], "ms_per_step": modelled[dom]["ms"],
                "share_of_step": round(modelled[dom]["ms"] / serial_ms, 3),
                "timed_in": "instrumented serial pass (see kernel_timing)"}
    total_alg = 32.0 * n_part + 48.0 * pairs + 8.0 * H * ncol
    kern_ms = sum(v["ms"] for v in kernels.values())

    # -------------------------------------------------------- cpu baseline + parity (rank 0)
    cpu_baseline = None
    parity = None
    if rank == 0 and not args.no_cpu_baseline and not hydro:
        try:
            src = host if not args.no_e2e else {t: {k: v.cpu() for k, v in d.items()} for t, d in data.items()}
            data_np = {t: {k: v.numpy() for k, v in d.items()} for t, d in src.items()}
            H_np = {k: v.cpu().numpy() for k, v in halos.items()}
            ds, Hs, side, margin, hidx, dropped = make_sample(data_np, H_np, L)
            r = cpu_reference_pass((ds, Hs), cp, so_list, cores, keep=True)
            t = r["t_mesh"] + r["t_halo"]
            cpu_baseline = {
                "value": r["halos"] / t, "unit": "halos/s", "cores": cores, "kind": "port",
                "pairs_per_s": r["pairs"] / t, "seconds": round(t, 2),
                "sample": sample_description(ds, Hs, side, margin, dropped, cores),
            }
            parity = parity_block(res, (ds, Hs), hidx, cp, so_list, cores, r)
        except Exception as e:  # the baseline is reported, never required
            cpu_baseline = {"value": None, "unit": "halos/s", "cores": cores, "kind": "port", "sample": f"failed: {e!r}"}

    if rank == 0:
        out = {
            "metric": "halos_per_s", "value": n_halos_total / (ms_step * 1e-3), "unit": "halos/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic",
            "config": {"workload": wl_name, "l2": "inputs (48 B x particles) larger than L2; no flush needed",
                       "halos_ok": n_ok_all, "halos": n_halos_total, "particles_incl_ghosts": int(n_part_all), "internal_mesh_res": int(stats.get("res", 0)),
                       "ladder_rounds": int(stats.get("rounds", 0)), "ncol": ncol,
                       "parallelism": f"{world} chunks, one per GPU, no data-path collective; float32 result tables gathered to "
                                      "rank 0 over NCCL asynchronously (overlapped with the next chunk)"},
            "pairs_per_s": pairs_all / (ms_step * 1e-3), "pairs_per_step": int(pairs_all),
            "candidates_per_step": int(stats.get("candidates", 0)),
            "algorithmic_gbs_whole_step": round(total_alg / (ms_step * 1e-3) / 1e9, 2),
            "algorithmic_frac_whole_step": round(total_alg / (ms_step * 1e-3) / 1e9 / peak, 4),
            "kernel_ms_per_step": round(kern_ms, 3), "serial_ms_per_step": round(serial_ms, 3),
            "kernel_timing": f"instrumented pass of {prof_steps} steps after the timed region, every kernel on one stream "
                             "(debug_flags bit 1) with CUDA events around each launch; the timed region overlaps tiers, scan "
                             "variants and the general path on seven streams",
            "roofline": roofline, "kernels": kernels, "phases_ms": {k: round(v, 4) for k, v in sorted(ph.items())},
            "stats": {k: int(v) for k, v in sorted(stats.items())}, "cpu_baseline": cpu_baseline, "parity": parity,
            "e2e": e2e,
            "gpu_launches": int(launches), "clocks": clocks,
        }
        emit(out)
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
