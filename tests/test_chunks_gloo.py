"""Host logic of the multi-GPU path (soap_b200/chunk_tasks.py) on CPU: Peano
decomposition, ghost shells, chunk -> rank mapping and the result gather, run
with torch.distributed (gloo) at world_size 2.  The per-chunk compute is the
oracle's periodic sphere count (no GPU here): a chunk carrying its own ghost
shell must give exactly the whole-box answer for each of its halos."""

import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import mesh as om
from soap_b200 import chunk_tasks as ct

L = 50.0


def _box(seed=5, n=40000, nh=60):
    rng = np.random.default_rng(seed)
    pos = rng.random((n, 3)) * L
    cofp = rng.random((nh, 3)) * L
    cofp[:6] = np.array([[0.1, 0.2, 49.9], [49.8, 25.0, 0.05], [25.0, 49.95, 25.0], [0.0, 0.0, 0.0], [49.99, 49.99, 49.99],
                         [0.3, 49.7, 10.0]])  # halos straddling the periodic faces / corners
    halo = {
        "cofp": cofp,
        "index": np.arange(nh, dtype=np.int64) * 3 + 7,
        "search_radius": 0.5 + 2.0 * rng.random(nh),
        "read_radius": np.full(nh, 3.0),
    }
    data = {1: {"Coordinates": pos, "Masses": rng.random(n).astype(np.float32)}}
    return data, halo


def _compute(cd, hc):
    """count + enclosed mass inside search_radius, from the chunk's particles only"""
    pos, m = cd[1]["Coordinates"], cd[1]["Masses"]
    out = np.zeros((len(hc["index"]), 2))
    for i in range(len(hc["index"])):
        idx = om.brute_force_query(pos, hc["cofp"][i], hc["search_radius"][i], L)
        out[i] = [len(idx), m[idx].astype(np.float64).sum()]
    return torch.as_tensor(out)


def _whole_box(data, halo):
    return _compute(data, halo).numpy()


def test_hilbert_keys_are_a_bijection_and_continuous():
    bits = 3
    g = np.arange(2**bits)
    ix, iy, iz = [a.ravel() for a in np.meshgrid(g, g, g, indexing="ij")]
    key = ct.hilbert_keys(ix, iy, iz, bits)
    assert sorted(key.tolist()) == list(range(8**bits))
    order = np.argsort(key)
    steps = np.abs(np.diff(ix[order])) + np.abs(np.diff(iy[order])) + np.abs(np.diff(iz[order]))
    assert np.all(steps == 1)  # consecutive cells of a Hilbert curve are face neighbours


def test_decomposition_sizes_and_assignment():
    _, halo = _box()
    hs, cs = ct.peano_decomposition(L, halo, 8)
    assert cs.sum() == 60 and cs.max() - cs.min() <= 1 and np.all(np.diff(cs) <= 0)
    assert sorted(hs["index"].tolist()) == sorted(halo["index"].tolist())
    _, cs1 = ct.peano_decomposition(L, halo, 1000)  # domain_decomposition.py:76-78
    assert len(cs1) == 60
    a = ct.assign_chunks(8, 3)
    assert a == [[0, 3, 6], [1, 4, 7], [2, 5]]


def test_single_process_chunks_equal_whole_box():
    data, halo = _box()
    t, i = ct.run_chunks(data, halo, L, 7, _compute)
    order = np.argsort(halo["index"])
    assert np.array_equal(i.numpy(), halo["index"][order])
    assert np.array_equal(t.numpy(), _whole_box(data, halo)[order])


def test_reread_loop_repeats_only_the_halos_whose_region_was_too_small():
    """chunk_tasks.py:188-367 / halo_tasks.py:386-402: status 1 -> read radius x factor, restart from
    the search radius that was reached; rows of finished halos are kept"""
    data, halo = _box(seed=9, nh=40)
    need = 1.0 + 5.0 * np.random.default_rng(3).random(40)  # radius each halo must reach
    need_of = dict(zip(halo["index"].tolist(), need.tolist()))
    calls = []

    def compute(cd, hc):
        n = len(hc["index"])
        calls.append(n)
        out = np.zeros((n, 8))
        for i in range(n):
            want = need_of[int(hc["index"][i])]
            if want > hc["read_radius"][i]:  # ladder hit the edge of what was read
                out[i, 0] = 1
                out[i, 4] = max(hc["search_radius"][i], hc["read_radius"][i])
                out[i, 5] = max(1.5 * hc["read_radius"][i], out[i, 4])
            else:
                idx = om.brute_force_query(cd[1]["Coordinates"], hc["cofp"][i], want, L)
                out[i, 6:] = [len(idx), want]
        return torch.as_tensor(out)

    t, i = ct.run_chunks(data, halo, L, 3, compute, reread=True)
    t = t.numpy()
    assert (t[:, 0] == 0).all() and len(calls) > 3 and calls[-1] < calls[0]
    order = np.argsort(halo["index"])
    for row, h in zip(t, order):
        ref = om.brute_force_query(data[1]["Coordinates"], halo["cofp"][h], need[h], L)
        assert row[6] == len(ref) and row[7] == need[h]
    # without the loop the same run leaves the unfinished halos flagged
    t0, _ = ct.run_chunks(data, halo, L, 3, compute, reread=False)
    assert (t0.numpy()[:, 0] == 1).sum() == (need > 3.0).sum()


def test_separate_chunks_isolate_the_largest_halos():
    """domain_decomposition.py:28-60,97-140: halos above the smallest threshold leave the curve and come last,
    largest first, in chunks of n_halo_per_chunk of the first threshold they exceed"""
    data, halo = _box()
    rng = np.random.default_rng(1)
    halo["nr_bound_part"] = rng.integers(20, 1000, size=60)
    halo["nr_bound_part"][[4, 17, 33, 50, 51]] = [90000, 2000000, 15000, 700000, 12000]
    sep = [{"n_bound_threshold": 500000, "n_halo_per_chunk": 1}, {"n_bound_threshold": 10000, "n_halo_per_chunk": 2}]
    hs, cs = ct.peano_decomposition(L, halo, 5, separate_chunks=sep)
    assert cs.tolist() == [11, 11, 11, 11, 11, 1, 1, 2, 1] and cs.sum() == 60
    assert hs["nr_bound_part"][55:].tolist() == [2000000, 700000, 90000, 15000, 12000]
    assert sorted(hs["index"].tolist()) == sorted(halo["index"].tolist())
    # the isolated chunks still give the whole-box answer
    t, i = ct.run_chunks(data, halo, L, 5, _compute, separate_chunks=sep)
    order = np.argsort(halo["index"])
    assert np.array_equal(t.numpy(), _whole_box(data, halo)[order])


def test_device_ghost_cut_equals_host_ghost_cut():
    data, halo = _box()
    hs, cs = ct.peano_decomposition(L, halo, 6)
    for c in range(6):
        hc = ct.chunk_halos(hs, cs, c)
        ref = ct.ghost_mask(data[1]["Coordinates"], hc["cofp"], hc["read_radius"], L)
        got = ct.ghost_mask_device(torch.as_tensor(data[1]["Coordinates"]), hc["cofp"], hc["read_radius"], L)
        assert np.array_equal(got.numpy(), ref) and 0 < ref.sum() < len(ref)


def test_stragglers_and_fatal_rows_are_not_passed_through_silently():
    data, halo = _box(nh=8)

    def never(cd, hc):  # every halo keeps asking for a larger region
        out = np.zeros((len(hc["index"]), 8))
        out[:, 0], out[:, 4], out[:, 5] = 1, hc["search_radius"], 1.5 * hc["read_radius"]
        return torch.as_tensor(out)

    with pytest.raises(RuntimeError, match="still ask for a larger read radius"):
        ct.run_chunks(data, halo, L, 2, never, reread=True, max_passes=3)

    def fatal(cd, hc):  # status 2: Ntot > nr_bound_part is a RuntimeError in the reference
        out = np.zeros((len(hc["index"]), 8))
        out[0, 0] = 2
        return torch.as_tensor(out)

    with pytest.raises(RuntimeError, match="failed with status 2"):
        ct.run_chunks(data, halo, L, 2, fatal, reread=True)


def _worker(rank, world, port, q, nr_chunks=7):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        data, halo = _box()
        got = ct.run_chunks(data, halo, L, nr_chunks, _compute, rank=rank, world_size=world)
        if rank == 0:
            q.put((got[0].numpy(), got[1].numpy()))
        else:
            assert got is None
        dist.barrier()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,nr_chunks", [(2, 7), (3, 2)])  # (3, 2): a rank that owns no chunk takes part in the gather
def test_ranks_gloo_gather_equals_whole_box(world, nr_chunks):
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q, nr_chunks)) for r in range(world)]
    for p in procs:
        p.start()
    t, i = q.get(timeout=120)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    data, halo = _box()
    order = np.argsort(halo["index"])
    assert np.array_equal(i, halo["index"][order])
    assert np.array_equal(t, _whole_box(data, halo)[order])
