"""
Chunk-level host logic of the multi-GPU path: one process per
GPU, independent spatial chunks, no data-path collective, one gather of result
tables.

Mirrors, for one box of GPUs,
  * ``peano_decomposition``     SOAP/core/domain_decomposition.py:9-142
    (halos sorted along a Peano-Hilbert curve, cut into chunks of equal halo
    count, remainder spread over the first chunks :136-139);
  * the per-chunk particle read with its ghost shell
                                SOAP/core/mask_cells.py:6-38,
                                SOAP/core/chunk_tasks.py:194-288
    (every particle within ``read_radius`` of a chunk's halos travels with the
    chunk: ghost particles are duplicated, never exchanged);
  * the chunk -> rank assignment of SOAP/core/chunk_tasks.py's task queue,
    static here: chunk c -> rank c mod N;
  * the merge of per-chunk results (``combine_chunks``), here a gather of the
    float64 result tables to rank 0 over torch.distributed (NCCL on GPUs, gloo in
    the CPU tests).

The Peano-Hilbert key itself comes from VirgoDC (``virgo.util.peano``), which is
not available here; ``hilbert_keys`` is Skilling's transform.  The curve only
decides which halos share a chunk -- no result depends on it.
"""

import numpy as np


def hilbert_keys(ix, iy, iz, bits):
    """3-D Hilbert curve index of integer cell coordinates (Skilling 2004)."""
    X = [np.asarray(a, dtype=np.int64).copy() for a in (ix, iy, iz)]
    M = 1 << (bits - 1)
    Q = M
    while Q > 1:  # inverse undo excess work
        P = Q - 1
        for i in range(3):
            hit = (X[i] & Q) != 0
            X[0] = np.where(hit, X[0] ^ P, X[0])
            if i > 0:
                t = np.where(hit, 0, (X[0] ^ X[i]) & P)
                X[0] ^= t
                X[i] ^= t
        Q >>= 1
    for i in range(1, 3):  # Gray encode
        X[i] ^= X[i - 1]
    t = np.zeros_like(X[0])
    Q = M
    while Q > 1:
        t = np.where((X[2] & Q) != 0, t ^ (Q - 1), t)
        Q >>= 1
    for i in range(3):
        X[i] ^= t
    key = np.zeros_like(X[0])
    for b in range(bits - 1, -1, -1):  # interleave, x most significant
        for i in range(3):
            key = (key << 1) | ((X[i] >> b) & 1)
    return key


def peano_decomposition(boxsize, halo, nr_chunks, bits_per_dimension=10, separate_chunks=None):
    """domain_decomposition.py:9-142 on one rank.  ``halo`` is the dict of halo
    arrays (cofp [H,3], ...).  Returns (halo in chunk order, chunk_size).

    ``separate_chunks`` (list of {"n_bound_threshold", "n_halo_per_chunk"}, sorted by descending threshold
    like the parameter file's, :28-60,97-140) takes the halos with more bound particles than the smallest
    threshold out of the curve and appends them, largest first, as extra chunks of at most
    ``n_halo_per_chunk`` halos of the first threshold they exceed -- a 10^6-particle cluster with its 5 Mpc
    ghost shell then does not share a GPU's memory with a thousand neighbours."""
    if separate_chunks:
        nb = np.asarray(halo["nr_bound_part"])
        large = nb > separate_chunks[-1]["n_bound_threshold"]  # :38-39
        if large.any():
            small = {k: np.asarray(v)[~large] for k, v in halo.items()}
            big = {k: np.asarray(v)[large] for k, v in halo.items()}
            order = np.argsort(big["nr_bound_part"], kind="stable")[::-1]  # :100-102
            big = {k: v[order] for k, v in big.items()}
            sizes, i_thr, i_h, n_big = [], 0, 0, len(order)
            while i_h < n_big:  # :105-117
                while separate_chunks[i_thr]["n_bound_threshold"] > big["nr_bound_part"][i_h]:
                    i_thr += 1
                c = min(int(separate_chunks[i_thr]["n_halo_per_chunk"]), n_big - i_h)
                sizes.append(c)
                i_h += c
            if len(small["index"]):
                out, chunk_size = peano_decomposition(boxsize, small, nr_chunks, bits_per_dimension)
                out = {k: np.concatenate([out[k], big[k]], axis=0) for k in out}  # :120-128
                return out, np.concatenate([chunk_size, np.array(sizes, dtype=np.int64)])
            return big, np.array(sizes, dtype=np.int64)
    centres = np.asarray(halo["cofp"], dtype=np.float64)
    nr_halos = centres.shape[0]
    nr_chunks = max(1, min(int(nr_chunks), nr_halos))  # :76-78
    cells = 2**bits_per_dimension
    grid_size = boxsize / cells
    ipos = np.clip(np.floor(centres / grid_size).astype(np.int64), 0, cells - 1)  # :84-85
    key = hilbert_keys(ipos[:, 0], ipos[:, 1], ipos[:, 2], bits_per_dimension)
    order = np.argsort(key, kind="stable")
    out = {k: np.asarray(v)[order] for k, v in halo.items()}
    chunk_size = np.zeros(nr_chunks, dtype=np.int64)
    chunk_size[:] = nr_halos // nr_chunks  # :136-139
    chunk_size[: nr_halos % nr_chunks] += 1
    return out, chunk_size


def assign_chunks(nr_chunks, world_size):
    """chunk c -> rank c mod N."""
    return [list(range(r, nr_chunks, world_size)) for r in range(world_size)]


def chunk_halos(halo, chunk_size, c):
    """The contiguous run of halos of chunk c."""
    off = np.concatenate([[0], np.cumsum(chunk_size)])
    return {k: v[off[c] : off[c + 1]] for k, v in halo.items()}


def cell_cover(cofp, read_radius, boxsize, cells_per_dim=64):
    """Which cells of a ``cells_per_dim``^3 grid of the box are touched by some halo's cube
    [c - r, c + r]^3 (periodic): bool [n, n, n].  This is the reference's read mask (mask_cells.py:6-38 marks,
    per halo, the SWIFT cells that overlap its read region), rasterised for all halos at once with a 3-D
    difference array."""
    cofp = np.asarray(cofp, dtype=np.float64).reshape(-1, 3)
    rr = np.asarray(read_radius, dtype=np.float64).reshape(-1, 1)
    n = int(cells_per_dim)
    cs = boxsize / n
    if len(cofp) == 0:
        return np.zeros((n, n, n), dtype=bool)
    lo = np.floor((cofp - rr) / cs).astype(np.int64)
    ext = np.floor((cofp + rr) / cs).astype(np.int64) - lo + 1
    full = ext >= n  # the cube spans the whole axis
    lo = np.where(full, 0, lo % n)  # in [0, n)
    hi1 = lo + np.where(full, n, ext)  # exclusive end, <= 2n: the wrapped part lies in the second period
    diff = np.zeros((2 * n + 1,) * 3, dtype=np.int32)
    for sx in (0, 1):
        for sy in (0, 1):
            for sz in (0, 1):
                ix = hi1[:, 0] if sx else lo[:, 0]
                iy = hi1[:, 1] if sy else lo[:, 1]
                iz = hi1[:, 2] if sz else lo[:, 2]
                np.add.at(diff, (ix, iy, iz), -1 if (sx + sy + sz) & 1 else 1)
    cov = diff.cumsum(axis=0).cumsum(axis=1).cumsum(axis=2)[: 2 * n, : 2 * n, : 2 * n] > 0
    del diff
    cov = cov[:n] | cov[n:]
    cov = cov[:, :n] | cov[:, n:]
    return np.ascontiguousarray(cov[:, :, :n] | cov[:, :, n:])


def _cells_of(pos, boxsize, n, xp):
    cs = boxsize / n
    if xp is np:
        return np.clip(np.floor((pos % boxsize) / cs).astype(np.int64), 0, n - 1)
    import torch

    return torch.clamp(torch.floor(torch.remainder(pos, boxsize) / cs).to(torch.int64), 0, n - 1)


def ghost_mask(pos, cofp, read_radius, boxsize, cells_per_dim=64):
    """Particles that must travel with a chunk: like the reference, which reads the SWIFT cells overlapping
    a halo's read region (mask_cells.py:6-38), the box is cut into cells and a particle is kept if its cell is
    touched by some halo's cube [c - r, c + r]^3 (``cell_cover``)."""
    pos = np.asarray(pos, dtype=np.float64)
    cover = cell_cover(cofp, read_radius, boxsize, cells_per_dim)
    cell = _cells_of(pos, boxsize, int(cells_per_dim), np)
    return cover[cell[:, 0], cell[:, 1], cell[:, 2]]


def ghost_mask_device(pos, cofp, read_radius, boxsize, cells_per_dim=64):
    """``ghost_mask`` for positions that are already on the device (torch [N,3]): the cell cover is computed on the
    host from the chunk's halos (a few MB), the per-particle test runs where the particles are.  Returns a bool
    tensor on pos.device."""
    import torch

    cover = torch.as_tensor(cell_cover(cofp, read_radius, boxsize, cells_per_dim), device=pos.device)
    cell = _cells_of(pos, boxsize, int(cells_per_dim), torch)
    return cover[cell[:, 0], cell[:, 1], cell[:, 2]]


def gather_tables(table, index, dst=0, group=None):
    """Gather per-rank [H_r, ncol] tables (+ their halo indices) to rank dst in ONE collective: the index
    travels as an extra float64 column (exact below 2^53).  Tables are torch tensors on the device the
    process group works on (CUDA for NCCL, CPU for gloo); a rank that owns no chunk (more ranks than
    chunks) takes part with an empty table of the common width.  Returns (table, index) on dst, else None."""
    import torch
    import torch.distributed as dist

    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return table, index
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    dev = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend(group) == "nccl" else torch.device("cpu")
    meta = torch.tensor([table.shape[0], table.shape[1] if table.ndim == 2 else 0], dtype=torch.int64, device=dev)
    metas = [torch.zeros_like(meta) for _ in range(world)]
    dist.all_gather(metas, meta, group=group)
    sizes = [int(m[0].item()) for m in metas]
    ncol = max(int(m[1].item()) for m in metas)
    nmax = max(sizes)
    if nmax == 0:
        return (torch.zeros((0, ncol), dtype=torch.float64, device=dev), torch.zeros(0, dtype=torch.int64, device=dev)) if rank == dst else None
    pad = torch.zeros((nmax, ncol + 1), dtype=torch.float64, device=dev)
    if table.shape[0]:
        pad[: table.shape[0], :ncol] = table.to(dev)
        pad[: index.shape[0], ncol] = index.to(dev).to(torch.float64)
    bufs = [torch.empty_like(pad) for _ in range(world)] if rank == dst else None
    dist.gather(pad, bufs, dst=dst, group=group)
    if rank != dst:
        return None
    t = torch.cat([b[:s] for b, s in zip(bufs, sizes)])
    return t[:, :ncol].contiguous(), t[:, ncol].to(torch.int64)


# columns of the result table the re-read loop looks at (soap_result_layout: InputHalos/status,
# InputHalos/search_radius, InputHalos/read_radius)
COL_STATUS, COL_SEARCH_RADIUS, COL_READ_RADIUS = 0, 4, 5
STATUS_RADIUS_TOO_SMALL = 1


def run_chunks(data, halo, boxsize, nr_chunks, compute, rank=0, world_size=1, group=None, reread=False,
               max_passes=20, separate_chunks=None):
    """Process every chunk owned by this rank with ``compute(chunk_data,
    chunk_halos) -> torch [H_c, ncol]`` and gather the rows on rank 0, ordered by
    halo index.  ``data[ptype]`` are numpy arrays of the whole box here (a real
    run reads only the chunk's cells from the snapshot).

    ``reread=True`` adds the repeat loop of ChunkTask.__call__
    (SOAP/core/chunk_tasks.py:188-367): halos that come back with status 1 (the
    search radius outgrew the region that was read, halo_tasks.py:386-402) are
    processed again, on the same rank, from a region re-cut with the enlarged
    read radius the device returned, starting from the search radius it had
    reached; everything else keeps its row from the pass that finished it."""
    import torch

    halo_s, chunk_size = peano_decomposition(boxsize, halo, nr_chunks, separate_chunks=separate_chunks)
    mine = assign_chunks(len(chunk_size), world_size)[rank]
    tables, indices = [], []
    for c in mine:
        hc = chunk_halos(halo_s, chunk_size, c)
        table_c = None
        sel = np.arange(len(hc["index"]))
        for _ in range(max_passes):
            cur = {k: np.ascontiguousarray(np.asarray(v)[sel]) for k, v in hc.items()}
            cd = {}
            for pt, d in data.items():
                m = ghost_mask(d["Coordinates"], cur["cofp"], cur["read_radius"], boxsize)
                cd[pt] = {k: np.ascontiguousarray(v[m]) for k, v in d.items()}
            t = compute(cd, cur)
            if table_c is None:
                table_c = t.clone()
            else:
                table_c[torch.as_tensor(sel, device=t.device)] = t
            if not reread:
                break
            again = (t[:, COL_STATUS] == STATUS_RADIUS_TOO_SMALL).cpu().numpy()
            if not again.any():
                break
            tn = t.detach().cpu().numpy()
            for k in ("search_radius", "read_radius"):
                hc[k] = np.array(hc[k], dtype=np.float64, copy=True)
            hc["search_radius"][sel[again]] = tn[again, COL_SEARCH_RADIUS]
            hc["read_radius"][sel[again]] = tn[again, COL_READ_RADIUS]
            sel = sel[again]
        else:
            # the reference repeats until no halo is left (chunk_tasks.py:188-367); give up loudly instead
            raise RuntimeError(f"chunk {c}: {len(sel)} halos still ask for a larger read radius after {max_passes} passes")
        t = table_c
        bad = (t[:, COL_STATUS] >= 2).cpu().numpy() if reread else np.zeros(0, dtype=bool)
        if bad.any():
            # count mismatch / SO not found within 20 Mpc / root bracket: RuntimeErrors in the reference
            # (subhalo_properties.py:2642-2646, SO_properties.py:150-153,208)
            i = int(np.flatnonzero(bad)[0])
            raise RuntimeError(f"chunk {c}: halo index={int(np.asarray(hc['index'])[i])} failed with status "
                               f"{int(t[i, COL_STATUS].item())} ({int(bad.sum())} halos in this chunk)")
        tables.append(t)
        indices.append(torch.as_tensor(np.asarray(hc["index"], dtype=np.int64), device=t.device))
    if tables:
        table, index = torch.cat(tables), torch.cat(indices)
    else:  # more ranks than chunks: this rank has nothing; gather_tables gives the table its width
        table, index = torch.zeros((0, 0), dtype=torch.float64), torch.zeros(0, dtype=torch.int64)
    got = gather_tables(table, index, 0, group)
    if got is None:
        return None
    t, i = got
    order = torch.argsort(i)
    return t[order], i[order]
