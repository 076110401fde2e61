// small.cu -- fused path for halos whose search sphere fits in shared memory.
//
// One CTA owns one halo from its first ladder rung to its result row:
//   sweep + count + density gate            halo_tasks.py:73-103,166-187
//   gather + halo-centred re-wrap           halo_tasks.py:106-117
//   radial sort (shared memory)             SO_properties.py:398
//   scans, SO / Vmax / half-mass solves     scan.cuh (same code as the general path)
//   moment banks + result row               moments.cuh (same code as the general path)
// and walks the search-radius ladder by itself, so the ~99 % of halos that are
// small never touch the per-rung kernel sequence of the general path (halos.cu)
// and never write their records to global memory.  Halos whose sphere does not
// fit (more than CAP particles, or too many mesh rows / candidates) are handed
// to the next tier with their ladder state intact.
#include "moments.cuh"
#include "scan.cuh"

namespace {

constexpr int SMALL_K = 4;  // records per thread of the scan tiles
constexpr size_t SMALL_SMEM_MAX_WARP = 216 * 1024;  // dynamic shared memory of a lock-step CTA

enum : int { ACT_TRY = 0, ACT_RETRY = 1, ACT_DONE = 2, ACT_OVERFLOW = 3 };

struct SmallCtl {
    uint32_t it;       // queue slot
    uint32_t total;    // candidates of the current rung
    unsigned int n_in, n_stage;
    double msum;
    int action, kacc;
    unsigned long long minr;
    int32_t minfof;
};

// ascending bitonic sort of (rec, pid) by (radius bits, particle slot)
template <int NT>
__device__ inline void sort_records(Rec* rec, uint32_t* pid, uint32_t n) {
    uint32_t np2 = 1;
    while (np2 < n) np2 <<= 1;
    auto cx = [&](uint32_t t, uint32_t p) {
        if (p > t && p < n) {
            const Rec a = rec[t], b = rec[p];
            const uint32_t ia = pid[t], ib = pid[p];
            if (b.rbits < a.rbits || (b.rbits == a.rbits && ib < ia)) {
                rec[t] = b; rec[p] = a;
                pid[t] = ib; pid[p] = ia;
            }
        }
    };
    for (uint32_t k = 2; k <= np2; k <<= 1) {
        for (uint32_t t = threadIdx.x; t < np2; t += NT) cx(t, t ^ (k - 1));
        __syncthreads();
        for (uint32_t j = k >> 2; j > 0; j >>= 1) {
            for (uint32_t t = threadIdx.x; t < np2; t += NT) cx(t, t ^ j);
            __syncthreads();
        }
    }
}

template <int NCH, int V, int NT, int CAP>
__global__ void __launch_bounds__(NT, (V <= 16 ? 512 : 256) / NT > 16 ? 16 : (V <= 16 ? 512 : 256) / NT) k_small_halos(ChunkView v, HaloArrays ha, DevCfg cfg,
                                                    const uint32_t* __restrict__ list,
                                                    const unsigned int* __restrict__ n_list,
                                                    uint32_t* __restrict__ overflow,
                                                    unsigned int* __restrict__ n_overflow,
                                                    unsigned int* __restrict__ queue_cursor, Counters* ctr,
                                                    int bank_stride) {
    constexpr int NW = NT / 32;
    constexpr int NTY = NCH == 2 ? 1 : 4;
    constexpr int VP = BankAcc<V>::VP;
    constexpr uint32_t CAND_MAX = 64u * CAP;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    Rec* rec = (Rec*)smem_raw;                        // [CAP] the sphere's records, radially sorted
    uint32_t* pid = (uint32_t*)(rec + CAP);           // [CAP] particle slot of each record
    double* banks = (double*)(pid + CAP);             // [NW][bank_stride]
    double* stage = banks + (size_t)NW * bank_stride;  // [NW][32][VP]
    int* skey = (int*)(stage + (size_t)NW * 32 * VP);  // [NW][32]
    __shared__ ScanShared<NCH, NT> S;
    __shared__ Cuts cuts;
    __shared__ DimRanges rg[3];
    __shared__ uint32_t row_s0[NT], row_off[NT + 1];
    __shared__ uint32_t w_u32[NW];
    __shared__ unsigned long long w_u64[NW];
    __shared__ int32_t w_i32[NW];
    __shared__ SmallCtl B;
    __shared__ uint32_t w_cnt[NW][4];
    __shared__ double w_msum[NW][4];
    __shared__ int nloop_s;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const double L = v.L, halfL = 0.5 * v.L;

    // candidate j of the current rung -> particle slot (rows are concatenated)
    auto cand_slot = [&](uint32_t j, int nrows) -> uint32_t {
        int lo = 0, hi = nrows;  // largest lo with row_off[lo] <= j
        while (hi - lo > 1) {
            const int mid = (lo + hi) >> 1;
            if (row_off[mid] <= j) lo = mid; else hi = mid;
        }
        return row_s0[lo] + (j - row_off[lo]);
    };

    while (true) {
        __syncthreads();
        if (threadIdx.x == 0) B.it = atomicAdd(queue_cursor, 1u);
        __syncthreads();
        if (B.it >= *n_list) break;
        const uint32_t h = list[B.it];
        const double cx = ha.cofp[3 * h], cy = ha.cofp[3 * h + 1], cz = ha.cofp[3 * h + 2];
        const int32_t hidx = (int32_t)ha.index[h];
        const bool central = ha.central[h] == 1;
        const int n_so = central ? cfg.n_so : 0;
        double cur = ha.cur_r[h];
        if (threadIdx.x == 0) nloop_s = ha.nloop[h];
        bool look1 = false;  // the look-ahead sphere did not fit: sweep the current rung alone
        int action = ACT_DONE;
        while (true) {
            // ------------------------ rows of the furthest of the next rungs (ladder look-ahead:
            // one sweep bins the sphere by rung, like k_count; if that sphere does not fit this
            // tier the current rung is swept alone)
            constexpr int LOOK = 4;
            double rr[LOOK];
            const int nr = ladder_radii(cur, ha.rr_in[h], look1 ? 1 : LOOK, rr);
            __syncthreads();
            if (threadIdx.x < 3) halo_ranges(v, cx, cy, cz, rr[nr - 1], rg, threadIdx.x);
            __syncthreads();
            const RowIter ri = row_iter(rg);
            bool too_big = ri.nrows > NT;
            uint32_t total = 0;
            if (!too_big) {
                uint32_t s0 = 0, s1 = 0;
                if ((int)threadIdx.x < ri.nrows) row_span(v, rg, ri, threadIdx.x, s0, s1);
                const uint32_t len = s1 - s0;
                uint32_t incl = len;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
                    if (lane >= o) incl += t;
                }
                if (lane == 31) w_u32[wid] = incl;
                __syncthreads();
                uint32_t base = 0, all = 0;
                for (int w = 0; w < NW; w++) {
                    if (w < wid) base += w_u32[w];
                    all += w_u32[w];
                }
                row_s0[threadIdx.x] = s0;
                row_off[threadIdx.x] = base + incl - len;
                if (threadIdx.x == NT - 1) row_off[NT] = all;
                total = all;
                too_big = total > CAND_MAX;
                __syncthreads();
            }
            if (too_big) {
                if (nr > 1) { look1 = true; continue; }
                action = ACT_OVERFLOW;
                break;
            }
            look1 = false;
            // ------------------------------------ count + enclosed mass per rung (halo_tasks.py:84-97)
            double r2k[LOOK];
#pragma unroll
            for (int k = 0; k < LOOK; k++) r2k[k] = k < nr ? __dmul_rn(rr[k], rr[k]) : -1.0;
            {
                uint32_t cnt[LOOK];
                double msum[LOOK];
#pragma unroll
                for (int k = 0; k < LOOK; k++) { cnt[k] = 0; msum[k] = 0.0; }
                for (uint32_t j = threadIdx.x; j < total; j += NT) {
                    const uint32_t t = cand_slot(j, ri.nrows);
                    const double r2 = periodic_r2(v.px[t], v.py[t], v.pz[t], cx, cy, cz, L, halfL);
                    if (r2 <= r2k[nr - 1]) {
                        const double m = (double)v.mass[t];
                        bool placed = false;
#pragma unroll
                        for (int k = 0; k < LOOK; k++)
                            if (!placed && k < nr && r2 <= r2k[k]) { cnt[k]++; msum[k] += m; placed = true; }
                    }
                }
#pragma unroll
                for (int k = 0; k < LOOK; k++) {
                    const uint32_t c = (uint32_t)warp_sum_u64(cnt[k]);
                    const double m = warp_sum(msum[k]);
                    if (lane == 0) { w_cnt[wid][k] = c; w_msum[wid][k] = m; }
                }
                __syncthreads();
                if (threadIdx.x == 0) {
                    // density gate and ladder steps (halo_tasks.py:73-103,166-187)
                    const bool has_target = central && cfg.target_density > 0.0;  // halo_tasks.py:381
                    uint32_t ccum = 0;
                    double mcum = 0.0;
                    int act = ACT_RETRY, kacc = 0;
                    bool pending = true;
                    for (int k = 0; k < nr && pending; k++) {
                        nloop_s++;  // halo_tasks.py:75
                        const double r = ha.cur_r[h];
                        for (int w = 0; w < NW; w++) { ccum += w_cnt[w][k]; mcum += w_msum[w][k]; }
                        const double density = mcum / (4.0 / 3.0 * SOAP_PI * (r * r * r));
                        if (!has_target || density <= cfg.target_density) {
                            kacc = k;
                            if (ccum > (uint32_t)CAP) {
                                act = ACT_OVERFLOW;
                                nloop_s--;  // the next tier repeats this rung
                            } else {
                                act = ACT_TRY;
                                // what k_plan_items / k_gate leave behind for the scan and moment stages
                                ha.cnt[h] = ccum;
                                ha.msum[h] = mcum;
                                ha.rung_r[h] = r;
                                ha.commit_lo[h] = ha.commit_hi[h] = ha.ndone[h];
                                ha.state[h] = ST_TRY;
                            }
                            break;
                        }
                        pending = ladder_step(ha, h, 0.0);
                        if (!pending) act = ACT_DONE;
                    }
                    B.n_in = ccum; B.msum = mcum; B.action = act; B.n_stage = 0; B.kacc = kacc;
                    atomicAdd(&ctr->candidates, (unsigned long long)total);
                    atomicAdd(&ctr->count_pairs, (unsigned long long)ccum);
                }
                __syncthreads();
            }
            action = B.action;
            if (action == ACT_RETRY) { cur = ha.cur_r[h]; continue; }
            if (action != ACT_TRY) break;
            cur = rr[B.kacc];
            const double r2max = r2k[B.kacc];
            const uint32_t n = B.n_in;
            // ------------------------------------- gather + re-wrap (halo_tasks.py:106-117)
            {
                unsigned long long minr = ~0ull;
                int32_t minfof = -1;
                for (uint32_t j0 = 0; j0 < total; j0 += NT) {
                    const uint32_t j = j0 + threadIdx.x;
                    bool in = false;
                    uint32_t t = 0;
                    if (j < total) {
                        t = cand_slot(j, ri.nrows);
                        const double r2 = periodic_r2(v.px[t], v.py[t], v.pz[t], cx, cy, cz, L, halfL);
                        in = r2 <= r2max;
                    }
                    const unsigned bal = __ballot_sync(0xffffffffu, in);
                    unsigned base = 0;
                    if (lane == 0 && bal) base = atomicAdd(&B.n_stage, (unsigned)__popc(bal));
                    base = __shfl_sync(0xffffffffu, base, 0);
                    if (in) {
                        const uint32_t slot = base + __popc(bal & ((1u << lane) - 1u));
                        const Part p = rel_part(v, t, cx, cy, cz, halfL);
                        Rec rc;
                        rc.rbits = (unsigned long long)__double_as_longlong(p.r);
                        rc.m = v.mass[t];
                        const uint32_t tc = NCH == 2 ? 1u : (uint32_t)v.type[t];
                        rc.flags = tc | ((v.grnr[t] == hidx) ? 4u : 0u);
                        rec[slot] = rc;
                        pid[slot] = t;
                        const int32_t f = v.fof[t];
                        if (rc.rbits < minr || (rc.rbits == minr && f < minfof)) { minr = rc.rbits; minfof = f; }
                    }
                }
                // fofid of the innermost particle (SO_properties.py:407-409)
                for (int o = 16; o > 0; o >>= 1) {
                    const unsigned long long orr = __shfl_xor_sync(0xffffffffu, minr, o);
                    const int32_t of = __shfl_xor_sync(0xffffffffu, minfof, o);
                    if (orr < minr || (orr == minr && of < minfof)) { minr = orr; minfof = of; }
                }
                if (lane == 0) { w_u64[wid] = minr; w_i32[wid] = minfof; }
                __syncthreads();
                if (threadIdx.x == 0) {
                    for (int w = 1; w < NW; w++)
                        if (w_u64[w] < minr || (w_u64[w] == minr && w_i32[w] < minfof)) { minr = w_u64[w]; minfof = w_i32[w]; }
                    B.minr = minr;
                    B.minfof = minfof;
                }
                __syncthreads();
            }
            // ------------------------------------------------ radial sort
            sort_records<NT>(rec, pid, n);
            // ------------------------------------------- scans and solves
            scan_solve_halo<NCH, 1, NT, SMALL_K>(S, ha, cfg, h, n, rec, nullptr, ctr, &B.minr, &B.minfof, 1u);
            __syncthreads();
            const int c_lo = S.commit_lo_, c_hi = S.commit_hi_, fail = S.fail_;
            // -------------------------------------- moments of the committed properties
            if (c_hi > c_lo && fail < 2) {
                const ScanRes* sr = ha.sres + h;
                const bool sub_c = cfg.do_sub && c_lo == 0;
                if (threadIdx.x == 32 % NT) build_cuts(cuts, cfg, sr, c_lo, c_hi, n_so);
                __syncthreads();
                const int ncut = cuts.n;
                const int nbank = (ncut + 1) * 2 * NTY;
                for (int w = 0; w < NW; w++)
                    for (int i = threadIdx.x; i < nbank * V; i += NT) banks[(size_t)w * bank_stride + i] = 0.0;
                __syncthreads();
                const int32_t cen_fof = sr->cen_fof;
                double* stage_w = stage + (size_t)wid * 32 * VP;
                int* skey_w = skey + wid * 32;
                double* bank_w = banks + (size_t)wid * bank_stride;
                BankAcc<V> ba;
                ba.init();
                // each warp takes a contiguous run of the sorted records: shells change rarely
                const uint32_t per = ((n + NW - 1) / NW + 31u) & ~31u;
                const uint32_t lo = wid * per, hi = lo + per < n ? lo + per : n;
                for (uint32_t b0 = lo; b0 < hi; b0 += 32) {
                    const uint32_t i = b0 + lane;
                    const bool in = i < hi;
                    int key = 0;
                    double val[V];
                    if (in) {
                        const uint32_t t = pid[i];
                        const double x = rewrap_rel(v.px[t], cx, L, halfL);
                        const double y = rewrap_rel(v.py[t], cy, L, halfL);
                        const double z = rewrap_rel(v.pz[t], cz, L, halfL);
                        const double r = radius3(x, y, z);
                        key = moment_terms<V, NTY>(cuts, ncut, cfg, x, y, z, r, (double)v.mass[t], (double)v.vx[t],
                                                   (double)v.vy[t], (double)v.vz[t], v.grnr[t], hidx, v.fof[t],
                                                   cen_fof, NTY == 1 ? 1u : (uint32_t)v.type[t], val);
                    }
                    ba.add(in, key, val, stage_w, skey_w, bank_w, 1, lane);
                }
                ba.flush(bank_w, 1, lane);
                __syncthreads();
                for (int i = threadIdx.x; i < nbank * V; i += NT) {
                    double s = banks[i];
                    for (int w = 1; w < NW; w++) s += banks[(size_t)w * bank_stride + i];
                    banks[i] = s;
                }
                __syncthreads();
                write_row<V, NTY>(banks, cuts, ncut, cfg, ha, h, sr, sub_c, n_so, cx, cy, cz, (int)threadIdx.x);
                __syncthreads();
            }
            if (fail == 1 && ha.state[h] == ST_PENDING) {
                cur = ha.cur_r[h];
                action = ACT_RETRY;
                continue;
            }
            action = ACT_DONE;
            break;
        }
        if (threadIdx.x == 0) {
            ha.nloop[h] = nloop_s;
            if (action == ACT_OVERFLOW) {
                ha.state[h] = ST_PENDING;
                overflow[atomicAdd(n_overflow, 1u)] = h;
            }
        }
    }
}

// ===================================================== lock-step warp tiers
// One WARP owns one halo; the W warps of a CTA walk the phases (ladder rungs,
// gather + sort, scan + solve, moments + row) in lock step, separated by CTA-wide
// alignment barriers, so that the large per-halo instruction stream is fetched
// once per CTA instead of once per warp (a warp-per-CTA version of this kernel
// spent 88 % of its issue slots waiting for instructions).
template <int NCH, int V, int CAP>
struct __align__(16) WarpSlot {
    Rec rec[CAP];
    uint32_t pid[CAP];
    // the scan scratch (phase 3) and the moment staging tile (phase 4) are never live together
    union {
        ScanShared<NCH, 32> S;
        struct {
            double stage[32 * BankAcc<V>::VP];
            int skey[32];
        };
    };
    Cuts cuts;
    DimRanges rg[3];
    uint32_t row_s0[32], row_off[33];
    unsigned long long minr;
    int32_t minfof;
    unsigned int n_stage;
    // followed by bank_stride doubles of banks
};

enum : int { WS_NEED = 0, WS_RUNG = 1, WS_TRY = 2, WS_EXHAUSTED = 3 };

// MAXW = most warps a CTA is launched with (sets the register budget); gbanks != nullptr: the
// moment banks of a warp live in global memory (hydro configurations: 4 types x 36 terms per
// shell do not fit next to the records of enough warps; the banks are touched only on a key change).
template <int NCH, int V, int CAP, int MAXW>
__global__ void __launch_bounds__(32 * MAXW, 1) k_small_warps(ChunkView v, HaloArrays ha, DevCfg cfg,
                                                        const uint32_t* __restrict__ list,
                                                        const unsigned int* __restrict__ n_list,
                                                        uint32_t* __restrict__ overflow,
                                                        unsigned int* __restrict__ n_overflow,
                                                        unsigned int* __restrict__ queue_cursor, Counters* ctr,
                                                        int bank_stride, int slot_bytes, double* gbanks) {
    constexpr int NTY = NCH == 2 ? 1 : 4;
    constexpr uint32_t CAND_MAX = 64u * CAP;
    using Slot = WarpSlot<NCH, V, CAP>;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    Slot& W = *(Slot*)(smem_raw + (size_t)wid * slot_bytes);
    double* banks = gbanks ? gbanks + ((size_t)blockIdx.x * (blockDim.x >> 5) + wid) * bank_stride
                           : (double*)(smem_raw + (size_t)wid * slot_bytes + sizeof(Slot));
    const double L = v.L, halfL = 0.5 * v.L;
    const unsigned int n_total = *n_list;

    auto cand_slot = [&](uint32_t j, int nrows) -> uint32_t {
        int lo = 0, hi = nrows;  // largest lo with row_off[lo] <= j
        while (hi - lo > 1) {
            const int mid = (lo + hi) >> 1;
            if (W.row_off[mid] <= j) lo = mid; else hi = mid;
        }
        return W.row_s0[lo] + (j - W.row_off[lo]);
    };

    // warp-uniform halo state
    int state = WS_NEED;
    uint32_t h = 0, n = 0, total = 0;
    double cx = 0, cy = 0, cz = 0, cur = 0, r2max = 0;
    int32_t hidx = 0;
    bool central = false;
    int n_so = 0, nloop = 0, nrows = 0;
    bool look1 = false;  // the look-ahead sphere did not fit: sweep the current rung alone

    while (true) {
        // ============ phase 1: fetch halos and walk their ladder until one needs a solve
        while (state == WS_NEED || state == WS_RUNG) {
            if (state == WS_NEED) {
                unsigned int it = 0;
                if (lane == 0) it = atomicAdd(queue_cursor, 1u);
                it = __shfl_sync(0xffffffffu, it, 0);
                if (it >= n_total) { state = WS_EXHAUSTED; break; }
                h = list[it];
                cx = ha.cofp[3 * h]; cy = ha.cofp[3 * h + 1]; cz = ha.cofp[3 * h + 2];
                hidx = (int32_t)ha.index[h];
                central = ha.central[h] == 1;
                n_so = central ? cfg.n_so : 0;
                cur = ha.cur_r[h];
                nloop = ha.nloop[h];
                state = WS_RUNG;
            }
            // ---- rows of the furthest of the next rungs (ladder look-ahead: one sweep bins the
            // sphere by rung, like k_count; a sphere too large for this tier is retried rung by rung)
            constexpr int LOOK = 4;
            double rr[LOOK];
            int nr = ladder_radii(cur, ha.rr_in[h], look1 ? 1 : LOOK, rr);
            const double rsweep = rr[nr - 1];
            __syncwarp();
            if (lane < 3) halo_ranges(v, cx, cy, cz, rsweep, W.rg, lane);
            __syncwarp();
            const RowIter ri = row_iter(W.rg);
            nrows = ri.nrows;
            bool too_big = nrows > 32;
            if (!too_big) {
                uint32_t s0 = 0, s1 = 0;
                if (lane < nrows) row_span(v, W.rg, ri, lane, s0, s1);
                const uint32_t len = s1 - s0;
                uint32_t incl = len;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
                    if (lane >= o) incl += t;
                }
                W.row_s0[lane] = s0;
                W.row_off[lane] = incl - len;
                total = __shfl_sync(0xffffffffu, incl, 31);
                if (lane == 31) W.row_off[32] = total;
                too_big = total > CAND_MAX;
                __syncwarp();
            }
            int action;
            if (too_big) {
                if (nr > 1) { look1 = true; continue; }  // try again with this rung alone
                action = ACT_OVERFLOW;
            } else {
                // ---- count + enclosed mass per rung (halo_tasks.py:84-97)
                double r2k[LOOK];
#pragma unroll
                for (int k = 0; k < LOOK; k++) r2k[k] = k < nr ? __dmul_rn(rr[k], rr[k]) : -1.0;
                uint32_t cnt[LOOK];
                double msum[LOOK];
#pragma unroll
                for (int k = 0; k < LOOK; k++) { cnt[k] = 0; msum[k] = 0.0; }
                for (uint32_t j = lane; j < total; j += 32) {
                    const uint32_t t = cand_slot(j, nrows);
                    const double r2 = periodic_r2(v.px[t], v.py[t], v.pz[t], cx, cy, cz, L, halfL);
                    if (r2 <= r2k[nr - 1]) {
                        const double m = (double)v.mass[t];
                        bool placed = false;
#pragma unroll
                        for (int k = 0; k < LOOK; k++)
                            if (!placed && k < nr && r2 <= r2k[k]) { cnt[k]++; msum[k] += m; placed = true; }
                    }
                }
#pragma unroll
                for (int k = 0; k < LOOK; k++) {
                    cnt[k] = (uint32_t)warp_sum_u64(cnt[k]);
                    msum[k] = warp_sum(msum[k]);
                }
                // ---- density gate and ladder steps (halo_tasks.py:73-103,166-187)
                action = ACT_RETRY;
                uint32_t ccum = 0;
                int kacc = 0;
                if (lane == 0) {
                    const bool has_target = central && cfg.target_density > 0.0;  // halo_tasks.py:381
                    double mcum = 0.0;
                    bool pending = true;
                    for (int k = 0; k < nr && pending; k++) {
                        nloop++;  // halo_tasks.py:75
                        const double r = ha.cur_r[h];
                        ccum += cnt[k];
                        mcum += msum[k];
                        const double density = mcum / (4.0 / 3.0 * SOAP_PI * (r * r * r));
                        if (!has_target || density <= cfg.target_density) {
                            kacc = k;
                            if (ccum > (uint32_t)CAP) {
                                action = ACT_OVERFLOW;
                                nloop--;  // the next tier repeats this rung
                            } else {
                                action = ACT_TRY;
                                // what k_plan_items / k_gate leave behind for the scan and moment stages
                                ha.cnt[h] = ccum;
                                ha.msum[h] = mcum;
                                ha.rung_r[h] = r;
                                ha.commit_lo[h] = ha.commit_hi[h] = ha.ndone[h];
                                ha.state[h] = ST_TRY;
                            }
                            break;
                        }
                        pending = ladder_step(ha, h, 0.0);
                        if (!pending) action = ACT_DONE;
                    }
                    atomicAdd(&ctr->candidates, (unsigned long long)total);
                    atomicAdd(&ctr->count_pairs, (unsigned long long)ccum);
                }
                action = __shfl_sync(0xffffffffu, action, 0);
                nloop = __shfl_sync(0xffffffffu, nloop, 0);
                n = __shfl_sync(0xffffffffu, ccum, 0);
                kacc = __shfl_sync(0xffffffffu, kacc, 0);
                if (action == ACT_TRY) {
                    cur = rr[kacc];
                    r2max = r2k[kacc];
                }
            }
            look1 = false;
            if (action == ACT_RETRY) {
                cur = __shfl_sync(0xffffffffu, lane == 0 ? ha.cur_r[h] : 0.0, 0);
            } else if (action == ACT_TRY) {
                state = WS_TRY;
            } else {
                if (lane == 0) {
                    ha.nloop[h] = nloop;
                    if (action == ACT_OVERFLOW) {
                        ha.state[h] = ST_PENDING;
                        overflow[atomicAdd(n_overflow, 1u)] = h;
                    }
                }
                state = WS_NEED;
            }
        }
        // every warp is now in WS_TRY or WS_EXHAUSTED
        {
            int all_done;
            asm volatile(
                "{\n .reg .pred p, q;\n setp.ne.s32 p, %1, 0;\n barrier.red.and.pred q, 1, %2, p;\n"
                " selp.s32 %0, 1, 0, q;\n}"
                : "=r"(all_done)
                : "r"((int)(state == WS_EXHAUSTED)), "r"(blockDim.x)
                : "memory");
            if (all_done) break;
        }
        // ============ phase 2: gather + re-wrap (halo_tasks.py:106-117) + radial sort
        if (state == WS_TRY) {
            if (lane == 0) W.n_stage = 0;
            __syncwarp();
            unsigned long long minr = ~0ull;
            int32_t minfof = -1;
            for (uint32_t j0 = 0; j0 < total; j0 += 32) {
                const uint32_t j = j0 + lane;
                bool in = false;
                uint32_t t = 0;
                if (j < total) {
                    t = cand_slot(j, nrows);
                    const double r2 = periodic_r2(v.px[t], v.py[t], v.pz[t], cx, cy, cz, L, halfL);
                    in = r2 <= r2max;
                }
                const unsigned bal = __ballot_sync(0xffffffffu, in);
                const unsigned base = W.n_stage;
                __syncwarp();
                if (lane == 0) W.n_stage = base + __popc(bal);
                if (in) {
                    const uint32_t slot = base + __popc(bal & ((1u << lane) - 1u));
                    const Part p = rel_part(v, t, cx, cy, cz, halfL);
                    Rec rc;
                    rc.rbits = (unsigned long long)__double_as_longlong(p.r);
                    rc.m = v.mass[t];
                    const uint32_t tc = NCH == 2 ? 1u : (uint32_t)v.type[t];
                    rc.flags = tc | ((v.grnr[t] == hidx) ? 4u : 0u);
                    W.rec[slot] = rc;
                    W.pid[slot] = t;
                    const int32_t f = v.fof[t];
                    if (rc.rbits < minr || (rc.rbits == minr && f < minfof)) { minr = rc.rbits; minfof = f; }
                }
                __syncwarp();
            }
            // fofid of the innermost particle (SO_properties.py:407-409)
            for (int o = 16; o > 0; o >>= 1) {
                const unsigned long long orr = __shfl_xor_sync(0xffffffffu, minr, o);
                const int32_t of = __shfl_xor_sync(0xffffffffu, minfof, o);
                if (orr < minr || (orr == minr && of < minfof)) { minr = orr; minfof = of; }
            }
            if (lane == 0) { W.minr = minr; W.minfof = minfof; }
            __syncwarp();
            // bitonic sort by (radius bits, particle slot), one warp
            uint32_t np2 = 1;
            while (np2 < n) np2 <<= 1;
            auto cxg = [&](uint32_t t, uint32_t p) {
                if (p > t && p < n) {
                    const Rec a = W.rec[t], b = W.rec[p];
                    const uint32_t ia = W.pid[t], ib = W.pid[p];
                    if (b.rbits < a.rbits || (b.rbits == a.rbits && ib < ia)) {
                        W.rec[t] = b; W.rec[p] = a;
                        W.pid[t] = ib; W.pid[p] = ia;
                    }
                }
            };
            for (uint32_t k = 2; k <= np2; k <<= 1) {
                for (uint32_t t = lane; t < np2; t += 32) cxg(t, t ^ (k - 1));
                __syncwarp();
                for (uint32_t j = k >> 2; j > 0; j >>= 1) {
                    for (uint32_t t = lane; t < np2; t += 32) cxg(t, t ^ j);
                    __syncwarp();
                }
            }
        }
        align_bar();
        // ============ phase 3: scans and solves (two more alignment barriers inside)
        if (state == WS_TRY) {
            // tiles of 64 records for the smallest spheres: more lanes busy per pass
            scan_solve_halo<NCH, 1, 32, (CAP <= 256 ? 2 : SMALL_K), true>(W.S, ha, cfg, h, n, W.rec, nullptr, ctr, &W.minr,
                                                                          &W.minfof, 1u);
            __syncwarp();
        } else {
            align_bar(2);
            align_bar(2);
        }
        align_bar();
        // ============ phase 4: moments of the committed properties + result row
        if (state == WS_TRY) {
            const int c_lo = W.S.commit_lo_, c_hi = W.S.commit_hi_, fail = W.S.fail_;
            __syncwarp();  // W.S is dead from here: its storage becomes the staging tile
            if (c_hi > c_lo && fail < 2) {
                const ScanRes* sr = ha.sres + h;
                const bool sub_c = cfg.do_sub && c_lo == 0;
                if (lane == 0) build_cuts(W.cuts, cfg, sr, c_lo, c_hi, n_so);
                __syncwarp();
                const int ncut = W.cuts.n;
                const int nbank = (ncut + 1) * 2 * NTY;
                for (int i = lane; i < nbank * V; i += 32) banks[i] = 0.0;
                __syncwarp();
                const int32_t cen_fof = sr->cen_fof;
                BankAcc<V> ba;
                ba.init();
                for (uint32_t b0 = 0; b0 < n; b0 += 32) {
                    const uint32_t i = b0 + lane;
                    const bool in = i < n;
                    int key = 0;
                    double val[V];
                    if (in) {
                        const uint32_t t = W.pid[i];
                        const double x = rewrap_rel(v.px[t], cx, L, halfL);
                        const double y = rewrap_rel(v.py[t], cy, L, halfL);
                        const double z = rewrap_rel(v.pz[t], cz, L, halfL);
                        const double r = radius3(x, y, z);
                        key = moment_terms<V, NTY>(W.cuts, ncut, cfg, x, y, z, r, (double)v.mass[t], (double)v.vx[t],
                                                   (double)v.vy[t], (double)v.vz[t], v.grnr[t], hidx, v.fof[t],
                                                   cen_fof, NTY == 1 ? 1u : (uint32_t)v.type[t], val);
                    }
                    ba.add(in, key, val, W.stage, W.skey, banks, 1, lane);
                }
                ba.flush(banks, 1, lane);
                __syncwarp();
                write_row<V, NTY>(banks, W.cuts, ncut, cfg, ha, h, sr, sub_c, n_so, cx, cy, cz, lane);
                __syncwarp();
                if constexpr (NCH == 8 && V >= V_FULL) if (cfg.flags & PF_KAPPA) {
                    // kappa_corot / DtoT / stellar rotation need the finished vcom and L of the row:
                    // a second pass over the gas and star records (scratch aliases the staging tile)
                    constexpr int KS = 1 + SOAP_MAX_APERTURES;
                    static_assert(sizeof(W.stage) >= KS * (sizeof(KapSel) + 11 * sizeof(double)), "kappa scratch");
                    KapSel* ksel = reinterpret_cast<KapSel*>(W.stage);
                    double(*kacc)[11] = reinterpret_cast<double(*)[11]>(ksel + KS);
                    int ns = 0;
                    if (lane == 0) ns = kappa_build_sels(ksel, cfg, ha, h, c_lo, c_hi);
                    ns = __shfl_sync(0xffffffffu, ns, 0);
                    if (ns > 0) {
                        for (int i = lane; i < ns * 11; i += 32) (&kacc[0][0])[i] = 0.0;
                        __syncwarp();
                        for (uint32_t i = lane; i < n; i += 32) {
                            const Rec rc = W.rec[i];
                            const uint32_t tc = rc.flags & 3u;
                            if (tc != 0u && tc != 2u) continue;
                            const uint32_t t = W.pid[i];
                            const double x = rewrap_rel(v.px[t], cx, L, halfL);
                            const double y = rewrap_rel(v.py[t], cy, L, halfL);
                            const double z = rewrap_rel(v.pz[t], cz, L, halfL);
                            kappa_add(ksel, ns, kacc, x, y, z, radius3(x, y, z), (double)v.mass[t], (double)v.vx[t],
                                      (double)v.vy[t], (double)v.vz[t], tc, (rc.flags & 4u) != 0u);
                        }
                        __syncwarp();
                        for (int i = lane; i < ns * 11; i += 32) {
                            const double a = kacc[i / 11][i % 11];
                            if (a != 0.0) ksel[i / 11].out[i % 11] += a;
                        }
                        __syncwarp();
                        if (lane == 0) kappa_finish_row(cfg, ha, h, c_lo, c_hi);
                        __syncwarp();
                    }
                }
            }
            int st = 0;
            if (lane == 0) st = ha.state[h];
            st = __shfl_sync(0xffffffffu, st, 0);
            if (fail == 1 && st == ST_PENDING) {
                cur = __shfl_sync(0xffffffffu, lane == 0 ? ha.cur_r[h] : 0.0, 0);
                state = WS_RUNG;
            } else {
                if (lane == 0) ha.nloop[h] = nloop;
                state = WS_NEED;
            }
        }
        align_bar();
    }
}

template <int NCH, int V, int CAP>
int launch_warp_tier(soap_chunk* c, const DevCfg& cfg, const HaloArrays& ha, const uint32_t* list,
                     const unsigned int* n_list, uint32_t* overflow, unsigned int* n_overflow,
                     unsigned int* queue_cursor, Counters* ctr, int bank_stride, unsigned int n_upper,
                     cudaStream_t stream) {
    soap_handle* h = c->h;
    constexpr int MAXW = (NCH == 2 && V <= 16) ? 16 : 8;  // fewer warps, more registers for the wide variants
    const size_t bank_bytes = (size_t)bank_stride * sizeof(double);
    size_t slot = (sizeof(WarpSlot<NCH, V, CAP>) + bank_bytes + 15) & ~(size_t)15;
    bool global_banks = false;
    if ((int)(SMALL_SMEM_MAX_WARP / slot) < MAXW && bank_bytes > 4096) {
        // banks to global memory: more warps per CTA matter more than the bank latency
        global_banks = true;
        slot = (sizeof(WarpSlot<NCH, V, CAP>) + 15) & ~(size_t)15;
    }
    int nw = (int)(SMALL_SMEM_MAX_WARP / slot);
    if (nw > MAXW) nw = MAXW;
    if (nw < 1) SOAP_FAIL("soap_process_halos: a warp slot of %zu bytes does not fit in shared memory", slot);
    auto kern = k_small_warps<NCH, V, CAP, MAXW>;
    const size_t smem = slot * nw;
    CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    unsigned int grid = (unsigned int)h->sm_count;
    const unsigned int need = (n_upper + nw - 1) / nw;
    if (grid > need) grid = need < 1 ? 1 : need;
    double* gbanks = nullptr;
    if (global_banks) {
        gbanks = (double*)h->get("h_tier_gbanks", bank_bytes * (size_t)grid * nw);
        if (!gbanks) return -1;
    }
    LAUNCH(h, kern, grid, 32 * nw, smem, stream, c->v, ha, cfg, list, n_list, overflow, n_overflow, queue_cursor, ctr,
           bank_stride, (int)slot, gbanks);
    return 0;
}

template <int V, int NT, int CAP>
size_t tier_smem(int bank_stride) {
    constexpr int NW = NT / 32;
    return (size_t)CAP * (sizeof(Rec) + sizeof(uint32_t)) + (size_t)NW * bank_stride * sizeof(double) +
           (size_t)NW * 32 * BankAcc<V>::VP * sizeof(double) + (size_t)NW * 32 * sizeof(int);
}
constexpr size_t SMALL_SMEM_MAX = 160 * 1024;

int bank_stride_of(const DevCfg& cfg) {
    const bool full = (cfg.flags & (PF_KIN | PF_KAPPA | PF_TENS)) != 0;
    return (cfg.n_so + cfg.n_ap + 3) * 2 * (cfg.dmo ? 1 : 4) * (full ? V_FULL : V_MIN);
}

template <int NCH, int V, int NT, int CAP>
int launch_tier(soap_chunk* c, const DevCfg& cfg, const HaloArrays& ha, const uint32_t* list,
                const unsigned int* n_list, uint32_t* overflow, unsigned int* n_overflow,
                unsigned int* queue_cursor, Counters* ctr, int bank_stride, unsigned int max_ctas,
                cudaStream_t stream) {
    soap_handle* h = c->h;
    const size_t smem = tier_smem<V, NT, CAP>(bank_stride);
    auto kern = k_small_halos<NCH, V, NT, CAP>;
    CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int per_sm = 0;
    CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, NT, smem));
    if (per_sm < 1) SOAP_FAIL("soap_process_halos: fused tier does not fit on an SM (%zu bytes of shared memory)", smem);
    unsigned int grid = (unsigned int)(per_sm * h->sm_count);
    if (grid > max_ctas) grid = max_ctas;
    if (grid < 1) grid = 1;
    LAUNCH(h, kern, grid, NT, smem, stream, c->v, ha, cfg, list, n_list, overflow, n_overflow, queue_cursor, ctr,
           bank_stride);
    return 0;
}

}  // namespace

// Tiers 0 and 1: spheres of up to 256 / 512 particles, one warp per halo, the
// warps of a CTA in lock step (k_small_warps); tier 2: up to 2048, one CTA of
// 256 threads per halo (k_small_halos).  A tier is used only if it fits.
int soap_small_tier_fits(const DevCfg& cfg, int tier) {
    const bool full = (cfg.flags & (PF_KIN | PF_KAPPA | PF_TENS)) != 0;
    const int stride = bank_stride_of(cfg);
    const size_t bank_bytes = (size_t)stride * sizeof(double);
    size_t slot = 0;
    if (tier == 2) {
        const size_t smem = full ? tier_smem<V_FULL, 256, 2048>(stride) : tier_smem<V_MIN, 256, 2048>(stride);
        return smem <= SMALL_SMEM_MAX ? 1 : 0;
    }
#define SLOT(NCH, VV) (tier == 0 ? sizeof(WarpSlot<NCH, VV, 256>) : sizeof(WarpSlot<NCH, VV, 512>))
    if (cfg.dmo) slot = full ? SLOT(2, V_FULL) : SLOT(2, V_MIN);
    else slot = full ? SLOT(8, V_FULL) : SLOT(8, V_MIN);
#undef SLOT
    // at least four warps per CTA, or lock step buys nothing
    (void)bank_bytes;  // large banks move to global memory (launch_warp_tier)
    return (slot + 16) * 4 <= SMALL_SMEM_MAX_WARP ? 1 : 0;
}

int soap_launch_small(soap_chunk* c, const DevCfg& cfg, const HaloArrays& ha, int tier, const uint32_t* list,
                      const unsigned int* n_list, unsigned int n_list_upper, uint32_t* overflow,
                      unsigned int* n_overflow, unsigned int* queue_cursor, Counters* ctr, cudaStream_t stream) {
    const bool full = (cfg.flags & (PF_KIN | PF_KAPPA | PF_TENS)) != 0;
    const int stride = bank_stride_of(cfg);
    const unsigned int cap = n_list_upper < 1 ? 1 : n_list_upper;
#define ARGS c, cfg, ha, list, n_list, overflow, n_overflow, queue_cursor, ctr, stride, cap, stream
#define GO(NCH, VV)                                                  \
    (tier == 0   ? launch_warp_tier<NCH, VV, 256>(ARGS)              \
     : tier == 1 ? launch_warp_tier<NCH, VV, 512>(ARGS)              \
                 : launch_tier<NCH, VV, 256, 2048>(ARGS))
    if (cfg.dmo) return full ? GO(2, V_FULL) : GO(2, V_MIN);
    return full ? GO(8, V_FULL) : GO(8, V_MIN);
#undef GO
#undef ARGS
}
