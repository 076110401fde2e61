// tier.cu -- staged path for halos whose search sphere fits in a warp's shared-memory slot.
//
// The ~99 % of halos that are small never enter the per-rung kernel sequence of the
// general path (halos.cu).  They run through four kernels per round, each of which keeps
// every lane busy:
//   k_tier_front    one WARP per halo: ladder rungs (sweep + count + density gate,
//                   halo_tasks.py:73-103,166-187), gather + halo-centred re-wrap
//                   (halo_tasks.py:106-117), radial sort (SO_properties.py:398) in shared
//                   memory; the sorted records and particle slots go to global memory
//   k_solve_seq     one THREAD per halo (seq.cuh): scans, SO / Vmax / half-mass solves,
//                   commit logic, shell cuts
//   k_tier_moments  one WARP per halo: moment banks of the committed selections from the
//                   halo's particle slots (moments.cuh)
//   k_rows          one THREAD per (halo, selection): the result row (moments.cu)
// (+ k_tier_kappa / k_kappa_finish for kappa_corot, DtoT and the stellar rotation).  A
// halo whose solve asks for a larger radius comes back in the next round with its ladder
// state; one whose sphere does not fit moves to the next tier / the general path.
#include "moments.cuh"
#include "scan.cuh"

int soap_bank_stride(const DevCfg& cfg);
int soap_launch_rows(soap_chunk* c, const DevCfg& cfg, const HaloArrays& ha, const uint32_t* list,
                     const unsigned int* n_list_dev, unsigned int n_list_host, cudaStream_t stream);
int soap_launch_solve_seq(soap_chunk* c, const DevCfg& cfg, const HaloArrays& ha, const uint32_t* list,
                          const unsigned int* n_list_dev, unsigned int n_list_host, const Rec* recs, uint32_t* next,
                          unsigned int* n_next, Counters* ctr, const unsigned long long* item_minr, const int32_t* item_minfof,
                          int multi, cudaStream_t stream);
int soap_launch_kappa_finish(soap_handle* h, const DevCfg& cfg, const HaloArrays& ha, const uint32_t* list,
                             const unsigned int* n_list_dev, unsigned int n_list_host, cudaStream_t stream);

namespace {

enum : int { ACT_TRY = 0, ACT_RETRY = 1, ACT_DONE = 2, ACT_OVERFLOW = 3 };

// one gathered particle -> its record (halo-centred re-wrap, halo_tasks.py:106-117); kept out of line:
// the unrolled sweep calls it from several sites and the kernel must stay small enough for the
// instruction cache
__device__ __noinline__ unsigned long long rel_radius_bits(double X, double Y, double Z, double cx, double cy,
                                                           double cz, double L, double halfL) {
    const double x = rewrap_rel(X, cx, L, halfL);
    const double y = rewrap_rel(Y, cy, L, halfL);
    const double z = rewrap_rel(Z, cz, L, halfL);
    return (unsigned long long)__double_as_longlong(radius3(x, y, z));
}

template <int CAP>
struct __align__(16) FrontSlot {
    Rec rec[CAP];
    uint32_t pid[CAP];
    uint32_t bin_off[CAP / 2];  // counting sort: radial bin histogram / running offsets
    uint16_t ord[CAP];          // sorted position -> slot in rec / pid
    DimRanges rg[3];
    uint32_t row_s0[32], row_off[33];
    unsigned int n_stage;
};

template <int NCH, int CAP, int NW>
__global__ void __launch_bounds__(32 * NW, CAP <= 256 ? 3 : 1) k_tier_front(ChunkView v, HaloArrays ha, DevCfg cfg,
                                                        const uint32_t* __restrict__ list,
                                                        const unsigned int* __restrict__ n_list,
                                                        uint32_t* __restrict__ overflow,
                                                        unsigned int* __restrict__ n_overflow,
                                                        unsigned int* __restrict__ queue_cursor,
                                                        uint32_t* __restrict__ try_list, Counters* ctr,
                                                        Rec* __restrict__ recs, uint32_t* __restrict__ pids,
                                                        unsigned long long* __restrict__ item_minr,
                                                        int32_t* __restrict__ item_minfof, int bank_stride, int multi,
                                                        unsigned int slot_cap, unsigned long long rec_capacity) {
    constexpr uint32_t CAND_MAX = 16u * CAP;  // larger sweeps belong to the CTA-wide kernels of the general path
    constexpr int U = 4;                      // candidate groups in flight per warp (memory-level parallelism)
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    FrontSlot<CAP>& W = reinterpret_cast<FrontSlot<CAP>*>(smem_raw)[wid];
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

    while (true) {
        unsigned int it = 0;
        if (lane == 0) it = atomicAdd(queue_cursor, 1u);
        it = __shfl_sync(0xffffffffu, it, 0);
        if (it >= n_total) break;
        const uint32_t h = list[it];
        const double cx = ha.cofp[3 * h], cy = ha.cofp[3 * h + 1], cz = ha.cofp[3 * h + 2];
        const int64_t hidx = ha.index[h];
        const bool central = ha.central[h] == 1;
        double cur = ha.cur_r[h], r2max = 0.0;
        // a list longer than the scratch provisioned for it (rounds are enqueued before their list sizes are known):
        // the surplus moves on like halos that do not fit the tier
        int nloop = ha.nloop[h], nrows = 0, action = it < slot_cap ? ACT_RETRY : ACT_OVERFLOW, n_rungs = 1;
        uint32_t n = 0, total = 0, n_gather = 0;
        double r2rung[4] = {-1.0, -1.0, -1.0, -1.0};  // squared radii of the gathered rungs, accepted rung first
        bool look1 = false;  // the look-ahead sphere did not fit: sweep the current rung alone
        // ---------------------------------------------------------------- ladder rungs
        while (action == ACT_RETRY) {
            // rows of the furthest of the next rungs (ladder look-ahead: one sweep bins the sphere by
            // rung, like k_count; a sphere too large for this tier is retried rung by rung)
            constexpr int LOOK = 4;
            double rr[LOOK];
            const int nr = ladder_radii(cur, ha.rr_in[h], look1 ? 1 : LOOK, rr);
            __syncwarp();
            if (lane < 3) halo_ranges(v, cx, cy, cz, rr[nr - 1], W.rg, lane);
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
                    const uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
                    if (lane >= o) incl += t;
                }
                W.row_s0[lane] = s0;
                W.row_off[lane] = incl - len;
                total = __shfl_sync(0xffffffffu, incl, 31);
                if (lane == 31) W.row_off[32] = total;
                too_big = total > CAND_MAX;
                __syncwarp();
            }
            if (too_big) {
                if (nr > 1) { look1 = true; continue; }  // try again with this rung alone
                action = ACT_OVERFLOW;
                break;
            }
            look1 = false;
            // count + enclosed mass per rung (halo_tasks.py:84-97)
            double r2k[LOOK];
#pragma unroll
            for (int k = 0; k < LOOK; k++) r2k[k] = k < nr ? __dmul_rn(rr[k], rr[k]) : -1.0;
            uint32_t cnt[LOOK];
            double msum[LOOK];
#pragma unroll
            for (int k = 0; k < LOOK; k++) { cnt[k] = 0; msum[k] = 0.0; }
            for (uint32_t j0 = lane; j0 < total; j0 += 32 * U) {
                bool ok[U];
                double X[U], Y[U], Z[U];
                float M[U];
#pragma unroll
                for (int u = 0; u < U; u++) {
                    const uint32_t j = j0 + 32 * u;
                    ok[u] = j < total;
                    const uint32_t t = cand_slot(ok[u] ? j : 0u, nrows);
                    X[u] = v.px[t]; Y[u] = v.py[t]; Z[u] = v.pz[t]; M[u] = v.mass[t];
                }
#pragma unroll
                for (int u = 0; u < U; u++) {
                    const double r2 = periodic_r2(X[u], Y[u], Z[u], cx, cy, cz, L, halfL);
                    if (ok[u] && r2 <= r2k[nr - 1]) {
                        const double m = (double)M[u];
                        bool placed = false;
#pragma unroll
                        for (int k = 0; k < LOOK; k++)
                            if (!placed && k < nr && r2 <= r2k[k]) { cnt[k]++; msum[k] += m; placed = true; }
                    }
                }
            }
#pragma unroll
            for (int k = 0; k < LOOK; k++) {
                cnt[k] = (uint32_t)warp_sum_u64(cnt[k]);
                msum[k] = warp_sum(msum[k]);
            }
            // density gate and ladder steps (halo_tasks.py:73-103,166-187)
            uint32_t ccum = 0;
            int kacc = 0;
            n_rungs = 1;
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
                            ha.cnt[h] = ccum;
                            ha.msum[h] = mcum;
                            ha.rung_r[h] = r;
                            ha.commit_lo[h] = ha.commit_hi[h] = ha.ndone[h];
                            ha.state[h] = ST_TRY;
                            // the rungs beyond the accepted one that this sweep covered travel with the records, so
                            // that a solve which needs a larger radius can try them at once (seq.cuh)
                            uint32_t call = ccum;
                            double mall = mcum;
                            ha.rung_cnt[(size_t)h * LOOK_MAX] = call;
                            ha.rung_msum[(size_t)h * LOOK_MAX] = mall;
                            int nrg = 1;
                            if (multi) {
                                uint32_t nall = ccum;
                                for (int kk = k + 1; kk < nr; kk++) nall += cnt[kk];
                                if (nall <= (uint32_t)CAP)
                                    for (int kk = k + 1; kk < nr; kk++) {
                                        call += cnt[kk];
                                        mall += msum[kk];
                                        ha.rung_cnt[(size_t)h * LOOK_MAX + nrg] = call;
                                        ha.rung_msum[(size_t)h * LOOK_MAX + nrg] = mall;
                                        nrg++;
                                    }
                            }
                            ha.look[h] = nrg;
                            n_rungs = nrg;
                            n_gather = call;
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
            n_rungs = __shfl_sync(0xffffffffu, n_rungs, 0);
            n = __shfl_sync(0xffffffffu, n_gather, 0);
            kacc = __shfl_sync(0xffffffffu, kacc, 0);
            if (action == ACT_TRY) {
                // radius of the sphere that is gathered, squared radii of its rungs (accepted rung first)
#pragma unroll
                for (int k = 0; k < LOOK; k++) {
                    r2rung[k] = -1.0;
#pragma unroll
                    for (int kk = 0; kk < LOOK; kk++) {
                        if (kk == kacc + k) r2rung[k] = r2k[kk];
                        if (kk == kacc + n_rungs - 1) { cur = rr[kk]; r2max = r2k[kk]; }
                    }
                }
            } else if (action == ACT_RETRY) {
                cur = __shfl_sync(0xffffffffu, lane == 0 ? ha.cur_r[h] : 0.0, 0);
            }
        }
        if (lane == 0) ha.nloop[h] = nloop;
        if (action != ACT_TRY) {
            if (lane == 0 && action == ACT_OVERFLOW) {
                ha.state[h] = ST_PENDING;
                overflow[atomicAdd(n_overflow, 1u)] = h;
            }
            continue;
        }
        // ---------------------------------------------- gather + re-wrap (halo_tasks.py:106-117)
        if (lane == 0) W.n_stage = 0;
        __syncwarp();
        unsigned long long minr = ~0ull;
        int32_t minfof = -1;
        for (uint32_t j0 = 0; j0 < total; j0 += 32 * U) {
            bool in[U];
            uint32_t T[U];
            double X[U], Y[U], Z[U];
#pragma unroll
            for (int u = 0; u < U; u++) {
                const uint32_t j = j0 + 32 * u + lane;
                in[u] = j < total;
                T[u] = cand_slot(in[u] ? j : 0u, nrows);
                X[u] = v.px[T[u]]; Y[u] = v.py[T[u]]; Z[u] = v.pz[T[u]];
            }
#pragma unroll
            for (int u = 0; u < U; u++) {
                if (j0 + 32 * u >= total) break;  // warp-uniform
                const uint32_t t = T[u];
                const double r2 = periodic_r2(X[u], Y[u], Z[u], cx, cy, cz, L, halfL);
                const bool isin = in[u] && r2 <= r2max;
                const unsigned bal = __ballot_sync(0xffffffffu, isin);
                const unsigned base = W.n_stage;
                __syncwarp();
                if (lane == 0) W.n_stage = base + __popc(bal);
                if (isin) {
                    const uint32_t slot = base + __popc(bal & ((1u << lane) - 1u));
                    Rec rc;
                    rc.rbits = rel_radius_bits(X[u], Y[u], Z[u], cx, cy, cz, L, halfL);
                    rc.m = v.mass[t];
                    const uint32_t tc = NCH == 2 ? 1u : (uint32_t)v.type[t];
                    // first gathered rung whose periodic r2 test includes the particle (0 = the accepted rung)
                    uint32_t rung = 0;
#pragma unroll
                    for (int k = 0; k < 3; k++) rung += (k + 1 < n_rungs && !(r2 <= r2rung[k])) ? 1u : 0u;
                    rc.flags = tc | ((v.grnr[t] == hidx) ? 4u : 0u) | (rung << 4);
                    SOAP_ASSERT(slot < (uint32_t)CAP && slot < n && t < (uint32_t)v.n && rung < 4u);
                    W.rec[slot] = rc;
                    W.pid[slot] = t;
                    const int32_t f = (int32_t)v.fof[t];
                    if (rc.rbits < minr || (rc.rbits == minr && f < minfof)) { minr = rc.rbits; minfof = f; }
                }
                __syncwarp();
            }
        }
        // fofid of the innermost particle (SO_properties.py:407-409)
        for (int o = 16; o > 0; o >>= 1) {
            const unsigned long long orr = __shfl_xor_sync(0xffffffffu, minr, o);
            const int32_t of = __shfl_xor_sync(0xffffffffu, minfof, o);
            if (orr < minr || (orr == minr && of < minfof)) { minr = orr; minfof = of; }
        }
        // ---------------------------------------------- radial sort by (radius bits, particle slot)
        // Counting sort on nb radial bins of the sphere (monotone in r, so the bins are ordered),
        // then every lane insertion-sorts whole bins; W.ord maps sorted position -> slot.  A sphere
        // whose records crowd into one bin falls back to the bitonic network.
        bool counted = false;
        if (n > 32) {
            uint32_t nb = 32;
            while (nb < n / 2) nb <<= 1;  // <= CAP / 2, ~2-4 records per bin
            const double scale = (double)nb / cur;
            auto bin_of = [&](unsigned long long rbits) -> uint32_t {
                const uint32_t b = (uint32_t)(__longlong_as_double((long long)rbits) * scale);
                return b < nb ? b : nb - 1;
            };
            for (uint32_t b = lane; b < nb; b += 32) W.bin_off[b] = 0;
            __syncwarp();
            for (uint32_t i = lane; i < n; i += 32) atomicAdd(&W.bin_off[bin_of(W.rec[i].rbits)], 1u);
            __syncwarp();
            // exclusive scan: lane l owns bins [l * per, (l + 1) * per)
            const uint32_t per = nb / 32;
            uint32_t loc = 0, mx = 0;
            for (uint32_t k = 0; k < per; k++) {
                const uint32_t c = W.bin_off[lane * per + k];
                loc += c;
                mx = c > mx ? c : mx;
            }
            uint32_t incl = loc;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
                if (lane >= o) incl += t;
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                const uint32_t t = __shfl_xor_sync(0xffffffffu, mx, o);
                mx = t > mx ? t : mx;
            }
            if (mx <= 24) {
                counted = true;
                uint32_t run = incl - loc;
                for (uint32_t k = 0; k < per; k++) {
                    const uint32_t c = W.bin_off[lane * per + k];
                    W.bin_off[lane * per + k] = run;
                    run += c;
                }
                __syncwarp();
                for (uint32_t i = lane; i < n; i += 32) {
                    const uint32_t pos = atomicAdd(&W.bin_off[bin_of(W.rec[i].rbits)], 1u);
                    SOAP_ASSERT(pos < n);
                    W.ord[pos] = (uint16_t)i;
                }
                __syncwarp();
                // bin b now spans [end of bin b - 1, bin_off[b])
                for (uint32_t b = lane; b < nb; b += 32) {
                    const uint32_t lo = b == 0 ? 0u : W.bin_off[b - 1], hi = W.bin_off[b];
                    for (uint32_t j = lo + 1; j < hi; j++) {
                        const uint16_t oj = W.ord[j];
                        const unsigned long long kj = W.rec[oj].rbits;
                        const uint32_t pj = W.pid[oj];
                        uint32_t q = j;
                        while (q > lo) {
                            const uint16_t oq = W.ord[q - 1];
                            const unsigned long long kq = W.rec[oq].rbits;
                            if (kq < kj || (kq == kj && W.pid[oq] < pj)) break;
                            W.ord[q] = oq;
                            q--;
                        }
                        W.ord[q] = oj;
                    }
                }
                __syncwarp();
            }
        }
        if (!counted) {
            if (n <= 32) {
                // one record per lane: rank = number of records in front of mine (a short rolled loop;
                // the unrolled shuffle network of the same job was a quarter of this kernel's code)
                unsigned long long key = ~0ull;
                uint32_t mp = 0xffffffffu;
                if (lane < (int)n) { key = W.rec[lane].rbits; mp = W.pid[lane]; }
                uint32_t rank = 0;
#pragma unroll 1
                for (uint32_t j = 0; j < n; j++) {
                    const unsigned long long kj = __shfl_sync(0xffffffffu, key, j);
                    const uint32_t pj = __shfl_sync(0xffffffffu, mp, j);
                    rank += (kj < key || (kj == key && pj < mp)) ? 1u : 0u;
                }
                if (lane < (int)n) W.ord[rank] = (uint16_t)lane;
                __syncwarp();
                counted = true;  // W.ord is set
            } else {
                uint32_t np2 = 64;
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
            if (!counted)
                for (uint32_t i = lane; i < n; i += 32) W.ord[i] = (uint16_t)i;
            __syncwarp();
        }
        // ---------------------------------------------- hand over to the solve / moment kernels
        unsigned long long off = 0;
        if (lane == 0) {
            off = atomicAdd(&ctr->rec_total, (unsigned long long)n);
            ha.rec_off[h] = off;
            ha.item_base[h] = it;
            ha.n_items[h] = 1;
            ha.bank_off[h] = (unsigned long long)it * (unsigned long long)bank_stride;
            item_minr[it] = minr;
            item_minfof[it] = minfof;
            try_list[atomicAdd(&ctr->n_try, 1u)] = h;
        }
        off = __shfl_sync(0xffffffffu, off, 0);
        SOAP_ASSERT(W.n_stage == n);  // the gather found exactly the particles the count sweep counted
        for (uint32_t i = lane; i < n; i += 32) {
            const uint32_t o = W.ord[i];
            SOAP_ASSERT(o < n && off + i < rec_capacity);
            recs[off + i] = W.rec[o];
            pids[off + i] = W.pid[o];
        }
        __syncwarp();
    }
}

// ------------------------------------------------------------------ moments
template <int V>
struct __align__(16) MomSlot {
    double stage[32 * BankAcc<V>::VP];
    int skey[32];
    Cuts cuts;
    // followed by the warp's banks when they live in shared memory
};

template <int V, int NTY, int NW>
__global__ void __launch_bounds__(32 * NW) k_tier_moments(ChunkView v, HaloArrays ha, DevCfg cfg,
                                                          const uint32_t* __restrict__ list,
                                                          const unsigned int* __restrict__ n_list,
                                                          const uint32_t* __restrict__ pids, int slot_bytes,
                                                          int smem_banks) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    MomSlot<V>& W = *reinterpret_cast<MomSlot<V>*>(smem_raw + (size_t)wid * slot_bytes);
    double* sbank = reinterpret_cast<double*>(smem_raw + (size_t)wid * slot_bytes + sizeof(MomSlot<V>));
    const double L = v.L, halfL = 0.5 * v.L;
    const unsigned int n_total = *n_list;
    for (unsigned int it = blockIdx.x * NW + wid; it < n_total; it += gridDim.x * NW) {
        const uint32_t h = list[it];
        const int c_lo = ha.commit_lo[h], c_hi = ha.commit_hi[h];
        if (c_hi <= c_lo || ha.status[h] >= 2) continue;
        const double cx = ha.cofp[3 * h], cy = ha.cofp[3 * h + 1], cz = ha.cofp[3 * h + 2];
        const int64_t hidx = ha.index[h];
        const uint32_t n = ha.cnt[h];
        const uint32_t* pid = pids + ha.rec_off[h];
        __syncwarp();
        {
            const int* src = reinterpret_cast<const int*>(ha.cuts + h);
            int* dst = reinterpret_cast<int*>(&W.cuts);
            for (int i = lane; i < (int)(sizeof(Cuts) / sizeof(int)); i += 32) dst[i] = src[i];
        }
        __syncwarp();
        const int ncut = W.cuts.n;
        const int nbank = (ncut + 1) * 2 * NTY;
        double* gb = ha.gbank + ha.bank_off[h];
        double* banks = smem_banks ? sbank : gb;
        constexpr int NCOPY = BankAcc<V>::HALF ? 2 : 1;  // V <= 16: one copy per half-warp (always in shared memory)
        const int hstride = nbank * V;
        for (int i = lane; i < NCOPY * nbank * V; i += 32) banks[i] = 0.0;
        __syncwarp();
        const int32_t cen_fof = ha.sres[h].cen_fof;
        BankAcc<V> ba;
        ba.init(hstride);
        for (uint32_t b0 = 0; b0 < n; b0 += 32) {
            const uint32_t i = b0 + lane;
            const bool in = i < n;
            int key = 0;
            double val[V];
            if (in) {
                const uint32_t t = pid[i];
                SOAP_ASSERT(t < (uint32_t)v.n);
                const double x = rewrap_rel(v.px[t], cx, L, halfL);
                const double y = rewrap_rel(v.py[t], cy, L, halfL);
                const double z = rewrap_rel(v.pz[t], cz, L, halfL);
                const double r = radius3(x, y, z);
                key = moment_terms<V, NTY>(W.cuts, ncut, cfg, x, y, z, r, (double)v.mass[t], (double)v.vx[t],
                                           (double)v.vy[t], (double)v.vz[t], v.grnr[t], hidx, v.fof[t], cen_fof,
                                           NTY == 1 ? 1u : (uint32_t)v.type[t], val);
                SOAP_ASSERT(key >= 0 && key < nbank);
            }
            ba.add(in, key, val, W.stage, W.skey, banks, 1, lane);
        }
        ba.flush(banks, 1, lane);
        __syncwarp();
        if (smem_banks)
            for (int i = lane; i < nbank * V; i += 32) gb[i] = NCOPY == 2 ? sbank[i] + sbank[hstride + i] : sbank[i];
        __syncwarp();
    }
}

// kappa_corot / DtoT / stellar rotation need the finished vcom and L of the row: a second pass
// over the gas and star records of the halo (kinematic_properties.py:266-425)
template <int NW>
__global__ void __launch_bounds__(32 * NW) k_tier_kappa(ChunkView v, HaloArrays ha, DevCfg cfg,
                                                        const uint32_t* __restrict__ list,
                                                        const unsigned int* __restrict__ n_list,
                                                        const Rec* __restrict__ recs,
                                                        const uint32_t* __restrict__ pids) {
    constexpr int KS = KAPPA_MAX_SEL;
    extern __shared__ __align__(16) unsigned char kappa_smem[];  // [NW] x (KapSel[KS], double[KS][11])
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    KapSel* ksel = reinterpret_cast<KapSel*>(kappa_smem + (size_t)wid * KS * (sizeof(KapSel) + 11 * sizeof(double)));
    double(*kacc)[11] = reinterpret_cast<double(*)[11]>(ksel + KS);
    const double L = v.L, halfL = 0.5 * v.L;
    const unsigned int n_total = *n_list;
    for (unsigned int it = blockIdx.x * NW + wid; it < n_total; it += gridDim.x * NW) {
        const uint32_t h = list[it];
        const int c_lo = ha.commit_lo[h], c_hi = ha.commit_hi[h];
        if (c_hi <= c_lo || ha.status[h] >= 2) continue;
        const double cx = ha.cofp[3 * h], cy = ha.cofp[3 * h + 1], cz = ha.cofp[3 * h + 2];
        const uint32_t n = ha.cnt[h];
        const Rec* R = recs + ha.rec_off[h];
        const uint32_t* pid = pids + ha.rec_off[h];
        __syncwarp();
        int ns = 0;
        if (lane == 0) ns = kappa_build_sels(ksel, cfg, ha, h, c_lo, c_hi);
        ns = __shfl_sync(0xffffffffu, ns, 0);
        if (ns == 0) continue;
        for (int i = lane; i < ns * 11; i += 32) (&kacc[0][0])[i] = 0.0;
        __syncwarp();
        for (uint32_t i = lane; i < n; i += 32) {
            const Rec rc = R[i];
            const uint32_t tc = rc.flags & 3u;
            if (tc != 0u && tc != 2u) continue;
            const uint32_t t = pid[i];
            const double x = rewrap_rel(v.px[t], cx, L, halfL);
            const double y = rewrap_rel(v.py[t], cy, L, halfL);
            const double z = rewrap_rel(v.pz[t], cz, L, halfL);
            kappa_add(ksel, ns, kacc, x, y, z, radius3(x, y, z), (double)v.mass[t], (double)v.vx[t], (double)v.vy[t],
                      (double)v.vz[t], tc, (rc.flags & 4u) != 0u);
        }
        __syncwarp();
        for (int i = lane; i < ns * 11; i += 32) {
            const double a = kacc[i / 11][i % 11];
            if (a != 0.0) *ksel[i / 11].slot(i % 11) += a;
        }
        __syncwarp();
    }
}

template <int NCH, int CAP, int NW>
int launch_front(soap_chunk* c, const DevCfg& cfg, const HaloArrays& ha, const uint32_t* list,
                 const unsigned int* n_list, unsigned int n_upper, uint32_t* overflow, unsigned int* n_overflow,
                 unsigned int* queue_cursor, uint32_t* try_list, Counters* ctr, Rec* recs, uint32_t* pids,
                 unsigned long long* item_minr, int32_t* item_minfof, int bank_stride, int multi, cudaStream_t stream) {
    soap_handle* h = c->h;
    auto kern = k_tier_front<NCH, CAP, NW>;
    const size_t smem = sizeof(FrontSlot<CAP>) * NW;
    CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int per_sm = 0;
    CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, 32 * NW, smem));
    if (per_sm < 1) SOAP_FAIL("soap_process_halos: tier front kernel does not fit on an SM (%zu bytes of shared memory)", smem);
    unsigned int grid = (unsigned int)(h->sm_count * per_sm);
    const unsigned int need = (n_upper + NW - 1) / NW;
    if (grid > need) grid = need < 1 ? 1 : need;
    LAUNCH_N(h, CAP <= 256 ? "k_tier_front<256>" : "k_tier_front<1024>", kern, grid, 32 * NW, smem, stream, c->v, ha, cfg,
             list, n_list, overflow, n_overflow, queue_cursor, try_list, ctr, recs, pids, item_minr, item_minfof,
             bank_stride, multi, n_upper, (unsigned long long)n_upper * CAP);
    return 0;
}

template <int V, int NTY>
int launch_tier_moments(soap_chunk* c, const DevCfg& cfg, const HaloArrays& ha, const uint32_t* list,
                        const unsigned int* n_list, unsigned int n_upper, const uint32_t* pids, int bank_stride,
                        cudaStream_t stream) {
    soap_handle* h = c->h;
    constexpr int NW = 8;
    const size_t bank_bytes = (size_t)bank_stride * sizeof(double);
    // banks in shared memory while a CTA of 8 warps stays below ~48 KB; else straight in global memory
    // (touched only on a key change).  The 15-term set keeps one copy per half-warp, always in shared memory.
    const int ncopy = V <= 16 ? 2 : 1;
    const int smem_banks = (V <= 16 || bank_bytes <= 4096) ? 1 : 0;
    const size_t slot = (sizeof(MomSlot<V>) + (smem_banks ? ncopy * bank_bytes : 0) + 15) & ~(size_t)15;
    const size_t smem = slot * NW;
    auto kern = k_tier_moments<V, NTY, NW>;
    CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int per_sm = 0;
    CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, 32 * NW, smem));
    if (per_sm < 1) SOAP_FAIL("soap_process_halos: tier moment kernel does not fit on an SM (%zu bytes of shared memory)", smem);
    unsigned int grid = (unsigned int)(h->sm_count * per_sm);
    const unsigned int need = (n_upper + NW - 1) / NW;
    if (grid > need) grid = need < 1 ? 1 : need;
    LAUNCH_N(h, "k_tier_moments", kern, grid, 32 * NW, smem, stream, c->v, ha, cfg, list, n_list, pids, (int)slot,
             smem_banks);
    return 0;
}

}  // namespace

int soap_tier_cap(int tier) { return tier == 0 ? 256 : 1024; }

// One round of tier `tier` over `list`: front -> solve -> moments -> rows (-> kappa).  Halos that need
// a larger radius are appended to `next` (*n_next), those that do not fit to `overflow`.
// try_list receives the halos handed to the solve (ctr->n_try).  ctr must be zeroed by the caller.
int soap_tier_round(soap_chunk* c, const DevCfg& cfg, HaloArrays& ha, int tier, const uint32_t* list,
                    const unsigned int* n_list, unsigned int n_upper, uint32_t* overflow, unsigned int* n_overflow,
                    unsigned int* queue_cursor, uint32_t* try_list, uint32_t* next, unsigned int* n_next, Counters* ctr,
                    unsigned long long* item_minr, int32_t* item_minfof, cudaStream_t stream) {
    soap_handle* h = c->h;
    if (n_upper == 0) return 0;
    const int cap = soap_tier_cap(tier);
    const int bank_stride = soap_bank_stride(cfg);
    const bool full = (cfg.flags & (PF_KIN | PF_KAPPA | PF_TENS)) != 0;
    // several ladder rungs per solve only where no selection depends on the sphere it was computed in: an inclusive
    // sphere's stellar tensors count every star loaded (aperture_properties.py:3579-3594), the iterative tensors
    // use the sphere of the rung that committed them
    const int multi = (cfg.n_ap == 0 && cfg.n_pj == 0 && !(cfg.flags & PF_ITER)) ? 1 : 0;
    // scratch per tier: the two tiers run concurrently (halos.cu)
    const char* nm[2][3] = {{"h_trecs0", "h_tpids0", "h_tbank0"}, {"h_trecs1", "h_tpids1", "h_tbank1"}};
    Rec* recs = (Rec*)h->get(nm[tier != 0][0], sizeof(Rec) * (size_t)n_upper * cap);
    uint32_t* pids = (uint32_t*)h->get(nm[tier != 0][1], sizeof(uint32_t) * (size_t)n_upper * cap);
    ha.gbank = (double*)h->get(nm[tier != 0][2], sizeof(double) * (size_t)bank_stride * ((size_t)n_upper + 1));
    if (!recs || !pids || !ha.gbank) return -1;
#define FRONT(NCH, CAP, NW)                                                                                            \
    launch_front<NCH, CAP, NW>(c, cfg, ha, list, n_list, n_upper, overflow, n_overflow, queue_cursor, try_list, ctr, recs, \
                               pids, item_minr, item_minfof, bank_stride, multi, stream)
    int rc;
    if (cfg.dmo) rc = tier == 0 ? FRONT(2, 256, 8) : FRONT(2, 1024, 8);
    else rc = tier == 0 ? FRONT(8, 256, 8) : FRONT(8, 1024, 8);
#undef FRONT
    if (rc) return -1;
    if (soap_launch_solve_seq(c, cfg, ha, try_list, &ctr->n_try, n_upper, recs, next, n_next, ctr, item_minr, item_minfof, multi, stream))
        return -1;
#define MOMS(V, NTY) launch_tier_moments<V, NTY>(c, cfg, ha, try_list, &ctr->n_try, n_upper, pids, bank_stride, stream)
    if (full && !cfg.dmo) rc = MOMS(V_FULL, 4);
    else if (full) rc = MOMS(V_FULL, 1);
    else if (!cfg.dmo) rc = MOMS(V_MIN, 4);
    else rc = MOMS(V_MIN, 1);
#undef MOMS
    if (rc) return -1;
    if (soap_launch_rows(c, cfg, ha, try_list, &ctr->n_try, n_upper, stream)) return -1;
    if ((cfg.flags & PF_KAPPA) && !cfg.dmo) {
        if (!(cfg.flags & PF_KIN)) SOAP_FAIL("soap_process_halos: kappa_corot needs the kinematics group (property_flags bit 0)");
        constexpr int NW = 8;
        unsigned int grid = (unsigned int)(h->sm_count * 4);
        const unsigned int need = (n_upper + NW - 1) / NW;
        if (grid > need) grid = need < 1 ? 1 : need;
        const size_t ksm = (size_t)NW * KAPPA_MAX_SEL * (sizeof(KapSel) + 11 * sizeof(double));
        CUDA_TRY(cudaFuncSetAttribute(k_tier_kappa<NW>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ksm));
        LAUNCH(h, k_tier_kappa<NW>, grid, 32 * NW, ksm, stream, c->v, ha, cfg, try_list, &ctr->n_try, recs, pids);
        if (soap_launch_kappa_finish(h, cfg, ha, try_list, &ctr->n_try, n_upper, stream)) return -1;
    }
    return 0;
}
