// halos.cu -- device-side halo batching: replaces the per-core task loop of
// process_halos / process_single_halo (SOAP/core/halo_tasks.py:23-430).
//
// One "round" = one rung of the search-radius ladder for every halo still
// pending:
//   k_count        periodic sphere count + enclosed mass (halo_tasks.py:84-97;
//                  shared_mesh.py:122-200 on the chunk's cell-ordered SoA)
//   k_gate         density gate + ladder step (halo_tasks.py:103,166-187)
//   k_fine_hist /
//   k_build_buckets  radial bucket plan for halos too large for one CTA sort
//   k_collect      gather (halo_tasks.py:106-117 re-wrap) -> 16-byte records
//   k_sort_bucket  segmented radial sort (np.argsort of SO_properties.py:398,
//                  half_mass_radius.py:45, kinematic_properties.py:581)
//   k_scan_solve   segmented scans over the sorted profile: SO radius/mass
//                  (SO_properties.py:80-217,356-513), Vmax
//                  (kinematic_properties.py:555-593), half-mass radii
//                  (half_mass_radius.py:16-97); decides retry vs final
//   k_moments      (moments.cu) masked moment sums + result row
#include "moments.cuh"
#include "scan.cuh"

#include <stdlib.h>

int soap_launch_moments(soap_chunk* c, const DevCfg& cfg, const HaloArrays& ha, const Item* items,
                        const unsigned int* n_items_dev, unsigned int n_items_host, unsigned int grid,
                        cudaStream_t stream);
int soap_bank_stride(const DevCfg& cfg);
int soap_launch_rows(soap_chunk* c, const DevCfg& cfg, const HaloArrays& ha, const uint32_t* list,
                     const unsigned int* n_list_dev, unsigned int n_list_host, cudaStream_t stream);
int soap_launch_solve_seq(soap_chunk* c, const DevCfg& cfg, const HaloArrays& ha, const uint32_t* list,
                          const unsigned int* n_list_dev, unsigned int n_list_host, const Rec* recs, uint32_t* next,
                          unsigned int* n_next, Counters* ctr, const unsigned long long* item_minr, const int32_t* item_minfof,
                          int multi, cudaStream_t stream);
int soap_write_input_cols(soap_handle* h, const HaloArrays& ha, int64_t nh, cudaStream_t stream);
int soap_launch_kappa(soap_chunk* c, const DevCfg& cfg, const HaloArrays& ha, const Item* items,
                      const unsigned int* n_items_dev, unsigned int n_items_host, const uint32_t* acc_list,
                      const unsigned int* n_acc_dev, unsigned int n_acc_host, unsigned int grid,
                      cudaStream_t stream);
int soap_launch_projected(soap_chunk* c, const DevCfg& cfg, const HaloArrays& ha, const Item* items,
                          const unsigned int* n_items_dev, unsigned int n_items_host, unsigned int n_mslot,
                          unsigned int grid, cudaStream_t stream);
int soap_iter_list(soap_handle* h, const HaloArrays& ha, int64_t nh, uint32_t* list, unsigned int* n_list_dev,
                   unsigned int* n_host, cudaStream_t stream);
int soap_launch_iter_tensors(soap_chunk* c, const DevCfg& cfg, const HaloArrays& ha, int64_t nh, const Item* items,
                             const unsigned int* n_items_dev, unsigned int n_items_host, const uint32_t* list,
                             const unsigned int* n_list_dev, unsigned int n_list_host, unsigned int grid,
                             cudaStream_t stream);
int soap_tier_round(soap_chunk* c, const DevCfg& cfg, HaloArrays& ha, int tier, const uint32_t* list,
                    const unsigned int* n_list, unsigned int n_upper, uint32_t* overflow, unsigned int* n_overflow,
                    unsigned int* queue_cursor, uint32_t* try_list, uint32_t* next, unsigned int* n_next, Counters* ctr,
                    unsigned long long* item_minr, int32_t* item_minfof, cudaStream_t stream);

namespace {

constexpr int TB = SWEEP_NT;
constexpr int SB_CAP = 4096;     // records of one bucket sorted in shared memory (64 KB)
constexpr int SMALL_CAP = 512;   // small-bucket class (8 KB)
constexpr int FINE_TARGET = 12;  // expected records per fine radial bin (sorted by one warp in registers)
constexpr uint32_t SCAN_BIG = 32768;  // halos with more records are scanned by a CTA cluster
constexpr uint32_t SEQ_MAX = 512;     // halos with up to this many records are scanned by one thread each (seq.cuh)
constexpr int SCAN_CS = 8;            // CTAs per cluster for those
constexpr uint32_t SCAN_HUGE = 524288;  // above this a cluster of 16 CTAs (non-portable size, B200 supports it)
constexpr int SCAN_CS_HUGE = 16;
// halos with up to this many bound particles start in tier 0 / 1 of the staged small-halo path (tier.cu:
// spheres of up to 256 / 1024 particles); larger ones take the general path
constexpr long long SMALL_NEXP_0 = 150, SMALL_NEXP_1 = 500;
constexpr int TIER_ROUNDS = 5;
constexpr unsigned int SEQ_MIN_LIST = 2048;  // general path: shorter lists of small halos are scanned CTA-wise

struct Bucket {
    unsigned long long start;
    uint32_t count;
    uint32_t halo;
};


// ------------------------------------------------------------------ k_init
// Also deals the halos to the tiers by their expected size: fused tier t
// (small.cu) takes halos with up to lim[t] bound particles, the rest goes to
// the general path.  tier_n = device counters of the four lists.
struct TierLims { long long lim[3]; };
constexpr int TIER_BUCKETS = 1024;  // one bucket per nr_bound_part value handled by a fused tier
__global__ void k_init(HaloArrays ha, int64_t nh, uint32_t* pend, unsigned int* tier_n, unsigned int* size_hist,
                       TierLims tl) {
    int64_t h = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (h >= nh) return;
    ha.cur_r[h] = ha.sr_in[h];
    ha.nloop[h] = 0;
    ha.state[h] = ST_PENDING;
    ha.status[h] = SOAP_HALO_OK;
    ha.ndone[h] = 0;
    ha.commit_lo[h] = 0;
    ha.commit_hi[h] = 0;
    ha.mslot[h] = -1;
    // the fused tiers take their halos largest first and size-sorted (k_tier_place): the warps
    // of a lock-step CTA then work on halos of similar size
    const long long ne = ha.nexp[h];
    if (ne <= tl.lim[2]) atomicAdd(&size_hist[TIER_BUCKETS - 1 - (ne < 0 ? 0 : (int)ne)], 1u);
    else pend[atomicAdd(&tier_n[3], 1u)] = (uint32_t)h;
    double* row = ha.out + h * ha.ncol;
    for (int64_t c = 0; c < ha.ncol; c++) row[c] = 0.0;
}

// exclusive scan of the size histogram (descending size) + the tier list sizes
__global__ void k_tier_scan(unsigned int* size_hist, unsigned int* tier_n, TierLims tl) {
    __shared__ unsigned int sh[TIER_BUCKETS];
    const int t = threadIdx.x;
    const unsigned int v = size_hist[t];
    sh[t] = v;
    __syncthreads();
    for (int o = 1; o < TIER_BUCKETS; o <<= 1) {
        unsigned int x = t >= o ? sh[t - o] : 0u;
        __syncthreads();
        sh[t] += x;
        __syncthreads();
    }
    size_hist[t] = sh[t] - v;  // position of the first halo of this size among all fused-tier halos
    if (t < 3) {
        // bucket b holds nr_bound_part = TIER_BUCKETS - 1 - b; tier k takes lim[k-1] < ne <= lim[k]
        const long long hi = tl.lim[t], lo = t == 0 ? -1 : tl.lim[t - 1];
        unsigned int cnt = 0;
        if (hi > lo) {
            const int b0 = TIER_BUCKETS - 1 - (int)hi, b1 = TIER_BUCKETS - 1 - (int)(lo + 1);  // inclusive bucket range
            cnt = sh[b1] - (b0 > 0 ? sh[b0 - 1] : 0u);
        }
        tier_n[t] = cnt;
    }
}

__global__ void k_tier_place(HaloArrays ha, int64_t nh, uint32_t* list0, uint32_t* list1, uint32_t* list2,
                             unsigned int* size_cursor, const unsigned int* __restrict__ size_hist, TierLims tl) {
    int64_t h = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (h >= nh) return;
    const long long ne = ha.nexp[h];
    if (ne > tl.lim[2]) return;
    const int b = TIER_BUCKETS - 1 - (ne < 0 ? 0 : (int)ne);
    const unsigned int pos = size_hist[b] + atomicAdd(&size_cursor[b], 1u);  // among all tier halos, largest first
    // a tier's list starts where its largest size starts
    if (ne <= tl.lim[0]) list0[pos - size_hist[TIER_BUCKETS - 1 - (int)tl.lim[0]]] = (uint32_t)h;
    else if (ne <= tl.lim[1]) list1[pos - size_hist[TIER_BUCKETS - 1 - (int)tl.lim[1]]] = (uint32_t)h;
    else list2[pos - size_hist[TIER_BUCKETS - 1 - (int)tl.lim[2]]] = (uint32_t)h;
}

// ------------------------------------------------------------- k_plan_items
// One WARP per pending halo: count the candidates of its rows at the current radius and cut the
// candidate stream into work items of ITEM_CAND particles (lanes stride over the rows: the spheres
// of ladder stragglers have tens of thousands of them).
constexpr int PLAN_NT = 128;
__global__ void __launch_bounds__(PLAN_NT) k_plan_items(ChunkView v, HaloArrays ha, const uint32_t* __restrict__ pend,
                                                        const unsigned int* __restrict__ n_pend,
                                                        Item* __restrict__ items, unsigned int items_cap,
                                                        Counters* ctr, int look, int replan, int bank_stride) {
    __shared__ DimRanges rg_s[PLAN_NT / 32][3];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const unsigned int it = blockIdx.x * (PLAN_NT / 32) + wid;
    if (it >= *n_pend) return;
    const uint32_t h = pend[it];
    DimRanges* rg = rg_s[wid];
    // round start: plan the sweep of the furthest of the next `look` ladder rungs;
    // replan: the sweep of the accepted rung (collect / moments)
    double r = ha.cur_r[h];
    if (!replan) {
        // a halo that has not failed a rung yet looks 4 rungs ahead even when the round's stragglers look further:
        // the sphere of the 12th rung holds ~700 times the volume
        double rr[LOOK_MAX];
        const int nr = ladder_radii(r, ha.rr_in[h], (look > 4 && ha.nloop[h] == 0) ? 4 : look, rr);
        r = rr[nr - 1];
        if (lane == 0) ha.look[h] = nr;
        if (lane < LOOK_MAX) { ha.rung_cnt[(size_t)h * LOOK_MAX + lane] = 0; ha.rung_msum[(size_t)h * LOOK_MAX + lane] = 0.0; }
    }
    if (lane < 3) dim_ranges(ha.cofp[3 * h + lane], r, v.L, v.pmin[lane], v.pmax[lane], v.cs[lane], v.res, rg[lane]);
    __syncwarp();
    const RowIter ri = row_iter(rg);
    unsigned long long cand = 0;
    for (int row = lane; row < ri.nrows; row += 32) {
        uint32_t s0, s1;
        row_span(v, rg, ri, row, s0, s1);
        cand += s1 - s0;
    }
    cand = warp_sum_u64(cand);
    if (cand > 0xfffffff0ull) cand = 0xfffffff0ull;
    uint32_t ni = (uint32_t)((cand + ITEM_CAND - 1) / ITEM_CAND);
    if (ni < 1) ni = 1;
    uint32_t base = 0;
    if (lane == 0) base = atomicAdd(&ctr->n_items, ni);
    base = __shfl_sync(0xffffffffu, base, 0);
    if (base + ni > items_cap) {
        // the host grows the item list to n_items and plans this rung again
        if (lane == 0) atomicExch(&ctr->items_overflow, 1u);
        return;
    }
    if (lane == 0) {
        ha.item_base[h] = base;
        ha.n_items[h] = ni;
        ha.cursor[h] = 0;
        ha.items_done[h] = 0;
        if (!replan) {
            ha.cnt[h] = 0;
            ha.msum[h] = 0.0;
            ha.commit_lo[h] = ha.commit_hi[h] = ha.ndone[h];
            atomicAdd(&ctr->candidates, cand);
        }
        ha.mslot[h] = ni > 1 ? (int32_t)atomicAdd(&ctr->n_mslot, 1u) : -1;
        if (replan) ha.bank_off[h] = (unsigned long long)atomicAdd(&ctr->n_bslot, 1u) * (unsigned long long)bank_stride;
        if (ni == 1) {
            Item im;
            im.halo = h; im.first = 0; im.count = (uint32_t)cand; im.row0 = 0;
            im.pos0 = 0; im.k = 0; im.pad0 = im.pad1 = 0;
            items[base] = im;
        }
    }
    if (ni > 1) {
        // second walk: item k starts in the row that holds candidate k * ITEM_CAND of the stream
        unsigned long long pos = 0;
        for (int row0 = 0; row0 < ri.nrows; row0 += 32) {
            const int row = row0 + lane;
            uint32_t s0 = 0, s1 = 0;
            if (row < ri.nrows) row_span(v, rg, ri, row, s0, s1);
            const unsigned long long len = s1 - s0;
            unsigned long long incl = len;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const unsigned long long t = __shfl_up_sync(0xffffffffu, incl, o);
                if (lane >= o) incl += t;
            }
            const unsigned long long end = pos + incl, start = end - len;
            if (len > 0) {
                for (unsigned long long k = (start + ITEM_CAND - 1) / ITEM_CAND; k < ni && k * ITEM_CAND < end; k++) {
                    Item im;
                    im.halo = h;
                    im.first = (uint32_t)(k * ITEM_CAND);
                    const unsigned long long left = cand - k * ITEM_CAND;
                    im.count = (uint32_t)(left < ITEM_CAND ? left : ITEM_CAND);
                    im.row0 = (uint32_t)row;
                    im.pos0 = (uint32_t)start;
                    im.k = (uint32_t)k; im.pad0 = im.pad1 = 0;
                    items[base + k] = im;
                }
            }
            pos += __shfl_sync(0xffffffffu, incl, 31);
        }
    }
}

// ----------------------------------------------------------------- k_count
// periodic sphere count + enclosed mass of every pending halo (halo_tasks.py:84-97;
// shared_mesh.py:122-200), for the next ha.look[h] ladder rungs in one sweep:
// the sphere of the furthest rung is swept once and every particle is binned by
// the first rung whose radius includes it (the same r2 <= radius^2 test).
template <int NLOOK>
__global__ void __launch_bounds__(TB, 3) k_count(ChunkView v, HaloArrays ha, const Item* __restrict__ items,
                                              Counters* ctr, double* __restrict__ item_msum) {
    __shared__ SweepShared S;
    __shared__ unsigned int s_cnt[TB / 32][NLOOK];
    __shared__ double s_m[TB / 32][NLOOK];
    const unsigned int n_items = ctr->n_items;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    for (unsigned int it = blockIdx.x; it < n_items; it += gridDim.x) {
        const Item im = items[it];
        const uint32_t h = im.halo;
        const double cx = ha.cofp[3 * h], cy = ha.cofp[3 * h + 1], cz = ha.cofp[3 * h + 2];
        double rr[NLOOK], r2k[NLOOK];
        const int nr = ladder_radii(ha.cur_r[h], ha.rr_in[h], ha.look[h] < NLOOK ? ha.look[h] : NLOOK, rr);
#pragma unroll
        for (int k = 0; k < NLOOK; k++) r2k[k] = k < nr ? __dmul_rn(rr[k], rr[k]) : -1.0;
        const double r = rr[nr - 1];
        const double halfL = 0.5 * v.L, L = v.L;
        unsigned int cnt[NLOOK];
        double msum[NLOOK];
#pragma unroll
        for (int k = 0; k < NLOOK; k++) { cnt[k] = 0; msum[k] = 0.0; }
        sweep_item<SW_MASS>(v, S, cx, cy, cz, r, im, [&](uint32_t t, bool ok) {
            if (ok) {
                const double r2 = periodic_r2(v.px[t], v.py[t], v.pz[t], cx, cy, cz, L, halfL);
                if (r2 <= r2k[nr - 1]) {
                    const double m = (double)v.mass[t];
                    bool placed = false;
#pragma unroll
                    for (int k = 0; k < NLOOK; k++)
                        if (!placed && k < nr && r2 <= r2k[k]) { cnt[k]++; msum[k] += m; placed = true; }
                }
            }
        });
#pragma unroll
        for (int k = 0; k < NLOOK; k++) {
            const unsigned int c = (unsigned int)warp_sum_u64(cnt[k]);
            const double m = warp_sum(msum[k]);
            if (lane == 0) { s_cnt[wid][k] = c; s_m[wid][k] = m; }
        }
        __syncthreads();
        if (threadIdx.x < LOOK_MAX && threadIdx.x >= (unsigned)nr) item_msum[(size_t)it * LOOK_MAX + threadIdx.x] = 0.0;
        if (threadIdx.x < (unsigned)nr) {
            const int k = threadIdx.x;
            unsigned int c = 0;
            double m = 0.0;
            for (int w = 0; w < TB / 32; w++) { c += s_cnt[w][k]; m += s_m[w][k]; }
            if (c) atomicAdd(&ha.rung_cnt[(size_t)h * LOOK_MAX + k], c);
            // the mass of a rung is summed over the halo's work items in item order by k_gate: the same chunk
            // gives the same float64 sum, hence the same density gate decision, on every run
            item_msum[(size_t)it * LOOK_MAX + k] = m;
            atomicAdd(&ctr->count_pairs, (unsigned long long)c);
        }
        __syncthreads();
    }
}

__global__ void __launch_bounds__(128) k_gate(HaloArrays ha, DevCfg cfg, const uint32_t* __restrict__ pend,
                                              const unsigned int* __restrict__ n_pend,
                                              uint32_t* __restrict__ try_list,
                                              uint32_t* __restrict__ big_list,
                                              uint32_t* __restrict__ acc_list,
                                              uint32_t* __restrict__ multi_list,
                                              uint32_t* __restrict__ seq_list,
                                              uint32_t* __restrict__ huge_list,
                                              uint32_t* __restrict__ next, Counters* ctr,
                                              const double* __restrict__ item_msum) {
    unsigned int it = blockIdx.x * blockDim.x + threadIdx.x;
    if (it >= *n_pend) return;
    const uint32_t h = pend[it];
    // walk the rungs covered by this round's sweep (halo_tasks.py:73-103,166-187)
    const int nr = ha.look[h];
    const bool has_target = ha.central[h] == 1 && cfg.target_density > 0.0;  // halo_tasks.py:381
    unsigned int ccum = 0;
    double mcum = 0.0;
    bool accepted = false, pending = true;
    for (int k = 0; k < nr && pending; k++) {
        ha.nloop[h] += 1;  // halo_tasks.py:75
        const double r = ha.cur_r[h];
        ccum += ha.rung_cnt[(size_t)h * LOOK_MAX + k];
        {
            const uint32_t ib = ha.item_base[h], ni = ha.n_items[h];
            double mk = 0.0;
            for (uint32_t j = 0; j < ni; j++) mk += item_msum[(size_t)(ib + j) * LOOK_MAX + k];
            mcum += mk;
        }
        // halo_tasks.py:97
        const double density = mcum / (4.0 / 3.0 * SOAP_PI * (r * r * r));
        if (!has_target || density <= cfg.target_density) {  // halo_tasks.py:103
            accepted = true;
            break;
        }
        pending = ladder_step(ha, h, 0.0);  // next rung, or out of read radius
    }
    if (accepted) {
        ha.cnt[h] = ccum;
        ha.msum[h] = mcum;
        ha.rung_r[h] = ha.cur_r[h];
        const uint32_t cnt = ha.cnt[h];
        int cls;
        if (cnt > SCAN_HUGE) { huge_list[atomicAdd(&ctr->n_huge, 1u)] = h; cls = 3; }  // scanned by a cluster of 16 CTAs
        else if (cnt > SCAN_BIG) { big_list[atomicAdd(&ctr->n_big, 1u)] = h; cls = 2; }  // by a cluster of 8
        else if (cnt <= SEQ_MAX) { seq_list[atomicAdd(&ctr->n_seq, 1u)] = h; cls = 0; }  // by one thread
        else { try_list[atomicAdd(&ctr->n_try, 1u)] = h; cls = 1; }
        atomicAdd(&ctr->rec_class[cls], (unsigned long long)cnt);
        acc_list[atomicAdd(&ctr->n_acc, 1u)] = h;  // every accepted halo: its sweep is planned again
        ha.state[h] = ST_TRY;
        if (cnt <= SMALL_CAP) {
            ha.rec_off[h] = atomicAdd(&ctr->rec_single, (unsigned long long)cnt);
            ha.nfine[h] = 0;
        } else {
            uint32_t nf = cnt / FINE_TARGET;
            if (nf < 64) nf = 64;
            if (nf > (1u << 22)) nf = 1u << 22;
            ha.nfine[h] = nf;
            ha.fine_off[h] = atomicAdd(&ctr->n_fine, nf);
            multi_list[atomicAdd(&ctr->n_multi, 1u)] = h;
        }
        atomicAdd(&ctr->rec_total, (unsigned long long)cnt);
    } else if (pending) {
        next[atomicAdd(&ctr->n_next, 1u)] = h;
    }
}

__device__ __forceinline__ uint32_t fine_bin(double r, double R, uint32_t nf) {
    double f = r * ((double)nf / R);
    uint32_t b = f <= 0.0 ? 0u : (uint32_t)f;
    return b >= nf ? nf - 1 : b;
}

// -------------------------------------------------------------- fine hist
__global__ void __launch_bounds__(TB) k_fine_hist_halo(ChunkView v, HaloArrays ha, const Item* __restrict__ items,
                                                       const Counters* __restrict__ ctr,
                                                       uint32_t* __restrict__ fine_cnt) {
    __shared__ SweepShared S;
    const unsigned int n_items = ctr->n_items;
    for (unsigned int it = blockIdx.x; it < n_items; it += gridDim.x) {
        const Item im = items[it];
        const uint32_t h = im.halo;
        if (ha.state[h] != ST_TRY || ha.nfine[h] == 0) continue;
        const double cx = ha.cofp[3 * h], cy = ha.cofp[3 * h + 1], cz = ha.cofp[3 * h + 2];
        const double R = ha.cur_r[h];
        const double r2max = __dmul_rn(R, R), halfL = 0.5 * v.L, L = v.L;
        const uint32_t nf = ha.nfine[h];
        uint32_t* fc = fine_cnt + ha.fine_off[h];
        sweep_item(v, S, cx, cy, cz, R, im, [&](uint32_t t, bool ok) {
            if (ok) {
                double r2 = periodic_r2(v.px[t], v.py[t], v.pz[t], cx, cy, cz, L, halfL);
                if (r2 <= r2max) {
                    Part p = rel_part(v, t, cx, cy, cz, halfL);
                    atomicAdd(&fc[fine_bin(p.r, R, nf)], 1u);
                }
            }
        });
    }
}

// Radial sort of the multi-bin halos: their records were scattered into fine
// radial bins of ~FINE_TARGET records by k_collect, so the stream is already
// ordered at bin granularity and every bin is sorted on its own.  One warp per
// bin: up to 32 records in registers (bitonic network over shuffles, no shared
// memory, no barriers), up to BIN_SMEM in the warp's shared-memory slice; the
// rare larger bins become buckets of the CTA-wide sort kernels.
constexpr int BIN_SMEM = 256;
__global__ void __launch_bounds__(256) k_sort_bins(const int64_t* __restrict__ fine_excl, uint32_t n_fine,
                                                   Counters* ctr, int use_single_base, Rec* __restrict__ recs,
                                                   Bucket* __restrict__ bkt_big, Bucket* __restrict__ bkt_huge) {
    __shared__ Rec sm[8][BIN_SMEM];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const unsigned long long base = use_single_base ? ctr->rec_single : 0ull;  // multi region follows the single region
    const uint32_t nwarp = gridDim.x * 8;
    // one bin of up to 32 / BIN_SMEM records by the whole warp
    auto sort_bin = [&](uint32_t b) {
        const unsigned long long e0 = (unsigned long long)fine_excl[b];
        const uint32_t cnt = (uint32_t)((unsigned long long)fine_excl[b + 1] - e0);
        if (cnt < 2) return;
        Rec* g = recs + base + e0;
        if (cnt <= 32) {
            Rec mine;
            if (lane < (int)cnt) mine = g[lane];
            else { mine.rbits = ~0ull; mine.m = 0.f; mine.flags = 0; }
#pragma unroll
            for (int k = 2; k <= 32; k <<= 1) {
#pragma unroll
                for (int j = k >> 1; j > 0; j >>= 1) {
                    Rec o;
                    o.rbits = __shfl_xor_sync(0xffffffffu, mine.rbits, j);
                    o.m = __shfl_xor_sync(0xffffffffu, mine.m, j);
                    o.flags = __shfl_xor_sync(0xffffffffu, mine.flags, j);
                    const bool take_min = ((lane & k) == 0) == ((lane & j) == 0);
                    if (take_min ? (o.rbits < mine.rbits) : (mine.rbits < o.rbits)) mine = o;
                }
            }
            if (lane < (int)cnt) g[lane] = mine;
        } else if (cnt <= BIN_SMEM) {
            Rec* s = sm[wid];
            for (uint32_t t = lane; t < cnt; t += 32) s[t] = g[t];
            __syncwarp();
            uint32_t np2 = 64;
            while (np2 < cnt) np2 <<= 1;
            auto cx = [&](uint32_t t, uint32_t p) {
                if (p > t && p < cnt) {
                    const Rec x = s[t], y = s[p];
                    if (y.rbits < x.rbits) { s[t] = y; s[p] = x; }
                }
            };
            for (uint32_t k = 2; k <= np2; k <<= 1) {
                for (uint32_t t = lane; t < np2; t += 32) cx(t, t ^ (k - 1));
                __syncwarp();
                for (uint32_t j = k >> 2; j > 0; j >>= 1) {
                    for (uint32_t t = lane; t < np2; t += 32) cx(t, t ^ j);
                    __syncwarp();
                }
            }
            for (uint32_t t = lane; t < cnt; t += 32) g[t] = s[t];
            __syncwarp();
        } else if (lane == 0) {
            Bucket bk;
            bk.start = base + e0;
            bk.count = cnt;
            bk.halo = 0;
            if (cnt <= SB_CAP) bkt_big[atomicAdd(&ctr->n_bkt_big, 1u)] = bk;
            else bkt_huge[atomicAdd(&ctr->n_bkt_huge, 1u)] = bk;
        }
    };
    // bins are made for ~FINE_TARGET = 12 records: two of them share a warp, one per half-warp, with the 16-wide
    // network (10 compare-exchange steps instead of 15, and half the warps); a pair with a fuller bin falls back
    for (uint32_t b0 = 2 * (blockIdx.x * 8 + wid); b0 < n_fine; b0 += 2 * nwarp) {
        const int half = lane >> 4, l16 = lane & 15;
        const uint32_t b = b0 + half;
        unsigned long long e0 = 0;
        uint32_t cnt = 0;
        if (b < n_fine) {
            e0 = (unsigned long long)fine_excl[b];
            cnt = (uint32_t)((unsigned long long)fine_excl[b + 1] - e0);
        }
        const uint32_t other = __shfl_xor_sync(0xffffffffu, cnt, 16);
        if (cnt <= 16 && other <= 16) {
            Rec* g = recs + base + e0;
            Rec mine;
            if (l16 < (int)cnt) mine = g[l16];
            else { mine.rbits = ~0ull; mine.m = 0.f; mine.flags = 0; }
#pragma unroll
            for (int k = 2; k <= 16; k <<= 1) {
#pragma unroll
                for (int j = k >> 1; j > 0; j >>= 1) {
                    Rec o;
                    o.rbits = __shfl_xor_sync(0xffffffffu, mine.rbits, j);
                    o.m = __shfl_xor_sync(0xffffffffu, mine.m, j);
                    o.flags = __shfl_xor_sync(0xffffffffu, mine.flags, j);
                    const bool take_min = ((l16 & k) == 0) == ((l16 & j) == 0);
                    if (take_min ? (o.rbits < mine.rbits) : (mine.rbits < o.rbits)) mine = o;
                }
            }
            if (l16 < (int)cnt && cnt > 1) g[l16] = mine;
        } else {
            sort_bin(b0);
            if (b0 + 1 < n_fine) sort_bin(b0 + 1);
        }
    }
}

// rec_off of the multi-bin halos (their bins are contiguous in the bin table)
__global__ void k_multi_offsets(HaloArrays ha, const uint32_t* __restrict__ multi_list,
                                const unsigned int* __restrict__ n_multi,
                                const int64_t* __restrict__ fine_excl, const Counters* __restrict__ ctr) {
    unsigned int it = blockIdx.x * blockDim.x + threadIdx.x;
    if (it >= *n_multi) return;
    const uint32_t h = multi_list[it];
    ha.rec_off[h] = ctr->rec_single + (unsigned long long)fine_excl[ha.fine_off[h]];
}

// single-bucket halos -> bucket lists (thread per accepted halo)
__global__ void k_single_buckets(HaloArrays ha, const uint32_t* __restrict__ try_list,
                                 const unsigned int* __restrict__ n_try, Counters* ctr,
                                 Bucket* __restrict__ bkt_small, Bucket* __restrict__ bkt_big) {
    unsigned int it = blockIdx.x * blockDim.x + threadIdx.x;
    if (it >= *n_try) return;
    const uint32_t h = try_list[it];
    if (ha.nfine[h] != 0) return;
    const uint32_t cnt = ha.cnt[h];
    if (cnt < 2) return;
    Bucket b;
    b.start = ha.rec_off[h];
    b.count = cnt;
    b.halo = h;
    if (cnt <= SMALL_CAP) bkt_small[atomicAdd(&ctr->n_bkt_small, 1u)] = b;
    else bkt_big[atomicAdd(&ctr->n_bkt_big, 1u)] = b;
}

// ---------------------------------------------------------------- k_collect
__global__ void __launch_bounds__(TB) k_collect(ChunkView v, HaloArrays ha, DevCfg cfg,
                                                const Item* __restrict__ items,
                                                const int64_t* __restrict__ fine_excl,
                                                uint32_t* __restrict__ fine_cursor,
                                                const Counters* __restrict__ ctr, Rec* __restrict__ recs,
                                                unsigned long long* __restrict__ item_minr,
                                                int32_t* __restrict__ item_minfof) {
    __shared__ SweepShared S;
    __shared__ unsigned long long s_minr[TB / 32];
    __shared__ int s_minfof[TB / 32];
    const unsigned int n_items = ctr->n_items;
    for (unsigned int it = blockIdx.x; it < n_items; it += gridDim.x) {
        const Item im = items[it];
        const uint32_t h = im.halo;
        if (ha.state[h] != ST_TRY) continue;
        const double cx = ha.cofp[3 * h], cy = ha.cofp[3 * h + 1], cz = ha.cofp[3 * h + 2];
        const double R = ha.cur_r[h];
        const double r2max = __dmul_rn(R, R), halfL = 0.5 * v.L, L = v.L;
        const uint32_t nf = ha.nfine[h];
        const int64_t hidx = ha.index[h];
        Rec* out = recs + (nf ? ctr->rec_single : ha.rec_off[h]);
        const int64_t* fex = fine_excl + (nf ? ha.fine_off[h] : 0);
        uint32_t* fcur = fine_cursor + (nf ? ha.fine_off[h] : 0);
        unsigned int* cursor = &ha.cursor[h];
        const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
        unsigned long long minr = ~0ull;
        int minfof = -1;
        const bool dmo = cfg.dmo != 0;
        sweep_item<SW_MASS | SW_IDS | SW_TYPE>(v, S, cx, cy, cz, R, im, [&](uint32_t t, bool ok) {
            bool in = false;
            Rec rec;
            uint32_t fb = 0;
            if (ok) {
                double r2 = periodic_r2(v.px[t], v.py[t], v.pz[t], cx, cy, cz, L, halfL);
                if (r2 <= r2max) {
                    in = true;
                    Part p = rel_part(v, t, cx, cy, cz, halfL);
                    rec.rbits = (unsigned long long)__double_as_longlong(p.r);
                    rec.m = v.mass[t];
                    uint32_t tc = dmo ? 1u : (uint32_t)v.type[t];
                    rec.flags = tc | ((v.grnr[t] == hidx) ? 4u : 0u);
                    if (rec.rbits < minr) { minr = rec.rbits; minfof = v.fof[t]; }
                    if (nf) fb = fine_bin(p.r, R, nf);
                }
            }
            if (nf == 0) {
                // warp-aggregated append to the halo's single bucket
                unsigned bal = __ballot_sync(0xffffffffu, in);
                unsigned base = 0;
                if (lane == 0 && bal) base = atomicAdd(cursor, (unsigned)__popc(bal));
                base = __shfl_sync(0xffffffffu, base, 0);
                SOAP_ASSERT(!in || base + __popc(bal & ((1u << lane) - 1u)) < ha.cnt[h]);
                if (in) out[base + __popc(bal & ((1u << lane) - 1u))] = rec;
            } else if (in) {
                unsigned slot = atomicAdd(&fcur[fb], 1u);
                SOAP_ASSERT(fb < nf && (unsigned long long)fex[fb] + slot < (unsigned long long)fex[fb + 1]);
                out[(unsigned long long)fex[fb] + slot] = rec;
            }
        });
        // fofid of the innermost particle (SO_properties.py:407-409): per-item
        // minimum, reduced over the halo's items in k_scan_solve
        for (int o = 16; o > 0; o >>= 1) {
            unsigned long long orr = __shfl_xor_sync(0xffffffffu, minr, o);
            int of = __shfl_xor_sync(0xffffffffu, minfof, o);
            if (orr < minr || (orr == minr && of < minfof)) { minr = orr; minfof = of; }
        }
        if (lane == 0) { s_minr[wid] = minr; s_minfof[wid] = minfof; }
        __syncthreads();
        if (threadIdx.x == 0) {
            for (int w = 1; w < TB / 32; w++)
                if (s_minr[w] < minr || (s_minr[w] == minr && s_minfof[w] < minfof)) { minr = s_minr[w]; minfof = s_minfof[w]; }
            item_minr[it] = minr;
            item_minfof[it] = minfof;
        }
        __syncthreads();
    }
}

// ------------------------------------------- projected half-mass radii
// get_half_weight_radius on the projected radius (projected_aperture_properties.py:
// 953-986 with half_mass_radius.py:16-97), per projection axis: the halo's bound
// particles are binned by projected radius, every bin is sorted (k_sort_bins), and
// a streaming per-type cumulative sum finds, for every aperture and type, the record
// at which the mass inside the aperture is half enclosed.
struct PjPlan {
    uint32_t* nfine;       // [H] bins of the halo (0 = not planned this rung)
    uint32_t* fine_off;    // [H] first bin
    unsigned int* n_fine;  // device total
};

__global__ void k_pj_plan(HaloArrays ha, DevCfg cfg, const uint32_t* __restrict__ acc_list,
                          const unsigned int* __restrict__ n_acc, PjPlan pl) {
    unsigned int it = blockIdx.x * blockDim.x + threadIdx.x;
    if (it >= *n_acc) return;
    const uint32_t h = acc_list[it];
    const int off_pj = (cfg.do_sub ? 1 : 0) + cfg.n_so + cfg.n_ap;
    const int c_lo = ha.commit_lo[h], c_hi = ha.commit_hi[h];
    pl.nfine[h] = 0;
    if (c_hi <= c_lo || ha.status[h] >= 2 || c_hi <= off_pj || c_lo > off_pj) return;
    const ScanRes* sr = ha.sres + h;
    const uint32_t nb = sr->bound_count[0] + sr->bound_count[1] + sr->bound_count[2] + sr->bound_count[3];
    uint32_t nf = nb / FINE_TARGET;
    if (nf < 16) nf = 16;
    pl.nfine[h] = nf;
    pl.fine_off[h] = atomicAdd(pl.n_fine, nf);
}

// FILL = false: histogram of the bound particles over the halo's bins of projected
// radius; FILL = true: scatter their records (projected radius, mass, type) into the bins
template <bool FILL>
__global__ void __launch_bounds__(TB) k_pj_bins(ChunkView v, HaloArrays ha, DevCfg cfg, const Item* __restrict__ items,
                                                const unsigned int* __restrict__ n_items_dev, PjPlan pl, int ax,
                                                uint32_t* __restrict__ fine_cnt,
                                                const int64_t* __restrict__ fine_excl, Rec* __restrict__ recs) {
    __shared__ SweepShared S;
    const unsigned int n_items = *n_items_dev;
    for (unsigned int it = blockIdx.x; it < n_items; it += gridDim.x) {
        const Item im = items[it];
        const uint32_t h = im.halo;
        const uint32_t nf = pl.nfine[h];
        if (nf == 0) continue;
        const double cx = ha.cofp[3 * h], cy = ha.cofp[3 * h + 1], cz = ha.cofp[3 * h + 2];
        const double R = ha.rung_r[h];
        const double r2max = __dmul_rn(R, R), halfL = 0.5 * v.L, L = v.L;
        const int64_t hidx = ha.index[h];
        uint32_t* fc = fine_cnt + pl.fine_off[h];
        const int64_t* fex = FILL ? fine_excl + pl.fine_off[h] : nullptr;
        const bool dmo = cfg.dmo != 0;
        sweep_item<SW_MASS | SW_TYPE>(v, S, cx, cy, cz, R, im, [&](uint32_t t, bool ok) {
            if (!ok || v.grnr[t] != hidx) return;
            const double r2 = periodic_r2(v.px[t], v.py[t], v.pz[t], cx, cy, cz, L, halfL);
            if (!(r2 <= r2max)) return;
            const Part p = rel_part(v, t, cx, cy, cz, halfL);
            const double q0 = ax == 0 ? p.y : p.x, q1 = ax == 2 ? p.y : p.z;
            const double rp = sqrt(__dadd_rn(__dmul_rn(q0, q0), __dmul_rn(q1, q1)));  // projected_aperture_properties.py:122-127
            const uint32_t fb = fine_bin(rp, R, nf);
            if (!FILL) {
                atomicAdd(&fc[fb], 1u);
            } else {
                Rec rc;
                rc.rbits = (unsigned long long)__double_as_longlong(rp);
                rc.m = v.mass[t];
                rc.flags = dmo ? 1u : (uint32_t)v.type[t];
                const uint32_t slot = atomicAdd(&fc[fb], 1u);  // fc is the zeroed cursor array here
                recs[(unsigned long long)fex[fb] + slot] = rc;
            }
        });
    }
}

// One CTA per halo: per-type running sums over the halo's records sorted by
// projected radius; writes HalfMassRadius{Gas,Dm,Star} of every aperture of axis ax.
__global__ void __launch_bounds__(256) k_pj_scan(HaloArrays ha, DevCfg cfg, const uint32_t* __restrict__ acc_list,
                                                 const unsigned int* __restrict__ n_acc, PjPlan pl, int ax,
                                                 const int64_t* __restrict__ fine_excl,
                                                 const Rec* __restrict__ recs) {
    constexpr int NT = 256, K = 4, NG = 3;
    __shared__ double wsum[NT / 32][NG];
    __shared__ double carry[NG];
    __shared__ double thr[SOAP_MAX_APERTURES][NG];
    __shared__ double cap_r[SOAP_MAX_APERTURES][NG], cap_in[SOAP_MAX_APERTURES][NG], cap_ex[SOAP_MAX_APERTURES][NG];
    __shared__ uint32_t cap_i[SOAP_MAX_APERTURES][NG];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int npj = cfg.n_pj;
    for (unsigned int it = blockIdx.x; it < *n_acc; it += gridDim.x) {
        const uint32_t h = acc_list[it];
        const uint32_t nf = pl.nfine[h];
        if (nf == 0) continue;
        const uint32_t fo = pl.fine_off[h];
        const unsigned long long e0 = (unsigned long long)fine_excl[fo];
        const uint32_t n = (uint32_t)((unsigned long long)fine_excl[fo + nf] - e0);
        const Rec* R = recs + e0;
        double* row = ha.out + (int64_t)h * ha.ncol;
        __syncthreads();
        if ((int)threadIdx.x < npj * NG) {
            const int p = threadIdx.x / NG, g = threadIdx.x % NG;
            // half the mass of type g inside aperture p of this projection (written by k_projected);
            // block columns 4.. are the masses of type codes gas, dm, star, bh
            thr[p][g] = 0.5 * row[cfg.lay.pj[p] + ax * cfg.lay.pjb + 4 + g];
            cap_i[p][g] = 0xffffffffu;
        }
        if (threadIdx.x < NG) carry[threadIdx.x] = 0.0;
        __syncthreads();
        for (uint32_t t0 = 0; t0 < n; t0 += NT * K) {
            const uint32_t i0 = t0 + threadIdx.x * K;
            Rec rc[K];
            double loc[NG] = {0.0, 0.0, 0.0};
#pragma unroll
            for (int k = 0; k < K; k++) {
                if (i0 + k < n) {
                    rc[k] = R[i0 + k];
#pragma unroll
                    for (int g = 0; g < NG; g++)
                        if ((rc[k].flags & 3u) == (uint32_t)g) loc[g] += (double)rc[k].m;
                } else {
                    rc[k].rbits = 0; rc[k].m = 0.f; rc[k].flags = 3u;
                }
            }
            double ex[NG];
#pragma unroll
            for (int g = 0; g < NG; g++) {
                double x = loc[g];
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const double y = __shfl_up_sync(0xffffffffu, x, o);
                    if (lane >= o) x += y;
                }
                if (lane == 31) wsum[wid][g] = x;
                ex[g] = x - loc[g];
            }
            __syncthreads();
#pragma unroll
            for (int g = 0; g < NG; g++) {
                double b = carry[g];
                for (int w = 0; w < wid; w++) b += wsum[w][g];
                ex[g] += b;
            }
#pragma unroll
            for (int k = 0; k < K; k++) {
                const uint32_t i = i0 + k;
                if (i >= n) break;
                const uint32_t g = rc[k].flags & 3u;
                if (g >= (uint32_t)NG) continue;
                const double m = (double)rc[k].m;
                const double cex = ex[g], cin = cex + m;
                ex[g] = cin;
                // half_mass_radius.py:63: first record of the type with cumulative weight >= target
                for (int p = 0; p < npj; p++)
                    if (thr[p][g] > 0.0 && cin >= thr[p][g] && !(cex >= thr[p][g])) {
                        cap_r[p][g] = __longlong_as_double((long long)rc[k].rbits);
                        cap_in[p][g] = cin; cap_ex[p][g] = cex; cap_i[p][g] = i;
                    }
            }
            __syncthreads();
            if (threadIdx.x < NG) {
                double b = carry[threadIdx.x];
                for (int w = 0; w < NT / 32; w++) b += wsum[w][threadIdx.x];
                carry[threadIdx.x] = b;
            }
            __syncthreads();
        }
        if ((int)threadIdx.x < npj * NG) {
            const int p = threadIdx.x / NG, g = threadIdx.x % NG;
            double hm = 0.0;
            const uint32_t i = cap_i[p][g];
            if (thr[p][g] > 0.0 && i != 0xffffffffu) {
                const double rmax_ = cap_r[p][g], Wmax = cap_in[p][g], Wmin = cap_ex[p][g];
                double rmin_ = 0.0;
                for (uint32_t j = i; j-- > 0;)
                    if ((R[j].flags & 3u) == (uint32_t)g) {
                        rmin_ = __longlong_as_double((long long)R[j].rbits);
                        break;
                    }
                // half_mass_radius.py:64-80
                if (Wmin == Wmax) hm = 0.5 * (rmin_ + rmax_);
                else hm = rmin_ + (thr[p][g] - Wmin) / (Wmax - Wmin) * (rmax_ - rmin_);
            }
            row[cfg.lay.pj[p] + ax * cfg.lay.pjb + 18 + g] = hm;
        }
        __syncthreads();
    }
}

// ------------------------------------------------------------ bucket sorts
struct LessRec {
    __device__ __forceinline__ bool operator()(const Rec& a, const Rec& b) const { return a.rbits < b.rbits; }
};

template <int CAP, int NT>
__global__ void __launch_bounds__(NT) k_sort_bucket(const Bucket* __restrict__ bkts,
                                                    const unsigned int* __restrict__ n_bkt,
                                                    Rec* __restrict__ recs) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    Rec* s = (Rec*)smem_raw;
    for (unsigned int it = blockIdx.x; it < *n_bkt; it += gridDim.x) {
        const Bucket b = bkts[it];
        Rec* g = recs + b.start;
        __syncthreads();
        if (CAP > 0) {
            for (uint32_t t = threadIdx.x; t < b.count; t += NT) s[t] = g[t];
            __syncthreads();
            block_bitonic_sort(s, b.count, LessRec());
            for (uint32_t t = threadIdx.x; t < b.count; t += NT) g[t] = s[t];
        } else {
            block_bitonic_sort(g, b.count, LessRec());  // oversize bucket: in global memory
        }
    }
}

// -------------------------------------------------------------- scan pass
// One CTA (CS == 1) or one cluster of CS CTAs (halos above SCAN_BIG records)
// per halo of the list; the per-halo work is scan_solve_halo (scan.cuh).
template <int NCH, int CS>
__global__ void __cluster_dims__(CS, 1, 1) __launch_bounds__(SCAN_NT, NCH == 2 ? (CS == 1 ? 3 : 2) : 1)
    k_scan_solve(HaloArrays ha, DevCfg cfg, const uint32_t* __restrict__ try_list,
                 const unsigned int* __restrict__ n_try, const Rec* __restrict__ recs,
                 uint32_t* __restrict__ next, Counters* ctr,
                 const unsigned long long* __restrict__ item_minr,
                 const int32_t* __restrict__ item_minfof, int cursor_slot) {
    __shared__ ScanShared<NCH, SCAN_NT> S;
    __shared__ unsigned int s_it;
    cg::cluster_group cluster = cg::this_cluster();
    const unsigned int n_list = *n_try;
    // dynamic queue: the lists are roughly largest-first, and a cluster that drew a 5-million-record halo must not
    // also own every 37th of the rest
    unsigned int* cursor = &ctr->scan_cursor[cursor_slot];
    while (true) {
        if (CS > 1) {
            if (cluster.block_rank() == 0 && threadIdx.x == 0) s_it = atomicAdd(cursor, 1u);
            cluster.sync();
        } else {
            __syncthreads();
            if (threadIdx.x == 0) s_it = atomicAdd(cursor, 1u);
            __syncthreads();
        }
        const unsigned int it = CS > 1 ? *cluster.map_shared_rank(&s_it, 0) : s_it;
        if (CS > 1) cluster.sync();  // everyone has read it before rank 0 draws again
        if (it >= n_list) break;
        const uint32_t h = try_list[it];
        const uint32_t ib = ha.item_base[h];
        scan_solve_halo<NCH, CS, SCAN_NT, (CS > 1 ? 2 * SCAN_K : SCAN_K)>(S, ha, cfg, h, ha.cnt[h], recs + ha.rec_off[h], next, ctr,
                                                  item_minr + ib, item_minfof + ib, ha.n_items[h]);
    }
}


DevCfg make_devcfg(const soap_halo_config& c) {
    DevCfg d;
    memset(&d, 0, sizeof(d));
    d.L = c.boxsize; d.halfL = 0.5 * c.boxsize; d.G = c.G; d.H = c.H; d.kpc = c.kpc_per_length;
    d.r20 = c.r_20mpc; d.nu = c.nu_density; d.mpc2c = c.phys_mpc_to_coord;
    const int pt[4] = {0, 1, 4, 5};
    for (int t = 0; t < 4; t++) d.soft[t] = c.softening[pt[t]];
    d.target_density = c.target_density;
    d.do_sub = c.do_subhalo; d.n_so = c.n_so; d.n_ap = c.n_apertures; d.n_pj = c.n_projected; d.dmo = c.dmo;
    for (int a = 0; a < SOAP_MAX_APERTURES; a++) d.pj_r[a] = c.proj_radius[a];
    for (int k = 0; k < SOAP_MAX_SO; k++) { d.so_rho[k] = c.so_reference_density[k]; d.so_virial[k] = c.so_virial[k]; }
    for (int a = 0; a < SOAP_MAX_APERTURES; a++) {
        d.ap_r[a] = c.ap_radius[a]; d.ap_mpc[a] = c.ap_physical_mpc[a]; d.ap_incl[a] = c.ap_inclusive[a];
    }
    d.flags = c.property_flags;
    d.n_filters = c.n_filters;
    for (int f = 0; f < SOAP_MAX_FILTERS; f++) { d.filter_limit[f] = c.filter_limit[f]; d.filter_types[f] = c.filter_types[f]; }
    for (int k = 0; k < SOAP_MAX_SO; k++) d.so_filter[k] = c.so_filter[k];
    for (int a = 0; a < SOAP_MAX_APERTURES; a++) {
        d.ap_filter[a] = c.ap_filter[a]; d.pj_filter[a] = c.proj_filter[a]; d.ap_prev[a] = c.ap_prev_radius[a];
    }
    d.lay = row_layout(c);
    return d;
}

int validate_cfg(const soap_halo_config* cfg) {
    if (cfg->n_so < 0 || cfg->n_so > SOAP_MAX_SO) SOAP_FAIL("config: n_so=%d outside [0,%d]", cfg->n_so, SOAP_MAX_SO);
    if (cfg->n_apertures < 0 || cfg->n_apertures > SOAP_MAX_APERTURES)
        SOAP_FAIL("config: n_apertures=%d outside [0,%d]", cfg->n_apertures, SOAP_MAX_APERTURES);
    if (cfg->n_projected < 0 || cfg->n_projected > SOAP_MAX_APERTURES)
        SOAP_FAIL("config: n_projected=%d outside [0,%d]", cfg->n_projected, SOAP_MAX_APERTURES);
    for (int a = 1; a < cfg->n_projected; a++)
        if (cfg->proj_radius[a] < cfg->proj_radius[a - 1]) SOAP_FAIL("config: projected aperture radii must ascend");
    if ((cfg->property_flags & PF_ITER) && !(cfg->property_flags & PF_TENS))
        SOAP_FAIL("config: iterative inertia tensors (property_flags bit 4) need the tensor group (bit 2)");
    if (cfg->n_projected > 0 && !cfg->do_subhalo)
        SOAP_FAIL("config: projected apertures need BoundSubhalo first (they assume every bound particle is loaded)");
    if (!(cfg->boxsize > 0.0)) SOAP_FAIL("config: boxsize must be positive");
    if (cfg->n_filters < 0 || cfg->n_filters > SOAP_MAX_FILTERS)
        SOAP_FAIL("config: n_filters=%d outside [0,%d]", cfg->n_filters, SOAP_MAX_FILTERS);
    {
        bool uses = false;
        auto chk = [&](int f) { if (f < 0 || (f > 0 && f >= cfg->n_filters)) return false; if (f > 0) uses = true; return true; };
        for (int k = 0; k < cfg->n_so; k++) if (!chk(cfg->so_filter[k])) SOAP_FAIL("config: so_filter[%d]=%d is not a filter index", k, cfg->so_filter[k]);
        for (int a = 0; a < cfg->n_apertures; a++) if (!chk(cfg->ap_filter[a])) SOAP_FAIL("config: ap_filter[%d]=%d is not a filter index", a, cfg->ap_filter[a]);
        for (int a = 0; a < cfg->n_projected; a++) if (!chk(cfg->proj_filter[a])) SOAP_FAIL("config: proj_filter[%d]=%d is not a filter index", a, cfg->proj_filter[a]);
        // the filters read BoundSubhalo's particle counts (category_filter.py:91-102)
        if (uses && !cfg->do_subhalo) SOAP_FAIL("config: category filters need BoundSubhalo (do_subhalo)");
    }
    for (int a = 1; a < cfg->n_apertures; a++)
        if (cfg->ap_radius[a] < cfg->ap_radius[a - 1]) SOAP_FAIL("config: aperture radii must ascend");
    return 0;
}

}  // namespace

extern "C" {

int64_t soap_result_layout(const soap_halo_config* cfg, char* buf, int64_t buflen) {
    if (!cfg) { snprintf(g_soap_err, sizeof(g_soap_err), "soap_result_layout: NULL config"); return -1; }
    if (validate_cfg(cfg)) return -1;
    RowLayout L = row_layout(*cfg);
    std::string s;
    auto add = [&](const std::string& name, int w) { s += name + ":" + std::to_string(w) + "\n"; };
    add("InputHalos/status", 1); add("InputHalos/n_loop", 1); add("InputHalos/radius", 1);
    add("InputHalos/n_pairs", 1); add("InputHalos/search_radius", 1); add("InputHalos/read_radius", 1);
    auto block = [&](const std::string& p, int kind) {
        const char* b17[] = {"Ngas", "Ndm", "Nstar", "Nbh", "Mgas", "Mdm", "Mstar", "Mbh"};
        for (int i = 0; i < 8; i++) add(p + b17[i], 1);
        add(p + "Mtot", 1); add(p + "com", 3); add(p + "vcom", 3); add(p + "Vmax_soft", 1); add(p + "R_vmax_soft", 1);
        if (cfg->property_flags & PF_KIN) {
            const char* g[] = {"gas", "dm", "star"};
            for (int i = 0; i < 3; i++) {
                add(p + "com_" + g[i], 3); add(p + "vcom_" + g[i], 3); add(p + "L" + g[i], 3);
                add(p + "veldisp_matrix_" + g[i], 6);
            }
            add(p + "Lbaryons", 3); add(p + "Ekin_tot", 1); add(p + "Ekin_gas", 1); add(p + "Ekin_star", 1);
        }
        if (cfg->property_flags & PF_KAPPA) {
            add(p + "kappa_corot_gas", 1); add(p + "kappa_corot_star", 1); add(p + "kappa_corot_baryons", 1);
            add(p + "DtoTgas", 1); add(p + "DtoTstar", 1);
            add(p + "StellarRotationalVelocity", 1); add(p + "StellarCylindricalVelocityDispersion", 1);
            add(p + "StellarCylindricalVelocityDispersionVertical", 1);
            add(p + "StellarCylindricalVelocityDispersionDiscPlane", 1);
        }
        if (cfg->property_flags & PF_TENS) {
            add(p + (kind == 2 ? "StellarInertiaTensorNoniterative" : "TotalInertiaTensorNoniterative"), 6);
            add(p + (kind == 2 ? "StellarInertiaTensorReducedNoniterative" : "TotalInertiaTensorReducedNoniterative"), 6);
            if (cfg->property_flags & PF_ITER) {
                add(p + (kind == 2 ? "StellarInertiaTensor" : "TotalInertiaTensor"), 6);
                add(p + (kind == 2 ? "StellarInertiaTensorReduced" : "TotalInertiaTensorReduced"), 6);
            }
        }
        if (cfg->property_flags & PF_HMR) {
            add(p + "HalfMassRadiusGas", 1); add(p + "HalfMassRadiusDM", 1); add(p + "HalfMassRadiusStar", 1);
            add(p + "HalfMassRadiusBaryon", 1);
        }
        if (kind == 0) {
            add(p + "HalfMassRadiusTot", 1); add(p + "EncloseRadius", 1); add(p + "Vmax_unsoft", 1);
            add(p + "R_vmax_unsoft", 1); add(p + "spin_parameter", 1);
        } else if (kind == 1) {
            add(p + "r", 1); add(p + "Mso", 1); add(p + "spin_parameter", 1); add(p + "Mfrac_satellites", 1);
            add(p + "Mfrac_external", 1); add(p + "concentration_unsoft", 1); add(p + "concentration_soft", 1);
            add(p + "concentration_dmo_unsoft", 1); add(p + "concentration_dmo_soft", 1);
        }
    };
    if (cfg->do_subhalo) block("BoundSubhalo/", 0);
    for (int k = 0; k < cfg->n_so; k++) block("SO/" + std::to_string(k) + "/", 1);
    for (int a = 0; a < cfg->n_apertures; a++) block("Aperture/" + std::to_string(a) + "/", 2);
    for (int a = 0; a < cfg->n_projected; a++)
        for (int ax = 0; ax < 3; ax++) {
            const std::string p = "ProjectedAperture/" + std::to_string(a) + "/proj" + std::string(1, "xyz"[ax]) + "/";
            const char* nm[] = {"Ngas", "Ndm", "Nstar", "Nbh", "Mgas", "Mdm", "Mstar", "Mbh"};
            for (int i = 0; i < 8; i++) add(p + nm[i], 1);
            add(p + "Mtot", 1); add(p + "com", 3); add(p + "vcom", 3);
            add(p + "proj_veldisp_gas", 1); add(p + "proj_veldisp_dm", 1); add(p + "proj_veldisp_star", 1);
            add(p + "HalfMassRadiusGas", 1); add(p + "HalfMassRadiusDm", 1); add(p + "HalfMassRadiusStar", 1);
            add(p + "ProjectedTotalInertiaTensorNoniterative", 3);
            add(p + "ProjectedTotalInertiaTensorReducedNoniterative", 3);
            if (cfg->property_flags & PF_ITER) {
                add(p + "ProjectedTotalInertiaTensor", 3);
                add(p + "ProjectedTotalInertiaTensorReduced", 3);
            }
        }
    if (buf && buflen > 0) {
        if ((int64_t)s.size() + 1 > buflen) { snprintf(g_soap_err, sizeof(g_soap_err), "soap_result_layout: buffer too small (%zu needed)", s.size() + 1); return -1; }
        memcpy(buf, s.data(), s.size());
        buf[s.size()] = 0;
    }
    return L.ncol;
}

int soap_process_halos(soap_chunk* c, const soap_halo_config* cfg, int64_t n_halo,
                       const double* cofp_dev, const double* search_radius_dev,
                       const double* read_radius_dev, const int64_t* index_dev,
                       const int32_t* is_central_dev, const int64_t* nr_bound_part_dev,
                       double* out_dev, int64_t ncol, int32_t* status_dev, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    if (!c || !cfg) SOAP_FAIL("soap_process_halos: NULL argument");
    if (validate_cfg(cfg)) return -1;
    if (n_halo <= 0) return 0;
    if (n_halo >= (1ll << 31)) SOAP_FAIL("soap_process_halos: too many halos");
    soap_handle* h = c->h;
    CUDA_TRY(cudaSetDevice(h->device));
    DevCfg dc = make_devcfg(*cfg);
    if (ncol != dc.lay.ncol) SOAP_FAIL("soap_process_halos: ncol=%lld but the layout has %d columns", (long long)ncol, dc.lay.ncol);
    if (cfg->dmo && (c->type_present[0] || c->type_present[2] || c->type_present[3]))
        SOAP_FAIL("soap_process_halos: dmo config on a chunk with baryonic particle types");
    const ChunkView& v = c->v;
    const uint32_t H = (uint32_t)n_halo;

    HaloArrays ha;
    ha.cofp = cofp_dev; ha.sr_in = search_radius_dev; ha.rr_in = read_radius_dev; ha.index = index_dev;
    ha.central = is_central_dev; ha.nexp = nr_bound_part_dev; ha.out = out_dev; ha.ncol = ncol;
    ha.status = status_dev;
    WS_GET(cur_r, double, h, "h_cur_r", H); ha.cur_r = cur_r;
    WS_GET(rung_r, double, h, "h_rung_r", H); ha.rung_r = rung_r;
    WS_GET(nloop, int32_t, h, "h_nloop", H); ha.nloop = nloop;
    WS_GET(state, int32_t, h, "h_state", H); ha.state = state;
    WS_GET(cnt, uint32_t, h, "h_cnt", H); ha.cnt = cnt;
    WS_GET(msum, double, h, "h_msum", H); ha.msum = msum;
    WS_GET(rec_off, unsigned long long, h, "h_rec_off", H); ha.rec_off = rec_off;
    WS_GET(fine_off, uint32_t, h, "h_fine_off", H); ha.fine_off = fine_off;
    WS_GET(nfine, uint32_t, h, "h_nfine", H); ha.nfine = nfine;
    WS_GET(ndone, int32_t, h, "h_ndone", H); ha.ndone = ndone;
    WS_GET(commit_lo, int32_t, h, "h_commit_lo", H); ha.commit_lo = commit_lo;
    WS_GET(commit_hi, int32_t, h, "h_commit_hi", H); ha.commit_hi = commit_hi;
    WS_GET(sres, ScanRes, h, "h_sres", H); ha.sres = sres;
    WS_GET(item_base, uint32_t, h, "h_item_base", H); ha.item_base = item_base;
    WS_GET(n_items_arr, uint32_t, h, "h_n_items", H); ha.n_items = n_items_arr;
    WS_GET(cursor, unsigned int, h, "h_cursor", H); ha.cursor = cursor;
    WS_GET(items_done, unsigned int, h, "h_items_done", H); ha.items_done = items_done;
    WS_GET(mslot, int32_t, h, "h_mslot", H); ha.mslot = mslot;
    WS_GET(listA, uint32_t, h, "h_listA", H);
    WS_GET(listB, uint32_t, h, "h_listB", H);
    WS_GET(try_list, uint32_t, h, "h_try", H);
    WS_GET(big_list, uint32_t, h, "h_big", H);
    WS_GET(acc_list, uint32_t, h, "h_acc", H);
    WS_GET(look_arr, int32_t, h, "h_look", H); ha.look = look_arr;
    WS_GET(rung_cnt, uint32_t, h, "h_rung_cnt", (size_t)H * LOOK_MAX); ha.rung_cnt = rung_cnt;
    WS_GET(rung_msum, double, h, "h_rung_msum", (size_t)H * LOOK_MAX); ha.rung_msum = rung_msum;
    WS_GET(multi_list, uint32_t, h, "h_multi", H);
    WS_GET(seq_list, uint32_t, h, "h_seq", H);
    WS_GET(huge_list, uint32_t, h, "h_huge", H);
    WS_GET(bank_off, unsigned long long, h, "h_bank_off", H); ha.bank_off = bank_off;
    WS_GET(cuts_arr, Cuts, h, "h_cuts", H); ha.cuts = cuts_arr;
    ha.gbank = nullptr;
    ha.kraw_nb = (dc.do_sub ? 1 : 0) + dc.n_so + dc.n_ap;
    ha.kraw = nullptr;
    if (dc.flags & PF_KAPPA) {
        WS_GET(kraw, double, h, "h_kraw", (size_t)H * ha.kraw_nb * 2 + 2);
        ha.kraw = kraw;
        CUDA_TRY(cudaMemsetAsync(kraw, 0, sizeof(double) * (size_t)H * ha.kraw_nb * 2, stream));
    }
    const int bank_stride = soap_bank_stride(dc);
    WS_GET(ctr, Counters, h, "h_ctr", 8);  // [0] current round, [2..4] the fused tiers
    WS_GET(n_pend_dev, unsigned int, h, "h_npend", 4);
    // work items: every halo has at least one; large spheres are cut every ITEM_CAND candidates
    size_t items_cap = (size_t)H + (size_t)(16 * (v.n / ITEM_CAND + 1)) + 1024;
    WS_GET(items, Item, h, "h_items", items_cap);
    WS_GET(item_minr, unsigned long long, h, "h_item_minr", items_cap);
    WS_GET(item_minfof, int32_t, h, "h_item_minfof", items_cap);
    WS_GET(item_msum, double, h, "h_item_msum", items_cap * LOOK_MAX);

    PhaseLog& log = c->halo_log;
    log.reset();
    c->last_pairs = 0;
    c->last_candidates = 0;
    for (int k = 0; k < 4; k++) c->last_rec_class[k] = 0;
    c->last_rounds = 0;
    uint32_t* pend = listA;
    uint32_t* next = listB;
    // tiers: the staged small-halo path (tier.cu) next to the rest; its overflow joins below
    constexpr int NTIER = 2;
    const long long tier_nexp[NTIER] = {SMALL_NEXP_0, SMALL_NEXP_1};
    uint32_t* tier_list[3 + 1];
    WS_GET(list0, uint32_t, h, "h_list0", H); tier_list[0] = list0;
    WS_GET(list1, uint32_t, h, "h_list1", H); tier_list[1] = list1;
    WS_GET(list2, uint32_t, h, "h_list2", H); tier_list[2] = list2;  // unused third size class of k_tier_place
    WS_GET(tier_n, unsigned int, h, "h_tier_n", 16);  // [0..3] list sizes, [8..] queue cursor, [12] round list size
    CUDA_TRY(cudaMemsetAsync(tier_n, 0, 16 * sizeof(unsigned int), stream));
    CUDA_TRY(cudaMemsetAsync(ctr + 2, 0, sizeof(Counters), stream));
    TierLims tl;
    // projected apertures and the general-path cross-check switch keep every halo out of the tiers
    const bool tiers_on = dc.n_pj == 0 && !(cfg->debug_flags & 1u);
    // measurement switch: every kernel on the caller's stream, one after the other (per-kernel event times then
    // measure work, not residency next to the kernels of other streams)
    const bool serial = (cfg->debug_flags & 2u) != 0;
    for (int t = 0; t < 3; t++) tl.lim[t] = tiers_on ? tier_nexp[t < NTIER ? t : NTIER - 1] : -1;
    for (int t = 0; t < 3; t++)
        if (tl.lim[t] >= TIER_BUCKETS) SOAP_FAIL("soap_process_halos: tier limit above %d", TIER_BUCKETS - 1);
    WS_GET(size_hist, unsigned int, h, "h_size_hist", 2 * TIER_BUCKETS);
    CUDA_TRY(cudaMemsetAsync(size_hist, 0, 2 * TIER_BUCKETS * sizeof(unsigned int), stream));
    LAUNCH(h, k_init, grid_for(H, 128), 128, 0, stream, ha, (int64_t)H, pend, tier_n, size_hist, tl);
    unsigned long long total_pairs = 0, total_cand = 0, total_count_pairs = 0, total_try_pairs = 0, total_mom_pairs = 0;
    for (int t = 0; t < 3; t++) c->last_tier_pairs[t] = 0;
    bool tiers_running = false;
    // a failed call must not leave tier kernels in flight on their own streams (they use the handle's workspace)
    struct TierGuard {
        const bool& running;
        ~TierGuard() { if (running) cudaDeviceSynchronize(); }
    } tier_guard{tiers_running};
    unsigned int n_pend_first = 0;
    if (tiers_on) {
        LAUNCH(h, k_tier_scan, 1, TIER_BUCKETS, 0, stream, size_hist, tier_n, tl);
        LAUNCH(h, k_tier_place, grid_for(H, 128), 128, 0, stream, ha, (int64_t)H, list0, list1, list2,
               size_hist + TIER_BUCKETS, size_hist, tl);
        // The tiers run asynchronously: all rounds of both tiers are enqueued up front on two streams of their own
        // (list sizes stay on the device, grids are sized for upper bounds), while the general path's first round
        // runs on the caller's stream over the halos that were too large for the tiers from the start.  Each
        // tier's solve is a thread-per-halo kernel whose tail leaves most of the GPU idle; the other tier's and the
        // general path's sweeps fill it.  A halo that outgrows tier 0 joins tier 1's next round (tier 1 waits for the
        // event that closes tier 0's round).  In a tier's last round overflow and stragglers (halos still asking for
        // a larger radius) go to `late`, which joins the general path's pending list after its first round.
        if (h->tier_init()) SOAP_FAIL("soap_process_halos: cannot create the tier streams");
        static_assert(TIER_ROUNDS + 1 <= soap_handle::TIER_EV, "one event per tier round");
        constexpr int NR = TIER_ROUNDS + 1;  // tier 0 runs TIER_ROUNDS rounds, tier 1 one more
        WS_GET(tctr_all, Counters, h, "h_tctr", NTIER * NR);
        WS_GET(tcur, unsigned int, h, "h_tcur", NTIER * NR);
        CUDA_TRY(cudaMemsetAsync(tctr_all, 0, NTIER * NR * sizeof(Counters), stream));
        CUDA_TRY(cudaMemsetAsync(tcur, 0, NTIER * NR * sizeof(unsigned int), stream));
        // one output list per (tier, round): tier 0 runs ahead of tier 1 and appends to tier 1's lists of later rounds
        // while tier 1 still reads its earlier ones
        uint32_t* t_next[NTIER][NR];
        uint32_t* t_try[NTIER];
        unsigned long long* t_minr[NTIER];
        int32_t* t_minfof[NTIER];
        {
            const char* nm[NTIER][4] = {{"h_t0_next", "h_t0_try", "h_t0_minr", "h_t0_minfof"},
                                        {"h_t1_next", "h_t1_try", "h_t1_minr", "h_t1_minfof"}};
            for (int t = 0; t < NTIER; t++) {
                uint32_t* lists = (uint32_t*)h->get(nm[t][0], sizeof(uint32_t) * (size_t)H * NR);
                t_try[t] = (uint32_t*)h->get(nm[t][1], sizeof(uint32_t) * (size_t)H);
                t_minr[t] = (unsigned long long*)h->get(nm[t][2], sizeof(unsigned long long) * (size_t)H);
                t_minfof[t] = (int32_t*)h->get(nm[t][3], sizeof(int32_t) * (size_t)H);
                if (!lists || !t_try[t] || !t_minr[t] || !t_minfof[t]) return -1;
                for (int r = 0; r < NR; r++) t_next[t][r] = lists + (size_t)H * r;
            }
        }
        uint32_t* late = list2;  // counted in tier_n[4]
        unsigned int n0[4] = {0, 0, 0, 0};
        CUDA_TRY(cudaMemcpyAsync(n0, tier_n, 4 * sizeof(unsigned int), cudaMemcpyDeviceToHost, stream));
        CUDA_TRY(cudaStreamSynchronize(stream));
        // upper bounds of the list sizes: tier 0's lists only shrink; tier 1 also receives what tier 0 hands over,
        // provisioned for an eighth of tier 0 (a longer list sends its surplus on to the general path)
        unsigned int extra = n0[0] / 8 > 4096u ? n0[0] / 8 : 4096u;
        if (extra > n0[0]) extra = n0[0];
        const unsigned int n_up[NTIER] = {n0[0], n0[1] + extra};
        if (n_up[0] + n_up[1] > 0) {
            CUDA_TRY(cudaEventRecord(h->ev_tfork, stream));
            for (int t = 0; t < NTIER; t++) {
                cudaStream_t ts = serial ? stream : h->tstream[t];
                CUDA_TRY(cudaStreamWaitEvent(ts, h->ev_tfork, 0));
                if (n_up[t] == 0) continue;
                const int rounds = TIER_ROUNDS + t;
                log.begin(t == 0 ? "tier_0" : "tier_1", ts);
                for (int round = 0; round < rounds; round++) {
                    Counters* tc = tctr_all + t * NR + round;
                    const bool last = round + 1 >= rounds;
                    const bool to_general = last || t + 1 >= NTIER;
                    // tier 0, round r -> tier 1's list of round r + 1 (the list tier 1's round r also appends to)
                    uint32_t* ovf = to_general ? late : t_next[t + 1][round];
                    unsigned int* n_ovf = to_general ? tier_n + 4 : &tctr_all[(t + 1) * NR + round].n_next;
                    const uint32_t* list = round == 0 ? tier_list[t] : t_next[t][round - 1];
                    const unsigned int* n_dev = round == 0 ? tier_n + t : &tctr_all[t * NR + round - 1].n_next;
                    if (t == 1 && round >= 1 && n_up[0] > 0 && round - 1 < TIER_ROUNDS)
                        CUDA_TRY(cudaStreamWaitEvent(ts, h->ev_tround[round - 1], 0));
                    if (soap_tier_round(c, dc, ha, t, list, n_dev, n_up[t], ovf, n_ovf, tcur + t * NR + round, t_try[t],
                                        last ? late : t_next[t][round], last ? tier_n + 4 : &tc->n_next, tc,
                                        t_minr[t], t_minfof[t], ts) < 0)
                        return -1;
                    if (t == 0) CUDA_TRY(cudaEventRecord(h->ev_tround[round], ts));
                }
                log.end(ts);
                CUDA_TRY(cudaEventRecord(h->ev_tjoin[t], ts));
            }
            tiers_running = true;
        }
        n_pend_first = n0[3];
    }
    // wait for the tiers, add up their counters; returns the number of halos in `late` (list2)
    auto join_tiers = [&](unsigned int* n_late) -> int {
        *n_late = 0;
        if (!tiers_running) return 0;
        tiers_running = false;
        constexpr int NR = TIER_ROUNDS + 1;
        Counters* tctr_all = (Counters*)h->get("h_tctr", NTIER * NR * sizeof(Counters));
        if (!tctr_all) return -1;
        for (int t = 0; t < NTIER; t++) CUDA_TRY(cudaStreamWaitEvent(stream, h->ev_tjoin[t], 0));
        Counters tc[NTIER * NR];
        CUDA_TRY(cudaMemcpyAsync(tc, tctr_all, sizeof(tc), cudaMemcpyDeviceToHost, stream));
        CUDA_TRY(cudaMemcpyAsync(n_late, tier_n + 4, sizeof(unsigned int), cudaMemcpyDeviceToHost, stream));
        CUDA_TRY(cudaStreamSynchronize(stream));
        for (int t = 0; t < NTIER; t++)
            for (int r = 0; r < NR; r++) {
                total_pairs += tc[t * NR + r].pairs;
                total_cand += tc[t * NR + r].candidates;
                c->last_tier_pairs[t] += (int64_t)tc[t * NR + r].pairs;
                c->last_small_pairs += (int64_t)tc[t * NR + r].pairs;
            }
        return 0;
    };
    unsigned int n_pend = 0;
    if (tiers_on) {
        n_pend = n_pend_first;
    } else {
        CUDA_TRY(cudaMemcpyAsync(&n_pend, tier_n + 3, sizeof(unsigned int), cudaMemcpyDeviceToHost, stream));
        CUDA_TRY(cudaStreamSynchronize(stream));
    }
    c->last_small_pairs = 0;
    if (n_pend == 0) {  // nothing but small halos: what the tiers hand over is the pending list
        unsigned int n_late = 0;
        if (join_tiers(&n_late)) return -1;
        if (n_late > 0) CUDA_TRY(cudaMemcpyAsync(pend, list2, n_late * sizeof(uint32_t), cudaMemcpyDeviceToDevice, stream));
        n_pend = n_late;
    }
    CUDA_TRY(cudaMemcpyAsync(n_pend_dev, &n_pend, sizeof(unsigned int), cudaMemcpyHostToDevice, stream));
    CUDA_TRY(cudaStreamSynchronize(stream));
    const int sm = h->sm_count;
    // per device, so set on every call (cheap)
    CUDA_TRY(cudaFuncSetAttribute(k_sort_bucket<SB_CAP, 512>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  (int)(SB_CAP * sizeof(Rec))));
    const unsigned int sweep_grid = (unsigned)(sm * 6);  // persistent CTAs striding over the item list
    // plan a sweep per halo of `list`; coarse meshes / huge spheres can need more
    // work items than provisioned: grow the list and plan again
    auto plan = [&](const uint32_t* list, const unsigned int* n_dev, unsigned int n_host, int look, int replan) -> int {
        for (int attempt = 0; attempt < 2; attempt++) {
            CUDA_TRY(cudaMemsetAsync(&ctr->n_items, 0, 4 * sizeof(unsigned int), stream));  // n_items, n_mslot, items_overflow, n_bslot
            LAUNCH(h, k_plan_items, grid_for(n_host, PLAN_NT / 32), PLAN_NT, 0, stream, v, ha, list, n_dev, items,
                   (unsigned int)items_cap, ctr, look, replan, bank_stride);
            Counters pc;
            CUDA_TRY(cudaMemcpyAsync(&pc, ctr, sizeof(Counters), cudaMemcpyDeviceToHost, stream));
            CUDA_TRY(cudaStreamSynchronize(stream));
            if (!pc.items_overflow) return 0;
            if (attempt == 1 || pc.n_items >= 0xfff00000u)
                SOAP_FAIL("soap_process_halos: work item list overflow (%u items)", pc.n_items);
            items_cap = (size_t)pc.n_items + 1024;
            items = (Item*)h->get("h_items", sizeof(Item) * items_cap);
            item_minr = (unsigned long long*)h->get("h_item_minr", sizeof(unsigned long long) * items_cap);
            item_minfof = (int32_t*)h->get("h_item_minfof", sizeof(int32_t) * items_cap);
            item_msum = (double*)h->get("h_item_msum", sizeof(double) * items_cap * LOOK_MAX);
            if (!items || !item_minr || !item_minfof || !item_msum) return -1;
            if (!replan) CUDA_TRY(cudaMemsetAsync(&ctr->candidates, 0, sizeof(unsigned long long), stream));
        }
        return 0;
    };
    while (n_pend > 0) {
        c->last_rounds++;
        if (c->last_rounds > 200) SOAP_FAIL("soap_process_halos: radius ladder did not terminate");
        CUDA_TRY(cudaMemsetAsync(ctr, 0, sizeof(Counters), stream));
        // ladder look-ahead: the first round covers 4 rungs per sweep, stragglers more
        const int look = c->last_rounds == 1 ? 4 : LOOK_MAX;
        log.begin("plan", stream);
        if (plan(pend, n_pend_dev, n_pend, look, 0)) return -1;
        log.end(stream);
        log.begin("count", stream);
        // the first round's sweep covers 4 rungs, the stragglers' LOOK_MAX: two instantiations, so that the bulk does
        // not carry twelve counters and mass sums per thread
        if (look <= 4) LAUNCH(h, k_count<4>, sweep_grid, TB, 0, stream, v, ha, items, ctr, item_msum);
        else LAUNCH(h, k_count<LOOK_MAX>, sweep_grid, TB, 0, stream, v, ha, items, ctr, item_msum);
        log.end(stream);
        log.begin("gate", stream);
        LAUNCH(h, k_gate, grid_for(n_pend, 128), 128, 0, stream, ha, dc, pend, n_pend_dev, try_list, big_list,
               acc_list, multi_list, seq_list, huge_list, next, ctr, item_msum);
        log.end(stream);
        Counters hc;
        CUDA_TRY(cudaMemcpyAsync(&hc, ctr, sizeof(Counters), cudaMemcpyDeviceToHost, stream));
        CUDA_TRY(cudaStreamSynchronize(stream));
        total_cand += hc.candidates;
        total_count_pairs += hc.count_pairs;
        total_try_pairs += hc.rec_total;
        for (int k = 0; k < 4; k++) c->last_rec_class[k] += (int64_t)hc.rec_class[k];
        if (hc.n_try + hc.n_big + hc.n_seq + hc.n_huge > 0) {
            const unsigned int n_try = hc.n_try + hc.n_big + hc.n_seq + hc.n_huge;
            // the accepted radius is generally smaller than the swept one: plan its sweep
            log.begin("plan", stream);
            if (plan(acc_list, &ctr->n_acc, n_try, look, 1)) return -1;
            log.end(stream);
            CUDA_TRY(cudaMemcpyAsync(&hc, ctr, sizeof(Counters), cudaMemcpyDeviceToHost, stream));
            CUDA_TRY(cudaStreamSynchronize(stream));
            // workspace for this round
            Rec* recs = (Rec*)h->get("h_recs", sizeof(Rec) * (size_t)(hc.rec_total + 1));
            if (!recs) return -1;
            ha.gbank = (double*)h->get("h_gbank", sizeof(double) * (size_t)bank_stride * (hc.n_bslot + 1));
            if (!ha.gbank) return -1;
            CUDA_TRY(cudaMemsetAsync(ha.gbank, 0, sizeof(double) * (size_t)bank_stride * hc.n_bslot, stream));
            // single-bin halos give one bucket each; bins above BIN_SMEM records (k_sort_bins) one each
            const size_t max_bkt = (size_t)n_try + (size_t)(hc.rec_total / BIN_SMEM) + 16;
            Bucket* bkt_small = (Bucket*)h->get("h_bkt_small", sizeof(Bucket) * max_bkt);
            Bucket* bkt_big = (Bucket*)h->get("h_bkt_big", sizeof(Bucket) * max_bkt);
            Bucket* bkt_huge = (Bucket*)h->get("h_bkt_huge", sizeof(Bucket) * (size_t)(hc.rec_total / SB_CAP + 16));
            uint32_t* fine_cnt = (uint32_t*)h->get("h_fine_cnt", sizeof(uint32_t) * (size_t)(hc.n_fine + 1));
            uint32_t* fine_cur = (uint32_t*)h->get("h_fine_cur", sizeof(uint32_t) * (size_t)(hc.n_fine + 1));
            int64_t* fine_excl = (int64_t*)h->get("h_fine_excl", sizeof(int64_t) * (size_t)(hc.n_fine + 1));
            if (!bkt_small || !bkt_big || !bkt_huge || !fine_cnt || !fine_cur || !fine_excl) return -1;
            unsigned int* n_try_dev = &ctr->n_try;
            if (hc.n_multi > 0) {
                log.begin("fine_hist", stream);
                CUDA_TRY(cudaMemsetAsync(fine_cnt, 0, sizeof(uint32_t) * (hc.n_fine + 1), stream));
                CUDA_TRY(cudaMemsetAsync(fine_cur, 0, sizeof(uint32_t) * (hc.n_fine + 1), stream));
                LAUNCH(h, k_fine_hist_halo, sweep_grid, TB, 0, stream, v, ha, items, ctr, fine_cnt);
                if (soap_exclusive_scan_u32(h, fine_cnt, nullptr, fine_excl, hc.n_fine + 1, nullptr, stream)) return -1;
                LAUNCH(h, k_multi_offsets, grid_for(hc.n_multi, 128), 128, 0, stream, ha, multi_list, &ctr->n_multi,
                       fine_excl, ctr);
                log.end(stream);
            }
            LAUNCH(h, k_single_buckets, grid_for(n_try, 128), 128, 0, stream, ha, acc_list, &ctr->n_acc, ctr, bkt_small,
                   bkt_big);
            log.begin("collect", stream);
            LAUNCH(h, k_collect, sweep_grid, TB, 0, stream, v, ha, dc, items, fine_excl, fine_cur, ctr, recs,
                   item_minr, item_minfof);
            log.end(stream);
            log.begin("sort", stream);
            if (hc.n_multi > 0) {
                // bins first: they feed the bucket lists of the CTA-wide kernels below
                const unsigned int nb = (hc.n_fine + 15) / 16;  // 16 bins per CTA: two per warp
                LAUNCH(h, k_sort_bins, nb < (unsigned)(sm * 8) ? nb : (unsigned)(sm * 8), 256, 0, stream, fine_excl,
                       hc.n_fine, ctr, 1, recs, bkt_big, bkt_huge);
            }
            {
                unsigned int gs = (unsigned)(max_bkt < (size_t)(sm * 16) ? max_bkt : (size_t)(sm * 16));
                LAUNCH(h, (k_sort_bucket<SMALL_CAP, 128>), gs, 128, SMALL_CAP * sizeof(Rec), stream, bkt_small,
                       &ctr->n_bkt_small, recs);
                unsigned int gb = (unsigned)(max_bkt < (size_t)(sm * 3) ? max_bkt : (size_t)(sm * 3));
                LAUNCH(h, (k_sort_bucket<SB_CAP, 512>), gb, 512, SB_CAP * sizeof(Rec), stream, bkt_big,
                       &ctr->n_bkt_big, recs);
                if (hc.n_multi > 0)
                    LAUNCH(h, (k_sort_bucket<0, 512>), (unsigned)(hc.n_multi < 64u ? hc.n_multi : 64u), 512, 16, stream,
                           bkt_huge, &ctr->n_bkt_huge, recs);
            }
            log.end(stream);
            log.begin("scan_solve", stream);
            {
                // the four scan kernels work on disjoint halos (by record count): thread per halo, CTA per halo,
                // clusters of 8 and of 16 CTAs.  Forked onto side streams they overlap -- the giants' clusters keep a
                // few SMs busy for milliseconds while the rest of the GPU does everything else.
                if (h->side_init()) SOAP_FAIL("soap_process_halos: cannot create side streams");
                cudaStream_t sd[3];
                for (int i = 0; i < 3; i++) sd[i] = serial ? stream : h->side[i];
                CUDA_TRY(cudaEventRecord(h->ev_fork, stream));
                for (int i = 0; i < 3; i++) CUDA_TRY(cudaStreamWaitEvent(sd[i], h->ev_fork, 0));
                if (hc.n_huge > 0) {
                    // clusters of 16 CTAs (non-portable size): the largest halo is the critical path of this phase
                    unsigned int ncl = hc.n_huge < (unsigned)(sm / SCAN_CS_HUGE) ? hc.n_huge : (unsigned)(sm / SCAN_CS_HUGE);
                    if (cfg->dmo) {
                        CUDA_TRY(cudaFuncSetAttribute(k_scan_solve<2, SCAN_CS_HUGE>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
                        LAUNCH(h, (k_scan_solve<2, SCAN_CS_HUGE>), ncl * SCAN_CS_HUGE, SCAN_NT, 0, sd[0], ha, dc, huge_list,
                               &ctr->n_huge, recs, next, ctr, item_minr, item_minfof, 2);
                    } else {
                        CUDA_TRY(cudaFuncSetAttribute(k_scan_solve<8, SCAN_CS_HUGE>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
                        LAUNCH(h, (k_scan_solve<8, SCAN_CS_HUGE>), ncl * SCAN_CS_HUGE, SCAN_NT, 0, sd[0], ha, dc, huge_list,
                               &ctr->n_huge, recs, next, ctr, item_minr, item_minfof, 2);
                    }
                }
                if (hc.n_big > 0) {
                    // one cluster of SCAN_CS CTAs per large halo
                    unsigned int ncl = hc.n_big < (unsigned)(sm * 2 / SCAN_CS) ? hc.n_big : (unsigned)(sm * 2 / SCAN_CS);
                    if (cfg->dmo)
                        LAUNCH(h, (k_scan_solve<2, SCAN_CS>), ncl * SCAN_CS, SCAN_NT, 0, sd[1], ha, dc, big_list,
                               &ctr->n_big, recs, next, ctr, item_minr, item_minfof, 1);
                    else
                        LAUNCH(h, (k_scan_solve<8, SCAN_CS>), ncl * SCAN_CS, SCAN_NT, 0, sd[1], ha, dc, big_list,
                               &ctr->n_big, recs, next, ctr, item_minr, item_minfof, 1);
                }
                if (hc.n_try > 0) {
                    unsigned int g = hc.n_try < (unsigned)(sm * 8) ? hc.n_try : (unsigned)(sm * 8);
                    if (cfg->dmo)
                        LAUNCH(h, (k_scan_solve<2, 1>), g, SCAN_NT, 0, sd[2], ha, dc, try_list, n_try_dev, recs, next,
                               ctr, item_minr, item_minfof, 0);
                    else
                        LAUNCH(h, (k_scan_solve<8, 1>), g, SCAN_NT, 0, sd[2], ha, dc, try_list, n_try_dev, recs, next,
                               ctr, item_minr, item_minfof, 0);
                }
                // the smallest halos: a thread each when there are thousands of them (tiers off, projected apertures);
                // a few dozen stragglers would be one long sequential tail on an idle GPU: a CTA each instead
                if (hc.n_seq >= SEQ_MIN_LIST) {
                    if (soap_launch_solve_seq(c, dc, ha, seq_list, &ctr->n_seq, hc.n_seq, recs, next, &ctr->n_next, ctr,
                                              item_minr, item_minfof, 0, stream))
                        return -1;
                } else if (hc.n_seq > 0) {
                    if (cfg->dmo)
                        LAUNCH(h, (k_scan_solve<2, 1>), hc.n_seq, SCAN_NT, 0, stream, ha, dc, seq_list, &ctr->n_seq, recs, next,
                               ctr, item_minr, item_minfof, 3);
                    else
                        LAUNCH(h, (k_scan_solve<8, 1>), hc.n_seq, SCAN_NT, 0, stream, ha, dc, seq_list, &ctr->n_seq, recs, next,
                               ctr, item_minr, item_minfof, 3);
                }
                for (int i = 0; i < 3; i++) {
                    CUDA_TRY(cudaEventRecord(h->ev_join[i], sd[i]));
                    CUDA_TRY(cudaStreamWaitEvent(stream, h->ev_join[i], 0));
                }
            }
            log.end(stream);
            log.begin("moments", stream);
            if (soap_launch_moments(c, dc, ha, items, &ctr->n_items, hc.n_items, sweep_grid, stream)) return -1;
            if (soap_launch_rows(c, dc, ha, acc_list, &ctr->n_acc, n_try, stream)) return -1;
            if (soap_launch_projected(c, dc, ha, items, &ctr->n_items, hc.n_items, hc.n_mslot, sweep_grid, stream)) return -1;
            if (dc.n_pj > 0 && (dc.flags & PF_HMR)) {
                // projected half-mass radii: per axis bin by projected radius, sort the bins, scan
                // (timed inside the enclosing "moments" phase)
                PjPlan pl;
                pl.nfine = (uint32_t*)h->get("h_pj_nfine", sizeof(uint32_t) * (size_t)H);
                pl.fine_off = (uint32_t*)h->get("h_pj_fine_off", sizeof(uint32_t) * (size_t)H);
                pl.n_fine = (unsigned int*)h->get("h_pj_nf", sizeof(unsigned int) * 4);
                if (!pl.nfine || !pl.fine_off || !pl.n_fine) return -1;
                CUDA_TRY(cudaMemsetAsync(pl.n_fine, 0, sizeof(unsigned int) * 4, stream));
                LAUNCH(h, k_pj_plan, grid_for(n_try, 128), 128, 0, stream, ha, dc, acc_list, &ctr->n_acc, pl);
                unsigned int nfp = 0;
                CUDA_TRY(cudaMemcpyAsync(&nfp, pl.n_fine, sizeof(unsigned int), cudaMemcpyDeviceToHost, stream));
                CUDA_TRY(cudaStreamSynchronize(stream));
                if (nfp > 0) {
                    uint32_t* pcnt = (uint32_t*)h->get("h_pj_cnt", sizeof(uint32_t) * (size_t)(nfp + 1));
                    int64_t* pexcl = (int64_t*)h->get("h_pj_excl", sizeof(int64_t) * (size_t)(nfp + 1));
                    // bound particles of the accepted halos: at most the records of this rung
                    Rec* precs = (Rec*)h->get("h_pj_recs", sizeof(Rec) * (size_t)(hc.rec_total + 1));
                    if (!pcnt || !pexcl || !precs) return -1;
                    for (int ax = 0; ax < 3; ax++) {
                        CUDA_TRY(cudaMemsetAsync(pcnt, 0, sizeof(uint32_t) * (nfp + 1), stream));
                        LAUNCH(h, k_pj_bins<false>, sweep_grid, TB, 0, stream, v, ha, dc, items, &ctr->n_items, pl, ax,
                               pcnt, pexcl, precs);
                        if (soap_exclusive_scan_u32(h, pcnt, nullptr, pexcl, nfp + 1, nullptr, stream)) return -1;
                        CUDA_TRY(cudaMemsetAsync(pcnt, 0, sizeof(uint32_t) * (nfp + 1), stream));
                        LAUNCH(h, k_pj_bins<true>, sweep_grid, TB, 0, stream, v, ha, dc, items, &ctr->n_items, pl, ax,
                               pcnt, pexcl, precs);
                        CUDA_TRY(cudaMemsetAsync(&ctr->n_bkt_small, 0, 3 * sizeof(unsigned int), stream));
                        const unsigned int nb = (nfp + 15) / 16;
                        LAUNCH(h, k_sort_bins, nb < (unsigned)(sm * 8) ? nb : (unsigned)(sm * 8), 256, 0, stream, pexcl,
                               nfp, ctr, 0, precs, bkt_big, bkt_huge);
                        LAUNCH(h, (k_sort_bucket<SB_CAP, 512>), (unsigned)(sm * 3), 512, SB_CAP * sizeof(Rec), stream,
                               bkt_big, &ctr->n_bkt_big, precs);
                        LAUNCH(h, (k_sort_bucket<0, 512>), 64, 512, 16, stream, bkt_huge, &ctr->n_bkt_huge, precs);
                        LAUNCH(h, k_pj_scan, n_try < (unsigned)(sm * 4) ? n_try : (unsigned)(sm * 4), 256, 0, stream, ha, dc,
                               acc_list, &ctr->n_acc, pl, ax, pexcl, precs);
                    }
                }
            }
            if (soap_launch_kappa(c, dc, ha, items, &ctr->n_items, hc.n_items, acc_list, &ctr->n_acc, n_try, sweep_grid,
                                  stream))
                return -1;
            log.end(stream);
        }
        // next round's pending list
        CUDA_TRY(cudaMemcpyAsync(n_pend_dev, &ctr->n_next, sizeof(unsigned int), cudaMemcpyDeviceToDevice, stream));
        Counters hc2;
        CUDA_TRY(cudaMemcpyAsync(&hc2, ctr, sizeof(Counters), cudaMemcpyDeviceToHost, stream));
        CUDA_TRY(cudaStreamSynchronize(stream));
        total_pairs += hc2.pairs;
        total_mom_pairs += hc2.mom_pairs;
        n_pend = hc2.n_next;
        if (tiers_running) {  // what the tiers hand over joins the second round
            unsigned int n_late = 0;
            if (join_tiers(&n_late)) return -1;
            if (n_late > 0) {
                CUDA_TRY(cudaMemcpyAsync(next + n_pend, list2, n_late * sizeof(uint32_t), cudaMemcpyDeviceToDevice, stream));
                n_pend += n_late;
                CUDA_TRY(cudaMemcpyAsync(n_pend_dev, &n_pend, sizeof(unsigned int), cudaMemcpyHostToDevice, stream));
                CUDA_TRY(cudaStreamSynchronize(stream));
            }
        }
        uint32_t* t = pend; pend = next; next = t;
    }
    if (dc.flags & PF_ITER) {
        // iterative inertia tensors: repeated sweeps of the finished halos (iter.cu)
        log.begin("iter_tensors", stream);
        unsigned int n_fin = 0;
        if (soap_iter_list(h, ha, (int64_t)H, acc_list, &ctr->n_acc, &n_fin, stream)) return -1;
        if (n_fin > 0) {
            if (plan(acc_list, &ctr->n_acc, n_fin, 1, 1)) return -1;
            Counters pc;
            CUDA_TRY(cudaMemcpyAsync(&pc, ctr, sizeof(Counters), cudaMemcpyDeviceToHost, stream));
            CUDA_TRY(cudaStreamSynchronize(stream));
            if (soap_launch_iter_tensors(c, dc, ha, (int64_t)H, items, &ctr->n_items, pc.n_items, acc_list, &ctr->n_acc,
                                         n_fin, sweep_grid, stream))
                return -1;
        }
        log.end(stream);
    }
    if (soap_write_input_cols(h, ha, (int64_t)H, stream)) return -1;
    CUDA_TRY(cudaStreamSynchronize(stream));
    log.collect();
    c->last_pairs = (int64_t)total_pairs;
    c->last_candidates = (int64_t)total_cand;
    c->last_count_pairs = (int64_t)total_count_pairs;
    c->last_try_pairs = (int64_t)total_try_pairs;
    c->last_mom_pairs = (int64_t)total_mom_pairs;
    return 0;
}

}  // extern "C"
