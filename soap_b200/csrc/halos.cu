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
#include "halos.cuh"

#include <cooperative_groups.h>
namespace cg = cooperative_groups;

int soap_launch_moments(soap_chunk* c, const DevCfg& cfg, const HaloArrays& ha, const Item* items,
                        const unsigned int* n_items_dev, unsigned int n_items_host,
                        unsigned int n_mslot, unsigned int grid, cudaStream_t stream);
int soap_write_input_cols(soap_handle* h, const HaloArrays& ha, int64_t nh, cudaStream_t stream);

namespace {

constexpr int TB = SWEEP_NT;
constexpr int SB_CAP = 4096;     // records of one bucket sorted in shared memory (64 KB)
constexpr int SMALL_CAP = 512;   // small-bucket class (8 KB)
constexpr int FINE_TARGET = 128; // expected records per fine radial bin
constexpr uint32_t SCAN_BIG = 32768;  // halos with more records are scanned by a CTA cluster
constexpr int SCAN_CS = 8;            // CTAs per cluster for those

struct Bucket {
    unsigned long long start;
    uint32_t count;
    uint32_t halo;
};

// device counters of one round
struct Counters {
    unsigned int n_try, n_big, n_next, n_multi, n_fine;
    unsigned long long rec_single, rec_total;
    unsigned int n_bkt_small, n_bkt_big, n_bkt_huge, n_items;
    unsigned int n_mslot, items_overflow;
    unsigned long long pairs, candidates, count_pairs, mom_pairs;
};

// ------------------------------------------------------------------ k_init
__global__ void k_init(HaloArrays ha, int64_t nh, uint32_t* pend) {
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
    pend[h] = (uint32_t)h;
    double* row = ha.out + h * ha.ncol;
    for (int64_t c = 0; c < ha.ncol; c++) row[c] = 0.0;
}

// ------------------------------------------------------------- k_plan_items
// One thread per pending halo: count the candidates of its rows at the current
// radius and cut the candidate stream into work items of ITEM_CAND particles.
__global__ void __launch_bounds__(128) k_plan_items(ChunkView v, HaloArrays ha, const uint32_t* __restrict__ pend,
                                                    const unsigned int* __restrict__ n_pend,
                                                    Item* __restrict__ items, unsigned int items_cap,
                                                    Counters* ctr) {
    unsigned int it = blockIdx.x * blockDim.x + threadIdx.x;
    if (it >= *n_pend) return;
    const uint32_t h = pend[it];
    DimRanges rg[3];
    const double r = ha.cur_r[h];
    for (int d = 0; d < 3; d++)
        dim_ranges(ha.cofp[3 * h + d], r, v.L, v.pmin[d], v.pmax[d], v.cs[d], v.res, rg[d]);
    const RowIter ri = row_iter(rg);
    unsigned long long cand = 0;
    for (int row = 0; row < ri.nrows; row++) {
        uint32_t s0, s1;
        row_span(v, rg, ri, row, s0, s1);
        cand += s1 - s0;
    }
    if (cand > 0xfffffff0ull) cand = 0xfffffff0ull;
    uint32_t ni = (uint32_t)((cand + ITEM_CAND - 1) / ITEM_CAND);
    if (ni < 1) ni = 1;
    uint32_t base = atomicAdd(&ctr->n_items, ni);
    if (base + ni > items_cap) {
        // the host grows the item list to n_items and plans this rung again
        atomicExch(&ctr->items_overflow, 1u);
        return;
    }
    ha.item_base[h] = base;
    ha.n_items[h] = ni;
    ha.cnt[h] = 0;
    ha.msum[h] = 0.0;
    ha.rung_r[h] = r;
    ha.cursor[h] = 0;
    ha.items_done[h] = 0;
    ha.commit_lo[h] = ha.commit_hi[h] = ha.ndone[h];
    ha.mslot[h] = ni > 1 ? (int32_t)atomicAdd(&ctr->n_mslot, 1u) : -1;
    if (ni == 1) {
        Item im;
        im.halo = h; im.first = 0; im.count = (uint32_t)cand; im.row0 = 0;
        im.pos0 = 0; im.k = 0; im.pad0 = im.pad1 = 0;
        items[base] = im;
    } else {
        // second walk: the row in which each item's range starts
        unsigned long long pos = 0;
        uint32_t k = 0;
        for (int row = 0; row < ri.nrows && k < ni; row++) {
            uint32_t s0, s1;
            row_span(v, rg, ri, row, s0, s1);
            const unsigned long long end = pos + (s1 - s0);
            while (k < ni && (unsigned long long)k * ITEM_CAND < end) {
                Item im;
                im.halo = h;
                im.first = k * ITEM_CAND;
                const unsigned long long left = cand - (unsigned long long)k * ITEM_CAND;
                im.count = (uint32_t)(left < ITEM_CAND ? left : ITEM_CAND);
                im.row0 = (uint32_t)row;
                im.pos0 = (uint32_t)pos;
                im.k = k; im.pad0 = im.pad1 = 0;
                items[base + k] = im;
                k++;
            }
            pos = end;
        }
    }
    atomicAdd(&ctr->candidates, cand);
}

// ----------------------------------------------------------------- k_count
// periodic sphere count + enclosed mass of every pending halo at its current
// radius (halo_tasks.py:84-97; shared_mesh.py:122-200)
__global__ void __launch_bounds__(TB) k_count(ChunkView v, HaloArrays ha, const Item* __restrict__ items,
                                              Counters* ctr) {
    __shared__ SweepShared S;
    __shared__ unsigned long long s_cnt[TB / 32];
    __shared__ double s_m[TB / 32];
    const unsigned int n_items = ctr->n_items;
    for (unsigned int it = blockIdx.x; it < n_items; it += gridDim.x) {
        const Item im = items[it];
        const uint32_t h = im.halo;
        const double cx = ha.cofp[3 * h], cy = ha.cofp[3 * h + 1], cz = ha.cofp[3 * h + 2];
        const double r = ha.cur_r[h];
        const double r2max = __dmul_rn(r, r);
        const double halfL = 0.5 * v.L, L = v.L;
        unsigned long long cnt = 0;
        double msum = 0.0;
        sweep_item(v, S, cx, cy, cz, r, im, [&](uint32_t t, bool ok) {
            if (ok) {
                double r2 = periodic_r2(v.px[t], v.py[t], v.pz[t], cx, cy, cz, L, halfL);
                if (r2 <= r2max) {
                    cnt++;
                    msum += (double)v.mass[t];
                }
            }
        });
        const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
        cnt = warp_sum_u64(cnt);
        msum = warp_sum(msum);
        if (lane == 0) { s_cnt[wid] = cnt; s_m[wid] = msum; }
        __syncthreads();
        if (threadIdx.x == 0) {
            unsigned long long c = 0;
            double m = 0.0;
            for (int w = 0; w < TB / 32; w++) { c += s_cnt[w]; m += s_m[w]; }
            if (c) atomicAdd(&ha.cnt[h], (unsigned int)c);
            if (m != 0.0) atomicAdd(&ha.msum[h], m);
            atomicAdd(&ctr->count_pairs, c);
        }
        __syncthreads();
    }
}

// ------------------------------------------------------------------ ladder
// halo_tasks.py:166-187 and :390-402.  Returns true if the halo stays pending.
__device__ inline bool ladder_step(const HaloArrays& ha, uint32_t h, double required) {
    const double search_radius = ha.sr_in[h], read_radius = ha.rr_in[h];
    double cur = ha.cur_r[h];
    double* row = ha.out + (int64_t)h * ha.ncol;
    if (required > read_radius || cur >= read_radius) {
        double sr = required > read_radius ? fmax(search_radius, required) : fmax(search_radius, cur);
        row[4] = sr;                                        // halo_tasks.py:173,179
        row[5] = fmax(__dmul_rn(read_radius, 1.5), sr);     // halo_tasks.py:393-396
        ha.status[h] = SOAP_HALO_RADIUS_TOO_SMALL;
        ha.state[h] = ST_DONE_FAIL;
        return false;
    }
    cur = fmin(__dmul_rn(cur, 1.2), read_radius);  // halo_tasks.py:184-186
    cur = fmax(cur, required);                      // halo_tasks.py:187
    ha.cur_r[h] = cur;
    ha.state[h] = ST_PENDING;
    return true;
}

__global__ void __launch_bounds__(128) k_gate(HaloArrays ha, DevCfg cfg, const uint32_t* __restrict__ pend,
                                              const unsigned int* __restrict__ n_pend,
                                              uint32_t* __restrict__ try_list,
                                              uint32_t* __restrict__ big_list,
                                              uint32_t* __restrict__ multi_list,
                                              uint32_t* __restrict__ next, Counters* ctr) {
    unsigned int it = blockIdx.x * blockDim.x + threadIdx.x;
    if (it >= *n_pend) return;
    const uint32_t h = pend[it];
    ha.nloop[h] += 1;  // halo_tasks.py:75
    const double r = ha.cur_r[h];
    // halo_tasks.py:97
    const double density = ha.msum[h] / (4.0 / 3.0 * SOAP_PI * (r * r * r));
    const bool has_target = ha.central[h] == 1 && cfg.target_density > 0.0;  // halo_tasks.py:381
    if (!has_target || density <= cfg.target_density) {  // halo_tasks.py:103
        const uint32_t cnt = ha.cnt[h];
        if (cnt > SCAN_BIG) big_list[atomicAdd(&ctr->n_big, 1u)] = h;  // scanned by a CTA cluster
        else try_list[atomicAdd(&ctr->n_try, 1u)] = h;
        ha.state[h] = ST_TRY;
        if (cnt <= SB_CAP) {
            ha.rec_off[h] = atomicAdd(&ctr->rec_single, (unsigned long long)cnt);
            ha.nfine[h] = 0;
        } else {
            uint32_t nf = cnt / FINE_TARGET;
            if (nf < 64) nf = 64;
            if (nf > 65536) nf = 65536;
            ha.nfine[h] = nf;
            ha.fine_off[h] = atomicAdd(&ctr->n_fine, nf);
            multi_list[atomicAdd(&ctr->n_multi, 1u)] = h;
        }
        atomicAdd(&ctr->rec_total, (unsigned long long)cnt);
    } else {
        if (ladder_step(ha, h, 0.0)) next[atomicAdd(&ctr->n_next, 1u)] = h;
    }
}

// ----------------------------------------------------------- record helper
struct Part {
    double x, y, z, r;
};
__device__ __forceinline__ Part rel_part(const ChunkView& v, uint32_t t, double cx, double cy,
                                         double cz, double halfL) {
    Part p;
    p.x = rewrap_rel(v.px[t], cx, v.L, halfL);
    p.y = rewrap_rel(v.py[t], cy, v.L, halfL);
    p.z = rewrap_rel(v.pz[t], cz, v.L, halfL);
    p.r = radius3(p.x, p.y, p.z);
    return p;
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

// Group a multi-bucket halo's fine radial bins into sort buckets: bins whose
// first record falls into the same BKT_SPAN-wide window of the halo's record
// range form one bucket (one CTA per halo, one thread per bin; a bin starts a
// bucket if its window differs from its predecessor's).  A bucket holds at
// most BKT_SPAN - 1 records plus its last bin.
constexpr uint32_t BKT_SPAN = SB_CAP / 2;
__global__ void __launch_bounds__(256) k_build_buckets(HaloArrays ha, const uint32_t* __restrict__ multi_list,
                                                       const int64_t* __restrict__ fine_excl, Counters* ctr,
                                                       Bucket* __restrict__ bkt_small, Bucket* __restrict__ bkt_big,
                                                       Bucket* __restrict__ bkt_huge) {
    const uint32_t h = multi_list[blockIdx.x];
    const uint32_t nf = ha.nfine[h], fo = ha.fine_off[h];
    const unsigned long long base = ctr->rec_single;  // multi region follows the single region
    const unsigned long long e0 = (unsigned long long)fine_excl[fo];
    if (threadIdx.x == 0) ha.rec_off[h] = base + e0;
    const unsigned long long e_end = e0 + ha.cnt[h];
    for (uint32_t f = threadIdx.x; f < nf; f += blockDim.x) {
        const unsigned long long ef = (unsigned long long)fine_excl[fo + f];
        const unsigned long long gf = (ef - e0) / BKT_SPAN;
        if (f > 0 && ((unsigned long long)fine_excl[fo + f - 1] - e0) / BKT_SPAN == gf) continue;
        // bucket start: find the first later bin in another window
        uint32_t f2 = f + 1;
        while (f2 < nf && ((unsigned long long)fine_excl[fo + f2] - e0) / BKT_SPAN == gf) f2++;
        const unsigned long long e2 = f2 < nf ? (unsigned long long)fine_excl[fo + f2] : e_end;
        const unsigned long long c = e2 - ef;
        if (c == 0) continue;
        Bucket b;
        b.start = base + ef;
        b.count = (uint32_t)c;
        b.halo = h;
        if (c <= SMALL_CAP) bkt_small[atomicAdd(&ctr->n_bkt_small, 1u)] = b;
        else if (c <= SB_CAP) bkt_big[atomicAdd(&ctr->n_bkt_big, 1u)] = b;
        else bkt_huge[atomicAdd(&ctr->n_bkt_huge, 1u)] = b;
    }
}

// single-bucket halos -> bucket lists (thread per try halo)
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
        const int32_t hidx = (int32_t)ha.index[h];
        Rec* out = recs + (nf ? ctr->rec_single : ha.rec_off[h]);
        const int64_t* fex = fine_excl + (nf ? ha.fine_off[h] : 0);
        uint32_t* fcur = fine_cursor + (nf ? ha.fine_off[h] : 0);
        unsigned int* cursor = &ha.cursor[h];
        const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
        unsigned long long minr = ~0ull;
        int minfof = -1;
        const bool dmo = cfg.dmo != 0;
        sweep_item(v, S, cx, cy, cz, R, im, [&](uint32_t t, bool ok) {
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
                if (in) out[base + __popc(bal & ((1u << lane) - 1u))] = rec;
            } else if (in) {
                unsigned slot = atomicAdd(&fcur[fb], 1u);
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
constexpr int SCAN_NT = 256;
constexpr int SCAN_K = 4;
constexpr int SCAN_TILE = SCAN_NT * SCAN_K;
constexpr uint32_t NONE = 0xffffffffu;

// first-index targets of the scan passes
enum {
    T_SO = 0,                                    // first record at or below each SO density
    T_NONNEG = T_SO + SOAP_MAX_SO,               // first non-negative cumulative mass
    T_SUBHMR = T_NONNEG + 1,                     // bound half-mass crossings (tot, gas, dm, star, baryon)
    T_APEDGE = T_SUBHMR + 5,                     // first record beyond each aperture
    T_APHMR = T_APEDGE + SOAP_MAX_APERTURES,     // aperture half-mass crossings [a][g]
    T_DMOUT = T_APHMR + 4 * SOAP_MAX_APERTURES,  // first dark matter particle outside each SO
    T_COUNT = T_DMOUT + SOAP_MAX_SO
};

template <int NCH>
struct ScanShared {
    double carry[NCH];
    uint32_t carryc[NCH];
    double wsum[SCAN_NT / 32][NCH];
    uint32_t wcnt[SCAN_NT / 32][NCH];
    // pass A results
    double tot[NCH], rmaxc[NCH];
    uint32_t cnt[NCH], cnt0[NCH];
    uint32_t n_zero;
    // first-index targets (T_*) and the values captured at them
    uint32_t tidx[T_COUNT];
    double tcap[T_COUNT][NCH < 3 ? 3 : NCH];
    // cluster merge scratch (rank 0)
    uint32_t m_idx[T_COUNT];
    uint8_t m_owner[T_COUNT];
    // local (this CTA's tile range) pass A results; tot/cnt/... above are the halo's
    double l_tot[NCH], l_rmaxc[NCH];
    uint32_t l_cnt[NCH], l_cnt0[NCH], l_n_zero;
    double carry_in[NCH];
    uint32_t carryc_in[NCH];
    // published block argmax results: 0 unsoftened, 1 softened subhalo Vmax, 2.. SO Vmax
    double pub_v[2 + SOAP_MAX_SO], pub_r[2 + SOAP_MAX_SO];
    uint32_t pub_i[2 + SOAP_MAX_SO];
    // outcome of the solve (rank 0), read by the other CTAs of the cluster
    int fail_;  // 0 ok, 1 retry, >=2 fatal status
    double required_;
    double so_r_[SOAP_MAX_SO];
    int commit_lo_, commit_hi_;
    double ap_thr[SOAP_MAX_APERTURES][4];
    // argmax block reduce
    double am_v[SCAN_NT / 32], am_r[SCAN_NT / 32];
    uint32_t am_i[SCAN_NT / 32];
};

// class of a record: type index * 2 + bound (DMO: type index 0)
template <int NCH>
__device__ __forceinline__ int rec_class(uint32_t flags) {
    return NCH == 2 ? (int)((flags >> 2) & 1u) : (int)(((flags & 3u) << 1) | ((flags >> 2) & 1u));
}
// group g (0 gas, 1 dm, 2 star, 3 baryon) membership of type code tc
__device__ __forceinline__ bool in_group(int g, uint32_t tc) {
    return g == 0 ? tc == 0 : (g == 1 ? tc == 1 : (g == 2 ? tc == 2 : (tc == 0 || tc == 2)));
}
// sum of the scan channels of group g, bound-only (b=1) or all (b=0)
template <int NCH>
__device__ __forceinline__ double group_sum(const double (&c)[NCH], int g, bool bound_only) {
    if (NCH == 2) {
        // DMO: only dark matter exists
        if (g == 0 || g == 2 || g == 3) return 0.0;
        return bound_only ? c[1] : c[0] + c[1];
    }
    double s = 0.0;
#pragma unroll
    for (int t = 0; t < 4; t++)
        if (in_group(g, (uint32_t)t)) s += bound_only ? c[2 * t + 1] : (c[2 * t] + c[2 * t + 1]);
    return s;
}
template <int NCH>
__device__ __forceinline__ double bound_sum(const double (&c)[NCH]) {
    double s = 0.0;
#pragma unroll
    for (int k = 1; k < NCH; k += 2) s += c[k];
    return s;
}
template <int NCH>
__device__ __forceinline__ double all_sum(const double (&c)[NCH]) {
    double s = 0.0;
#pragma unroll
    for (int k = 0; k < NCH; k++) s += c[k];
    return s;
}
template <int NCH>
__device__ __forceinline__ uint32_t bound_cnt(const uint32_t (&c)[NCH]) {
    uint32_t s = 0;
#pragma unroll
    for (int k = 1; k < NCH; k += 2) s += c[k];
    return s;
}
template <int NCH>
__device__ __forceinline__ uint32_t all_cnt(const uint32_t (&c)[NCH]) {
    uint32_t s = 0;
#pragma unroll
    for (int k = 0; k < NCH; k++) s += c[k];
    return s;
}

// SO cumulative mass at a record: float64 cumsum rounded to float32, neutrino
// background added in place in float32 (SO_properties.py:400-406)
__device__ __forceinline__ float so_cm32(double cum, double r, double nu) {
    float c = (float)cum;
    double r3 = r * r * r;
    return (float)((double)c + nu * 4.0 / 3.0 * SOAP_PI * r3);
}
__device__ __forceinline__ double so_density(float cm, double r) {
    return (double)cm / (4.0 / 3.0 * SOAP_PI * (r * r * r));  // SO_properties.py:420
}

// scipy.optimize.brentq (scipy/optimize/Zeros/brentq.c) with scipy's defaults
// xtol=2e-12, rtol=8.881784197001252e-16, maxiter=100, on the reference's
// cumulative_mass_intersection (SO_properties.py:50-77,206-210).
__device__ inline double cmi(double u, double rho_dim, double slope_dim) {
    return 4.0 * SOAP_PI / 3.0 * rho_dim * (u * u * u) - slope_dim * u + slope_dim - 1.0;
}
__device__ inline int brentq_dev(double xa, double xb, double rho_dim, double slope_dim, double* root) {
    const double xtol = 2e-12, rtol = 8.881784197001252e-16;
    double xpre = xa, xcur = xb, xblk = 0., fpre, fcur, fblk = 0., spre = 0., scur = 0., sbis;
    double delta, stry, dpre, dblk;
    fpre = cmi(xpre, rho_dim, slope_dim);
    fcur = cmi(xcur, rho_dim, slope_dim);
    if (fpre == 0) { *root = xpre; return 0; }
    if (fcur == 0) { *root = xcur; return 0; }
    if (signbit(fpre) == signbit(fcur)) return -1;  // ValueError in scipy
    for (int i = 0; i < 100; i++) {
        if (fpre != 0 && fcur != 0 && (signbit(fpre) != signbit(fcur))) {
            xblk = xpre; fblk = fpre; spre = scur = xcur - xpre;
        }
        if (fabs(fblk) < fabs(fcur)) {
            xpre = xcur; xcur = xblk; xblk = xpre;
            fpre = fcur; fcur = fblk; fblk = fpre;
        }
        delta = (xtol + rtol * fabs(xcur)) / 2;
        sbis = (xblk - xcur) / 2;
        if (fcur == 0 || fabs(sbis) < delta) { *root = xcur; return 0; }
        if (fabs(spre) > delta && fabs(fcur) < fabs(fpre)) {
            if (xpre == xblk) {
                stry = -fcur * (xcur - xpre) / (fcur - fpre);  // secant
            } else {
                dpre = (fpre - fcur) / (xpre - xcur);          // inverse quadratic
                dblk = (fblk - fcur) / (xblk - xcur);
                stry = -fcur * (fblk * dblk - fpre * dpre) / (dblk * dpre * (fblk - fpre));
            }
            if (2 * fabs(stry) < fmin(fabs(spre), 3 * fabs(sbis) - delta)) {
                spre = scur; scur = stry;
            } else {
                spre = sbis; scur = sbis;
            }
        } else {
            spre = sbis; scur = sbis;
        }
        xpre = xcur; fpre = fcur;
        if (fabs(scur) > delta) xcur += scur;
        else xcur += (sbis > 0 ? delta : -delta);
        fcur = cmi(xcur, rho_dim, slope_dim);
    }
    *root = xcur;  // scipy raises RuntimeError (convergence); keep the iterate
    return 0;
}

// block exclusive scan of NCH double + NCH uint32 channels (thread totals)
template <int NCH>
__device__ __forceinline__ void block_scan_channels(double (&v)[NCH], uint32_t (&c)[NCH],
                                                    ScanShared<NCH>& S) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    double inc[NCH];
    uint32_t incc[NCH];
#pragma unroll
    for (int ch = 0; ch < NCH; ch++) {
        double x = v[ch];
        uint32_t y = c[ch];
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            double tx = __shfl_up_sync(0xffffffffu, x, o);
            uint32_t ty = __shfl_up_sync(0xffffffffu, y, o);
            if (lane >= o) { x += tx; y += ty; }
        }
        inc[ch] = x;
        incc[ch] = y;
        if (lane == 31) { S.wsum[wid][ch] = x; S.wcnt[wid][ch] = y; }
    }
    __syncthreads();
#pragma unroll
    for (int ch = 0; ch < NCH; ch++) {
        double base = S.carry[ch];
        uint32_t basec = S.carryc[ch];
        for (int w = 0; w < wid; w++) { base += S.wsum[w][ch]; basec += S.wcnt[w][ch]; }
        double ex = base + (inc[ch] - v[ch]);
        uint32_t exc = basec + (incc[ch] - c[ch]);
        v[ch] = ex;
        c[ch] = exc;
    }
    __syncthreads();
    if (threadIdx.x == SCAN_NT - 1) {
        // carry for the next tile = inclusive total of the last thread
#pragma unroll
        for (int ch = 0; ch < NCH; ch++) {
            double base = S.carry[ch];
            uint32_t basec = S.carryc[ch];
            for (int w = 0; w < SCAN_NT / 32; w++) { base += S.wsum[w][ch]; basec += S.wcnt[w][ch]; }
            S.carry[ch] = base;
            S.carryc[ch] = basec;
        }
    }
    // callers sync before the next tile touches carry / wsum
}

struct ArgMax {
    double v, r;
    uint32_t i;
    __device__ __forceinline__ void init() { v = -1.0; r = 0.0; i = NONE; }
    __device__ __forceinline__ void offer(double v_, double r_, uint32_t i_) {
        if (v_ > v || (v_ == v && i_ < i)) { v = v_; r = r_; i = i_; }
    }
};
template <int NCH>
__device__ inline void argmax_reduce(ArgMax& a, ScanShared<NCH>& S) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        double ov = __shfl_xor_sync(0xffffffffu, a.v, o);
        double orr = __shfl_xor_sync(0xffffffffu, a.r, o);
        uint32_t oi = __shfl_xor_sync(0xffffffffu, a.i, o);
        a.offer(ov, orr, oi);
    }
    __syncthreads();
    if (lane == 0) { S.am_v[wid] = a.v; S.am_r[wid] = a.r; S.am_i[wid] = a.i; }
    __syncthreads();
    a.v = S.am_v[0]; a.r = S.am_r[0]; a.i = S.am_i[0];
    for (int w = 1; w < SCAN_NT / 32; w++) a.offer(S.am_v[w], S.am_r[w], S.am_i[w]);
    __syncthreads();
}

// One CTA (CS == 1) or one cluster of CS CTAs (halos above SCAN_BIG records) per
// halo of the list.  Three streaming passes over the halo's radially sorted
// records.  In a cluster every CTA owns a contiguous range of tiles: pass A
// totals give each CTA its carry-in, the first-index targets and argmax
// candidates found locally in passes B and C are merged by rank 0 through
// distributed shared memory, and rank 0 alone runs the solve.
template <int NCH, int CS>
__global__ void __cluster_dims__(CS, 1, 1) __launch_bounds__(SCAN_NT)
    k_scan_solve(HaloArrays ha, DevCfg cfg, const uint32_t* __restrict__ try_list,
                 const unsigned int* __restrict__ n_try, const Rec* __restrict__ recs,
                 uint32_t* __restrict__ next, Counters* ctr,
                 const unsigned long long* __restrict__ item_minr,
                 const int32_t* __restrict__ item_minfof) {
    __shared__ ScanShared<NCH> S;
    cg::cluster_group cluster = cg::this_cluster();
    const unsigned int crank = CS > 1 ? cluster.block_rank() : 0u;
    const unsigned int cid = blockIdx.x / CS, ncl = gridDim.x / CS;
    auto csync = [&]() { if (CS > 1) cluster.sync(); else __syncthreads(); };
    auto peer = [&](unsigned int rk) -> ScanShared<NCH>* { return CS > 1 ? cluster.map_shared_rank(&S, rk) : &S; };
    // rank 0: merge the first-index targets found by the other CTAs into S
    auto merge_targets = [&]() {
        if (CS > 1 && crank == 0) {
            constexpr int CW = NCH < 3 ? 3 : NCH;
            for (int j = threadIdx.x; j < T_COUNT; j += SCAN_NT) {
                uint32_t best = S.tidx[j];
                unsigned int owner = 0;
                for (unsigned int rk = 1; rk < CS; rk++) {
                    const uint32_t o = peer(rk)->tidx[j];
                    if (o < best) { best = o; owner = rk; }
                }
                S.m_idx[j] = best;
                S.m_owner[j] = (uint8_t)owner;
            }
            __syncthreads();
            for (int e = threadIdx.x; e < T_COUNT * CW; e += SCAN_NT) {
                const int j = e / CW, k = e % CW;
                if (S.m_owner[j]) S.tcap[j][k] = peer(S.m_owner[j])->tcap[j][k];
            }
            for (int j = threadIdx.x; j < T_COUNT; j += SCAN_NT) S.tidx[j] = S.m_idx[j];
            __syncthreads();
        }
    };
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    for (unsigned int it = cid; it < *n_try; it += ncl) {
        const uint32_t h = try_list[it];
        const uint32_t n = ha.cnt[h];
        const Rec* R = recs + ha.rec_off[h];
        ScanRes* sr = ha.sres + h;
        const bool central = ha.central[h] == 1;
        const int n_so = central ? cfg.n_so : 0;  // SO_properties.py:3627
        const int n_ap = cfg.n_ap;
        const bool want_hmr = (cfg.flags & PF_HMR) != 0;
        // this CTA's tiles
        const uint32_t ntile = (n + SCAN_TILE - 1) / SCAN_TILE;
        const uint32_t tiles_per = (ntile + CS - 1) / CS;
        const uint32_t t_lo = crank * tiles_per < ntile ? crank * tiles_per : ntile;
        const uint32_t t_hi = t_lo + tiles_per < ntile ? t_lo + tiles_per : ntile;
        const uint32_t i_lo = t_lo * SCAN_TILE;
        const uint32_t i_hi = (unsigned long long)t_hi * SCAN_TILE < n ? t_hi * SCAN_TILE : n;
        __syncthreads();
        // ------------------------------------------------------------ pass A
        if (threadIdx.x < NCH) {
            S.l_tot[threadIdx.x] = 0.0; S.l_rmaxc[threadIdx.x] = 0.0;
            S.l_cnt[threadIdx.x] = 0; S.l_cnt0[threadIdx.x] = 0;
        }
        if (threadIdx.x == 0) S.l_n_zero = 0;
        __syncthreads();
        {
            double tot[NCH], rmx[NCH];
            uint32_t cnt[NCH], cnt0[NCH], nz = 0;
#pragma unroll
            for (int ch = 0; ch < NCH; ch++) { tot[ch] = 0.0; rmx[ch] = 0.0; cnt[ch] = 0; cnt0[ch] = 0; }
            for (uint32_t i = i_lo + threadIdx.x; i < i_hi; i += SCAN_NT) {
                Rec rc = R[i];
                double r = __longlong_as_double((long long)rc.rbits);
                int c = rec_class<NCH>(rc.flags);
                nz += (r == 0.0);
#pragma unroll
                for (int ch = 0; ch < NCH; ch++)
                    if (c == ch) {
                        tot[ch] += (double)rc.m;
                        cnt[ch]++;
                        cnt0[ch] += (r <= 1e-8);
                        rmx[ch] = fmax(rmx[ch], r);
                    }
            }
#pragma unroll
            for (int ch = 0; ch < NCH; ch++) {
                double t = warp_sum(tot[ch]);
                double m = warp_max(rmx[ch]);
                uint32_t a = (uint32_t)warp_sum_u64(cnt[ch]);
                uint32_t b = (uint32_t)warp_sum_u64(cnt0[ch]);
                if (lane == 0) {
                    atomicAdd(&S.l_tot[ch], t);
                    atomicAdd(&S.l_cnt[ch], a);
                    atomicAdd(&S.l_cnt0[ch], b);
                    // non-negative doubles order like their bit patterns
                    atomicMax((unsigned long long*)&S.l_rmaxc[ch], (unsigned long long)__double_as_longlong(m));
                }
            }
            nz = (uint32_t)warp_sum_u64(nz);
            if (lane == 0) atomicAdd(&S.l_n_zero, nz);
        }
        csync();
        // halo totals and this CTA's carry-in (sum over the lower ranks)
        if (threadIdx.x < NCH) {
            const int ch = threadIdx.x;
            double tot = 0.0, rmx = 0.0, cin = 0.0;
            uint32_t cnt = 0, cnt0 = 0, ccin = 0;
            for (unsigned int rk = 0; rk < CS; rk++) {
                const ScanShared<NCH>* P = peer(rk);
                const double t = P->l_tot[ch];
                const uint32_t c = P->l_cnt[ch];
                if (rk < crank) { cin += t; ccin += c; }
                tot += t; cnt += c; cnt0 += P->l_cnt0[ch];
                rmx = fmax(rmx, P->l_rmaxc[ch]);
            }
            S.tot[ch] = tot; S.cnt[ch] = cnt; S.cnt0[ch] = cnt0; S.rmaxc[ch] = rmx;
            S.carry_in[ch] = cin; S.carryc_in[ch] = ccin;
        }
        if (threadIdx.x == NCH) {
            uint32_t nz = 0;
            for (unsigned int rk = 0; rk < CS; rk++) nz += peer(rk)->l_n_zero;
            S.n_zero = nz;
        }
        __syncthreads();
        // bound totals
        double Mb_g[5];  // tot, gas, dm, star, baryon
        uint32_t NB = 0, NB0 = 0;
        {
            double t[NCH];
#pragma unroll
            for (int ch = 0; ch < NCH; ch++) t[ch] = S.tot[ch];
            Mb_g[0] = bound_sum<NCH>(t);
            Mb_g[1] = group_sum<NCH>(t, 0, true);
            Mb_g[2] = group_sum<NCH>(t, 1, true);
            Mb_g[3] = group_sum<NCH>(t, 2, true);
            Mb_g[4] = group_sum<NCH>(t, 3, true);
#pragma unroll
            for (int ch = 1; ch < NCH; ch += 2) { NB += S.cnt[ch]; NB0 += S.cnt0[ch]; }
        }
        // SO_properties.py:416: nskip = max(1, argmax(r > 0))
        uint32_t nskip_so = S.n_zero >= n ? 1u : (S.n_zero > 1u ? S.n_zero : 1u);
        // kinematic_properties.py:584-586 on the bound subset
        const uint32_t fnc_u = NB0 < NB ? NB0 : 0u;
        const uint32_t nskip_u = fnc_u > 1u ? fnc_u : 1u;
        // softened radii: isclose(max(soft, r), 0) needs soft <= 1e-8
        double min_soft = fmin(fmin(cfg.soft[0], cfg.soft[1]), fmin(cfg.soft[2], cfg.soft[3]));
        const uint32_t nskip_s = (min_soft <= 1e-8) ? fnc_u : 0u;

        // ------------------------------------------------------------ pass B
        if (threadIdx.x < NCH) { S.carry[threadIdx.x] = S.carry_in[threadIdx.x]; S.carryc[threadIdx.x] = S.carryc_in[threadIdx.x]; }
        for (int j = threadIdx.x; j < T_COUNT; j += SCAN_NT) S.tidx[j] = NONE;
        __syncthreads();
        ArgMax amU, amS;
        amU.init();
        amS.init();
        for (uint32_t tile = t_lo; tile < t_hi; tile++) {
            const uint32_t i0 = tile * SCAN_TILE + threadIdx.x * SCAN_K;
            Rec rc[SCAN_K];
            double base[NCH];
            uint32_t basec[NCH];
#pragma unroll
            for (int ch = 0; ch < NCH; ch++) { base[ch] = 0.0; basec[ch] = 0; }
#pragma unroll
            for (int k = 0; k < SCAN_K; k++) {
                if (i0 + k < n) {
                    rc[k] = R[i0 + k];
                    int c = rec_class<NCH>(rc[k].flags);
#pragma unroll
                    for (int ch = 0; ch < NCH; ch++)
                        if (c == ch) { base[ch] += (double)rc[k].m; basec[ch]++; }
                } else {
                    rc[k].rbits = 0; rc[k].m = 0.f; rc[k].flags = 0;
                }
            }
            block_scan_channels<NCH>(base, basec, S);  // base = exclusive prefix at i0
            double b0[NCH];
            uint32_t bc0[NCH];
#pragma unroll
            for (int ch = 0; ch < NCH; ch++) { b0[ch] = base[ch]; bc0[ch] = basec[ch]; }
            // detection sweep
#pragma unroll
            for (int k = 0; k < SCAN_K; k++) {
                const uint32_t i = i0 + k;
                if (i >= n) break;
                const double r = __longlong_as_double((long long)rc[k].rbits);
                const double m = (double)rc[k].m;
                const int c = rec_class<NCH>(rc[k].flags);
                const uint32_t tc = NCH == 2 ? 1u : (rc[k].flags & 3u);
                const bool bound = (rc[k].flags & 4u) != 0;
                const double call_ex = all_sum<NCH>(base);
                const double cb_ex = bound_sum<NCH>(base);
                const uint32_t posb = bound_cnt<NCH>(basec);
#pragma unroll
                for (int ch = 0; ch < NCH; ch++)
                    if (c == ch) { base[ch] += m; basec[ch]++; }
                const double call_in = call_ex + m;
                // SO first-below (SO_properties.py:140-147)
                if (i >= nskip_so) {
                    float cm = so_cm32(call_in, r, cfg.nu);
                    double dens = so_density(cm, r);
                    for (int q = 0; q < n_so; q++)
                        if (!(dens > cfg.so_rho[q]) && i < S.tidx[T_SO + (q)]) atomicMin(&S.tidx[T_SO + (q)], i);
                    if (!(cm < 0.f) && i < S.tidx[T_NONNEG]) atomicMin(&S.tidx[T_NONNEG], i);
                }
                if (bound) {
                    const double cb_in = cb_ex + m;
                    // Vmax of the bound subhalo (subhalo_properties.py:982-1045)
                    if (cfg.do_sub) {
                        if (posb >= nskip_u && r > 0.0) amU.offer(cb_in / r, r, i);
                        double rs = fmax(cfg.soft[tc], r);
                        if (posb >= nskip_s && rs > 0.0) amS.offer(cb_in / rs, rs, i);
                        // half-mass radii (half_mass_radius.py:63)
                        if (want_hmr || true) {
                            if (cb_in >= 0.5 * Mb_g[0] && i < S.tidx[T_SUBHMR + (0)]) atomicMin(&S.tidx[T_SUBHMR + (0)], i);
                        }
                        if (want_hmr) {
#pragma unroll
                            for (int g = 0; g < 4; g++)
                                if (in_group(g, tc)) {
                                    double w = group_sum<NCH>(base, g, true);
                                    if (w >= 0.5 * Mb_g[1 + g] && i < S.tidx[T_SUBHMR + (1 + g)])
                                        atomicMin(&S.tidx[T_SUBHMR + (1 + g)], i);
                                }
                        }
                    }
                }
                // first record beyond each aperture radius (aperture_properties.py:310)
                for (int a = 0; a < n_ap; a++)
                    if (want_hmr && r > cfg.ap_r[a] && i < S.tidx[T_APEDGE + (a)]) atomicMin(&S.tidx[T_APEDGE + (a)], i);
            }
            __syncthreads();
            // capture sweep: the owner of a newly found index re-derives its values
            {
                double bb[NCH];
#pragma unroll
                for (int ch = 0; ch < NCH; ch++) bb[ch] = b0[ch];
#pragma unroll
                for (int k = 0; k < SCAN_K; k++) {
                    const uint32_t i = i0 + k;
                    if (i >= n) break;
                    const double r = __longlong_as_double((long long)rc[k].rbits);
                    const double m = (double)rc[k].m;
                    const int c = rec_class<NCH>(rc[k].flags);
                    const uint32_t tc = NCH == 2 ? 1u : (rc[k].flags & 3u);
                    const double call_ex = all_sum<NCH>(bb);
                    double ex[NCH];
#pragma unroll
                    for (int ch = 0; ch < NCH; ch++) ex[ch] = bb[ch];
#pragma unroll
                    for (int ch = 0; ch < NCH; ch++)
                        if (c == ch) bb[ch] += m;
                    for (int q = 0; q < n_so; q++)
                        if (S.tidx[T_SO + (q)] == i) {
                            S.tcap[T_SO + (q)][0] = r; S.tcap[T_SO + (q)][1] = call_ex + m; S.tcap[T_SO + (q)][2] = call_ex;
                        }
                    if (S.tidx[T_NONNEG] == i) {
                        S.tcap[T_NONNEG][0] = r;
                        S.tcap[T_NONNEG][1] = (double)so_cm32(call_ex + m, r, cfg.nu);
                    }
                    if (cfg.do_sub) {
                        if (S.tidx[T_SUBHMR + (0)] == i) {
                            S.tcap[T_SUBHMR + (0)][0] = r;
                            S.tcap[T_SUBHMR + (0)][1] = bound_sum<NCH>(bb);
                            S.tcap[T_SUBHMR + (0)][2] = bound_sum<NCH>(ex);
                        }
                        if (want_hmr)
                            for (int g = 0; g < 4; g++)
                                if (S.tidx[T_SUBHMR + (1 + g)] == i && in_group(g, tc)) {
                                    S.tcap[T_SUBHMR + (1 + g)][0] = r;
                                    S.tcap[T_SUBHMR + (1 + g)][1] = group_sum<NCH>(bb, g, true);
                                    S.tcap[T_SUBHMR + (1 + g)][2] = group_sum<NCH>(ex, g, true);
                                }
                    }
                    for (int a = 0; a < n_ap; a++)
                        if (S.tidx[T_APEDGE + (a)] == i) {
#pragma unroll
                            for (int ch = 0; ch < NCH; ch++) S.tcap[T_APEDGE + (a)][ch] = ex[ch];
                        }
                }
            }
            __syncthreads();
        }
        if (cfg.do_sub) {
            argmax_reduce<NCH>(amU, S);
            argmax_reduce<NCH>(amS, S);
        }
        if (CS > 1) {
            if (threadIdx.x == 0) {
                S.pub_v[0] = amU.v; S.pub_r[0] = amU.r; S.pub_i[0] = amU.i;
                S.pub_v[1] = amS.v; S.pub_r[1] = amS.r; S.pub_i[1] = amS.i;
            }
            csync();  // every CTA's pass B results are visible
            merge_targets();
            if (crank == 0 && threadIdx.x == 0)
                for (unsigned int rk = 1; rk < CS; rk++) {
                    const ScanShared<NCH>* P = peer(rk);
                    amU.offer(P->pub_v[0], P->pub_r[0], P->pub_i[0]);
                    amS.offer(P->pub_v[1], P->pub_r[1], P->pub_i[1]);
                }
        }
        __syncthreads();

        // ------------------------- rank 0, thread 0: SO solve + checks
        if (threadIdx.x == 0 && crank == 0) {
            {
                // innermost particle over the halo's work items (SO_properties.py:407-409)
                unsigned long long mr = ~0ull;
                int mf = -1;
                const uint32_t ib = ha.item_base[h], ni = ha.n_items[h];
                for (uint32_t k = 0; k < ni; k++)
                    if (item_minr[ib + k] < mr || (item_minr[ib + k] == mr && item_minfof[ib + k] < mf)) {
                        mr = item_minr[ib + k];
                        mf = item_minfof[ib + k];
                    }
                sr->cen_fof = mf;
            }
            int fail = 0;
            double required = 0.0;
            int status = SOAP_HALO_OK;
            // halo_prop_list order: BoundSubhalo, SO..., apertures.  Properties
            // done at an earlier rung are not recomputed (halo_tasks.py:120-123),
            // and the done set is always a prefix of the list.
            const int off_so = cfg.do_sub ? 1 : 0, off_ap = off_so + cfg.n_so, nprops = off_ap + n_ap;
            const int p0 = ha.ndone[h];
            int p = p0;
            const double r_last = n > 0 ? __longlong_as_double((long long)R[n - 1].rbits) : 0.0;
            for (int q = 0; q < SOAP_MAX_SO; q++) { sr->so_r[q] = 0.0; sr->so_mass[q] = 0.0; sr->so_exists[q] = 0; S.so_r_[q] = 0.0; }
            while (p < nprops && !fail) {
                if (p < off_so) {
                    // BoundSubhalo particle count (subhalo_properties.py:2632-2646)
                    long long Ntot = NB, Nexp = ha.nexp[h];
                    if (Ntot < Nexp) { fail = 1; required = 0.0; }
                    else if (Ntot > Nexp) { fail = 2; status = SOAP_HALO_COUNT_MISMATCH; }
                } else if (p < off_ap) {
                    const int q = p - off_so;
                    if (central) {
                const double rho = cfg.so_rho[q];
                double SO_r = 0.0, SO_mass = 0.0;
                const uint32_t nr_parts = n > nskip_so ? n - nskip_so : 0u;
                if (nr_parts > 0) {
                    uint32_t i = S.tidx[T_SO + (q)];
                    if (i == NONE) {
                        // no particle below the threshold (SO_properties.py:147-156)
                        if (r_last > cfg.r20) { fail = 2; status = SOAP_HALO_SO_NOT_FOUND; }
                        else { fail = 1; required = 0.0; }
                    } else if (i == nskip_so) {
                        // all below: SO_properties.py:157-177
                        uint32_t ip = S.tidx[T_NONNEG];
                        if (ip == NONE) { fail = 2; status = SOAP_HALO_SO_NOT_FOUND; }
                        else {
                            double rp = S.tcap[T_NONNEG][0], cmp = S.tcap[T_NONNEG][1];
                            SO_r = sqrt(0.75 * cmp / (SOAP_PI * rp * rho));
                            SO_mass = cmp * SO_r / rp;
                        }
                    } else {
                        // intersecting interval (SO_properties.py:180-201)
                        double r2 = S.tcap[T_SO + (q)][0];
                        double cum2 = S.tcap[T_SO + (q)][1], cum1 = S.tcap[T_SO + (q)][2];
                        double r1 = __longlong_as_double((long long)R[i - 1].rbits);
                        float M1 = so_cm32(cum1, r1, cfg.nu), M2 = so_cm32(cum2, r2, cfg.nu);
                        bool ab1 = so_density(M1, r1) > rho, ab2 = so_density(M2, r2) > rho;
                        double cum = cum2;
                        bool ran_out = false;
                        while (r1 == r2 || ab1 == ab2) {
                            i++;
                            if (i >= n) { ran_out = true; break; }
                            r1 = r2; M1 = M2; ab1 = ab2;
                            r2 = __longlong_as_double((long long)R[i].rbits);
                            cum += (double)R[i].m;
                            M2 = so_cm32(cum, r2, cfg.nu);
                            ab2 = so_density(M2, r2) > rho;
                        }
                        if (ran_out) {
                            if (r_last > cfg.r20) { fail = 2; status = SOAP_HALO_SO_NOT_FOUND; }
                            else { fail = 1; required = 0.0; }
                        } else {
                            // SO_properties.py:206-215 (float32 M promoted to float64)
                            double dM1 = (double)M1, dM2 = (double)M2;
                            double rho_dim = rho * (r1 * r1 * r1) / dM1;
                            double slope_dim = (dM2 - dM1) / (r2 - r1) * (r1 / dM1);
                            double root;
                            if (brentq_dev(1.0, r2 / r1, rho_dim, slope_dim, &root)) {
                                fail = 2; status = SOAP_HALO_ROOT_FAILED;
                            } else {
                                SO_r = r1 * root;
                                SO_mass = 4.0 / 3.0 * SOAP_PI * (SO_r * SO_r * SO_r) * rho;
                            }
                        }
                    }
                }
                if (!fail) {
                    sr->so_r[q] = SO_r;
                    sr->so_mass[q] = SO_mass;
                    sr->so_exists[q] = (SO_r > 0.0 && SO_mass > 0.0) ? 1 : 0;  // SO_properties.py:457
                    S.so_r_[q] = sr->so_exists[q] ? SO_r : 0.0;
                }
                    }
                } else {
                    // apertures ascending (aperture_properties.py:4140-4143)
                    const int a = p - off_ap;
                    if (ha.cur_r[h] < cfg.ap_r[a]) { fail = 1; required = cfg.ap_mpc[a] * cfg.mpc2c; }
                }
                if (!fail) p++;
            }
            ha.commit_lo[h] = p0;
            ha.commit_hi[h] = p;
            ha.ndone[h] = p;
            S.commit_lo_ = p0;
            S.commit_hi_ = p;
            if (p > p0 && fail < 2) atomicAdd(&ctr->mom_pairs, (unsigned long long)n);
            if (!fail && p >= nprops) {
                (ha.out + (int64_t)h * ha.ncol)[3] = (double)n;
            }
            S.fail_ = fail;
            S.required_ = required;
            if (fail >= 2) {
                ha.status[h] = status;
                ha.state[h] = ST_DONE_FAIL;
            } else if (fail == 1) {
                if (ladder_step(ha, h, required)) next[atomicAdd(&ctr->n_next, 1u)] = h;
            } else {
                ha.state[h] = ST_FINAL;
                atomicAdd(&ctr->pairs, (unsigned long long)n);
            }
            if (fail < 2 && cfg.do_sub && S.commit_lo_ == 0 && S.commit_hi_ >= 1) {
                // subhalo scan results
                sr->sub_vmax_u_r = amU.i == NONE ? 0.0 : amU.r;
                sr->sub_vmax_u_v = amU.i == NONE ? 0.0 : amU.v;
                sr->sub_vmax_s_r = amS.i == NONE ? 0.0 : amS.r;
                sr->sub_vmax_s_v = amS.i == NONE ? 0.0 : amS.v;
                double enc = 0.0;
                for (int ch = 1; ch < NCH; ch += 2) enc = fmax(enc, S.rmaxc[ch]);
                sr->sub_enclose = enc;
                {
                    double t[NCH];
                    uint32_t cn[NCH];
                    for (int ch = 0; ch < NCH; ch++) { t[ch] = S.tot[ch]; cn[ch] = S.cnt[ch]; }
                    for (int ty = 0; ty < 4; ty++) {
                        if (NCH == 2) {
                            sr->bound_mass[ty] = ty == 1 ? t[1] : 0.0;
                            sr->bound_count[ty] = ty == 1 ? cn[1] : 0u;
                        } else {
                            sr->bound_mass[ty] = t[(2 * ty + 1) % NCH];
                            sr->bound_count[ty] = cn[(2 * ty + 1) % NCH];
                        }
                    }
                }
                // half-mass radii of the bound subhalo (half_mass_radius.py:64-80)
                for (int g = 0; g < 5; g++) {
                    double hm = 0.0;
                    uint32_t i = S.tidx[T_SUBHMR + (g)];
                    if (cfg.do_sub && Mb_g[g] != 0.0 && i != NONE) {
                        double rmax_ = S.tcap[T_SUBHMR + (g)][0], Wmax = S.tcap[T_SUBHMR + (g)][1], Wmin = S.tcap[T_SUBHMR + (g)][2];
                        double rmin_ = 0.0;
                        // previous member of the subset (walk back over the sorted records)
                        for (uint32_t j = i; j-- > 0;) {
                            uint32_t f = R[j].flags;
                            uint32_t tcj = NCH == 2 ? 1u : (f & 3u);
                            if ((f & 4u) && (g == 0 || in_group(g - 1, tcj))) {
                                rmin_ = __longlong_as_double((long long)R[j].rbits);
                                break;
                            }
                        }
                        double target = 0.5 * Mb_g[g];
                        if (Wmin == Wmax) hm = 0.5 * (rmin_ + rmax_);
                        else hm = rmin_ + (target - Wmin) / (Wmax - Wmin) * (rmax_ - rmin_);
                    }
                    sr->sub_hmr[g] = hm;
                }
            }
        }
        __syncthreads();
        // aperture half-mass thresholds from the edge captures (totals inside)
        if (crank == 0 && threadIdx.x < n_ap * 4) {
            int a = threadIdx.x / 4, g = threadIdx.x % 4;
            double t[NCH];
            if (S.tidx[T_APEDGE + (a)] == NONE) {
                for (int ch = 0; ch < NCH; ch++) t[ch] = S.tot[ch];
            } else {
                for (int ch = 0; ch < NCH; ch++) t[ch] = S.tcap[T_APEDGE + (a)][ch];
            }
            S.ap_thr[a][g] = 0.5 * group_sum<NCH>(t, g, cfg.ap_incl[a] == 0);
        }
        if (CS > 1) {
            csync();  // rank 0's solve outcome is visible
            if (crank != 0) {
                const ScanShared<NCH>* P = peer(0);
                if (threadIdx.x == 0) {
                    S.fail_ = P->fail_; S.commit_lo_ = P->commit_lo_; S.commit_hi_ = P->commit_hi_;
                }
                if (threadIdx.x < SOAP_MAX_SO) S.so_r_[threadIdx.x] = P->so_r_[threadIdx.x];
                if (threadIdx.x < SOAP_MAX_APERTURES * 4)
                    S.ap_thr[threadIdx.x / 4][threadIdx.x % 4] = P->ap_thr[threadIdx.x / 4][threadIdx.x % 4];
            }
        }
        __syncthreads();
        // pass C serves the SOs and apertures committed at this rung
        const int c_so_lo = cfg.do_sub ? 1 : 0, c_ap_lo = c_so_lo + cfg.n_so;
        const bool so_committed = n_so > 0 && S.commit_hi_ > S.commit_lo_ && S.commit_lo_ < c_ap_lo && S.commit_hi_ > c_so_lo;
        const bool ap_committed = n_ap > 0 && S.commit_hi_ > c_ap_lo && S.commit_hi_ > S.commit_lo_;
        const bool need_c = S.fail_ < 2 && (so_committed || (ap_committed && want_hmr));
        // (uniform over the cluster: everyone leaves or everyone stays)
        if (!need_c) { csync(); continue; }

        // ------------------------------------------------------------ pass C
        if (threadIdx.x < NCH) { S.carry[threadIdx.x] = S.carry_in[threadIdx.x]; S.carryc[threadIdx.x] = S.carryc_in[threadIdx.x]; }
        __syncthreads();
        ArgMax amSO[SOAP_MAX_SO];
#pragma unroll
        for (int q = 0; q < SOAP_MAX_SO; q++) amSO[q].init();
        for (uint32_t tile = t_lo; tile < t_hi; tile++) {
            const uint32_t i0 = tile * SCAN_TILE + threadIdx.x * SCAN_K;
            Rec rc[SCAN_K];
            double base[NCH];
            uint32_t basec[NCH];
#pragma unroll
            for (int ch = 0; ch < NCH; ch++) { base[ch] = 0.0; basec[ch] = 0; }
#pragma unroll
            for (int k = 0; k < SCAN_K; k++) {
                if (i0 + k < n) {
                    rc[k] = R[i0 + k];
                    int c = rec_class<NCH>(rc[k].flags);
#pragma unroll
                    for (int ch = 0; ch < NCH; ch++)
                        if (c == ch) { base[ch] += (double)rc[k].m; basec[ch]++; }
                } else {
                    rc[k].rbits = 0; rc[k].m = 0.f; rc[k].flags = 0;
                }
            }
            block_scan_channels<NCH>(base, basec, S);
            double b0[NCH];
#pragma unroll
            for (int ch = 0; ch < NCH; ch++) b0[ch] = base[ch];
#pragma unroll
            for (int k = 0; k < SCAN_K; k++) {
                const uint32_t i = i0 + k;
                if (i >= n) break;
                const double r = __longlong_as_double((long long)rc[k].rbits);
                const double m = (double)rc[k].m;
                const int c = rec_class<NCH>(rc[k].flags);
                const uint32_t tc = NCH == 2 ? 1u : (rc[k].flags & 3u);
                const bool bound = (rc[k].flags & 4u) != 0;
                const uint32_t pos_all = all_cnt<NCH>(basec);
#pragma unroll
                for (int ch = 0; ch < NCH; ch++)
                    if (c == ch) { base[ch] += m; basec[ch]++; }
                const double call_in = all_sum<NCH>(base);
                const double rs = fmax(cfg.soft[tc], r);
#pragma unroll
                for (int q = 0; q < SOAP_MAX_SO; q++)
                    if (q < n_so && S.so_r_[q] > 0.0) {
                        // Vmax_soft inside the SO (SO_properties.py:573-600)
                        if (r < S.so_r_[q] && rs > 0.0 && (min_soft > 1e-8 || pos_all >= S.n_zero))
                            amSO[q].offer(call_in / rs, rs, i);
                        // first dark matter particle outside (SO_properties.py:471-482)
                        if (tc == 1u && r > S.so_r_[q] && i < S.tidx[T_DMOUT + (q)]) atomicMin(&S.tidx[T_DMOUT + (q)], i);
                    }
                if (want_hmr)
                    for (int a = 0; a < n_ap; a++) {
                        if (r > cfg.ap_r[a]) continue;
                        if (!cfg.ap_incl[a] && !bound) continue;
#pragma unroll
                        for (int g = 0; g < 4; g++)
                            if (in_group(g, tc) && S.ap_thr[a][g] > 0.0) {
                                double w = group_sum<NCH>(base, g, cfg.ap_incl[a] == 0);
                                if (w >= S.ap_thr[a][g] && i < S.tidx[T_APHMR + (a) * 4 + (g)]) atomicMin(&S.tidx[T_APHMR + (a) * 4 + (g)], i);
                            }
                    }
            }
            __syncthreads();
            {
                double bb[NCH];
#pragma unroll
                for (int ch = 0; ch < NCH; ch++) bb[ch] = b0[ch];
#pragma unroll
                for (int k = 0; k < SCAN_K; k++) {
                    const uint32_t i = i0 + k;
                    if (i >= n) break;
                    const double r = __longlong_as_double((long long)rc[k].rbits);
                    const double m = (double)rc[k].m;
                    const int c = rec_class<NCH>(rc[k].flags);
                    const uint32_t tc = NCH == 2 ? 1u : (rc[k].flags & 3u);
                    double ex[NCH];
#pragma unroll
                    for (int ch = 0; ch < NCH; ch++) ex[ch] = bb[ch];
#pragma unroll
                    for (int ch = 0; ch < NCH; ch++)
                        if (c == ch) bb[ch] += m;
                    for (int q = 0; q < n_so; q++)
                        if (S.tidx[T_DMOUT + (q)] == i) { S.tcap[T_DMOUT + (q)][0] = r; S.tcap[T_DMOUT + (q)][1] = m; }
                    if (want_hmr)
                        for (int a = 0; a < n_ap; a++)
                            for (int g = 0; g < 4; g++)
                                if (S.tidx[T_APHMR + (a) * 4 + (g)] == i && in_group(g, tc)) {
                                    S.tcap[T_APHMR + (a) * 4 + (g)][0] = r;
                                    S.tcap[T_APHMR + (a) * 4 + (g)][1] = group_sum<NCH>(bb, g, cfg.ap_incl[a] == 0);
                                    S.tcap[T_APHMR + (a) * 4 + (g)][2] = group_sum<NCH>(ex, g, cfg.ap_incl[a] == 0);
                                }
                }
            }
            __syncthreads();
        }
#pragma unroll
        for (int q = 0; q < SOAP_MAX_SO; q++)
            if (q < n_so) {
                argmax_reduce<NCH>(amSO[q], S);
                if (CS > 1 && threadIdx.x == 0) {
                    S.pub_v[2 + q] = amSO[q].v; S.pub_r[2 + q] = amSO[q].r; S.pub_i[2 + q] = amSO[q].i;
                }
            }
        if (CS > 1) {
            csync();  // every CTA's pass C results are visible
            merge_targets();
        }
        __syncthreads();
        if (threadIdx.x == 0 && crank == 0) {
#pragma unroll
            for (int q = 0; q < SOAP_MAX_SO; q++)
                if (q < n_so) {
                    for (unsigned int rk = 1; rk < CS; rk++) {
                        const ScanShared<NCH>* P = peer(rk);
                        amSO[q].offer(P->pub_v[2 + q], P->pub_r[2 + q], P->pub_i[2 + q]);
                    }
                    sr->so_vmax_r[q] = amSO[q].i == NONE ? 0.0 : amSO[q].r;
                    sr->so_vmax_v[q] = amSO[q].i == NONE ? 0.0 : amSO[q].v;
                }
            for (int q = 0; q < n_so; q++) {
                // dm_missed_mass (SO_properties.py:471-482)
                double missed = 0.0;
                uint32_t i = S.tidx[T_DMOUT + (q)];
                if (S.so_r_[q] > 0.0 && i != NONE) {
                    double r2 = S.tcap[T_DMOUT + (q)][0], m2 = S.tcap[T_DMOUT + (q)][1];
                    for (uint32_t j = i; j-- > 0;) {
                        uint32_t f = R[j].flags;
                        uint32_t tcj = NCH == 2 ? 1u : (f & 3u);
                        if (tcj == 1u) {
                            double r1 = __longlong_as_double((long long)R[j].rbits);
                            missed = m2 * (S.so_r_[q] - r1) / (r2 - r1);
                            break;
                        }
                    }
                }
                sr->so_dm_missed[q] = missed;
            }
            for (int a = 0; a < n_ap; a++)
                for (int g = 0; g < 4; g++) {
                    double hm = 0.0;
                    uint32_t i = S.tidx[T_APHMR + (a) * 4 + (g)];
                    if (want_hmr && S.ap_thr[a][g] > 0.0 && i != NONE) {
                        double rmax_ = S.tcap[T_APHMR + (a) * 4 + (g)][0], Wmax = S.tcap[T_APHMR + (a) * 4 + (g)][1], Wmin = S.tcap[T_APHMR + (a) * 4 + (g)][2];
                        double rmin_ = 0.0;
                        for (uint32_t j = i; j-- > 0;) {
                            uint32_t f = R[j].flags;
                            uint32_t tcj = NCH == 2 ? 1u : (f & 3u);
                            if ((cfg.ap_incl[a] || (f & 4u)) && in_group(g, tcj)) {
                                rmin_ = __longlong_as_double((long long)R[j].rbits);
                                break;
                            }
                        }
                        double target = S.ap_thr[a][g];
                        if (Wmin == Wmax) hm = 0.5 * (rmin_ + rmax_);
                        else hm = rmin_ + (target - Wmin) / (Wmax - Wmin) * (rmax_ - rmin_);
                    }
                    sr->ap_hmr[a][g] = hm;
                }
        }
        csync();  // rank 0 is done reading the other CTAs' shared memory
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
    d.do_sub = c.do_subhalo; d.n_so = c.n_so; d.n_ap = c.n_apertures; d.dmo = c.dmo;
    for (int k = 0; k < SOAP_MAX_SO; k++) { d.so_rho[k] = c.so_reference_density[k]; d.so_virial[k] = c.so_virial[k]; }
    for (int a = 0; a < SOAP_MAX_APERTURES; a++) {
        d.ap_r[a] = c.ap_radius[a]; d.ap_mpc[a] = c.ap_physical_mpc[a]; d.ap_incl[a] = c.ap_inclusive[a];
    }
    d.flags = c.property_flags;
    d.lay = row_layout(c);
    return d;
}

int validate_cfg(const soap_halo_config* cfg) {
    if (cfg->n_so < 0 || cfg->n_so > SOAP_MAX_SO) SOAP_FAIL("config: n_so=%d outside [0,%d]", cfg->n_so, SOAP_MAX_SO);
    if (cfg->n_apertures < 0 || cfg->n_apertures > SOAP_MAX_APERTURES)
        SOAP_FAIL("config: n_apertures=%d outside [0,%d]", cfg->n_apertures, SOAP_MAX_APERTURES);
    if (cfg->n_projected != 0) SOAP_FAIL("config: projected apertures are not implemented in this build");
    if (!(cfg->boxsize > 0.0)) SOAP_FAIL("config: boxsize must be positive");
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
        }
        if (cfg->property_flags & PF_TENS) {
            add(p + (kind == 2 ? "StellarInertiaTensorNoniterative" : "TotalInertiaTensorNoniterative"), 6);
            add(p + (kind == 2 ? "StellarInertiaTensorReducedNoniterative" : "TotalInertiaTensorReducedNoniterative"), 6);
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
    WS_GET(required, double, h, "h_required", H); ha.required = required;
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
    WS_GET(multi_list, uint32_t, h, "h_multi", H);
    WS_GET(ctr, Counters, h, "h_ctr", 2);
    WS_GET(n_pend_dev, unsigned int, h, "h_npend", 4);
    // work items: every halo has at least one; large spheres are cut every ITEM_CAND candidates
    size_t items_cap = (size_t)H + (size_t)(16 * (v.n / ITEM_CAND + 1)) + 1024;
    WS_GET(items, Item, h, "h_items", items_cap);
    WS_GET(item_minr, unsigned long long, h, "h_item_minr", items_cap);
    WS_GET(item_minfof, int32_t, h, "h_item_minfof", items_cap);

    PhaseLog& log = c->halo_log;
    log.reset();
    c->last_pairs = 0;
    c->last_candidates = 0;
    c->last_rounds = 0;
    uint32_t* pend = listA;
    uint32_t* next = listB;
    LAUNCH(h, k_init, grid_for(H, 128), 128, 0, stream, ha, (int64_t)H, pend);
    unsigned int n_pend = H;
    CUDA_TRY(cudaMemcpyAsync(n_pend_dev, &n_pend, sizeof(unsigned int), cudaMemcpyHostToDevice, stream));
    const int sm = h->sm_count;
    static bool attr_done = false;
    if (!attr_done) {
        CUDA_TRY(cudaFuncSetAttribute(k_sort_bucket<SB_CAP, 512>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      (int)(SB_CAP * sizeof(Rec))));
        attr_done = true;
    }
    const unsigned int sweep_grid = (unsigned)(sm * 6);  // persistent CTAs striding over the item list
    unsigned long long total_pairs = 0, total_cand = 0, total_count_pairs = 0, total_try_pairs = 0, total_mom_pairs = 0;
    while (n_pend > 0) {
        c->last_rounds++;
        if (c->last_rounds > 200) SOAP_FAIL("soap_process_halos: radius ladder did not terminate");
        CUDA_TRY(cudaMemsetAsync(ctr, 0, sizeof(Counters), stream));
        log.begin("plan", stream);
        LAUNCH(h, k_plan_items, grid_for(n_pend, 128), 128, 0, stream, v, ha, pend, n_pend_dev, items,
               (unsigned int)items_cap, ctr);
        log.end(stream);
        {
            // coarse meshes / huge spheres can need more work items than provisioned
            Counters pc;
            CUDA_TRY(cudaMemcpyAsync(&pc, ctr, sizeof(Counters), cudaMemcpyDeviceToHost, stream));
            CUDA_TRY(cudaStreamSynchronize(stream));
            if (pc.items_overflow) {
                if (pc.n_items >= 0xfff00000u) SOAP_FAIL("soap_process_halos: work item list overflow (%u items)", pc.n_items);
                items_cap = (size_t)pc.n_items + 1024;
                items = (Item*)h->get("h_items", sizeof(Item) * items_cap);
                item_minr = (unsigned long long*)h->get("h_item_minr", sizeof(unsigned long long) * items_cap);
                item_minfof = (int32_t*)h->get("h_item_minfof", sizeof(int32_t) * items_cap);
                if (!items || !item_minr || !item_minfof) return -1;
                CUDA_TRY(cudaMemsetAsync(ctr, 0, sizeof(Counters), stream));
                LAUNCH(h, k_plan_items, grid_for(n_pend, 128), 128, 0, stream, v, ha, pend, n_pend_dev, items,
                       (unsigned int)items_cap, ctr);
            }
        }
        log.begin("count", stream);
        LAUNCH(h, k_count, sweep_grid, TB, 0, stream, v, ha, items, ctr);
        log.end(stream);
        log.begin("gate", stream);
        LAUNCH(h, k_gate, grid_for(n_pend, 128), 128, 0, stream, ha, dc, pend, n_pend_dev, try_list, big_list,
               multi_list, next, ctr);
        log.end(stream);
        Counters hc;
        CUDA_TRY(cudaMemcpyAsync(&hc, ctr, sizeof(Counters), cudaMemcpyDeviceToHost, stream));
        CUDA_TRY(cudaStreamSynchronize(stream));
        if (hc.items_overflow) SOAP_FAIL("soap_process_halos: work item list overflow (%u items)", hc.n_items);
        total_cand += hc.candidates;
        total_count_pairs += hc.count_pairs;
        total_try_pairs += hc.rec_total;
        if (hc.n_try + hc.n_big > 0) {
            const unsigned int n_try = hc.n_try + hc.n_big;
            // workspace for this round
            Rec* recs = (Rec*)h->get("h_recs", sizeof(Rec) * (size_t)(hc.rec_total + 1));
            if (!recs) return -1;
            size_t max_bkt = (size_t)n_try + (size_t)(hc.rec_total / 16) + hc.n_fine + 16;
            Bucket* bkt_small = (Bucket*)h->get("h_bkt_small", sizeof(Bucket) * max_bkt);
            Bucket* bkt_big = (Bucket*)h->get("h_bkt_big", sizeof(Bucket) * max_bkt);
            Bucket* bkt_huge = (Bucket*)h->get("h_bkt_huge", sizeof(Bucket) * (size_t)(hc.n_multi + hc.n_fine + 16));
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
                LAUNCH(h, k_build_buckets, hc.n_multi, 256, 0, stream, ha, multi_list, fine_excl, ctr, bkt_small,
                       bkt_big, bkt_huge);
                log.end(stream);
            }
            if (hc.n_try > 0)
                LAUNCH(h, k_single_buckets, grid_for(hc.n_try, 128), 128, 0, stream, ha, try_list, n_try_dev, ctr,
                       bkt_small, bkt_big);
            log.begin("collect", stream);
            LAUNCH(h, k_collect, sweep_grid, TB, 0, stream, v, ha, dc, items, fine_excl, fine_cur, ctr, recs,
                   item_minr, item_minfof);
            log.end(stream);
            log.begin("sort", stream);
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
            if (hc.n_try > 0) {
                unsigned int g = hc.n_try < (unsigned)(sm * 8) ? hc.n_try : (unsigned)(sm * 8);
                if (cfg->dmo)
                    LAUNCH(h, (k_scan_solve<2, 1>), g, SCAN_NT, 0, stream, ha, dc, try_list, n_try_dev, recs, next,
                           ctr, item_minr, item_minfof);
                else
                    LAUNCH(h, (k_scan_solve<8, 1>), g, SCAN_NT, 0, stream, ha, dc, try_list, n_try_dev, recs, next,
                           ctr, item_minr, item_minfof);
            }
            if (hc.n_big > 0) {
                // one cluster of SCAN_CS CTAs per large halo
                unsigned int ncl = hc.n_big < (unsigned)(sm * 2 / SCAN_CS) ? hc.n_big : (unsigned)(sm * 2 / SCAN_CS);
                if (cfg->dmo)
                    LAUNCH(h, (k_scan_solve<2, SCAN_CS>), ncl * SCAN_CS, SCAN_NT, 0, stream, ha, dc, big_list,
                           &ctr->n_big, recs, next, ctr, item_minr, item_minfof);
                else
                    LAUNCH(h, (k_scan_solve<8, SCAN_CS>), ncl * SCAN_CS, SCAN_NT, 0, stream, ha, dc, big_list,
                           &ctr->n_big, recs, next, ctr, item_minr, item_minfof);
            }
            log.end(stream);
            log.begin("moments", stream);
            if (soap_launch_moments(c, dc, ha, items, &ctr->n_items, hc.n_items, hc.n_mslot, sweep_grid, stream)) return -1;
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
        uint32_t* t = pend; pend = next; next = t;
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
