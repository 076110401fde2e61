// halos.cuh -- shared definitions of the batched halo pipeline (halos.cu, moments.cu)
#pragma once
#include "chunk.cuh"

constexpr double SOAP_PI = 3.141592653589793;

// property_flags bits (include/soap_b200.h)
constexpr uint32_t PF_KIN = 1u, PF_KAPPA = 2u, PF_TENS = 4u, PF_HMR = 8u, PF_ITER = 16u;

// One in-sphere particle of one halo, the unit of the segmented radial sort.
struct __align__(16) Rec {
    unsigned long long rbits;  // IEEE bits of the float64 radius (non-negative: order-preserving)
    float m;
    uint32_t flags;            // bits 0-1 type code, bit 2 bound to this halo
};

// halo states of the radius ladder (SOAP/core/halo_tasks.py:73-187)
enum : int32_t { ST_PENDING = 0, ST_TRY = 1, ST_FINAL = 2, ST_DONE_FAIL = 3 };

// Result of the sorted-profile pass of one halo (global memory).
struct ScanRes {
    double so_r[SOAP_MAX_SO], so_mass[SOAP_MAX_SO];
    double so_vmax_r[SOAP_MAX_SO], so_vmax_v[SOAP_MAX_SO];  // v = cum/r (times G later)
    double so_dm_missed[SOAP_MAX_SO];
    double sub_vmax_u_r, sub_vmax_u_v, sub_vmax_s_r, sub_vmax_s_v;
    double sub_hmr[5];  // tot, gas, dm, star, baryon
    double sub_enclose;
    double ap_hmr[SOAP_MAX_APERTURES][4];  // gas, dm, star, baryon
    double ap_vmax_r[SOAP_MAX_APERTURES], ap_vmax_v[SOAP_MAX_APERTURES];  // Vmax_soft of the aperture (v = cum / r)
    double bound_mass[4];
    uint32_t bound_count[4];
    int32_t cen_fof;
    int32_t so_exists[SOAP_MAX_SO];
    // apertures / projected apertures committed at this rung that are actually computed (not filtered out,
    // not skipped by the EncloseRadius shortcut); the others keep their zeros
    uint32_t ap_on, pj_on;
};

// Selection block layout inside a result row (offsets relative to block start).
struct BlockLayout {
    int kin, kappa, tens, hmr, extra, size;
};

__host__ __device__ inline BlockLayout block_layout(uint32_t flags, int n_extra) {
    BlockLayout b;
    int o = 17;
    b.kin = (flags & PF_KIN) ? o : -1;
    if (flags & PF_KIN) o += 51;
    b.kappa = (flags & PF_KAPPA) ? o : -1;
    if (flags & PF_KAPPA) o += 9;  // kappa x3, DtoT x2, stellar rotation + cylindrical dispersions x4
    b.tens = (flags & PF_TENS) ? o : -1;
    if (flags & PF_TENS) o += (flags & PF_ITER) ? 24 : 12;  // non-iterative pair, then the iterative pair
    b.hmr = (flags & PF_HMR) ? o : -1;
    if (flags & PF_HMR) o += 4;
    b.extra = o;
    b.size = o + n_extra;
    return b;
}
constexpr int N_INPUT_COLS = 6;  // status, n_loop, radius, n_pairs, search_radius_out, read_radius_out
constexpr int N_SUB_EXTRA = 5;   // HalfMassRadiusTot, EncloseRadius, Vmax_unsoft, R_vmax_unsoft, spin
constexpr int N_SO_EXTRA = 9;    // r, SO_mass, spin, Mfrac_sat, Mfrac_ext, conc_unsoft, conc_soft, conc_dmo_unsoft, conc_dmo_soft

// One projected aperture = three blocks (projx, projy, projz) of PJ_BLOCK columns:
// N[4] M[4] Mtot com[3] vcom[3] proj_veldisp{gas,dm,star} HalfMassRadius{gas,dm,star}
// ProjectedTotalInertiaTensorNoniterative[3] ...ReducedNoniterative[3]
constexpr int PJ_BLOCK = 27;

struct RowLayout {
    int ncol;
    int sub;                       // -1 if absent
    int so[SOAP_MAX_SO];
    int ap[SOAP_MAX_APERTURES];
    int pj[SOAP_MAX_APERTURES];
    int pjb;  // columns of one projection block: PJ_BLOCK (+ the iterative tensor pair)
    BlockLayout bsub, bso, bap;
};

inline RowLayout row_layout(const soap_halo_config& cfg) {
    RowLayout L;
    L.bsub = block_layout(cfg.property_flags, N_SUB_EXTRA);
    L.bso = block_layout(cfg.property_flags, N_SO_EXTRA);
    L.bap = block_layout(cfg.property_flags, 0);
    L.pjb = PJ_BLOCK + ((cfg.property_flags & PF_ITER) ? 6 : 0);
    int o = N_INPUT_COLS;
    L.sub = -1;
    if (cfg.do_subhalo) { L.sub = o; o += L.bsub.size; }
    for (int k = 0; k < SOAP_MAX_SO; k++) {
        L.so[k] = -1;
        if (k < cfg.n_so) { L.so[k] = o; o += L.bso.size; }
    }
    for (int a = 0; a < SOAP_MAX_APERTURES; a++) {
        L.ap[a] = -1;
        if (a < cfg.n_apertures) { L.ap[a] = o; o += L.bap.size; }
    }
    for (int a = 0; a < SOAP_MAX_APERTURES; a++) {
        L.pj[a] = -1;
        if (a < cfg.n_projected) { L.pj[a] = o; o += 3 * L.pjb; }
    }
    L.ncol = o;
    return L;
}

// Device-side copy of what the kernels need from soap_halo_config.
struct DevCfg {
    double L, halfL, G, H, kpc, r20, nu, mpc2c;
    double soft[4];  // by type code
    double target_density;
    int do_sub, n_so, n_ap, n_pj, dmo;
    double pj_r[SOAP_MAX_APERTURES];
    double so_rho[SOAP_MAX_SO];
    int so_virial[SOAP_MAX_SO];
    double ap_r[SOAP_MAX_APERTURES], ap_mpc[SOAP_MAX_APERTURES];
    int ap_incl[SOAP_MAX_APERTURES];
    uint32_t flags;
    // CategoryFilter and the EncloseRadius shortcut (include/soap_b200.h)
    int n_filters;
    long long filter_limit[SOAP_MAX_FILTERS];
    uint32_t filter_types[SOAP_MAX_FILTERS];
    int so_filter[SOAP_MAX_SO], ap_filter[SOAP_MAX_APERTURES], pj_filter[SOAP_MAX_APERTURES];
    double ap_prev[SOAP_MAX_APERTURES];
    RowLayout lay;
};

struct Cuts;  // moments.cuh

// Per-halo arrays (inputs are caller memory, the rest is workspace).
struct HaloArrays {
    const double* cofp;
    const double* sr_in;
    const double* rr_in;
    const int64_t* index;
    const int32_t* central;
    const int64_t* nexp;
    double* cur_r;
    double* rung_r;  // radius of the rung being processed (cur_r may already hold the next rung)
    int32_t* nloop;
    int32_t* state;
    int32_t* status;
    uint32_t* cnt;
    double* msum;
    unsigned long long* rec_off;
    uint32_t* fine_off;
    uint32_t* nfine;
    int32_t* ndone;    // leading halo_prop_list entries already done (halo_tasks.py:62,121-123)
    int32_t* commit_lo;  // properties [lo, hi) computed at this rung
    int32_t* commit_hi;
    ScanRes* sres;
    double* out;
    int64_t ncol;
    // work items of the current rung (sweeps of large halos are split)
    uint32_t* item_base;
    uint32_t* n_items;
    unsigned int* cursor;      // append cursor of single-bucket halos (k_collect)
    unsigned int* items_done;  // last-arriver counter (k_moments)
    int32_t* mslot;            // global bank slot of multi-item halos, -1 otherwise (projected apertures)
    // moment banks of every halo accepted this round (written by the moment kernels, read by k_rows)
    double* gbank;
    double* kraw;  // [H][kraw_nb][2] raw sums of the stellar cylindrical dispersions that have no row column
    int kraw_nb;   // selection blocks per halo: BoundSubhalo + SO variations + apertures
    unsigned long long* bank_off;  // [H] offset of the halo's banks in gbank (doubles)
    Cuts* cuts;                    // [H] shell cuts of the selections committed this round
    // ladder look-ahead: one count sweep bins the sphere of the furthest rung by rung
    int32_t* look;             // rungs covered by this round's sweep (1..LOOK_MAX)
    uint32_t* rung_cnt;        // [H][LOOK_MAX] particles first included at rung k
    double* rung_msum;         // [H][LOOK_MAX] their mass
};

constexpr int LOOK_MAX = 12;  // ladder rungs one count sweep can cover

// A work item = a contiguous range [first, first+count) of a halo's candidate
// stream (its rows concatenated in row order).  row0 / pos0 locate the first
// row that overlaps the range (pos0 = stream offset at the start of row0), so a
// sweep never walks the rows in front of its range.
struct Item {
    uint32_t halo, first, count, row0;
    uint32_t pos0, k, pad0, pad1;
};
constexpr uint32_t ITEM_CAND = 16384;  // candidates per item
constexpr int SWEEP_NT = 256;          // threads of every sweeping kernel = rows per batch

// device counters of one round
struct Counters {
    unsigned int n_try, n_big, n_acc, n_next, n_multi, n_fine;
    unsigned long long rec_single, rec_total;
    unsigned int n_bkt_small, n_bkt_big, n_bkt_huge, n_items;
    unsigned int n_mslot, items_overflow, n_bslot, n_seq;
    unsigned int scan_cursor[4], n_huge;  // dynamic queues of k_scan_solve (CTA / cluster of 8 / cluster of 16 lists,
                                          // [3]: a short list of small halos scanned CTA-wise)
    unsigned long long pairs, candidates, count_pairs, mom_pairs;
    unsigned long long rec_class[4];  // records of the halos scanned by: thread, CTA, cluster of 8, cluster of 16
};

#ifdef __CUDACC__
// CategoryFilter.get_do_calculation (category_filter.py:69-110) for filter f: bc = BoundSubhalo particle
// counts by type code (gas, dm, star, bh)
__device__ __forceinline__ bool filter_ok(const DevCfg& cfg, int f, const uint32_t* bc) {
    if (f <= 0 || f >= cfg.n_filters) return true;  // "basic"
    long long v = 0;
#pragma unroll
    for (int t = 0; t < 4; t++)
        if ((cfg.filter_types[f] >> t) & 1u) v += bc[t];
    return v >= cfg.filter_limit[f];
}
// What the commit logic does with aperture a: 0 = leave the zeros (filtered out, or an inclusive sphere the
// previous radius of which already held every bound particle: aperture_properties.py:4082-4127),
// 1 = compute it with the radius check of :4140-4143, 2 = compute it without (an exclusive sphere whose
// previous radius held every bound particle equals that previous sphere; all its particles are loaded).
__device__ __forceinline__ int aperture_mode(const DevCfg& cfg, int a, const uint32_t* bc, double enclose) {
    const bool ok = filter_ok(cfg, cfg.ap_filter[a], bc);
    // the reference compares with the float32 BoundSubhalo/EncloseRadius it stored
    const bool skip = cfg.do_sub && cfg.ap_prev[a] >= 0.0 && cfg.ap_prev[a] > (double)(float)enclose;
    if (!ok) return 0;
    if (skip) return cfg.ap_incl[a] ? 0 : 2;
    return 1;
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

// Radii of the next rungs of the ladder starting at cur, exactly as repeated
// gate failures would produce them (halo_tasks.py:184-187 with required = 0):
// r[k+1] = min(1.2 r[k], read_radius).  Stops at read_radius; returns the count.
__device__ __forceinline__ int ladder_radii(double cur, double read_radius, int kmax, double* r) {
    int n = 1;
    r[0] = cur;
    while (n < kmax && r[n - 1] < read_radius) {
        r[n] = fmin(__dmul_rn(r[n - 1], 1.2), read_radius);
        n++;
    }
    return n;
}

__device__ __forceinline__ int range_total(const DimRanges& r) {
    int t = 0;
    for (int a = 0; a < r.n; a++) t += r.hi[a] - r.lo[a] + 1;
    return t;
}
__device__ __forceinline__ int range_cell(const DimRanges& r, int idx) {
    for (int a = 0; a < r.n; a++) {
        int len = r.hi[a] - r.lo[a] + 1;
        if (idx < len) return r.lo[a] + idx;
        idx -= len;
    }
    return r.lo[0];
}

// Row = one run of cells along x inside one block, for fixed (y, z): a contiguous particle span.
// nx = number of such pieces over the x ranges of the sphere.
struct RowIter {
    int ny, nx, nrows;
};
__device__ __forceinline__ int range_pieces(const DimRanges& r) {
    int t = 0;
    for (int a = 0; a < r.n; a++) t += (r.hi[a] >> BLK_SHIFT) - (r.lo[a] >> BLK_SHIFT) + 1;
    return t;
}
__device__ __forceinline__ RowIter row_iter(const DimRanges* rg) {
    RowIter it;
    it.ny = range_total(rg[1]);
    it.nx = range_pieces(rg[0]);
    it.nrows = range_total(rg[2]) * it.ny * it.nx;
    return it;
}
__device__ __forceinline__ void row_span(const ChunkView& v, const DimRanges* rg, const RowIter& it,
                                         int row, uint32_t& s0, uint32_t& s1) {
    int px = row % it.nx;
    int t = row / it.nx;
    int jy = range_cell(rg[1], t % it.ny);
    int kz = range_cell(rg[2], t / it.ny);
    // piece px of the x ranges: cells [x0, x1] of one block
    int x0 = rg[0].lo[0], x1 = rg[0].hi[0];
    for (int a = 0; a < rg[0].n; a++) {
        const int lo = rg[0].lo[a], hi = rg[0].hi[a];
        const int np = (hi >> BLK_SHIFT) - (lo >> BLK_SHIFT) + 1;
        if (px < np || a == rg[0].n - 1) {
            const int b0 = ((lo >> BLK_SHIFT) + px) << BLK_SHIFT;
            x0 = b0 > lo ? b0 : lo;
            x1 = b0 + BLK - 1 < hi ? b0 + BLK - 1 : hi;
            break;
        }
        px -= np;
    }
    const uint32_t c0 = blocked_cell(x0, jy, kz, v.nb);
    s0 = v.cell_off[c0];
    s1 = v.cell_off[c0 + (uint32_t)(x1 - x0) + 1u];
}
__device__ __forceinline__ void halo_ranges(const ChunkView& v, double cx, double cy, double cz,
                                            double r, DimRanges* rg, int d) {
    // called by three threads, one per dimension d
    double c = d == 0 ? cx : (d == 1 ? cy : cz);
    dim_ranges(c, r, v.L, v.pmin[d], v.pmax[d], v.cs[d], v.res, rg[d]);
}

// Load-balanced sweep state: one batch = up to SWEEP_NT rows; every row piece
// inside the item's range is cut into 32-candidate chunks and the chunks are
// dealt round-robin to the warps (no warp waits for a long row).
struct SweepShared {
    DimRanges rg[3];
    uint32_t s0[SWEEP_NT];        // first particle of the row piece
    uint32_t len[SWEEP_NT];       // candidates of the row piece
    uint32_t cpre[SWEEP_NT + 1];  // exclusive prefix of chunk counts
    unsigned long long wlen[SWEEP_NT / 32];
    uint32_t wchk[SWEEP_NT / 32];
    unsigned long long pos_next;
};

// Fields of the next chunk a sweep asks to have prefetched into L1 while it works on the current one
// (positions always; the rest is what the kernel reads of the particles that pass the radius test).
enum : unsigned { SW_MASS = 1u, SW_VEL = 2u, SW_IDS = 4u, SW_TYPE = 8u };
__device__ __forceinline__ void prefetch_l1(const void* p) { asm volatile("prefetch.global.L1 [%0];" ::"l"(p)); }

// Sweep the candidates of work item im with the whole CTA (SWEEP_NT threads).
// f(t, ok) is called warp-synchronously: all 32 lanes of a warp call it
// together, ok tells whether particle slot t is a real candidate of the lane.
// A warp's chunks ascend, so the row piece of a chunk is found by walking on
// from the previous one; the chunk after the current one is located first and
// its particle fields are prefetched, so that f's loads do not wait on HBM.
template <unsigned PREF = 0u, class F>
__device__ inline void sweep_item(const ChunkView& v, SweepShared& S, double cx, double cy, double cz,
                                  double r, const Item& im, F f) {
    constexpr int NT = SWEEP_NT, NW = SWEEP_NT / 32;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const unsigned long long a = im.first, b = (unsigned long long)im.first + im.count;
    __syncthreads();
    if (threadIdx.x < 3) halo_ranges(v, cx, cy, cz, r, S.rg, threadIdx.x);
    __syncthreads();
    const RowIter ri = row_iter(S.rg);
    int row_base = (int)im.row0;
    unsigned long long pos_base = im.pos0;
    while (true) {
        const int row = row_base + (int)threadIdx.x;
        uint32_t s0 = 0, s1 = 0;
        if (row < ri.nrows) row_span(v, S.rg, ri, row, s0, s1);
        const uint32_t len = s1 - s0;
        unsigned long long incl = len;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            unsigned long long t = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += t;
        }
        if (lane == 31) S.wlen[wid] = incl;
        __syncthreads();
        unsigned long long base = pos_base;
        for (int w = 0; w < wid; w++) base += S.wlen[w];
        const unsigned long long end = base + incl, start = end - len;
        const unsigned long long lo = a > start ? a : start, hi = b < end ? b : end;
        const uint32_t clen = hi > lo ? (uint32_t)(hi - lo) : 0u;
        uint32_t cincl = (clen + 31u) >> 5;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            uint32_t t = __shfl_up_sync(0xffffffffu, cincl, o);
            if (lane >= o) cincl += t;
        }
        if (lane == 31) S.wchk[wid] = cincl;
        S.s0[threadIdx.x] = s0 + (uint32_t)(lo - start);
        S.len[threadIdx.x] = clen;
        if (threadIdx.x == NT - 1) S.pos_next = end;
        __syncthreads();
        uint32_t cbase = 0;
        for (int w = 0; w < wid; w++) cbase += S.wchk[w];
        S.cpre[threadIdx.x + 1] = cbase + cincl;
        if (threadIdx.x == 0) S.cpre[0] = 0;
        __syncthreads();
        const uint32_t total = S.cpre[NT];
        int jl = 0;  // largest jl with cpre[jl] <= c; cpre[NT] = total > c ends the walk
        auto locate = [&](uint32_t c, uint32_t& t, bool& ok) {
            while (S.cpre[jl + 1] <= c) jl++;
            const uint32_t off = (c - S.cpre[jl]) * 32u + (uint32_t)lane;
            t = S.s0[jl] + off;
            ok = off < S.len[jl];
            if (ok) {
                prefetch_l1(v.px + t); prefetch_l1(v.py + t); prefetch_l1(v.pz + t);
                if (PREF & SW_MASS) prefetch_l1(v.mass + t);
                if (PREF & SW_VEL) { prefetch_l1(v.vx + t); prefetch_l1(v.vy + t); prefetch_l1(v.vz + t); }
                if (PREF & SW_IDS) { prefetch_l1(v.grnr + t); prefetch_l1(v.fof + t); }
                if (PREF & SW_TYPE) prefetch_l1(v.type + t);
            }
        };
        uint32_t c = (uint32_t)wid, t_cur = 0, t_nxt = 0;
        bool ok_cur = false, ok_nxt = false;
        if (c < total) locate(c, t_cur, ok_cur);
        while (c < total) {
            const uint32_t cn = c + NW;
            if (cn < total) locate(cn, t_nxt, ok_nxt);
            f(t_cur, ok_cur);
            c = cn; t_cur = t_nxt; ok_cur = ok_nxt;
        }
        pos_base = S.pos_next;
        row_base += NT;
        __syncthreads();
        if (row_base >= ri.nrows || pos_base >= b) break;
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
#endif
