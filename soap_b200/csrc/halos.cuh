// halos.cuh -- shared definitions of the batched halo pipeline (halos.cu, moments.cu)
#pragma once
#include "chunk.cuh"

constexpr double SOAP_PI = 3.141592653589793;

// property_flags bits (include/soap_b200.h)
constexpr uint32_t PF_KIN = 1u, PF_KAPPA = 2u, PF_TENS = 4u, PF_HMR = 8u;

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
    double bound_mass[4];
    uint32_t bound_count[4];
    int32_t cen_fof;
    int32_t so_exists[SOAP_MAX_SO];
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
    if (flags & PF_KAPPA) o += 5;
    b.tens = (flags & PF_TENS) ? o : -1;
    if (flags & PF_TENS) o += 12;
    b.hmr = (flags & PF_HMR) ? o : -1;
    if (flags & PF_HMR) o += 4;
    b.extra = o;
    b.size = o + n_extra;
    return b;
}
constexpr int N_INPUT_COLS = 6;  // status, n_loop, radius, n_pairs, search_radius_out, read_radius_out
constexpr int N_SUB_EXTRA = 5;   // HalfMassRadiusTot, EncloseRadius, Vmax_unsoft, R_vmax_unsoft, spin
constexpr int N_SO_EXTRA = 9;    // r, SO_mass, spin, Mfrac_sat, Mfrac_ext, conc_unsoft, conc_soft, conc_dmo_unsoft, conc_dmo_soft

struct RowLayout {
    int ncol;
    int sub;                       // -1 if absent
    int so[SOAP_MAX_SO];
    int ap[SOAP_MAX_APERTURES];
    BlockLayout bsub, bso, bap;
};

inline RowLayout row_layout(const soap_halo_config& cfg) {
    RowLayout L;
    L.bsub = block_layout(cfg.property_flags, N_SUB_EXTRA);
    L.bso = block_layout(cfg.property_flags, N_SO_EXTRA);
    L.bap = block_layout(cfg.property_flags, 0);
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
    L.ncol = o;
    return L;
}

// Device-side copy of what the kernels need from soap_halo_config.
struct DevCfg {
    double L, halfL, G, H, kpc, r20, nu, mpc2c;
    double soft[4];  // by type code
    double target_density;
    int do_sub, n_so, n_ap, dmo;
    double so_rho[SOAP_MAX_SO];
    int so_virial[SOAP_MAX_SO];
    double ap_r[SOAP_MAX_APERTURES], ap_mpc[SOAP_MAX_APERTURES];
    int ap_incl[SOAP_MAX_APERTURES];
    uint32_t flags;
    RowLayout lay;
};

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
    double* required;  // required radius of the failing property of this rung
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
    int32_t* mslot;            // global bank slot of multi-item halos, -1 otherwise
};

// A work item = a contiguous range [first, first+count) of a halo's candidate
// stream (its rows concatenated in row order).
struct Item {
    uint32_t halo, first, count, pad;
};
constexpr uint32_t ITEM_CAND = 32768;  // candidates per item
constexpr int SWEEP_MAXP = 64;         // row pieces per batch
constexpr uint32_t LONG_PIECE = 1024;  // longer pieces are swept by the whole CTA

#ifdef __CUDACC__
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

// Row = one run of cells along x for fixed (y, z): a contiguous particle span.
struct RowIter {
    int ny, nx, nrows;
};
__device__ __forceinline__ RowIter row_iter(const DimRanges* rg) {
    RowIter it;
    it.ny = range_total(rg[1]);
    it.nx = rg[0].n;
    it.nrows = range_total(rg[2]) * it.ny * it.nx;
    return it;
}
__device__ __forceinline__ void row_span(const ChunkView& v, const DimRanges* rg, const RowIter& it,
                                         int row, uint32_t& s0, uint32_t& s1) {
    int rx = row % it.nx;
    int t = row / it.nx;
    int jy = range_cell(rg[1], t % it.ny);
    int kz = range_cell(rg[2], t / it.ny);
    uint32_t c0 = (uint32_t)rg[0].lo[rx] + (uint32_t)v.res * ((uint32_t)jy + (uint32_t)v.res * (uint32_t)kz);
    uint32_t c1 = c0 + (uint32_t)(rg[0].hi[rx] - rg[0].lo[rx]);
    s0 = v.cell_off[c0];
    s1 = v.cell_off[c1 + 1];
}
__device__ __forceinline__ void halo_ranges(const ChunkView& v, double cx, double cy, double cz,
                                            double r, DimRanges* rg) {
    // called by threads 0..2
    int d = threadIdx.x;
    double c = d == 0 ? cx : (d == 1 ? cy : cz);
    dim_ranges(c, r, v.L, v.pmin[d], v.pmax[d], v.cs[d], v.res, rg[d]);
}

struct Piece {
    uint32_t s0, s1;
};
struct SweepShared {
    DimRanges rg[3];
    Piece pieces[SWEEP_MAXP];
    int np, row, done;
    uint32_t pos;
};

// warp 0: collect the next batch of row pieces of the stream range [a, b)
__device__ inline void sweep_build_batch(const ChunkView& v, SweepShared& S, const RowIter& ri,
                                         uint32_t a, uint32_t b) {
    const int lane = threadIdx.x & 31;
    int np = 0, row = S.row;
    uint32_t pos = S.pos;
    while (row < ri.nrows && pos < b && np < SWEEP_MAXP) {
        const int r = row + lane;
        uint32_t s0 = 0, s1 = 0;
        if (r < ri.nrows) row_span(v, S.rg, ri, r, s0, s1);
        const uint32_t len = s1 - s0;
        uint32_t incl = len;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += t;
        }
        const uint32_t end = pos + incl, start = end - len;
        const uint32_t lo = a > start ? a : start, hi = b < end ? b : end;
        const bool has = (r < ri.nrows) && hi > lo;
        const unsigned bal = __ballot_sync(0xffffffffu, has);
        const int my = np + __popc(bal & ((1u << lane) - 1u));
        const int tot = __popc(bal);
        if (has && my < SWEEP_MAXP) {
            Piece pc;
            pc.s0 = s0 + (lo - start);
            pc.s1 = s0 + (hi - start);
            S.pieces[my] = pc;
        }
        if (np + tot <= SWEEP_MAXP) {
            np += tot;
            pos = __shfl_sync(0xffffffffu, end, 31);
            row += 32;
        } else {
            const unsigned lastm = __ballot_sync(0xffffffffu, has && my == SWEEP_MAXP - 1);
            const int Ln = __ffs(lastm) - 1;
            np = SWEEP_MAXP;
            pos = __shfl_sync(0xffffffffu, end, Ln);
            row += Ln + 1;
        }
    }
    if (lane == 0) {
        S.np = np;
        S.row = row;
        S.pos = pos;
        S.done = !(row < ri.nrows && pos < b);
    }
}

// Sweep the candidates [a, b) of a halo's stream with the whole CTA.  f(t, ok)
// is called warp-synchronously: every lane of a warp calls it together, ok
// tells whether t is a real candidate.
template <int NT, class F>
__device__ inline void sweep_item(const ChunkView& v, SweepShared& S, double cx, double cy, double cz,
                                  double r, uint32_t a, uint32_t b, F f) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    __syncthreads();
    if (threadIdx.x < 3) halo_ranges(v, cx, cy, cz, r, S.rg);
    if (threadIdx.x == 0) { S.row = 0; S.pos = 0; S.done = 0; S.np = 0; }
    __syncthreads();
    const RowIter ri = row_iter(S.rg);
    while (true) {
        if (wid == 0) sweep_build_batch(v, S, ri, a, b);
        __syncthreads();
        const bool done = S.done != 0;
        const int np = S.np;
        for (int p = wid; p < np; p += NT / 32) {
            const Piece pc = S.pieces[p];
            if (pc.s1 - pc.s0 > LONG_PIECE) continue;
            for (uint32_t t0 = pc.s0; t0 < pc.s1; t0 += 32) f(t0 + lane, t0 + lane < pc.s1);
        }
        for (int p = 0; p < np; p++) {
            const Piece pc = S.pieces[p];
            if (pc.s1 - pc.s0 <= LONG_PIECE) continue;
            for (uint32_t t0 = pc.s0 + wid * 32; t0 < pc.s1; t0 += NT) f(t0 + lane, t0 + lane < pc.s1);
        }
        __syncthreads();
        if (done) break;
    }
}
#endif
