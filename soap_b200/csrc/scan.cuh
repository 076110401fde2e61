// scan.cuh -- segmented scans over a halo's radially sorted records and the
// SO / Vmax / half-mass solves that read them.  Shared by the streaming kernel
// of the general path (halos.cu: k_scan_solve, records in global memory) and the
// fused small-halo kernel (small.cu, records in shared memory).
//   SO radius/mass        SO_properties.py:80-217,356-513
//   Vmax                  kinematic_properties.py:555-593
//   half-mass radii       half_mass_radius.py:16-97
#pragma once
#include "halos.cuh"

#ifdef __CUDACC__
#include <cooperative_groups.h>
namespace cg = cooperative_groups;

// A "group" = the threads that own one halo: the whole CTA (NT >= 64) or one
// warp of a multi-warp CTA (NT == 32, small.cu's lock-step warp-per-halo tiers).
template <int NT>
__device__ __forceinline__ void gsync() {
    if (NT == 32) __syncwarp(); else __syncthreads();
}
template <int NT>
__device__ __forceinline__ int gtid() {
    return NT == 32 ? (int)(threadIdx.x & 31) : (int)threadIdx.x;
}
// CTA-wide alignment barrier of the lock-step tiers (named barrier 1): keeps the
// warps of a CTA in the same code region so that they share instruction fetches.
#ifndef SOAP_ALIGN_LEVEL
#define SOAP_ALIGN_LEVEL 2
#endif
__device__ __forceinline__ void align_bar(int level = 1) {
    if (level <= SOAP_ALIGN_LEVEL) asm volatile("barrier.sync 1, %0;" ::"r"(blockDim.x) : "memory");
}

// -------------------------------------------------------------- scan pass
constexpr int SCAN_NT = 256;  // the streaming kernel's block size and records per thread
constexpr int SCAN_K = 4;
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

template <int NCH, int NT>
struct ScanShared {
    double carry[NCH];
    uint32_t carryc[NCH];
    double wsum[NT / 32][NCH];
    uint32_t wcnt[NT / 32][NCH];
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
    // per-SO solve results (computed by one lane each)
    int par_fail[SOAP_MAX_SO], par_status[SOAP_MAX_SO];
    double par_r[SOAP_MAX_SO], par_mass[SOAP_MAX_SO];
    double ap_thr[SOAP_MAX_APERTURES][4];
    // argmax block reduce
    double am_v[NT / 32], am_r[NT / 32];
    uint32_t am_i[NT / 32];
    // aperture Vmax_soft: first maximum of cum / max(soft, r) per radial shell between aperture edges,
    // kind 0 = all records, 1 = bound records (published per CTA, merged by rank 0)
    double apv_v[2][SOAP_MAX_APERTURES], apv_r[2][SOAP_MAX_APERTURES];
    uint32_t apv_i[2][SOAP_MAX_APERTURES];
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
template <int NCH, int NT>
__device__ __forceinline__ void block_scan_channels(double (&v)[NCH], uint32_t (&c)[NCH],
                                                    ScanShared<NCH, NT>& S) {
    const int gt = gtid<NT>();
    const int lane = gt & 31, wid = gt >> 5;
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
    gsync<NT>();
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
    gsync<NT>();
    if (gt == NT - 1) {
        // carry for the next tile = inclusive total of the last thread
#pragma unroll
        for (int ch = 0; ch < NCH; ch++) {
            double base = S.carry[ch];
            uint32_t basec = S.carryc[ch];
            for (int w = 0; w < NT / 32; w++) { base += S.wsum[w][ch]; basec += S.wcnt[w][ch]; }
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
template <int NCH, int NT>
__device__ inline void argmax_reduce(ArgMax& a, ScanShared<NCH, NT>& S) {
    const int gt = gtid<NT>();
    const int lane = gt & 31, wid = gt >> 5;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        double ov = __shfl_xor_sync(0xffffffffu, a.v, o);
        double orr = __shfl_xor_sync(0xffffffffu, a.r, o);
        uint32_t oi = __shfl_xor_sync(0xffffffffu, a.i, o);
        a.offer(ov, orr, oi);
    }
    gsync<NT>();
    if (lane == 0) { S.am_v[wid] = a.v; S.am_r[wid] = a.r; S.am_i[wid] = a.i; }
    gsync<NT>();
    a.v = S.am_v[0]; a.r = S.am_r[0]; a.i = S.am_i[0];
    for (int w = 1; w < NT / 32; w++) a.offer(S.am_v[w], S.am_r[w], S.am_i[w]);
    gsync<NT>();
}

// One CTA (CS == 1) or one cluster of CS CTAs (halos above SCAN_BIG records) per
// halo of the list.  Three streaming passes over the halo's radially sorted
// records.  In a cluster every CTA owns a contiguous range of tiles: pass A
// totals give each CTA its carry-in, the first-index targets and argmax
// candidates found locally in passes B and C are merged by rank 0 through
// distributed shared memory, and rank 0 alone runs the solve.
template <int NCH, int CS, int NT, int K, bool ALIGN = false>
__device__ void scan_solve_halo(ScanShared<NCH, NT>& S, const HaloArrays& ha, const DevCfg& cfg, const uint32_t h,
                                const uint32_t n, const Rec* __restrict__ R, uint32_t* __restrict__ next,
                                Counters* ctr, const unsigned long long* __restrict__ minr,
                                const int32_t* __restrict__ minfof, const uint32_t n_min) {
    constexpr int TILE = NT * K;
    const int gt = gtid<NT>();
    cg::cluster_group cluster = cg::this_cluster();
    const unsigned int crank = CS > 1 ? cluster.block_rank() : 0u;
    auto csync = [&]() { if (CS > 1) cluster.sync(); else gsync<NT>(); };
    auto peer = [&](unsigned int rk) -> ScanShared<NCH, NT>* { return CS > 1 ? cluster.map_shared_rank(&S, rk) : &S; };
    // rank 0: merge the first-index targets found by the other CTAs into S
    auto merge_targets = [&]() {
        if (CS > 1 && crank == 0) {
            constexpr int CW = NCH < 3 ? 3 : NCH;
            for (int j = gt; j < T_COUNT; j += NT) {
                uint32_t best = S.tidx[j];
                unsigned int owner = 0;
                for (unsigned int rk = 1; rk < CS; rk++) {
                    const uint32_t o = peer(rk)->tidx[j];
                    if (o < best) { best = o; owner = rk; }
                }
                S.m_idx[j] = best;
                S.m_owner[j] = (uint8_t)owner;
            }
            gsync<NT>();
            for (int e = gt; e < T_COUNT * CW; e += NT) {
                const int j = e / CW, k = e % CW;
                if (S.m_owner[j]) S.tcap[j][k] = peer(S.m_owner[j])->tcap[j][k];
            }
            for (int j = gt; j < T_COUNT; j += NT) S.tidx[j] = S.m_idx[j];
            gsync<NT>();
        }
    };
    const int lane = gt & 31;
    {
        ScanRes* sr = ha.sres + h;
        const bool central = ha.central[h] == 1;
        const int n_so = central ? cfg.n_so : 0;  // SO_properties.py:3627
        const int n_ap = cfg.n_ap;
        const bool want_hmr = (cfg.flags & PF_HMR) != 0;
        // this CTA's tiles
        const uint32_t ntile = (n + TILE - 1) / TILE;
        const uint32_t tiles_per = (ntile + CS - 1) / CS;
        const uint32_t t_lo = crank * tiles_per < ntile ? crank * tiles_per : ntile;
        const uint32_t t_hi = t_lo + tiles_per < ntile ? t_lo + tiles_per : ntile;
        const uint32_t i_lo = t_lo * TILE;
        const uint32_t i_hi = (unsigned long long)t_hi * TILE < n ? t_hi * TILE : n;
        gsync<NT>();
        // ------------------------------------------------------------ pass A
        if (gt < NCH) {
            S.l_tot[gt] = 0.0; S.l_rmaxc[gt] = 0.0;
            S.l_cnt[gt] = 0; S.l_cnt0[gt] = 0;
        }
        if (gt == 0) S.l_n_zero = 0;
        gsync<NT>();
        {
            double tot[NCH], rmx[NCH];
            uint32_t cnt[NCH], cnt0[NCH], nz = 0;
#pragma unroll
            for (int ch = 0; ch < NCH; ch++) { tot[ch] = 0.0; rmx[ch] = 0.0; cnt[ch] = 0; cnt0[ch] = 0; }
            for (uint32_t i = i_lo + gt; i < i_hi; i += NT) {
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
        if (gt < NCH) {
            const int ch = gt;
            double tot = 0.0, rmx = 0.0, cin = 0.0;
            uint32_t cnt = 0, cnt0 = 0, ccin = 0;
            for (unsigned int rk = 0; rk < CS; rk++) {
                const ScanShared<NCH, NT>* P = peer(rk);
                const double t = P->l_tot[ch];
                const uint32_t c = P->l_cnt[ch];
                if (CS > 1 && rk < crank) { cin += t; ccin += c; }
                tot += t; cnt += c; cnt0 += P->l_cnt0[ch];
                rmx = fmax(rmx, P->l_rmaxc[ch]);
            }
            S.tot[ch] = tot; S.cnt[ch] = cnt; S.cnt0[ch] = cnt0; S.rmaxc[ch] = rmx;
            S.carry_in[ch] = cin; S.carryc_in[ch] = ccin;
        }
        if (gt == NCH) {
            uint32_t nz = 0;
            for (unsigned int rk = 0; rk < CS; rk++) nz += peer(rk)->l_n_zero;
            S.n_zero = nz;
        }
        gsync<NT>();
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
        double so_rho_max = 0.0;
        for (int q = 0; q < n_so; q++) so_rho_max = fmax(so_rho_max, cfg.so_rho[q]);

        // ------------------------------------------------------------ pass B
        if (gt < NCH) { S.carry[gt] = S.carry_in[gt]; S.carryc[gt] = S.carryc_in[gt]; }
        for (int j = gt; j < T_COUNT; j += NT) S.tidx[j] = NONE;
        gsync<NT>();
        ArgMax amU, amS;
        amU.init();
        amS.init();
        for (uint32_t tile = t_lo; tile < t_hi; tile++) {
            const uint32_t i0 = tile * TILE + gt * K;
            Rec rc[K];
            double base[NCH];
            uint32_t basec[NCH];
#pragma unroll
            for (int ch = 0; ch < NCH; ch++) { base[ch] = 0.0; basec[ch] = 0; }
#pragma unroll
            for (int k = 0; k < K; k++) {
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
            block_scan_channels<NCH, NT>(base, basec, S);  // base = exclusive prefix at i0
            double b0[NCH];
            uint32_t bc0[NCH];
#pragma unroll
            for (int ch = 0; ch < NCH; ch++) { b0[ch] = base[ch]; bc0[ch] = basec[ch]; }
            // detection sweep
            const uint32_t seen_nn = S.tidx[T_NONNEG], seen_hm0 = S.tidx[T_SUBHMR];  // stale values are safe (atomicMin)
#pragma unroll
            for (int k = 0; k < K; k++) {
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
                // SO first-below (SO_properties.py:140-147).  Records whose density is far above
                // every threshold (the bulk of a halo) skip the exact division.
                if (n_so > 0 && i >= nskip_so) {
                    const float cm = so_cm32(call_in, r, cfg.nu);
                    const double vol = 4.0 / 3.0 * SOAP_PI * (r * r * r);
                    if (!((double)cm > so_rho_max * vol * (1.0 + 1e-9))) {
                        const double dens = (double)cm / vol;  // SO_properties.py:420
                        for (int q = 0; q < n_so; q++)
                            if (!(dens > cfg.so_rho[q]) && i < S.tidx[T_SO + (q)]) atomicMin(&S.tidx[T_SO + (q)], i);
                    }
                    if (i < seen_nn && !(cm < 0.f)) atomicMin(&S.tidx[T_NONNEG], i);
                }
                if (bound) {
                    const double cb_in = cb_ex + m;
                    // Vmax of the bound subhalo (subhalo_properties.py:982-1045); a record that cannot
                    // reach this thread's running maximum skips the division
                    if (cfg.do_sub) {
                        if (posb >= nskip_u && r > 0.0 && cb_in >= amU.v * r - 1e-12 * fabs(amU.v * r)) amU.offer(cb_in / r, r, i);
                        double rs = fmax(cfg.soft[tc], r);
                        if (posb >= nskip_s && rs > 0.0 && cb_in >= amS.v * rs - 1e-12 * fabs(amS.v * rs)) amS.offer(cb_in / rs, rs, i);
                        // half-mass radii (half_mass_radius.py:63)
                        if (i < seen_hm0 && cb_in >= 0.5 * Mb_g[0]) atomicMin(&S.tidx[T_SUBHMR + (0)], i);
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
            gsync<NT>();
            // capture: the thread whose records hold a newly found index re-derives its values
            {
                auto mine = [&](uint32_t idx) { return idx >= i0 && idx < i0 + K && idx < n; };
                // exclusive class prefix at record i0 + kk of this thread
                auto prefix_at = [&](int kk, double (&ex)[NCH]) {
#pragma unroll
                    for (int ch = 0; ch < NCH; ch++) ex[ch] = b0[ch];
#pragma unroll
                    for (int k = 0; k < K; k++)
                        if (k < kk) {
                            const int c = rec_class<NCH>(rc[k].flags);
#pragma unroll
                            for (int ch = 0; ch < NCH; ch++)
                                if (c == ch) ex[ch] += (double)rc[k].m;
                        }
                };
                auto rec_at = [&](int kk, double& r, double& m, uint32_t& tc, int& c) {
#pragma unroll
                    for (int k = 0; k < K; k++)
                        if (k == kk) {
                            r = __longlong_as_double((long long)rc[k].rbits);
                            m = (double)rc[k].m;
                            tc = NCH == 2 ? 1u : (rc[k].flags & 3u);
                            c = rec_class<NCH>(rc[k].flags);
                        }
                };
                auto capture = [&](int target, int kind, int g) {
                    const uint32_t idx = S.tidx[target];
                    if (!mine(idx)) return;
                    const int kk = (int)(idx - i0);
                    double ex[NCH], in[NCH], r = 0.0, m = 0.0;
                    uint32_t tc = 0;
                    int c = 0;
                    prefix_at(kk, ex);
                    rec_at(kk, r, m, tc, c);
#pragma unroll
                    for (int ch = 0; ch < NCH; ch++) in[ch] = ex[ch] + (c == ch ? m : 0.0);
                    const double call_ex = all_sum<NCH>(ex);
                    if (kind == 0) {  // SO crossing
                        S.tcap[target][0] = r; S.tcap[target][1] = call_ex + m; S.tcap[target][2] = call_ex;
                    } else if (kind == 1) {  // first non-negative cumulative mass
                        S.tcap[target][0] = r;
                        S.tcap[target][1] = (double)so_cm32(call_ex + m, r, cfg.nu);
                    } else if (kind == 2) {  // bound half mass, all types
                        S.tcap[target][0] = r;
                        S.tcap[target][1] = bound_sum<NCH>(in);
                        S.tcap[target][2] = bound_sum<NCH>(ex);
                    } else if (kind == 3) {  // bound half mass of group g
                        if (in_group(g, tc)) {
                            S.tcap[target][0] = r;
                            S.tcap[target][1] = group_sum<NCH>(in, g, true);
                            S.tcap[target][2] = group_sum<NCH>(ex, g, true);
                        }
                    } else {  // aperture edge: class sums in front of it
#pragma unroll
                        for (int ch = 0; ch < NCH; ch++) S.tcap[target][ch] = ex[ch];
                    }
                };
                for (int q = 0; q < n_so; q++) capture(T_SO + q, 0, 0);
                if (n_so > 0) capture(T_NONNEG, 1, 0);
                if (cfg.do_sub) {
                    capture(T_SUBHMR, 2, 0);
                    if (want_hmr)
                        for (int g = 0; g < 4; g++) capture(T_SUBHMR + 1 + g, 3, g);
                }
                for (int a = 0; a < n_ap; a++) capture(T_APEDGE + a, 4, 0);
            }
            gsync<NT>();
        }
        if (cfg.do_sub) {
            argmax_reduce<NCH, NT>(amU, S);
            argmax_reduce<NCH, NT>(amS, S);
        }
        if (CS > 1) {
            if (gt == 0) {
                S.pub_v[0] = amU.v; S.pub_r[0] = amU.r; S.pub_i[0] = amU.i;
                S.pub_v[1] = amS.v; S.pub_r[1] = amS.r; S.pub_i[1] = amS.i;
            }
            csync();  // every CTA's pass B results are visible
            merge_targets();
            if (crank == 0 && gt == 0)
                for (unsigned int rk = 1; rk < CS; rk++) {
                    const ScanShared<NCH, NT>* P = peer(rk);
                    amU.offer(P->pub_v[0], P->pub_r[0], P->pub_i[0]);
                    amS.offer(P->pub_v[1], P->pub_r[1], P->pub_i[1]);
                }
        }
        gsync<NT>();

        if (ALIGN) align_bar(2);
        // ------------- rank 0: the SO solves, one lane per SO variation (they are independent;
        // the sequential commit logic below consumes them in halo_prop_list order)
        if (crank == 0 && gt < n_so) {
            const int q = gt;
            const double r_last = n > 0 ? __longlong_as_double((long long)R[n - 1].rbits) : 0.0;
            const double rho = cfg.so_rho[q];
            double SO_r = 0.0, SO_mass = 0.0;
            int fail = 0, status = SOAP_HALO_OK;
            const uint32_t nr_parts = n > nskip_so ? n - nskip_so : 0u;
            if (nr_parts > 0) {
                uint32_t i = S.tidx[T_SO + (q)];
                if (i == NONE) {
                    // no particle below the threshold (SO_properties.py:147-156)
                    if (r_last > cfg.r20) { fail = 2; status = SOAP_HALO_SO_NOT_FOUND; }
                    else { fail = 1; }
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
                        else { fail = 1; }
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
            S.par_fail[q] = fail; S.par_status[q] = status; S.par_r[q] = SO_r; S.par_mass[q] = SO_mass;
        }
        gsync<NT>();
        // ------------------------- rank 0, thread 0: commit logic + checks
        if (gt == 0 && crank == 0) {
            {
                // innermost particle over the halo's work items (SO_properties.py:407-409)
                unsigned long long mr = ~0ull;
                int mf = -1;
                for (uint32_t k = 0; k < n_min; k++)
                    if (minr[k] < mr || (minr[k] == mr && minfof[k] < mf)) {
                        mr = minr[k];
                        mf = minfof[k];
                    }
                sr->cen_fof = mf;
            }
            int fail = 0;
            double required = 0.0;
            int status = SOAP_HALO_OK;
            // halo_prop_list order: BoundSubhalo, SO..., apertures.  Properties
            // done at an earlier rung are not recomputed (halo_tasks.py:120-123),
            // and the done set is always a prefix of the list.
            const int off_so = cfg.do_sub ? 1 : 0, off_ap = off_so + cfg.n_so, off_pj = off_ap + n_ap;
            const int nprops = off_pj + cfg.n_pj;
            const int p0 = ha.ndone[h];
            int p = p0;
            const double r_last = n > 0 ? __longlong_as_double((long long)R[n - 1].rbits) : 0.0;
            for (int q = 0; q < SOAP_MAX_SO; q++) { sr->so_r[q] = 0.0; sr->so_mass[q] = 0.0; sr->so_exists[q] = 0; S.so_r_[q] = 0.0; }
            // BoundSubhalo's particle counts and enclosing radius: of this rung if it is committed now, else as stored
            uint32_t bc[4];
            double enclose = 0.0;
            if (p0 == 0) {
                for (int ty = 0; ty < 4; ty++) bc[ty] = NCH == 2 ? (ty == 1 ? S.cnt[1] : 0u) : S.cnt[(2 * ty + 1) % NCH];
                for (int ch = 1; ch < NCH; ch += 2) enclose = fmax(enclose, S.rmaxc[ch]);
            } else {
                for (int ty = 0; ty < 4; ty++) bc[ty] = sr->bound_count[ty];
                enclose = sr->sub_enclose;
            }
            uint32_t ap_on = 0, pj_on = 0;
            while (p < nprops && !fail) {
                if (p < off_so) {
                    // BoundSubhalo particle count (subhalo_properties.py:2632-2646)
                    long long Ntot = NB, Nexp = ha.nexp[h];
                    if (Ntot < Nexp) { fail = 1; required = 0.0; }
                    else if (Ntot > Nexp) { fail = 2; status = SOAP_HALO_COUNT_MISMATCH; }
                } else if (p < off_ap) {
                    const int q = p - off_so;
                    if (central && filter_ok(cfg, cfg.so_filter[q], bc)) {  // SO_properties.py:3627
                const int sf = S.par_fail[q];
                double SO_r = S.par_r[q], SO_mass = S.par_mass[q];
                if (sf) { fail = sf; status = S.par_status[q]; required = 0.0; }
                if (!fail) {
                    sr->so_r[q] = SO_r;
                    sr->so_mass[q] = SO_mass;
                    sr->so_exists[q] = (SO_r > 0.0 && SO_mass > 0.0) ? 1 : 0;  // SO_properties.py:457
                    S.so_r_[q] = sr->so_exists[q] ? SO_r : 0.0;
                }
                    }
                } else if (p < off_pj) {
                    // apertures ascending (aperture_properties.py:4082-4143)
                    const int a = p - off_ap;
                    const int mode = aperture_mode(cfg, a, bc, enclose);
                    if (mode == 1 && ha.cur_r[h] < cfg.ap_r[a]) { fail = 1; required = cfg.ap_mpc[a] * cfg.mpc2c; }
                    if (mode != 0 && !fail) ap_on |= 1u << a;
                } else {
                    // projected apertures use bound particles only and never ask for a
                    // larger radius (projected_aperture_properties.py:1888-1892)
                    if (filter_ok(cfg, cfg.pj_filter[p - off_pj], bc)) pj_on |= 1u << (p - off_pj);
                }
                if (!fail) p++;
            }
            sr->ap_on = ap_on;
            sr->pj_on = pj_on;
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
                if (ladder_step(ha, h, required) && next) next[atomicAdd(&ctr->n_next, 1u)] = h;
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
        gsync<NT>();
        // aperture half-mass thresholds from the edge captures (totals inside)
        if (crank == 0)
            for (int e = gt; e < n_ap * 4; e += NT) {
                const int a = e / 4, g = e % 4;
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
                const ScanShared<NCH, NT>* P = peer(0);
                if (gt == 0) {
                    S.fail_ = P->fail_; S.commit_lo_ = P->commit_lo_; S.commit_hi_ = P->commit_hi_;
                }
                if (gt < SOAP_MAX_SO) S.so_r_[gt] = P->so_r_[gt];
                if (gt < SOAP_MAX_APERTURES * 4)
                    S.ap_thr[gt / 4][gt % 4] = P->ap_thr[gt / 4][gt % 4];
            }
        }
        gsync<NT>();
        // pass C serves the SOs and apertures committed at this rung
        const int c_so_lo = cfg.do_sub ? 1 : 0, c_ap_lo = c_so_lo + cfg.n_so;
        const bool so_committed = n_so > 0 && S.commit_hi_ > S.commit_lo_ && S.commit_lo_ < c_ap_lo && S.commit_hi_ > c_so_lo;
        const bool ap_committed = n_ap > 0 && S.commit_hi_ > c_ap_lo && S.commit_hi_ > S.commit_lo_;
        const bool need_c = S.fail_ < 2 && (so_committed || ap_committed);
        if (ALIGN) align_bar(2);
        // (uniform over the cluster: everyone leaves or everyone stays)
        if (!need_c) { csync(); return; }

        // ------------------------------------------------------------ pass C
        if (gt < NCH) { S.carry[gt] = S.carry_in[gt]; S.carryc[gt] = S.carryc_in[gt]; }
        gsync<NT>();
        ArgMax amSO[SOAP_MAX_SO];
#pragma unroll
        for (int q = 0; q < SOAP_MAX_SO; q++) amSO[q].init();
        // aperture Vmax_soft (aperture_properties.py:3553-3577): a record lies in exactly one shell between
        // aperture edges, so one tracker per (kind, shell) is touched per record (local memory); an aperture's
        // maximum is the first maximum over its shells
        ArgMax amAP[2][SOAP_MAX_APERTURES];
        for (int a = 0; a < n_ap; a++) { amAP[0][a].init(); amAP[1][a].init(); }
        uint32_t NA0 = 0;
#pragma unroll
        for (int ch = 0; ch < NCH; ch++) NA0 += S.cnt0[ch];
        const uint32_t nskip_all = (min_soft <= 1e-8) ? (NA0 < n ? NA0 : 0u) : 0u;
        for (uint32_t tile = t_lo; tile < t_hi; tile++) {
            const uint32_t i0 = tile * TILE + gt * K;
            Rec rc[K];
            double base[NCH];
            uint32_t basec[NCH];
#pragma unroll
            for (int ch = 0; ch < NCH; ch++) { base[ch] = 0.0; basec[ch] = 0; }
#pragma unroll
            for (int k = 0; k < K; k++) {
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
            block_scan_channels<NCH, NT>(base, basec, S);
            double b0[NCH];
#pragma unroll
            for (int ch = 0; ch < NCH; ch++) b0[ch] = base[ch];
#pragma unroll
            for (int k = 0; k < K; k++) {
                const uint32_t i = i0 + k;
                if (i >= n) break;
                const double r = __longlong_as_double((long long)rc[k].rbits);
                const double m = (double)rc[k].m;
                const int c = rec_class<NCH>(rc[k].flags);
                const uint32_t tc = NCH == 2 ? 1u : (rc[k].flags & 3u);
                const bool bound = (rc[k].flags & 4u) != 0;
                const uint32_t pos_all = all_cnt<NCH>(basec);
                const uint32_t posb = bound_cnt<NCH>(basec);
#pragma unroll
                for (int ch = 0; ch < NCH; ch++)
                    if (c == ch) { base[ch] += m; basec[ch]++; }
                const double call_in = all_sum<NCH>(base);
                const double rs = fmax(cfg.soft[tc], r);
                if (ap_committed && rs > 0.0) {
                    int sh = 0;
                    while (sh < n_ap && r > cfg.ap_r[sh]) sh++;
                    if (sh < n_ap) {
                        if (pos_all >= nskip_all) {
                            ArgMax& t = amAP[0][sh];
                            if (call_in >= t.v * rs - 1e-12 * fabs(t.v * rs)) t.offer(call_in / rs, rs, i);
                        }
                        if (bound && posb >= nskip_s) {
                            const double cb_in = bound_sum<NCH>(base);
                            ArgMax& t = amAP[1][sh];
                            if (cb_in >= t.v * rs - 1e-12 * fabs(t.v * rs)) t.offer(cb_in / rs, rs, i);
                        }
                    }
                }
#pragma unroll
                for (int q = 0; q < SOAP_MAX_SO; q++)
                    if (q < n_so && S.so_r_[q] > 0.0) {
                        // Vmax_soft inside the SO (SO_properties.py:573-600)
                        if (r < S.so_r_[q] && rs > 0.0 && (min_soft > 1e-8 || pos_all >= S.n_zero) &&
                            call_in >= amSO[q].v * rs - 1e-12 * fabs(amSO[q].v * rs))
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
            gsync<NT>();
            {
                auto mine = [&](uint32_t idx) { return idx >= i0 && idx < i0 + K && idx < n; };
                for (int q = 0; q < n_so; q++) {
                    const uint32_t idx = S.tidx[T_DMOUT + q];
                    if (!mine(idx)) continue;
#pragma unroll
                    for (int k = 0; k < K; k++)
                        if (k == (int)(idx - i0)) {
                            S.tcap[T_DMOUT + q][0] = __longlong_as_double((long long)rc[k].rbits);
                            S.tcap[T_DMOUT + q][1] = (double)rc[k].m;
                        }
                }
                if (want_hmr)
                    for (int a = 0; a < n_ap; a++)
                        for (int g = 0; g < 4; g++) {
                            const uint32_t idx = S.tidx[T_APHMR + a * 4 + g];
                            if (!mine(idx)) continue;
                            const int kk = (int)(idx - i0);
                            double ex[NCH], in[NCH], r = 0.0;
                            uint32_t tc = 0;
#pragma unroll
                            for (int ch = 0; ch < NCH; ch++) ex[ch] = b0[ch];
#pragma unroll
                            for (int k = 0; k < K; k++) {
                                const int c = rec_class<NCH>(rc[k].flags);
                                if (k < kk) {
#pragma unroll
                                    for (int ch = 0; ch < NCH; ch++)
                                        if (c == ch) ex[ch] += (double)rc[k].m;
                                }
                                if (k == kk) {
                                    r = __longlong_as_double((long long)rc[k].rbits);
                                    tc = NCH == 2 ? 1u : (rc[k].flags & 3u);
#pragma unroll
                                    for (int ch = 0; ch < NCH; ch++) in[ch] = (c == ch) ? (double)rc[k].m : 0.0;
                                }
                            }
#pragma unroll
                            for (int ch = 0; ch < NCH; ch++) in[ch] += ex[ch];
                            if (in_group(g, tc)) {
                                S.tcap[T_APHMR + a * 4 + g][0] = r;
                                S.tcap[T_APHMR + a * 4 + g][1] = group_sum<NCH>(in, g, cfg.ap_incl[a] == 0);
                                S.tcap[T_APHMR + a * 4 + g][2] = group_sum<NCH>(ex, g, cfg.ap_incl[a] == 0);
                            }
                        }
            }
            gsync<NT>();
        }
#pragma unroll
        for (int q = 0; q < SOAP_MAX_SO; q++)
            if (q < n_so) {
                argmax_reduce<NCH, NT>(amSO[q], S);
                if (CS > 1 && gt == 0) {
                    S.pub_v[2 + q] = amSO[q].v; S.pub_r[2 + q] = amSO[q].r; S.pub_i[2 + q] = amSO[q].i;
                }
            }
        if (ap_committed)
            for (int kd = 0; kd < 2; kd++)
                for (int a = 0; a < n_ap; a++) {
                    ArgMax t = amAP[kd][a];
                    argmax_reduce<NCH, NT>(t, S);
                    if (gt == 0) { S.apv_v[kd][a] = t.v; S.apv_r[kd][a] = t.r; S.apv_i[kd][a] = t.i; }
                }
        if (CS > 1) {
            csync();  // every CTA's pass C results are visible
            merge_targets();
        }
        gsync<NT>();
        if (gt == 0 && crank == 0) {
            if (ap_committed)
                for (int a = 0; a < n_ap; a++) {
                    const int kd = cfg.ap_incl[a] ? 0 : 1;
                    ArgMax best;
                    best.init();
                    for (int sh = 0; sh <= a; sh++)
                        for (unsigned int rk = 0; rk < CS; rk++) {
                            const ScanShared<NCH, NT>* P = peer(rk);
                            best.offer(P->apv_v[kd][sh], P->apv_r[kd][sh], P->apv_i[kd][sh]);
                        }
                    sr->ap_vmax_r[a] = best.i == NONE ? 0.0 : best.r;
                    sr->ap_vmax_v[a] = best.i == NONE ? 0.0 : best.v;
                }
#pragma unroll
            for (int q = 0; q < SOAP_MAX_SO; q++)
                if (q < n_so) {
                    for (unsigned int rk = 1; rk < CS; rk++) {
                        const ScanShared<NCH, NT>* P = peer(rk);
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
#endif  // __CUDACC__
