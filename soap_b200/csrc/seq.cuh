// seq.cuh -- one THREAD per halo: the sorted-profile scans and the SO / Vmax /
// half-mass solves of scan.cuh as three sequential streaming passes.
//
// The warp- and CTA-cooperative version (scan.cuh: scan_solve_halo) spends most
// of its issue slots in code only one lane runs (the brentq solves, the commit
// logic, the half-mass interpolation) and in cross-lane bookkeeping (first-index
// targets found with shared-memory atomics, captures re-derived by the owning
// thread).  For a halo of n <= a few thousand records a single thread walking
// the radially sorted records is far cheaper in issue slots: "first record
// that ..." is just the first time a condition holds, the cumulative mass is a
// running sum in the reference's own order (np.cumsum, SO_properties.py:400),
// and every lane of a warp is busy with its own halo.  Same outputs as
// scan_solve_halo: ScanRes, commit range, status / ladder step, retry list, and
// the shell cuts of the moment stage (moments.cuh: Cuts) stored per halo.
//   SO radius/mass        SO_properties.py:80-217,356-513
//   Vmax                  kinematic_properties.py:555-593
//   half-mass radii       half_mass_radius.py:16-97
#pragma once
#include "moments.cuh"
#include "scan.cuh"

#ifdef __CUDACC__

__device__ __forceinline__ Rec ld_rec(const Rec* p) {
    const uint4 q = __ldg(reinterpret_cast<const uint4*>(p));
    Rec r;
    r.rbits = ((unsigned long long)q.y << 32) | q.x;
    r.m = __uint_as_float(q.z);
    r.flags = q.w;
    return r;
}

// Per-thread double-buffered record stream: batches of SEQ_B records are copied into the thread's
// shared-memory slots with cp.async while the previous batch is being processed, so the walk over
// the sorted records never waits for a global load (every thread reads only its own slots).
constexpr int SEQ_NT = 64, SEQ_B = 8;
constexpr size_t SEQ_SMEM = (size_t)SEQ_NT * 2 * SEQ_B * sizeof(uint4);
struct RecStream {
    const Rec* R;
    uint32_t n;
    uint4* s;  // slot (buf, k) of this thread at s[(buf * SEQ_B + k) * SEQ_NT]
    __device__ __forceinline__ void issue(uint32_t buf, uint32_t i0) {
#pragma unroll
        for (int k = 0; k < SEQ_B; k++)
            if (i0 + k < n) {
                const unsigned dst = (unsigned)__cvta_generic_to_shared(s + (buf * SEQ_B + k) * SEQ_NT);
                asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(R + i0 + k) : "memory");
            }
        asm volatile("cp.async.commit_group;" ::: "memory");
    }
    __device__ __forceinline__ void start() {
        issue(0, 0);
        issue(1, SEQ_B);
    }
    // records must be taken in order i = 0, 1, 2, ...
    __device__ __forceinline__ Rec get(uint32_t i) {
        const uint32_t k = i & (SEQ_B - 1), b = (i / SEQ_B) & 1u;
        if (k == 0) asm volatile("cp.async.wait_group 1;" ::: "memory");
        const uint4 q = s[(b * SEQ_B + k) * SEQ_NT];
        if (k == SEQ_B - 1) issue(b, i + 1 + SEQ_B);  // this buffer is consumed: refill it with the batch after next
        Rec r;
        r.rbits = ((unsigned long long)q.y << 32) | q.x;
        r.m = __uint_as_float(q.z);
        r.flags = q.w;
        return r;
    }
    __device__ __forceinline__ void finish() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }
};

// Is the quotient c / r larger than bc / br (all positive)?  Decided by cross-multiplication; the two
// divisions are done only when the products agree to ~1e-15 (the answer then is that of the rounded
// quotients, which is what the reference's argmax over cum / r compares).
__device__ __forceinline__ bool quotient_greater(double c, double r, double bc, double br) {
    const double lhs = c * br, rhs = bc * r;
    const double d = lhs - rhs, tol = 1e-15 * fabs(rhs);
    if (d > tol) return true;
    if (d < -tol) return false;
    return c / r > bc / br;
}

// running mass of the four half-mass groups (gas, dm, star, baryon = gas + star) : add a particle of type code tc
__device__ __forceinline__ void group_add(double (&g)[4], uint32_t tc, double m) {
    if (tc == 0u) { g[0] += m; g[3] += m; }
    else if (tc == 1u) g[1] += m;
    else if (tc == 2u) { g[2] += m; g[3] += m; }
}

__device__ __forceinline__ double hm_interp(double rmin_, double rmax_, double Wmin, double Wmax, double target) {
    // half_mass_radius.py:64-80
    if (Wmin == Wmax) return 0.5 * (rmin_ + rmax_);
    return rmin_ + (target - Wmin) / (Wmax - Wmin) * (rmax_ - rmin_);
}

// One attempt at the rung whose sphere is the first n records.  iter > 0: a further rung tried within the same
// call (solve_seq_halo): results of the selections committed by the earlier attempts are kept and the commit range
// reported to the moment stage starts at c_lo_first.  defer: on "needs a larger radius" do the ladder step but leave
// the decision where the halo goes to the caller (*pending_out, *required_out).
template <int NCH>
__device__ void solve_seq_once(const HaloArrays& ha, const DevCfg& cfg, const uint32_t h, const uint32_t n,
                               const Rec* __restrict__ R, uint32_t* __restrict__ next, unsigned int* __restrict__ n_next, Counters* ctr,
                               const unsigned long long* __restrict__ minr, const int32_t* __restrict__ minfof,
                               const uint32_t n_min, Cuts* __restrict__ cuts_out, uint4* __restrict__ slots, const int iter,
                               const int c_lo_first, const bool defer, int* fail_out, bool* pending_out,
                               double* required_out) {
    ScanRes* sr = ha.sres + h;
    RecStream rs;
    rs.R = R; rs.n = n; rs.s = slots;
    const bool central = ha.central[h] == 1;
    const int n_so = central ? cfg.n_so : 0;  // SO_properties.py:3627
    const int n_ap = cfg.n_ap;
    const bool want_hmr = (cfg.flags & PF_HMR) != 0;
    // ------------------------------------------------------------ pass A: class totals
    double tot[NCH], rmaxc[NCH];
    uint32_t cnt[NCH], cnt0[NCH], n_zero = 0;
#pragma unroll
    for (int ch = 0; ch < NCH; ch++) { tot[ch] = 0.0; rmaxc[ch] = 0.0; cnt[ch] = 0; cnt0[ch] = 0; }
    rs.start();
    for (uint32_t i = 0; i < n; i++) {
        const Rec rc = rs.get(i);
        const double r = __longlong_as_double((long long)rc.rbits);
        const int c = rec_class<NCH>(rc.flags);
        n_zero += (r == 0.0);
#pragma unroll
        for (int ch = 0; ch < NCH; ch++)
            if (c == ch) {
                tot[ch] += (double)rc.m;
                cnt[ch]++;
                cnt0[ch] += (r <= 1e-8);
                rmaxc[ch] = fmax(rmaxc[ch], r);
            }
    }
    double Mb_g[5];  // bound mass: tot, gas, dm, star, baryon
    Mb_g[0] = bound_sum<NCH>(tot);
    Mb_g[1] = group_sum<NCH>(tot, 0, true);
    Mb_g[2] = group_sum<NCH>(tot, 1, true);
    Mb_g[3] = group_sum<NCH>(tot, 2, true);
    Mb_g[4] = group_sum<NCH>(tot, 3, true);
    uint32_t NB = 0, NB0 = 0;
#pragma unroll
    for (int ch = 1; ch < NCH; ch += 2) { NB += cnt[ch]; NB0 += cnt0[ch]; }
    // SO_properties.py:416: nskip = max(1, argmax(r > 0))
    const uint32_t nskip_so = n_zero >= n ? 1u : (n_zero > 1u ? n_zero : 1u);
    // kinematic_properties.py:584-586 on the bound subset
    const uint32_t fnc_u = NB0 < NB ? NB0 : 0u;
    const uint32_t nskip_u = fnc_u > 1u ? fnc_u : 1u;
    const double min_soft = fmin(fmin(cfg.soft[0], cfg.soft[1]), fmin(cfg.soft[2], cfg.soft[3]));
    const uint32_t nskip_s = (min_soft <= 1e-8) ? fnc_u : 0u;
    double so_rho_max = 0.0;
    for (int q = 0; q < n_so; q++) so_rho_max = fmax(so_rho_max, cfg.so_rho[q]);
    const double r_last = n > 0 ? __longlong_as_double((long long)ld_rec(R + n - 1).rbits) : 0.0;

    // ------------------------------------------------------------ pass B
    // SO state machine per variation (SO_properties.py:140-201): 0 searching the first record at or
    // below the threshold, 1 walking on (equal radii / same side), 2 bracket found, 3 first
    // considered record already below, 4 ran out of records while walking
    int so_st[SOAP_MAX_SO];
    double so_r1[SOAP_MAX_SO], so_r2[SOAP_MAX_SO];
    float so_M1[SOAP_MAX_SO], so_M2[SOAP_MAX_SO];
    bool so_ab1[SOAP_MAX_SO];
    for (int q = 0; q < SOAP_MAX_SO; q++) { so_st[q] = 0; so_r1[q] = so_r2[q] = 0.0; so_M1[q] = so_M2[q] = 0.f; so_ab1[q] = false; }
    int n_walking = 0, n_searching = n_so;
    double rho_search_max = so_rho_max;  // largest threshold among the variations still searching (state 0)
    bool nn_found = false;
    double nn_r = 0.0, nn_cm = 0.0;
    double amU_c = 0.0, amU_r = 0.0, amS_c = 0.0, amS_r = 0.0;  // best (cumulative mass, radius) so far
    bool amU_ok = false, amS_ok = false;
    double sub_hm[5] = {0.0, 0.0, 0.0, 0.0, 0.0};
    unsigned sub_hm_found = 0;
    double last_b[5] = {0.0, 0.0, 0.0, 0.0, 0.0};  // radius of the previous bound record (all, gas, dm, star, baryon)
    int a_next = 0;                                // first aperture whose edge has not been passed
    double ap_thr[SOAP_MAX_APERTURES][4];
    {
        double ga[4] = {0.0, 0.0, 0.0, 0.0}, gb[4] = {0.0, 0.0, 0.0, 0.0};  // group masses so far: all / bound
        uint32_t nbound = 0;
        double cum_all = 0.0, cum_b = 0.0, prev_r = 0.0, prev_cum = 0.0;
        const bool do_so = n_so > 0;
        rs.finish();
        rs.start();
        for (uint32_t i = 0; i < n; i++) {
            const Rec rc = rs.get(i);
            const double r = __longlong_as_double((long long)rc.rbits);
            const double m = (double)rc.m;
            const uint32_t tc = NCH == 2 ? 1u : (rc.flags & 3u);
            const bool bound = (rc.flags & 4u) != 0;
            // first record beyond each aperture radius: group masses in front of it (aperture_properties.py:310)
            while (want_hmr && a_next < n_ap && r > cfg.ap_r[a_next]) {
#pragma unroll
                for (int g = 0; g < 4; g++) ap_thr[a_next][g] = 0.5 * (cfg.ap_incl[a_next] == 0 ? gb[g] : ga[g]);
                a_next++;
            }
            double w_ex[4];
#pragma unroll
            for (int g = 0; g < 4; g++) w_ex[g] = gb[g];
            const double cb_ex_cls = cum_b;
            group_add(ga, tc, m);
            if (bound) group_add(gb, tc, m);
            cum_all += m;  // np.cumsum order (SO_properties.py:400)
            if (do_so && i >= nskip_so && (n_searching > 0 || n_walking > 0 || !nn_found)) {
                const float cm = so_cm32(cum_all, r, cfg.nu);
                const double vol = 4.0 / 3.0 * SOAP_PI * (r * r * r);
                // records whose density is far above every threshold still searched (the bulk of a halo)
                // skip the exact division and the per-variation state (local memory)
                if (n_walking > 0 || (n_searching > 0 && !((double)cm > rho_search_max * vol * (1.0 + 1e-9)))) {
                    const double dens = (double)cm / vol;  // SO_properties.py:420
                    bool changed = false;
                    for (int q = 0; q < n_so; q++) {
                        const bool above = dens > cfg.so_rho[q];
                        if (so_st[q] == 0) {
                            if (!above) {
                                n_searching--;
                                changed = true;
                                if (i == nskip_so) {
                                    so_st[q] = 3;  // all below: SO_properties.py:157-177
                                } else {
                                    const float M1 = so_cm32(prev_cum, prev_r, cfg.nu);
                                    const bool ab1 = so_density(M1, prev_r) > cfg.so_rho[q];
                                    if (prev_r == r || ab1 == above) {
                                        so_st[q] = 1;
                                        n_walking++;
                                        so_r1[q] = r; so_M1[q] = cm; so_ab1[q] = above;
                                    } else {
                                        so_st[q] = 2;
                                        so_r1[q] = prev_r; so_M1[q] = M1; so_r2[q] = r; so_M2[q] = cm;
                                    }
                                }
                            }
                        } else if (so_st[q] == 1) {
                            if (so_r1[q] == r || so_ab1[q] == above) {
                                so_r1[q] = r; so_M1[q] = cm; so_ab1[q] = above;
                            } else {
                                so_st[q] = 2;
                                n_walking--;
                                so_r2[q] = r; so_M2[q] = cm;
                            }
                        }
                    }
                    if (changed) {
                        rho_search_max = 0.0;
                        for (int q = 0; q < n_so; q++)
                            if (so_st[q] == 0) rho_search_max = fmax(rho_search_max, cfg.so_rho[q]);
                    }
                }
                if (!nn_found && !(cm < 0.f)) { nn_found = true; nn_r = r; nn_cm = (double)cm; }
            }
            prev_r = r;
            prev_cum = cum_all;
            if (bound) {
                cum_b += m;
                const double cb_in_cls = cum_b;
                if (cfg.do_sub) {
                    // Vmax of the bound subhalo (subhalo_properties.py:982-1045): first maximum of cum / r
                    if (nbound >= nskip_u && r > 0.0 && (!amU_ok || quotient_greater(cum_b, r, amU_c, amU_r))) {
                        amU_c = cum_b; amU_r = r; amU_ok = true;
                    }
                    const double rs = fmax(cfg.soft[tc], r);
                    if (nbound >= nskip_s && rs > 0.0 && (!amS_ok || quotient_greater(cum_b, rs, amS_c, amS_r))) {
                        amS_c = cum_b; amS_r = rs; amS_ok = true;
                    }
                    // half-mass radii (half_mass_radius.py:63-80)
                    if (!(sub_hm_found & 1u) && Mb_g[0] != 0.0 && cb_in_cls >= 0.5 * Mb_g[0]) {
                        sub_hm_found |= 1u;
                        sub_hm[0] = hm_interp(last_b[0], r, cb_ex_cls, cb_in_cls, 0.5 * Mb_g[0]);
                    }
                    if (want_hmr) {
#pragma unroll
                        for (int g = 0; g < 4; g++)
                            if (in_group(g, tc)) {
                                if (!(sub_hm_found & (2u << g)) && Mb_g[1 + g] != 0.0) {
                                    const double w = gb[g];
                                    if (w >= 0.5 * Mb_g[1 + g]) {
                                        sub_hm_found |= 2u << g;
                                        sub_hm[1 + g] = hm_interp(last_b[1 + g], r, w_ex[g], w, 0.5 * Mb_g[1 + g]);
                                    }
                                }
                                last_b[1 + g] = r;
                            }
                    }
                }
                last_b[0] = r;
                nbound++;
            }
        }
        // apertures whose edge lies beyond the last record: totals inside
        while (want_hmr && a_next < n_ap) {
#pragma unroll
            for (int g = 0; g < 4; g++) ap_thr[a_next][g] = 0.5 * (cfg.ap_incl[a_next] == 0 ? gb[g] : ga[g]);
            a_next++;
        }
    }

    // ------------------------------------------------------------ SO solves
    int par_fail[SOAP_MAX_SO], par_status[SOAP_MAX_SO];
    double par_r[SOAP_MAX_SO], par_mass[SOAP_MAX_SO];
    for (int q = 0; q < n_so; q++) {
        const double rho = cfg.so_rho[q];
        double SO_r = 0.0, SO_mass = 0.0;
        int fail = 0, status = SOAP_HALO_OK;
        const uint32_t nr_parts = n > nskip_so ? n - nskip_so : 0u;
        if (nr_parts > 0) {
            const int st = so_st[q];
            if (st == 0 || st == 1) {
                // no particle below the threshold / ran out while walking (SO_properties.py:147-156,190-193)
                if (r_last > cfg.r20) { fail = 2; status = SOAP_HALO_SO_NOT_FOUND; }
                else { fail = 1; }
            } else if (st == 3) {
                if (!nn_found) { fail = 2; status = SOAP_HALO_SO_NOT_FOUND; }
                else {
                    SO_r = sqrt(0.75 * nn_cm / (SOAP_PI * nn_r * rho));
                    SO_mass = nn_cm * SO_r / nn_r;
                }
            } else {
                // SO_properties.py:206-215 (float32 M promoted to float64)
                const double r1 = so_r1[q], r2 = so_r2[q];
                const double dM1 = (double)so_M1[q], dM2 = (double)so_M2[q];
                const double rho_dim = rho * (r1 * r1 * r1) / dM1;
                const double slope_dim = (dM2 - dM1) / (r2 - r1) * (r1 / dM1);
                double root;
                if (brentq_dev(1.0, r2 / r1, rho_dim, slope_dim, &root)) {
                    fail = 2; status = SOAP_HALO_ROOT_FAILED;
                } else {
                    SO_r = r1 * root;
                    SO_mass = 4.0 / 3.0 * SOAP_PI * (SO_r * SO_r * SO_r) * rho;
                }
            }
        }
        par_fail[q] = fail; par_status[q] = status; par_r[q] = SO_r; par_mass[q] = SO_mass;
    }

    // ------------------------------------------------------------ commit logic + checks
    {
        unsigned long long mr = ~0ull;
        int mf = -1;
        for (uint32_t k = 0; k < n_min; k++)
            if (minr[k] < mr || (minr[k] == mr && minfof[k] < mf)) { mr = minr[k]; mf = minfof[k]; }
        sr->cen_fof = mf;  // innermost particle (SO_properties.py:407-409)
    }
    int fail = 0;
    double required = 0.0;
    int status = SOAP_HALO_OK;
    // halo_prop_list order: BoundSubhalo, SO..., apertures, projected apertures.  Properties done at an
    // earlier rung are not recomputed (halo_tasks.py:120-123); the done set is a prefix of the list.
    const int off_so = cfg.do_sub ? 1 : 0, off_ap = off_so + cfg.n_so, off_pj = off_ap + n_ap;
    const int nprops = off_pj + cfg.n_pj;
    const int p0 = ha.ndone[h];
    int p = p0;
    double so_r_[SOAP_MAX_SO];
    for (int q = 0; q < SOAP_MAX_SO; q++) {
        if (iter == 0) { sr->so_r[q] = 0.0; sr->so_mass[q] = 0.0; sr->so_exists[q] = 0; }
        so_r_[q] = 0.0;
    }
    // BoundSubhalo's particle counts and enclosing radius: of this rung if it is committed now, else as stored
    uint32_t bc[4];
    double enclose = 0.0;
    if (p0 == 0) {
#pragma unroll
        for (int ty = 0; ty < 4; ty++) bc[ty] = NCH == 2 ? (ty == 1 ? cnt[1] : 0u) : cnt[(2 * ty + 1) % NCH];
#pragma unroll
        for (int ch = 1; ch < NCH; ch += 2) enclose = fmax(enclose, rmaxc[ch]);
    } else {
        for (int ty = 0; ty < 4; ty++) bc[ty] = sr->bound_count[ty];
        enclose = sr->sub_enclose;
    }
    uint32_t ap_on = 0, pj_on = 0;
    while (p < nprops && !fail) {
        if (p < off_so) {
            // BoundSubhalo particle count (subhalo_properties.py:2632-2646)
            const long long Ntot = NB, Nexp = ha.nexp[h];
            if (Ntot < Nexp) { fail = 1; required = 0.0; }
            else if (Ntot > Nexp) { fail = 2; status = SOAP_HALO_COUNT_MISMATCH; }
        } else if (p < off_ap) {
            const int q = p - off_so;
            if (central && filter_ok(cfg, cfg.so_filter[q], bc)) {  // SO_properties.py:3627
                const int sf = par_fail[q];
                if (sf) { fail = sf; status = par_status[q]; required = 0.0; }
                if (!fail) {
                    sr->so_r[q] = par_r[q];
                    sr->so_mass[q] = par_mass[q];
                    const int ex = (par_r[q] > 0.0 && par_mass[q] > 0.0) ? 1 : 0;  // SO_properties.py:457
                    sr->so_exists[q] = ex;
                    so_r_[q] = ex ? par_r[q] : 0.0;
                }
            }
        } else if (p < off_pj) {
            // apertures ascending (aperture_properties.py:4082-4143)
            const int a = p - off_ap;
            const int mode = aperture_mode(cfg, a, bc, enclose);
            if (mode == 1 && ha.cur_r[h] < cfg.ap_r[a]) { fail = 1; required = cfg.ap_mpc[a] * cfg.mpc2c; }
            if (mode != 0 && !fail) ap_on |= 1u << a;
        } else {
            // projected apertures use bound particles only and never ask for a larger radius
            // (projected_aperture_properties.py:1888-1892)
            if (filter_ok(cfg, cfg.pj_filter[p - off_pj], bc)) pj_on |= 1u << (p - off_pj);
        }
        if (!fail) p++;
    }
    const int c_lo = p0, c_hi = p;
    const int c_lo_u = iter ? c_lo_first : c_lo;  // first property committed by this call
    sr->ap_on = (iter ? sr->ap_on : 0u) | ap_on;
    sr->pj_on = (iter ? sr->pj_on : 0u) | pj_on;
    ha.commit_lo[h] = c_lo_u;
    ha.commit_hi[h] = c_hi;
    ha.ndone[h] = p;
    if (p > p0 && fail < 2) atomicAdd(&ctr->mom_pairs, (unsigned long long)n);
    if (!fail && p >= nprops) (ha.out + (int64_t)h * ha.ncol)[3] = (double)n;
    if (fail >= 2) {
        ha.status[h] = status;
        ha.state[h] = ST_DONE_FAIL;
    } else if (fail == 1) {
        const bool pending = ladder_step(ha, h, required);
        if (pending && !defer && next) next[atomicAdd(n_next, 1u)] = h;
        *pending_out = pending;
    } else {
        ha.state[h] = ST_FINAL;
        atomicAdd(&ctr->pairs, (unsigned long long)n);
    }
    if (fail < 2 && cfg.do_sub && c_lo == 0 && c_hi >= 1) {
        sr->sub_vmax_u_r = amU_ok ? amU_r : 0.0;
        sr->sub_vmax_u_v = amU_ok ? amU_c / amU_r : 0.0;
        sr->sub_vmax_s_r = amS_ok ? amS_r : 0.0;
        sr->sub_vmax_s_v = amS_ok ? amS_c / amS_r : 0.0;
        double enc = 0.0;
#pragma unroll
        for (int ch = 1; ch < NCH; ch += 2) enc = fmax(enc, rmaxc[ch]);
        sr->sub_enclose = enc;
#pragma unroll
        for (int ty = 0; ty < 4; ty++) {
            if (NCH == 2) {
                sr->bound_mass[ty] = ty == 1 ? tot[1] : 0.0;
                sr->bound_count[ty] = ty == 1 ? cnt[1] : 0u;
            } else {
                sr->bound_mass[ty] = tot[(2 * ty + 1) % NCH];
                sr->bound_count[ty] = cnt[(2 * ty + 1) % NCH];
            }
        }
        for (int g = 0; g < 5; g++) sr->sub_hmr[g] = sub_hm[g];
    }
    // pass C serves the SOs and apertures committed at this rung
    const bool so_committed = n_so > 0 && c_hi > c_lo && c_lo < off_ap && c_hi > off_so;
    const bool ap_committed = n_ap > 0 && c_hi > off_ap && c_hi > c_lo;
    const bool need_c = fail < 2 && (so_committed || ap_committed);
    if (need_c) {
        // ------------------------------------------------------------ pass C
        // the SO radii ascending: the records offered to a variation are a prefix of the sorted profile,
        // so one running maximum serves them all and is snapshotted when the walk passes each radius
        int qs[SOAP_MAX_SO], nq = 0;
        for (int q = 0; q < n_so; q++) {
            if (iter == 0) {
                sr->so_dm_missed[q] = 0.0;
                sr->so_vmax_r[q] = 0.0;
                sr->so_vmax_v[q] = 0.0;
            }
            if (so_r_[q] > 0.0) {
                int k = nq++;
                while (k > 0 && so_r_[qs[k - 1]] > so_r_[q]) { qs[k] = qs[k - 1]; k--; }
                qs[k] = q;
            }
        }
        int next_v = 0, next_dm = 0;
        double bound_v = nq > 0 ? so_r_[qs[0]] : 0.0, bound_dm = bound_v;  // radius of the next variation to close
        double best_c = 0.0, best_r = 0.0;
        bool best_ok = false;
        if (want_hmr && iter == 0)
            for (int a = 0; a < n_ap; a++)
                for (int g = 0; g < 4; g++) sr->ap_hmr[a][g] = 0.0;
        unsigned long long ap_found = 0ull, ap_want = 0ull;  // bit a * 4 + g
        if (want_hmr)
            for (int a = 0; a < n_ap; a++)
                for (int g = 0; g < 4; g++)
                    if (ap_thr[a][g] > 0.0) ap_want |= 1ull << (a * 4 + g);
        double ga[4] = {0.0, 0.0, 0.0, 0.0}, gb[4] = {0.0, 0.0, 0.0, 0.0};  // group masses so far: all / bound
        double cum_all = 0.0, last_dm_r = 0.0;
        bool have_dm = false;
        double last_m[2][4];  // radius of the previous member of (all / bound-only, group)
        for (int b = 0; b < 2; b++)
            for (int g = 0; g < 4; g++) last_m[b][g] = 0.0;
        int a_lo = 0;  // apertures below a_lo lie inside the current radius
        // Vmax_soft of the apertures (aperture_properties.py:3553-3577): running first maximum of cum / max(soft, r)
        // over all records (inclusive spheres) and over the bound ones (exclusive), closed at each aperture edge
        uint32_t NA0 = 0;
#pragma unroll
        for (int ch = 0; ch < NCH; ch++) NA0 += cnt0[ch];
        const uint32_t nskip_all = (min_soft <= 1e-8) ? (NA0 < n ? NA0 : 0u) : 0u;
        double bA_c = 0.0, bA_r = 0.0, bB_c = 0.0, bB_r = 0.0, cum_bnd = 0.0;
        bool bA_ok = false, bB_ok = false;
        uint32_t nb_seen = 0;
        int a_v = 0;
        auto close_ap = [&](int a) {
            const bool incl = cfg.ap_incl[a] != 0;
            const bool ok = incl ? bA_ok : bB_ok;
            sr->ap_vmax_r[a] = ok ? (incl ? bA_r : bB_r) : 0.0;
            sr->ap_vmax_v[a] = ok ? (incl ? bA_c / bA_r : bB_c / bB_r) : 0.0;
        };
        rs.finish();
        rs.start();
        for (uint32_t i = 0; i < n; i++) {
            const Rec rc = rs.get(i);
            const double r = __longlong_as_double((long long)rc.rbits);
            const double m = (double)rc.m;
            const uint32_t tc = NCH == 2 ? 1u : (rc.flags & 3u);
            const bool bound = (rc.flags & 4u) != 0;
            double w_ex[2][4];
#pragma unroll
            for (int g = 0; g < 4; g++) { w_ex[0][g] = ga[g]; w_ex[1][g] = gb[g]; }
            group_add(ga, tc, m);
            if (bound) group_add(gb, tc, m);
            cum_all += m;
            const double rs = fmax(cfg.soft[tc], r);
            if (n_ap > 0) {
                while (a_v < n_ap && r > cfg.ap_r[a_v]) close_ap(a_v++);
                const uint32_t nb_before = nb_seen;
                if (bound) { cum_bnd += m; nb_seen++; }
                if (a_v < n_ap && rs > 0.0) {
                    if (i >= nskip_all && (!bA_ok || quotient_greater(cum_all, rs, bA_c, bA_r))) {
                        bA_c = cum_all; bA_r = rs; bA_ok = true;
                    }
                    if (bound && nb_before >= nskip_s && (!bB_ok || quotient_greater(cum_bnd, rs, bB_c, bB_r))) {
                        bB_c = cum_bnd; bB_r = rs; bB_ok = true;
                    }
                }
            }
            // Vmax_soft inside each SO (SO_properties.py:573-600): close the variations this record lies outside of
            while (next_v < nq && !(r < bound_v)) {
                const int q = qs[next_v];
                sr->so_vmax_r[q] = best_ok ? best_r : 0.0;
                sr->so_vmax_v[q] = best_ok ? best_c / best_r : 0.0;
                next_v++;
                bound_v = next_v < nq ? so_r_[qs[next_v]] : 0.0;
            }
            if (next_v < nq && rs > 0.0 && (min_soft > 1e-8 || i >= n_zero) &&
                (!best_ok || quotient_greater(cum_all, rs, best_c, best_r))) {
                best_c = cum_all; best_r = rs; best_ok = true;
            }
            // first dark matter particle outside each SO (SO_properties.py:471-482)
            if (tc == 1u)
                while (next_dm < nq && r > bound_dm) {
                    const int q = qs[next_dm];
                    if (have_dm) sr->so_dm_missed[q] = m * (so_r_[q] - last_dm_r) / (r - last_dm_r);
                    next_dm++;
                    bound_dm = next_dm < nq ? so_r_[qs[next_dm]] : 0.0;
                }
            if (tc == 1u) { last_dm_r = r; have_dm = true; }
            if (want_hmr && n_ap > 0) {
                while (a_lo < n_ap && r > cfg.ap_r[a_lo]) a_lo++;
                for (int a = a_lo; a < n_ap && ap_found != ap_want; a++) {
                    const bool excl = cfg.ap_incl[a] == 0;
                    if (excl && !bound) continue;
#pragma unroll
                    for (int g = 0; g < 4; g++)
                        if (in_group(g, tc) && ap_thr[a][g] > 0.0 && !((ap_found >> (a * 4 + g)) & 1ull)) {
                            const double w = excl ? gb[g] : ga[g];
                            if (w >= ap_thr[a][g]) {
                                ap_found |= 1ull << (a * 4 + g);
                                sr->ap_hmr[a][g] = hm_interp(last_m[excl ? 1 : 0][g], r, w_ex[excl ? 1 : 0][g], w, ap_thr[a][g]);
                            }
                        }
                }
#pragma unroll
                for (int g = 0; g < 4; g++)
                    if (in_group(g, tc)) {
                        last_m[0][g] = r;
                        if (bound) last_m[1][g] = r;
                    }
            }
        }
        while (a_v < n_ap) close_ap(a_v++);
        for (; next_v < nq; next_v++) {
            const int q = qs[next_v];
            sr->so_vmax_r[q] = best_ok ? best_r : 0.0;
            sr->so_vmax_v[q] = best_ok ? best_c / best_r : 0.0;
        }
    }
    rs.finish();
    // shell cuts of the moment stage for the selections committed by this call
    if (cuts_out && c_hi > c_lo_u && fail < 2) build_cuts(cuts_out[h], cfg, sr, c_lo_u, c_hi, n_so);
    *fail_out = fail;
    *required_out = required;
}

// The solve of one halo.  multi: the records at hand are those of the sphere of the furthest of ha.look[h] ladder
// rungs, sorted by radius, each carrying in flags bits 4-7 the first rung whose periodic r2 test includes it
// (tier.cu); ha.rung_cnt / ha.rung_msum hold the cumulative count and mass of every rung.  When the attempt at a
// rung ends with "needs a larger radius" (an SO that no particle of the sphere reaches, halo_tasks.py:125-140) the
// next rungs are tried right here on longer prefixes of the same records -- density gate, n_loop and ladder exactly as
// if the halo had come back in a new round (halo_tasks.py:73-103,166-187) -- instead of costing a kernel sequence
// each.  Only offered for configurations in which no selection depends on the sphere it was computed in.
template <int NCH>
__device__ void solve_seq_halo(const HaloArrays& ha, const DevCfg& cfg, const uint32_t h, uint32_t n,
                               const Rec* __restrict__ R, uint32_t* __restrict__ next, unsigned int* __restrict__ n_next, Counters* ctr,
                               const unsigned long long* __restrict__ minr, const int32_t* __restrict__ minfof,
                               const uint32_t n_min, Cuts* __restrict__ cuts_out, uint4* __restrict__ slots, const bool multi) {
    const int look = multi ? ha.look[h] : 1;
    const uint32_t n_all = multi ? ha.rung_cnt[(size_t)h * LOOK_MAX + (look - 1)] : n;
    const int c_lo_first = ha.ndone[h];
    const bool has_target = ha.central[h] == 1 && cfg.target_density > 0.0;  // halo_tasks.py:381
    int j = 0;
    for (int iter = 0;; iter++) {
        int fail = 0;
        bool pending = false;
        double required = 0.0;
        solve_seq_once<NCH>(ha, cfg, h, n, R, next, n_next, ctr, minr, minfof, n_min, cuts_out, slots, iter, c_lo_first,
                            j + 1 < look, &fail, &pending, &required);
        if (fail != 1 || !pending) return;
        if (!(j + 1 < look)) return;  // not deferred: the halo is already on the next list
        bool served = false;
        while (j + 1 < look && required == 0.0) {
            j++;
            const uint32_t nj = ha.rung_cnt[(size_t)h * LOOK_MAX + j];
            SOAP_ASSERT(nj <= n_all && nj >= n);
            // the radial order must agree with the rung binning at the prefix boundary (the two come from
            // different roundings of the same distance)
            const bool ok = (nj == 0 || (int)((ld_rec(R + nj - 1).flags >> 4) & 15u) <= j) &&
                            (nj >= n_all || (int)((ld_rec(R + nj).flags >> 4) & 15u) > j);
            if (!ok) break;
            ha.nloop[h] += 1;  // halo_tasks.py:75
            const double r = ha.cur_r[h];
            const double mj = ha.rung_msum[(size_t)h * LOOK_MAX + j];
            const double density = mj / (4.0 / 3.0 * SOAP_PI * (r * r * r));
            if (!has_target || density <= cfg.target_density) {
                ha.cnt[h] = nj;
                ha.msum[h] = mj;
                ha.rung_r[h] = r;
                ha.state[h] = ST_TRY;
                n = nj;
                served = true;
                break;
            }
            if (!ladder_step(ha, h, 0.0)) return;  // out of read radius: status set
        }
        if (!served) {
            if (next) next[atomicAdd(n_next, 1u)] = h;
            return;
        }
    }
}

#endif  // __CUDACC__
