// moments.cuh -- selection banks, moment terms and the result row shared by the
// sweeping kernel of the general path (moments.cu) and the fused small-halo
// kernel (small.cu).
#pragma once
#include "halos.cuh"

#ifdef __CUDACC__


constexpr int MAX_CUTS = SOAP_MAX_SO + SOAP_MAX_APERTURES + 2;

enum { V_N = 0, V_M, V_MX, V_MV = 5, V_ML = 8, V_MR = 11, V_MRS, V_SAT, V_EXT, V_MIN = 15,
       V_VV = 15, V_XV = 21, V_XX = 22, V_XXR = 28, V_M0 = 34, V_N0 = 35, V_FULL = 36 };

struct Cuts {
    int n;
    double r[MAX_CUTS];
    int strict[MAX_CUTS];
    int pos_so[SOAP_MAX_SO], pos_ap[SOAP_MAX_APERTURES], pos_vmax, pos_tens;
};

__device__ inline void add_cut(Cuts& c, double r, int strict, int* pos) {
    // insertion keeping (r, strict-first) ascending
    int k = c.n;
    while (k > 0 && (c.r[k - 1] > r || (c.r[k - 1] == r && c.strict[k - 1] < strict))) k--;
    for (int j = c.n; j > k; j--) { c.r[j] = c.r[j - 1]; c.strict[j] = c.strict[j - 1]; }
    c.r[k] = r;
    c.strict[k] = strict;
    c.n++;
    (void)pos;
}
__device__ inline int find_cut(const Cuts& c, double r, int strict) {
    for (int k = 0; k < c.n; k++)
        if (c.r[k] == r && c.strict[k] == strict) return k;
    return -1;
}

// Sum of the banks of shells [0, pos] over the bound states and types selected by the masks.  The banks
// handed to the row writer are CUMULATIVE over shells (k_bank_prefix), so shell `pos` holds that sum.
template <int V>
__device__ inline void sel_sum(const double* banks, int NTY, int pos, bool bound_only, unsigned typemask,
                               double* out) {
    for (int i = 0; i < V; i++) out[i] = 0.0;
    for (int s = pos; s <= pos; s++)
        for (int b = bound_only ? 1 : 0; b < 2; b++)
            for (int t = 0; t < NTY; t++) {
                int tcode = NTY == 1 ? 1 : t;
                if (!((typemask >> tcode) & 1u)) continue;
                const double* bk = banks + (size_t)((s * 2 + b) * NTY + t) * V;
                for (int i = 0; i < V; i++) out[i] += bk[i];
            }
}

__device__ inline void cross3(const double* a, const double* b, double* c) {
    c[0] = a[1] * b[2] - a[2] * b[1];
    c[1] = a[2] * b[0] - a[0] * b[2];
    c[2] = a[0] * b[1] - a[1] * b[0];
}

__device__ inline double conc_from_R1(double R1) {
    // SO_properties.py:2724-2735
    const double poly_rev[6] = {-5.07, -43.59, -140.17, -250.14, -222.46, -79.71};
    double x = log10(R1), c = 0.0, xp = 1.0;
    for (int i = 0; i < 6; i++) { c += poly_rev[i] * xp; xp *= x; }
    c = fmax(fmin(c, 3.0), 0.0);
    return (double)(float)pow(10.0, c);
}

// Generic part of a selection block: sums per type code S[t][V]
template <int V>
__device__ void write_block(double* blk, const BlockLayout& bl, const double (*S)[V], const double* centre,
                            const DevCfg& cfg, uint32_t flags) {
    double M[4], Mtot = 0.0;
    for (int t = 0; t < 4; t++) { blk[t] = S[t][V_N]; M[t] = S[t][V_M]; blk[4 + t] = M[t]; Mtot += M[t]; }
    blk[8] = Mtot;
    double vcom[3] = {0, 0, 0};
    if (Mtot != 0.0) {
        for (int d = 0; d < 3; d++) {
            double mx = 0.0, mv = 0.0;
            for (int t = 0; t < 4; t++) { mx += S[t][V_MX + d]; mv += S[t][V_MV + d]; }
            blk[9 + d] = floored_mod(mx / Mtot + centre[d], cfg.L);
            vcom[d] = mv / Mtot;
            blk[12 + d] = vcom[d];
        }
    }
    if constexpr (V >= V_FULL) if (flags & PF_KIN) {
        double* k = blk + bl.kin;
        const int gt[3] = {0, 1, 2};
        for (int g = 0; g < 3; g++) {
            const double* s = S[gt[g]];
            double Mg = s[V_M];
            double* o = k + 15 * g;
            if (Mg != 0.0) {
                double vc[3], mxv[3], L[3];
                for (int d = 0; d < 3; d++) {
                    o[d] = floored_mod(s[V_MX + d] / Mg + centre[d], cfg.L);
                    vc[d] = s[V_MV + d] / Mg;
                    o[3 + d] = vc[d];
                }
                cross3(&s[V_MX], vc, mxv);
                for (int d = 0; d < 3; d++) { L[d] = s[V_ML + d] - mxv[d]; o[6 + d] = L[d]; }
                const int ia[6] = {0, 1, 2, 0, 0, 1}, ib[6] = {0, 1, 2, 1, 2, 2};
                for (int q = 0; q < 6; q++) o[9 + q] = s[V_VV + q] / Mg - vc[ia[q]] * vc[ib[q]];
            }
        }
        // baryons = gas + star (aperture_properties.py:1663-1700, SO_properties.py:1267-1277)
        {
            double Mb = S[0][V_M] + S[2][V_M];
            if (Mb != 0.0) {
                double mx[3], vc[3], ml[3], mxv[3];
                for (int d = 0; d < 3; d++) {
                    mx[d] = S[0][V_MX + d] + S[2][V_MX + d];
                    vc[d] = (S[0][V_MV + d] + S[2][V_MV + d]) / Mb;
                    ml[d] = S[0][V_ML + d] + S[2][V_ML + d];
                }
                cross3(mx, vc, mxv);
                for (int d = 0; d < 3; d++) k[45 + d] = ml[d] - mxv[d];
            }
        }
        // kinetic energies about the total vcom with Hubble flow
        if (Mtot != 0.0) {
            auto ekin = [&](const double* s) {
                double tr = s[V_VV] + s[V_VV + 1] + s[V_VV + 2];
                double vdotmv = vcom[0] * s[V_MV] + vcom[1] * s[V_MV + 1] + vcom[2] * s[V_MV + 2];
                double v2 = vcom[0] * vcom[0] + vcom[1] * vcom[1] + vcom[2] * vcom[2];
                double vdotmx = vcom[0] * s[V_MX] + vcom[1] * s[V_MX + 1] + vcom[2] * s[V_MX + 2];
                double trx = s[V_XX] + s[V_XX + 1] + s[V_XX + 2];
                return 0.5 * (tr - 2.0 * vdotmv + s[V_M] * v2 + 2.0 * cfg.H * (s[V_XV] - vdotmx) +
                              cfg.H * cfg.H * trx);
            };
            double tot[V];
            for (int i = 0; i < V; i++) tot[i] = S[0][i] + S[1][i] + S[2][i] + S[3][i];
            k[48] = ekin(tot);
            if (S[0][V_M] != 0.0) k[49] = ekin(S[0]);
            if (S[2][V_M] != 0.0) k[50] = ekin(S[2]);
        }
    }
}

// inertia tensor (max_iterations=1) from sums inside the cut; n_passed is the
// number of particles handed to the reference function (inertia_tensors.py:58)
template <int V>
__device__ void write_tensor(double* o, const double* s, double n_passed, const DevCfg& cfg) {
    const double k2 = cfg.kpc * cfg.kpc;
    if constexpr (V >= V_FULL) if (n_passed >= 20.0) {
        if (s[V_N] >= 20.0 && s[V_M] != 0.0)  // inertia_tensors.py:103-104
            for (int q = 0; q < 6; q++) o[q] = s[V_XX + q] * k2 / s[V_M];
        double nred = s[V_N] - s[V_N0], mred = s[V_M] - s[V_M0];
        if (nred >= 20.0 && mred != 0.0)
            for (int q = 0; q < 6; q++) o[6 + q] = s[V_XXR + q] / mred;
    }
}


// Shell boundaries of one halo at this rung: the radial cuts of every selection
// committed now (properties [c_lo, c_hi) of halo_prop_list), ascending.
__device__ inline void build_cuts(Cuts& c, const DevCfg& cfg, const ScanRes* sr, int c_lo, int c_hi, int n_so) {
    const int off_so = cfg.do_sub ? 1 : 0, off_ap = off_so + cfg.n_so;
    const bool sub_c = cfg.do_sub && c_lo == 0;
    c.n = 0;
    auto so_c = [&](int q) { return q < n_so && off_so + q >= c_lo && off_so + q < c_hi && sr->so_exists[q]; };
    auto ap_c = [&](int a) { return a < cfg.n_ap && off_ap + a >= c_lo && off_ap + a < c_hi && ((sr->ap_on >> a) & 1u); };
    for (int q = 0; q < n_so; q++)
        if (so_c(q)) add_cut(c, sr->so_r[q], 1, nullptr);
    if (sub_c && sr->sub_vmax_s_r > 0.0) add_cut(c, sr->sub_vmax_s_r, 0, nullptr);
    for (int a = 0; a < cfg.n_ap; a++)
        if (ap_c(a)) add_cut(c, cfg.ap_r[a], 0, nullptr);
    if (sub_c && (cfg.flags & PF_TENS)) add_cut(c, 10.0 * sr->sub_hmr[0], 0, nullptr);
    for (int q = 0; q < n_so; q++) c.pos_so[q] = so_c(q) ? find_cut(c, sr->so_r[q], 1) : -1;  // read for q < n_so only
    c.pos_vmax = (sub_c && sr->sub_vmax_s_r > 0.0) ? find_cut(c, sr->sub_vmax_s_r, 0) : -1;
    for (int a = 0; a < cfg.n_ap; a++) c.pos_ap[a] = ap_c(a) ? find_cut(c, cfg.ap_r[a], 0) : -1;
    c.pos_tens = (sub_c && (cfg.flags & PF_TENS)) ? find_cut(c, 10.0 * sr->sub_hmr[0], 0) : -1;
}

// The V moment terms of one selected particle (halo-centred x, y, z, r) and
// its bank key (shell, bound, type).  aperture_properties.py:1098-1270,
// SO_properties.py:531-692, kinematic_properties.py:91-263.
template <int V, int NTY>
__device__ __forceinline__ int moment_terms(const Cuts& cuts, int ncut, const DevCfg& cfg, double x, double y,
                                            double z, double r, double m, double vx, double vy, double vz,
                                            int32_t g, int64_t hidx, int32_t fof, int32_t cen_fof, uint32_t tc,
                                            double (&val)[V]) {
                    int shell = 0;
                    for (int k = 0; k < ncut; k++) shell += cuts.strict[k] ? !(r < cuts.r[k]) : !(r <= cuts.r[k]);
                    const int bound = g == hidx;
                    const int key = (shell * 2 + bound) * NTY + (NTY == 1 ? 0 : (int)tc);
    #pragma unroll
                    for (int i = 0; i < V; i++) val[i] = 0.0;
                    val[V_N] = 1.0;
                    val[V_M] = m;
                    val[V_MX] = m * x; val[V_MX + 1] = m * y; val[V_MX + 2] = m * z;
                    val[V_MV] = m * vx; val[V_MV + 1] = m * vy; val[V_MV + 2] = m * vz;
                    val[V_ML] = m * (y * vz - z * vy);
                    val[V_ML + 1] = m * (z * vx - x * vz);
                    val[V_ML + 2] = m * (x * vy - y * vx);
                    val[V_MR] = m * r;
                    val[V_MRS] = m * fmax(cfg.soft[tc], r);
                    if (!bound && g >= 0) {
                        // SO_properties.py:461-466
                        if (fof == cen_fof) val[V_SAT] = m; else val[V_EXT] = m;
                    }
                    if constexpr (V >= V_FULL) {
                        val[V_VV] = m * vx * vx; val[V_VV + 1] = m * vy * vy; val[V_VV + 2] = m * vz * vz;
                        val[V_VV + 3] = m * vx * vy; val[V_VV + 4] = m * vx * vz; val[V_VV + 5] = m * vy * vz;
                        val[V_XV] = m * (x * vx + y * vy + z * vz);
                        const double xx = x * x, yy = y * y, zz = z * z, xy = x * y, xz = x * z, yz = y * z;
                        val[V_XX] = m * xx; val[V_XX + 1] = m * yy; val[V_XX + 2] = m * zz;
                        val[V_XX + 3] = m * xy; val[V_XX + 4] = m * xz; val[V_XX + 5] = m * yz;
                        const double nrm = r * r;
                        if (nrm <= 1e-8) {  // np.isclose(norm, 0): inertia_tensors.py:62-64
                            val[V_M0] = m; val[V_N0] = 1.0;
                        } else {
                            const double w = m / nrm;
                            val[V_XXR] = w * xx; val[V_XXR + 1] = w * yy; val[V_XXR + 2] = w * zz;
                            val[V_XXR + 3] = w * xy; val[V_XXR + 4] = w * xz; val[V_XXR + 5] = w * yz;
                        }
                    }
    return key;
}

// Transposed accumulation of staged moment terms: a lane computes the V terms
// of its own particle and stages them in its warp's tile; then lane l owns
// value l and the warp walks the staged particles one by one, so values of
// one bank are summed in registers and a bank is touched only when the key
// changes.  priv != 0: bank_w is private to the warp (plain read-modify-write).
template <int V>
struct BankAcc {
    static constexpr int NA = (V + 31) / 32;
    static constexpr int VP = V | 1;  // odd row stride: conflict-free 64-bit stores
    // V <= 16: two staged particles are added per step, one by each half-warp (lane & 15 = term), into the
    // half-warp's OWN copy of the banks (bank_w holds two copies, `hstride` doubles apart: callers add them up).
    // Plain read-modify-write, no running sums, no flush on a key change -- in cell order the (shell, bound) key
    // changes with almost every particle, and flushing 15 running sums each time was most of this kernel.
    static constexpr bool HALF = V <= 16;
    double acc[NA];
    int cur;
    int hstride;
    __device__ __forceinline__ void init(int half_stride = 0) {
        cur = -1;
        hstride = half_stride;
#pragma unroll
        for (int q = 0; q < NA; q++) acc[q] = 0.0;
    }
    __device__ __forceinline__ void flush(double* bank_w, int priv, int lane) {
        if (HALF) return;
        if (cur >= 0) {
            double* b = bank_w + (size_t)cur * V;
#pragma unroll
            for (int q = 0; q < NA; q++) {
                const int vi = q * 32 + lane;
                if (vi < V && acc[q] != 0.0) {
                    if (priv) b[vi] += acc[q]; else atomicAdd(&b[vi], acc[q]);
                }
                acc[q] = 0.0;
            }
        }
    }
    // warp-synchronous: every lane calls with its own (in, key, val)
    __device__ __forceinline__ void add(bool in, int key, const double (&val)[V], double* stage_w, int* skey_w,
                                        double* bank_w, int priv, int lane) {
        const unsigned bal = __ballot_sync(0xffffffffu, in);
        if (bal == 0u) return;
        if (in) {
            const int slot = __popc(bal & ((1u << lane) - 1u));
            double* st = stage_w + slot * VP;
#pragma unroll
            for (int i = 0; i < V; i++) st[i] = val[i];
            skey_w[slot] = key;
        }
        __syncwarp();
        const int cnt = __popc(bal);
        if (HALF) {
            const int half = lane >> 4, t = lane & 15;
            double* b = bank_w + (size_t)half * hstride;
            if (t < V)
                for (int p = half; p < cnt; p += 2) {
                    double* e = b + (size_t)skey_w[p] * V + t;
                    *e += stage_w[p * VP + t];
                }
            __syncwarp();
            return;
        }
        for (int p = 0; p < cnt; p++) {
            const int k = skey_w[p];
            if (k != cur) { flush(bank_w, priv, lane); cur = k; }
#pragma unroll
            for (int q = 0; q < NA; q++) {
                const int vi = q * 32 + lane;
                if (vi < V) acc[q] += stage_w[p * VP + vi];
            }
        }
        __syncwarp();
    }
};

// ------------------------------------------------------------- result row
// banks = [(ncut + 1) shells][2 bound states][NTY types][V] sums of one halo.
// One thread per selection: 0 = subhalo, 1.. = SO, then apertures.
template <int V, int NTY>
__device__ void write_row(const double* banks, const Cuts& cuts, int ncut, const DevCfg& cfg, const HaloArrays& ha,
                          uint32_t h, const ScanRes* sr, bool sub_c, int n_so, double cx, double cy, double cz,
                          int gt) {
    const int nsel = 1 + SOAP_MAX_SO + SOAP_MAX_APERTURES;
    if (gt < nsel) {
        double* row = ha.out + (int64_t)h * ha.ncol;
        const double centre[3] = {cx, cy, cz};
        const RowLayout& L = cfg.lay;
        double S[4][V];
        const int sel = gt;
        if (sel == 0) {
            if (sub_c) {
                for (int t = 0; t < 4; t++) sel_sum<V>(banks, NTY, ncut, true, 1u << t, S[t]);
                double* blk = row + L.sub;
                write_block<V>(blk, L.bsub, S, centre, cfg, cfg.flags);
                if ((cfg.flags & PF_ITER) && L.bsub.tens >= 0) blk[L.bsub.tens + 12] = ha.rung_r[h];  // iter.cu
                const double Mtot = blk[8];
                blk[15] = Mtot != 0.0 ? sqrt(sr->sub_vmax_s_v * cfg.G) : 0.0;
                blk[16] = Mtot != 0.0 ? sr->sub_vmax_s_r : 0.0;
                double* ex = blk + L.bsub.extra;
                ex[0] = sr->sub_hmr[0];
                ex[1] = Mtot != 0.0 ? sr->sub_enclose : 0.0;
                ex[2] = Mtot != 0.0 ? sqrt(sr->sub_vmax_u_v * cfg.G) : 0.0;
                ex[3] = Mtot != 0.0 ? sr->sub_vmax_u_r : 0.0;
                if (cfg.flags & PF_HMR)
                    for (int g = 0; g < 4; g++) blk[L.bsub.hmr + g] = sr->sub_hmr[1 + g];
                // spin (subhalo_properties.py:1049-1073)
                if (Mtot != 0.0 && blk[16] > 0.0 && blk[15] > 0.0 && cuts.pos_vmax >= 0) {
                    double q[V];
                    sel_sum<V>(banks, NTY, cuts.pos_vmax, true, 0xfu, q);
                    if (q[V_M] > 0.0) {
                        double vc[3] = {blk[12], blk[13], blk[14]}, mxv[3];
                        cross3(&q[V_MX], vc, mxv);
                        double lx = q[V_ML] - mxv[0], ly = q[V_ML + 1] - mxv[1], lz = q[V_ML + 2] - mxv[2];
                        ex[4] = sqrt(lx * lx + ly * ly + lz * lz) / (sqrt(2.0) * q[V_M] * blk[15] * blk[16]);
                    }
                }
                if ((cfg.flags & PF_TENS) && Mtot != 0.0 && cuts.pos_tens >= 0) {
                    double q[V];
                    sel_sum<V>(banks, NTY, cuts.pos_tens, true, 0xfu, q);
                    double npass = blk[0] + blk[1] + blk[2] + blk[3];
                    write_tensor<V>(blk + L.bsub.tens, q, npass, cfg);
                }
            }
        } else if (sel <= SOAP_MAX_SO) {
            const int q = sel - 1;
            if (q < n_so && cuts.pos_so[q] >= 0) {
                const int pos = cuts.pos_so[q];
                for (int t = 0; t < 4; t++) sel_sum<V>(banks, NTY, pos, false, 1u << t, S[t]);
                double* blk = row + L.so[q];
                write_block<V>(blk, L.bso, S, centre, cfg, cfg.flags);
                if ((cfg.flags & PF_ITER) && L.bso.tens >= 0) blk[L.bso.tens + 12] = ha.rung_r[h];  // iter.cu
                const double Mpart = blk[8];
                const double SO_r = sr->so_r[q], SO_m = sr->so_mass[q];
                const double vmax = Mpart != 0.0 ? sqrt(sr->so_vmax_v[q] * cfg.G) : 0.0;
                blk[15] = vmax;
                blk[16] = Mpart != 0.0 ? sr->so_vmax_r[q] : 0.0;
                double* ex = blk + L.bso.extra;
                ex[0] = SO_r;
                ex[1] = SO_m;
                double tot[V];
                for (int i = 0; i < V; i++) tot[i] = S[0][i] + S[1][i] + S[2][i] + S[3][i];
                if (Mpart != 0.0 && vmax > 0.0) {  // SO_properties.py:602-618
                    double vc[3] = {blk[12], blk[13], blk[14]}, mxv[3];
                    cross3(&tot[V_MX], vc, mxv);
                    double lx = tot[V_ML] - mxv[0], ly = tot[V_ML + 1] - mxv[1], lz = tot[V_ML + 2] - mxv[2];
                    ex[2] = sqrt(lx * lx + ly * ly + lz * lz) / (sqrt(2.0) * Mpart * SO_r * vmax);
                }
                ex[3] = tot[V_SAT] / SO_m;
                ex[4] = tot[V_EXT] / SO_m;
                if (cfg.so_virial[q]) {
                    // SO_properties.py:2737-2790
                    const double nu = cfg.nu;
                    if (tot[V_N] >= 10.0) {
                        for (int w = 0; w < 2; w++) {
                            double R1 = w == 0 ? tot[V_MR] : tot[V_MRS];
                            double missed = SO_m - tot[V_M];
                            R1 += SOAP_PI * nu * (SO_r * SO_r * SO_r * SO_r);
                            missed -= nu * 4.0 / 3.0 * SOAP_PI * (SO_r * SO_r * SO_r);
                            R1 += missed * SO_r;
                            R1 /= SO_r * SO_m;
                            ex[5 + w] = conc_from_R1(R1);
                        }
                    }
                    if (S[1][V_N] >= 10.0) {
                        const double dmm = sr->so_dm_missed[q];
                        for (int w = 0; w < 2; w++) {
                            double R1 = w == 0 ? S[1][V_MR] : S[1][V_MRS];
                            R1 += dmm * SO_r;
                            R1 /= SO_r * (S[1][V_M] + dmm);
                            ex[7 + w] = conc_from_R1(R1);
                        }
                    }
                }
                if ((cfg.flags & PF_TENS) && Mpart != 0.0)
                    write_tensor<V>(blk + L.bso.tens, tot, tot[V_N], cfg);
            }
        } else {
            const int a = sel - 1 - SOAP_MAX_SO;
            if (a < cfg.n_ap && cuts.pos_ap[a] >= 0) {
                const int pos = cuts.pos_ap[a];
                const bool excl = cfg.ap_incl[a] == 0;
                for (int t = 0; t < 4; t++) sel_sum<V>(banks, NTY, pos, excl, 1u << t, S[t]);
                double* blk = row + L.ap[a];
                write_block<V>(blk, L.bap, S, centre, cfg, cfg.flags);
                if ((cfg.flags & PF_ITER) && L.bap.tens >= 0) blk[L.bap.tens + 12] = ha.rung_r[h];  // iter.cu
                // aperture_properties.py:3553-3577
                blk[15] = blk[8] != 0.0 ? sqrt(sr->ap_vmax_v[a] * cfg.G) : 0.0;
                blk[16] = blk[8] != 0.0 ? sr->ap_vmax_r[a] : 0.0;
                if (cfg.flags & PF_HMR)
                    for (int g = 0; g < 4; g++) blk[L.bap.hmr + g] = sr->ap_hmr[a][g];
                if ((cfg.flags & PF_TENS) && S[2][V_M] != 0.0) {
                    // all stars of the halo mask (aperture_properties.py:3579-3594)
                    double all[V];
                    sel_sum<V>(banks, NTY, ncut, excl, 1u << 2, all);
                    write_tensor<V>(blk + L.bap.tens, S[2], all[V_N], cfg);
                }
            }
        }
    }
}

// ------------------------------------------------------------ kappa_corot
// get_angular_momentum_and_kappa_corot_weighted (kinematic_properties.py:266-425)
// needs the direction of L before it can split the kinetic energy, so it runs
// after k_moments has written L and vcom of every type: a second sweep adds up
//   Kcorot = sum_{Li > 0, Ri2 != 0} 0.5 Li^2 / (m Ri2)   and   Mcounterrot = sum_{Li < 0} m
// per selection (BoundSubhalo, apertures) and group (gas, stars, baryons) into the
// kappa slots of the row; k_kappa_finish turns them into kappa_corot = Kcorot / K
// and DtoT = 1 - 2 Mcounterrot / M (aperture_properties.py:1147-1270).
struct KapSel {
    double vc[3][3], lh[3][3];  // per group: reference velocity, unit angular momentum
    int ok[3];
    double ex[3], ey[3];  // in-plane axes of the stellar frame (cylindrical_coordinates.py:13-42)
    int cyl_ok;           // Nstar >= 2 and sum(Lstar) != 0 (aperture_properties.py:1483-1490)
    double R;
    int incl, is_sub, strict;  // strict: r < R (SO selections, SO_properties.py:485) instead of r <= R
    double* out;   // the block's 9 kappa / rotation columns, holding raw sums 0..8 until kappa_finish_row
    double* out2;  // raw sums 9, 10 (sum m v_phi^2, sum m v_z^2) in HaloArrays::kraw
    __device__ double* slot(int j) const { return j < 9 ? out + j : out2 + (j - 9); }
};
constexpr int KAPPA_MAX_SEL = 1 + SOAP_MAX_SO + SOAP_MAX_APERTURES;

__device__ inline void kappa_refs(KapSel& k, double* blk, const BlockLayout& bl, double* raw2) {
    const double* kin = blk + bl.kin;
    const double Mg = blk[4], Ms = blk[6];
    for (int g = 0; g < 3; g++) {
        double L[3];
        if (g < 2) {
            const double* o = kin + 15 * (g == 0 ? 0 : 2);
            for (int d = 0; d < 3; d++) { k.vc[g][d] = o[3 + d]; L[d] = o[6 + d]; }
        } else {
            for (int d = 0; d < 3; d++) {
                L[d] = kin[45 + d];
                k.vc[2][d] = (Mg + Ms) != 0.0 ? (Mg * kin[3 + d] + Ms * kin[30 + 3 + d]) / (Mg + Ms) : 0.0;
            }
        }
        const double nrm = sqrt(L[0] * L[0] + L[1] * L[1] + L[2] * L[2]);
        k.ok[g] = nrm > 0.0;
        for (int d = 0; d < 3; d++) k.lh[g][d] = nrm > 0.0 ? L[d] / nrm : 0.0;
    }
    k.out = blk + bl.kappa;
    k.out2 = raw2;
    // stellar frame: z = L_star / |L_star|, x = helper x z normalised, y = z x x
    {
        const double* o = kin + 30;
        const double Lx = o[6], Ly = o[7], Lz = o[8];
        k.cyl_ok = blk[2] >= 2.0 && (Lx + Ly + Lz) != 0.0 && k.ok[1];
        const double* z = k.lh[1];
        // np.allclose(z_axis, [1, 0, 0], rtol=0.1): |z - h| <= 1e-8 + 0.1 |h| per component
        const bool near_x = fabs(z[0] - 1.0) <= 1e-8 + 0.1 && fabs(z[1]) <= 1e-8 && fabs(z[2]) <= 1e-8;
        const double hx = near_x ? 0.0 : 1.0, hy = near_x ? 1.0 : 0.0, hz = 0.0;
        double x[3] = {hy * z[2] - hz * z[1], hz * z[0] - hx * z[2], hx * z[1] - hy * z[0]};
        const double xn = sqrt(x[0] * x[0] + x[1] * x[1] + x[2] * x[2]);
        for (int d = 0; d < 3; d++) k.ex[d] = xn > 0.0 ? x[d] / xn : 0.0;
        k.ey[0] = z[1] * k.ex[2] - z[2] * k.ex[1];
        k.ey[1] = z[2] * k.ex[0] - z[0] * k.ex[2];
        k.ey[2] = z[0] * k.ex[1] - z[1] * k.ex[0];
        if (!(xn > 0.0)) k.cyl_ok = 0;  // the reference divides by zero here (L anti-parallel to x)
    }
}


// the selections of halo h committed at this rung that carry kappa slots; returns their number
__device__ inline int kappa_build_sels(KapSel* sel, const DevCfg& cfg, const HaloArrays& ha, uint32_t h, int c_lo,
                                       int c_hi) {
    const int off_ap = (cfg.do_sub ? 1 : 0) + cfg.n_so;
    double* row = ha.out + (int64_t)h * ha.ncol;
    double* raw2 = ha.kraw + (int64_t)h * ha.kraw_nb * 2;  // two doubles per (halo, selection block)
    int n = 0;
    if (cfg.do_sub && c_lo == 0) {
        kappa_refs(sel[n], row + cfg.lay.sub, cfg.lay.bsub, raw2);
        sel[n].is_sub = 1; sel[n].incl = 0; sel[n].R = 0.0; sel[n].strict = 0;
        n++;
    }
    // SO variations: every particle inside the SO radius (DtoTgas / DtoTstar of SOProperties)
    {
        const int off_so = cfg.do_sub ? 1 : 0;
        const ScanRes* sr = ha.sres + h;
        for (int q = 0; q < cfg.n_so; q++)
            if (off_so + q >= c_lo && off_so + q < c_hi && sr->so_exists[q]) {
                kappa_refs(sel[n], row + cfg.lay.so[q], cfg.lay.bso, raw2 + 2 * (off_so + q));
                sel[n].is_sub = 0; sel[n].incl = 1; sel[n].R = sr->so_r[q]; sel[n].strict = 1;
                n++;
            }
    }
    for (int a = 0; a < cfg.n_ap; a++)
        if (off_ap + a >= c_lo && off_ap + a < c_hi && ((ha.sres[h].ap_on >> a) & 1u)) {
            kappa_refs(sel[n], row + cfg.lay.ap[a], cfg.lay.bap, raw2 + 2 * (off_ap + a));
            sel[n].is_sub = 0; sel[n].incl = cfg.ap_incl[a]; sel[n].R = cfg.ap_r[a]; sel[n].strict = 0;
            n++;
        }
    return n;
}

// one gas or star particle (halo-centred x, y, z, r) into the raw sums acc[selection][11]
__device__ __forceinline__ void kappa_add(const KapSel* sel, int ns, double (*acc)[11], double x, double y, double z,
                                          double r, double m, double vx, double vy, double vz, uint32_t tc,
                                          bool bound) {
    const double rr2 = x * x + y * y + z * z;
    const int g0 = tc == 0u ? 0 : 1;
    for (int s = 0; s < ns; s++) {
        const KapSel& k = sel[s];
        const bool in = k.is_sub ? bound : ((k.strict ? r < k.R : r <= k.R) && (k.incl || bound));
        if (!in) continue;
        for (int gi = 0; gi < 2; gi++) {
            const int g = gi == 0 ? g0 : 2;
            if (!k.ok[g]) continue;
            const double ux = vx - k.vc[g][0], uy = vy - k.vc[g][1], uz = vz - k.vc[g][2];
            const double lx = m * (y * uz - z * uy), ly = m * (z * ux - x * uz), lz = m * (x * uy - y * ux);
            const double Li = lx * k.lh[g][0] + ly * k.lh[g][1] + lz * k.lh[g][2];
            const double rdl = x * k.lh[g][0] + y * k.lh[g][1] + z * k.lh[g][2];
            const double Ri2 = rr2 - rdl * rdl;
            if (Ri2 != 0.0 && Li > 0.0) atomicAdd(&acc[s][g], 0.5 * (Li * Li / (m * Ri2)));
            if (g < 2 && Li < 0.0) atomicAdd(&acc[s][3 + g], m);
        }
        if (tc == 2u && k.cyl_ok) {
            // cylindrical velocity of a star in the frame of L_star, about vcom_star
            // (calculate_cylindrical_velocities, cylindrical_coordinates.py:45-93)
            const double ux = vx - k.vc[1][0], uy = vy - k.vc[1][1], uz = vz - k.vc[1][2];
            const double X = x * k.ex[0] + y * k.ex[1] + z * k.ex[2];
            const double Y = x * k.ey[0] + y * k.ey[1] + z * k.ey[2];
            const double VX = ux * k.ex[0] + uy * k.ex[1] + uz * k.ex[2];
            const double VY = ux * k.ey[0] + uy * k.ey[1] + uz * k.ey[2];
            const double VZ = ux * k.lh[1][0] + uy * k.lh[1][1] + uz * k.lh[1][2];
            const double Rp = sqrt(X * X + Y * Y);
            const double cph = Rp > 0.0 ? X / Rp : 1.0, sph = Rp > 0.0 ? Y / Rp : 0.0;  // arctan2(0, 0) = 0
            const double vr = VX * cph + VY * sph, vp = -VX * sph + VY * cph;
            atomicAdd(&acc[s][5], m * vr); atomicAdd(&acc[s][6], m * vp); atomicAdd(&acc[s][7], m * VZ);
            atomicAdd(&acc[s][8], m * vr * vr); atomicAdd(&acc[s][9], m * vp * vp); atomicAdd(&acc[s][10], m * VZ * VZ);
        }
    }
}

// raw sums -> kappa_corot / DtoT / stellar rotation, once per selection, at the rung that committed it
__device__ inline void kappa_finish_row(const DevCfg& cfg, const HaloArrays& ha, uint32_t h, int c_lo, int c_hi) {
    const int off_ap = (cfg.do_sub ? 1 : 0) + cfg.n_so;
    double* row = ha.out + (int64_t)h * ha.ncol;
    const double* raw2 = ha.kraw + (int64_t)h * ha.kraw_nb * 2;
    auto fin = [&](double* blk, const BlockLayout& bl, int c) {
        const double* kin = blk + bl.kin;
        double* o = blk + bl.kappa;
        const double o9 = raw2[2 * c], o10 = raw2[2 * c + 1];
        const double Mg = blk[4], Ms = blk[6];
        const double* gk = kin;        // gas: com 3, vcom 3, L 3, veldisp 6
        const double* sk = kin + 30;   // stars
        const double trg = gk[9] + gk[10] + gk[11], trs = sk[9] + sk[10] + sk[11];
        const double Kg = 0.5 * Mg * trg, Ks = 0.5 * Ms * trs;
        double Kb = 0.0;
        if (Mg + Ms != 0.0) {
            double dg = 0.0, ds = 0.0;
            for (int d = 0; d < 3; d++) {
                const double vb = (Mg * gk[3 + d] + Ms * sk[3 + d]) / (Mg + Ms);
                dg += (gk[3 + d] - vb) * (gk[3 + d] - vb);
                ds += (sk[3 + d] - vb) * (sk[3 + d] - vb);
            }
            Kb = 0.5 * (Mg * (trg + dg) + Ms * (trs + ds));
        }
        const double kc_g = o[0], kc_s = o[1], kc_b = o[2], mc_g = o[3], mc_s = o[4];
        o[0] = Kg > 0.0 ? kc_g / Kg : 0.0;
        o[1] = Ks > 0.0 ? kc_s / Ks : 0.0;
        o[2] = Kb > 0.0 ? kc_b / Kb : 0.0;
        o[3] = Mg != 0.0 ? 1.0 - 2.0 * mc_g / Mg : 0.0;
        o[4] = Ms != 0.0 ? 1.0 - 2.0 * mc_s / Ms : 0.0;
        // stellar rotation and cylindrical dispersions (kinematic_properties.py:17-51,130-178;
        // aperture_properties.py:1502-1536): mean v_phi, sqrt(sum sigma^2 / 3), sigma_z, sqrt(sigma_r^2 + sigma_phi^2)
        {
            double mean[3], var[3];
            const bool have = Ms != 0.0 && (o[5] != 0.0 || o[6] != 0.0 || o[7] != 0.0 || o[8] != 0.0 || o9 != 0.0 || o10 != 0.0);
            for (int c = 0; c < 3; c++) {
                mean[c] = have ? o[5 + c] / Ms : 0.0;
                const double s2 = c == 0 ? o[8] : c == 1 ? o9 : o10;
                var[c] = have ? fmax(s2 / Ms - mean[c] * mean[c], 0.0) : 0.0;
            }
            o[5] = mean[1];
            o[6] = sqrt((var[0] + var[1] + var[2]) / 3.0);
            o[7] = sqrt(var[2]);
            o[8] = sqrt(var[0] + var[1]);
        }
    };
    if (cfg.do_sub && c_lo == 0) fin(row + cfg.lay.sub, cfg.lay.bsub, 0);
    for (int q = 0; q < cfg.n_so; q++)
        if ((cfg.do_sub ? 1 : 0) + q >= c_lo && (cfg.do_sub ? 1 : 0) + q < c_hi && ha.sres[h].so_exists[q])
            fin(row + cfg.lay.so[q], cfg.lay.bso, (cfg.do_sub ? 1 : 0) + q);
    for (int a = 0; a < cfg.n_ap; a++)
        if (off_ap + a >= c_lo && off_ap + a < c_hi && ((ha.sres[h].ap_on >> a) & 1u)) fin(row + cfg.lay.ap[a], cfg.lay.bap, off_ap + a);
}

#endif  // __CUDACC__
