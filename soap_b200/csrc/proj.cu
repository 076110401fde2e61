// proj.cu -- projected apertures (ProjectedApertureProperties,
// SOAP/particle_selection/projected_aperture_properties.py:98-198, 339-374,
// 738-875, 953-986, 1491-1577).
//
// Bound particles only (:115), three projections at once: for the projection
// along axis a the radius is that of the two perpendicular components (:122-127)
// and the mask is rproj <= R (:140-142).  The radii of one halo are nested, so a
// particle belongs, per axis, to exactly one shell; it is added there once and
// every aperture is a prefix sum over shells.  One sweep of the halo's sphere
// serves all radii and all three axes (48 B per pair, one read).
//   masses, counts, com, vcom           :339-374, 738-786
//   proj_veldisp_{gas,dm,star}          :865-875   sqrt(sum mf (v_a - vcom_a)^2)
//   ProjectedTotalInertiaTensor[Reduced]Noniterative
//                                       :789-852 with inertia_tensors.py:226-343
//                                       (max_iterations = 1: circle of the aperture
//                                       radius, all bound particles, >= 20 inside)
#include "moments.cuh"

namespace {

constexpr int TB = SWEEP_NT;
// per (axis, shell, type) bank
enum { P_N = 0, P_M, P_MX, P_MV = 5, P_MVA = 8, P_MVVA, P_XX, P_XXR = 13, P_M0 = 16, P_N0, PV = 18 };

template <int NTY>
__global__ void __launch_bounds__(TB) k_projected(ChunkView v, HaloArrays ha, DevCfg cfg,
                                                  const Item* __restrict__ items,
                                                  const unsigned int* __restrict__ n_items_dev,
                                                  double* __restrict__ gbanks, int stride, int priv) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    constexpr int NW = TB / 32;
    constexpr int VP = BankAcc<PV>::VP;
    double* banks = (double*)smem_raw;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    double* stage_w = banks + (size_t)(priv ? NW : 1) * stride + (size_t)wid * 32 * VP;
    int* skey_w = (int*)(banks + (size_t)(priv ? NW : 1) * stride + (size_t)NW * 32 * VP) + wid * 32;
    double* bank_w = banks + (priv ? (size_t)wid * stride : 0);
    __shared__ SweepShared SW;
    __shared__ int s_last;
    const unsigned int n_items = *n_items_dev;
    const int npj = cfg.n_pj, nsh = npj + 1;  // shell npj = outside the largest aperture
    const int nbank = 3 * nsh * NTY;
    const int off_pj = (cfg.do_sub ? 1 : 0) + cfg.n_so + cfg.n_ap;
    for (unsigned int it = blockIdx.x; it < n_items; it += gridDim.x) {
        const Item im = items[it];
        const uint32_t h = im.halo;
        // every projected aperture is committed together (they never fail)
        const int c_lo = ha.commit_lo[h], c_hi = ha.commit_hi[h];
        if (c_hi <= c_lo || ha.status[h] >= 2 || c_hi <= off_pj || c_lo > off_pj) continue;
        const double cx = ha.cofp[3 * h], cy = ha.cofp[3 * h + 1], cz = ha.cofp[3 * h + 2];
        const double R = ha.rung_r[h];
        const double r2max = __dmul_rn(R, R), halfL = 0.5 * v.L, L = v.L;
        const int64_t hidx = ha.index[h];
        __syncthreads();
        for (int w = 0; w < (priv ? NW : 1); w++)
            for (int i = threadIdx.x; i < nbank * PV; i += TB) banks[(size_t)w * stride + i] = 0.0;
        __syncthreads();
        BankAcc<PV> ba;
        ba.init();
        sweep_item<SW_MASS | SW_VEL | SW_IDS | SW_TYPE>(v, SW, cx, cy, cz, R, im, [&](uint32_t t, bool ok) {
            bool in = false;
            double x = 0, y = 0, z = 0, m = 0, vx = 0, vy = 0, vz = 0;
            int tcode = 0;
            if (ok && v.grnr[t] == hidx) {  // bound only
                const double r2 = periodic_r2(v.px[t], v.py[t], v.pz[t], cx, cy, cz, L, halfL);
                in = r2 <= r2max;
            }
            if (in) {
                x = rewrap_rel(v.px[t], cx, L, halfL);
                y = rewrap_rel(v.py[t], cy, L, halfL);
                z = rewrap_rel(v.pz[t], cz, L, halfL);
                m = (double)v.mass[t];
                vx = (double)v.vx[t]; vy = (double)v.vy[t]; vz = (double)v.vz[t];
                tcode = NTY == 1 ? 0 : (int)v.type[t];
            }
#pragma unroll
            for (int ax = 0; ax < 3; ax++) {
                int key = 0;
                double val[PV];
                if (in) {
                    // in-plane components in the reference's order (inertia_tensors.py:272-280):
                    // axis 0 -> (y, z), 1 -> (z, x), 2 -> (x, y)
                    const double pa = ax == 0 ? y : (ax == 1 ? z : x);
                    const double pb = ax == 0 ? z : (ax == 1 ? x : y);
                    const double va = ax == 0 ? vx : (ax == 1 ? vy : vz);
                    // projected_aperture_properties.py:122-127: sqrt of the two perpendicular squares
                    const double q0 = ax == 0 ? y : x, q1 = ax == 2 ? y : z;
                    const double rp = sqrt(__dadd_rn(__dmul_rn(q0, q0), __dmul_rn(q1, q1)));
                    int shell = 0;
                    for (int k = 0; k < npj; k++) shell += !(rp <= cfg.pj_r[k]);
                    key = (ax * nsh + shell) * NTY + tcode;
#pragma unroll
                    for (int i = 0; i < PV; i++) val[i] = 0.0;
                    val[P_N] = 1.0;
                    val[P_M] = m;
                    val[P_MX] = m * x; val[P_MX + 1] = m * y; val[P_MX + 2] = m * z;
                    val[P_MV] = m * vx; val[P_MV + 1] = m * vy; val[P_MV + 2] = m * vz;
                    val[P_MVA] = m * va;
                    val[P_MVVA] = m * va * va;
                    val[P_XX] = m * pa * pa; val[P_XX + 1] = m * pb * pb; val[P_XX + 2] = m * pa * pb;
                    const double nrm = pa * pa + pb * pb;
                    if (nrm <= 1e-8) {  // np.isclose(norm, 0): inertia_tensors.py:283-285
                        val[P_M0] = m; val[P_N0] = 1.0;
                    } else {
                        const double w = m / nrm;
                        val[P_XXR] = w * pa * pa; val[P_XXR + 1] = w * pb * pb; val[P_XXR + 2] = w * pa * pb;
                    }
                }
                ba.add(in, key, val, stage_w, skey_w, bank_w, priv, lane);
            }
        });
        ba.flush(bank_w, priv, lane);
        __syncthreads();
        if (priv) {
            for (int i = threadIdx.x; i < nbank * PV; i += TB) {
                double s = banks[i];
                for (int w = 1; w < NW; w++) s += banks[(size_t)w * stride + i];
                banks[i] = s;
            }
            __syncthreads();
        }
        const uint32_t n_it = ha.n_items[h];
        if (n_it > 1) {
            double* gb = gbanks + (size_t)ha.mslot[h] * stride;
            for (int i = threadIdx.x; i < nbank * PV; i += TB)
                if (banks[i] != 0.0) atomicAdd(&gb[i], banks[i]);
            __threadfence();
            __syncthreads();
            if (threadIdx.x == 0) s_last = (atomicAdd(&ha.items_done[h], 1u) == n_it - 1) ? 1 : 0;
            __syncthreads();
            if (!s_last) continue;
            __threadfence();
            for (int i = threadIdx.x; i < nbank * PV; i += TB) banks[i] = __ldcg(&gb[i]);
            __syncthreads();
        }
        // ------------------------------------------------- rows: one thread per (radius, axis)
        if ((int)threadIdx.x < npj * 3 && ((ha.sres[h].pj_on >> (threadIdx.x / 3)) & 1u)) {
            const int p = threadIdx.x / 3, ax = threadIdx.x % 3;
            double* blk = ha.out + (int64_t)h * ha.ncol + cfg.lay.pj[p] + ax * cfg.lay.pjb;
            double S[4][PV], T[PV];
            double n_all = 0.0;
            for (int t = 0; t < 4; t++)
                for (int i = 0; i < PV; i++) S[t][i] = 0.0;
            for (int s = 0; s < nsh; s++)
                for (int t = 0; t < NTY; t++) {
                    const double* bk = banks + (size_t)((ax * nsh + s) * NTY + t) * PV;
                    n_all += bk[P_N];
                    if (s <= p)
                        for (int i = 0; i < PV; i++) S[NTY == 1 ? 1 : t][i] += bk[i];
                }
            for (int i = 0; i < PV; i++) T[i] = S[0][i] + S[1][i] + S[2][i] + S[3][i];
            for (int t = 0; t < 4; t++) { blk[t] = S[t][P_N]; blk[4 + t] = S[t][P_M]; }
            const double Mtot = T[P_M];
            blk[8] = Mtot;
            if (Mtot != 0.0) {
                const double centre[3] = {cx, cy, cz};
                for (int d = 0; d < 3; d++) {
                    blk[9 + d] = floored_mod(T[P_MX + d] / Mtot + centre[d], cfg.L);
                    blk[12 + d] = T[P_MV + d] / Mtot;
                }
            }
            const int gt[3] = {0, 1, 2};
            for (int g = 0; g < 3; g++) {
                const double* s = S[gt[g]];
                if (s[P_M] > 0.0) {
                    const double vc = s[P_MVA] / s[P_M];
                    const double var = s[P_MVVA] / s[P_M] - vc * vc;
                    blk[15 + g] = var > 0.0 ? sqrt(var) : 0.0;
                }
            }
            // tensors over all bound particles, particles inside the aperture weighted
            if (Mtot != 0.0 && n_all >= 20.0) {
                const double k2 = cfg.kpc * cfg.kpc;
                if (T[P_N] >= 20.0) {
                    blk[21] = T[P_XX] * k2 / Mtot; blk[22] = T[P_XX + 1] * k2 / Mtot; blk[23] = T[P_XX + 2] * k2 / Mtot;
                }
                const double nred = T[P_N] - T[P_N0], mred = T[P_M] - T[P_M0];
                if (nred >= 20.0 && mred != 0.0) {
                    blk[24] = T[P_XXR] / mred; blk[25] = T[P_XXR + 1] / mred; blk[26] = T[P_XXR + 2] / mred;
                }
            }
        }
        __syncthreads();
    }
}

}  // namespace

int soap_launch_projected(soap_chunk* c, const DevCfg& cfg, const HaloArrays& ha, const Item* items,
                          const unsigned int* n_items_dev, unsigned int n_items_host, unsigned int n_mslot,
                          unsigned int grid, cudaStream_t stream) {
    if (cfg.n_pj <= 0) return 0;
    soap_handle* h = c->h;
    const int nty = cfg.dmo ? 1 : 4;
    const int stride = 3 * (cfg.n_pj + 1) * nty * PV;
    const int NW = TB / 32, VP = BankAcc<PV>::VP;
    const size_t stage_bytes = (size_t)NW * 32 * VP * sizeof(double) + (size_t)NW * 32 * sizeof(int);
    const int priv = ((size_t)NW * stride * sizeof(double) + stage_bytes <= 128 * 1024) ? 1 : 0;
    const size_t smem = (size_t)(priv ? NW : 1) * stride * sizeof(double) + stage_bytes;
    if (smem > 220 * 1024) SOAP_FAIL("soap_process_halos: %d projected apertures need %zu bytes of shared memory", cfg.n_pj, smem);
    double* gbanks = (double*)h->get("h_gbanks_pj", sizeof(double) * (size_t)stride * (n_mslot + 1));
    if (!gbanks) return -1;
    if (n_mslot > 0) CUDA_TRY(cudaMemsetAsync(gbanks, 0, sizeof(double) * (size_t)stride * n_mslot, stream));
    unsigned int g = n_items_host < grid ? n_items_host : grid;
    if (g < 1) g = 1;
    if (nty == 4) {
        CUDA_TRY(cudaFuncSetAttribute(k_projected<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        LAUNCH(h, k_projected<4>, g, TB, smem, stream, c->v, ha, cfg, items, n_items_dev, gbanks, stride, priv);
    } else {
        CUDA_TRY(cudaFuncSetAttribute(k_projected<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        LAUNCH(h, k_projected<1>, g, TB, smem, stream, c->v, ha, cfg, items, n_items_dev, gbanks, stride, priv);
    }
    return 0;
}
