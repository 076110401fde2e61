// moments.cu -- masked moment sums of every accepted halo and its result row.
//
// Every selection of the reference's property classes is a radial cut on the
// halo-centred radius, optionally restricted to bound particles and/or a
// particle type:
//   SO                r <  R_SO          all particles (SO_properties.py:485-489)
//   Exclusive sphere  r <= R_ap, bound   (aperture_properties.py:285-288,310)
//   Inclusive sphere  r <= R_ap
//   BoundSubhalo      bound, no cut      (subhalo_properties.py:144)
// The cuts of one halo are nested, so a particle belongs to exactly one
// (shell, bound, type) bank; it is added there once and every selection is a
// prefix sum over shells at the end.  Raw moments about the halo centre are
// accumulated in float64 and the reference's central moments (velocity
// dispersion about vcom, L about vcom, kinetic energy with Hubble flow) are
// derived from them:
//   com, vcom     aperture_properties.py:1098-1126, SO_properties.py:557-571
//   L             kinematic_properties.py:222-263
//   veldisp       kinematic_properties.py:91-127
//   Ekin          subhalo_properties.py:848-858
//   spin          SO_properties.py:602-618, subhalo_properties.py:1049-1073
//   concentration SO_properties.py:2724-2790
//   tensors       inertia_tensors.py:19-132 with max_iterations=1
#include "moments.cuh"

namespace {

constexpr int TB = SWEEP_NT;

template <int V, int NTY>
__global__ void __launch_bounds__(TB, (V <= 16 && NTY == 1) ? 3 : 1) k_moments(ChunkView v, HaloArrays ha, DevCfg cfg,
                                                const Item* __restrict__ items,
                                                const unsigned int* __restrict__ n_items_dev,
                                                double* __restrict__ gbanks, int gbank_stride, int priv) {
    // dynamic shared memory: [priv ? NW : 1][gbank_stride] banks, then the
    // per-warp staging tiles [NW][32][VP] and their keys [NW][32]
    extern __shared__ __align__(16) unsigned char smem_raw[];
    constexpr int NW = TB / 32;
    constexpr int VP = BankAcc<V>::VP;
    double* banks = (double*)smem_raw;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    double* stage_w = banks + (size_t)(priv ? NW : 1) * gbank_stride + (size_t)wid * 32 * VP;
    int* skey_w = (int*)(banks + (size_t)(priv ? NW : 1) * gbank_stride + (size_t)NW * 32 * VP) + wid * 32;
    double* bank_w = banks + (priv ? (size_t)wid * gbank_stride : 0);
    __shared__ SweepShared SW;
    __shared__ Cuts cuts;
    __shared__ int s_last;
    const unsigned int n_items = *n_items_dev;
    for (unsigned int it = blockIdx.x; it < n_items; it += gridDim.x) {
        const Item im = items[it];
        const uint32_t h = im.halo;
        // properties [lo, hi) of halo_prop_list were computed at this rung
        const int c_lo = ha.commit_lo[h], c_hi = ha.commit_hi[h];
        if (c_hi <= c_lo || ha.status[h] >= 2) continue;
        const bool sub_c = cfg.do_sub && c_lo == 0;
        const ScanRes* sr = ha.sres + h;
        const double cx = ha.cofp[3 * h], cy = ha.cofp[3 * h + 1], cz = ha.cofp[3 * h + 2];
        const double R = ha.rung_r[h];
        const double r2max = __dmul_rn(R, R), halfL = 0.5 * v.L, L = v.L;
        const int32_t hidx = (int32_t)ha.index[h];
        const bool central = ha.central[h] == 1;
        const int n_so = central ? cfg.n_so : 0;
        __syncthreads();
        if (threadIdx.x == 32) build_cuts(cuts, cfg, sr, c_lo, c_hi, n_so);
        __syncthreads();
        const int ncut = cuts.n;
        const int nbank = (ncut + 1) * 2 * NTY;
        if (priv) {
            for (int w = 0; w < NW; w++)
                for (int i = threadIdx.x; i < nbank * V; i += TB) banks[(size_t)w * gbank_stride + i] = 0.0;
        } else {
            for (int i = threadIdx.x; i < nbank * V; i += TB) banks[i] = 0.0;
        }
        __syncthreads();
        const int32_t cen_fof = sr->cen_fof;
        BankAcc<V> ba;
        ba.init();
        sweep_item(v, SW, cx, cy, cz, R, im, [&](uint32_t t, bool ok) {
            bool in = false;
            int key = 0;
            double val[V];
            if (ok) {
                double r2 = periodic_r2(v.px[t], v.py[t], v.pz[t], cx, cy, cz, L, halfL);
                in = r2 <= r2max;
            }
            if (in) {
                const double x = rewrap_rel(v.px[t], cx, L, halfL);
                const double y = rewrap_rel(v.py[t], cy, L, halfL);
                const double z = rewrap_rel(v.pz[t], cz, L, halfL);
                const double r = radius3(x, y, z);
                key = moment_terms<V, NTY>(cuts, ncut, cfg, x, y, z, r, (double)v.mass[t], (double)v.vx[t],
                                           (double)v.vy[t], (double)v.vz[t], v.grnr[t], hidx, v.fof[t], cen_fof,
                                           NTY == 1 ? 1u : (uint32_t)v.type[t], val);
            }
            ba.add(in, key, val, stage_w, skey_w, bank_w, priv, lane);
        });
        ba.flush(bank_w, priv, lane);
        __syncthreads();
        if (priv) {
            for (int i = threadIdx.x; i < nbank * V; i += TB) {
                double s = banks[i];
                for (int w = 1; w < NW; w++) s += banks[(size_t)w * gbank_stride + i];
                banks[i] = s;
            }
            __syncthreads();
        }
        // halos swept by several work items: combine in global banks; the last
        // item to arrive writes the result row
        const uint32_t n_it = ha.n_items[h];
        if (n_it > 1) {
            double* gb = gbanks + (size_t)ha.mslot[h] * gbank_stride;
            for (int i = threadIdx.x; i < nbank * V; i += TB)
                if (banks[i] != 0.0) atomicAdd(&gb[i], banks[i]);
            __threadfence();
            __syncthreads();
            if (threadIdx.x == 0) s_last = (atomicAdd(&ha.items_done[h], 1u) == n_it - 1) ? 1 : 0;
            __syncthreads();
            if (!s_last) continue;
            __threadfence();
            for (int i = threadIdx.x; i < nbank * V; i += TB) banks[i] = __ldcg(&gb[i]);
            __syncthreads();
        }
        write_row<V, NTY>(banks, cuts, ncut, cfg, ha, h, sr, sub_c, n_so, cx, cy, cz, (int)threadIdx.x);
        __syncthreads();
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
    int incl, is_sub;
    double* out;  // the block's 11 kappa / rotation slots
};

__device__ inline void kappa_refs(KapSel& k, double* blk, const BlockLayout& bl) {
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

__global__ void __launch_bounds__(TB) k_kappa(ChunkView v, HaloArrays ha, DevCfg cfg, const Item* __restrict__ items,
                                              const unsigned int* __restrict__ n_items_dev) {
    __shared__ SweepShared SW;
    __shared__ KapSel sel[1 + SOAP_MAX_APERTURES];
    __shared__ double acc[1 + SOAP_MAX_APERTURES][11];
    __shared__ int nsel;
    const unsigned int n_items = *n_items_dev;
    for (unsigned int it = blockIdx.x; it < n_items; it += gridDim.x) {
        const Item im = items[it];
        const uint32_t h = im.halo;
        const int c_lo = ha.commit_lo[h], c_hi = ha.commit_hi[h];
        if (c_hi <= c_lo || ha.status[h] >= 2) continue;
        const int off_ap = (cfg.do_sub ? 1 : 0) + cfg.n_so;
        const double cx = ha.cofp[3 * h], cy = ha.cofp[3 * h + 1], cz = ha.cofp[3 * h + 2];
        const double R = ha.rung_r[h];
        const double r2max = __dmul_rn(R, R), halfL = 0.5 * v.L, L = v.L;
        const int32_t hidx = (int32_t)ha.index[h];
        __syncthreads();
        if (threadIdx.x == 0) {
            double* row = ha.out + (int64_t)h * ha.ncol;
            int n = 0;
            if (cfg.do_sub && c_lo == 0) {
                kappa_refs(sel[n], row + cfg.lay.sub, cfg.lay.bsub);
                sel[n].is_sub = 1; sel[n].incl = 0; sel[n].R = 0.0;
                n++;
            }
            for (int a = 0; a < cfg.n_ap; a++)
                if (off_ap + a >= c_lo && off_ap + a < c_hi) {
                    kappa_refs(sel[n], row + cfg.lay.ap[a], cfg.lay.bap);
                    sel[n].is_sub = 0; sel[n].incl = cfg.ap_incl[a]; sel[n].R = cfg.ap_r[a];
                    n++;
                }
            nsel = n;
        }
        for (int i = threadIdx.x; i < (1 + SOAP_MAX_APERTURES) * 11; i += TB) (&acc[0][0])[i] = 0.0;
        __syncthreads();
        const int ns = nsel;
        if (ns == 0) continue;
        sweep_item(v, SW, cx, cy, cz, R, im, [&](uint32_t t, bool ok) {
            if (!ok) return;
            const uint32_t tc = (uint32_t)v.type[t];
            if (tc != 0u && tc != 2u) return;  // gas and stars only
            const double r2 = periodic_r2(v.px[t], v.py[t], v.pz[t], cx, cy, cz, L, halfL);
            if (!(r2 <= r2max)) return;
            const double x = rewrap_rel(v.px[t], cx, L, halfL);
            const double y = rewrap_rel(v.py[t], cy, L, halfL);
            const double z = rewrap_rel(v.pz[t], cz, L, halfL);
            const double r = radius3(x, y, z);
            const bool bound = v.grnr[t] == hidx;
            const double m = (double)v.mass[t];
            const double vx = (double)v.vx[t], vy = (double)v.vy[t], vz = (double)v.vz[t];
            const double rr2 = x * x + y * y + z * z;
            const int g0 = tc == 0u ? 0 : 1;
            for (int s = 0; s < ns; s++) {
                const KapSel& k = sel[s];
                const bool in = k.is_sub ? bound : (r <= k.R && (k.incl || bound));
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
        });
        __syncthreads();
        for (int i = threadIdx.x; i < ns * 11; i += TB) {
            const double a = acc[i / 11][i % 11];
            if (a != 0.0) atomicAdd(&sel[i / 11].out[i % 11], a);
        }
        __syncthreads();
    }
}

// raw sums -> kappa_corot / DtoT, once per selection, at the rung that committed it
__global__ void k_kappa_finish(HaloArrays ha, DevCfg cfg, const uint32_t* __restrict__ list,
                               const unsigned int* __restrict__ n_list) {
    unsigned int it = blockIdx.x * blockDim.x + threadIdx.x;
    if (it >= *n_list) return;
    const uint32_t h = list[it];
    const int c_lo = ha.commit_lo[h], c_hi = ha.commit_hi[h];
    if (c_hi <= c_lo || ha.status[h] >= 2) return;
    const int off_ap = (cfg.do_sub ? 1 : 0) + cfg.n_so;
    double* row = ha.out + (int64_t)h * ha.ncol;
    auto fin = [&](double* blk, const BlockLayout& bl) {
        const double* kin = blk + bl.kin;
        double* o = blk + bl.kappa;
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
            const bool have = Ms != 0.0 && (o[5] != 0.0 || o[6] != 0.0 || o[7] != 0.0 || o[8] != 0.0 || o[9] != 0.0 || o[10] != 0.0);
            for (int c = 0; c < 3; c++) {
                mean[c] = have ? o[5 + c] / Ms : 0.0;
                var[c] = have ? fmax(o[8 + c] / Ms - mean[c] * mean[c], 0.0) : 0.0;
            }
            o[5] = mean[1];
            o[6] = sqrt((var[0] + var[1] + var[2]) / 3.0);
            o[7] = sqrt(var[2]);
            o[8] = sqrt(var[0] + var[1]);
            o[9] = 0.0;
            o[10] = 0.0;
        }
    };
    if (cfg.do_sub && c_lo == 0) fin(row + cfg.lay.sub, cfg.lay.bsub);
    for (int a = 0; a < cfg.n_ap; a++)
        if (off_ap + a >= c_lo && off_ap + a < c_hi) fin(row + cfg.lay.ap[a], cfg.lay.bap);
}

__global__ void k_write_input_cols(HaloArrays ha, int64_t nh) {
    int64_t h = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (h >= nh) return;
    double* row = ha.out + h * ha.ncol;
    row[0] = (double)ha.status[h];
    row[1] = (double)ha.nloop[h];
    row[2] = ha.cur_r[h];
}

}  // namespace

int soap_launch_moments(soap_chunk* c, const DevCfg& cfg, const HaloArrays& ha, const Item* items,
                        const unsigned int* n_items_dev, unsigned int n_items_host,
                        unsigned int n_mslot, unsigned int grid, cudaStream_t stream) {
    soap_handle* h = c->h;
    const bool full = (cfg.flags & (PF_KIN | PF_KAPPA | PF_TENS)) != 0;
    const int nty = cfg.dmo ? 1 : 4;
    const int V = full ? V_FULL : V_MIN;
    const int stride = (cfg.n_so + cfg.n_ap + 3) * 2 * nty * V;
    const int NW = TB / 32, VP = V | 1;
    const size_t stage_bytes = (size_t)NW * 32 * VP * sizeof(double) + (size_t)NW * 32 * sizeof(int);
    // warp-private banks when they fit next to the staging tiles
    const int priv = ((size_t)NW * stride * sizeof(double) + stage_bytes <= 96 * 1024) ? 1 : 0;
    const size_t smem = (size_t)(priv ? NW : 1) * stride * sizeof(double) + stage_bytes;
    if (smem > 220 * 1024) SOAP_FAIL("soap_process_halos: %d SO + %d aperture variations need %zu bytes of shared memory", cfg.n_so, cfg.n_ap, smem);
    double* gbanks = (double*)h->get("h_gbanks", sizeof(double) * (size_t)stride * (n_mslot + 1));
    if (!gbanks) return -1;
    if (n_mslot > 0) CUDA_TRY(cudaMemsetAsync(gbanks, 0, sizeof(double) * (size_t)stride * n_mslot, stream));
    unsigned int g = n_items_host < grid ? n_items_host : grid;
    if (g < 1) g = 1;
#define MOM(VV, NT)                                                                                   \
    do {                                                                                              \
        CUDA_TRY(cudaFuncSetAttribute(k_moments<VV, NT>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                                      (int)smem));                                                    \
        LAUNCH(h, (k_moments<VV, NT>), g, TB, smem, stream, c->v, ha, cfg, items, n_items_dev, gbanks, \
               stride, priv);                                                                         \
    } while (0)
    if (full && nty == 4) MOM(V_FULL, 4);
    else if (full) MOM(V_FULL, 1);
    else if (nty == 4) MOM(V_MIN, 4);
    else MOM(V_MIN, 1);
#undef MOM
    return 0;
}

int soap_launch_kappa(soap_chunk* c, const DevCfg& cfg, const HaloArrays& ha, const Item* items,
                      const unsigned int* n_items_dev, unsigned int n_items_host, const uint32_t* acc_list,
                      const unsigned int* n_acc_dev, unsigned int n_acc_host, unsigned int grid,
                      cudaStream_t stream) {
    soap_handle* h = c->h;
    if (!(cfg.flags & PF_KAPPA) || cfg.dmo) return 0;  // no gas or stars: every kappa slot stays zero
    if (!(cfg.flags & PF_KIN)) SOAP_FAIL("soap_process_halos: kappa_corot needs the kinematics group (property_flags bit 0)");
    unsigned int g = n_items_host < grid ? n_items_host : grid;
    if (g < 1) g = 1;
    LAUNCH(h, k_kappa, g, TB, 0, stream, c->v, ha, cfg, items, n_items_dev);
    LAUNCH(h, k_kappa_finish, grid_for(n_acc_host, 128), 128, 0, stream, ha, cfg, acc_list, n_acc_dev);
    return 0;
}

int soap_write_input_cols(soap_handle* h, const HaloArrays& ha, int64_t nh, cudaStream_t stream) {
    LAUNCH(h, k_write_input_cols, grid_for(nh, 128), 128, 0, stream, ha, nh);
    return 0;
}
