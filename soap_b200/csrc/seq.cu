// seq.cu -- thread-per-halo scan + solve kernel (seq.cuh) for halos of up to a few thousand records.
#include "seq.cuh"

namespace {

template <int NCH>
__global__ void __launch_bounds__(SEQ_NT) k_solve_seq(HaloArrays ha, DevCfg cfg, const uint32_t* __restrict__ list,
                                                  const unsigned int* __restrict__ n_list,
                                                  const Rec* __restrict__ recs, uint32_t* __restrict__ next,
                                                  unsigned int* __restrict__ n_next, Counters* ctr, const unsigned long long* __restrict__ item_minr,
                                                  const int32_t* __restrict__ item_minfof, int multi) {
    extern __shared__ __align__(16) uint4 seq_slots[];
    const unsigned int it = blockIdx.x * blockDim.x + threadIdx.x;
    if (it >= *n_list) return;
    const uint32_t h = list[it];
    const uint32_t ib = ha.item_base[h];
    solve_seq_halo<NCH>(ha, cfg, h, ha.cnt[h], recs + ha.rec_off[h], next, n_next, ctr, item_minr + ib, item_minfof + ib,
                        ha.n_items[h], ha.cuts, seq_slots + threadIdx.x, multi != 0);
}

}  // namespace

int soap_launch_solve_seq(soap_chunk* c, const DevCfg& cfg, const HaloArrays& ha, const uint32_t* list,
                          const unsigned int* n_list_dev, unsigned int n_list_host, const Rec* recs, uint32_t* next,
                          unsigned int* n_next, Counters* ctr, const unsigned long long* item_minr, const int32_t* item_minfof,
                          int multi, cudaStream_t stream) {
    soap_handle* h = c->h;
    if (n_list_host == 0) return 0;
    if (cfg.dmo)
        LAUNCH(h, k_solve_seq<2>, grid_for(n_list_host, SEQ_NT), SEQ_NT, SEQ_SMEM, stream, ha, cfg, list, n_list_dev, recs, next, n_next, ctr,
               item_minr, item_minfof, multi);
    else
        LAUNCH(h, k_solve_seq<8>, grid_for(n_list_host, SEQ_NT), SEQ_NT, SEQ_SMEM, stream, ha, cfg, list, n_list_dev, recs, next, n_next, ctr,
               item_minr, item_minfof, multi);
    return 0;
}
