// vos_affinity_topk<split, n_sub>
#include "launch.h"
#include "affinity_topk.cuh"

namespace vosk {

cudaError_t launch_affinity_topk(bool split, int n_sub, int grid, cudaStream_t st, const CUtensorMap& tmap_hi, const CUtensorMap& tmap_lo,
                                 const AffinityParams& prm) {
    void (*kern)(CUtensorMap, CUtensorMap, AffinityParams) =
        n_sub == 4 ? (split ? vos_affinity_topk<true, 4> : vos_affinity_topk<false, 4>)
      : n_sub == 2 ? (split ? vos_affinity_topk<true, 2> : vos_affinity_topk<false, 2>)
                   : (split ? vos_affinity_topk<true, 1> : vos_affinity_topk<false, 1>);
    const int smem = n_sub == 4 ? TopkCfg<4>::kSmem : (n_sub == 2 ? TopkCfg<2>::kSmem : TopkCfg<1>::kSmem);
    const int threads = n_sub == 4 ? TopkCfg<4>::kThreads : (n_sub == 2 ? TopkCfg<2>::kThreads : TopkCfg<1>::kThreads);
    cudaError_t ce = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (ce != cudaSuccess) return ce;
    kern<<<grid, threads, smem, st>>>(tmap_hi, tmap_lo, prm);
    return cudaGetLastError();
}

cudaError_t launch_topk_finish(const TopkFinishParams& fp, cudaStream_t st) {
    cudaError_t ce = cudaFuncSetAttribute(vos_topk_finish, cudaFuncAttributeMaxDynamicSharedMemorySize, kFinishSmem);
    if (ce != cudaSuccess) return ce;
    vos_topk_finish<<<(fp.mp.n_pixels + kFinishWarps - 1) / kFinishWarps, kFinishWarps * 32, kFinishSmem, st>>>(fp);
    return cudaGetLastError();
}

cudaError_t launch_upsample_mask(const uint8_t* low, uint8_t* out, int h_lowres, int w_lowres, int H, int W, cudaStream_t st) {
    vos_upsample_mask<<<H, 256, 0, st>>>(low, out, h_lowres, w_lowres, H, W);
    return cudaGetLastError();
}

}  // namespace vosk
