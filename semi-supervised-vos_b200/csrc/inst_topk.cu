// top-k extension: vos_topk_scan<split, pass>, vos_topk_threshold, vos_topk_finish, vos_upsample_mask
#include "launch.h"
#include "affinity_topk.cuh"

namespace vosk {

cudaError_t launch_topk_scan(bool split, int pass, int grid, cudaStream_t st, const CUtensorMap& tmap_hi, const CUtensorMap& tmap_lo,
                             const AffinityParams& prm) {
    void (*kern)(CUtensorMap, CUtensorMap, AffinityParams) =
        pass == 1 ? (split ? vos_topk_scan<true, 1> : vos_topk_scan<false, 1>) : (split ? vos_topk_scan<true, 2> : vos_topk_scan<false, 2>);
    cudaError_t ce = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kIdxSmem);
    if (ce != cudaSuccess) return ce;
    return launch_pdl(kern, grid, kIdxThreads, kIdxSmem, st, tmap_hi, tmap_lo, prm);
}

// rows per CTA: 8 (whole 32-byte sectors of the block maxima) or fewer when the row of keys is long
cudaError_t launch_topk_threshold(const float* bound, float* tau, int n_pixels, int n_tiles, int tile_step, int k, cudaStream_t st) {
    const size_t budget = 200 * 1024, hist = kThrWarps * 256 * 4;
    const int n_blocks = (n_tiles + tile_step - 1) / tile_step * kIdxSub;
    int rows_log2 = 3;
    while (rows_log2 > 0 && (static_cast<size_t>(n_blocks | 1) << rows_log2) * 4 + hist > budget) --rows_log2;
    const size_t smem = (static_cast<size_t>(n_blocks | 1) << rows_log2) * 4 + hist;
    if (smem > budget) return cudaErrorInvalidValue;
    cudaError_t ce = cudaFuncSetAttribute(vos_topk_threshold, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(budget));
    if (ce != cudaSuccess) return ce;
    const int rows = 1 << rows_log2;
    vos_topk_threshold<<<(n_pixels + rows - 1) / rows, kThrWarps * 32, smem, st>>>(bound, tau, n_pixels, n_tiles, tile_step, k, rows_log2);
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
