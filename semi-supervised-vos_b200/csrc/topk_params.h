// Shapes and parameter block of the top-k extension shared by the host code (vos_prop.cu) and the kernels (affinity_topk.cuh).
#pragma once
#include "kernels.cuh"

namespace vosk {

constexpr int kTopkMax = 64;                   // largest supported k
constexpr int kTopkCap = kTopkMax + kQCap;     // slots of one thread's candidate list: pruned to k when fewer than 16 are free
constexpr int kTopkMaxCand = 2048;             // candidates merged per target pixel by the finish kernel (lists x k)

struct TopkFinishParams {
    MergeParams mp;               // geometry + outputs shared with vos_merge_writeback
    int32_t topk;
    int32_t ref_slot[32];
    float ref_coef[32];           // log2(e) / sigma^2 ; 0 = no prior
    const uint32_t* cand_key;     // candidate lists of vos_topk_scan<pass 2>: slot e of record rec at topk_slot(rec, e)
    const int32_t* cand_idx;
    const int32_t* cand_cnt;      // [records]: entries of each list (<= k)
    int32_t* out_topk_idx;        // (P, topk) int32 or null; value-descending, -1 where fewer than k references exist
};

constexpr int kFinishWarps = 4;
constexpr int kFinishSmem = kFinishWarps * (kTopkMaxCand * 8 + kTopkMax * 8);

// host launchers (inst_topk.cu)
cudaError_t launch_topk_scan(bool split, int pass, int grid, cudaStream_t st, const CUtensorMap& tmap_hi, const CUtensorMap& tmap_lo,
                             const AffinityParams& prm);
cudaError_t launch_topk_threshold(const float* bound, float* tau, int n_pixels, int n_tiles, int tile_step, int k, cudaStream_t st);
cudaError_t launch_topk_finish(const TopkFinishParams& fp, cudaStream_t st);
cudaError_t launch_upsample_mask(const uint8_t* low, uint8_t* out, int h_lowres, int w_lowres, int H, int W, cudaStream_t st);

}  // namespace vosk
