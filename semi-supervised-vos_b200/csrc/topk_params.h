// Shapes and parameter block of the top-k extension shared by the host code (vos_prop.cu) and the kernels (affinity_topk.cuh).
#pragma once
#include "kernels.cuh"

namespace vosk {

constexpr int kTopkGroup = 2;                  // chunks per pipeline stage (32 KiB)
constexpr int kTopkMax = 64;                   // largest supported k
constexpr int kTopkK16 = 8;                    // k <= this: 16 epilogue warps, 36 slots per thread
constexpr int kTopkK8 = 24;                    // k <= this:  8 epilogue warps, 64 slots per thread; larger k: 4 warps, 112 slots
constexpr int kTopkMaxCand = 2048;             // candidates merged per target pixel by the finish kernel (lists x k)

struct TopkFinishParams {
    MergeParams mp;               // geometry + outputs shared with vos_merge_writeback
    int32_t topk;
    int32_t ref_slot[32];
    float ref_coef[32];           // log2(e) / sigma^2 ; 0 = no prior
    const uint32_t* cand_key;     // [grid * max_segs][128][kTopkMax]
    const int32_t* cand_idx;
    const int32_t* cand_cnt;      // [grid * max_segs][128]
    int32_t* out_topk_idx;        // (P, topk) int32 or null; value-descending, -1 where fewer than k references exist
};

constexpr int kFinishWarps = 4;
constexpr int kFinishSmem = kFinishWarps * (kTopkMaxCand * 8 + kTopkMax * 8);

// host launchers (inst_topk.cu)
cudaError_t launch_topk_finish(const TopkFinishParams& fp, cudaStream_t st);
cudaError_t launch_upsample_mask(const uint8_t* low, uint8_t* out, int h_lowres, int w_lowres, int H, int W, cudaStream_t st);

}  // namespace vosk
