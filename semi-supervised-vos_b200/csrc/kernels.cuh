// Device code of libvosprop: the fused affinity kernel (tcgen05/TMEM/TMA), its on-device fp32
// checker, and the bandwidth kernels around it (ring append, label set, merge + write-back).
//
// Reference semantics implemented (paths relative to the reference root):
//   src/model/predict.py:46-49   affinity  S[n,m] = <ref[n,:], target[:,m]>          (raw dot product)
//   src/model/predict.py:52      S *= temperature
//   src/model/predict.py:55      softmax over ALL n (all reference pixels of all frames) per target pixel m
//   src/model/predict.py:58-66   post-softmax Gaussian prior W_sigma[n_pixel, m]  (sigma per reference frame)
//   src/model/predict.py:70      prediction[c,m] = sum_n label[c,n] * S[n,m]
//   src/model/predict.py:158-175 W[i,j] = exp(-((i/W_d - j/W_d)^2 + (i%W_d - j%W_d)^2) / sigma^2), fractional row
// The (N x P) affinity and the (P x P) priors never exist in memory: logits live in TMEM for one
// 128x128 tile, the softmax is streamed (running max / sum per target pixel) and the prior is
// evaluated in closed form in the epilogue.
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <cstdint>

#include "decompose.h"
#include "ptx.cuh"

namespace vosk {

using namespace vosptx;

constexpr int kTile = 128;           // target / reference pixels per tile (UMMA M = N = 128)
constexpr int kK = 256;              // embedding width
constexpr int kKC = 64;              // K elements per smem chunk = one 128-byte swizzle row of bf16
constexpr int kNKC = kK / kKC;       // 4
constexpr int kChunkBytes = kTile * kKC * 2;   // 16 KiB
constexpr int kQBytes = 2 * kNKC * kChunkBytes;  // hi + lo target tile = 128 KiB
constexpr int kStages = 5;           // reference-chunk ring
constexpr int kMetaFloats = 16;      // per reference pixel: rowf, xf, V[0..13]
constexpr int kMetaTileBytes = kTile * kMetaFloats * 4;  // 8 KiB
constexpr int kMetaStages = 2;
constexpr int kAccBufs = 4;          // TMEM accumulator ring: 4 x 128 columns = all 512
constexpr int kTcThreads = 384;      // warp 0 TMA, 1 MMA, 2 meta, 3 idle, 4-11 epilogue
constexpr int kEpiThreads = 256;
constexpr int kSmemTc = kQBytes + kStages * kChunkBytes + kMetaStages * kMetaTileBytes + 512 + 1024;
constexpr int kMetaClasses = 14;      // label floats a meta record holds (dense / top-k / probability paths)
constexpr int kMaxClasses = 24;       // index-label path only: class bytes, no meta record needed (validation, d = 22)
constexpr int kPartFloats = (2 + kMaxClasses) * kTile;  // one partial record: m, l, acc[<= 24] x 128 rows
constexpr int kIdxSub = 4;             // vos_affinity_idx: partial records per (CTA, segment), one per 32-column quarter
constexpr int kIdxEpiWarpsHost = 16;   // epilogue warps of vos_affinity_idx (= 32 x 32 blocks per 128 x 128 tile)
constexpr int kQCap = 16;              // logit columns per epilogue step (a TMEM load of 16 columns)
constexpr float kLog2e = 1.4426950408889634f;
constexpr float kNegBig = -1.0e30f;  // finite "minus infinity" for the running max

struct AffinityParams {
    int32_t n_pixels;     // P
    int32_t p_pad;        // ring rows per slot (multiple of 128)
    int32_t w_lowres;     // W_d
    int32_t n_refs;
    int32_t q_slot;       // ring slot of the target frame
    int32_t num_sms;
    int32_t ref_slot[32];
    float ref_coef[32];   // log2(e) / sigma_r^2 ; 0 = no prior
    float scale2;         // temperature * log2(e)
    const float* meta;    // [slots * p_pad][16]
    float* partials;      // [grid * max_segs * 2][kPartFloats]
    const __nv_bfloat16* ring_hi;  // read directly by the SIMT checker and by vos_affinity_idx's TMEM staging
    const __nv_bfloat16* ring_lo;
    const uint8_t* cls;   // [slots * p_pad] class id per reference pixel (0xFF = padding), index-label mode
    float inv_w;          // 1 / W_d
    uint32_t idesc;       // tcgen05 instruction descriptor (f16 or bf16 operands, fp32 accumulate, 128x128)
    int32_t feat_fmt;     // kFmtSplit: ring_hi/ring_lo = bf16 hi + lo; kFmtF16 / kFmtBF16: ring_hi only, one pass
    // top-k mode (vos_affinity_topk): per (CTA, segment, target pixel) candidate lists
    float temperature;    // predict.py:52 -- the selection compares logit * temperature as torch computes it
    int32_t topk;
    uint32_t* cand_key;   // candidate lists (pass 2): order-preserving keys of logit * temperature, topk_slot(record, slot)
    int32_t* cand_idx;    // same shape: reference index r*P + pixel
    int32_t* cand_cnt;    // [records]
    float* topk_bound;    // pass 1: [target tile][reference tile][4 column groups][128 rows] block maxima of logit * temperature
    const float* topk_tau;  // pass 2: [P] lower bound of each row's k-th largest logit * temperature
    int32_t tile_stride;  // 0/1, or (block skipping) the stride of the permuted tile order inside a reference frame
    int32_t* skip_scratch;  // block skipping: device {sum of dead blocks, CTA ticket} of the launch in flight (null: no report)
    volatile int32_t* skip_report;  // host-mapped {tag, dead blocks} slot the last CTA of the launch fills in (vosprop_block_skip auto mode)
    int32_t skip_tag;
    int32_t tile_step;    // 0/1, or (top-k pass 1) only every tile_step-th reference tile of a target tile's row is visited
    int32_t dbg;          // development aid (vosprop_debug_flags): disables parts of the epilogue; 0 in production
    long long* dbg_clk;   // development aid: per-CTA cycle counters of the role warps' waits (null in production)
};

enum : int32_t { kFmtSplit = 0, kFmtF16 = 1, kFmtBF16 = 2 };

// feature (row, k) as fp32, whatever the ring's format (SIMT checker, top-k finish)
__device__ __forceinline__ float2 ring_feat2(const AffinityParams& prm, size_t row, int k2) {
    const size_t off = row * (kK / 2) + k2;
    if (prm.feat_fmt == kFmtF16) return __half22float2(reinterpret_cast<const __half2*>(prm.ring_hi)[off]);
    float2 h = __bfloat1622float2(reinterpret_cast<const __nv_bfloat162*>(prm.ring_hi)[off]);
    if (prm.feat_fmt == kFmtSplit) {
        const float2 l = __bfloat1622float2(reinterpret_cast<const __nv_bfloat162*>(prm.ring_lo)[off]);
        h.x += l.x;
        h.y += l.y;
    }
    return h;
}

// -------------------------------------------------------------------------------------------
// Streaming softmax + prior + label gather for one target pixel (one thread).
// -------------------------------------------------------------------------------------------
template <int D>
struct RowAcc {
    float m;       // running max of scale2 * s
    float l;       // running sum of exp2(scale2*s - m)      (softmax denominator, no prior)
    float acc[D];  // running sum of exp2(scale2*s - m) * prior * V[c]
    __device__ __forceinline__ void init() {
        m = kNegBig;
        l = 0.f;
#pragma unroll
        for (int c = 0; c < D; ++c) acc[c] = 0.f;
    }
};

// Consumes 32 logits v[0..31] of reference pixels whose meta records start at `meta_col`
// (16 floats each).  `n_valid`: number of leading valid columns (>= 32 when kPartial is false).
// kPrior = false: no spatial prior (coef == 0: probability propagation, predict.py:58) -- one exp2 per logit.
template <int D, bool kPartial, bool kPrior = true>
__device__ __forceinline__ void consume32(RowAcc<D>& st, float (&v)[32], const float4* __restrict__ meta_col,
                                          int n_valid, float scale2, float coef, float rm, float xm) {
    if (kPartial) {
#pragma unroll
        for (int j = 0; j < 32; ++j)
            if (j >= n_valid) v[j] = -INFINITY;
    }
    float cmax = v[0];
#pragma unroll
    for (int j = 1; j < 32; ++j) cmax = fmaxf(cmax, v[j]);
    const float m_new = fmaxf(st.m, cmax * scale2);  // fmaxf drops a NaN from (-inf * 0)
    if (m_new > st.m) {
        const float corr = ex2(st.m - m_new);
        st.l *= corr;
#pragma unroll
        for (int c = 0; c < D; ++c) st.acc[c] *= corr;
        st.m = m_new;
    }
    const float neg_m = -st.m;
#pragma unroll
    for (int j = 0; j < 32; ++j) {
        const float4 a = meta_col[j * 4];
        float e = fmaf(v[j], scale2, neg_m);
        if (kPartial && j >= n_valid) e = -INFINITY;
        const float p = ex2(e);
        st.l += p;
        float pw = p;
        if (kPrior) {
            const float dr = a.x - rm;
            const float dx = a.y - xm;
            const float d2 = fmaf(dx, dx, dr * dr);
            pw = ex2(fmaf(d2, -coef, e));
        }
        st.acc[0] = fmaf(pw, a.z, st.acc[0]);
        if (D > 1) st.acc[1] = fmaf(pw, a.w, st.acc[1]);
        if (D > 2) {
            const float4 b = meta_col[j * 4 + 1];
            st.acc[2] = fmaf(pw, b.x, st.acc[2]);
            if (D > 3) st.acc[3] = fmaf(pw, b.y, st.acc[3]);
            if (D > 4) st.acc[4] = fmaf(pw, b.z, st.acc[4]);
            if (D > 5) st.acc[5] = fmaf(pw, b.w, st.acc[5]);
        }
        if (D > 6) {
            const float4 b = meta_col[j * 4 + 2];
            st.acc[6] = fmaf(pw, b.x, st.acc[6]);
            if (D > 7) st.acc[7] = fmaf(pw, b.y, st.acc[7]);
            if (D > 8) st.acc[8] = fmaf(pw, b.z, st.acc[8]);
            if (D > 9) st.acc[9] = fmaf(pw, b.w, st.acc[9]);
        }
        if (D > 10) {
            const float4 b = meta_col[j * 4 + 3];
            st.acc[10] = fmaf(pw, b.x, st.acc[10]);
            if (D > 11) st.acc[11] = fmaf(pw, b.y, st.acc[11]);
            if (D > 12) st.acc[12] = fmaf(pw, b.z, st.acc[12]);
            if (D > 13) st.acc[13] = fmaf(pw, b.w, st.acc[13]);
        }
    }
}

template <int D>
__device__ __forceinline__ void store_partial(const RowAcc<D>& st, float* __restrict__ rec, int row) {
    rec[row] = st.m;
    rec[kTile + row] = st.l;
#pragma unroll
    for (int c = 0; c < D; ++c) rec[(2 + c) * kTile + row] = st.acc[c];
}

// fractional row / column of a pixel exactly as the reference builds them (predict.py:167-168)
__device__ __forceinline__ void pixel_coord(int pix, int w_lowres, float& rowf, float& xf) {
    rowf = __fdiv_rn(static_cast<float>(pix), static_cast<float>(w_lowres));
    xf = static_cast<float>(pix % w_lowres);
}

// -------------------------------------------------------------------------------------------
// Product kernel: persistent, warp-specialised, TMA -> smem -> tcgen05.mma -> TMEM -> epilogue.
//   bf16x3: features are stored as hi + lo bf16 pairs; S = Qhi.Rhi + Qlo.Rhi + Qhi.Rlo with fp32
//   accumulation in TMEM (single-pass bf16 misses the 1e-3 bar by 45x).
// -------------------------------------------------------------------------------------------
template <int D>
__global__ void __launch_bounds__(kTcThreads, 1)
vos_affinity_tc(const __grid_constant__ CUtensorMap tmap_hi, const __grid_constant__ CUtensorMap tmap_lo,
                const AffinityParams prm) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* q_smem = smem;                                  // [hi kc0..3][lo kc0..3] x 16 KiB
    uint8_t* r_smem = smem + kQBytes;                        // kStages x 16 KiB
    float4* meta_smem = reinterpret_cast<float4*>(r_smem + kStages * kChunkBytes);  // kMetaStages x 8 KiB
    uint64_t* bars = reinterpret_cast<uint64_t*>(reinterpret_cast<uint8_t*>(meta_smem) + kMetaStages * kMetaTileBytes);
    uint64_t* full = bars;                     // [kStages]   TMA  -> MMA
    uint64_t* empty = full + kStages;          // [kStages]   MMA  -> TMA
    uint64_t* q_full = empty + kStages;        // [1]
    uint64_t* q_empty = q_full + 1;            // [1]
    uint64_t* acc_full = q_empty + 1;          // [kAccBufs]  MMA  -> epilogue
    uint64_t* acc_empty = acc_full + kAccBufs; // [kAccBufs]  epilogue -> MMA
    uint64_t* meta_full = acc_empty + kAccBufs;   // [kMetaStages]
    uint64_t* meta_empty = meta_full + kMetaStages;  // [kMetaStages]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(meta_empty + kMetaStages);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const vosd::Decomp dec = vosd::make_decomp(prm.n_pixels, prm.n_refs, prm.num_sms);
    const bool split = prm.feat_fmt == kFmtSplit;          // bf16 hi+lo (3 passes) or one exact 16-bit pass
    const int n_chunks = split ? 2 * kNKC : kNKC;

    if (warp == 0 && lane == 0) {
        prefetch_tmap(&tmap_hi);
        prefetch_tmap(&tmap_lo);
        for (int i = 0; i < kStages; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
        mbar_init(q_full, 1);
        mbar_init(q_empty, 1);
        for (int i = 0; i < kAccBufs; ++i) { mbar_init(&acc_full[i], 1); mbar_init(&acc_empty[i], kEpiThreads / 32); }
        for (int i = 0; i < kMetaStages; ++i) { mbar_init(&meta_full[i], 1); mbar_init(&meta_empty[i], kEpiThreads / 32); }
        fence_mbar_init();
    }
    if (warp == 1) tmem_alloc<512>(tmem_slot);
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ================= TMA producer: target tile per segment, 8 (split) or 4 (single pass) reference chunks per tile.
        // Whole warp in the (uniform) control flow, elect.sync picks the issuing lane.
        vosd::SegIter it(dec, blockIdx.x);
        int m_tile, n0, n1;
        uint32_t stage = 0, phase = 0;
        while (it.next(m_tile, n0, n1)) {
            if (it.seg > 0) mbar_wait(q_empty, (it.seg - 1) & 1);
            const int q_row = prm.q_slot * prm.p_pad + m_tile * kTile;
            if (elect_one()) {
                mbar_arrive_expect_tx(q_full, split ? kQBytes : kQBytes / 2);
                for (int kc = 0; kc < kNKC; ++kc) {
                    tma_load_2d(q_smem + kc * kChunkBytes, &tmap_hi, kc * kKC, q_row, q_full);
                    if (split) tma_load_2d(q_smem + (kNKC + kc) * kChunkBytes, &tmap_lo, kc * kKC, q_row, q_full);
                }
            }
            __syncwarp();
            for (int nt = n0; nt < n1; ++nt) {
                const int r = nt / dec.tpf;
                const int row0 = prm.ref_slot[r] * prm.p_pad + (nt - r * dec.tpf) * kTile;
                for (int c = 0; c < n_chunks; ++c) {
                    mbar_wait_relaxed(&empty[stage], phase ^ 1, 64);
                    if (elect_one()) {
                        mbar_arrive_expect_tx(&full[stage], kChunkBytes);
                        tma_load_2d(r_smem + stage * kChunkBytes, (split && (c & 1)) ? &tmap_lo : &tmap_hi,
                                    (split ? (c >> 1) : c) * kKC, row0, &full[stage]);
                    }
                    __syncwarp();
                    if (++stage == kStages) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // ================= MMA issuer (one elected lane issues; SS form: A and B from shared memory)
        const uint32_t idesc = prm.idesc;
        vosd::SegIter it(dec, blockIdx.x);
        int m_tile, n0, n1;
        uint32_t stage = 0, phase = 0, tile_count = 0;
        const uint64_t q_desc0 = umma_desc_kmajor_sw128(smem_u32(q_smem));
        const uint64_t r_desc0 = umma_desc_kmajor_sw128(smem_u32(r_smem));
        while (it.next(m_tile, n0, n1)) {
            mbar_wait(q_full, it.seg & 1);
            tc_fence_after_sync();
            for (int nt = n0; nt < n1; ++nt, ++tile_count) {
                const uint32_t buf = tile_count % kAccBufs;
                const uint32_t aphase = (tile_count / kAccBufs) & 1;
                mbar_wait_relaxed(&acc_empty[buf], aphase ^ 1, 32);
                tc_fence_after_sync();
                const uint32_t d_tmem = tmem_base + buf * kTile;
                for (int c = 0; c < n_chunks; ++c) {
                    const int kc = split ? (c >> 1) : c;
                    const bool lo_chunk = split && (c & 1);
                    mbar_wait(&full[stage], phase);
                    tc_fence_after_sync();
                    if (elect_one()) {
                        // descriptors advance in the (addr >> 4) field: 16 KiB chunk = +1024, K-step = +2
                        const uint64_t b_desc = r_desc0 + static_cast<uint64_t>(stage * (kChunkBytes >> 4));
                        const uint64_t a_hi = q_desc0 + static_cast<uint64_t>(kc * (kChunkBytes >> 4));
                        const uint64_t a_lo = q_desc0 + static_cast<uint64_t>((kNKC + kc) * (kChunkBytes >> 4));
                        if (!lo_chunk) {      // reference hi chunk: Qhi.Rhi (+ Qlo.Rhi)
#pragma unroll
                            for (int k = 0; k < kKC / 16; ++k) umma_bf16_ss(d_tmem, a_hi + 2 * k, b_desc + 2 * k, idesc, (c | k) != 0);
                            if (split) {
#pragma unroll
                                for (int k = 0; k < kKC / 16; ++k) umma_bf16_ss(d_tmem, a_lo + 2 * k, b_desc + 2 * k, idesc, 1);
                            }
                        } else {              // reference lo chunk: Qhi.Rlo
#pragma unroll
                            for (int k = 0; k < kKC / 16; ++k) umma_bf16_ss(d_tmem, a_hi + 2 * k, b_desc + 2 * k, idesc, 1);
                        }
                        umma_commit(&empty[stage]);                       // smem stage free once these MMAs retire
                        if (c == n_chunks - 1) umma_commit(&acc_full[buf]);  // accumulator complete -> epilogue
                    }
                    __syncwarp();
                    if (++stage == kStages) { stage = 0; phase ^= 1; }
                }
            }
            if (elect_one()) umma_commit(q_empty);                        // target tile may be overwritten
            __syncwarp();
        }
    } else if (warp == 2) {
        // ================= meta producer: per reference tile 128 x {rowf, xf, V[14]} via 1-D bulk copy
        vosd::SegIter it(dec, blockIdx.x);
        int m_tile, n0, n1;
        uint32_t count = 0;
        while (it.next(m_tile, n0, n1)) {
            for (int nt = n0; nt < n1; ++nt, ++count) {
                const uint32_t ms = count % kMetaStages;
                const uint32_t mph = (count / kMetaStages) & 1;
                const int r = nt / dec.tpf;
                const size_t row0 = static_cast<size_t>(prm.ref_slot[r]) * prm.p_pad + (nt - r * dec.tpf) * kTile;
                mbar_wait_relaxed(&meta_empty[ms], mph ^ 1, 128);
                if (elect_one()) {
                    mbar_arrive_expect_tx(&meta_full[ms], kMetaTileBytes);
                    bulk_load_1d(meta_smem + ms * (kMetaTileBytes / 16), prm.meta + row0 * kMetaFloats,
                                 kMetaTileBytes, &meta_full[ms]);
                }
                __syncwarp();
            }
        }
    } else if (warp >= 4) {
        // ================= epilogue: 8 warps; warp (w%4) owns TMEM lanes [32*(w%4), +32),
        // warps 4-7 take logit columns [0,64), warps 8-11 take [64,128).
        const int quarter = warp & 3;
        const int half = (warp - 4) >> 2;
        const int row = quarter * 32 + lane;
        vosd::SegIter it(dec, blockIdx.x);
        int m_tile, n0, n1;
        uint32_t tile_count = 0;
        while (it.next(m_tile, n0, n1)) {
            RowAcc<D> st;
            st.init();
            float rm, xm;
            pixel_coord(m_tile * kTile + row, prm.w_lowres, rm, xm);
            for (int nt = n0; nt < n1; ++nt, ++tile_count) {
                const uint32_t buf = tile_count % kAccBufs;
                const uint32_t aphase = (tile_count / kAccBufs) & 1;
                const uint32_t ms = tile_count % kMetaStages;
                const uint32_t mph = (tile_count / kMetaStages) & 1;
                const int r = nt / dec.tpf;
                const int j = nt - r * dec.tpf;
                const int n_valid = min(kTile, prm.n_pixels - j * kTile) - half * 64;  // relative to this half
                const float coef = prm.ref_coef[r];
                float v0[32], v1[32];
                mbar_wait(&acc_full[buf], aphase);
                tc_fence_after_sync();
                const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + buf * kTile + half * 64;
                tmem_ld_32x32b_x32(taddr, v0);
                tmem_ld_32x32b_x32(taddr + 32, v1);
                tmem_ld_wait();
                tc_fence_before_sync();
                __syncwarp();
                if (lane == 0) mbar_arrive(&acc_empty[buf]);  // accumulator drained into registers (one arrival per warp)
                mbar_wait(&meta_full[ms], mph);
                const float4* mcol = meta_smem + ms * (kMetaTileBytes / 16) + half * 64 * (kMetaFloats / 4);
                if (n_valid >= 64 && coef == 0.f) {
                    consume32<D, false, false>(st, v0, mcol, 32, prm.scale2, coef, rm, xm);
                    consume32<D, false, false>(st, v1, mcol + 32 * (kMetaFloats / 4), 32, prm.scale2, coef, rm, xm);
                } else if (n_valid >= 64) {
                    consume32<D, false>(st, v0, mcol, 32, prm.scale2, coef, rm, xm);
                    consume32<D, false>(st, v1, mcol + 32 * (kMetaFloats / 4), 32, prm.scale2, coef, rm, xm);
                } else {
                    consume32<D, true>(st, v0, mcol, n_valid, prm.scale2, coef, rm, xm);
                    consume32<D, true>(st, v1, mcol + 32 * (kMetaFloats / 4), n_valid - 32, prm.scale2, coef, rm, xm);
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(&meta_empty[ms]);
            }
            float* rec = prm.partials +
                         (static_cast<size_t>(blockIdx.x * dec.max_segs + it.seg) * 2 + half) * kPartFloats;
            store_partial<D>(st, rec, row);
        }
    }
    tc_fence_before_sync();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after_sync();
        tmem_dealloc<512>(tmem_base);
    }
}

// -------------------------------------------------------------------------------------------
// On-device fp32 checker: same ring, same meta, same decomposition, same epilogue math and the
// same partial format, but the logits are plain fp32 FMA dot products of (hi + lo) features on
// CUDA cores.  Tests use it to separate "tensor-core path is wrong" from "everything else is".
// -------------------------------------------------------------------------------------------
constexpr int kSimtThreads = 256;
constexpr int kQsStride = 260;  // floats; float4 reads of 8 consecutive rows hit 32 distinct banks
constexpr int kSmemSimt = (kTile * kQsStride + 2 * 32 * kK) * 4;

template <int D>
__global__ void __launch_bounds__(kSimtThreads, 1) vos_affinity_simt(const AffinityParams prm) {
    extern __shared__ float smem_f[];
    float* qs = smem_f;                      // [128][260]
    float* rs = smem_f + kTile * kQsStride;  // [2 halves][32 cols][256]
    const int tid = threadIdx.x;
    const int row = tid & 127;
    const int half = tid >> 7;
    const vosd::Decomp dec = vosd::make_decomp(prm.n_pixels, prm.n_refs, prm.num_sms);
    vosd::SegIter it(dec, blockIdx.x);
    int m_tile, n0, n1;
    while (it.next(m_tile, n0, n1)) {
        __syncthreads();
        const size_t q_row = static_cast<size_t>(prm.q_slot) * prm.p_pad + m_tile * kTile;
        for (int i = tid; i < kTile * (kK / 2); i += kSimtThreads) {
            const int rr = i / (kK / 2), k2 = i % (kK / 2);
            const float2 f = ring_feat2(prm, q_row + rr, k2);
            qs[rr * kQsStride + 2 * k2] = f.x;
            qs[rr * kQsStride + 2 * k2 + 1] = f.y;
        }
        RowAcc<D> st;
        st.init();
        float rm, xm;
        pixel_coord(m_tile * kTile + row, prm.w_lowres, rm, xm);
        for (int nt = n0; nt < n1; ++nt) {
            const int r = nt / dec.tpf;
            const int j = nt - r * dec.tpf;
            const size_t row0 = static_cast<size_t>(prm.ref_slot[r]) * prm.p_pad + j * kTile;
            const float coef = prm.ref_coef[r];
            for (int ch = 0; ch < 2; ++ch) {
                __syncthreads();
                // columns {h*64 + ch*32 + c : h in 0..1, c in 0..31}
                for (int i = tid; i < 64 * (kK / 2); i += kSimtThreads) {
                    const int cc = i / (kK / 2), k2 = i % (kK / 2);
                    const int col = (cc >> 5) * 64 + ch * 32 + (cc & 31);
                    const float2 f = ring_feat2(prm, row0 + col, k2);
                    rs[cc * kK + 2 * k2] = f.x;
                    rs[cc * kK + 2 * k2 + 1] = f.y;
                }
                __syncthreads();
                float v[32];
#pragma unroll
                for (int c = 0; c < 32; ++c) v[c] = 0.f;
                const float4* q4 = reinterpret_cast<const float4*>(qs + row * kQsStride);
                const float4* r4 = reinterpret_cast<const float4*>(rs + half * 32 * kK);
                for (int k4 = 0; k4 < kK / 4; ++k4) {
                    const float4 q = q4[k4];
#pragma unroll
                    for (int c = 0; c < 32; ++c) {
                        const float4 b = r4[c * (kK / 4) + k4];
                        v[c] = fmaf(q.x, b.x, v[c]);
                        v[c] = fmaf(q.y, b.y, v[c]);
                        v[c] = fmaf(q.z, b.z, v[c]);
                        v[c] = fmaf(q.w, b.w, v[c]);
                    }
                }
                const int col0 = half * 64 + ch * 32;
                const int n_valid = min(kTile, prm.n_pixels - j * kTile) - col0;
                const float4* mcol = reinterpret_cast<const float4*>(prm.meta + (row0 + col0) * kMetaFloats);
                if (n_valid >= 32) consume32<D, false>(st, v, mcol, 32, prm.scale2, coef, rm, xm);
                else consume32<D, true>(st, v, mcol, n_valid, prm.scale2, coef, rm, xm);
            }
        }
        float* rec = prm.partials + (static_cast<size_t>(blockIdx.x * dec.max_segs + it.seg) * 2 + half) * kPartFloats;
        store_partial<D>(st, rec, row);
    }
}

// Parameters of vos_merge_writeback (side_kernels.cuh), shared with the top-k finish kernel.
struct MergeParams {
    int32_t n_pixels, p_pad, w_lowres, h_lowres, n_refs, num_sms;
    int32_t d;                 // real class count
    int32_t H, W;              // full resolution
    int32_t q_slot;
    int32_t write_labels, probability;
    int32_t n_sub;             // partial records per (CTA, segment): 2 (half tiles) or 4 (quarter tiles)
    const float* partials;
    const int32_t* tables;     // host-computed decomposition tables (no 64-bit divisions in the kernel):
                               // [0, tpf) first CTA of each target tile, [tpf, 2 tpf) last CTA, [2 tpf, 2 tpf + grid) first target tile of each CTA
    int32_t tpf, max_segs;
    int32_t tables_fresh;      // the tables were written by the kernel launched right before this one: wait before reading them
    float* meta;
    uint8_t* cls;              // class-id ring (index-label mode of the next steps)
    float* out_prediction;     // (d, P) or null
    uint8_t* out_mask_lowres;  // (P) or null
    uint8_t* out_mask_fullres; // (H, W) or null
};

__device__ __forceinline__ int nearest_src(int dst, float scale, int in_size) {
    // ATen nearest_neighbor_compute_source_index (legacy 'nearest'): min(floor(dst * scale), in - 1)
    return min(static_cast<int>(floorf(static_cast<float>(dst) * scale)), in_size - 1);
}

}  // namespace vosk
