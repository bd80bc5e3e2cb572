// vos_affinity_idx: the product kernel for index-label propagation (the reference's default,
// --no-probability): every reference pixel carries ONE class id, so the label gather
// (predict.py:70) is "add the weight to the accumulator of that pixel's class".
//
// Versus vos_affinity_tc (kept for dense/probability labels) this version is shaped by the ncu
// captures in profiles/ (README there lists what each one changed):
//   * the target tile (A operand) lives in TMEM (tcgen05.st once per segment; the MMA is the TS form),
//     halving the MMA's smem reads and freeing 128 KiB of smem -> 13-stage B ring;
//   * the Gaussian prior needs no per-column data: inside a 16-column step the reference pixel index
//     is n_c + j, so  -coef*((dr_c + j/W)^2 + (bx + j)^2) = alpha + beta*j + gamma*j^2 ; on the common
//     path the prior is carried by a multiplicative recurrence  g(j+2) = g(j)*rho(j), rho(j+2) = rho(j)*k
//     on packed fp32 pairs (FMUL2), so the MUFU computes ONE exp2 per logit (+4 per 16 columns);
//   * labels are one byte per reference pixel, fetched by lane j for column j and turned into
//     warp-uniform class bit masks with ballots; a step whose 16 pixels share one class (the
//     common case) takes a path with a single running sum.
// Requires W_d >= 32 (at most one row wrap per 16-column step, x bookkeeping with single subtractions);
// smaller maps use vos_affinity_tc.
#pragma once
#include "kernels.cuh"

// Profiling aids (vosprop_debug_flags / vosprop_debug_clocks) are compiled in only with -DVOS_KERNEL_DEBUG (tools/epilogue_ablation.py
// builds such a library); the product kernels carry no clock reads and no flag tests.
#ifdef VOS_KERNEL_DEBUG
#define VOS_DBG(prm, bits) (((prm).dbg & (bits)) != 0)
#define VOS_DBG_ANY(prm) ((prm).dbg != 0)
#define VOS_CLK() clock64()
#else
#define VOS_DBG(prm, bits) false
#define VOS_DBG_ANY(prm) false
#define VOS_CLK() 0ll
#endif


namespace vosk {

constexpr int kIdxEpiWarps = 16;     // 4 per scheduler: each owns 32 TMEM lanes x 32 logit columns of a tile
constexpr int kIdxEpiThreads = kIdxEpiWarps * 32;
constexpr int kIdxThreads = 64 + kIdxEpiThreads;   // warp 0 TMA, warp 1 MMA, warps 2-17 epilogue
// (A 20-warp layout with setmaxnreg -- role warpgroup down to 24..48 registers, epilogue warpgroups up to 104..112 -- was
// built in round 2: ptxas uses at most 100 registers in the epilogue even when allowed 168, so the 96 of this layout cost
// nothing, while the role warps spill below 56.  Not kept.)
constexpr int kIdxRingChunks = 12;   // 12 x 16 KiB of reference chunks in flight, grouped into stages (IdxCfg)
constexpr int kIdxMaxAccBufs = 3;
constexpr int kIdxRowMaxBytes = 4 * kTile * 4;   // block skipping: 4 segment-parity slots x 128 rows of shared running maxima
constexpr int kIdxSmem = kIdxRingChunks * kChunkBytes + 512 + kIdxRowMaxBytes + 1024;

// Precision-dependent shape of the pipeline.
//   kSplit = true : fp32 features stored as bf16 hi + lo; S = Qhi.Rhi + Qlo.Rhi + Qhi.Rlo (3 MMA passes, 8 chunks of
//                   16 KiB per reference tile).  TMEM: [0,256) two accumulators, [256,384) Q hi, [384,512) Q lo.
//   kSplit = false: features arrive as fp16 (the reference's own CUDA path runs the backbone under autocast,
//                   inference_utils.py:52-53) or bf16: the product of two 16-bit floats is exact in the fp32
//                   accumulator, so ONE pass reproduces the fp32 contraction (4 chunks per tile).
//                   TMEM: [0,384) three accumulators, [384,512) Q.
template <bool kSplit>
struct IdxCfg {
    static constexpr int kChunks = kSplit ? 2 * kNKC : kNKC;   // smem chunks per reference tile
#ifndef VOS_ACC_BUFS
#define VOS_ACC_BUFS 3               // experiment hook: -DVOS_ACC_BUFS=2 measures what one accumulator less costs (profiles/README.md)
#endif
    static constexpr int kAccBufs = kSplit ? 2 : VOS_ACC_BUFS;
    static constexpr uint32_t kTmemQ = kAccBufs * kTile;       // first TMEM column of the target tile
    static constexpr int kQChunks = kSplit ? 8 : 4;            // 32-column TMEM chunks of the target tile
    static constexpr int kGroup = kSplit ? 2 : 4;              // chunks per pipeline stage: a hi+lo pair / a whole tile
    static constexpr int kStages = kIdxRingChunks / kGroup;    // 6 x 32 KiB / 3 x 64 KiB
};

struct IdxPipe {              // 32-bit shared-window addresses, computed once per thread
    uint32_t r_smem;          // kStages stages of kGroup x 16 KiB reference chunks
    uint32_t full, empty;     // [kStages] TMA -> MMA, MMA -> TMA
    uint32_t q_full, q_empty; // epilogue threads -> MMA (target tile is in TMEM), MMA -> epilogue
    uint32_t acc_full, acc_empty;  // [kIdxMaxAccBufs] MMA -> epilogue, epilogue -> MMA
    uint32_t tmem_base;
};

// ------------------------------------------------------------------------------------------------
// Roles shared by vos_affinity_idx and vos_affinity_topk
// ------------------------------------------------------------------------------------------------

// The reference tiles stream through a ring of kStages stages of kGroup 16-KiB chunks each (one mbarrier round trip
// per stage).  The first ncu-guided versions used one chunk per stage; the profile of the single-pass kernel showed
// the MMA-issuing warp itself as the limiter: ~45 dependent instructions per barrier round trip for only 4 UMMAs
// (1650 cycles per tile with the loads switched off, against 1024 cycles of tensor work).  Grouping a whole tile
// (fp16: 4 chunks, 16 UMMAs) or a hi+lo pair (split: 2 chunks, 12 UMMAs) per stage amortises the round trip.
//
// TMA producer (one warp).  The whole warp walks the (uniform) control flow; elect.sync picks the issuing lane,
// which lets ptxas keep addresses in uniform registers instead of a per-instruction waterfall.
template <bool kSplit, int kGroup, int kStages>
__device__ __forceinline__ void idx_role_producer(const IdxPipe& pp, const CUtensorMap* tmap_hi, const CUtensorMap* tmap_lo,
                                                  const AffinityParams& prm, const vosd::Decomp& dec) {
    constexpr int kChunks = IdxCfg<kSplit>::kChunks;
    constexpr uint32_t kStageBytes = kGroup * kChunkBytes;
    static_assert(kChunks % kGroup == 0, "a stage holds whole chunks of one tile");
    vosd::SegIter it(dec, blockIdx.x);
    int m_tile, n0, n1;
    uint32_t stage = 0, phase = 0;
    long long t_wait = 0, t_begin = VOS_CLK();
    const int tstep = prm.tile_step > 1 ? prm.tile_step : 1;                     // top-k pass 1: every tstep-th tile of a row
    while (it.next(m_tile, n0, n1)) {
        for (int nt = n0 + (tstep - n0 % tstep) % tstep; nt < n1; nt += tstep) {
            const int r = nt / dec.tpf;
            int jt = nt - r * dec.tpf;
            if (prm.tile_stride > 1) jt = (jt * prm.tile_stride) % dec.tpf;      // work-balancing tile order (block skipping)
            int row0 = prm.ref_slot[r] * prm.p_pad + jt * kTile;
            if (VOS_DBG(prm, 64)) row0 = 0;              // profiling: every CTA streams the same tile
#pragma unroll
            for (int g = 0; g < kChunks / kGroup; ++g) {
                const long long t0 = VOS_CLK();
                mbar_wait_relaxed_s(pp.empty + 8 * stage, phase ^ 1, 64);
                t_wait += VOS_CLK() - t0;
                if (elect_one()) {
                    const uint32_t bar = pp.full + 8 * stage;
                    if (VOS_DBG(prm, 128)) {             // profiling: no loads at all (MMA on stale shared memory)
                        mbar_arrive_s(bar);
                    } else {
                        mbar_arrive_expect_tx_s(bar, kStageBytes);
#pragma unroll
                        for (int i = 0; i < kGroup; ++i) {
                            const int c = g * kGroup + i;
                            const uint32_t dst = pp.r_smem + stage * kStageBytes + i * kChunkBytes;
                            if (kSplit) tma_load_2d_s(dst, (c & 1) ? tmap_lo : tmap_hi, (c >> 1) * kKC, row0, bar);
                            else tma_load_2d_s(dst, tmap_hi, c * kKC, row0, bar);
                        }
                    }
                }
                __syncwarp();
                if (++stage == kStages) { stage = 0; phase ^= 1; }
            }
        }
    }
#ifdef VOS_KERNEL_DEBUG
    if (prm.dbg_clk && (threadIdx.x & 31) == 0) {
        prm.dbg_clk[blockIdx.x * 16 + 0] = clock64() - t_begin;
        prm.dbg_clk[blockIdx.x * 16 + 1] = t_wait;
    }
#else
    (void)t_wait; (void)t_begin;
#endif
}

// MMA issuer (one warp, one elected lane issues): D[tmem] += Q[tmem] . R[smem]^T
template <bool kSplit, int kGroup, int kStages>
__device__ __forceinline__ void idx_role_mma(const IdxPipe& pp, const AffinityParams& prm, const vosd::Decomp& dec) {
    using Cfg = IdxCfg<kSplit>;
    constexpr uint32_t kStageBytes = kGroup * kChunkBytes;
    const uint32_t idesc = prm.idesc;
    vosd::SegIter it(dec, blockIdx.x);
    int m_tile, n0, n1;
    uint32_t stage = 0, phase = 0, buf = 0, aphase = 0;
    const uint32_t q_hi = pp.tmem_base + Cfg::kTmemQ, q_lo = pp.tmem_base + Cfg::kTmemQ + 128;
    const uint64_t desc0 = umma_desc_kmajor_sw128(pp.r_smem);
    long long t_q = 0, t_acc = 0, t_full = 0, t_begin = VOS_CLK();
#ifdef VOS_KERNEL_DEBUG
    unsigned long long ns_begin;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(ns_begin));
#endif
    const int tstep = prm.tile_step > 1 ? prm.tile_step : 1;
    while (it.next(m_tile, n0, n1)) {
        long long t0 = VOS_CLK();
        mbar_wait_s(pp.q_full, it.seg & 1);
        t_q += VOS_CLK() - t0;
        tc_fence_after_sync();
        for (int nt = n0 + (tstep - n0 % tstep) % tstep; nt < n1; nt += tstep) {
            t0 = VOS_CLK();
            mbar_wait_relaxed_s(pp.acc_empty + 8 * buf, aphase ^ 1, 32);
            t_acc += VOS_CLK() - t0;
            tc_fence_after_sync();
            const uint32_t d_tmem = pp.tmem_base + buf * kTile;
#pragma unroll
            for (int g = 0; g < Cfg::kChunks / kGroup; ++g) {
                t0 = VOS_CLK();
                mbar_wait_s(pp.full + 8 * stage, phase);
                t_full += VOS_CLK() - t0;
                tc_fence_after_sync();
                if (elect_one()) {
                    // stage s starts s*kStageBytes after stage 0 ((addr >> 4) field); chunk i: +1024; K-step k: +2
                    const uint64_t s_desc = desc0 + static_cast<uint64_t>(stage * (kStageBytes >> 4));
#pragma unroll
                    for (int i = 0; i < kGroup; ++i) {
                        const int c = g * kGroup + i;
                        const uint64_t b_desc = s_desc + static_cast<uint64_t>(i * (kChunkBytes >> 4));
                        if (kSplit) {
                            const int kc = c >> 1;
                            if ((c & 1) == 0) {   // reference hi chunk: Qhi.Rhi + Qlo.Rhi
#pragma unroll
                                for (int k = 0; k < kKC / 16; ++k)
                                    umma_bf16_ts(d_tmem, q_hi + (kc * 4 + k) * 8, b_desc + 2 * k, idesc, (c | k) != 0);
#pragma unroll
                                for (int k = 0; k < kKC / 16; ++k)
                                    umma_bf16_ts(d_tmem, q_lo + (kc * 4 + k) * 8, b_desc + 2 * k, idesc, 1);
                            } else {              // reference lo chunk: Qhi.Rlo
#pragma unroll
                                for (int k = 0; k < kKC / 16; ++k)
                                    umma_bf16_ts(d_tmem, q_hi + (kc * 4 + k) * 8, b_desc + 2 * k, idesc, 1);
                            }
                        } else if (VOS_DBG(prm, 256)) {   // profiling: no tensor work at all
                        } else if (VOS_DBG(prm, 512)) {   // profiling: consecutive UMMAs into different accumulators
#pragma unroll
                            for (int k = 0; k < kKC / 16; ++k)
                                umma_bf16_ts(pp.tmem_base + ((k & 1) ? 128u : 0u), q_hi + (c * 4 + k) * 8, b_desc + 2 * k, idesc, (c | k) != 0);
                        } else {
#pragma unroll
                            for (int k = 0; k < kKC / 16; ++k)
                                umma_bf16_ts(d_tmem, q_hi + (c * 4 + k) * 8, b_desc + 2 * k, idesc, (c | k) != 0);
                        }
                    }
                    umma_commit_s(pp.empty + 8 * stage);
                    if (g == Cfg::kChunks / kGroup - 1) umma_commit_s(pp.acc_full + 8 * buf);
                }
                __syncwarp();
                if (++stage == kStages) { stage = 0; phase ^= 1; }
            }
            if (++buf == Cfg::kAccBufs) { buf = 0; aphase ^= 1; }
        }
        if (elect_one()) umma_commit_s(pp.q_empty);
        __syncwarp();
    }
#ifdef VOS_KERNEL_DEBUG
    if (prm.dbg_clk && (threadIdx.x & 31) == 0) {
        prm.dbg_clk[blockIdx.x * 16 + 2] = clock64() - t_begin;
        prm.dbg_clk[blockIdx.x * 16 + 3] = t_q;
        prm.dbg_clk[blockIdx.x * 16 + 4] = t_acc;
        prm.dbg_clk[blockIdx.x * 16 + 5] = t_full;
        unsigned long long ns_end;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(ns_end));
        prm.dbg_clk[blockIdx.x * 16 + 6] = static_cast<long long>(ns_end - ns_begin);
        prm.dbg_clk[blockIdx.x * 16 + 7] = static_cast<long long>(ns_begin);
        prm.dbg_clk[blockIdx.x * 16 + 8] = static_cast<long long>(ns_end);
        uint32_t smid;
        asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
        prm.dbg_clk[blockIdx.x * 16 + 9] = static_cast<long long>(smid);
    }
#else
    (void)t_q; (void)t_acc; (void)t_full; (void)t_begin;
#endif
}

// Epilogue threads stage one segment's target tile into TMEM.  `n_sub` column groups (warps with the same
// TMEM lane quarter) share the kQChunks 32-column chunks of the row; chunk ch covers bytes [128*(ch%4), +128)
// of the pixel's hi (ch < 4) or lo (ch >= 4) feature row.
template <bool kSplit>
__device__ __forceinline__ void idx_stage_target(const IdxPipe& pp, const AffinityParams& prm, int seg, int m_tile,
                                                 int row, uint32_t lane_base, int sub, int n_sub) {
    using Cfg = IdxCfg<kSplit>;
    if (seg > 0) {
        mbar_wait_s(pp.q_empty, (seg - 1) & 1);
        tc_fence_after_sync();
    }
    const size_t q_row = (static_cast<size_t>(prm.q_slot) * prm.p_pad + m_tile * kTile + row) * kK;
    const int per = Cfg::kQChunks / n_sub;
    for (int ch = sub * per; ch < (sub + 1) * per; ++ch) {
        const __nv_bfloat16* src = ((ch & 4) ? prm.ring_lo : prm.ring_hi) + q_row;
        const uint4* src4 = reinterpret_cast<const uint4*>(src) + (ch & 3) * 8;
        uint32_t regs[32];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const uint4 q = __ldg(src4 + i);
            regs[4 * i] = q.x; regs[4 * i + 1] = q.y; regs[4 * i + 2] = q.z; regs[4 * i + 3] = q.w;
        }
        tmem_st_32x32b_x32(pp.tmem_base + lane_base + Cfg::kTmemQ + ch * 32, regs);
    }
    tmem_st_wait();
    tc_fence_before_sync();
    __syncwarp();
    if ((threadIdx.x & 31) == 0) mbar_arrive_s(pp.q_full);    // one arrival per epilogue warp
}

template <int kGroup, int kStages>
__device__ __forceinline__ IdxPipe idx_setup(uint8_t* smem_raw, const CUtensorMap* tmap_hi, const CUtensorMap* tmap_lo,
                                             int n_acc_bufs, int epi_warps) {
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    IdxPipe pp;
    pp.r_smem = base;
    pp.full = base + kStages * kGroup * kChunkBytes;
    pp.empty = pp.full + 8 * kStages;
    pp.q_full = pp.empty + 8 * kStages;
    pp.q_empty = pp.q_full + 8;
    pp.acc_full = pp.q_empty + 8;
    pp.acc_empty = pp.acc_full + 8 * kIdxMaxAccBufs;
    const uint32_t tmem_slot = pp.acc_empty + 8 * kIdxMaxAccBufs;
    const int warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) {
        prefetch_tmap(tmap_hi);
        prefetch_tmap(tmap_lo);
        for (int i = 0; i < kStages; ++i) { mbar_init_s(pp.full + 8 * i, 1); mbar_init_s(pp.empty + 8 * i, 1); }
        mbar_init_s(pp.q_full, epi_warps);
        mbar_init_s(pp.q_empty, 1);
        for (int i = 0; i < n_acc_bufs; ++i) { mbar_init_s(pp.acc_full + 8 * i, 1); mbar_init_s(pp.acc_empty + 8 * i, epi_warps); }
        fence_mbar_init();
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(tmem_slot) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(pp.tmem_base) : "r"(tmem_slot) : "memory");
    return pp;
}

__device__ __forceinline__ void idx_teardown(const IdxPipe& pp) {
    tc_fence_before_sync();
    __syncthreads();
    if ((threadIdx.x >> 5) == 1) {
        tc_fence_after_sync();
        tmem_dealloc<512>(pp.tmem_base);
    }
}

// ------------------------------------------------------------------------------------------------
// Epilogue arithmetic
// ------------------------------------------------------------------------------------------------

// alpha + beta*j + gamma*j^2 = -coef*((drc + j/W)^2 + (bx + j)^2); `shift` is folded into alpha
__device__ __forceinline__ void quad_coeffs(float drc, float bx, float inv_w, float coef, float shift, float& alpha, float& beta) {
    alpha = fmaf(-coef, fmaf(bx, bx, drc * drc), shift);
    beta = -2.f * coef * fmaf(drc, inv_w, bx);
}

template <int D>
__device__ __forceinline__ void add_to_class(RowAcc<D>& st, int cls, float s) {
#pragma unroll
    for (int c = 0; c < D; ++c)
        if (c == cls) st.acc[c] += s;
}

// 16 logits load: 32 lanes x 16 consecutive fp32 columns
__device__ __forceinline__ void tmem_ld_32x32b_x16(uint32_t taddr, float (&v)[16]) {
    uint32_t* r = reinterpret_cast<uint32_t*>(v);
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr)
        : "memory");
}

constexpr int kQC = 16;   // columns per epilogue step
// (Moving 1/4 or 1/2 of the exponentials from the MUFU to the FMA pipe -- vosptx::ex2_poly2, the FlashAttention-4 trick -- was
// measured at 480p, R = 9, F16: 225 -> 234 / 242 us per launch.  The epilogue is bound by its instruction count, not by the MUFU.)

__device__ __forceinline__ float max16(const float (&v)[kQC]) {
    const float a = fmax3(v[0], v[1], v[2]), b = fmax3(v[3], v[4], v[5]), c = fmax3(v[6], v[7], v[8]);
    const float d = fmax3(v[9], v[10], v[11]), e = fmax3(v[12], v[13], v[14]);
    return fmax3(fmax3(a, b, c), fmax3(d, e, v[15]), a);
}

// Per-reference-frame constants of the prior (warp-uniform)
struct PriorConst {
    float coef;    // log2(e) / sigma^2   (0: no prior)
    float gamma;   // -coef * (1 + 1/W^2)
    float k8;      // 2^(8*gamma)
    float k2;      // 2^(2*gamma)
    float k8inv;   // 2^(-8*gamma)   Horner form (fast_tile32, !kWide): ratio step of the per-pair multipliers
    float k52;     // 2^(52*gamma)
    float k1;      // 2^gamma
    bool chain_always;   // the recurrence of step16_chain is safe for every (target, reference) pair of this frame
};
__device__ __forceinline__ PriorConst prior_const(float coef, float inv_w, float w_lowres, float h_lowres) {
    PriorConst pc;
    pc.coef = coef;
    pc.gamma = -coef * fmaf(inv_w, inv_w, 1.0f);
    pc.k8 = ex2(8.f * pc.gamma);
    pc.k2 = ex2(2.f * pc.gamma);
    pc.k8inv = ex2(-8.f * pc.gamma);
    pc.k52 = ex2(52.f * pc.gamma);
    pc.k1 = ex2(pc.gamma);
    // |beta| <= 2*coef*(|drc|/W + |bx|) <= 2*coef*(H_d/W + W + 16): exponent spread inside a 16-column step
    // (+16: the masked Horner passes of a step with an image-row wrap evaluate the parabola of either row over all 16 columns,
    //  i.e. up to 16 virtual columns beyond the row's end)
    pc.chain_always = fmaf(30.f * coef, fmaf(h_lowres, inv_w, w_lowres + 16.f), -225.f * pc.gamma) < 100.f;
    return pc;
}

// Running softmax maximum of one target pixel: fold in the 16 new logits, rescale the sums when it moves.
template <int D>
__device__ __forceinline__ void update_max(RowAcc<D>& st, const float (&v)[kQC], float scale2) {
    const float m_new = fmaxf(st.m, max16(v) * scale2);
    if (m_new > st.m) {
        const float corr = ex2(st.m - m_new);
        st.l *= corr;
#pragma unroll
        for (int c = 0; c < D; ++c) st.acc[c] *= corr;
        st.m = m_new;
    }
}

// ---- Fast step: 16 consecutive reference pixels on one image row, prior by recurrence.
// Per column j:  p = 2^(s*scale2 - m) ;  l += p ;  v[j] <- p * g(j)   (returns 2^sh: the factor still missing from v[])
//   g(j) = 2^(t_j - sh),  t_j = a0 + b0*j + gamma*j^2 = -coef*((dn + j)^2/W^2 + (bx + j)^2)   (predict.py:158-175),
// carried on column pairs by  G = (g(j), g(j+1)),  G *= Rho,  Rho *= (k8, k8),  Rho = (g(j+2)/g(j), g(j+3)/g(j+1)):
// five exp2 per step for the prior instead of one per column (relative drift <= ~2e-6 over the 7 steps).
// `ok` = this lane may use the recurrence (see chain_safe); lanes with ok == false (whole step underflows) get zeros.
template <int D>
__device__ __forceinline__ float step16_chain(RowAcc<D>& st, float (&v)[kQC], float a0, float b0, float sh, bool live,
                                              const PriorConst& pc, float scale2, float& sum) {
    const float neg_m = -st.m;
    const float2 s2 = make_float2(scale2, scale2);
    const float2 nm2 = make_float2(neg_m, neg_m);
    float2 G = make_float2(0.f, 0.f), Rho = make_float2(0.f, 0.f);
    float scale = 0.f;
    if (live) {
        const float a1 = a0 - sh;
        const float r0 = fmaf(4.f, pc.gamma, 2.f * b0);                 // log2 rho(0) = 2*beta + 4*gamma
        G = make_float2(ex2(a1), ex2(a1 + b0 + pc.gamma));
        Rho = make_float2(ex2(r0), ex2(fmaf(4.f, pc.gamma, r0)));
        scale = ex2(sh);
    }
    const float2 K8 = make_float2(pc.k8, pc.k8);
    float2 l2 = make_float2(0.f, 0.f), sum2 = make_float2(0.f, 0.f);
#pragma unroll
    for (int j = 0; j < kQC; j += 2) {
        const float2 e2 = ffma2(make_float2(v[j], v[j + 1]), s2, nm2);
        const float2 p2 = make_float2(ex2(e2.x), ex2(e2.y));
        l2 = fadd2(l2, p2);
        const float2 pw2 = fmul2(p2, G);
        sum2 = fadd2(sum2, pw2);
        v[j] = pw2.x;
        v[j + 1] = pw2.y;
        if (j + 2 < kQC) {
            G = fmul2(G, Rho);
            Rho = fmul2(Rho, K8);
        }
    }
    st.l += l2.x + l2.y;
    sum = sum2.x + sum2.y;
    return scale;
}

// ---- Generic step: row wrap inside the step (columns >= jw are on the next image row: bx - W), ragged frame tail
// (`valid` mask), or a prior too steep for the recurrence (tiny sigma).  Every exponent is evaluated directly, the
// second exp2 per column on the MUFU.  v[j] <- p * prior (complete; factor 1).
template <int D>
__device__ __forceinline__ float step16_direct(RowAcc<D>& st, float (&v)[kQC], float drc, float bx, int jw, uint32_t valid,
                                               const PriorConst& pc, float inv_w, float scale2, float w_lowres, float& sum) {
    const float neg_m = -st.m;
    const float2 s2 = make_float2(scale2, scale2);
    const float2 nm2 = make_float2(neg_m, neg_m);
    const float2 g2 = make_float2(pc.gamma, pc.gamma);
    float aA, bA, aB, bB;
    quad_coeffs(drc, bx, inv_w, pc.coef, neg_m, aA, bA);
    quad_coeffs(drc, bx - w_lowres, inv_w, pc.coef, neg_m, aB, bB);
    float2 l2 = make_float2(0.f, 0.f), sum2 = make_float2(0.f, 0.f);
#pragma unroll
    for (int j = 0; j < kQC; j += 2) {
        const bool w0 = j >= jw, w1 = j + 1 >= jw;
        const float2 v2 = make_float2(v[j], v[j + 1]);      // -inf on padding columns -> both exponentials are 0
        const float2 e2 = ffma2(v2, s2, nm2);
        float2 t2 = ffma2(make_float2(w0 ? bB : bA, w1 ? bB : bA), make_float2(float(j), float(j + 1)),
                          make_float2(w0 ? aB : aA, w1 ? aB : aA));
        t2 = ffma2(g2, make_float2(float(j * j), float((j + 1) * (j + 1))), t2);
        const float2 u2 = ffma2(v2, s2, t2);
        l2 = fadd2(l2, make_float2(ex2(e2.x), ex2(e2.y)));
        const float2 pw2 = make_float2(ex2(u2.x), ex2(u2.y));
        sum2 = fadd2(sum2, pw2);
        v[j] = pw2.x;
        v[j + 1] = pw2.y;
    }
    st.l += l2.x + l2.y;
    sum = sum2.x + sum2.y;
    return 1.f;
}

// Label gather of a class-mixed step: v[] holds the 16 weights (times `scale`), `cls_lane` the class byte of the warp's
// column `lane` (this step: lanes [lane_shift, +16)), `mk` the columns of class `c` (both warp-uniform).  One masked sum
// per class present -- no class takes "the step total minus the others": that cancellation would leave an absolute error
// of ~1e-7 x total in a class whose own weights are tiny, which the validation loss (log of the true class's probability,
// loss.py:60) would see.
template <int D>
__device__ __forceinline__ void gather_mixed(RowAcc<D>& st, const float (&v)[kQC], uint32_t cls_lane, int lane_shift,
                                             uint32_t valid, uint32_t c, uint32_t mk, float /*sum*/, float scale) {
    const uint32_t full = 0xffffffffu;
    uint32_t rem = valid;
    while (true) {
        float s = 0.f;
#pragma unroll
        for (int j = 0; j < kQC; ++j)
            if ((mk >> j) & 1u) s += v[j];
        add_to_class<D>(st, static_cast<int>(c), s * scale);
        rem &= ~mk;
        if (rem == 0u) break;
        c = __shfl_sync(full, cls_lane, lane_shift + __ffs(rem) - 1);
        mk = (__ballot_sync(full, cls_lane == c) >> lane_shift) & rem;
    }
}

// Chain initialisation of one 16-column step (see step16_chain): G, Rho and the factor 2^sh still missing from the sums.
// Three exponentials instead of five: with q = g(1)/g(0) = 2^(beta + gamma),  g(1) = g(0) q,  rho(0) = g(2)/g(0) =
// q^2 2^(2 gamma),  rho(1) = rho(0) 2^(4 gamma)  -- the MUFU is the busiest pipe of this kernel.
__device__ __forceinline__ void chain_init(float a0, float b0, const PriorConst& pc, float2& G, float2& Rho, float& scale) {
    const float sh = fmaxf(a0, fmaf(15.f, b0, fmaf(225.f, pc.gamma, a0)));
    const float g0 = ex2(a0 - sh), q = ex2(b0 + pc.gamma);
    const float r0 = q * q * pc.k2;
    G = make_float2(g0, g0 * q);
    Rho = make_float2(r0, r0 * (pc.k2 * pc.k2));
    scale = ex2(sh);
}

// ------------------------------------------------------------------------------------------------
// Horner form of the prior-weighted sum of one 16-column step (the 480p product path, !kWide).
//   sum_j p_j g(j),  g(j) = 2^(a + b j + gamma j^2)  (j = 0..15: consecutive reference pixels on one image row)
// With r(j) = g(j+2)/g(j) = 2^(2b + 4 gamma + 4 gamma j) the even and the odd columns are two nested products
//   S_k = P_k + R_k * S_{k+1},  P_k = (p_2k, p_2k+1),  R_k = (r(2k), r(2k+1)),  R_{k-1} = R_k * 2^(-8 gamma),  k = 6..0,  S_7 = P_7
//   sum = g(0) * (S_0.x + 2^(b + gamma) * S_0.y)
// i.e. ONE FFMA2 + ONE FMUL2 per column pair, against four packed operations for the forward recurrence of step16_chain
// (weight, running sum, G, Rho).  The FMA pipe carries as many cycles per tile as the tensor pipe in this kernel, so the
// instruction count is what matters.  Set-up: two MUFU per step (2^b, 2^(a/2)).  The partial sums are relative to g(2k):
// inside PriorConst::chain_always' bound (exponent spread of a step < 100) they stay far from fp32 overflow, and g(0) is
// applied as h*h with h = 2^(a/2), so a very negative `a` underflows only where g(0)*S itself is below fp32 (where
// exp(-d^2/sigma^2) of the reference, predict.py:173, is a denormal or 0 as well).  Relative error of the multipliers:
// 7 x that of one MUFU.EX2 (~1.5e-6), the same as the forward recurrence.
struct HornerStep {
    float2 R;     // (r(12), r(13)) on entry to the chain
    float qk;     // g(1)/g(0) = 2^(b + gamma)
    float h;      // 2^(a/2)
};
__device__ __forceinline__ HornerStep horner_init(float a, float b, const PriorConst& pc) {
    HornerStep hs;
    const float q = ex2(b);
    hs.h = ex2(0.5f * a);
    const float r12 = q * q * pc.k52;                  // 2^(2b + 52 gamma)
    hs.R = make_float2(r12, r12 * (pc.k2 * pc.k2));    // r(13) = r(12) * 2^(4 gamma)
    hs.qk = q * pc.k1;
    return hs;
}
__device__ __forceinline__ float horner_finish(const HornerStep& hs, float2 S) {
    return fmaf(hs.qk, S.y, S.x) * hs.h * hs.h;
}

// One 16-column step whose columns do not all share (class, image row): for each image-row segment ([0, jwh) on the
// row of column 0, [jwh, 16) on the next one: x jumps back by W) and each class present in it, a Horner pass over the
// exponentials p[] masked to those columns.  cls_lane / lane_shift as in gather_mixed; everything but p[] and the
// coefficients is warp-uniform.  One masked sum per class present -- no class takes "the step total minus the others"
// (see gather_mixed).
template <int D>
__device__ __forceinline__ void horner_general16(RowAcc<D>& st, const float (&p)[kQC], uint32_t cls_lane, int lane_shift, int jwh,
                                                 float drc, float bx, const PriorConst& pc, float inv_w, float w_lowres) {
    const uint32_t full = 0xffffffffu;
    const float2 K8i = make_float2(pc.k8inv, pc.k8inv);
    const uint32_t row0 = jwh >= kQC ? 0xffffu : ((1u << jwh) - 1u);      // columns on the image row of column 0
#pragma unroll 1
    for (int seg = 0; seg < 2; ++seg) {
        uint32_t rem = seg == 0 ? row0 : (0xffffu & ~row0);
        if (rem == 0u) continue;
        float a, b;
        quad_coeffs(drc, seg == 0 ? bx : bx - w_lowres, inv_w, pc.coef, 0.f, a, b);
        const HornerStep hs = horner_init(a, b, pc);
        while (rem != 0u) {
            const uint32_t c = __shfl_sync(full, cls_lane, lane_shift + __ffs(rem) - 1);
            const uint32_t mk = (__ballot_sync(full, cls_lane == c) >> lane_shift) & rem;
            float2 R = hs.R;
            float2 S = make_float2(((mk >> 14) & 1u) ? p[14] : 0.f, ((mk >> 15) & 1u) ? p[15] : 0.f);
#pragma unroll
            for (int k = 6; k >= 0; --k) {
                const float2 P = make_float2(((mk >> (2 * k)) & 1u) ? p[2 * k] : 0.f, ((mk >> (2 * k + 1)) & 1u) ? p[2 * k + 1] : 0.f);
                S = ffma2(R, S, P);
                if (k > 0) R = fmul2(R, K8i);
            }
            add_to_class<D>(st, static_cast<int>(c), horner_finish(hs, S));
            rem &= ~mk;
        }
    }
}

// ---- One 16-column half of a tile that is not all-simple (fast_tile32, !kWide): p[] holds the exponentials.
// The columns of a half form GROUPS of equal (image row, class).  One group: a single Horner chain.  Two groups (an
// image-row wrap inside the half -- every 128-pixel tile of a 107-pixel-wide map has one -- or a class boundary): both
// chains run side by side over all 16 columns, each fed the exponentials of its own columns (P_B = P - P_A is exact: one
// of the two is 0), each normalised to its own parabola's (virtual) column 0 -- 5 packed operations per column pair
// instead of two complete masked passes through the loops of horner_general16, which stays as the fallback for three
// or more groups.  The per-scheduler instruction count is what bounds this kernel (profiles/README.md, round 2: with
// every tile forced down the simple path the launch took 187 instead of 236 us), and before this path existed a tile
// with a wrap cost 2.3 x a simple one.
// jwh: first column of the half on the next image row (>= 16: none; never 0).  (drc, bx): geometry of column 0.
template <int D>
__device__ __forceinline__ void horner_half16(RowAcc<D>& st, const float (&p)[kQC], uint32_t cls_lane, int lane_shift, int jwh,
                                              float drc, float bx, const PriorConst& pc, float inv_w, float w_lowres,
                                              uint32_t c0, uint32_t m0) {
    // c0: class of the half's column 0; m0: the half's columns of that class (both warp-uniform, from the caller's ballots)
    const uint32_t full = 0xffffffffu;
    const float2 K8i = make_float2(pc.k8inv, pc.k8inv);
    const uint32_t row0 = jwh >= kQC ? 0xffffu : ((1u << jwh) - 1u);      // columns on the image row of column 0
    const uint32_t gA = m0 & row0;                                          // group of column 0
    const uint32_t rest = 0xffffu & ~gA;
    float aA, bA;
    quad_coeffs(drc, bx, inv_w, pc.coef, 0.f, aA, bA);
    HornerStep hA = horner_init(aA, bA, pc);
    if (rest == 0u) {                                                      // ---- one group
        float2 S = make_float2(p[14], p[15]);
#pragma unroll
        for (int k = 6; k >= 0; --k) {
            S = ffma2(hA.R, S, make_float2(p[2 * k], p[2 * k + 1]));
            if (k > 0) hA.R = fmul2(hA.R, K8i);
        }
        add_to_class<D>(st, static_cast<int>(c0), horner_finish(hA, S));
        return;
    }
    const int j1 = __ffs(rest) - 1;                                        // first column outside group A
    const uint32_t c1 = __shfl_sync(full, cls_lane, lane_shift + j1);
    const bool next_row = j1 >= jwh;
    const uint32_t gB = (__ballot_sync(full, cls_lane == c1) >> lane_shift) & (next_row ? 0xffffu & ~row0 : row0);
    if (gB != rest) {                                                      // ---- three or more groups (rare)
        horner_general16<D>(st, p, cls_lane, lane_shift, jwh, drc, bx, pc, inv_w, w_lowres);
        return;
    }
    // ---- two groups: both chains side by side
    const float2 neg1 = make_float2(-1.f, -1.f);
    float2 SA = make_float2(0.f, 0.f), SB = make_float2(0.f, 0.f);
    float tA, tB;
    if (!next_row) {
        // a class boundary on one image row: one parabola, the chains share their multipliers
#pragma unroll
        for (int k = kQC / 2 - 1; k >= 0; --k) {
            const float2 P = make_float2(p[2 * k], p[2 * k + 1]);
            const float2 PA = make_float2(((gA >> (2 * k)) & 1u) ? P.x : 0.f, ((gA >> (2 * k + 1)) & 1u) ? P.y : 0.f);
            const float2 PB = ffma2(PA, neg1, P);
            if (k == kQC / 2 - 1) {
                SA = PA;
                SB = PB;
            } else {
                SA = ffma2(hA.R, SA, PA);
                SB = ffma2(hA.R, SB, PB);
                if (k > 0) hA.R = fmul2(hA.R, K8i);
            }
        }
        tA = horner_finish(hA, SA);
        tB = horner_finish(hA, SB);
    } else {
        // an image-row wrap: group B lies on the next row (x jumps back by W) and has its own parabola
        float aB, bB;
        quad_coeffs(drc, bx - w_lowres, inv_w, pc.coef, 0.f, aB, bB);
        HornerStep hB = horner_init(aB, bB, pc);
#pragma unroll
        for (int k = kQC / 2 - 1; k >= 0; --k) {
            const float2 P = make_float2(p[2 * k], p[2 * k + 1]);
            const float2 PA = make_float2(((gA >> (2 * k)) & 1u) ? P.x : 0.f, ((gA >> (2 * k + 1)) & 1u) ? P.y : 0.f);
            const float2 PB = ffma2(PA, neg1, P);
            if (k == kQC / 2 - 1) {
                SA = PA;
                SB = PB;
            } else {
                SA = ffma2(hA.R, SA, PA);
                SB = ffma2(hB.R, SB, PB);
                if (k > 0) {
                    hA.R = fmul2(hA.R, K8i);
                    hB.R = fmul2(hB.R, K8i);
                }
            }
        }
        tA = horner_finish(hA, SA);
        tB = horner_finish(hB, SB);
    }
    if (c0 == c1) {
        add_to_class<D>(st, static_cast<int>(c0), tA + tB);
    } else {
        add_to_class<D>(st, static_cast<int>(c0), tA);
        add_to_class<D>(st, static_cast<int>(c1), tB);
    }
}

// ---- 480p product path (!kWide), phase 1: running maximum, exponentials IN PLACE (va / vb <- 2^(s*scale2 - m)), softmax
// denominator.  Returns false for a dead block: nothing to add (kSkip).
// While one warp of a scheduler sits in this MUFU-bound phase the other three run their FMA-bound Horner chains (phase 2).
// (A "lazy maximum" -- no maximum tree; the reference value m follows the tile sums, tiles whose exponentials leave the
// fp32 range are reloaded from TMEM and redone exactly -- was built and measured in round 2: 6 % fewer instructions, the
// same launch time, because the accumulator buffer can be handed back to the MMA warp only after the tile has been
// judged and the epilogue warps then wait for the tensor pipe.  Not kept; profiles/README.md.)
template <int D, bool kSkip>
__device__ __forceinline__ bool tile_exps32(RowAcc<D>& st, float (&va)[kQC], float (&vb)[kQC], float scale2, uint32_t& probe,
                                            volatile float* rm_row, bool row_real) {
    const uint32_t full = 0xffffffffu;
    {
        const float bm = fmaxf(max16(va), max16(vb)) * scale2;
        // Block skipping (kSkip instantiation, vosprop_block_skip).  If every logit of this warp's 32 x 32 block lies more
        // than 127 (log2 units) below a lower bound of the final maximum of its row, each of its exponentials is below
        // 2^-127 of the row's denominator: the whole affinity row can lose at most N x 2^-127 ~ 1e-34 of its mass that way,
        // 26 orders of magnitude under fp32 resolution (and under the 1e-14 of the validation loss), so the block is left
        // out.  The bound is the largest running maximum any of the four column-quarter warps of this row has published in
        // shared memory (a racy, monotone-enough max: any value ever written is a true running maximum, hence a valid bound).
        // Trained embeddings (|f|^2 ~ 256) make most of the affinity matrix such blocks; low-contrast synthetic clips make
        // none.  The test costs a shared load and a warp vote: it runs on the 3rd, 19th, ... block of a segment and, after a dead block, on each
        // of the next 64 blocks (`probe`: bits 16.. = blocks left in that window, low bits = block counter).
        if constexpr (kSkip) {
            probe = (probe & 0xffff0000u) | ((probe + 1u) & 0xffffu);
            if ((probe >> 16) != 0u || (probe & 15u) == 3u) {
                const float bound = fmaxf(st.m, *rm_row);
                // rows beyond the frame's last pixel (ragged last target tile) are discarded by the merge: they never veto
                const bool dead = __all_sync(full, !row_real || bm - bound < -127.f);
                const uint32_t window = dead ? 64u : max(probe >> 16, 1u) - 1u;
                probe = (probe & 0xffffu) | (window << 16);
                if (dead) return false;
            }
            if (bm > st.m && bm > *rm_row) *rm_row = bm;      // publish (bm becomes this warp's running maximum below)
        }
        const float m_new = fmaxf(st.m, bm);
        if (m_new > st.m) {
            const float corr = ex2(st.m - m_new);
            st.l *= corr;
#pragma unroll
            for (int c = 0; c < D; ++c) st.acc[c] *= corr;
            st.m = m_new;
        }
    }
    const float neg_m = -st.m;
    const float2 s2 = make_float2(scale2, scale2);
    const float2 nm2 = make_float2(neg_m, neg_m);
    float2 l2 = make_float2(0.f, 0.f);
#pragma unroll
    for (int j = 0; j < kQC; j += 2) {
        const float2 ea = ffma2(make_float2(va[j], va[j + 1]), s2, nm2);
        const float2 eb = ffma2(make_float2(vb[j], vb[j + 1]), s2, nm2);
        const float2 pa = make_float2(ex2(ea.x), ex2(ea.y)), pb = make_float2(ex2(eb.x), ex2(eb.y));
        l2 = fadd2(l2, fadd2(pa, pb));
        va[j] = pa.x; va[j + 1] = pa.y;
        vb[j] = pb.x; vb[j + 1] = pb.y;
    }
    st.l += l2.x + l2.y;
    return true;
}

// ---- phase 2: prior-weighted class sums of the 32 exponentials (Horner form, PriorConst::chain_always holds for every
// reference of the frame).
// jw = first of the 32 columns that lies on the next image row (>= 32: none).
template <int D>
__device__ __forceinline__ void tile_prior32(RowAcc<D>& st, const float (&va)[kQC], const float (&vb)[kQC], uint32_t cls_lane,
                                             uint32_t cls0, uint32_t same32, int dn, float bx, int jw, const PriorConst& pc,
                                             float inv_w, float w_lowres) {
    const uint32_t full = 0xffffffffu;
    const float drc = static_cast<float>(dn) * inv_w;
    const uint32_t cB = __shfl_sync(full, cls_lane, kQC);
    const uint32_t sameB = __ballot_sync(full, cls_lane == cB) >> kQC;
    const bool simple = jw >= 2 * kQC && (same32 & 0xffffu) == 0xffffu && sameB == 0xffffu;   // warp-uniform
    if (simple) {
        // no image-row wrap inside the 32 columns and one class per 16-column half (the common case): two independent chains
        float a0, b0;
        quad_coeffs(drc, bx, inv_w, pc.coef, 0.f, a0, b0);
        const float a1 = fmaf(16.f, b0, fmaf(256.f, pc.gamma, a0)), b1 = fmaf(32.f, pc.gamma, b0);   // the same parabola at column 16
        HornerStep ha = horner_init(a0, b0, pc), hb = horner_init(a1, b1, pc);
        const float2 K8i = make_float2(pc.k8inv, pc.k8inv);
        float2 Sa = make_float2(va[kQC - 2], va[kQC - 1]), Sb = make_float2(vb[kQC - 2], vb[kQC - 1]);
#pragma unroll
        for (int k = kQC / 2 - 2; k >= 0; --k) {
            Sa = ffma2(ha.R, Sa, make_float2(va[2 * k], va[2 * k + 1]));
            Sb = ffma2(hb.R, Sb, make_float2(vb[2 * k], vb[2 * k + 1]));
            if (k > 0) {
                ha.R = fmul2(ha.R, K8i);
                hb.R = fmul2(hb.R, K8i);
            }
        }
        const float ta = horner_finish(ha, Sa), tb = horner_finish(hb, Sb);
        if (cls0 == cB) {
            add_to_class<D>(st, static_cast<int>(cls0), ta + tb);
        } else {
            add_to_class<D>(st, static_cast<int>(cls0), ta);
            add_to_class<D>(st, static_cast<int>(cB), tb);
        }
    } else {
        // an image-row wrap or a class boundary inside the 32 columns: each half by its groups of equal (image row, class)
        horner_half16<D>(st, va, cls_lane, 0, jw, drc, bx, pc, inv_w, w_lowres, cls0, same32 & 0xffffu);
        // the second half starts 16 pixels further: on the next image row if the wrap lies in the first half
        horner_half16<D>(st, vb, cls_lane, kQC, jw > kQC ? jw - kQC : kQC, static_cast<float>(dn + kQC) * inv_w,
                         bx + static_cast<float>(kQC) - (jw <= kQC ? w_lowres : 0.f), pc, inv_w, w_lowres, cB, sameB);
    }
}

// ---- Fast tile path of the wide instantiation (kWide: 1080p / narrow sigma / 15..24 classes): forward recurrence with
// per-tile safety tests.  This warp's 32 columns are all real.  Per-step bookkeeping (validity masks, path selection) is
// decided once per tile by the caller.  jw = first of the 32 columns that lies on the next image row (>= 32: none).
template <int D>
__device__ __forceinline__ void fast_tile32_wide(RowAcc<D>& st, float (&va)[kQC], float (&vb)[kQC], uint32_t cls_lane, uint32_t cls0,
                                                 uint32_t same32, int dn, float bx, int jw, const PriorConst& pc, float inv_w,
                                                 float scale2, float w_lowres) {
    constexpr bool kWide = true;
    const uint32_t full = 0xffffffffu;
    {
        const float m_new = fmaxf(st.m, fmaxf(max16(va), max16(vb)) * scale2);
        if (m_new > st.m) {
            const float corr = ex2(st.m - m_new);
            st.l *= corr;
#pragma unroll
            for (int c = 0; c < D; ++c) st.acc[c] *= corr;
            st.m = m_new;
        }
    }
    const float drc = static_cast<float>(dn) * inv_w;
    float ta, tb, sa, sb;           // step sums (of v[]) and the factors still missing from them
    bool direct = false;            // evaluate both halves directly (recurrence unsafe somewhere in the warp)
    bool far_a = false, far_b = false;
    if (kWide && !pc.chain_always) {
        // Wide maps / narrow sigma (1080p with sigma = 8): the frame-level bound does not hold, so each tile tests its own
        // two steps.  `far`: every weight of the step underflows fp32 (exp(-d^2/sigma^2) == 0 in the reference as well);
        // unsafe: the exponent spread inside a step is too large for the recurrence and the step is not far.
        float a0, b0;
        quad_coeffs(drc, bx, inv_w, pc.coef, 0.f, a0, b0);
        const float a1 = fmaf(16.f, b0, fmaf(256.f, pc.gamma, a0)), b1 = fmaf(32.f, pc.gamma, b0);
        far_a = fmaf(-57.f, pc.gamma, fmaxf(a0, fmaf(15.f, b0, fmaf(225.f, pc.gamma, a0)))) < -150.f;
        far_b = fmaf(-57.f, pc.gamma, fmaxf(a1, fmaf(15.f, b1, fmaf(225.f, pc.gamma, a1)))) < -150.f;
        const bool ok_a = fmaf(fabsf(b0), 15.f, -225.f * pc.gamma) < 100.f, ok_b = fmaf(fabsf(b1), 15.f, -225.f * pc.gamma) < 100.f;
        direct = jw < 2 * kQC || !__all_sync(full, (ok_a || far_a) && (ok_b || far_b));
        if (!direct && __all_sync(full, far_a && far_b)) {
            // nothing of this tile reaches the numerators: only the softmax denominator grows
            const float neg_m = -st.m;
            const float2 s2 = make_float2(scale2, scale2), nm2 = make_float2(neg_m, neg_m);
            float2 l2 = make_float2(0.f, 0.f);
#pragma unroll
            for (int j = 0; j < kQC; j += 2) {
                const float2 ea = ffma2(make_float2(va[j], va[j + 1]), s2, nm2);
                const float2 eb = ffma2(make_float2(vb[j], vb[j + 1]), s2, nm2);
                l2 = fadd2(l2, fadd2(make_float2(ex2(ea.x), ex2(ea.y)), make_float2(ex2(eb.x), ex2(eb.y))));
            }
            st.l += l2.x + l2.y;
            return;
        }
    }
    if (jw >= 2 * kQC && !direct) {
        const float neg_m = -st.m;
        const float2 s2 = make_float2(scale2, scale2);
        const float2 nm2 = make_float2(neg_m, neg_m);
        const float2 K8 = make_float2(pc.k8, pc.k8);
        float a0, b0;
        quad_coeffs(drc, bx, inv_w, pc.coef, 0.f, a0, b0);
        const float a1 = fmaf(16.f, b0, fmaf(256.f, pc.gamma, a0)), b1 = fmaf(32.f, pc.gamma, b0);   // the same parabola at column 16
        float2 Ga, Ra, Gb, Rb;
        chain_init(a0, b0, pc, Ga, Ra, sa);
        chain_init(a1, b1, pc, Gb, Rb, sb);
        if (kWide && far_a) { Ga = Ra = make_float2(0.f, 0.f); sa = 0.f; }
        if (kWide && far_b) { Gb = Rb = make_float2(0.f, 0.f); sb = 0.f; }
        float2 l2 = make_float2(0.f, 0.f), suma = make_float2(0.f, 0.f), sumb = make_float2(0.f, 0.f);
#pragma unroll
        for (int j = 0; j < kQC; j += 2) {
            const float2 ea = ffma2(make_float2(va[j], va[j + 1]), s2, nm2);
            const float2 eb = ffma2(make_float2(vb[j], vb[j + 1]), s2, nm2);
            const float2 pa = make_float2(ex2(ea.x), ex2(ea.y)), pb = make_float2(ex2(eb.x), ex2(eb.y));
            l2 = fadd2(l2, fadd2(pa, pb));
            const float2 wa = fmul2(pa, Ga), wb = fmul2(pb, Gb);
            suma = fadd2(suma, wa);
            sumb = fadd2(sumb, wb);
            va[j] = wa.x; va[j + 1] = wa.y;
            vb[j] = wb.x; vb[j + 1] = wb.y;
            if (j + 2 < kQC) {
                Ga = fmul2(Ga, Ra); Ra = fmul2(Ra, K8);
                Gb = fmul2(Gb, Rb); Rb = fmul2(Rb, K8);
            }
        }
        st.l += l2.x + l2.y;
        ta = suma.x + suma.y;
        tb = sumb.x + sumb.y;
    } else {
        // one image-row wrap inside the 32 columns: the half that contains it evaluates its exponents directly,
        // the other half runs its chain with the geometry of its own row
        float a0, b0;
        if (jw < kQC || (kWide && direct)) {
            sa = step16_direct<D>(st, va, drc, bx, jw, 0xffffu, pc, inv_w, scale2, w_lowres, ta);
        } else {
            quad_coeffs(drc, bx, inv_w, pc.coef, 0.f, a0, b0);
            sa = step16_chain<D>(st, va, a0, b0, fmaxf(a0, fmaf(15.f, b0, fmaf(225.f, pc.gamma, a0))), true, pc, scale2, ta);
        }
        const float drc_b = static_cast<float>(dn + kQC) * inv_w;
        if (jw > kQC || (kWide && direct)) {
            sb = step16_direct<D>(st, vb, drc_b, bx + static_cast<float>(kQC) - (jw <= kQC ? w_lowres : 0.f), jw > kQC ? jw - kQC : 2 * kQC,
                                  0xffffu, pc, inv_w, scale2, w_lowres, tb);
        } else {
            quad_coeffs(drc_b, bx + static_cast<float>(kQC) - w_lowres, inv_w, pc.coef, 0.f, a0, b0);
            sb = step16_chain<D>(st, vb, a0, b0, fmaxf(a0, fmaf(15.f, b0, fmaf(225.f, pc.gamma, a0))), true, pc, scale2, tb);
        }
    }
    if (same32 == full) {                                   // one class in all 32 columns (the common case)
        add_to_class<D>(st, static_cast<int>(cls0), fmaf(ta, sa, tb * sb));
        return;
    }
    {
        const uint32_t mk = same32 & 0xffffu;
        if (mk == 0xffffu) add_to_class<D>(st, static_cast<int>(cls0), ta * sa);
        else gather_mixed<D>(st, va, cls_lane, 0, 0xffffu, cls0, mk, ta, sa);
    }
    {
        const uint32_t c = __shfl_sync(full, cls_lane, kQC);
        const uint32_t mk = __ballot_sync(full, cls_lane == c) >> kQC;
        if (mk == 0xffffu) add_to_class<D>(st, static_cast<int>(c), tb * sb);
        else gather_mixed<D>(st, vb, cls_lane, kQC, 0xffffu, c, mk, tb, sb);
    }
    return;
}

// kWide: the frame-level bound that makes the prior recurrence safe everywhere (PriorConst::chain_always) fails for
// some reference of this step (1080p with sigma = 8): tiles test themselves.  A separate instantiation, because the extra
// live values cost the 96-register kernel 3.6 % at 480p.
template <int D, bool kSplit, bool kWide, bool kSkip = false>
__global__ void __launch_bounds__(kIdxThreads, 1)   // 18 warps: 5 on two of the schedulers -> 16384 / (5 * 32) = 96 registers
vos_affinity_idx(const __grid_constant__ CUtensorMap tmap_hi, const __grid_constant__ CUtensorMap tmap_lo,
                 const __grid_constant__ AffinityParams prm) {
    using Cfg = IdxCfg<kSplit>;
    extern __shared__ uint8_t smem_raw[];
    pdl_launch_dependents();
    const IdxPipe pp = idx_setup<Cfg::kGroup, Cfg::kStages>(smem_raw, &tmap_hi, &tmap_lo, Cfg::kAccBufs, kIdxEpiWarps);
    pdl_wait();       // barriers, TMEM and tensor-map prefetch are set up while the kernel before us drains
    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const vosd::Decomp dec = vosd::make_decomp(prm.n_pixels, prm.n_refs, prm.num_sms);

    if (warp == 0) {
        idx_role_producer<kSplit, Cfg::kGroup, Cfg::kStages>(pp, &tmap_hi, &tmap_lo, prm, dec);
    } else if (warp == 1) {
        idx_role_mma<kSplit, Cfg::kGroup, Cfg::kStages>(pp, prm, dec);
    } else {
        // ================= epilogue: warps 2-17; TMEM lanes [32*(warp%4), +32); logit columns [32*sub, +32)
        const uint32_t full = 0xffffffffu;
        const int quarter = warp & 3;
        const int sub = (warp - 2) >> 2;
        const int row = quarter * 32 + lane;
        const uint32_t lane_base = static_cast<uint32_t>(quarter * 32) << 16;
        // values the tile loop needs every iteration, pinned to registers: ptxas otherwise recomputes them from %tid / the shared
        // window base each time (~25 of a tile's ~200 bookkeeping instructions in the round-2 profile)
        const uint32_t bar_full = pin_reg(pp.acc_full), bar_empty = pin_reg(pp.acc_empty);
        const uint32_t tbase = pin_reg(pp.tmem_base + lane_base + static_cast<uint32_t>(sub * 32));
        const uint32_t lane_is0 = pin_reg(lane == 0 ? 1u : 0u);
        const int W = prm.w_lowres;
        const float w_f = static_cast<float>(W);
        const float h_f = static_cast<float>((prm.n_pixels + W - 1) / W);
        const float inv_w = prm.inv_w, scale2 = prm.scale2;
        const int x_step = kTile % W;
        const int last_valid = prm.n_pixels - (dec.tpf - 1) * kTile - sub * 32;   // real columns of this warp in a frame's last tile
        const uint32_t ragged = pin_reg(last_valid >= 32 ? 0xffffffffu : (last_valid <= 0 ? 0u : ((1u << last_valid) - 1u)));
        // tile index (inside a frame) from which this warp's columns are ragged: the last tile, or none at all
        const int ragged_tile = pin_reg(ragged == 0xffffffffu ? dec.tpf : dec.tpf - 1);
        const uint32_t sub_lane = pin_reg(static_cast<uint32_t>(sub * 32 + lane));
        vosd::SegIter it(dec, blockIdx.x);
        int m_tile, n0, n1;
        uint32_t buf = 0, aphase = 0;
        int n_dead = 0;                    // kSkip: 32 x 32 blocks this warp left out (warp-uniform)
        while (it.next(m_tile, n0, n1)) {
            idx_stage_target<kSplit>(pp, prm, it.seg, m_tile, row, lane_base, sub, kIdxSub);
            RowAcc<D> st;
            st.init();
            uint32_t probe = 0;            // block-skipping state of fast_tile32 (warp-uniform)
            volatile float* rm_row = nullptr;
            if constexpr (kSkip) {
                // this row's slot of the shared running maxima; 4 slots by segment parity because the warps of a CTA drift
                // apart by fewer than 3 tiles (accumulator barriers), hence by fewer than 4 segments
                uint8_t* aligned = smem_raw + (((smem_u32(smem_raw) + 1023u) & ~1023u) - smem_u32(smem_raw));
                rm_row = reinterpret_cast<volatile float*>(aligned + kIdxRingChunks * kChunkBytes + 512) + (it.seg & 3) * kTile + row;
                *rm_row = kNegBig;
            }
            const int m = m_tile * kTile + row;
            const int xm = m % W;
            // (reference frame r, tile j inside it) of the segment's first tile; afterwards incremental
            int r = n0 / dec.tpf;
            int j = n0 - r * dec.tpf;
            // kSkip: the tiles of a reference frame are visited in a strided order (tile jp = j * stride mod tiles-per-frame,
            // stride coprime: a permutation), so that the live tiles around the diagonal are spread evenly over the linear
            // tile space and every CTA's range holds its share of them.  jp and everything derived from it advance
            // incrementally (one add and one conditional subtract per tile; the modulos are taken once per segment).
            int jp = j;
            int xs_step = 0, xs_wrap = 0;
            if constexpr (kSkip) {
                jp = (j * prm.tile_stride) % dec.tpf;
                xs_step = (prm.tile_stride * kTile) % W;
                xs_wrap = (dec.tpf * kTile) % W;
            }
            int x_sub = (jp * kTile + sub * 32) % W;      // image column of this warp's first logit column
            int dn = jp * kTile + sub * 32 - m;           // pixel-index difference of that column to the target pixel
            const uint8_t* cls_p = prm.cls + static_cast<size_t>(prm.ref_slot[r]) * prm.p_pad + jp * kTile + sub * 32 + lane;
            PriorConst pc = prior_const(prm.ref_coef[r], inv_w, w_f, h_f);
            uint32_t cls_next = __ldg(cls_p);                        // class byte of logit column `lane`, one tile ahead
            const bool row_real = m < prm.n_pixels;
            for (int nt = n0; nt < n1; ++nt) {
                const uint32_t cls_lane = cls_next;
                {   // prefetch the next tile's class bytes: the load's latency hides behind this tile's arithmetic
                    const bool wrap_ref = j + 1 == dec.tpf;
                    const uint8_t* nxt = wrap_ref ? prm.cls + static_cast<size_t>(prm.ref_slot[min(r + 1, prm.n_refs - 1)]) * prm.p_pad + sub_lane
                                                  : cls_p + kTile;
                    if constexpr (kSkip) {
                        if (!wrap_ref) {
                            const int jn = jp + prm.tile_stride;
                            nxt = cls_p + ((jn >= dec.tpf ? jn - dec.tpf : jn) - jp) * kTile;
                        }
                    }
                    if (nt + 1 < n1) cls_next = __ldg(nxt);
                }
                const bool all_valid = jp < ragged_tile;               // every column of this warp is a real pixel
                const uint32_t valid32 = all_valid ? full : ragged;
                mbar_wait_hint_s(bar_full + 8 * buf, aphase, 20000u);
                tc_fence_after_sync();
                const uint32_t taddr = tbase + buf * kTile;
                const uint32_t cls0 = __shfl_sync(full, cls_lane, 0);
                const uint32_t same32 = __ballot_sync(full, cls_lane == cls0);
                if (all_valid && (kWide || pc.chain_always) && !VOS_DBG_ANY(prm)) {
                    float va[kQC], vb[kQC];
                    if constexpr (kWide) {
                        tmem_ld_32x32b_x16(taddr, va);
                        tmem_ld_32x32b_x16(taddr + kQC, vb);
                        tmem_ld_wait();
                        tc_fence_before_sync();
                        __syncwarp();
                        if (lane_is0) mbar_arrive_s(bar_empty + 8 * buf);
                        fast_tile32_wide<D>(st, va, vb, cls_lane, cls0, same32, dn, static_cast<float>(x_sub - xm), W - x_sub, pc, inv_w, scale2, w_f);
                    } else {
                        tmem_ld_32x32b_x16(taddr, va);
                        tmem_ld_32x32b_x16(taddr + kQC, vb);
                        tmem_ld_wait();
                        tc_fence_before_sync();
                        __syncwarp();
                        if (lane_is0) mbar_arrive_s(bar_empty + 8 * buf);
                        if (tile_exps32<D, kSkip>(st, va, vb, scale2, probe, rm_row, row_real))
                            tile_prior32<D>(st, va, vb, cls_lane, cls0, same32, dn, static_cast<float>(x_sub - xm), W - x_sub, pc, inv_w, w_f);
                        else if constexpr (kSkip) ++n_dead;
                    }
                } else {
                int xq = x_sub;
#pragma unroll 1
                for (int q = 0; q < 2; ++q) {
                    float v[kQC];
                    if (VOS_DBG(prm, 32)) {
#pragma unroll
                        for (int i = 0; i < kQC; ++i) v[i] = 0.f;
                    } else {
                        tmem_ld_32x32b_x16(taddr + q * kQC, v);
                        tmem_ld_wait();
                    }
                    if (q == 1) {                                    // both halves of this warp's columns are in registers
                        tc_fence_before_sync();
                        __syncwarp();
                        if (lane_is0) mbar_arrive_s(bar_empty + 8 * buf);       // one arrival per warp: 512 per-thread
                    }                                                            // arrivals serialise on the barrier word
                    const int lane_shift = q * kQC;
                    const uint32_t valid = (valid32 >> lane_shift) & 0xffffu;
                    const int jw = W - xq;                            // first column of the step on the next image row
                    const float bx = static_cast<float>(xq - xm);
                    const float drc = static_cast<float>(dn + lane_shift) * inv_w;
                    xq += kQC;
                    if (xq >= W) xq -= W;
                    if (valid == 0u) continue;                       // beyond the frame's last pixel
                    if (VOS_DBG(prm, 1)) { st.l += v[0]; continue; }
                    if (valid != 0xffffu) {
#pragma unroll
                        for (int i = 0; i < kQC; ++i)
                            if (!((valid >> i) & 1u)) v[i] = -INFINITY;
                    }
                    if (!VOS_DBG(prm, 8)) update_max<D>(st, v, scale2);
                    float sum, scale;
                    bool fast = jw >= kQC && valid == 0xffffu;
                    float a0, b0, sh;
                    bool live = true;
                    if (fast) {
                        // prior exponent (log2) of column i: t_i = a0 + b0*i + gamma*i^2 (concave); sh = max(t_0, t_15)
                        quad_coeffs(drc, bx, inv_w, pc.coef, 0.f, a0, b0);
                        sh = fmaxf(a0, fmaf(15.f, b0, fmaf(225.f, pc.gamma, a0)));
                        if (!pc.chain_always) {
                            // the vertex exceeds the end points by < 57*|gamma|: below 2^-150 every weight of the step is 0
                            // in fp32 (as is exp(-d^2/sigma^2) in the reference, predict.py:173) -> dead lane;
                            // exponent spread inside the step < 100: the recurrence cannot cross an underflow
                            live = !(fmaf(-57.f, pc.gamma, sh) < -150.f);
                            const bool chain_ok = fmaf(fabsf(b0), 15.f, -225.f * pc.gamma) < 100.f;
                            fast = __all_sync(full, !live || chain_ok);
                        }
                    }
                    if (VOS_DBG(prm, 4)) { a0 = 0.f; b0 = 0.f; sh = 0.f; live = true; fast = true; }
                    if (VOS_DBG(prm, 16)) {
                        float l = 0.f;
#pragma unroll
                        for (int i = 0; i < kQC; ++i) l += ex2(fmaf(v[i], scale2, -st.m));
                        st.l += l; sum = l; scale = 1.f;
                    } else
                    if (fast) scale = step16_chain<D>(st, v, a0, b0, sh, live, pc, scale2, sum);
                    else scale = step16_direct<D>(st, v, drc, bx, jw, valid, pc, inv_w, scale2, w_f, sum);
                    // ---- label gather: add each column's weight to its class (class bytes are warp-uniform per column)
                    if (VOS_DBG(prm, 2)) { st.acc[0] += sum * scale; continue; }
                    uint32_t c = cls0, mk = same32 & 0xffffu;
                    if (q == 1 || valid != 0xffffu) {
                        c = __shfl_sync(full, cls_lane, lane_shift + __ffs(valid) - 1);
                        mk = (__ballot_sync(full, cls_lane == c) >> lane_shift) & valid;
                    }
                    if (mk == valid) add_to_class<D>(st, static_cast<int>(c), sum * scale);   // one class (common)
                    else gather_mixed<D>(st, v, cls_lane, lane_shift, valid, c, mk, sum, scale);
                }
                }
                if (++buf == Cfg::kAccBufs) { buf = 0; aphase ^= 1; }
                // next tile: 128 pixels further in the same frame, or tile 0 of the next reference frame
                if (++j == dec.tpf) {
                    j = 0;
                    jp = 0;
                    ++r;
                    x_sub = (sub * 32) % W;
                    dn = sub * 32 - m;
                    if (nt + 1 < n1) {
                        pc = prior_const(prm.ref_coef[r], inv_w, w_f, h_f);
                        cls_p = prm.cls + static_cast<size_t>(prm.ref_slot[r]) * prm.p_pad + sub_lane;
                    }
                } else if constexpr (kSkip) {
                    int step = prm.tile_stride;               // tiles forward; minus a whole frame when the index wraps
                    x_sub += xs_step;
                    if (jp + step >= dec.tpf) {
                        step -= dec.tpf;
                        x_sub -= xs_wrap;
                    }
                    jp += step;
                    x_sub = x_sub >= W ? x_sub - W : (x_sub < 0 ? x_sub + W : x_sub);
                    dn += step * kTile;
                    cls_p += step * kTile;
                } else {
                    jp = j;
                    x_sub += x_step;
                    if (x_sub >= W) x_sub -= W;
                    dn += kTile;
                    cls_p += kTile;
                }
            }
            float* rec = prm.partials +
                         (static_cast<size_t>(blockIdx.x * dec.max_segs + it.seg) * kIdxSub + sub) * kPartFloats;
            store_partial<D>(st, rec, row);
        }
        if constexpr (kSkip) {
            // report for the host's auto mode (vosprop_block_skip): how many blocks the launch skipped.  Warps add to the
            // device sum; the last CTA to arrive hands the total to the host-mapped slot (the host reads it some launches
            // later, without ever synchronising) and re-arms the scratch words for the next launch.
            if (prm.skip_scratch != nullptr && lane == 0 && n_dead != 0) atomicAdd(prm.skip_scratch, n_dead);
        }
    }
    idx_teardown(pp);
    if constexpr (kSkip) {
        if (prm.skip_scratch != nullptr && threadIdx.x == 0) {
            __threadfence();
            if (atomicAdd(prm.skip_scratch + 1, 1) == static_cast<int>(gridDim.x) - 1) {
                const int total = atomicExch(prm.skip_scratch, 0);
                prm.skip_scratch[1] = 0;
                prm.skip_report[1] = total;
                __threadfence_system();
                prm.skip_report[0] = prm.skip_tag;
                __threadfence_system();
            }
        }
    }
}

}  // namespace vosk
