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

namespace vosk {

constexpr int kIdxEpiWarps = 16;     // 4 per scheduler: each owns 32 TMEM lanes x 32 logit columns of a tile
constexpr int kIdxEpiThreads = kIdxEpiWarps * 32;
constexpr int kIdxThreads = 64 + kIdxEpiThreads;   // warp 0 TMA, warp 1 MMA, warps 2-17 epilogue
constexpr int kIdxSub = 4;           // partial records per (CTA, segment): one per 32-column quarter
constexpr int kIdxStages = 13;       // 13 x 16 KiB reference chunks in flight
constexpr int kIdxMaxAccBufs = 3;
constexpr int kIdxSmem = kIdxStages * kChunkBytes + 512 + 1024;

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
    static constexpr int kAccBufs = kSplit ? 2 : 3;
    static constexpr uint32_t kTmemQ = kAccBufs * kTile;       // first TMEM column of the target tile
    static constexpr int kQChunks = kSplit ? 8 : 4;            // 32-column TMEM chunks of the target tile
};

struct IdxPipe {
    uint8_t* r_smem;
    uint64_t *full, *empty, *q_full, *q_empty, *acc_full, *acc_empty;
    uint32_t tmem_base;
};

// ------------------------------------------------------------------------------------------------
// Roles shared by vos_affinity_idx and vos_affinity_topk
// ------------------------------------------------------------------------------------------------

// TMA producer (one warp): kChunks reference chunks per tile through the kIdxStages ring.
// The whole warp walks the (uniform) control flow; elect.sync picks the issuing lane, which lets
// ptxas keep addresses in uniform registers instead of a per-instruction waterfall.
template <bool kSplit>
__device__ __forceinline__ void idx_role_producer(const IdxPipe& pp, const CUtensorMap* tmap_hi, const CUtensorMap* tmap_lo,
                                                  const AffinityParams& prm, const vosd::Decomp& dec) {
    constexpr int kChunks = IdxCfg<kSplit>::kChunks;
    vosd::SegIter it(dec, blockIdx.x);
    int m_tile, n0, n1;
    uint32_t stage = 0, phase = 0;
    while (it.next(m_tile, n0, n1)) {
        for (int nt = n0; nt < n1; ++nt) {
            const int r = nt / dec.tpf;
            const int row0 = prm.ref_slot[r] * prm.p_pad + (nt - r * dec.tpf) * kTile;
            for (int c = 0; c < kChunks; ++c) {
                mbar_wait_relaxed(&pp.empty[stage], phase ^ 1, 64);
                if (elect_one()) {
                    mbar_arrive_expect_tx(&pp.full[stage], kChunkBytes);
                    if (kSplit)
                        tma_load_2d(pp.r_smem + stage * kChunkBytes, (c & 1) ? tmap_lo : tmap_hi, (c >> 1) * kKC, row0,
                                    &pp.full[stage]);
                    else
                        tma_load_2d(pp.r_smem + stage * kChunkBytes, tmap_hi, c * kKC, row0, &pp.full[stage]);
                }
                __syncwarp();
                if (++stage == kIdxStages) { stage = 0; phase ^= 1; }
            }
        }
    }
}

// MMA issuer (one warp, one elected lane issues): D[tmem] += Q[tmem] . R[smem]^T
template <bool kSplit>
__device__ __forceinline__ void idx_role_mma(const IdxPipe& pp, const AffinityParams& prm, const vosd::Decomp& dec) {
    using Cfg = IdxCfg<kSplit>;
    const uint32_t idesc = prm.idesc;
    vosd::SegIter it(dec, blockIdx.x);
    int m_tile, n0, n1;
    uint32_t stage = 0, phase = 0, buf = 0, aphase = 0;
    const uint32_t q_hi = pp.tmem_base + Cfg::kTmemQ, q_lo = pp.tmem_base + Cfg::kTmemQ + 128;
    const uint64_t desc0 = umma_desc_kmajor_sw128(smem_u32(pp.r_smem));
    while (it.next(m_tile, n0, n1)) {
        mbar_wait(pp.q_full, it.seg & 1);
        tc_fence_after_sync();
        for (int nt = n0; nt < n1; ++nt) {
            mbar_wait_relaxed(&pp.acc_empty[buf], aphase ^ 1, 32);
            tc_fence_after_sync();
            const uint32_t d_tmem = pp.tmem_base + buf * kTile;
#pragma unroll
            for (int c = 0; c < Cfg::kChunks; ++c) {
                mbar_wait(&pp.full[stage], phase);
                tc_fence_after_sync();
                if (elect_one()) {
                    // stage s starts s*16 KiB after stage 0: +1024 in the (addr >> 4) field; K-step k: +2
                    const uint64_t b_desc = desc0 + static_cast<uint64_t>(stage * (kChunkBytes >> 4));
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
                    } else {
#pragma unroll
                        for (int k = 0; k < kKC / 16; ++k)
                            umma_bf16_ts(d_tmem, q_hi + (c * 4 + k) * 8, b_desc + 2 * k, idesc, (c | k) != 0);
                    }
                    umma_commit(&pp.empty[stage]);
                    if (c == Cfg::kChunks - 1) umma_commit(&pp.acc_full[buf]);
                }
                __syncwarp();
                if (++stage == kIdxStages) { stage = 0; phase ^= 1; }
            }
            if (++buf == Cfg::kAccBufs) { buf = 0; aphase ^= 1; }
        }
        if (elect_one()) umma_commit(pp.q_empty);
        __syncwarp();
    }
}

// Epilogue threads stage one segment's target tile into TMEM.  `n_sub` column groups (warps with the same
// TMEM lane quarter) share the kQChunks 32-column chunks of the row; chunk ch covers bytes [128*(ch%4), +128)
// of the pixel's hi (ch < 4) or lo (ch >= 4) feature row.
template <bool kSplit>
__device__ __forceinline__ void idx_stage_target(const IdxPipe& pp, const AffinityParams& prm, int seg, int m_tile,
                                                 int row, uint32_t lane_base, int sub, int n_sub) {
    using Cfg = IdxCfg<kSplit>;
    if (seg > 0) {
        mbar_wait(pp.q_empty, (seg - 1) & 1);
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
    mbar_arrive(pp.q_full);
}

__device__ __forceinline__ IdxPipe idx_setup(uint8_t* smem_raw, const CUtensorMap* tmap_hi, const CUtensorMap* tmap_lo,
                                             int n_acc_bufs, int epi_threads) {
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    IdxPipe pp;
    pp.r_smem = smem;                                   // kIdxStages x 16 KiB
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kIdxStages * kChunkBytes);
    pp.full = bars;                              // [kIdxStages] TMA -> MMA
    pp.empty = pp.full + kIdxStages;             // [kIdxStages] MMA -> TMA
    pp.q_full = pp.empty + kIdxStages;           // epilogue threads -> MMA : target tile is in TMEM
    pp.q_empty = pp.q_full + 1;                  // MMA -> epilogue : target tile may be replaced
    pp.acc_full = pp.q_empty + 1;                // [kIdxMaxAccBufs] MMA -> epilogue
    pp.acc_empty = pp.acc_full + kIdxMaxAccBufs; // [kIdxMaxAccBufs] epilogue -> MMA
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(pp.acc_empty + kIdxMaxAccBufs);
    const int warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) {
        prefetch_tmap(tmap_hi);
        prefetch_tmap(tmap_lo);
        for (int i = 0; i < kIdxStages; ++i) { mbar_init(&pp.full[i], 1); mbar_init(&pp.empty[i], 1); }
        mbar_init(pp.q_full, epi_threads);
        mbar_init(pp.q_empty, 1);
        for (int i = 0; i < n_acc_bufs; ++i) { mbar_init(&pp.acc_full[i], 1); mbar_init(&pp.acc_empty[i], epi_threads); }
        fence_mbar_init();
    }
    if (warp == 1) tmem_alloc<512>(tmem_slot);
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    pp.tmem_base = *tmem_slot;
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
struct ChunkGeom {
    float drc;   // (n_c - m) / W : row-coordinate difference of the step's first column (fractional rows)
    float bx;    // x(n_c) - x(m)  : column difference of the step's first column
    int jw;      // first column of the step that belongs to the next image row (>= kQC: none)
};

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

constexpr int kQC = 16;   // columns per epilogue step (kept small: the three unrolled paths must fit the I-cache)

__device__ __forceinline__ float max16(const float (&v)[kQC]) {
    const float a = fmax3(v[0], v[1], v[2]), b = fmax3(v[3], v[4], v[5]), c = fmax3(v[6], v[7], v[8]);
    const float d = fmax3(v[9], v[10], v[11]), e = fmax3(v[12], v[13], v[14]);
    return fmax3(fmax3(a, b, c), fmax3(d, e, v[15]), a);
}

// One 16-column step of one target pixel.  cls_lane: class byte of column (lane - lane_shift) for the 16
// lanes [lane_shift, lane_shift+16).  n_valid: valid leading columns (>= 16 unless the tile is ragged).
// Arithmetic per column j:  p = 2^(s*scale2 - m) ;  l += p ;  acc[class(j)] += p * 2^(alpha + beta*j + gamma*j^2)
// in packed fp32 pairs (FFMA2/FMUL2/FADD2).
template <int D>
__device__ __forceinline__ void consume16_idx(RowAcc<D>& st, float (&v)[kQC], uint32_t cls_lane, int lane_shift,
                                              int n_valid, const ChunkGeom& g, float inv_w, float coef, float gamma,
                                              float k8, float scale2, float w_lowres) {
    const uint32_t full = 0xffffffffu;
    const uint32_t window = 0xffffu << lane_shift;
    const bool partial = n_valid < kQC;
    const uint32_t valid = partial ? (n_valid <= 0 ? 0u : ((1u << n_valid) - 1u)) : 0xffffu;
    const uint32_t first = __shfl_sync(full, cls_lane, lane_shift);
    const bool homog = !partial && ((__ballot_sync(full, cls_lane == first) & window) == window);
    if (partial) {
#pragma unroll
        for (int j = 0; j < kQC; ++j)
            if (!((valid >> j) & 1u)) v[j] = -INFINITY;
    }
    const float cmax = max16(v);
    const float m_new = fmaxf(st.m, cmax * scale2);
    if (m_new > st.m) {
        const float corr = ex2(st.m - m_new);
        st.l *= corr;
#pragma unroll
        for (int c = 0; c < D; ++c) st.acc[c] *= corr;
        st.m = m_new;
    }
    const float neg_m = -st.m;
    const float2 s2 = make_float2(scale2, scale2);
    const float2 nm2 = make_float2(neg_m, neg_m);
    float2 l2 = make_float2(0.f, 0.f), sum2 = make_float2(0.f, 0.f);
    // prior exponent (log2) of column j: t_j = a0 + b0*j + gamma*j^2 (concave); sh = max(t_0, t_15)
    float a0, b0;
    quad_coeffs(g.drc, g.bx, inv_w, coef, 0.f, a0, b0);
    const float sh = fmaxf(a0, fmaf(15.f, b0, fmaf(225.f, gamma, a0)));
    // the vertex exceeds the end points by < 57*|gamma|: below 2^-150 every weight of the step is 0 in fp32
    // (as is exp(-d^2/sigma^2) in the reference, predict.py:173)
    const bool far = fmaf(-57.f, gamma, sh) < -150.f;
    // exponent spread inside the step < 100: the recurrence below cannot cross an underflow
    const bool chain_ok = fmaf(fabsf(b0), 15.f, -225.f * gamma) < 100.f;
    if (homog && g.jw >= kQC && (far || chain_ok)) {
        // ---- path A: one class, no row wrap.  The prior g(j) = 2^(t_j - sh) is carried on column pairs by
        //   G = (g(j), g(j+1)),  G *= Rho,  Rho *= (k8, k8),  Rho = (g(j+2)/g(j), g(j+3)/g(j+1)),  k8 = 2^(8*gamma)
        // five exp2 per step instead of one per column (relative drift <= ~2e-6 over the 7 steps); the step
        // total is scaled by 2^sh once.
        float2 G = make_float2(0.f, 0.f), Rho = make_float2(0.f, 0.f);
        float scale = 0.f;
        if (!far) {
            const float a1 = a0 - sh;
            const float r0 = fmaf(4.f, gamma, 2.f * b0);                    // log2 rho(0) = 2*beta + 4*gamma
            G = make_float2(ex2(a1), ex2(a1 + b0 + gamma));
            Rho = make_float2(ex2(r0), ex2(fmaf(4.f, gamma, r0)));
            scale = ex2(sh);
        }
        const float2 K8 = make_float2(k8, k8);
#pragma unroll
        for (int j = 0; j < kQC; j += 2) {
            const float2 e2 = ffma2(make_float2(v[j], v[j + 1]), s2, nm2);
            const float2 p2 = make_float2(ex2(e2.x), ex2(e2.y));
            l2 = fadd2(l2, p2);
            sum2 = ffma2(p2, G, sum2);
            if (j + 2 < kQC) {
                G = fmul2(G, Rho);
                Rho = fmul2(Rho, K8);
            }
        }
        st.l += l2.x + l2.y;
        add_to_class<D>(st, static_cast<int>(first), (sum2.x + sum2.y) * scale);
        return;
    }
    const float2 g2 = make_float2(gamma, gamma);
    float aA, bA, aB, bB;
    quad_coeffs(g.drc, g.bx, inv_w, coef, neg_m, aA, bA);              // alpha already contains -m
    quad_coeffs(g.drc, g.bx - w_lowres, inv_w, coef, neg_m, aB, bB);   // columns >= jw: next image row
    if (homog) {
        // ---- path B: one class, row wrap inside the step
#pragma unroll
        for (int j = 0; j < kQC; j += 2) {
            const bool w0 = j >= g.jw, w1 = j + 1 >= g.jw;
            const float2 v2 = make_float2(v[j], v[j + 1]);
            const float2 e2 = ffma2(v2, s2, nm2);
            float2 t2 = ffma2(make_float2(w0 ? bB : bA, w1 ? bB : bA), make_float2(float(j), float(j + 1)),
                              make_float2(w0 ? aB : aA, w1 ? aB : aA));
            t2 = ffma2(g2, make_float2(float(j * j), float((j + 1) * (j + 1))), t2);
            const float2 u2 = ffma2(v2, s2, t2);
            l2 = fadd2(l2, make_float2(ex2(e2.x), ex2(e2.y)));
            sum2 = fadd2(sum2, make_float2(ex2(u2.x), ex2(u2.y)));
        }
        st.l += l2.x + l2.y;
        add_to_class<D>(st, static_cast<int>(first), sum2.x + sum2.y);
        return;
    }
    // ---- path C: mixed classes and/or ragged tile.  Classes >= 1 get predicated adds; class 0 receives
    // the remainder of the step total (exact up to one rounding of the total).
    uint32_t mask[D];
    float part[D];
#pragma unroll
    for (int c = 1; c < D; ++c) {
        mask[c] = (__ballot_sync(full, cls_lane == static_cast<uint32_t>(c)) >> lane_shift) & valid;
        part[c] = 0.f;
    }
#pragma unroll
    for (int j = 0; j < kQC; j += 2) {
        const bool w0 = j >= g.jw, w1 = j + 1 >= g.jw;
        const float2 v2 = make_float2(v[j], v[j + 1]);
        float2 e2 = ffma2(v2, s2, nm2);
        float2 t2 = ffma2(make_float2(w0 ? bB : bA, w1 ? bB : bA), make_float2(float(j), float(j + 1)),
                          make_float2(w0 ? aB : aA, w1 ? aB : aA));
        t2 = ffma2(g2, make_float2(float(j * j), float((j + 1) * (j + 1))), t2);
        float2 u2 = ffma2(v2, s2, t2);
        if (partial) {   // padding columns: weight exactly 0 in both sums
            if (!((valid >> j) & 1u)) { e2.x = -INFINITY; u2.x = -INFINITY; }
            if (!((valid >> (j + 1)) & 1u)) { e2.y = -INFINITY; u2.y = -INFINITY; }
        }
        l2 = fadd2(l2, make_float2(ex2(e2.x), ex2(e2.y)));
        const float pw0 = ex2(u2.x), pw1 = ex2(u2.y);
        sum2 = fadd2(sum2, make_float2(pw0, pw1));
#pragma unroll
        for (int c = 1; c < D; ++c) {
            if ((mask[c] >> j) & 1u) part[c] += pw0;
            if ((mask[c] >> (j + 1)) & 1u) part[c] += pw1;
        }
    }
    st.l += l2.x + l2.y;
    float rest = sum2.x + sum2.y;
#pragma unroll
    for (int c = 1; c < D; ++c) {
        st.acc[c] += part[c];
        rest -= part[c];
    }
    st.acc[0] += fmaxf(rest, 0.f);
}

template <int D, bool kSplit>
__global__ void __launch_bounds__(kIdxThreads, 1)
vos_affinity_idx(const __grid_constant__ CUtensorMap tmap_hi, const __grid_constant__ CUtensorMap tmap_lo,
                 const AffinityParams prm) {
    using Cfg = IdxCfg<kSplit>;
    extern __shared__ uint8_t smem_raw[];
    const IdxPipe pp = idx_setup(smem_raw, &tmap_hi, &tmap_lo, Cfg::kAccBufs, kIdxEpiThreads);
    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const vosd::Decomp dec = vosd::make_decomp(prm.n_pixels, prm.n_refs, prm.num_sms);

    if (warp == 0) {
        idx_role_producer<kSplit>(pp, &tmap_hi, &tmap_lo, prm, dec);
    } else if (warp == 1) {
        idx_role_mma<kSplit>(pp, prm, dec);
    } else {
        // ================= epilogue: warps 2-17; TMEM lanes [32*(warp%4), +32); logit columns [32*sub, +32)
        const int quarter = warp & 3;
        const int sub = (warp - 2) >> 2;
        const int row = quarter * 32 + lane;
        const uint32_t lane_base = static_cast<uint32_t>(quarter * 32) << 16;
        const int W = prm.w_lowres;
        vosd::SegIter it(dec, blockIdx.x);
        int m_tile, n0, n1;
        uint32_t buf = 0, aphase = 0;
        while (it.next(m_tile, n0, n1)) {
            idx_stage_target<kSplit>(pp, prm, it.seg, m_tile, row, lane_base, sub, kIdxSub);
            RowAcc<D> st;
            st.init();
            const int m = m_tile * kTile + row;
            const int xm = m % W;
            // (reference frame r, tile j inside it) of the segment's first tile; afterwards incremental
            int r = n0 / dec.tpf;
            int j = n0 - r * dec.tpf;
            int x_sub = (j * kTile + sub * 32) % W;       // image column of this warp's first logit column
            const int x_step = kTile % W;
            float coef = prm.ref_coef[r];
            float gamma = -coef * fmaf(prm.inv_w, prm.inv_w, 1.0f);
            float k8 = ex2(8.f * gamma);
            for (int nt = n0; nt < n1; ++nt) {
                const size_t row0 = static_cast<size_t>(prm.ref_slot[r]) * prm.p_pad + j * kTile + sub * 32;
                const uint32_t cls_lane = prm.cls[row0 + lane];      // class byte of logit column `lane`
                const int n_sub = j * kTile + sub * 32;              // pixel index (in its frame) of column 0
                const int n_valid = min(kTile, prm.n_pixels - j * kTile) - sub * 32;
                mbar_wait(&pp.acc_full[buf], aphase);
                tc_fence_after_sync();
                const uint32_t taddr = pp.tmem_base + lane_base + buf * kTile + sub * 32;
                float v0[kQC], v1[kQC];
                tmem_ld_32x32b_x16(taddr, v0);
                tmem_ld_32x32b_x16(taddr + kQC, v1);
                tmem_ld_wait();
                tc_fence_before_sync();
                mbar_arrive(&pp.acc_empty[buf]);                     // this warp's columns are in registers
                if (++buf == Cfg::kAccBufs) { buf = 0; aphase ^= 1; }
                int xq = x_sub;
#pragma unroll 1
                for (int q = 0; q < 2; ++q) {
                    ChunkGeom g;
                    g.drc = static_cast<float>(n_sub + q * kQC - m) * prm.inv_w;
                    g.bx = static_cast<float>(xq - xm);
                    g.jw = W - xq;
                    if (q == 0)
                        consume16_idx<D>(st, v0, cls_lane, 0, n_valid, g, prm.inv_w, coef, gamma, k8, prm.scale2, static_cast<float>(W));
                    else
                        consume16_idx<D>(st, v1, cls_lane, kQC, n_valid - kQC, g, prm.inv_w, coef, gamma, k8, prm.scale2, static_cast<float>(W));
                    xq += kQC;
                    if (xq >= W) xq -= W;
                }
                // next tile: 128 pixels further in the same frame, or tile 0 of the next reference frame
                if (++j == dec.tpf) {
                    j = 0;
                    ++r;
                    x_sub = (sub * 32) % W;
                    if (nt + 1 < n1) {
                        coef = prm.ref_coef[r];
                        gamma = -coef * fmaf(prm.inv_w, prm.inv_w, 1.0f);
                        k8 = ex2(8.f * gamma);
                    }
                } else {
                    x_sub += x_step;
                    if (x_sub >= W) x_sub -= W;
                }
            }
            float* rec = prm.partials +
                         (static_cast<size_t>(blockIdx.x * dec.max_segs + it.seg) * kIdxSub + sub) * kPartFloats;
            store_partial<D>(st, rec, row);
        }
    }
    idx_teardown(pp);
}

}  // namespace vosk
