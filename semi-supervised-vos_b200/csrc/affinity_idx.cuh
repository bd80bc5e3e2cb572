// vos_affinity_idx: the product kernel for index-label propagation (the reference's default,
// --no-probability): every reference pixel carries ONE class id, so the label gather
// (predict.py:70) is "add the weight to the accumulator of that pixel's class".
//
// Versus vos_affinity_tc (kept for dense/probability labels) this version is shaped by the first
// ncu capture (profiles/r1_affinity_v1_*.txt): the v1 tile time was the SUM of three shared-memory
// consumers -- SS-mode MMA operand reads (A+B = 128 B/clk), TMA writes, and per-column broadcast
// loads of {rowf, xf, V[..]} records in the epilogue.  Here:
//   * the target tile (A operand, hi + lo) lives in TMEM (tcgen05.st once per segment; the MMA is the
//     TS form), halving the MMA's smem reads and freeing 128 KiB of smem -> 13-stage B ring;
//   * the Gaussian prior needs no per-column data: inside a 32-column chunk the reference pixel index
//     is n_c + j, so  -coef*((dr_c + j/W)^2 + (bx + j)^2) = alpha + beta*j + gamma*j^2  with
//     per-thread alpha/beta (two variants around the single possible image-row wrap) and j, j^2
//     compile-time immediates;
//   * labels are one byte per reference pixel, fetched by lane j for column j and turned into
//     warp-uniform class bit masks with ballots; a chunk whose 32 pixels share one class (the
//     common case) takes a path with a single running sum.
// Requires W_d >= 32 (at most one row wrap per 16-column step, x bookkeeping with single subtractions); smaller maps use vos_affinity_tc.
#pragma once
#include "kernels.cuh"

namespace vosk {

constexpr int kIdxEpiWarps = 16;     // 4 per scheduler: each owns 32 TMEM lanes x 32 logit columns of a tile
constexpr int kIdxEpiThreads = kIdxEpiWarps * 32;
constexpr int kIdxThreads = 64 + kIdxEpiThreads;   // warp 0 TMA, warp 1 MMA, warps 2-17 epilogue
constexpr int kIdxSub = 4;           // partial records per (CTA, segment): one per 32-column quarter
constexpr int kIdxStages = 13;       // 13 x 16 KiB reference chunks in flight
constexpr int kIdxAccBufs = 2;       // TMEM: [0,256) two accumulators, [256,384) Q hi, [384,512) Q lo
constexpr int kIdxSmem = kIdxStages * kChunkBytes + 512 + 1024;
constexpr uint32_t kTmemQ = 256;

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

// One 16-column step of one target pixel.  cls_lane: class byte of column (lane - lane_shift) for the 16
// lanes [lane_shift, lane_shift+16).  n_valid: valid leading columns (>= 16 unless the tile is ragged).
// Arithmetic per column j:  e = s*scale2 - m ;  l += 2^e ;  pw = 2^(e + alpha + beta*j + gamma*j^2) ;
// acc[class(j)] += pw  -- in packed fp32 pairs (FFMA2/FADD2), the two exp2 per column on the MUFU.
template <int D>
__device__ __forceinline__ void consume16_idx(RowAcc<D>& st, float (&v)[kQC], uint32_t cls_lane, int lane_shift,
                                              int n_valid, const ChunkGeom& g, float inv_w, float coef, float gamma,
                                              float scale2, float w_lowres) {
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
    float cmax = v[0];
#pragma unroll
    for (int j = 1; j < kQC; ++j) cmax = fmaxf(cmax, v[j]);
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
    const float2 g2 = make_float2(gamma, gamma);
    float aA, bA;
    quad_coeffs(g.drc, g.bx, inv_w, coef, neg_m, aA, bA);      // alpha already contains -m
    float2 l2 = make_float2(0.f, 0.f), sum2 = make_float2(0.f, 0.f);
    if (homog && g.jw >= kQC) {
        // ---- path A: one class, no row wrap.  t_j = alpha + beta*j + gamma*j^2 by forward differences on
        // column pairs: T = (t_j, t_j+1), T += dT, dT += 8*gamma  (no per-column constants to materialise)
        float2 T = make_float2(aA, aA + bA + gamma);
        float2 dT = make_float2(2.f * bA + 4.f * gamma, 2.f * bA + 8.f * gamma);
        const float2 c8 = make_float2(8.f * gamma, 8.f * gamma);
#pragma unroll
        for (int j = 0; j < kQC; j += 2) {
            const float2 v2 = make_float2(v[j], v[j + 1]);
            const float2 e2 = ffma2(v2, s2, nm2);
            const float2 u2 = ffma2(v2, s2, T);
            T = fadd2(T, dT);
            dT = fadd2(dT, c8);
            l2 = fadd2(l2, make_float2(ex2(e2.x), ex2(e2.y)));
            sum2 = fadd2(sum2, make_float2(ex2(u2.x), ex2(u2.y)));
        }
        st.l += l2.x + l2.y;
        add_to_class<D>(st, static_cast<int>(first), sum2.x + sum2.y);
        return;
    }
    float aB, bB;
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

template <int D>
__global__ void __launch_bounds__(kIdxThreads, 1)
vos_affinity_idx(const __grid_constant__ CUtensorMap tmap_hi, const __grid_constant__ CUtensorMap tmap_lo,
                 const AffinityParams prm) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* r_smem = smem;                                   // kIdxStages x 16 KiB
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kIdxStages * kChunkBytes);
    uint64_t* full = bars;                       // [kIdxStages] TMA -> MMA
    uint64_t* empty = full + kIdxStages;         // [kIdxStages] MMA -> TMA
    uint64_t* q_full = empty + kIdxStages;       // epilogue (256 threads) -> MMA : target tile is in TMEM
    uint64_t* q_empty = q_full + 1;              // MMA -> epilogue : target tile may be replaced
    uint64_t* acc_full = q_empty + 1;            // [2] MMA -> epilogue
    uint64_t* acc_empty = acc_full + kIdxAccBufs;  // [2] epilogue -> MMA
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + kIdxAccBufs);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const vosd::Decomp dec = vosd::make_decomp(prm.n_pixels, prm.n_refs, prm.num_sms);

    if (warp == 0 && lane == 0) {
        prefetch_tmap(&tmap_hi);
        prefetch_tmap(&tmap_lo);
        for (int i = 0; i < kIdxStages; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
        mbar_init(q_full, kIdxEpiThreads);
        mbar_init(q_empty, 1);
        for (int i = 0; i < kIdxAccBufs; ++i) { mbar_init(&acc_full[i], 1); mbar_init(&acc_empty[i], kIdxEpiThreads); }
        fence_mbar_init();
    }
    if (warp == 1) tmem_alloc<512>(tmem_slot);
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ================= TMA producer: 8 reference chunks (hi/lo x 4 K-chunks) per tile.
        // The whole warp walks the (uniform) control flow; elect.sync picks the issuing lane, which
        // lets ptxas keep addresses in uniform registers instead of a per-instruction waterfall.
        vosd::SegIter it(dec, blockIdx.x);
        int m_tile, n0, n1;
        uint32_t stage = 0, phase = 0;
        while (it.next(m_tile, n0, n1)) {
            for (int nt = n0; nt < n1; ++nt) {
                const int r = nt / dec.tpf;
                const int row0 = prm.ref_slot[r] * prm.p_pad + (nt - r * dec.tpf) * kTile;
                for (int c = 0; c < 2 * kNKC; ++c) {
                    mbar_wait_relaxed(&empty[stage], phase ^ 1, 64);
                    if (elect_one()) {
                        mbar_arrive_expect_tx(&full[stage], kChunkBytes);
                        tma_load_2d(r_smem + stage * kChunkBytes, (c & 1) ? &tmap_lo : &tmap_hi, (c >> 1) * kKC,
                                    row0, &full[stage]);
                    }
                    __syncwarp();
                    if (++stage == kIdxStages) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // ================= MMA issuer: D[tmem] += Q[tmem] . R[smem]^T, bf16x3 (one elected lane issues)
        constexpr uint32_t idesc = umma_idesc_bf16_f32(kTile, kTile);
        vosd::SegIter it(dec, blockIdx.x);
        int m_tile, n0, n1;
        uint32_t stage = 0, phase = 0, tile_count = 0;
        const uint32_t q_hi = tmem_base + kTmemQ, q_lo = tmem_base + kTmemQ + 128;
        const uint64_t desc0 = umma_desc_kmajor_sw128(smem_u32(r_smem));
        while (it.next(m_tile, n0, n1)) {
            mbar_wait(q_full, it.seg & 1);
            tc_fence_after_sync();
            for (int nt = n0; nt < n1; ++nt, ++tile_count) {
                const uint32_t buf = tile_count % kIdxAccBufs;
                const uint32_t aphase = (tile_count / kIdxAccBufs) & 1;
                mbar_wait_relaxed(&acc_empty[buf], aphase ^ 1, 32);
                tc_fence_after_sync();
                const uint32_t d_tmem = tmem_base + buf * kTile;
#pragma unroll
                for (int c = 0; c < 2 * kNKC; ++c) {
                    const int kc = c >> 1;
                    mbar_wait(&full[stage], phase);
                    tc_fence_after_sync();
                    if (elect_one()) {
                        // stage s starts s*16 KiB after stage 0: +1024 in the (addr >> 4) field; K-step k: +2
                        const uint64_t b_desc = desc0 + static_cast<uint64_t>(stage * (kChunkBytes >> 4));
                        if ((c & 1) == 0) {
#pragma unroll
                            for (int k = 0; k < kKC / 16; ++k)
                                umma_bf16_ts(d_tmem, q_hi + (kc * 4 + k) * 8, b_desc + 2 * k, idesc, (c | k) != 0);
#pragma unroll
                            for (int k = 0; k < kKC / 16; ++k)
                                umma_bf16_ts(d_tmem, q_lo + (kc * 4 + k) * 8, b_desc + 2 * k, idesc, 1);
                        } else {
#pragma unroll
                            for (int k = 0; k < kKC / 16; ++k)
                                umma_bf16_ts(d_tmem, q_hi + (kc * 4 + k) * 8, b_desc + 2 * k, idesc, 1);
                        }
                        umma_commit(&empty[stage]);
                        if (c == 2 * kNKC - 1) umma_commit(&acc_full[buf]);
                    }
                    __syncwarp();
                    if (++stage == kIdxStages) { stage = 0; phase ^= 1; }
                }
            }
            if (elect_one()) umma_commit(q_empty);
            __syncwarp();
        }
    } else {
        // ================= epilogue: warps 2-17; TMEM lanes [32*(warp%4), +32); logit columns [32*sub, +32)
        const int quarter = warp & 3;
        const int sub = (warp - 2) >> 2;
        const int row = quarter * 32 + lane;
        const uint32_t lane_base = static_cast<uint32_t>(quarter * 32) << 16;
        const int W = prm.w_lowres;
        vosd::SegIter it(dec, blockIdx.x);
        int m_tile, n0, n1;
        uint32_t tile_count = 0;
        while (it.next(m_tile, n0, n1)) {
            // ---- stage this segment's target tile into TMEM: column quarters 0,1 write the two halves of
            // Q hi (TMEM columns [256,384)), quarters 2,3 the two halves of Q lo ([384,512))
            if (it.seg > 0) {
                mbar_wait(q_empty, (it.seg - 1) & 1);
                tc_fence_after_sync();
            }
            {
                const __nv_bfloat16* src = ((sub & 2) ? prm.ring_lo : prm.ring_hi) +
                                           (static_cast<size_t>(prm.q_slot) * prm.p_pad + m_tile * kTile + row) * kK;
                const uint4* src4 = reinterpret_cast<const uint4*>(src) + (sub & 1) * 16;
#pragma unroll
                for (int pass = 0; pass < 2; ++pass) {
                    uint32_t regs[32];
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const uint4 q = __ldg(src4 + pass * 8 + i);
                        regs[4 * i] = q.x; regs[4 * i + 1] = q.y; regs[4 * i + 2] = q.z; regs[4 * i + 3] = q.w;
                    }
                    tmem_st_32x32b_x32(tmem_base + lane_base + kTmemQ + sub * 64 + pass * 32, regs);
                }
                tmem_st_wait();
                tc_fence_before_sync();
                mbar_arrive(q_full);
            }
            RowAcc<D> st;
            st.init();
            const int m = m_tile * kTile + row;
            const int xm = m % W;
            // (reference frame r, tile j inside it) of the segment's first tile; afterwards incremental
            int r = n0 / dec.tpf;
            int j = n0 - r * dec.tpf;
            int x_sub = (j * kTile + sub * 32) % W;       // image column of this warp's first logit column
            const int x_step = kTile % W;
            for (int nt = n0; nt < n1; ++nt, ++tile_count) {
                const uint32_t buf = tile_count % kIdxAccBufs;
                const uint32_t aphase = (tile_count / kIdxAccBufs) & 1;
                const size_t row0 = static_cast<size_t>(prm.ref_slot[r]) * prm.p_pad + j * kTile + sub * 32;
                const uint32_t cls_lane = prm.cls[row0 + lane];      // class byte of logit column `lane`
                const float coef = prm.ref_coef[r];
                const float gamma = -coef * fmaf(prm.inv_w, prm.inv_w, 1.0f);
                const int n_sub = j * kTile + sub * 32;              // pixel index (in its frame) of column 0
                const int n_valid = min(kTile, prm.n_pixels - j * kTile) - sub * 32;
                mbar_wait(&acc_full[buf], aphase);
                tc_fence_after_sync();
                const uint32_t taddr = tmem_base + lane_base + buf * kTile + sub * 32;
                float v0[kQC], v1[kQC];
                tmem_ld_32x32b_x16(taddr, v0);
                tmem_ld_32x32b_x16(taddr + kQC, v1);
                tmem_ld_wait();
                tc_fence_before_sync();
                mbar_arrive(&acc_empty[buf]);                        // this warp's columns are in registers
                int xq = x_sub;
#pragma unroll 1
                for (int q = 0; q < 2; ++q) {
                    ChunkGeom g;
                    g.drc = static_cast<float>(n_sub + q * kQC - m) * prm.inv_w;
                    g.bx = static_cast<float>(xq - xm);
                    g.jw = W - xq;
                    if (q == 0)
                        consume16_idx<D>(st, v0, cls_lane, 0, n_valid, g, prm.inv_w, coef, gamma, prm.scale2, static_cast<float>(W));
                    else
                        consume16_idx<D>(st, v1, cls_lane, kQC, n_valid - kQC, g, prm.inv_w, coef, gamma, prm.scale2, static_cast<float>(W));
                    xq += kQC;
                    if (xq >= W) xq -= W;
                }
                // next tile: 128 pixels further in the same frame, or tile 0 of the next reference frame
                if (++j == dec.tpf) {
                    j = 0;
                    ++r;
                    x_sub = (sub * 32) % W;
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
    tc_fence_before_sync();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after_sync();
        tmem_dealloc<512>(tmem_base);
    }
}

}  // namespace vosk
