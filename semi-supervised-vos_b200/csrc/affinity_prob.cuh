// vos_affinity_prob: probability propagation (--probability-propagation, inference_utils.py:67-68: the raw (d, P) prediction
// of a frame is stored as its label) on the pipeline of vos_affinity_idx.
//
// The reference applies no spatial prior in this mode (predict.py:58), so per logit the epilogue needs one exponential and
// D multiply-adds with the reference pixel's label probabilities:  acc[c] += 2^(s*scale2 - m) * V[c][n]   (predict.py:70).
// The labels are dense fp32 records (64 bytes per reference pixel: {rowf, xf, V[14]}, the `meta` ring).  Versus
// vos_affinity_tc, the round-1 kernel that served this mode (SS-form MMA, 8 epilogue warps at 168 registers, 273 us per 480p
// launch), this one has the target tile in TMEM (TS-form MMA), the 12-chunk reference ring and 16 epilogue warps of
// vos_affinity_idx, plus one warp that streams each tile's 128 records (8 KiB, cp.async.bulk) through a two-stage shared
// buffer; an epilogue thread reads the records of its 32 columns as warp-uniform (broadcast) 8-byte loads.
// References WITH a prior and dense labels (float label histories handed to the stateless predict() adapter), and maps of
// any width, stay on vos_affinity_tc.
#pragma once
#include "affinity_idx.cuh"

namespace vosk {

constexpr int kProbThreads = 128 + kIdxEpiThreads;     // warp 0 TMA, 1 MMA, 2 label records, 3 idle; warps 4-19 epilogue
constexpr int kProbLabelStages = 2;
constexpr int kProbSmem = kIdxRingChunks * kChunkBytes + 512 + kProbLabelStages * kMetaTileBytes + 1024;

template <int D, bool kSplit>
__global__ void __launch_bounds__(kProbThreads, 1)
vos_affinity_prob(const __grid_constant__ CUtensorMap tmap_hi, const __grid_constant__ CUtensorMap tmap_lo,
                  const __grid_constant__ AffinityParams prm) {
    using Cfg = IdxCfg<kSplit>;
    extern __shared__ uint8_t smem_raw[];
    pdl_launch_dependents();
    const IdxPipe pp = idx_setup<Cfg::kGroup, Cfg::kStages>(smem_raw, &tmap_hi, &tmap_lo, Cfg::kAccBufs, kIdxEpiWarps);
    // label stages behind the 512 bytes of pipeline barriers; their own barriers in the second half of those 512 bytes
    const uint32_t lab_smem = pp.full + 512;
    const uint32_t lab_full = pp.full + 256, lab_empty = pp.full + 256 + 8 * kProbLabelStages;
    if (threadIdx.x == 32) {
        for (int i = 0; i < kProbLabelStages; ++i) { mbar_init_s(lab_full + 8 * i, 1); mbar_init_s(lab_empty + 8 * i, kIdxEpiWarps); }
        fence_mbar_init();
    }
    __syncthreads();
    pdl_wait();
    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const vosd::Decomp dec = vosd::make_decomp(prm.n_pixels, prm.n_refs, prm.num_sms);

    if (warp == 0) {
        idx_role_producer<kSplit, Cfg::kGroup, Cfg::kStages>(pp, &tmap_hi, &tmap_lo, prm, dec);
    } else if (warp == 1) {
        idx_role_mma<kSplit, Cfg::kGroup, Cfg::kStages>(pp, prm, dec);
    } else if (warp == 2) {
        // ================= label records: 128 x 64 bytes per reference tile, one bulk copy
        vosd::SegIter it(dec, blockIdx.x);
        int m_tile, n0, n1;
        uint32_t count = 0;
        while (it.next(m_tile, n0, n1)) {
            for (int nt = n0; nt < n1; ++nt, ++count) {
                const uint32_t ms = count % kProbLabelStages, mph = (count / kProbLabelStages) & 1;
                const int r = nt / dec.tpf;
                const size_t row0 = static_cast<size_t>(prm.ref_slot[r]) * prm.p_pad + (nt - r * dec.tpf) * kTile;
                mbar_wait_relaxed_s(lab_empty + 8 * ms, mph ^ 1, 128);
                if (elect_one()) {
                    mbar_arrive_expect_tx_s(lab_full + 8 * ms, kMetaTileBytes);
                    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                                 ::"r"(lab_smem + ms * kMetaTileBytes), "l"(reinterpret_cast<uint64_t>(prm.meta + row0 * kMetaFloats)),
                                   "r"(static_cast<uint32_t>(kMetaTileBytes)), "r"(lab_full + 8 * ms)
                                 : "memory");
                }
                __syncwarp();
            }
        }
    } else if (warp >= 4) {
        // ================= epilogue: TMEM lanes [32*(warp%4), +32); logit columns [32*sub, +32)
        const int quarter = warp & 3;
        const int sub = (warp - 4) >> 2;
        const int row = quarter * 32 + lane;
        const uint32_t lane_base = static_cast<uint32_t>(quarter * 32) << 16;
        const uint32_t bar_full = pin_reg(pp.acc_full), bar_empty = pin_reg(pp.acc_empty);
        const uint32_t tbase = pin_reg(pp.tmem_base + lane_base + static_cast<uint32_t>(sub * 32));
        const uint32_t lane_is0 = pin_reg(lane == 0 ? 1u : 0u);
        const uint32_t lab_mine = pin_reg(lab_smem + static_cast<uint32_t>(sub * 32) * (kMetaFloats * 4) + 8);   // V[0] of this warp's column 0
        const float scale2 = prm.scale2;
        const int last_valid = prm.n_pixels - (dec.tpf - 1) * kTile - sub * 32;   // real columns of this warp in a frame's last tile
        vosd::SegIter it(dec, blockIdx.x);
        int m_tile, n0, n1;
        uint32_t buf = 0, aphase = 0, count = 0;
        while (it.next(m_tile, n0, n1)) {
            idx_stage_target<kSplit>(pp, prm, it.seg, m_tile, row, lane_base, sub, kIdxSub);
            RowAcc<D> st;
            st.init();
            int j = n0 % dec.tpf;
            for (int nt = n0; nt < n1; ++nt, ++count) {
                mbar_wait_s(bar_full + 8 * buf, aphase);
                tc_fence_after_sync();
                const uint32_t taddr = tbase + buf * kTile;
                float va[kQC], vb[kQC];
                tmem_ld_32x32b_x16(taddr, va);
                tmem_ld_32x32b_x16(taddr + kQC, vb);
                tmem_ld_wait();
                tc_fence_before_sync();
                __syncwarp();
                if (lane_is0) mbar_arrive_s(bar_empty + 8 * buf);
                if (j == dec.tpf - 1 && last_valid < 32) {                     // ragged last tile of a frame: pad columns weigh nothing
#pragma unroll
                    for (int i = 0; i < kQC; ++i) {
                        if (i >= last_valid) va[i] = -INFINITY;
                        if (i + kQC >= last_valid) vb[i] = -INFINITY;
                    }
                }
                // ---- running maximum, exponentials in place, denominator (as tile_exps32)
                const float m_new = fmaxf(st.m, fmaxf(max16(va), max16(vb)) * scale2);
                if (m_new > st.m) {
                    const float corr = ex2(st.m - m_new);
                    st.l *= corr;
#pragma unroll
                    for (int c = 0; c < D; ++c) st.acc[c] *= corr;
                    st.m = m_new;
                }
                const float neg_m = -st.m;
                const float2 s2 = make_float2(scale2, scale2), nm2 = make_float2(neg_m, neg_m);
                float2 l2 = make_float2(0.f, 0.f);
#pragma unroll
                for (int i = 0; i < kQC; i += 2) {
                    const float2 ea = ffma2(make_float2(va[i], va[i + 1]), s2, nm2);
                    const float2 eb = ffma2(make_float2(vb[i], vb[i + 1]), s2, nm2);
                    const float2 pa = make_float2(ex2(ea.x), ex2(ea.y)), pb = make_float2(ex2(eb.x), ex2(eb.y));
                    l2 = fadd2(l2, fadd2(pa, pb));
                    va[i] = pa.x; va[i + 1] = pa.y;
                    vb[i] = pb.x; vb[i + 1] = pb.y;
                }
                st.l += l2.x + l2.y;
                // ---- label gather: warp-uniform 8-byte loads of the records of this warp's 32 columns
                const uint32_t ms = count % kProbLabelStages, mph = (count / kProbLabelStages) & 1;
                mbar_wait_s(lab_full + 8 * ms, mph);
                const uint32_t rec0 = lab_mine + ms * kMetaTileBytes;
#pragma unroll
                for (int i = 0; i < 2 * kQC; ++i) {
                    const float p = i < kQC ? va[i] : vb[i - kQC];
#pragma unroll
                    for (int c = 0; c < D; c += 2) {
                        float2 v;
                        asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"(rec0 + i * (kMetaFloats * 4) + c * 4));
                        st.acc[c] = fmaf(p, v.x, st.acc[c]);
                        if (c + 1 < D) st.acc[c + 1] = fmaf(p, v.y, st.acc[c + 1]);
                    }
                }
                __syncwarp();
                if (lane_is0) mbar_arrive_s(lab_empty + 8 * ms);
                if (++buf == Cfg::kAccBufs) { buf = 0; aphase ^= 1; }
                if (++j == dec.tpf) j = 0;
            }
            float* rec = prm.partials +
                         (static_cast<size_t>(blockIdx.x * dec.max_segs + it.seg) * kIdxSub + sub) * kPartFloats;
            store_partial<D>(st, rec, row);
        }
    }
    idx_teardown(pp);
}

}  // namespace vosk
