// Top-k propagation (EXTENSION of the reference, BASELINE.json north star / SURVEY.md H3): the softmax of
// predict.py:55 is restricted, per target pixel, to the k reference pixels with the largest logit
// (ties -> lowest reference index first); prior and label gather are unchanged.  Oracle: oracle.predict(topk=k).
//
//   vos_affinity_topk : same TMA -> tcgen05 -> TMEM pipeline as vos_affinity_idx, but the epilogue keeps a
//                       streaming top-k of (logit * temperature, reference index) per target pixel in shared
//                       memory -- the (N x P) affinity never reaches HBM.  One epilogue thread owns one target
//                       pixel and all 128 columns of a tile; candidates above the thread's running k-th value are
//                       appended to its column of a [slot][128] buffer; when a buffer fills, the warp prunes all of
//                       its 32 buffers to k entries (bisection on order-preserving integer keys + stable compaction).
//                       Output: per (CTA, segment) and target pixel the k best (key, index) pairs, index-ascending.
//   vos_topk_finish   : one warp per target pixel: merges the per-segment lists, selects the global top-k, orders
//                       it (value descending, index ascending), soft-maxes over the k logits, applies the prior
//                       in closed form, gathers the label records of the k reference pixels (64-byte coalesced
//                       loads) and writes prediction / arg-max / new labels / top-k indices.
//   vos_upsample_mask : stride-8 class map -> full-resolution uint8 mask (ATen legacy 'nearest').
#pragma once
#include "affinity_idx.cuh"
#include "topk_params.h"

namespace vosk {


// Shape of the top-k epilogue: kSub column groups per tile, 4 * kSub warps, one thread = one target pixel x 128 / kSub
// columns with its own buffer.  kSub = 1: 112 slots (any k <= 64), one warp per scheduler -- latency bound
// (profiles/README.md).  More warps hide that latency but shared memory caps the slots per thread (a buffer needs
// k + 16 slots plus slack between prunes): kSub = 2: 64 slots (k <= 24), kSub = 4: 36 slots (k <= 8) -- at the price
// of kSub lists per (CTA, segment, pixel) for the finish kernel to merge.
template <int kSub>
struct TopkCfg {
    static constexpr int kEpiWarps = 4 * kSub;
    static constexpr int kThreads = 64 + 32 * kEpiWarps;      // warp 0 TMA, warp 1 MMA, then the epilogue warps
    static constexpr int kBuf = kSub == 1 ? 112 : (kSub == 2 ? 64 : 36);   // candidate slots per thread
    static constexpr int kStages = kSub == 1 ? 3 : 2;         // x 32 KiB of reference chunks in flight
    static constexpr uint32_t kStride = 512u * kSub;          // bytes between consecutive slots of one thread
    static constexpr int kCols = kTile / kSub;                // logit columns per thread and tile
    static constexpr int kSmem = kStages * kTopkGroup * kChunkBytes + 512 + 1024 + kBuf * kTile * kSub * 8;
};

// order-preserving map float -> uint32 (larger float <=> larger key); -0.0 must be normalised to +0.0 by the caller
__device__ __forceinline__ uint32_t f2key(float f) {
    const uint32_t u = __float_as_uint(f);
    return u ^ ((u >> 31) ? 0xffffffffu : 0x80000000u);
}
__device__ __forceinline__ float key2f(uint32_t k) {
    return __uint_as_float((k & 0x80000000u) ? (k ^ 0x80000000u) : ~k);
}
__device__ __forceinline__ uint32_t lds_u32(uint32_t addr) {
    uint32_t v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr));
    return v;
}
__device__ __forceinline__ void sts_u32(uint32_t addr, uint32_t v) {
    asm volatile("st.shared.u32 [%0], %1;" ::"r"(addr), "r"(v) : "memory");
}

// Warp-synchronous prune of the 32 candidate buffers of a warp to their k best entries.
// Buffer of a thread: slots e = 0..cnt-1 at kbase + kStride*e (keys) / ibase + kStride*e (indices), index-ascending.
// Order: key descending, then index ascending (= slot order among equal keys).  tau <- the k-th best key.
// On entry tau is a lower bound of every key in the buffer (the k-th best of the last prune, or the key of -inf) and
// kmax the largest key ever pushed, so no pass is needed to find the search interval; pivots alternate between
// interpolation on the counts and plain bisection, and the bounds snap to actual keys after every pass (typically 4-6
// passes over the buffer instead of ~14 with min/max + pure bisection + a separate count).
template <uint32_t kStride>
__device__ __forceinline__ void topk_prune(uint32_t kbase, uint32_t ibase, int& cnt, int k, uint32_t& tau, uint32_t kmax) {
    const uint32_t full = 0xffffffffu;
    const int cmax = __reduce_max_sync(full, cnt);
    const bool active = cnt > k;
    // f(t) = #{key >= t}.  Invariant: f(lo) >= k (cge = f(lo)), #{key > hi} < k (cgt), hi is a key.
    uint32_t lo = tau, hi = kmax;
    int cge = cnt, cgt = 0;
    bool interp = true;
    while (__any_sync(full, active && lo < hi)) {
        uint32_t mid = lo + ((hi - lo) >> 1) + 1u;           // lo < mid <= hi
        if (interp) {   // keys roughly uniform in (lo, hi]: the k-th best sits (cge - k) / (cge - cgt) of the way up
            const float t = (static_cast<float>(cge - k) + 0.5f) / static_cast<float>(cge - cgt);
            const uint32_t off = static_cast<uint32_t>(t * static_cast<float>(hi - lo));
            mid = lo + min(max(off, 1u), hi - lo);
        }
        interp = !interp;
        int c = 0;
        uint32_t mn = 0xffffffffu, mxb = 0u;                 // smallest key >= mid, largest key < mid
        for (int e = 0; e < cmax; ++e) {
            if (e < cnt) {
                const uint32_t key = lds_u32(kbase + kStride * e);
                if (key >= mid) { ++c; mn = min(mn, key); }
                else mxb = max(mxb, key);
            }
        }
        if (active && lo < hi) {
            if (c >= k) { lo = mn; cge = c; }
            else { hi = mxb; cgt = c; }
        }
    }
    if (!__any_sync(full, active)) return;
    int need = k - cgt, w = 0;                               // ties at the k-th key: the first `need` in slot order
    for (int e = 0; e < cmax; ++e) {
        if (active && e < cnt) {
            const uint32_t key = lds_u32(kbase + kStride * e);
            const uint32_t idx = lds_u32(ibase + kStride * e);
            bool keep = key > lo;
            if (key == lo && need > 0) { keep = true; --need; }
            if (keep) {
                sts_u32(kbase + kStride * w, key);
                sts_u32(ibase + kStride * w, idx);
                ++w;
            }
        }
    }
    if (active) { cnt = w; tau = lo; }
}

template <bool kSplit, int kSub>
__global__ void __launch_bounds__(TopkCfg<kSub>::kThreads, 1)
vos_affinity_topk(const __grid_constant__ CUtensorMap tmap_hi, const __grid_constant__ CUtensorMap tmap_lo,
                  const __grid_constant__ AffinityParams prm) {
    using Cfg = IdxCfg<kSplit>;
    using TC = TopkCfg<kSub>;
    constexpr int kTopkStages = TC::kStages;
    constexpr uint32_t kStride = TC::kStride;
    extern __shared__ uint8_t smem_raw[];
    const IdxPipe pp = idx_setup<kTopkGroup, kTopkStages>(smem_raw, &tmap_hi, &tmap_lo, Cfg::kAccBufs, TC::kEpiWarps);
    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const vosd::Decomp dec = vosd::make_decomp(prm.n_pixels, prm.n_refs, prm.num_sms);

    if (warp == 0) {
        idx_role_producer<kSplit, kTopkGroup, kTopkStages>(pp, &tmap_hi, &tmap_lo, prm, dec);
    } else if (warp == 1) {
        idx_role_mma<kSplit, kTopkGroup, kTopkStages>(pp, prm, dec);
    } else {
        // ================= epilogue: warp w owns TMEM lanes [32*(w%4), +32) = 32 target pixels and the logit columns
        // [kCols*sub, +kCols) of every tile
        const uint32_t full = 0xffffffffu;
        const int quarter = warp & 3;
        const int sub = (warp - 2) >> 2;
        const int row = quarter * 32 + lane;
        const uint32_t lane_base = static_cast<uint32_t>(quarter * 32) << 16;
        const uint32_t buf_base = (pp.acc_empty + 8 * kIdxMaxAccBufs + 16 + 127) & ~127u;
        const uint32_t kbase = buf_base + 4u * (sub * kTile + row);                                  // keys    [kBuf][128 * kSub]
        const uint32_t ibase = buf_base + TC::kBuf * kTile * kSub * 4 + 4u * (sub * kTile + row);    // indices [kBuf][128 * kSub]
        const int k = prm.topk;
        const float temperature = prm.temperature;
        vosd::SegIter it(dec, blockIdx.x);
        int m_tile, n0, n1;
        uint32_t buf = 0, aphase = 0;
        long long n_steps = 0, n_slow = 0, n_push = 0, n_prune = 0;     // profiling counters (vosprop_debug_clocks)
        while (it.next(m_tile, n0, n1)) {
            idx_stage_target<kSplit>(pp, prm, it.seg, m_tile, row, lane_base, sub, kSub);
            int cnt = 0;
            uint32_t tau = f2key(-INFINITY);       // key of the running k-th best (a lower bound of every buffered key)
            uint32_t kmax = 0u;                    // largest key pushed in this segment
            float tau_f = -INFINITY;
            int r = n0 / dec.tpf;
            int j = n0 - r * dec.tpf;
            for (int nt = n0; nt < n1; ++nt) {
                const int n_tile = r * prm.n_pixels + j * kTile + sub * TC::kCols;   // reference index (r*P + pixel) of this thread's column 0
                const int cols = min(kTile, prm.n_pixels - j * kTile) - sub * TC::kCols;  // its real columns in this tile
                mbar_wait_s(pp.acc_full + 8 * buf, aphase);
                tc_fence_after_sync();
                const uint32_t taddr = pp.tmem_base + lane_base + buf * kTile + sub * TC::kCols;
#pragma unroll 1
                for (int s = 0; s < TC::kCols / kQC; ++s) {
                    float v[kQC];
                    tmem_ld_32x32b_x16(taddr + s * kQC, v);
                    tmem_ld_wait();
                    if (s == TC::kCols / kQC - 1) {                           // this thread's columns of the tile are consumed
                        tc_fence_before_sync();
                        __syncwarp();
                        if (lane == 0) mbar_arrive_s(pp.acc_empty + 8 * buf);  // one arrival per warp
                    }
                    const int nv = cols - s * kQC;
                    if (nv <= 0) continue;
                    ++n_steps;
#pragma unroll
                    for (int i = 0; i < kQC; ++i) {
                        v[i] = v[i] * temperature + 0.0f;                     // predict.py:52 (fp32 product); -0 -> +0
                        if (i >= nv) v[i] = -INFINITY;
                    }
                    if (!__any_sync(full, max16(v) > tau_f)) continue;        // nothing beats any lane's k-th best
                    ++n_slow;
#pragma unroll
                    for (int i = 0; i < kQC; ++i) {
                        if (v[i] > tau_f) {
                            const uint32_t key = f2key(v[i]);
                            kmax = max(kmax, key);
                            ++n_push;
                            sts_u32(kbase + kStride * cnt, key);
                            sts_u32(ibase + kStride * cnt, static_cast<uint32_t>(n_tile + s * kQC + i));
                            ++cnt;
                        }
                    }
                    if (__any_sync(full, cnt > TC::kBuf - kQC)) {
                        ++n_prune;
                        topk_prune<kStride>(kbase, ibase, cnt, k, tau, kmax);
                        tau_f = key2f(tau);
                    }
                }
                if (++buf == Cfg::kAccBufs) { buf = 0; aphase ^= 1; }
                if (++j == dec.tpf) { j = 0; ++r; }
            }
            topk_prune<kStride>(kbase, ibase, cnt, k, tau, kmax);
            // ---- this thread's list of the target pixel for this segment: cnt <= k entries, index-ascending
            const size_t rec = (static_cast<size_t>(blockIdx.x * dec.max_segs + it.seg) * kSub + sub) * kTile + row;
            prm.cand_cnt[rec] = cnt;
            uint32_t* ck = prm.cand_key + rec * kTopkMax;
            int32_t* ci = prm.cand_idx + rec * kTopkMax;
            for (int e = 0; e < cnt; ++e) {
                ck[e] = lds_u32(kbase + kStride * e);
                ci[e] = static_cast<int32_t>(lds_u32(ibase + kStride * e));
            }
        }
        if (prm.dbg_clk && warp == 2 && lane == 0) {
            prm.dbg_clk[blockIdx.x * 16 + 9] = n_steps;
            prm.dbg_clk[blockIdx.x * 16 + 10] = n_slow;
            prm.dbg_clk[blockIdx.x * 16 + 11] = n_push;
            prm.dbg_clk[blockIdx.x * 16 + 12] = n_prune;
        }
    }
    idx_teardown(pp);
}

// -------------------------------------------------------------------------------------------
// Finish: one warp per target pixel.
// -------------------------------------------------------------------------------------------


__global__ void __launch_bounds__(kFinishWarps * 32) vos_topk_finish(const TopkFinishParams fp) {
    extern __shared__ uint32_t fsm[];
    const MergeParams& prm = fp.mp;
    const uint32_t full = 0xffffffffu;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int pix = blockIdx.x * kFinishWarps + warp;
    if (pix >= prm.n_pixels) return;
    uint32_t* keys = fsm + warp * (kTopkMaxCand * 2 + kTopkMax * 2);   // [C]
    uint32_t* idxs = keys + kTopkMaxCand;                               // [C]
    uint32_t* skey = idxs + kTopkMaxCand;                               // [k] selected
    uint32_t* sidx = skey + kTopkMax;
    const vosd::Decomp dec = vosd::make_decomp(prm.n_pixels, prm.n_refs, prm.num_sms);
    const int k = fp.topk;
    const int mt = pix / kTile, row = pix % kTile;
    const int64_t lin_lo = static_cast<int64_t>(mt) * dec.nt;
    const int c_first = vosd::cta_of(dec, lin_lo), c_last = vosd::cta_of(dec, lin_lo + dec.nt - 1);
    // ---- gather the lists of every (CTA, segment, column group) that saw this pixel's row
    int C = 0;
    for (int c = c_first; c <= c_last; ++c) {
        const int seg = mt - static_cast<int>(vosd::cta_begin(dec, c) / dec.nt);
        for (int sub = 0; sub < prm.n_sub; ++sub) {
            const size_t rec = (static_cast<size_t>(c * dec.max_segs + seg) * prm.n_sub + sub) * kTile + row;
            const int n = fp.cand_cnt[rec];
            for (int e = lane; e < n; e += 32) {
                keys[C + e] = fp.cand_key[rec * kTopkMax + e];
                idxs[C + e] = static_cast<uint32_t>(fp.cand_idx[rec * kTopkMax + e]);
            }
            C += n;
        }
    }
    __syncwarp();
    // ---- k-th best key by bisection over the keys (warp-cooperative counts)
    int n_sel = C;
    if (C > k) {
        uint32_t lo = 0xffffffffu, hi = 0u;
        for (int e = lane; e < C; e += 32) { lo = min(lo, keys[e]); hi = max(hi, keys[e]); }
        lo = __reduce_min_sync(full, lo);
        hi = __reduce_max_sync(full, hi);
        while (lo < hi) {
            const uint32_t mid = lo + ((hi - lo) >> 1) + 1u;
            int c = 0;
            uint32_t mn = 0xffffffffu, mxb = 0u;
            for (int e = lane; e < C; e += 32) {
                const uint32_t key = keys[e];
                if (key >= mid) { ++c; mn = min(mn, key); }
                else mxb = max(mxb, key);
            }
            c = __reduce_add_sync(full, c);
            if (c >= k) lo = __reduce_min_sync(full, mn);
            else hi = __reduce_max_sync(full, mxb);
        }
        int gt = 0;
        for (int e = lane; e < C; e += 32) gt += keys[e] > lo;
        gt = __reduce_add_sync(full, gt);
        // ties at the k-th key: the `need` lowest reference indices among them (the lists interleave column groups, so
        // list order is not index order): largest index still kept = the need-th smallest tie index, by bisection
        const int need = k - gt;
        int n_tie = 0;
        for (int e = lane; e < C; e += 32) n_tie += keys[e] == lo;
        n_tie = __reduce_add_sync(full, n_tie);
        uint32_t idx_cut = 0xffffffffu;
        if (n_tie > need) {
            uint32_t ilo = 0u, ihi = 0xffffffffu;               // smallest cut with #{tie, idx <= cut} >= need
            while (ilo < ihi) {
                const uint32_t mid = ilo + ((ihi - ilo) >> 1);
                int c = 0;
                for (int e = lane; e < C; e += 32) c += keys[e] == lo && idxs[e] <= mid;
                c = __reduce_add_sync(full, c);
                if (c >= need) ihi = mid; else ilo = mid + 1u;
            }
            idx_cut = ilo;
        }
        int w = 0;
        for (int e0 = 0; e0 < C; e0 += 32) {
            const int e = e0 + lane;
            const uint32_t key = e < C ? keys[e] : 0u;
            const bool keep = e < C && (key > lo || (key == lo && idxs[e] <= idx_cut));
            const uint32_t kmask = __ballot_sync(full, keep);
            if (keep) {
                const int pos = w + __popc(kmask & ((1u << lane) - 1u));
                skey[pos] = key;
                sidx[pos] = idxs[e];
            }
            w += __popc(kmask);
        }
        n_sel = w;
    } else {
        for (int e = lane; e < C; e += 32) { skey[e] = keys[e]; sidx[e] = idxs[e]; }
    }
    __syncwarp();
    // ---- order for the index output: value descending, reference index ascending
    if (fp.out_topk_idx) {
        for (int i = lane; i < k; i += 32) {
            if (i >= n_sel) fp.out_topk_idx[static_cast<size_t>(pix) * k + i] = -1;
        }
        for (int i = lane; i < n_sel; i += 32) {
            const uint32_t ki = skey[i];
            int rank = 0;
            const uint32_t ii = sidx[i];
            for (int jj = 0; jj < n_sel; ++jj) rank += (skey[jj] > ki) || (skey[jj] == ki && sidx[jj] < ii);
            fp.out_topk_idx[static_cast<size_t>(pix) * k + rank] = static_cast<int32_t>(sidx[i]);
        }
    }
    // ---- softmax over the selected logits, prior, label gather
    float tmax = -INFINITY;
    for (int i = lane; i < n_sel; i += 32) tmax = fmaxf(tmax, key2f(skey[i]));
    tmax = fmaxf(tmax, __shfl_xor_sync(full, tmax, 16));
    tmax = fmaxf(tmax, __shfl_xor_sync(full, tmax, 8));
    tmax = fmaxf(tmax, __shfl_xor_sync(full, tmax, 4));
    tmax = fmaxf(tmax, __shfl_xor_sync(full, tmax, 2));
    tmax = fmaxf(tmax, __shfl_xor_sync(full, tmax, 1));
    float L = 0.f, acc[kMetaClasses];
#pragma unroll
    for (int c = 0; c < kMetaClasses; ++c) acc[c] = 0.f;
    float rm, xm;
    pixel_coord(pix, prm.w_lowres, rm, xm);
    for (int i = lane; i < n_sel; i += 32) {
        const float p = vosptx::ex2((key2f(skey[i]) - tmax) * kLog2e);
        L += p;
        const int n = static_cast<int>(sidx[i]);
        const int r = n / prm.n_pixels, px = n - r * prm.n_pixels;
        float rn, xn;
        pixel_coord(px, prm.w_lowres, rn, xn);
        const float dr = rn - rm, dx = xn - xm;
        const float pw = p * vosptx::ex2(-fp.ref_coef[r] * fmaf(dx, dx, dr * dr));
        const float4* rec = reinterpret_cast<const float4*>(prm.meta + (static_cast<size_t>(fp.ref_slot[r]) * prm.p_pad + px) * kMetaFloats);
        const float4 a = rec[0], b = rec[1], c4 = rec[2], d4 = rec[3];     // {rowf, xf, V0, V1}, V2..5, V6..9, V10..13
        acc[0] = fmaf(pw, a.z, acc[0]);  acc[1] = fmaf(pw, a.w, acc[1]);
        acc[2] = fmaf(pw, b.x, acc[2]);  acc[3] = fmaf(pw, b.y, acc[3]);  acc[4] = fmaf(pw, b.z, acc[4]);  acc[5] = fmaf(pw, b.w, acc[5]);
        acc[6] = fmaf(pw, c4.x, acc[6]); acc[7] = fmaf(pw, c4.y, acc[7]); acc[8] = fmaf(pw, c4.z, acc[8]); acc[9] = fmaf(pw, c4.w, acc[9]);
        acc[10] = fmaf(pw, d4.x, acc[10]); acc[11] = fmaf(pw, d4.y, acc[11]); acc[12] = fmaf(pw, d4.z, acc[12]); acc[13] = fmaf(pw, d4.w, acc[13]);
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
        L += __shfl_xor_sync(full, L, off);
#pragma unroll
        for (int c = 0; c < kMetaClasses; ++c) acc[c] += __shfl_xor_sync(full, acc[c], off);
    }
    if (lane != 0) return;
    const float inv = 1.0f / L;
    int best = 0;
    float best_v = -INFINITY;
    float* mrec = prm.meta + (static_cast<size_t>(prm.q_slot) * prm.p_pad + pix) * kMetaFloats + 2;
#pragma unroll
    for (int c = 0; c < kMetaClasses; ++c) {
        if (c < prm.d) {
            const float pk = acc[c] * inv;
            acc[c] = pk;
            if (pk > best_v) { best_v = pk; best = c; }   // strict '>' : first maximum wins
            if (prm.out_prediction) prm.out_prediction[static_cast<size_t>(c) * prm.n_pixels + pix] = pk;
        }
    }
    if (prm.write_labels) {
#pragma unroll
        for (int c = 0; c < kMetaClasses; ++c)
            mrec[c] = (c < prm.d) ? (prm.probability ? acc[c] : (c == best ? 1.f : 0.f)) : 0.f;
        prm.cls[static_cast<size_t>(prm.q_slot) * prm.p_pad + pix] = static_cast<uint8_t>(best);
    }
    if (prm.out_mask_lowres) prm.out_mask_lowres[pix] = static_cast<uint8_t>(best);
}

// stride-8 class map (P) -> full-resolution mask (H, W): src = min(floor(dst * in/out), in - 1) in fp32 (ATen legacy 'nearest')
__global__ void __launch_bounds__(256) vos_upsample_mask(const uint8_t* __restrict__ low, uint8_t* __restrict__ out,
                                                         int h_lowres, int w_lowres, int H, int W) {
    const float sy = static_cast<float>(h_lowres) / static_cast<float>(H);
    const float sx = static_cast<float>(w_lowres) / static_cast<float>(W);
    const int dy = blockIdx.x;
    const uint8_t* src = low + static_cast<size_t>(nearest_src(dy, sy, h_lowres)) * w_lowres;
    for (int dx = threadIdx.x; dx < W; dx += blockDim.x) out[static_cast<size_t>(dy) * W + dx] = src[nearest_src(dx, sx, w_lowres)];
}

}  // namespace vosk
