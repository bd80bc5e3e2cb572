// Top-k propagation (EXTENSION of the reference, BASELINE.json north star / SURVEY.md H3): the softmax of
// predict.py:55 is restricted, per target pixel, to the k reference pixels with the largest logit
// (ties -> lowest reference index first); prior and label gather are unchanged.  Oracle: oracle.predict(topk=k).
//
// Two passes over the tensor cores (the (N x P) affinity never reaches HBM in either), because what made the one-pass
// streaming selection of round 1 slow was not the selection but the start: until a row has seen its k best logits
// nearly every logit is a candidate, and every candidate costs a divergent push.  A scan of the affinity without
// arithmetic is cheap (1 300 cycles per 128 x 128 tile against 2 700 for the full softmax), so:
//
//   vos_topk_scan<pass 1>  : TMA -> tcgen05 -> TMEM pipeline of vos_affinity_idx; the epilogue (16 warps, one thread = one
//                            target pixel x 32 columns per tile) only takes the maximum of its 32 logits:
//                            bound[(tile * 4 + column group) * 128 + row].  17 instructions per thread and tile.
//   vos_topk_threshold     : per target pixel the k-th largest of its block maxima, tau.  At least k logits of the row are
//                            >= tau (k different blocks hold one), so tau is a LOWER bound of the row's k-th largest logit:
//                            every top-k member is >= tau.  Measured on the synthetic 480p clips (57 780 logits per row):
//                            5.5 / 36 / 147 logits >= tau for k = 5 / 20 / 50.
//   vos_topk_scan<pass 2>  : the same contraction again (bit-identical logits: same tiles, same MMA order); a logit is
//                            looked at only if it is >= tau (a maximum tree and one vote per 16 columns otherwise) and then
//                            appended to its thread's list in global memory.  A list that fills up is pruned to its k best
//                            (key descending, index ascending) and the thread's own threshold rises.
//   vos_topk_finish        : one warp per target pixel: merges the lists (4 column groups x the CTAs of the row), selects
//                            and orders the global top-k, soft-maxes over the k logits, applies the prior in closed form,
//                            gathers the label records (64-byte loads), writes prediction / arg-max / new labels / indices.
//   vos_upsample_mask      : stride-8 class map -> full-resolution uint8 mask (ATen legacy 'nearest').
// Selection compares fl(logit * temperature) exactly as torch computes it (predict.py:52); ties at the k-th value go to
// the lowest reference indices (each list keeps its first k ties in index order: it scans indices in ascending order).
#pragma once
#include "affinity_idx.cuh"
#include "topk_params.h"

namespace vosk {

// order-preserving map float -> uint32 (larger float <=> larger key); -0.0 must be normalised to +0.0 by the caller
__device__ __forceinline__ uint32_t f2key(float f) {
    const uint32_t u = __float_as_uint(f);
    return u ^ ((u >> 31) ? 0xffffffffu : 0x80000000u);
}
__device__ __forceinline__ float key2f(uint32_t k) {
    return __uint_as_float((k & 0x80000000u) ? (k ^ 0x80000000u) : ~k);
}

// Smallest raw logit x with  fl(x * T) + 0 >= tau  (T >= 0; fl(x * T) is monotone in x): the scan compares raw logits
// against it instead of multiplying every logit by the temperature.
__device__ __forceinline__ float raw_threshold(float tau, float T) {
    // "everything passes" is the lowest FINITE float: the -inf that masks the pad columns of a ragged tile must fail the test
    if (T == 0.f) return tau <= 0.f ? -3.402823466e38f : INFINITY;  // every product is +0
    if (tau == -INFINITY) return -3.402823466e38f;
    if (!(tau < INFINITY)) return INFINITY;                          // +inf (rows beyond the frame) or NaN: nothing passes
    float x = __fdiv_rn(tau, T);
    for (int i = 0; i < 4 && !(x * T + 0.f >= tau); ++i) x = nextafterf(x, INFINITY);
    for (int i = 0; i < 4; ++i) {
        const float y = nextafterf(x, -INFINITY);
        if (!(y * T + 0.f >= tau)) break;
        x = y;
    }
    return x;
}

// A thread's candidate list lives in global memory, interleaved with the lists of its warp: slot e of lane l at
// (list_group * kTopkCap + e) * 32 + l  with list_group = record / 32 -- slot-uniform accesses of a warp coalesce.
__device__ __forceinline__ size_t topk_slot(size_t rec, int e) {
    return ((rec >> 5) * kTopkCap + static_cast<size_t>(e)) * 32 + (rec & 31);
}

// Warp-synchronous prune of the 32 lists of a warp to their k best entries (lists with more than k entries only).
// Order: key descending, then index ascending (= slot order among equal keys: a thread scans indices in ascending order).
// tau <- the k-th best key.  On entry tau is a lower bound of every key in the list and kmax the largest key ever pushed,
// so no pass is needed to find the search interval; pivots alternate between interpolation on the counts and plain
// bisection, and the bounds snap to actual keys after every pass (typically 4-6 passes).
__device__ __forceinline__ void topk_prune(uint32_t* __restrict__ keys, int32_t* __restrict__ idxs, int& cnt, int k,
                                           uint32_t& tau, uint32_t kmax) {
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
                const uint32_t key = keys[e * 32];
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
            const uint32_t key = keys[e * 32];
            const int32_t idx = idxs[e * 32];
            bool keep = key > lo;
            if (key == lo && need > 0) { keep = true; --need; }
            if (keep) {
                keys[w * 32] = key;
                idxs[w * 32] = idx;
                ++w;
            }
        }
    }
    if (active) { cnt = w; tau = lo; }
}

// kPass = 1: block maxima (prm.topk_bound).  kPass = 2: candidate lists (prm.cand_*), thresholds from prm.topk_tau.
template <bool kSplit, int kPass>
__global__ void __launch_bounds__(kIdxThreads, 1)
vos_topk_scan(const __grid_constant__ CUtensorMap tmap_hi, const __grid_constant__ CUtensorMap tmap_lo,
              const __grid_constant__ AffinityParams prm) {
    using Cfg = IdxCfg<kSplit>;
    extern __shared__ uint8_t smem_raw[];
    pdl_launch_dependents();
    const IdxPipe pp = idx_setup<Cfg::kGroup, Cfg::kStages>(smem_raw, &tmap_hi, &tmap_lo, Cfg::kAccBufs, kIdxEpiWarps);
    pdl_wait();
    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const vosd::Decomp dec = vosd::make_decomp(prm.n_pixels, prm.n_refs, prm.num_sms);

    if (warp == 0) {
        idx_role_producer<kSplit, Cfg::kGroup, Cfg::kStages>(pp, &tmap_hi, &tmap_lo, prm, dec);
    } else if (warp == 1) {
        idx_role_mma<kSplit, Cfg::kGroup, Cfg::kStages>(pp, prm, dec);
    } else {
        // ================= epilogue: warp w owns TMEM lanes [32*(w%4), +32) = 32 target pixels and the logit columns
        // [32*sub, +32) of every tile
        const uint32_t full = 0xffffffffu;
        const int quarter = warp & 3;
        const int sub = (warp - 2) >> 2;
        const int row = quarter * 32 + lane;
        const uint32_t lane_base = static_cast<uint32_t>(quarter * 32) << 16;
        const uint32_t bar_full = pin_reg(pp.acc_full), bar_empty = pin_reg(pp.acc_empty);
        const uint32_t tbase = pin_reg(pp.tmem_base + lane_base + static_cast<uint32_t>(sub * 32));
        const uint32_t lane_is0 = pin_reg(lane == 0 ? 1u : 0u);
        const int k = prm.topk;
        const float temperature = prm.temperature;
        const int last_valid = prm.n_pixels - (dec.tpf - 1) * kTile - sub * 32;   // real columns of this warp in a frame's last tile
        vosd::SegIter it(dec, blockIdx.x);
        int m_tile, n0, n1;
        uint32_t buf = 0, aphase = 0;
        while (it.next(m_tile, n0, n1)) {
            idx_stage_target<kSplit>(pp, prm, it.seg, m_tile, row, lane_base, sub, kIdxSub);
            const size_t rec = (static_cast<size_t>(blockIdx.x * dec.max_segs + it.seg) * kIdxSub + sub) * kTile + row;
            // ---- pass 2 state: this thread's list and thresholds
            uint32_t* lkeys = nullptr;
            int32_t* lidx = nullptr;
            int cnt = 0;
            uint32_t tau_key = 0u, kmax = 0u;
            float tcmp = INFINITY;             // raw logits >= tcmp are candidates
            if constexpr (kPass == 2) {
                lkeys = prm.cand_key + topk_slot(rec, 0);
                lidx = prm.cand_idx + topk_slot(rec, 0);
                const int pix = m_tile * kTile + row;
                const float tau = pix < prm.n_pixels ? prm.topk_tau[pix] : INFINITY;     // rows beyond the frame collect nothing
                tau_key = f2key(tau + 0.f);
                tcmp = raw_threshold(tau, temperature);
            }
            float* bound = nullptr;
            const int tstep = kPass == 1 ? prm.tile_step : 1;            // pass 1 may visit every tstep-th reference tile only
            const int nt_first = n0 + (tstep - n0 % tstep) % tstep;
            if constexpr (kPass == 1)
                bound = prm.topk_bound + ((static_cast<size_t>(m_tile) * dec.nt + nt_first) * kIdxSub + sub) * kTile + row;
            int r = nt_first / dec.tpf;
            int j = nt_first - r * dec.tpf;
            for (int nt = nt_first; nt < n1; nt += tstep) {
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
                if (j == dec.tpf - 1 && last_valid < 32) {                     // ragged last tile of a frame: pad columns never win
#pragma unroll
                    for (int i = 0; i < kQC; ++i) {
                        if (i >= last_valid) va[i] = -INFINITY;
                        if (i + kQC >= last_valid) vb[i] = -INFINITY;
                    }
                }
                const float ma = max16(va), mb = max16(vb);
                if constexpr (kPass == 1) {
                    *bound = fmaxf(ma, mb) * temperature + 0.0f;            // predict.py:52 (fp32 product); -0 -> +0
                    bound += tstep * kIdxSub * kTile;
                } else {
                    const int n_tile = r * prm.n_pixels + j * kTile + sub * 32;   // reference index (r*P + pixel) of this thread's column 0
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        const float (&v)[kQC] = h ? vb : va;
                        if (!__any_sync(full, (h ? mb : ma) >= tcmp)) continue;     // nothing reaches any lane's threshold
#pragma unroll
                        for (int i = 0; i < kQC; ++i) {
                            const bool hit = v[i] >= tcmp;
                            if (__any_sync(full, hit)) {                            // warp-uniform: columns without a hit cost a vote
                                if (hit) {
                                    const uint32_t key = f2key(v[i] * temperature + 0.0f);
                                    kmax = max(kmax, key);
                                    lkeys[cnt * 32] = key;
                                    lidx[cnt * 32] = n_tile + h * kQC + i;
                                    ++cnt;
                                }
                            }
                        }
                        if (__any_sync(full, cnt > kTopkCap - kQC)) {
                            const int before = cnt;
                            topk_prune(lkeys, lidx, cnt, k, tau_key, kmax);
                            // after a prune the list holds k entries >= tau, all with lower indices than anything still to
                            // come: a later tie at tau cannot enter any more -> strictly greater from now on
                            if (cnt < before) tcmp = raw_threshold(nextafterf(key2f(tau_key), INFINITY), temperature);
                        }
                    }
                }
                if (++buf == Cfg::kAccBufs) { buf = 0; aphase ^= 1; }
                j += tstep;
                while (j >= dec.tpf) { j -= dec.tpf; ++r; }
            }
            if constexpr (kPass == 2) {
                if (__any_sync(full, cnt > k)) topk_prune(lkeys, lidx, cnt, k, tau_key, kmax);
                prm.cand_cnt[rec] = cnt;          // this thread's list of the target pixel for this segment: index-ascending
            }
        }
    }
    idx_teardown(pp);
}

// Per target pixel the k-th largest of its block maxima (pass 1) -> tau (a lower bound of the row's k-th largest
// fl(logit * temperature)); -inf when the row has fewer than k blocks.  One CTA per `rows` consecutive pixels of one
// target tile: the maxima are brought into shared memory transposed (rows x blocks; the global layout is blocks x 128 rows,
// so `rows` = 8 pixels read whole 32-byte sectors), then one warp per row runs a radix select on the order-preserving
// keys: four passes of 8 bits, each a 256-bin histogram of the keys that still match the prefix (shared-memory atomics)
// and a suffix scan over the bins (8 per lane) for the digit that holds the k-th largest.
// tile_step: pass 1 visited every tile_step-th reference tile of a row only (blocks of the others are not read).
constexpr int kThrWarps = 8;
__global__ void __launch_bounds__(kThrWarps * 32) vos_topk_threshold(const float* __restrict__ bound, float* __restrict__ tau,
                                                                     int n_pixels, int n_tiles, int tile_step, int k, int rows_log2) {
    extern __shared__ uint32_t tkeys[];                    // [rows][stride] keys, then [kThrWarps][256] histograms
    const uint32_t full = 0xffffffffu;
    const int rows = 1 << rows_log2;
    const int n_vis = (n_tiles + tile_step - 1) / tile_step;          // visited tiles per row
    const int n_blocks = n_vis * kIdxSub;
    const int stride = n_blocks | 1;                       // odd: the transposing stores spread over the banks
    uint32_t* hist_all = tkeys + rows * stride;
    const int pix0 = blockIdx.x << rows_log2;
    const int mt = pix0 / kTile, row0 = pix0 % kTile;
    const float* src = bound + static_cast<size_t>(mt) * n_tiles * kIdxSub * kTile + row0;
    for (int i = threadIdx.x; i < n_blocks << rows_log2; i += blockDim.x) {
        const int b = i >> rows_log2, rr = i & (rows - 1);
        const int t = (b >> 2) * tile_step, sub = b & 3;
        tkeys[rr * stride + b] = f2key(src[static_cast<size_t>(t * kIdxSub + sub) * kTile + rr] + 0.0f);
    }
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    uint32_t* hist = hist_all + warp * 256;
    for (int rr = warp; rr < rows; rr += kThrWarps) {
        const int pix = pix0 + rr;
        if (pix >= n_pixels) continue;                     // warp-uniform
        const uint32_t* keys = tkeys + rr * stride;
        float result = -INFINITY;
        if (n_blocks >= k) {
            uint32_t prefix = 0u, mask = 0u;
            int kk = k;                                    // rank (from the top) of the wanted key among the keys matching the prefix
#pragma unroll 1
            for (int shift = 24; shift >= 0; shift -= 8) {
#pragma unroll
                for (int i = 0; i < 8; ++i) hist[lane * 8 + i] = 0u;
                __syncwarp();
                for (int e = lane; e < n_blocks; e += 32) {
                    const uint32_t key = keys[e];
                    if ((key & mask) == prefix) atomicAdd(&hist[(key >> shift) & 255u], 1u);
                }
                __syncwarp();
                uint32_t c[8], mine = 0u;
#pragma unroll
                for (int i = 0; i < 8; ++i) { c[i] = hist[lane * 8 + i]; mine += c[i]; }
                uint32_t above = mine;                     // inclusive suffix sum over the lanes: bins >= 8 * lane
#pragma unroll
                for (int off = 1; off < 32; off <<= 1) {
                    const uint32_t o = __shfl_down_sync(full, above, off);
                    if (lane + off < 32) above += o;
                }
                above -= mine;                             // keys in the bins of higher lanes
                const bool here = above < static_cast<uint32_t>(kk) && static_cast<uint32_t>(kk) <= above + mine;
                uint32_t digit = 0u, higher = 0u;
                if (here) {
                    uint32_t acc = above;
#pragma unroll
                    for (int i = 7; i >= 0; --i) {
                        if (acc < static_cast<uint32_t>(kk) && static_cast<uint32_t>(kk) <= acc + c[i]) { digit = lane * 8 + i; higher = acc; }
                        acc += c[i];
                    }
                }
                const int src_lane = __ffs(__ballot_sync(full, here)) - 1;
                digit = __shfl_sync(full, digit, src_lane);
                higher = __shfl_sync(full, higher, src_lane);
                prefix |= digit << shift;
                mask |= 255u << shift;
                kk -= static_cast<int>(higher);
                __syncwarp();
            }
            result = key2f(prefix);
        }
        if (lane == 0) tau[pix] = result;
    }
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
    // ---- gather the lists of every (CTA, segment, column group) that saw this pixel's row: lane l owns lists l, l + 32, ...
    // (all counts are fetched at once, a warp scan places the lists, every lane copies its own -- the loads of different
    // lists are independent of one another)
    const int n_lists = (c_last - c_first + 1) * prm.n_sub;
    constexpr int kListsPerLane = 4;                       // <= 128 lists per pixel (the host keeps lists x k <= kTopkMaxCand)
    size_t recs[kListsPerLane];
    int cnts[kListsPerLane], offs[kListsPerLane];
    int C = 0;
#pragma unroll
    for (int q = 0; q < kListsPerLane; ++q) {
        const int li = lane + 32 * q;
        cnts[q] = 0;
        recs[q] = 0;
        if (li < n_lists) {
            const int c = c_first + li / prm.n_sub, sub = li % prm.n_sub;
            const int seg = mt - static_cast<int>(vosd::cta_begin(dec, c) / dec.nt);
            recs[q] = (static_cast<size_t>(c * dec.max_segs + seg) * prm.n_sub + sub) * kTile + row;
            cnts[q] = fp.cand_cnt[recs[q]];
        }
    }
#pragma unroll
    for (int q = 0; q < kListsPerLane; ++q) {
        int incl = cnts[q];
#pragma unroll
        for (int off = 1; off < 32; off <<= 1) {
            const int o = __shfl_up_sync(full, incl, off);
            if (lane >= off) incl += o;
        }
        offs[q] = C + incl - cnts[q];
        C += __shfl_sync(full, incl, 31);
    }
#pragma unroll
    for (int q = 0; q < kListsPerLane; ++q) {
        for (int e = 0; e < cnts[q]; ++e) {
            keys[offs[q] + e] = fp.cand_key[topk_slot(recs[q], e)];
            idxs[offs[q] + e] = static_cast<uint32_t>(fp.cand_idx[topk_slot(recs[q], e)]);
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
