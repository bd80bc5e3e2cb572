// Bandwidth kernels around the fused affinity kernel: merge + write-back, ring append, label records, decomposition
// tables, input normalisation.  Included by vos_prop.cu only (non-template kernels: one definition per library).
#pragma once
#include "kernels.cuh"

namespace vosk {

// -------------------------------------------------------------------------------------------
// Merge + write-back (HBM-bound): combines the per-segment online-softmax partials of each target
// pixel, normalises, writes prediction (predict()'s return value), arg-maxes over classes (first
// maximum wins, like torch.argmax), updates the ring's label record of the target frame
// (one-hot, or the raw prediction in probability mode -- inference_utils.py:67-71) and writes the
// stride-8 and the nearest-upsampled full-resolution uint8 masks (inference_utils.py:74-75).
// kMergeSplit blocks per low-resolution row (column halves), each emitting the full-resolution pixels that sample its columns.
// -------------------------------------------------------------------------------------------


// Decomposition tables for vos_merge_writeback (layout: MergeParams::tables); 64-bit divisions happen here, once per
// (video, reference count), instead of per pixel in the merge.
__global__ void vos_decomp_tables(int32_t* __restrict__ tab, int n_pixels, int n_refs, int num_sms) {
    const vosd::Decomp dec = vosd::make_decomp(n_pixels, n_refs, num_sms);
    for (int mt = threadIdx.x; mt < dec.tpf; mt += blockDim.x) {
        tab[mt] = vosd::cta_of(dec, static_cast<int64_t>(mt) * dec.nt);
        tab[dec.tpf + mt] = vosd::cta_of(dec, static_cast<int64_t>(mt) * dec.nt + dec.nt - 1);
    }
    for (int c = threadIdx.x; c < dec.grid; c += blockDim.x)
        tab[2 * dec.tpf + c] = static_cast<int32_t>(vosd::cta_begin(dec, c) / dec.nt);
}

constexpr int kMergeThreads = 512;    // 64 pixel groups of 8 lanes
constexpr int kMergeSplit = 2;        // blocks per low-resolution row (column halves): 120 blocks at 480p instead of 60 on 148 SMs
constexpr int kMergeLanes = 8;        // lanes cooperating on one target pixel

constexpr int kMergeSmall = 6;        // class capacity of the small instantiation (DAVIS: at most 5 objects + background)

template <int kCap>   // class capacity of this instantiation: kMergeSmall, kMetaClasses or kMaxClasses
__global__ void __launch_bounds__(kMergeThreads) vos_merge_writeback(const MergeParams prm) {
    extern __shared__ uint8_t row_cls[];  // [w_lowres]
    pdl_launch_dependents();
    const int y = blockIdx.x;
    const int x_lo = (prm.w_lowres * blockIdx.y) / kMergeSplit, x_hi = (prm.w_lowres * (blockIdx.y + 1)) / kMergeSplit;   // this block's columns
    const int sublane = threadIdx.x & (kMergeLanes - 1);
    const unsigned gmask = 0xffu << ((threadIdx.x & 31) & ~(kMergeLanes - 1));   // the 8 lanes of this pixel group
    const int32_t* mt0 = prm.tables + 2 * prm.tpf;
    const int sub_shift = prm.n_sub == 4 ? 2 : 1;        // n_sub is 2 (general kernel) or 4 (index-label kernel)
    constexpr int kMaxRec = 4;                           // records a lane folds without looping (8 lanes x 4 = 32 per pixel)
    bool waited = false;
    if (prm.tables_fresh) {
        // vos_decomp_tables is our immediate predecessor on this step and never triggers its dependents early: only
        // griddepcontrol.wait guarantees that its writes are visible
        pdl_wait();
        waited = true;
    }
    for (int x = x_lo + threadIdx.x / kMergeLanes; x < x_hi; x += kMergeThreads / kMergeLanes) {
        const int pix = y * prm.w_lowres + x;
        const int mt = pix / kTile, row = pix % kTile;
        // Everything up to here depends only on the decomposition tables.  On all but the first step of a reference count
        // they were written by a kernel that completed before the affinity kernel started, so this runs while the affinity
        // kernel is still executing (tables_fresh covers the first step); the partials are touched after pdl_wait().
        const int c_first = prm.tables[mt], c_last = prm.tables[prm.tpf + mt];
        const int n_rec = (c_last - c_first + 1) << sub_shift;
        const float* recs[kMaxRec];
#pragma unroll
        for (int j = 0; j < kMaxRec; ++j) {
            const int i = sublane + j * kMergeLanes;
            recs[j] = nullptr;
            if (i < n_rec) {
                const int c = c_first + (i >> sub_shift), h = i & (prm.n_sub - 1);
                recs[j] = prm.partials + (static_cast<size_t>(c * prm.max_segs + (mt - mt0[c])) * prm.n_sub + h) * kPartFloats + row;
            }
        }
        if (!waited) {
            pdl_wait();                   // the affinity kernel's partials (and everything before it) are complete
            waited = true;
        }
        // All of this lane's records are fetched in ONE batch of independent loads (absent records read the buffer's first
        // record and are zeroed afterwards), then folded in the order j = 0..3: a single L2 round trip on the chain
        // affinity -> merge -> affinity.  (Loads under a per-record branch were issued one record after the other: four
        // dependent round trips.)
        float M = kNegBig, L = 0.f, acc[kCap];
        float m_r[kMaxRec], l_r[kMaxRec], a_r[kMaxRec][kCap];
#pragma unroll
        for (int j = 0; j < kMaxRec; ++j) {
            const float* rp = recs[j] ? recs[j] : prm.partials;
            m_r[j] = __ldcg(rp);
            l_r[j] = __ldcg(rp + kTile);
#pragma unroll
            for (int k = 0; k < kCap; ++k) a_r[j][k] = (k < prm.d) ? __ldcg(rp + (2 + k) * kTile) : 0.f;
        }
#pragma unroll
        for (int j = 0; j < kMaxRec; ++j) {
            if (!recs[j]) {
                m_r[j] = kNegBig;
                l_r[j] = 0.f;
#pragma unroll
                for (int k = 0; k < kCap; ++k) a_r[j][k] = 0.f;
            }
            M = fmaxf(M, m_r[j]);
        }
#pragma unroll
        for (int k = 0; k < kCap; ++k) acc[k] = 0.f;
#pragma unroll
        for (int j = 0; j < kMaxRec; ++j) {
            const float w = vosptx::ex2(m_r[j] - M);       // absent record: its sums are exact zeros, L and acc stay as they are
            L = fmaf(l_r[j], w, L);
#pragma unroll
            for (int k = 0; k < kCap; ++k) acc[k] = fmaf(a_r[j][k], w, acc[k]);
        }
        for (int i = sublane + kMaxRec * kMergeLanes; i < n_rec; i += kMergeLanes) {      // (very long segment lists only)
            const int c = c_first + (i >> sub_shift), h = i & (prm.n_sub - 1);
            const float* rec = prm.partials + (static_cast<size_t>(c * prm.max_segs + (mt - mt0[c])) * prm.n_sub + h) * kPartFloats + row;
            const float mr = rec[0];
            const float M_new = fmaxf(M, mr);
            const float w_old = vosptx::ex2(M - M_new), w_new = vosptx::ex2(mr - M_new);
            L = fmaf(rec[kTile], w_new, L * w_old);
#pragma unroll
            for (int k = 0; k < kCap; ++k)
                if (k < prm.d) acc[k] = fmaf(rec[(2 + k) * kTile], w_new, acc[k] * w_old);
            M = M_new;
        }
#pragma unroll
        for (int off = 1; off < kMergeLanes; off <<= 1) {
            const float M_o = __shfl_xor_sync(gmask, M, off);
            const float M_new = fmaxf(M, M_o);
            const float w_a = vosptx::ex2(M - M_new), w_b = vosptx::ex2(M_o - M_new);
            L = fmaf(__shfl_xor_sync(gmask, L, off), w_b, L * w_a);
#pragma unroll
            for (int k = 0; k < kCap; ++k) acc[k] = fmaf(__shfl_xor_sync(gmask, acc[k], off), w_b, acc[k] * w_a);   // rows >= d: zeros
            M = M_new;
        }
        if (sublane != 0) continue;
        const float inv = 1.0f / L;
        int best = 0;
        float best_v = -INFINITY;
        float* mrec = prm.meta + (static_cast<size_t>(prm.q_slot) * prm.p_pad + pix) * kMetaFloats + 2;
#pragma unroll
        for (int k = 0; k < kCap; ++k) {
            if (k < prm.d) {
                const float pk = acc[k] * inv;
                acc[k] = pk;
                if (pk > best_v) { best_v = pk; best = k; }   // strict '>' : first maximum wins
                if (prm.out_prediction) prm.out_prediction[static_cast<size_t>(k) * prm.n_pixels + pix] = pk;
            }
        }
        if (prm.write_labels) {
            if constexpr (kCap <= kMetaClasses) {   // wider class sets live in the class bytes only
                float rec_v[kMetaClasses];          // the whole record, also the classes beyond this instantiation's capacity
#pragma unroll
                for (int k = 0; k < kMetaClasses; ++k)
                    rec_v[k] = (k < kCap && k < prm.d) ? (prm.probability ? acc[k < kCap ? k : 0] : (k == best ? 1.f : 0.f)) : 0.f;
#pragma unroll
                for (int k = 0; k < kMetaClasses; k += 2)          // the record's class part starts 8 bytes into a 64-byte record
                    *reinterpret_cast<float2*>(mrec + k) = make_float2(rec_v[k], rec_v[k + 1]);
            }
            prm.cls[static_cast<size_t>(prm.q_slot) * prm.p_pad + pix] = static_cast<uint8_t>(best);
        }
        if (prm.out_mask_lowres) prm.out_mask_lowres[pix] = static_cast<uint8_t>(best);
        row_cls[x] = static_cast<uint8_t>(best);
    }
    if (!waited) pdl_wait();
    if (!prm.out_mask_fullres) return;
    __syncthreads();
    const float sy = static_cast<float>(prm.h_lowres) / static_cast<float>(prm.H);
    const float sx = static_cast<float>(prm.w_lowres) / static_cast<float>(prm.W);
    // full-res rows that sample low-res row y: a window around y/sy, filtered by the exact rule
    const int dy0 = max(0, static_cast<int>(static_cast<float>(y) / sy) - 1);
    const int dy1 = min(prm.H, static_cast<int>(static_cast<float>(y + 1) / sy) + 2);
    for (int dx = threadIdx.x; dx < prm.W; dx += kMergeThreads) {
        const int xs = nearest_src(dx, sx, prm.w_lowres);
        if (xs < x_lo || xs >= x_hi) continue;                 // the other column half's block writes this pixel
        const uint8_t c = row_cls[xs];
        for (int dy = dy0; dy < dy1; ++dy)
            if (nearest_src(dy, sy, prm.h_lowres) == y) prm.out_mask_fullres[static_cast<size_t>(dy) * prm.W + dx] = c;
    }
}

// -------------------------------------------------------------------------------------------
// Ring append (HBM-bound): one frame's embedding -> bf16 hi/lo, pixel-major (P_pad, 256).
// Replaces torch.cat of feats_history (inference_utils.py:72) + the permute/reshape copy of
// predict.py:47.  Source may be fp32 / fp16 / bf16, channel-major (NCHW) or pixel-major (NHWC).
// -------------------------------------------------------------------------------------------
template <typename T> __device__ __forceinline__ float to_f32(T v);
template <> __device__ __forceinline__ float to_f32<float>(float v) { return v; }
template <> __device__ __forceinline__ float to_f32<__half>(__half v) { return __half2float(v); }
template <> __device__ __forceinline__ float to_f32<__nv_bfloat16>(__nv_bfloat16 v) { return __bfloat162float(v); }

// Stores one channel pair in the ring's format: bf16 hi + lo (kFmtSplit), or a single fp16 / bf16 value
// (exact when the source already has that type -- the host refuses anything else in those modes).
__device__ __forceinline__ uint32_t pack_hi(float x0, float x1, int fmt, uint32_t& lo_bits) {
    if (fmt == kFmtF16) {
        const __half2 h = __floats2half2_rn(x0, x1);
        return reinterpret_cast<const uint32_t&>(h);
    }
    const __nv_bfloat16 h0 = __float2bfloat16_rn(x0), h1 = __float2bfloat16_rn(x1);
    const __nv_bfloat162 h = __halves2bfloat162(h0, h1);
    if (fmt == kFmtSplit) {
        const __nv_bfloat162 l = __halves2bfloat162(__float2bfloat16_rn(x0 - __bfloat162float(h0)),
                                                    __float2bfloat16_rn(x1 - __bfloat162float(h1)));
        lo_bits = reinterpret_cast<const uint32_t&>(l);
    }
    return reinterpret_cast<const uint32_t&>(h);
}

// A launch appends a batch of consecutive frames (grid z / y = frame of the batch): frame first_frame + i goes to ring slot
// (first_frame + i) % ring_slots.  One launch per batch instead of one per frame takes the append off the per-frame chain
// affinity -> merge -> affinity (a 3.3 MB copy alone is launch-latency-bound: 6-8 us at 0.8 TB/s).
struct AppendSlots {
    int32_t first_frame, ring_slots, p_pad;
    __device__ __forceinline__ size_t row0(unsigned i) const {
        return static_cast<size_t>((first_frame + static_cast<int>(i)) % ring_slots) * p_pad;
    }
};

// Channel-major source (torch default): 32 pixels x 64 channels per block, transposed through shared memory.
template <typename T>
__global__ void __launch_bounds__(256) vos_append_nchw(const T* __restrict__ src, __nv_bfloat16* __restrict__ hi,
                                                       __nv_bfloat16* __restrict__ lo, int n_pixels, AppendSlots sl, int fmt) {
    __shared__ float tile[64][33];
    pdl_launch_dependents();
    pdl_wait();                           // the ring slot being overwritten is no longer read by earlier kernels
    const size_t slot_row0 = sl.row0(blockIdx.z);
    src += static_cast<size_t>(blockIdx.z) * kK * n_pixels;      // frame blockIdx.z of the batch
    const int p0 = blockIdx.x * 32, c0 = blockIdx.y * 64;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int p = p0 + lane;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int c = w * 8 + i;
        tile[c][lane] = p < n_pixels ? to_f32<T>(src[static_cast<size_t>(c0 + c) * n_pixels + p]) : 0.f;
    }
    __syncthreads();
    uint32_t* hi32 = reinterpret_cast<uint32_t*>(hi);
    uint32_t* lo32 = reinterpret_cast<uint32_t*>(lo);
#pragma unroll
    for (int i = threadIdx.x; i < 32 * 32; i += 256) {
        const int pix = i >> 5, cp = i & 31;
        if (p0 + pix < n_pixels) {
            uint32_t lo_bits = 0;
            const size_t off = ((slot_row0 + p0 + pix) * kK + c0) / 2 + cp;
            hi32[off] = pack_hi(tile[2 * cp][pix], tile[2 * cp + 1][pix], fmt, lo_bits);
            if (fmt == kFmtSplit) lo32[off] = lo_bits;
        }
    }
}

// Pixel-major source (channels_last): elementwise convert, 8 channels per thread (16-byte stores).
// Requires a 16-byte aligned source (the host falls back to the pair kernel otherwise).
template <typename T>
__global__ void __launch_bounds__(256) vos_append_nhwc8(const T* __restrict__ src, __nv_bfloat16* __restrict__ hi,
                                                        __nv_bfloat16* __restrict__ lo, int n_pixels, AppendSlots sl, int fmt) {
    pdl_launch_dependents();
    pdl_wait();
    const size_t i = static_cast<size_t>(blockIdx.x) * 256 + threadIdx.x;  // group of 8 channels
    if (i >= static_cast<size_t>(n_pixels) * (kK / 8)) return;
    const size_t slot_row0 = sl.row0(blockIdx.y);
    src += static_cast<size_t>(blockIdx.y) * kK * n_pixels;      // frame blockIdx.y of the batch
    float x[8];
    if (sizeof(T) == 4) {
        const float4 a = reinterpret_cast<const float4*>(src)[2 * i], b = reinterpret_cast<const float4*>(src)[2 * i + 1];
        x[0] = a.x; x[1] = a.y; x[2] = a.z; x[3] = a.w; x[4] = b.x; x[5] = b.y; x[6] = b.z; x[7] = b.w;
    } else {
        const uint4 raw = reinterpret_cast<const uint4*>(src)[i];
        const T* e = reinterpret_cast<const T*>(&raw);
#pragma unroll
        for (int k = 0; k < 8; ++k) x[k] = to_f32<T>(e[k]);
    }
    uint4 h, l = make_uint4(0, 0, 0, 0);
    h.x = pack_hi(x[0], x[1], fmt, l.x);
    h.y = pack_hi(x[2], x[3], fmt, l.y);
    h.z = pack_hi(x[4], x[5], fmt, l.z);
    h.w = pack_hi(x[6], x[7], fmt, l.w);
    const size_t off = slot_row0 * (kK / 8) + i;
    reinterpret_cast<uint4*>(hi)[off] = h;
    if (fmt == kFmtSplit) reinterpret_cast<uint4*>(lo)[off] = l;
}

template <typename T>
__global__ void __launch_bounds__(256) vos_append_nhwc(const T* __restrict__ src, __nv_bfloat16* __restrict__ hi,
                                                       __nv_bfloat16* __restrict__ lo, int n_pixels, AppendSlots sl, int fmt) {
    pdl_launch_dependents();
    pdl_wait();
    const size_t i = static_cast<size_t>(blockIdx.x) * 256 + threadIdx.x;  // channel pair index
    if (i >= static_cast<size_t>(n_pixels) * (kK / 2)) return;
    const size_t slot_row0 = sl.row0(blockIdx.y);
    src += static_cast<size_t>(blockIdx.y) * kK * n_pixels;
    uint32_t lo_bits = 0;
    const size_t off = slot_row0 * (kK / 2) + i;
    reinterpret_cast<uint32_t*>(hi)[off] = pack_hi(to_f32<T>(src[2 * i]), to_f32<T>(src[2 * i + 1]), fmt, lo_bits);
    if (fmt == kFmtSplit) reinterpret_cast<uint32_t*>(lo)[off] = lo_bits;
}

// -------------------------------------------------------------------------------------------
// Meta records: {rowf, xf} are geometry (same for every slot), V[14] are the labels.
// -------------------------------------------------------------------------------------------
__global__ void vos_init_meta(float* __restrict__ meta, int slots, int p_pad, int n_pixels, int w_lowres) {
    const size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= static_cast<size_t>(slots) * p_pad) return;
    const int pix = static_cast<int>(i % p_pad);
    float rowf = 0.f, xf = 0.f;
    if (pix < n_pixels) pixel_coord(pix, w_lowres, rowf, xf);
    float4* rec = reinterpret_cast<float4*>(meta + i * kMetaFloats);
    rec[0] = make_float4(rowf, xf, 0.f, 0.f);
    rec[1] = rec[2] = rec[3] = make_float4(0.f, 0.f, 0.f, 0.f);
}

__global__ void vos_set_labels_index(float* __restrict__ meta_slot, uint8_t* __restrict__ cls_slot,
                                     const uint8_t* __restrict__ cls, int n_pixels, int d) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n_pixels) return;
    const int c = cls[p];
    cls_slot[p] = static_cast<uint8_t>(c);
    if (d > kMetaClasses) return;   // more classes than a meta record holds: class bytes only
    float* rec = meta_slot + static_cast<size_t>(p) * kMetaFloats + 2;
#pragma unroll
    for (int k = 0; k < kMetaClasses; ++k) rec[k] = (k == c) ? 1.f : 0.f;
}

__global__ void vos_set_labels_dense(float* __restrict__ meta_slot, const float* __restrict__ labels, int n_pixels, int d) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n_pixels) return;
    float* rec = meta_slot + static_cast<size_t>(p) * kMetaFloats + 2;
#pragma unroll
    for (int k = 0; k < kMetaClasses; ++k) rec[k] = (k < d) ? labels[static_cast<size_t>(k) * n_pixels + p] : 0.f;
}

// -------------------------------------------------------------------------------------------
// Input normalisation (HBM-bound, caller side of the path): uint8 RGB pixel-interleaved -> (x/255 - mean)/std, same
// element order (= channels-last NCHW).  fp32 arithmetic in the reference's order (datasets.py:128-131: ToTensor divides
// by 255, Normalize subtracts the mean, then divides by std); each thread converts 4 pixels (12 bytes in, 3 x 128-bit or
// 3 x 64-bit out).
// -------------------------------------------------------------------------------------------
template <typename T>
__global__ void vos_normalize_u8(const uint8_t* __restrict__ rgb, T* __restrict__ out, int64_t n_pixels,
                                 float m0, float m1, float m2, float s0, float s1, float s2) {
    const int64_t quad = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    const int64_t p0 = quad * 4;
    if (p0 >= n_pixels) return;
    const float mean[3] = {m0, m1, m2}, sd[3] = {s0, s1, s2};
    if (p0 + 4 <= n_pixels) {
        const uint32_t* src = reinterpret_cast<const uint32_t*>(rgb + p0 * 3);      // p0 * 3 is a multiple of 12
        const uint32_t w[3] = {__ldg(src), __ldg(src + 1), __ldg(src + 2)};
        T v[12];
#pragma unroll
        for (int i = 0; i < 12; ++i) {
            const float x = static_cast<float>((w[i >> 2] >> (8 * (i & 3))) & 0xffu);
            v[i] = static_cast<T>(__fdiv_rn(__fdiv_rn(x, 255.f) - mean[i % 3], sd[i % 3]));
        }
        T* dst = out + p0 * 3;
#pragma unroll
        for (int i = 0; i < 12; ++i) dst[i] = v[i];       // 48 / 24 contiguous bytes per thread: the compiler vectorises
    } else {
        for (int64_t i = p0 * 3; i < n_pixels * 3; ++i)
            out[i] = static_cast<T>(__fdiv_rn(__fdiv_rn(static_cast<float>(rgb[i]), 255.f) - mean[i % 3], sd[i % 3]));
    }
}

}  // namespace vosk
