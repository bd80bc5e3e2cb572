// libvosprop: C ABI (include/vos_prop.h) over the kernels in kernels.cuh.
// Host side only: engine state (ring buffer, TMA descriptors, scratch), argument checking,
// launch plumbing.  No CPU fallback anywhere: without an sm_100 device vosprop_create() fails.
#include "../../include/vos_prop.h"

#include <cuda.h>
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <new>
#include <utility>
#include <vector>

#include "decompose.h"
#include "kernels.cuh"
#include "side_kernels.cuh"
#include "launch.h"
#include "topk_params.h"

static_assert(VOSPROP_PREC_SPLIT3 == vosk::kFmtSplit && VOSPROP_PREC_F16 == vosk::kFmtF16 && VOSPROP_PREC_BF16 == vosk::kFmtBF16,
              "precision enum and ring formats must coincide");

namespace {

thread_local char g_err[512] = "";

int fail(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}

#define VOS_CUDA(expr)                                                                              \
    do {                                                                                            \
        cudaError_t _e = (expr);                                                                    \
        if (_e != cudaSuccess)                                                                      \
            return fail(VOSPROP_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
    } while (0)

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

}  // namespace

struct vosprop_engine {
    vosprop_config cfg{};
    int num_sms = 0;
    int p_pad_cap = 0;      // ring rows per slot at capacity
    // geometry of the current video
    int H_d = 0, W_d = 0, H = 0, W = 0, d = 0, P = 0, p_pad = 0;
    int precision = VOSPROP_PREC_SPLIT3;
    int dbg = 0;
    long long* dbg_clk = nullptr;
    __nv_bfloat16* ring_hi = nullptr;
    __nv_bfloat16* ring_lo = nullptr;
    float* meta = nullptr;
    uint8_t* cls = nullptr;   // class-id ring, [slots * p_pad]
    float* partials = nullptr;
    size_t partial_records = 0;
    // top-k mode scratch (allocated on first use)
    uint32_t* cand_key = nullptr;
    int32_t* cand_idx = nullptr;
    int32_t* cand_cnt = nullptr;
    uint8_t* low_scratch = nullptr;   // stride-8 class map when the caller wants only the full-resolution mask
    float* topk_bound = nullptr;      // block maxima of the first scan (grows on demand)
    size_t topk_bound_floats = 0;
    float* topk_tau = nullptr;        // [p_pad_cap] per-pixel threshold of the second scan
    // decomposition tables of the merge kernel, one per reference count (cached until the next reset)
    int32_t* tables = nullptr;        // device [VOSPROP_MAX_REFS + 1][table_stride]
    int block_skip = 2;               // vosprop_block_skip: 0 never, 1 always, 2 auto (the default)
    // auto mode: a launch of the skipping kernel reports how many blocks it left out into host-mapped memory; the host looks
    // at the newest report before each launch (no synchronisation: reports arrive some launches late)
    int32_t* skip_scratch = nullptr;          // device {sum, ticket}
    volatile int32_t* skip_report = nullptr;  // host-mapped [kSkipSlots][2] = {tag, dead blocks}
    int32_t* skip_report_dev = nullptr;
    int last_dense_kernel = -1;       // 0 vos_affinity_tc, 1 simt checker, 2 vos_affinity_prob
    int skip_on = 0;                  // auto: current decision
    int64_t skip_launch = 0;          // launches of the fused index kernel so far
    int64_t skip_probe_until = 0;     // launches < this run the skipping kernel to probe
    int skip_next_tag = 1;
    int64_t skip_blocks[8] = {0};     // blocks of the launch that carried tag t (slot t % 8)
    size_t table_stride = 0;
    std::vector<char> table_valid;
    CUtensorMap tmap_hi{}, tmap_lo{};
    EncodeTiledFn encode = nullptr;
    std::vector<int> slot_frame;
    std::vector<char> slot_labels;   // 0: none, 1: index labels (cls ring + one-hot record), 2: dense record only
    int64_t launches = 0;
    // optional per-kernel timing (bench.py roofline)
    std::vector<cudaEvent_t> ev;      // 2 events per timed launch
    std::vector<int> ev_kind;
    size_t ev_used = 0;
    bool timing = false;
    int timing_mask = 7;
};

namespace {
using vosk::launch_pdl;

struct TimedLaunch {   // RAII: records an event pair around one kernel launch when timing is on
    vosprop_engine* e; cudaStream_t st; bool on;
    TimedLaunch(vosprop_engine* e_, int kind, cudaStream_t st_) : e(e_), st(st_), on(false) {
        if (e->timing && ((e->timing_mask >> kind) & 1) && e->ev_used + 2 <= e->ev.size()) {
            on = true;
            e->ev_kind[e->ev_used / 2] = kind;
            cudaEventRecord(e->ev[e->ev_used], st);
        }
    }
    ~TimedLaunch() {
        if (on) { cudaEventRecord(e->ev[e->ev_used + 1], st); e->ev_used += 2; }
    }
};
}  // namespace

namespace {

int encode_maps(vosprop_engine* e) {
    const cuuint64_t dims[2] = {static_cast<cuuint64_t>(vosk::kK),
                                static_cast<cuuint64_t>(e->cfg.ring_slots) * static_cast<cuuint64_t>(e->p_pad)};
    const cuuint64_t strides[1] = {static_cast<cuuint64_t>(vosk::kK) * 2};
    const cuuint32_t box[2] = {static_cast<cuuint32_t>(vosk::kKC), static_cast<cuuint32_t>(vosk::kTile)};
    const cuuint32_t estr[2] = {1, 1};
    CUresult r1 = e->encode(&e->tmap_hi, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, e->ring_hi, dims, strides, box, estr,
                            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                            CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    CUresult r2 = e->encode(&e->tmap_lo, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, e->ring_lo, dims, strides, box, estr,
                            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                            CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r1 != CUDA_SUCCESS || r2 != CUDA_SUCCESS)
        return fail(VOSPROP_ERR_CUDA, "cuTensorMapEncodeTiled failed (%d, %d)", (int)r1, (int)r2);
    return VOSPROP_OK;
}

int dispatch_affinity(vosprop_engine* e, const vosk::AffinityParams& prm, int grid, int kernel, cudaStream_t st) {
    // smallest instantiated class capacity >= d
    static const int caps[] = {2, 3, 4, 6, 8, 11, 14, vosk::kMaxClasses};
    int D = vosk::kMaxClasses;
    for (int c : caps)
        if (e->d <= c) { D = c; break; }
    cudaError_t ce;
    if (kernel == VOSPROP_KERNEL_TC) {
        // same bound as vosk::prior_const: is the prior recurrence safe for every reference of this step?
        bool wide = false;
        const float inv_w = prm.inv_w, w = static_cast<float>(prm.w_lowres), h = static_cast<float>((prm.n_pixels + prm.w_lowres - 1) / prm.w_lowres);
        for (int r = 0; r < prm.n_refs; ++r) {
            const float coef = prm.ref_coef[r], gamma = -coef * (inv_w * inv_w + 1.0f);
            wide = wide || !(30.f * coef * (h * inv_w + w + 16.f) - 225.f * gamma < 100.f);
        }
        if (D > vosk::kMetaClasses) wide = true;      // 15..24 classes: only the per-tile-tested form is instantiated
        const bool split = prm.feat_fmt == vosk::kFmtSplit;
        vosk::AffinityParams prm_k = prm;
        prm_k.tile_stride = 1;
        bool skip = e->block_skip == 1 && !wide;
        if (e->block_skip == 2 && !wide && e->skip_report) {
            // Auto mode.  The skipping kernel is 6-9 % slower than the plain one where nothing can be skipped (probing, strided
            // tile order) and 6-17 % faster on embeddings as peaked as a trained network's.  Policy: the first launches of an
            // engine and two of every 256 afterwards run it as a probe; a report of >= 5 % skipped blocks switches it on, one
            // of < 2 % off.  Reports are read as they arrive (the host may be many launches ahead of the device).
            constexpr int kSlots = 8;
            for (int sl = 0; sl < kSlots; ++sl) {
                const int tag = e->skip_report[2 * sl];
                if (tag > 0 && e->skip_blocks[sl] > 0 && tag % kSlots == sl) {
                    const double frac = static_cast<double>(e->skip_report[2 * sl + 1]) / static_cast<double>(e->skip_blocks[sl]);
                    if (frac >= 0.05) e->skip_on = 1;
                    else if (frac < 0.02) e->skip_on = 0;
                    e->skip_report[2 * sl] = 0;              // consumed
                    e->skip_blocks[sl] = 0;
                }
            }
            if (e->skip_launch % 256 == 0) e->skip_probe_until = e->skip_launch + (e->skip_launch == 0 ? 4 : 2);
            skip = e->skip_on || e->skip_launch < e->skip_probe_until;
            ++e->skip_launch;
            if (skip) {
                const int tag = e->skip_next_tag;
                e->skip_next_tag = tag == (1 << 30) ? 1 : tag + 1;
                const int sl = tag % kSlots;
                const int64_t tiles = static_cast<int64_t>((prm.n_pixels + vosk::kTile - 1) / vosk::kTile);
                e->skip_blocks[sl] = tiles * tiles * prm.n_refs * vosk::kIdxEpiWarpsHost;
                prm_k.skip_scratch = e->skip_scratch;
                prm_k.skip_report = reinterpret_cast<volatile int32_t*>(e->skip_report_dev) + 2 * sl;
                prm_k.skip_tag = tag;
            }
        }
        // explicit mode 1 only: golden-section stride, coprime with the tiles per frame, so that the live (near-diagonal) tiles
        // spread evenly over the CTAs' ranges (-10..-17 % instead of -6..-9 % on peaked embeddings).  Auto mode keeps the
        // natural tile order: then the skipping kernel adds the same numbers in the same order as the plain one minus exact
        // zeros, and a switch between the two is invisible in the results (sharding determinism, tests/test_gpu_shard.py).
        if (skip && e->block_skip == 1) {
            const int tpf = (prm.n_pixels + vosk::kTile - 1) / vosk::kTile;
            int s = std::max(1, static_cast<int>(tpf * 0.381966f + 0.5f));
            auto gcd = [](int a, int b) { while (b) { const int t = a % b; a = b; b = t; } return a; };
            while (s < tpf && gcd(s, tpf) != 1) ++s;
            prm_k.tile_stride = s < tpf ? s : 1;
        }
        ce = vosk::launch_affinity_idx(D, split, wide, skip, grid, st, e->tmap_hi, e->tmap_lo, prm_k);
    } else {
        // dense label records: without any prior (probability propagation, predict.py:58) the TMEM-A kernel with 16 epilogue
        // warps; with a prior (float label histories of the stateless predict() adapter) or on request the round-1 kernel
        bool no_prior = true;
        for (int r = 0; r < prm.n_refs; ++r) no_prior = no_prior && prm.ref_coef[r] == 0.f;
        const int which = kernel == VOSPROP_KERNEL_SIMT ? 1 : (no_prior ? 2 : 0);
        e->last_dense_kernel = which;
        ce = vosk::launch_affinity_dense(D, which, grid, st, e->tmap_hi, e->tmap_lo, prm);
    }
    if (ce != cudaSuccess) return fail(VOSPROP_ERR_CUDA, "affinity kernel launch failed: %s", cudaGetErrorString(ce));
    VOS_CUDA(cudaGetLastError());
    return VOSPROP_OK;
}

int check_frame(const vosprop_engine* e, int frame_idx) {
    if (!e) return fail(VOSPROP_ERR_INVALID, "null engine");
    if (e->P == 0) return fail(VOSPROP_ERR_STATE, "vosprop_reset() has not been called");
    if (frame_idx < 0) return fail(VOSPROP_ERR_INVALID, "negative frame index %d", frame_idx);
    return VOSPROP_OK;
}


// Top-k extension (not in the reference): two scans of the affinity (block maxima -> per-row threshold -> candidate lists),
// then the per-pixel finish, then the up-sampling of the mask (affinity_topk.cuh).
int propagate_topk(vosprop_engine* e, const vosprop_step* s, vosk::AffinityParams ap, const vosd::Decomp& dec_in, cudaStream_t st) {
    // Lists merged per target pixel = 4 column groups x the CTAs whose ranges intersect one target tile's row of the tile
    // grid.  Small maps with many references would spread one row over dozens of CTAs: shrink the grid until the finish
    // kernel's merge buffer (kTopkMaxCand) holds k entries of every list.
    int grid_cap = e->num_sms;
    const int64_t max_lists = std::max<int64_t>(1, vosk::kTopkMaxCand / (static_cast<int64_t>(vosk::kIdxSub) * s->topk));
    {
        const int64_t min_per_cta = max_lists > 1 ? (dec_in.nt + max_lists - 2) / (max_lists - 1) : dec_in.total;
        if (dec_in.total / grid_cap < min_per_cta) grid_cap = static_cast<int>(std::max<int64_t>(1, dec_in.total / min_per_cta));
    }
    const vosd::Decomp dec = vosd::make_decomp(e->P, s->n_refs, grid_cap);
    const int64_t per_cta = dec.total / dec.grid;
    const int64_t lists = (dec.nt + per_cta - 1) / per_cta + 1;
    if (lists * vosk::kIdxSub * s->topk > vosk::kTopkMaxCand)
        return fail(VOSPROP_ERR_UNSUPPORTED, "top-k: %lld lists x k=%d candidates per target pixel (max %d)", (long long)(lists * vosk::kIdxSub),
                    s->topk, vosk::kTopkMaxCand);
    if (static_cast<size_t>(dec.grid) * dec.max_segs * vosk::kIdxSub > e->partial_records)
        return fail(VOSPROP_ERR_UNSUPPORTED, "top-k: list table too small (grid %d x segs %d)", dec.grid, dec.max_segs);
    ap.num_sms = grid_cap;
    if (!e->cand_key) {
        const size_t rows = e->partial_records * vosk::kTile;     // partial_records counts 4 column groups per (CTA, segment)
        cudaError_t a1 = cudaMalloc(&e->cand_key, rows * vosk::kTopkCap * 4);
        cudaError_t a2 = cudaMalloc(&e->cand_idx, rows * vosk::kTopkCap * 4);
        cudaError_t a3 = cudaMalloc(&e->cand_cnt, rows * 4);
        cudaError_t a4 = cudaMalloc(&e->low_scratch, static_cast<size_t>(e->cfg.max_pixels));
        cudaError_t a5 = cudaMalloc(&e->topk_tau, static_cast<size_t>(e->p_pad_cap) * 4);
        if (a1 != cudaSuccess || a2 != cudaSuccess || a3 != cudaSuccess || a4 != cudaSuccess || a5 != cudaSuccess)
            return fail(VOSPROP_ERR_CUDA, "cudaMalloc of the top-k candidate lists failed (%zu rows)", rows);
    }
    // block maxima of pass 1: 4 x 128 floats per 128 x 128 logit tile (480p, 9 references: 48 MB); grows on demand
    const size_t bound_floats = static_cast<size_t>(dec.total) * vosk::kIdxSub * vosk::kTile;
    if (bound_floats > e->topk_bound_floats) {
        if (e->topk_bound) {
            VOS_CUDA(cudaStreamSynchronize(st));          // a scan of an earlier step may still be writing the old buffer
            cudaFree(e->topk_bound);
            e->topk_bound = nullptr;
            e->topk_bound_floats = 0;
        }
        if (cudaMalloc(&e->topk_bound, bound_floats * 4) != cudaSuccess)
            return fail(VOSPROP_ERR_CUDA, "cudaMalloc of the top-k block maxima failed (%zu MB)", bound_floats * 4 >> 20);
        e->topk_bound_floats = bound_floats;
    }
    ap.temperature = s->temperature;
    ap.topk = s->topk;
    ap.cand_key = e->cand_key; ap.cand_idx = e->cand_idx; ap.cand_cnt = e->cand_cnt;
    ap.topk_bound = e->topk_bound; ap.topk_tau = e->topk_tau;
    {
        TimedLaunch timed(e, VOSPROP_T_AFFINITY, st);
        const bool split = ap.feat_fmt == vosk::kFmtSplit;
        // The first scan visits every tile_step-th reference tile of a row only: any subset of a row's logits gives a valid
        // lower bound of its k-th largest.  A third of the tiles costs a third of the scan and about triples the candidates
        // of the second one, which is cheap while the lists stay short (measured at 480p, 9 references, frames/s for
        // steps 1 / 2 / 3 / 4 / 6:  k = 5: 2583 / 3474 / 3926 / 4067 / 4148;  k = 20: 2550 / 3382 / 3788 / 3764 / 3230;
        // k = 50: 2468 / 3181 / 3384 / 2591 / 328 -- past the knee the lists overflow and every overflow is a prune).
        int tile_step = s->topk <= 8 ? 4 : (s->topk <= 24 ? 3 : 2);
        while (tile_step > 1 && (dec.nt / tile_step) * vosk::kIdxSub < 8 * s->topk) --tile_step;    // small maps: keep >= 8 k blocks per row
        vosk::AffinityParams ap1 = ap;
        ap1.tile_step = tile_step;
        VOS_CUDA(vosk::launch_topk_scan(split, 1, dec.grid, st, e->tmap_hi, e->tmap_lo, ap1));
        VOS_CUDA(vosk::launch_topk_threshold(e->topk_bound, e->topk_tau, e->P, dec.nt, tile_step, s->topk, st));
        VOS_CUDA(vosk::launch_topk_scan(split, 2, dec.grid, st, e->tmap_hi, e->tmap_lo, ap));
        VOS_CUDA(cudaGetLastError());
    }
    e->launches += 2;
    if (s->record_event) VOS_CUDA(cudaEventRecord(static_cast<cudaEvent_t>(s->record_event), st));
    vosk::TopkFinishParams fp{};
    vosk::MergeParams& mp = fp.mp;
    const int q_slot = ap.q_slot;
    mp.n_pixels = e->P; mp.p_pad = e->p_pad; mp.w_lowres = e->W_d; mp.h_lowres = e->H_d; mp.n_refs = s->n_refs;
    mp.num_sms = grid_cap; mp.d = e->d; mp.H = e->H; mp.W = e->W; mp.q_slot = q_slot;
    mp.write_labels = s->write_labels; mp.probability = s->probability_propagation; mp.n_sub = vosk::kIdxSub;
    mp.partials = nullptr; mp.meta = e->meta; mp.cls = e->cls;
    mp.out_prediction = s->out_prediction;
    mp.out_mask_lowres = s->out_mask_lowres ? s->out_mask_lowres : (s->out_mask_fullres ? e->low_scratch : nullptr);
    mp.out_mask_fullres = nullptr;
    fp.topk = s->topk;
    for (int r = 0; r < s->n_refs; ++r) { fp.ref_slot[r] = ap.ref_slot[r]; fp.ref_coef[r] = ap.ref_coef[r]; }
    fp.cand_key = e->cand_key; fp.cand_idx = e->cand_idx; fp.cand_cnt = e->cand_cnt;
    fp.out_topk_idx = s->out_topk_idx;
    {
        TimedLaunch timed(e, VOSPROP_T_MERGE, st);
        VOS_CUDA(vosk::launch_topk_finish(fp, st));
        if (s->out_mask_fullres) {
            VOS_CUDA(vosk::launch_upsample_mask(mp.out_mask_lowres, s->out_mask_fullres, e->H_d, e->W_d, e->H, e->W, st));
            e->launches++;
        }
    }
    if (s->write_labels) e->slot_labels[q_slot] = s->probability_propagation ? 2 : 1;
    e->launches += 2;
    return VOSPROP_OK;
}

}  // namespace

extern "C" {

const char* vosprop_last_error(void) { return g_err; }
int vosprop_abi_version(void) { return VOSPROP_ABI_VERSION; }

int vosprop_create(const vosprop_config* cfg, vosprop_engine** out) {
    if (!cfg || !out) return fail(VOSPROP_ERR_INVALID, "null argument");
    *out = nullptr;
    if (cfg->max_pixels <= 0 || cfg->ring_slots < 2 || cfg->max_fullres_pixels < 0)
        return fail(VOSPROP_ERR_INVALID, "bad config: max_pixels=%d ring_slots=%d", cfg->max_pixels, cfg->ring_slots);
    int n_dev = 0;
    cudaError_t ce = cudaGetDeviceCount(&n_dev);
    if (ce != cudaSuccess || n_dev == 0)
        return fail(VOSPROP_ERR_UNSUPPORTED, "no CUDA device (%s); libvosprop has no CPU fallback",
                    ce == cudaSuccess ? "count = 0" : cudaGetErrorString(ce));
    if (cfg->device < 0 || cfg->device >= n_dev) return fail(VOSPROP_ERR_INVALID, "device %d out of range", cfg->device);
    cudaDeviceProp prop;
    VOS_CUDA(cudaGetDeviceProperties(&prop, cfg->device));
    if (prop.major != 10)
        return fail(VOSPROP_ERR_UNSUPPORTED, "device %d is sm_%d%d; this library is built for sm_100a only", cfg->device,
                    prop.major, prop.minor);
    VOS_CUDA(cudaSetDevice(cfg->device));
    vosprop_engine* e = new (std::nothrow) vosprop_engine();
    if (!e) return fail(VOSPROP_ERR_INVALID, "out of host memory");
    e->cfg = *cfg;
    e->num_sms = prop.multiProcessorCount;
    e->p_pad_cap = (cfg->max_pixels + vosk::kTile - 1) / vosk::kTile * vosk::kTile;
    e->slot_frame.assign(cfg->ring_slots, -1);
    e->slot_labels.assign(cfg->ring_slots, 0);

    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    ce = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres);
    if (ce != cudaSuccess || qres != cudaDriverEntryPointSuccess || !fn) {
        delete e;
        return fail(VOSPROP_ERR_CUDA, "cuTensorMapEncodeTiled not available from the driver");
    }
    e->encode = reinterpret_cast<EncodeTiledFn>(fn);

    const size_t rows = static_cast<size_t>(cfg->ring_slots) * e->p_pad_cap;
    const int tiles_cap = e->p_pad_cap / vosk::kTile;
    e->partial_records = static_cast<size_t>(e->num_sms) * ((tiles_cap + e->num_sms - 1) / e->num_sms + 2) * vosk::kIdxSub;
    cudaError_t a1 = cudaMalloc(&e->ring_hi, rows * vosk::kK * 2);
    cudaError_t a2 = cudaMalloc(&e->ring_lo, rows * vosk::kK * 2);
    cudaError_t a3 = cudaMalloc(&e->meta, rows * vosk::kMetaFloats * 4);
    cudaError_t a4 = cudaMalloc(&e->partials, e->partial_records * vosk::kPartFloats * 4);
    cudaError_t a5 = cudaMalloc(&e->cls, rows);
    e->table_stride = static_cast<size_t>(2 * tiles_cap + e->num_sms + 8);
    e->table_valid.assign(VOSPROP_MAX_REFS + 1, 0);
    cudaError_t a6 = cudaMalloc(&e->tables, (VOSPROP_MAX_REFS + 1) * e->table_stride * sizeof(int32_t));
    if (a6 != cudaSuccess) a1 = a6;
    {   // block-skipping auto mode: device scratch + host-mapped report slots (optional: without them auto mode never skips)
        void* host = nullptr;
        if (cudaMalloc(&e->skip_scratch, 2 * sizeof(int32_t)) == cudaSuccess && cudaHostAlloc(&host, 8 * 2 * sizeof(int32_t), cudaHostAllocMapped) == cudaSuccess) {
            std::memset(host, 0, 8 * 2 * sizeof(int32_t));
            cudaMemset(e->skip_scratch, 0, 2 * sizeof(int32_t));
            e->skip_report = static_cast<volatile int32_t*>(host);
            void* dev = nullptr;
            if (cudaHostGetDevicePointer(&dev, host, 0) == cudaSuccess) e->skip_report_dev = static_cast<int32_t*>(dev);
            else { cudaFreeHost(host); e->skip_report = nullptr; }
        }
        cudaGetLastError();
    }
    if (a1 != cudaSuccess || a2 != cudaSuccess || a3 != cudaSuccess || a4 != cudaSuccess || a5 != cudaSuccess) {
        vosprop_destroy(e);
        return fail(VOSPROP_ERR_CUDA, "cudaMalloc of the reference-memory ring failed (%zu rows)", rows);
    }
    // pad rows are read by TMA (and masked in the epilogue); keep them finite
    cudaMemset(e->ring_hi, 0, rows * vosk::kK * 2);
    cudaMemset(e->ring_lo, 0, rows * vosk::kK * 2);
    cudaMemset(e->meta, 0, rows * vosk::kMetaFloats * 4);
    ce = cudaDeviceSynchronize();
    if (ce != cudaSuccess) {
        vosprop_destroy(e);
        return fail(VOSPROP_ERR_CUDA, "ring initialisation failed: %s", cudaGetErrorString(ce));
    }
    *out = e;
    return VOSPROP_OK;
}

void vosprop_destroy(vosprop_engine* e) {
    if (!e) return;
    cudaFree(e->ring_hi);
    cudaFree(e->ring_lo);
    cudaFree(e->meta);
    cudaFree(e->partials);
    cudaFree(e->cls);
    cudaFree(e->cand_key);
    cudaFree(e->cand_idx);
    cudaFree(e->cand_cnt);
    cudaFree(e->low_scratch);
    cudaFree(e->skip_scratch);
    if (e->skip_report) cudaFreeHost(const_cast<int32_t*>(e->skip_report));
    cudaFree(e->topk_bound);
    cudaFree(e->topk_tau);
    cudaFree(e->tables);
    for (cudaEvent_t ev : e->ev) cudaEventDestroy(ev);
    delete e;
}

int vosprop_reset(vosprop_engine* e, int32_t H_d, int32_t W_d, int32_t H, int32_t W, int32_t d, int32_t precision,
                  void* stream) {
    if (!e) return fail(VOSPROP_ERR_INVALID, "null engine");
    if (precision < VOSPROP_PREC_SPLIT3 || precision > VOSPROP_PREC_BF16) return fail(VOSPROP_ERR_INVALID, "unknown precision %d", precision);
    if (H_d <= 0 || W_d <= 0 || H <= 0 || W <= 0) return fail(VOSPROP_ERR_INVALID, "bad geometry %dx%d / %dx%d", H_d, W_d, H, W);
    if (d < 1 || d > VOSPROP_MAX_CLASSES) return fail(VOSPROP_ERR_UNSUPPORTED, "d=%d classes; supported 1..%d", d, VOSPROP_MAX_CLASSES);
    const int64_t P = static_cast<int64_t>(H_d) * W_d;
    if (P > e->cfg.max_pixels) return fail(VOSPROP_ERR_UNSUPPORTED, "%lld pixels > engine capacity %d", (long long)P, e->cfg.max_pixels);
    if (W_d > 4096) return fail(VOSPROP_ERR_UNSUPPORTED, "W_d=%d too wide", W_d);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    e->H_d = H_d; e->W_d = W_d; e->H = H; e->W = W; e->d = d;
    e->precision = precision;
    e->P = static_cast<int>(P);
    e->p_pad = (e->P + vosk::kTile - 1) / vosk::kTile * vosk::kTile;
    int rc = encode_maps(e);
    if (rc) return rc;
    std::fill(e->slot_frame.begin(), e->slot_frame.end(), -1);
    std::fill(e->slot_labels.begin(), e->slot_labels.end(), 0);
    std::fill(e->table_valid.begin(), e->table_valid.end(), 0);
    const size_t n = static_cast<size_t>(e->cfg.ring_slots) * e->p_pad;
    vosk::vos_init_meta<<<static_cast<unsigned>((n + 255) / 256), 256, 0, st>>>(e->meta, e->cfg.ring_slots, e->p_pad, e->P, W_d);
    VOS_CUDA(cudaGetLastError());
    VOS_CUDA(cudaMemsetAsync(e->cls, 0xFF, n, st));   // padding pixels keep class 0xFF
    e->launches++;
    return VOSPROP_OK;
}

// One launch appends n consecutive frames (side_kernels.cuh: AppendSlots); the source holds them back to back.
static int append_batch(vosprop_engine* e, int32_t first_frame, int32_t n, const void* features, int32_t dtype, int32_t layout,
                        cudaStream_t st) {
    int rc = check_frame(e, first_frame);
    if (rc) return rc;
    if (!features) return fail(VOSPROP_ERR_INVALID, "null features");
    if (n < 1 || n > e->cfg.ring_slots) return fail(VOSPROP_ERR_INVALID, "n_frames=%d outside 1..ring_slots=%d", n, e->cfg.ring_slots);
    const int P = e->P;
    if (dtype != VOSPROP_F32 && dtype != VOSPROP_F16 && dtype != VOSPROP_BF16) return fail(VOSPROP_ERR_INVALID, "unknown dtype %d", dtype);
    if ((e->precision == VOSPROP_PREC_F16 && dtype != VOSPROP_F16) || (e->precision == VOSPROP_PREC_BF16 && dtype != VOSPROP_BF16))
        return fail(VOSPROP_ERR_INVALID, "this video was reset in %s precision: embeddings must arrive in that dtype (got dtype %d); "
                    "use VOSPROP_PREC_SPLIT3 for fp32 embeddings", e->precision == VOSPROP_PREC_F16 ? "F16" : "BF16", dtype);
    const int fmt = e->precision;   // enum values coincide with vosk::kFmt*
    const vosk::AppendSlots sl{first_frame, e->cfg.ring_slots, e->p_pad};
    const unsigned nf = static_cast<unsigned>(n);
    TimedLaunch timed(e, VOSPROP_T_APPEND, st);
    if (layout == VOSPROP_NCHW) {
        const dim3 grid((P + 31) / 32, vosk::kK / 64, nf);
        if (dtype == VOSPROP_F32) VOS_CUDA(launch_pdl(vosk::vos_append_nchw<float>, grid, 256, 0, st, static_cast<const float*>(features), e->ring_hi, e->ring_lo, P, sl, fmt));
        else if (dtype == VOSPROP_F16) VOS_CUDA(launch_pdl(vosk::vos_append_nchw<__half>, grid, 256, 0, st, static_cast<const __half*>(features), e->ring_hi, e->ring_lo, P, sl, fmt));
        else VOS_CUDA(launch_pdl(vosk::vos_append_nchw<__nv_bfloat16>, grid, 256, 0, st, static_cast<const __nv_bfloat16*>(features), e->ring_hi, e->ring_lo, P, sl, fmt));
    } else if (layout == VOSPROP_NHWC) {
        const size_t frame_bytes = static_cast<size_t>(vosk::kK) * P * (dtype == VOSPROP_F32 ? 4 : 2);
        if (reinterpret_cast<uintptr_t>(features) % 16 == 0 && (n == 1 || frame_bytes % 16 == 0)) {
            const dim3 grid(static_cast<unsigned>((static_cast<size_t>(P) * (vosk::kK / 8) + 255) / 256), nf);
            if (dtype == VOSPROP_F32) VOS_CUDA(launch_pdl(vosk::vos_append_nhwc8<float>, grid, 256, 0, st, static_cast<const float*>(features), e->ring_hi, e->ring_lo, P, sl, fmt));
            else if (dtype == VOSPROP_F16) VOS_CUDA(launch_pdl(vosk::vos_append_nhwc8<__half>, grid, 256, 0, st, static_cast<const __half*>(features), e->ring_hi, e->ring_lo, P, sl, fmt));
            else VOS_CUDA(launch_pdl(vosk::vos_append_nhwc8<__nv_bfloat16>, grid, 256, 0, st, static_cast<const __nv_bfloat16*>(features), e->ring_hi, e->ring_lo, P, sl, fmt));
        } else {
            const dim3 grid(static_cast<unsigned>((static_cast<size_t>(P) * (vosk::kK / 2) + 255) / 256), nf);
            if (dtype == VOSPROP_F32) VOS_CUDA(launch_pdl(vosk::vos_append_nhwc<float>, grid, 256, 0, st, static_cast<const float*>(features), e->ring_hi, e->ring_lo, P, sl, fmt));
            else if (dtype == VOSPROP_F16) VOS_CUDA(launch_pdl(vosk::vos_append_nhwc<__half>, grid, 256, 0, st, static_cast<const __half*>(features), e->ring_hi, e->ring_lo, P, sl, fmt));
            else VOS_CUDA(launch_pdl(vosk::vos_append_nhwc<__nv_bfloat16>, grid, 256, 0, st, static_cast<const __nv_bfloat16*>(features), e->ring_hi, e->ring_lo, P, sl, fmt));
        }
    } else {
        return fail(VOSPROP_ERR_INVALID, "unknown layout %d", layout);
    }
    VOS_CUDA(cudaGetLastError());
    for (int i = 0; i < n; ++i) {
        const int slot = (first_frame + i) % e->cfg.ring_slots;
        e->slot_frame[slot] = first_frame + i;
        e->slot_labels[slot] = 0;
    }
    e->launches++;
    return VOSPROP_OK;
}

int vosprop_append_features(vosprop_engine* e, int32_t frame_idx, const void* features, int32_t dtype,
                            int32_t layout, void* stream) {
    return append_batch(e, frame_idx, 1, features, dtype, layout, static_cast<cudaStream_t>(stream));
}

static int labels_target(vosprop_engine* e, int frame_idx, float** meta_slot, int kind) {
    int rc = check_frame(e, frame_idx);
    if (rc) return rc;
    const int slot = frame_idx % e->cfg.ring_slots;
    if (e->slot_frame[slot] != frame_idx)
        return fail(VOSPROP_ERR_STATE, "frame %d is not in the ring (slot %d holds %d): append its features first", frame_idx, slot, e->slot_frame[slot]);
    *meta_slot = e->meta + static_cast<size_t>(slot) * e->p_pad * vosk::kMetaFloats;
    e->slot_labels[slot] = static_cast<char>(kind);
    return VOSPROP_OK;
}

int vosprop_set_labels_index(vosprop_engine* e, int32_t frame_idx, const uint8_t* class_idx, void* stream) {
    float* ms = nullptr;
    int rc = labels_target(e, frame_idx, &ms, 1);
    if (rc) return rc;
    if (!class_idx) return fail(VOSPROP_ERR_INVALID, "null class_idx");
    uint8_t* cls_slot = e->cls + static_cast<size_t>(frame_idx % e->cfg.ring_slots) * e->p_pad;
    vosk::vos_set_labels_index<<<(e->P + 255) / 256, 256, 0, static_cast<cudaStream_t>(stream)>>>(ms, cls_slot, class_idx, e->P, e->d);
    VOS_CUDA(cudaGetLastError());
    e->launches++;
    return VOSPROP_OK;
}

int vosprop_append_frames(vosprop_engine* e, int32_t first_frame_idx, int32_t n_frames, const void* features, int32_t dtype,
                          int32_t layout, const uint8_t* class_idx, void* stream) {
    if (!e) return fail(VOSPROP_ERR_INVALID, "null engine");
    if (n_frames < 0 || n_frames > e->cfg.ring_slots)
        return fail(VOSPROP_ERR_INVALID, "n_frames=%d outside 0..ring_slots=%d", n_frames, e->cfg.ring_slots);
    if (n_frames == 0) return VOSPROP_OK;
    int rc = append_batch(e, first_frame_idx, n_frames, features, dtype, layout, static_cast<cudaStream_t>(stream));
    if (rc) return rc;
    for (int i = 0; class_idx && i < n_frames; ++i) {
        rc = vosprop_set_labels_index(e, first_frame_idx + i, class_idx + static_cast<size_t>(i) * e->P, stream);
        if (rc) return rc;
    }
    return VOSPROP_OK;
}

int vosprop_set_labels_dense(vosprop_engine* e, int32_t frame_idx, const float* labels, void* stream) {
    float* ms = nullptr;
    if (e && e->d > VOSPROP_MAX_DENSE_CLASSES)
        return fail(VOSPROP_ERR_UNSUPPORTED, "dense labels hold at most %d classes (d=%d): use index labels", VOSPROP_MAX_DENSE_CLASSES, e->d);
    int rc = labels_target(e, frame_idx, &ms, 2);
    if (rc) return rc;
    if (!labels) return fail(VOSPROP_ERR_INVALID, "null labels");
    vosk::vos_set_labels_dense<<<(e->P + 255) / 256, 256, 0, static_cast<cudaStream_t>(stream)>>>(ms, labels, e->P, e->d);
    VOS_CUDA(cudaGetLastError());
    e->launches++;
    return VOSPROP_OK;
}

int vosprop_propagate(vosprop_engine* e, const vosprop_step* s, void* stream) {
    if (!s) return fail(VOSPROP_ERR_INVALID, "null step");
    int rc = check_frame(e, s->frame_idx);
    if (rc) return rc;
    if (s->n_refs < 1 || s->n_refs > VOSPROP_MAX_REFS) return fail(VOSPROP_ERR_INVALID, "n_refs=%d outside 1..%d", s->n_refs, VOSPROP_MAX_REFS);
    if (!(s->temperature >= 0.f) || !std::isfinite(s->temperature))
        return fail(VOSPROP_ERR_UNSUPPORTED, "temperature %g: only finite temperature >= 0 is supported", (double)s->temperature);
    if (s->topk < 0 || s->topk > VOSPROP_MAX_TOPK)
        return fail(VOSPROP_ERR_UNSUPPORTED, "topk=%d outside 0..%d (0 = full softmax, the reference)", s->topk, VOSPROP_MAX_TOPK);
    if (s->out_topk_idx && s->topk == 0) return fail(VOSPROP_ERR_INVALID, "out_topk_idx needs topk > 0");
    if (s->kernel < VOSPROP_KERNEL_TC || s->kernel > VOSPROP_KERNEL_TC_DENSE) return fail(VOSPROP_ERR_INVALID, "unknown kernel %d", s->kernel);
    if (s->out_mask_fullres && static_cast<int64_t>(e->H) * e->W > e->cfg.max_fullres_pixels && e->cfg.max_fullres_pixels > 0)
        return fail(VOSPROP_ERR_UNSUPPORTED, "full-resolution frame larger than configured");
    const int S = e->cfg.ring_slots;
    const int q_slot = s->frame_idx % S;
    if (e->slot_frame[q_slot] != s->frame_idx)
        return fail(VOSPROP_ERR_STATE, "target frame %d is not in the ring: append its features first", s->frame_idx);

    bool all_index = true;
    vosk::AffinityParams ap{};
    ap.n_pixels = e->P; ap.p_pad = e->p_pad; ap.w_lowres = e->W_d; ap.n_refs = s->n_refs; ap.q_slot = q_slot;
    ap.num_sms = e->num_sms;
    for (int r = 0; r < s->n_refs; ++r) {
        const int f = s->ref_frames[r];
        if (f < 0) return fail(VOSPROP_ERR_INVALID, "negative reference frame %d", f);
        const int slot = f % S;
        if (e->slot_frame[slot] != f || !e->slot_labels[slot])
            return fail(VOSPROP_ERR_STATE, "reference frame %d is not in the ring (slot %d holds frame %d, labels=%d); ring_slots=%d too small or labels never set",
                        f, slot, e->slot_frame[slot], (int)e->slot_labels[slot], S);
        if (slot == q_slot && s->write_labels)
            return fail(VOSPROP_ERR_STATE, "reference frame %d aliases the target's ring slot", f);
        ap.ref_slot[r] = slot;
        all_index = all_index && e->slot_labels[slot] == 1;
        const float sg = s->ref_sigma[r];
        ap.ref_coef[r] = sg > 0.f ? static_cast<float>(1.4426950408889634 / (static_cast<double>(sg) * sg)) : 0.f;
    }
    ap.scale2 = static_cast<float>(static_cast<double>(s->temperature) * 1.4426950408889634);
    ap.meta = e->meta; ap.partials = e->partials; ap.ring_hi = e->ring_hi; ap.ring_lo = e->ring_lo;
    ap.cls = e->cls; ap.inv_w = 1.0f / static_cast<float>(e->W_d);
    ap.feat_fmt = e->precision;
    ap.dbg = e->dbg;
    ap.dbg_clk = e->dbg_clk;
    ap.idesc = vosptx::umma_idesc_f32acc(vosk::kTile, vosk::kTile, e->precision == VOSPROP_PREC_F16 ? 0u : 1u);
    // the index-label kernel needs one class byte per reference pixel and W_d >= 32; anything else
    // (dense / probability labels, tiny maps) runs on the general tensor-core kernel
    int kernel = s->kernel;
    if (kernel == VOSPROP_KERNEL_TC && (!all_index || e->W_d < 32)) kernel = VOSPROP_KERNEL_TC_DENSE;
    if (e->d > VOSPROP_MAX_DENSE_CLASSES && (kernel != VOSPROP_KERNEL_TC || s->topk > 0 || s->probability_propagation))
        return fail(VOSPROP_ERR_UNSUPPORTED, "d=%d > %d classes runs only on the index-label kernel: index labels, W_d >= 32 (got %d), "
                    "topk = 0, no probability propagation", e->d, VOSPROP_MAX_DENSE_CLASSES, e->W_d);
    const vosd::Decomp dec = vosd::make_decomp(e->P, s->n_refs, e->num_sms);
    if (static_cast<size_t>(dec.grid) * dec.max_segs * vosk::kIdxSub > e->partial_records)
        return fail(VOSPROP_ERR_UNSUPPORTED, "partial buffer too small (grid %d x segs %d)", dec.grid, dec.max_segs);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (s->wait_event) VOS_CUDA(cudaStreamWaitEvent(st, static_cast<cudaEvent_t>(s->wait_event), 0));
    if (s->topk > 0) return propagate_topk(e, s, ap, dec, st);
    {
        TimedLaunch timed(e, VOSPROP_T_AFFINITY, st);
        rc = dispatch_affinity(e, ap, dec.grid, kernel, st);
    }
    if (rc) return rc;
    if (s->record_event) VOS_CUDA(cudaEventRecord(static_cast<cudaEvent_t>(s->record_event), st));

    vosk::MergeParams mp{};
    mp.n_pixels = e->P; mp.p_pad = e->p_pad; mp.w_lowres = e->W_d; mp.h_lowres = e->H_d; mp.n_refs = s->n_refs;
    mp.num_sms = e->num_sms; mp.d = e->d; mp.H = e->H; mp.W = e->W; mp.q_slot = q_slot;
    mp.write_labels = s->write_labels; mp.probability = s->probability_propagation;
    mp.n_sub = (kernel == VOSPROP_KERNEL_TC || e->last_dense_kernel == 2) ? vosk::kIdxSub : 2;
    {   // which CTAs hold partials of which target tile: computed once per reference count and video (tiny kernel,
        // stream-ordered, so a host that runs clips ahead of the device never races with it)
        int32_t* dev = e->tables + static_cast<size_t>(s->n_refs) * e->table_stride;
        if (!e->table_valid[s->n_refs]) {
            mp.tables_fresh = 1;
            vosk::vos_decomp_tables<<<1, 256, 0, st>>>(dev, e->P, s->n_refs, e->num_sms);
            VOS_CUDA(cudaGetLastError());
            e->table_valid[s->n_refs] = 1;
            e->launches++;
        }
        mp.tables = dev; mp.tpf = dec.tpf; mp.max_segs = dec.max_segs;
    }
    mp.partials = e->partials; mp.meta = e->meta; mp.cls = e->cls;
    mp.out_prediction = s->out_prediction; mp.out_mask_lowres = s->out_mask_lowres; mp.out_mask_fullres = s->out_mask_fullres;
    {
        TimedLaunch timed(e, VOSPROP_T_MERGE, st);
        VOS_CUDA(launch_pdl(e->d > vosk::kMetaClasses ? vosk::vos_merge_writeback<vosk::kMaxClasses>
                            : e->d > vosk::kMergeSmall ? vosk::vos_merge_writeback<vosk::kMetaClasses> : vosk::vos_merge_writeback<vosk::kMergeSmall>,
                            dim3(e->H_d, vosk::kMergeSplit), vosk::kMergeThreads, e->W_d, st, mp));
    }
    VOS_CUDA(cudaGetLastError());
    if (s->write_labels) e->slot_labels[q_slot] = s->probability_propagation ? 2 : 1;
    e->launches += 2;
    return VOSPROP_OK;
}

int vosprop_normalize_u8(const uint8_t* rgb, int64_t n_pixels, const float* mean3, const float* std3, void* out,
                         int32_t out_dtype, void* stream) {
    if (!rgb || !out || !mean3 || !std3) return fail(VOSPROP_ERR_INVALID, "null pointer");
    if (n_pixels < 0) return fail(VOSPROP_ERR_INVALID, "negative pixel count");
    if (out_dtype != VOSPROP_F32 && out_dtype != VOSPROP_F16) return fail(VOSPROP_ERR_INVALID, "output dtype %d: F32 or F16", out_dtype);
    if ((reinterpret_cast<uintptr_t>(rgb) & 3) || (reinterpret_cast<uintptr_t>(out) & 15))
        return fail(VOSPROP_ERR_INVALID, "normalize: input must be 4-byte and output 16-byte aligned");
    if (n_pixels == 0) return VOSPROP_OK;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int threads = 256;
    const int64_t quads = (n_pixels + 3) / 4;
    const unsigned blocks = static_cast<unsigned>((quads + threads - 1) / threads);
    if (out_dtype == VOSPROP_F32)
        vosk::vos_normalize_u8<float><<<blocks, threads, 0, st>>>(rgb, static_cast<float*>(out), n_pixels, mean3[0], mean3[1], mean3[2],
                                                                  std3[0], std3[1], std3[2]);
    else
        vosk::vos_normalize_u8<__half><<<blocks, threads, 0, st>>>(rgb, static_cast<__half*>(out), n_pixels, mean3[0], mean3[1], mean3[2],
                                                                   std3[0], std3[1], std3[2]);
    VOS_CUDA(cudaGetLastError());
    return VOSPROP_OK;
}

int vosprop_sample_frames(int32_t frame_idx, int32_t take_range, int32_t num_refs, int32_t* out_idx) {
    // src/model/predict.py:74-89
    if (!out_idx) return fail(VOSPROP_ERR_INVALID, "null out_idx");
    if (frame_idx < 0) return fail(VOSPROP_ERR_INVALID, "negative frame index");
    if (frame_idx <= num_refs) {
        if (frame_idx > VOSPROP_MAX_REFS) return fail(VOSPROP_ERR_UNSUPPORTED, "more than %d references", VOSPROP_MAX_REFS);
        for (int i = 0; i < frame_idx; ++i) out_idx[i] = i;
        return frame_idx;
    }
    const int dense_num = 4 - 1;  // Config.CONTINUOUS_FRAME - 1
    const int sparse_num = num_refs - dense_num;
    if (sparse_num < 0) return fail(VOSPROP_ERR_INVALID, "num_refs=%d < 3 with frame_idx=%d: the reference raises ValueError here (np.linspace with a negative count)", num_refs, frame_idx);
    if (num_refs > VOSPROP_MAX_REFS) return fail(VOSPROP_ERR_UNSUPPORTED, "more than %d references", VOSPROP_MAX_REFS);
    const int ref_end = frame_idx - dense_num - 1;
    const int ref_start = ref_end - take_range > 0 ? ref_end - take_range : 0;
    // numpy.linspace(ref_start, ref_end, sparse_num) in float64, then astype(int) (truncation)
    const double start = ref_start, stop = ref_end, delta = stop - start;
    const int div = sparse_num - 1;
    int n = 0;
    for (int i = 0; i < sparse_num; ++i) {
        volatile double y;  // volatile: keep the separately rounded multiply and add of numpy (no FMA contraction)
        if (div > 0) {
            const double step = delta / div;
            if (step == 0.0) { y = static_cast<double>(i) / div; y = y * delta; }
            else y = static_cast<double>(i) * step;
        } else {
            y = static_cast<double>(i) * delta;
        }
        y = y + start;
        if (sparse_num > 1 && i == sparse_num - 1) y = stop;
        out_idx[n++] = static_cast<int32_t>(y);
    }
    for (int j = 0; j < dense_num; ++j) out_idx[n++] = frame_idx - dense_num + j;
    return n;
}

int vosprop_plan_step(int32_t frame_idx, int32_t take_range, int32_t num_refs, float sigma_dense, float sigma_sparse,
                      int32_t probability_propagation, vosprop_step* step) {
    if (!step) return fail(VOSPROP_ERR_INVALID, "null step");
    const int n = vosprop_sample_frames(frame_idx, take_range, num_refs, step->ref_frames);
    if (n < 0) return n;
    step->frame_idx = frame_idx;
    step->n_refs = n;
    // src/model/predict.py:59-66: frame_idx > 15 -> refs[:-4] sparse sigma, refs[-4:] dense sigma; else all dense
    for (int r = 0; r < n; ++r) {
        float sg = sigma_dense;
        if (frame_idx > 15 && r < n - 4) sg = sigma_sparse;
        step->ref_sigma[r] = probability_propagation ? 0.f : sg;
    }
    step->probability_propagation = probability_propagation;
    return VOSPROP_OK;
}

int vosprop_block_skip(vosprop_engine* e, int32_t mode) {
    if (!e) return fail(VOSPROP_ERR_INVALID, "null engine");
    if (mode < 0 || mode > 2) return fail(VOSPROP_ERR_INVALID, "block skip mode %d: 0 never, 1 always, 2 auto", mode);
    e->block_skip = mode;
    return VOSPROP_OK;
}

int vosprop_block_skip_state(const vosprop_engine* e) {
    if (!e) return VOSPROP_ERR_INVALID;
    return e->block_skip == 2 ? e->skip_on : e->block_skip;
}

int vosprop_debug_flags(vosprop_engine* e, int32_t flags) {
    if (!e) return fail(VOSPROP_ERR_INVALID, "null engine");
    e->dbg = flags;
    return VOSPROP_OK;
}

int vosprop_debug_clocks(vosprop_engine* e, void* device_buffer) {
    if (!e) return fail(VOSPROP_ERR_INVALID, "null engine");
    e->dbg_clk = static_cast<long long*>(device_buffer);
    return VOSPROP_OK;
}

int vosprop_ring_slots(const vosprop_engine* e) { return e ? e->cfg.ring_slots : VOSPROP_ERR_INVALID; }
int vosprop_num_sms(const vosprop_engine* e) { return e ? e->num_sms : VOSPROP_ERR_INVALID; }
int64_t vosprop_launch_count(const vosprop_engine* e) { return e ? e->launches : -1; }

int vosprop_timing_enable(vosprop_engine* e, int32_t capacity) {
    if (!e || capacity < 0) return fail(VOSPROP_ERR_INVALID, "bad timing request");
    while (e->ev.size() < static_cast<size_t>(capacity) * 2) {
        cudaEvent_t ev;
        VOS_CUDA(cudaEventCreate(&ev));
        e->ev.push_back(ev);
    }
    e->ev_kind.resize(e->ev.size() / 2);
    e->ev_used = 0;
    e->timing = capacity > 0;
    return VOSPROP_OK;
}

int vosprop_timing_select(vosprop_engine* e, int32_t class_mask) {
    if (!e) return fail(VOSPROP_ERR_INVALID, "null engine");
    e->timing_mask = class_mask;
    return VOSPROP_OK;
}

int vosprop_timing_read(vosprop_engine* e, double* totals_ms, int64_t* counts) {
    if (!e || !totals_ms || !counts) return fail(VOSPROP_ERR_INVALID, "null argument");
    for (int k = 0; k < 3; ++k) { totals_ms[k] = 0.0; counts[k] = 0; }
    for (size_t i = 0; i + 1 < e->ev_used; i += 2) {
        VOS_CUDA(cudaEventSynchronize(e->ev[i + 1]));
        float ms = 0.f;
        VOS_CUDA(cudaEventElapsedTime(&ms, e->ev[i], e->ev[i + 1]));
        const int k = e->ev_kind[i / 2];
        totals_ms[k] += ms;
        counts[k] += 1;
    }
    e->ev_used = 0;
    return VOSPROP_OK;
}

int vosprop_debug_decompose(int32_t n_pixels, int32_t n_refs, int32_t num_sms, int32_t* grid, int64_t* cta_begin,
                            int32_t* max_segments) {
    if (n_pixels <= 0 || n_refs <= 0 || num_sms <= 0) return fail(VOSPROP_ERR_INVALID, "bad decomposition query");
    const vosd::Decomp d = vosd::make_decomp(n_pixels, n_refs, num_sms);
    if (grid) *grid = d.grid;
    if (max_segments) *max_segments = d.max_segs;
    if (cta_begin)
        for (int c = 0; c <= d.grid; ++c) cta_begin[c] = vosd::cta_begin(d, c);
    return VOSPROP_OK;
}

}  // extern "C"
