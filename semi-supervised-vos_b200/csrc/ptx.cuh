// Thin inline-PTX wrappers for sm_100a: mbarrier, TMA (tensor + 1-D bulk), tcgen05 (UMMA) and
// TMEM.  Nothing here is generic -- it is exactly what vos_affinity_tc needs.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace vosptx {

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "elect.sync _|p, 0xffffffff;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}\n"
        : "=r"(pred));
    return pred != 0;
}

// ---------------------------------------------------------------- mbarrier
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void fence_mbar_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)),
                 "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}\n"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
// Bounded wait: a protocol bug must surface as a trapped kernel (cudaErrorLaunchFailure), never
// as a hung GPU.  ~2 s at 2 GHz is far beyond any legitimate wait in this kernel.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    if (mbar_try_wait(bar, parity)) return;
    const long long t0 = clock64();
    while (!mbar_try_wait(bar, parity)) {
        if (clock64() - t0 > 4000000000LL) __trap();
    }
}

// Same, for single-thread producer roles that are usually far ahead: back off between polls so the
// spin does not steal issue slots from the epilogue warps sharing the scheduler.
__device__ __forceinline__ void mbar_wait_relaxed(uint64_t* bar, uint32_t parity, uint32_t ns) {
    if (mbar_try_wait(bar, parity)) return;
    const long long t0 = clock64();
    while (!mbar_try_wait(bar, parity)) {
        __nanosleep(ns);
        if (clock64() - t0 > 4000000000LL) __trap();
    }
}

// ---- the same on 32-bit shared-window addresses (kept in registers once, no generic->shared conversion per use)
__device__ __forceinline__ void mbar_init_s(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_arrive_s(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx_s(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait_s(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}\n"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
    return ok != 0;
}
// Bounded like mbar_wait, but the bound is a poll count (each try_wait suspends for a hardware time slice),
// so the spin body is three instructions.
__device__ __forceinline__ void mbar_wait_s(uint32_t bar, uint32_t parity) {
    if (mbar_try_wait_s(bar, parity)) return;
    uint32_t polls = 0;
    while (!mbar_try_wait_s(bar, parity)) {
        if (++polls > (1u << 26)) __trap();
    }
}
// try_wait with a suspend-time hint: the warp sleeps in hardware until the phase completes or `hint_ns` pass, instead of
// coming back after the default time slice (16 epilogue warps polling cost 8 % of the fused kernel's issue slots, round 2)
__device__ __forceinline__ bool mbar_try_wait_hint_s(uint32_t bar, uint32_t parity, uint32_t hint_ns) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}\n"
        : "=r"(ok)
        : "r"(bar), "r"(parity), "r"(hint_ns)
        : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait_hint_s(uint32_t bar, uint32_t parity, uint32_t hint_ns) {
    if (mbar_try_wait_s(bar, parity)) return;
    uint32_t polls = 0;
    while (!mbar_try_wait_hint_s(bar, parity, hint_ns)) {
        if (++polls > (1u << 24)) __trap();
    }
}
__device__ __forceinline__ void mbar_wait_relaxed_s(uint32_t bar, uint32_t parity, uint32_t ns) {
    if (mbar_try_wait_s(bar, parity)) return;
    uint32_t polls = 0;
    while (!mbar_try_wait_s(bar, parity)) {
        __nanosleep(ns);
        if (++polls > (1u << 25)) __trap();
    }
}
__device__ __forceinline__ void tma_load_2d_s(uint32_t dst, const void* tmap, int x, int y, uint32_t bar) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes "
        "[%0], [%1, {%3, %4}], [%2];"
        ::"r"(dst), "l"(reinterpret_cast<uint64_t>(tmap)), "r"(bar), "r"(x), "r"(y)
        : "memory");
}
__device__ __forceinline__ void umma_commit_s(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}

// ---------------------------------------------------------------- fences
__device__ __forceinline__ void fence_proxy_async_smem() {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_before_sync() {
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_after_sync() {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}

// ---------------------------------------------------------------- TMA
__device__ __forceinline__ void prefetch_tmap(const void* tmap) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(tmap)) : "memory");
}
// 2-D tiled load: box lands at `dst` (swizzled as the tensor map says), bytes are counted on `bar`.
__device__ __forceinline__ void tma_load_2d(void* dst, const void* tmap, int x, int y, uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes "
        "[%0], [%1, {%3, %4}], [%2];"
        ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(tmap)), "r"(smem_u32(bar)), "r"(x), "r"(y)
        : "memory");
}
// 1-D bulk copy global -> shared (16-byte aligned, size multiple of 16).
__device__ __forceinline__ void bulk_load_1d(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
        ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(src)), "r"(bytes), "r"(smem_u32(bar))
        : "memory");
}

// ---------------------------------------------------------------- TMEM
template <uint32_t kCols>
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_result) {  // whole warp
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;"
                 ::"r"(smem_u32(smem_result)), "n"(kCols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
template <uint32_t kCols>
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr) {  // whole warp (the allocating one)
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "n"(kCols) : "memory");
}
// 32 lanes x 32 consecutive fp32 columns; thread i of the warp gets lane (base_lane + i).
__device__ __forceinline__ void tmem_ld_32x32b_x32(uint32_t taddr, float (&v)[32]) {
    uint32_t* r = reinterpret_cast<uint32_t*>(v);
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
          "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
          "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr)
        : "memory");
}
// store 32 consecutive 32-bit columns of this thread's lane
__device__ __forceinline__ void tmem_st_32x32b_x32(uint32_t taddr, const uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
        "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
        "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
        ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]),
          "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]),
          "r"(r[16]), "r"(r[17]), "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]),
          "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
        : "memory");
}
__device__ __forceinline__ void tmem_st_wait() {
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_ld_wait() {
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// ---------------------------------------------------------------- UMMA (tcgen05.mma)
// Shared-memory matrix descriptor, K-major operand, 128-byte swizzle, rows at 128 B pitch,
// 8-row groups at 1024 B pitch (exactly what a {64 x rows} bf16 TMA box with SWIZZLE_128B writes).
//   [0,14) start>>4 | [16,30) LBO>>4 (=1, unused for swizzled K-major) | [32,46) SBO>>4 (=64)
//   [46,48) version=1 (Blackwell) | [61,64) layout=2 (SWIZZLE_128B)
__device__ __forceinline__ uint64_t umma_desc_kmajor_sw128(uint32_t smem_addr) {
    return static_cast<uint64_t>((smem_addr & 0x3FFFFu) >> 4) | (1ull << 16) | (64ull << 32) |
           (1ull << 46) | (2ull << 61);
}
// Instruction descriptor, kind::f16: D=f32, A and B of the same 16-bit format, both K-major, dense.
//   [4,6) c_format=1(f32) | [7,10) a_format (0 = f16, 1 = bf16) | [10,13) b_format
//   [15] a_major=0(K) | [16] b_major=0(K) | [17,23) N>>3 | [24,29) M>>4
__host__ __device__ constexpr uint32_t umma_idesc_f32acc(uint32_t M, uint32_t N, uint32_t ab_format) {
    return (1u << 4) | (ab_format << 7) | (ab_format << 10) | ((N >> 3) << 17) | ((M >> 4) << 24);
}
__host__ __device__ constexpr uint32_t umma_idesc_bf16_f32(uint32_t M, uint32_t N) { return umma_idesc_f32acc(M, N, 1u); }
// D[tmem] (+)= A[smem] * B[smem]^T   (one thread issues for the CTA)
__device__ __forceinline__ void umma_bf16_ss(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc,
                                             uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n"
        ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// D[tmem] (+)= A[tmem] * B[smem]^T : A = 128 lanes x 8 columns per K=16 step (two bf16 per 32-bit column)
__device__ __forceinline__ void umma_bf16_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc,
                                             uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}\n"
        ::"r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// Arrive on `bar` when every previously issued tcgen05.mma of this thread has completed
// (implies tcgen05.fence::before_thread_sync).
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];"
                 ::"r"(smem_u32(bar)) : "memory");
}

// an opaque copy: the compiler must keep the value in a register instead of recomputing it where it is used
__device__ __forceinline__ uint32_t pin_reg(uint32_t x) {
    asm volatile("" : "+r"(x));
    return x;
}

// ---------------------------------------------------------------- programmatic dependent launch
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

// ---------------------------------------------------------------- math
// packed fp32 pairs (FFMA2 / FADD2 on sm_100): two lanes of work per issue slot
__device__ __forceinline__ float2 ffma2(float2 a, float2 b, float2 c) {
    float2 d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;"
        : "=l"(reinterpret_cast<uint64_t&>(d))
        : "l"(reinterpret_cast<uint64_t&>(a)), "l"(reinterpret_cast<uint64_t&>(b)), "l"(reinterpret_cast<uint64_t&>(c)));
    return d;
}
__device__ __forceinline__ float2 fadd2(float2 a, float2 b) {
    float2 d;
    asm("add.rn.f32x2 %0, %1, %2;"
        : "=l"(reinterpret_cast<uint64_t&>(d))
        : "l"(reinterpret_cast<uint64_t&>(a)), "l"(reinterpret_cast<uint64_t&>(b)));
    return d;
}
__device__ __forceinline__ float2 fmul2(float2 a, float2 b) {
    float2 d;
    asm("mul.rn.f32x2 %0, %1, %2;"
        : "=l"(reinterpret_cast<uint64_t&>(d))
        : "l"(reinterpret_cast<uint64_t&>(a)), "l"(reinterpret_cast<uint64_t&>(b)));
    return d;
}
__device__ __forceinline__ float fmax3(float a, float b, float c) {  // FMNMX3
    float d;
    asm("max.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c));
    return d;
}
// 2^x for a pair of x <= 0 on the FMA / ALU pipes instead of the MUFU (the trick of FlashAttention-4): round to the
// nearest integer with the 1.5*2^23 magic constant, degree-5 polynomial for 2^f on [-0.5, 0.5] (relative error < 3e-6),
// exponent added as an integer.  Inputs below -126 (incl. -inf) give 0 like ex2.approx.ftz.
__device__ __forceinline__ float2 ex2_poly2(float2 x) {
    const float kMagic = 12582912.f;
    x.x = fmaxf(x.x, -127.f);
    x.y = fmaxf(x.y, -127.f);
    const float2 r = fadd2(x, make_float2(kMagic, kMagic));
    const float2 n = fadd2(r, make_float2(-kMagic, -kMagic));
    const float2 f = fadd2(x, make_float2(-n.x, -n.y));
    float2 p = ffma2(f, make_float2(1.33335581e-3f, 1.33335581e-3f), make_float2(9.61812911e-3f, 9.61812911e-3f));
    p = ffma2(p, f, make_float2(5.55041087e-2f, 5.55041087e-2f));
    p = ffma2(p, f, make_float2(2.40226507e-1f, 2.40226507e-1f));
    p = ffma2(p, f, make_float2(6.93147182e-1f, 6.93147182e-1f));
    p = ffma2(p, f, make_float2(1.f, 1.f));
    float2 out;
    out.x = __int_as_float(__float_as_int(p.x) + (__float_as_int(r.x) << 23));
    out.y = __int_as_float(__float_as_int(p.y) + (__float_as_int(r.y) << 23));
    if (x.x <= -127.f) out.x = 0.f;
    if (x.y <= -127.f) out.y = 0.f;
    return out;
}
__device__ __forceinline__ float ex2(float x) {  // MUFU.EX2, flushes denormals, ex2(-inf) = 0
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

}  // namespace vosptx
