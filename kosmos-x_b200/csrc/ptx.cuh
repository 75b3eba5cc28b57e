// Thin inline-PTX wrappers for the sm_100a features the Kosmos-X kernels use:
// mbarrier, TMA (cp.async.bulk.tensor), tcgen05 (alloc / mma / commit / ld / fences).
// Everything here is a one-instruction wrapper; policy lives in the kernels.
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cuda.h>
#include <stdint.h>

namespace kx {

// ----------------------------------------------------------------------------- misc
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ uint32_t lane_id() {
    uint32_t l;
    asm volatile("mov.u32 %0, %%laneid;" : "=r"(l));
    return l;
}
__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ bool elect_one() {
    uint32_t pred = 0;
    asm volatile(
        "{\n\t.reg .pred P;\n\t"
        "elect.sync _|P, 0xffffffff;\n\t"
        "selp.b32 %0, 1, 0, P;\n\t}"
        : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\t"
                 "barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void named_bar_sync(uint32_t id, uint32_t nthreads) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}
__device__ __forceinline__ uint64_t globaltimer_ns() {
    uint64_t t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

// ----------------------------------------------------------------------------- mbarrier
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void fence_mbar_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// arrive on the barrier at the same smem offset in CTA `cta` of the cluster
__device__ __forceinline__ void mbar_arrive_cluster(uint64_t* bar, uint32_t cta) {
    asm volatile(
        "{\n\t.reg .b32 ra;\n\t"
        "mapa.shared::cluster.u32 ra, %0, %1;\n\t"
        "mbarrier.arrive.shared::cluster.b64 _, [ra];\n\t}" ::"r"(smem_u32(bar)),
        "r"(cta)
        : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)),
                 "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred P;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 P, [%1], %2;\n\t"
        "selp.b32 %0, 1, 0, P;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
// Bounded wait: a pipeline bug must trap (visible error) rather than hang the GPU.
#ifndef KX_WAIT_TIMEOUT_NS
#define KX_WAIT_TIMEOUT_NS 4000000000ull
#endif
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    if (mbar_try_wait(bar, parity)) return;
    uint64_t t0 = globaltimer_ns();
    uint32_t spins = 0;
    while (!mbar_try_wait(bar, parity)) {
        if ((++spins & 0x3ff) == 0 && globaltimer_ns() - t0 > KX_WAIT_TIMEOUT_NS) {
            printf("kx: mbarrier timeout block (%d,%d) thread %d bar@%u parity %u\n", blockIdx.x,
                   blockIdx.y, threadIdx.x, smem_u32(bar), parity);
            __trap();
        }
    }
}
// Same wait with the diagnostic out of line: keeps the printf call's register / stack footprint out of kernels whose
// hot loops wait on many barriers (attention backward runs at the 96-register launch-bounds budget).
static __device__ __noinline__ void mbar_timeout_trap(uint32_t bar_addr, uint32_t parity) {
    printf("kx: mbarrier timeout block (%d,%d) thread %d bar@%u parity %u\n", blockIdx.x, blockIdx.y, threadIdx.x, bar_addr, parity);
    __trap();
}
__device__ __forceinline__ void mbar_wait_lean(uint64_t* bar, uint32_t parity) {
    if (mbar_try_wait(bar, parity)) return;
    uint64_t t0 = globaltimer_ns();
    uint32_t spins = 0;
    while (!mbar_try_wait(bar, parity)) {
        if ((++spins & 0x3ff) == 0 && globaltimer_ns() - t0 > KX_WAIT_TIMEOUT_NS) mbar_timeout_trap(smem_u32(bar), parity);
    }
}

// ----------------------------------------------------------------------------- TMA
__device__ __forceinline__ void prefetch_tmap(const CUtensorMap* m) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(m)) : "memory");
}
constexpr uint64_t kEvictNormal = 0x1000000000000000ull;
constexpr uint64_t kEvictFirst = 0x12F0000000000000ull;
constexpr uint64_t kEvictLast = 0x14F0000000000000ull;

__device__ __forceinline__ void tma_load_2d(const CUtensorMap* m, uint64_t* bar, void* dst, int c0,
                                            int c1, uint64_t hint = kEvictNormal) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint"
        " [%0], [%1, {%3, %4}], [%2], %5;" ::"r"(smem_u32(dst)),
        "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "l"(hint)
        : "memory");
}
// 2-CTA form: data lands in this CTA's smem, transaction bytes complete on the LEADER CTA's
// barrier (same smem offset, peer bit cleared).
__device__ __forceinline__ void tma_load_2d_cg2(const CUtensorMap* m, uint64_t* bar, void* dst,
                                                int c0, int c1, uint64_t hint = kEvictNormal) {
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes"
        ".L2::cache_hint [%0], [%1, {%3, %4}], [%2], %5;" ::"r"(smem_u32(dst)),
        "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar) & 0xFEFFFFFFu), "r"(c0), "r"(c1),
        "l"(hint)
        : "memory");
}
// Pull a tensor-map box into L2 only (no shared-memory destination, no barrier).
__device__ __forceinline__ void tma_prefetch_l2_2d(const CUtensorMap* m, int c0, int c1) {
    asm volatile("cp.async.bulk.prefetch.tensor.2d.L2.global.tile [%0, {%1, %2}];" ::"l"(reinterpret_cast<uint64_t>(m)),
                 "r"(c0), "r"(c1)
                 : "memory");
}
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* m, const void* src, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(
                     reinterpret_cast<uint64_t>(m)),
                 "r"(smem_u32(src)), "r"(c0), "r"(c1)
                 : "memory");
}
// TMA reduction: global[box] += smem tile (element type and swizzle come from the tensor map)
__device__ __forceinline__ void tma_reduce_add_2d(const CUtensorMap* m, const void* src, int c0, int c1) {
    asm volatile("cp.reduce.async.bulk.tensor.2d.global.shared::cta.add.tile.bulk_group [%0, {%2, %3}], [%1];" ::"l"(
                     reinterpret_cast<uint64_t>(m)),
                 "r"(smem_u32(src)), "r"(c0), "r"(c1)
                 : "memory");
}
__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void tma_store_wait_read() {
    asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
template <int N>
__device__ __forceinline__ void tma_store_wait() {
    asm volatile("cp.async.bulk.wait_group %0;" ::"n"(N) : "memory");
}
// generic-proxy smem writes -> visible to the async proxy (TMA store / tcgen05.mma operand reads)
__device__ __forceinline__ void fence_proxy_async_smem() {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

// ----------------------------------------------------------------------------- tcgen05
template <int CG>
__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem, uint32_t ncols) {
    if constexpr (CG == 1)
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                         smem_u32(dst_smem)),
                     "r"(ncols)
                     : "memory");
    else
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                         smem_u32(dst_smem)),
                     "r"(ncols)
                     : "memory");
}
template <int CG>
__device__ __forceinline__ void tmem_relinquish() {
    if constexpr (CG == 1)
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    else
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
template <int CG>
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
    if constexpr (CG == 1)
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols)
                     : "memory");
    else
        asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols)
                     : "memory");
}
__device__ __forceinline__ void tc_fence_before() {
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_after() {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}

// D[tmem] (+)= A[smem] * B[smem]^T, bf16 inputs, fp32 accumulate.  One thread issues.
template <int CG>
__device__ __forceinline__ void umma_bf16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc,
                                          uint32_t idesc, uint32_t accumulate) {
    if constexpr (CG == 1)
        asm volatile(
            "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
            "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
            "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
            : "memory");
    else
        asm volatile(
            "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
            "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
            "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
            : "memory");
}
// Same MMAs with the descriptors given as (low word, high word): only the low word (start address) changes between the
// K steps of a tile, so an unrolled issue sequence costs one 32-bit add per operand and no 64-bit register pairs.
__device__ __forceinline__ void umma_bf16_lohi(uint32_t d_tmem, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo, uint32_t b_hi,
                                               uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\tsetp.ne.b32 p, %6, 0;\n\t"
        "mov.b64 da, {%1, %2};\n\tmov.b64 db, {%3, %4};\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %5, p;\n\t}" ::"r"(d_tmem),
        "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void umma_bf16_ts_lohi(uint32_t d_tmem, uint32_t a_tmem, uint32_t b_lo, uint32_t b_hi, uint32_t idesc,
                                                  uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t.reg .b64 db;\n\tsetp.ne.b32 p, %5, 0;\n\t"
        "mov.b64 db, {%2, %3};\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], db, %4, p;\n\t}" ::"r"(d_tmem),
        "r"(a_tmem), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(accumulate)
        : "memory");
}
// Arrive on an mbarrier once all previously issued tcgen05.mma of this thread have completed.
// (implies tcgen05.fence::before_thread_sync)
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(
                     smem_u32(bar))
                 : "memory");
}
// 2-CTA form: arrive on the barrier at this smem offset in every CTA of `mask`.
__device__ __forceinline__ void umma_commit_cg2(uint64_t* bar, uint16_t mask) {
    asm volatile(
        "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 "
        "[%0], %1;" ::"r"(smem_u32(bar)),
        "h"(mask)
        : "memory");
}

// TMEM -> registers: this warp's 32 lanes x 32 consecutive fp32 columns (lane i gets row i).
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]),
          "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]),
          "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]),
          "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]),
          "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr)
        : "memory");
}
// 16-column variants (lane i gets / gives 16 consecutive 32-bit columns)
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&v)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
          "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&v)[16]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
        "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(taddr),
        "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]),
        "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15])
        : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() {
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// registers -> TMEM: lane i of this warp's 32 lanes writes 32 consecutive 32-bit columns.
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t (&v)[32]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
        "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
        "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};" ::"r"(taddr),
        "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]),
        "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]), "r"(v[16]),
        "r"(v[17]), "r"(v[18]), "r"(v[19]), "r"(v[20]), "r"(v[21]), "r"(v[22]), "r"(v[23]), "r"(v[24]),
        "r"(v[25]), "r"(v[26]), "r"(v[27]), "r"(v[28]), "r"(v[29]), "r"(v[30]), "r"(v[31])
        : "memory");
}
__device__ __forceinline__ void tmem_st_wait() {
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
}
// D[tmem] (+)= A[tmem] * B[smem]: A is 128 lanes x (K/2) 32-bit columns of packed bf16 pairs (the
// softmax probabilities written with tcgen05.st), B a shared-memory descriptor.
__device__ __forceinline__ void umma_bf16_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc,
                                             uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d_tmem),
        "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}

// ----------------------------------------------------------------------------- packed fp32 math (sm_100)
__device__ __forceinline__ uint64_t pack_f32x2(float lo, float hi) {
    uint64_t r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ void unpack_f32x2(uint64_t v, float& lo, float& hi) {
    asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ uint64_t ffma2(uint64_t a, uint64_t b, uint64_t c) {
    uint64_t r;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
    return r;
}
__device__ __forceinline__ uint64_t fadd2(uint64_t a, uint64_t b) {
    uint64_t r;
    asm("add.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ uint64_t fmul2(uint64_t a, uint64_t b) {
    uint64_t r;
    asm("mul.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ float fmax3(float a, float b, float c) {
    float r;
    asm("max.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c));
    return r;
}
__device__ __forceinline__ float ex2_approx(float x) {
    float r;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}

// 2^x for two values on the FMA pipes (no MUFU): Cody-Waite split x = j + f, f in [-0.5, 0.5], degree-3
// near-minimax polynomial for 2^f (max relative error 1.6e-4, far below bf16's 3.9e-3), exponent patched in
// with an integer add.  Valid for x <= ~120; inputs below -125 (masked scores) are clamped and return ~0.
__device__ __forceinline__ void exp2_poly_x2(float x0, float x1, float& e0, float& e1) {
    x0 = fmaxf(x0, -125.0f);
    x1 = fmaxf(x1, -125.0f);
    const uint64_t x = pack_f32x2(x0, x1);
    const uint64_t t = fadd2(x, pack_f32x2(12582912.0f, 12582912.0f));          // 1.5 * 2^23: integer part in the low mantissa bits
    const uint64_t jf = fadd2(t, pack_f32x2(-12582912.0f, -12582912.0f));
    const uint64_t f = ffma2(jf, pack_f32x2(-1.0f, -1.0f), x);
    uint64_t p = ffma2(pack_f32x2(0.05360212177038193f, 0.05360212177038193f), f,
                       pack_f32x2(0.24237291514873505f, 0.24237291514873505f));
    p = ffma2(p, f, pack_f32x2(0.6935023665428162f, 0.6935023665428162f));
    p = ffma2(p, f, pack_f32x2(0.9999481439590454f, 0.9999481439590454f));
    float p0, p1, t0, t1;
    unpack_f32x2(p, p0, p1);
    unpack_f32x2(t, t0, t1);
    // exponent patch on the ALU pipe (SHF + IADD3): ptxas turns `p + (t << 23)` into one IMAD, and the FMA pipe is the one
    // that bounds the softmax loops (one warp instruction per 2 cycles and scheduler; profiles/r2_attn_softmax_scheduling_experiments.md)
    uint32_t s0, s1, r0, r1;
    asm("shf.l.clamp.b32 %0, 0, %1, 23;" : "=r"(s0) : "r"(__float_as_uint(t0)));       // (t << 23): funnel shift of {t, 0}
    asm("shf.l.clamp.b32 %0, 0, %1, 23;" : "=r"(s1) : "r"(__float_as_uint(t1)));
    asm("add.u32 %0, %1, %2;" : "=r"(r0) : "r"(__float_as_uint(p0)), "r"(s0));
    asm("add.u32 %0, %1, %2;" : "=r"(r1) : "r"(__float_as_uint(p1)), "r"(s1));
    e0 = __uint_as_float(r0);
    e1 = __uint_as_float(r1);
}

// ----------------------------------------------------------------------------- descriptors
// Shared-memory matrix descriptor (SM100 UMMA), K-major or MN-major tile stored as rows of
// exactly 128 bytes with the 128B swizzle that TMA's CU_TENSOR_MAP_SWIZZLE_128B produces.
//   bits [0,14)  start address >> 4        bits [16,30) leading byte offset >> 4
//   bits [32,46) stride byte offset >> 4   bits [46,48) version = 1
//   bits [61,64) layout type (2 = SWIZZLE_128B)
// For one 128-byte swizzle atom along the contiguous dimension, LBO is unused and
// SBO = 8 rows * 128 B = 1024 (distance between 8-row groups).
__device__ __forceinline__ uint64_t make_desc_sw128(uint32_t smem_addr, uint32_t lbo_bytes = 0) {
    uint64_t d = 0;
    d |= static_cast<uint64_t>((smem_addr & 0x3FFFF) >> 4);
    d |= static_cast<uint64_t>((lbo_bytes >> 4) & 0x3FFF) << 16;
    d |= static_cast<uint64_t>(1024 >> 4) << 32;
    d |= static_cast<uint64_t>(1) << 46;
    d |= static_cast<uint64_t>(2) << 61;
    return d;
}
// Instruction descriptor for kind::f16: bf16 x bf16 -> fp32.
//   [4,6) c_format=1 (f32)  [7,10) a_format=1 (bf16)  [10,13) b_format=1 (bf16)
//   [15] a_major (0 = K)    [16] b_major (0 = K, 1 = MN)
//   [17,23) N >> 3          [24,29) M >> 4
__host__ __device__ constexpr uint32_t make_idesc_bf16(uint32_t M, uint32_t N, uint32_t a_mn_major = 0,
                                                       uint32_t b_mn_major = 0) {
    return (1u << 4) | (1u << 7) | (1u << 10) | (a_mn_major << 15) | (b_mn_major << 16) |
           ((N >> 3) << 17) | ((M >> 4) << 24);
}

// ----------------------------------------------------------------------------- numerics
__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {
    __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<uint32_t*>(&v);
}
__device__ __forceinline__ float gelu_erf(float x) {
    return 0.5f * x * (1.0f + erff(x * 0.70710678118654752440f));
}
// erf-GELU on two values at once, ONE MUFU per value.  erf(a) = 1 - 2^(-Q(a)) for a = |x|/sqrt2 >= 0 with Q a
// degree-5 polynomial fitted to -log2(erfc(a)) (max |erf error| 8.9e-7 on [0,4]; beyond, 2^-Q underflows to the
// right limit), Horner chain on packed FFMA2.  Written as
//   gelu(x) = 0.5*(x + |x|) - 0.5*|x| * 2^(-Q(a))
// which needs no sign handling and has no cancellation in the negative tail.  |gelu error| <= 0.5|x| * 8.9e-7.
__device__ __forceinline__ void gelu_erf_x2(float& a, float& b) {
    const uint64_t z = pack_f32x2(fabsf(a) * 0.70710678118654752440f, fabsf(b) * 0.70710678118654752440f);
    uint64_t q = ffma2(z, pack_f32x2(-0.003024620935320854f, -0.003024620935320854f),
                       pack_f32x2(0.029882797971367836f, 0.029882797971367836f));
    q = ffma2(q, z, pack_f32x2(-0.14901681244373322f, -0.14901681244373322f));
    q = ffma2(q, z, pack_f32x2(-0.9183504581451416f, -0.9183504581451416f));
    q = ffma2(q, z, pack_f32x2(-1.6279104948043823f, -1.6279104948043823f));
    q = fmul2(q, z);                                                   // -Q(a)
    float q0, q1;
    unpack_f32x2(q, q0, q1);
    const uint64_t e = pack_f32x2(ex2_approx(q0), ex2_approx(q1));     // erfc(a)
    const uint64_t h = fmul2(z, pack_f32x2(0.70710678118654752440f, 0.70710678118654752440f));      // 0.5|x|
    const uint64_t r = ffma2(pack_f32x2(a, b), pack_f32x2(0.5f, 0.5f), h);                          // 0.5(x + |x|)
    const uint64_t g = ffma2(fmul2(h, e), pack_f32x2(-1.0f, -1.0f), r);
    unpack_f32x2(g, a, b);
}
__device__ __forceinline__ float quick_gelu(float x) { return x / (1.0f + __expf(-1.702f * x)); }

}  // namespace kx
