// sm_100a building blocks for the tensor-core path: mbarrier, bulk-TMA, TMEM and tcgen05 wrappers (inline PTX),
// UMMA shared-memory / instruction descriptors and the bf16 hi/lo operand split.
//
// Operand layout used everywhere in this library (the UMMA "K-major, no swizzle / interleaved" canonical form):
//   a [rows x K] bf16 operand is a sequence of K/8 "k-core slabs"; slab kc holds, for every row r, the 8
//   consecutive K elements 8kc..8kc+7 as one 16-byte unit at byte offset  kc*LBO + r*16   (LBO = rows*16).
//   8 consecutive rows therefore form the 128-byte UMMA core matrix, row groups are 128 B apart (SBO = 128),
//   and one tcgen05.mma (K = 16) consumes two adjacent slabs.  A thread that owns one row writes one 16-byte
//   vector per slab and a warp (32 consecutive rows) writes 512 contiguous bytes: bank-conflict free.
//
// fp32 accuracy on bf16 tensor cores: x = hi + lo with hi = bf16(x), lo = bf16(x - hi) (|x - hi - lo| <= 2^-18 |x|);
// a*b ~= a_hi*b_hi + a_hi*b_lo + a_lo*b_hi, three tcgen05.mma into the same fp32 TMEM accumulator, relative
// error ~2^-17 per product -- inside the 1e-4 budget of BASELINE.json where single-pass bf16 or tf32 is not.
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace tc {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// one lane of a converged warp (CUTLASS elect_one_sync): keeps the warp's control flow uniform around single-thread issue
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "elect.sync _|p, 0xffffffff;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(pred));
    return pred != 0;
}

// register rebalancing between warp roles (whole warpgroups): data-movement / issue warps give registers to the gather warps
template <int N>
__device__ __forceinline__ void reg_dec() { asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(N)); }
template <int N>
__device__ __forceinline__ void reg_inc() { asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(N)); }

// ---- mbarrier ------------------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
#ifndef KAGNN_MBAR_HINT_NS
#define KAGNN_MBAR_HINT_NS 0
#endif
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
#if KAGNN_MBAR_HINT_NS > 0
    // suspend-time hint: upper bound on how long the hardware may park the warp before re-checking
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity), "r"((uint32_t)KAGNN_MBAR_HINT_NS)
        : "memory");
#else
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
#endif
    return ok != 0;
}
// Bounded wait: a protocol bug must trap (launch error) instead of hanging the GPU.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    uint32_t tries = 0;
    while (!mbar_try_wait(bar, parity)) {
        if (++tries > (1u << 26)) __trap();
    }
}

// Wait of a warp that is not on the critical path (it is ahead of whoever it waits for): back off with nanosleep between
// probes so that the probe loop does not take issue slots from the warps that are working.
#ifndef KAGNN_MBAR_SLEEP_NS
#define KAGNN_MBAR_SLEEP_NS 96
#endif
__device__ __forceinline__ void mbar_wait_relaxed(uint64_t* bar, uint32_t parity) {
    if (mbar_try_wait(bar, parity)) return;
    uint32_t tries = 0;
    do {
        __nanosleep(KAGNN_MBAR_SLEEP_NS);
        if (++tries > (1u << 24)) __trap();
    } while (!mbar_try_wait(bar, parity));
}

// ---- proxies / fences ---------------------------------------------------------------------------------
// generic-proxy st.shared -> visible to the async proxy (UMMA operand fetch, bulk copies)
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before_sync() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after_sync() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// ---- bulk TMA: contiguous global -> shared, completion on an mbarrier ---------------------------------
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst_smem)),
                 "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

// ---- bulk TMA: shared -> global (also peer-mapped global: posted NVLink writes), tracked by bulk async-groups ----------
__device__ __forceinline__ void bulk_s2g(void* dst_gmem, const void* src_smem, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst_gmem), "r"(smem_u32(src_smem)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
// the source shared memory of every committed group of this thread has been read (the slots may be overwritten)
__device__ __forceinline__ void bulk_wait_read_all() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
// every committed group of this thread is complete (the writes are performed)
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
// generic-proxy st.global -> visible to bulk copies (async proxy) that read the same global memory afterwards
__device__ __forceinline__ void fence_proxy_async_global() { asm volatile("fence.proxy.async.global;" ::: "memory"); }

// fire-and-forget request to bring `bytes` of global memory into L2, one 128-byte line per instruction through the LSU path
// (cp.async.bulk.prefetch.L2 would queue behind the W loader's bulk copies in the TMA unit: measured 1.7x slower layers)
__device__ __forceinline__ void prefetch_l2(const void* src_gmem, uint32_t bytes) {
    const char* p = static_cast<const char*>(src_gmem);
    for (uint32_t o = 0; o < bytes; o += 128) asm volatile("prefetch.global.L2 [%0];" ::"l"(p + o) : "memory");
}

// ---- TMEM ---------------------------------------------------------------------------------------------
// one full warp; writes the base address of `ncols` (power of two >= 32) columns to *dst_smem
__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(ncols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
// 32 lanes x 8 consecutive fp32 columns: thread t of the warp receives lane (lane_base + t), columns col..col+7
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, float (&v)[8]) {
    uint32_t r[8];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "r"(taddr)
                 : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tmem_ld4(uint32_t taddr, float (&v)[4]) {
    uint32_t r[4];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
                 : "r"(taddr)
                 : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 4; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tmem_ld2(uint32_t taddr, float (&v)[2]) {
    uint32_t r[2];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x2.b32 {%0,%1}, [%2];" : "=r"(r[0]), "=r"(r[1]) : "r"(taddr) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    v[0] = __uint_as_float(r[0]);
    v[1] = __uint_as_float(r[1]);
}
template <int N>
__device__ __forceinline__ void tmem_ldn(uint32_t taddr, float (&v)[N]);
template <>
__device__ __forceinline__ void tmem_ldn<2>(uint32_t taddr, float (&v)[2]) { tmem_ld2(taddr, v); }
template <>
__device__ __forceinline__ void tmem_ldn<4>(uint32_t taddr, float (&v)[4]) { tmem_ld4(taddr, v); }
template <>
__device__ __forceinline__ void tmem_ldn<8>(uint32_t taddr, float (&v)[8]) { tmem_ld8(taddr, v); }
__device__ __forceinline__ void tmem_ld16_nowait(uint32_t taddr, uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_ld8_nowait(uint32_t taddr, uint32_t (&r)[8]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "r"(taddr)
                 : "memory");
}
// registers -> TMEM: thread t of the warp writes lane (lane_base + t), 32-bit columns col..col+N-1.
// Used to hand the A operand (bf16 pairs, K-major: column c holds K elements 2c, 2c+1) to tcgen05.mma directly.
__device__ __forceinline__ void tmem_st4(uint32_t taddr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1,%2,%3,%4};" ::"r"(taddr), "r"(a), "r"(b), "r"(c), "r"(d)
                 : "memory");
}
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t (&r)[8]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"r"(taddr), "r"(r[0]), "r"(r[1]),
                 "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
                 : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// ---- UMMA descriptors ---------------------------------------------------------------------------------
// shared-memory matrix descriptor, K-major, SWIZZLE_NONE (bit layout: cute/arch/mma_sm100_desc.hpp SmemDescriptor):
//   [0,14) start>>4 | [16,30) LBO>>4 (K-direction core-matrix stride) | [32,46) SBO>>4 (8-row group stride)
//   [46,48) version = 1 (Blackwell) | [49,52) base offset = 0 | [61,64) layout type = 0 (no swizzle)
__device__ __forceinline__ uint64_t smem_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16) |
           ((uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32) | ((uint64_t)1 << 46);
}
// instruction descriptor for kind::f16 with bf16 A/B (both K-major) and fp32 accumulate
// (cute InstrDescriptor: c_format [4,6)=1 F32, a_format [7,10)=1 BF16, b_format [10,13)=1 BF16, n>>3 [17,23), m>>4 [24,29))
__host__ __device__ __forceinline__ uint32_t idesc_bf16_f32(int m, int n) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}
// D[tmem] (+)= A[smem] * B[smem]^T ; issued by ONE thread
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// D[tmem] (+)= A[tmem] * B[smem]^T : A operand read from tensor memory (lane = row, 32-bit column = 2 bf16 of K)
__device__ __forceinline__ void umma_bf16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
        ::"r"(tmem_d), "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// all previously issued MMAs of this thread -> arrive(1) on the mbarrier when they have completed
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// ---- bf16 hi/lo split ---------------------------------------------------------------------------------
// two fp32 -> packed bf16x2 hi and lo words (element `a` in the low half = lower address)
__device__ __forceinline__ void split2(float a, float b, uint32_t& hi, uint32_t& lo) {
    __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
    __nv_bfloat162 l = __floats2bfloat162_rn(a - __low2float(h), b - __high2float(h));
    hi = *reinterpret_cast<uint32_t*>(&h);
    lo = *reinterpret_cast<uint32_t*>(&l);
}
__device__ __forceinline__ void split8(const float (&v)[8], uint4& hi, uint4& lo) {
    split2(v[0], v[1], hi.x, lo.x);
    split2(v[2], v[3], hi.y, lo.y);
    split2(v[4], v[5], hi.z, lo.z);
    split2(v[6], v[7], hi.w, lo.w);
}

__host__ __device__ __forceinline__ uint32_t tmem_cols_pow2(uint32_t n) {
    uint32_t c = 32;
    while (c < n) c <<= 1;
    return c;
}

}  // namespace tc
