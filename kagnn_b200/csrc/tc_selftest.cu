// kagnn_tc_selftest: one 128-row tile D = A . B^T through exactly the machinery the tensor-core fused kernel uses
// (bf16 hi/lo split written with generic stores in the canonical K-major layout, fence.proxy.async, B streamed by
// bulk TMA onto an mbarrier, tcgen05.mma x3 into TMEM, tcgen05.commit, tcgen05.ld epilogue).  It exists so the
// GPU test-suite can pin descriptor encodings and TMEM addressing independently of the KAN arithmetic.
#include "common.cuh"
#include "tc_common.cuh"

namespace {

// B (N x K fp32, row-major) -> [hi slabs | lo slabs], slab kc = N rows x 8 bf16 (16 B per row)
__global__ void tc_pack_b_kernel(const float* __restrict__ B, int N, int K, uint4* __restrict__ packed) {
    int idx = blockIdx.x * blockDim.x + threadIdx.x;
    int kcs = K / 8;
    if (idx >= kcs * N) return;
    int kc = idx / N, n = idx - kc * N;
    float v[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = B[(size_t)n * K + kc * 8 + i];
    uint4 hi, lo;
    tc::split8(v, hi, lo);
    packed[(size_t)kc * N + n] = hi;
    packed[(size_t)kcs * N + (size_t)kc * N + n] = lo;
}

__global__ void __launch_bounds__(128) tc_selftest_kernel(const float* __restrict__ A, const uint4* __restrict__ Bp, int N,
                                                          int K, float* __restrict__ D, int nprod) {
    extern __shared__ __align__(128) uint8_t smem[];
    const int kcs = K / 8;
    const uint32_t a_bytes = (uint32_t)kcs * 2048u;       // one of hi / lo
    const uint32_t b_bytes = (uint32_t)kcs * N * 16u;
    uint8_t* a_hi = smem;
    uint8_t* a_lo = a_hi + a_bytes;
    uint8_t* b_hi = a_lo + a_bytes;
    uint8_t* b_lo = b_hi + b_bytes;
    uint64_t* bars = reinterpret_cast<uint64_t*>(b_lo + b_bytes);   // [0] B landed, [1] MMAs done
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2);
    const int tid = threadIdx.x, warp = tid >> 5;
    const bool a_in_tmem = nprod >= 10;      // 11 / 13: A operand handed over in tensor memory (tcgen05.st + TS-form MMA)
    if (a_in_tmem) nprod -= 10;
    const uint32_t a_col = (uint32_t)((N + 31) & ~31);
    const uint32_t ncols = tc::tmem_cols_pow2(a_in_tmem ? a_col + (uint32_t)K : (uint32_t)N);

    if (warp == 0) tc::tmem_alloc(tmem_slot, ncols);
    if (tid == 32) {
        tc::mbar_init(&bars[0], 1);
        tc::mbar_init(&bars[1], 1);
        tc::mbar_fence_init();
    }
    tc::tc_fence_before_sync();
    __syncthreads();
    tc::tc_fence_after_sync();
    const uint32_t tmem_base = *tmem_slot;

    if (tid == 0) {
        tc::mbar_arrive_expect_tx(&bars[0], 2 * b_bytes);
        tc::bulk_g2s(b_hi, Bp, 2 * b_bytes, &bars[0]);     // hi and lo are contiguous in the packed buffer
    }
    // A: thread = row, one 16-byte vector per slab
    const uint32_t lane_base_st = (uint32_t)(warp * 32) << 16;
    for (int kc = 0; kc < kcs; ++kc) {
        float v[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) v[i] = A[(size_t)tid * K + kc * 8 + i];
        uint4 hi, lo;
        tc::split8(v, hi, lo);
        if (a_in_tmem) {   // hi at columns a_col + 4kc.., lo at a_col + K/2 + 4kc..
            tc::tmem_st4(tmem_base + lane_base_st + a_col + 4u * kc, hi.x, hi.y, hi.z, hi.w);
            tc::tmem_st4(tmem_base + lane_base_st + a_col + (uint32_t)K / 2 + 4u * kc, lo.x, lo.y, lo.z, lo.w);
        } else {
            *reinterpret_cast<uint4*>(a_hi + (size_t)kc * 2048 + tid * 16) = hi;
            *reinterpret_cast<uint4*>(a_lo + (size_t)kc * 2048 + tid * 16) = lo;
        }
    }
    if (a_in_tmem) tc::tmem_st_wait();
    tc::fence_proxy_async_smem();
    tc::tc_fence_before_sync();
    __syncthreads();

    if (tid == 0) {
        tc::mbar_wait(&bars[0], 0);
        tc::tc_fence_after_sync();
        const uint32_t idesc = tc::idesc_bf16_f32(128, N);
        const uint32_t lbo_a = 2048, lbo_b = (uint32_t)N * 16u;
        uint32_t acc = 0;
        for (int ks = 0; ks < K / 16; ++ks) {
            const uint64_t dah = tc::smem_desc(tc::smem_u32(a_hi) + ks * 2 * lbo_a, lbo_a, 128);
            const uint64_t dal = tc::smem_desc(tc::smem_u32(a_lo) + ks * 2 * lbo_a, lbo_a, 128);
            const uint64_t dbh = tc::smem_desc(tc::smem_u32(b_hi) + ks * 2 * lbo_b, lbo_b, 128);
            const uint64_t dbl = tc::smem_desc(tc::smem_u32(b_lo) + ks * 2 * lbo_b, lbo_b, 128);
            if (a_in_tmem) {
                const uint32_t tah = tmem_base + a_col + 8u * ks, tal = tah + (uint32_t)K / 2;
                tc::umma_bf16_ts(tmem_base, tah, dbh, idesc, acc);
                acc = 1;
                if (nprod >= 3) {
                    tc::umma_bf16_ts(tmem_base, tah, dbl, idesc, 1);
                    tc::umma_bf16_ts(tmem_base, tal, dbh, idesc, 1);
                }
                continue;
            }
            tc::umma_bf16(tmem_base, dah, dbh, idesc, acc);
            acc = 1;
            if (nprod >= 3) {
                tc::umma_bf16(tmem_base, dah, dbl, idesc, 1);
                tc::umma_bf16(tmem_base, dal, dbh, idesc, 1);
            }
        }
        tc::umma_commit(&bars[1]);
    }
    tc::mbar_wait(&bars[1], 0);
    tc::tc_fence_after_sync();
    const uint32_t lane_base = (uint32_t)(warp * 32) << 16;
    for (int c = 0; c < N; c += 8) {
        float v[8];
        tc::tmem_ld8(tmem_base + lane_base + (uint32_t)c, v);
#pragma unroll
        for (int i = 0; i < 8; ++i) D[(size_t)tid * N + c + i] = v[i];
    }
    tc::tc_fence_before_sync();
    __syncthreads();
    if (warp == 0) tc::tmem_dealloc(tmem_base, ncols);
}

}  // namespace

extern "C" size_t kagnn_tc_selftest_workspace(int32_t N, int32_t K) {
    if (N <= 0 || K <= 0) return 0;
    return (size_t)2 * N * K * sizeof(uint16_t);
}

extern "C" int kagnn_tc_selftest(const float* A, const float* B, int32_t N, int32_t K, float* D, int32_t nprod,
                                 void* workspace, size_t workspace_bytes, void* stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    if (!A || !B || !D || !workspace) return KAGNN_EINVAL;
    if (N < 16 || N > 256 || (N % 16) != 0 || K < 16 || (K % 16) != 0) return KAGNN_EUNSUPPORTED;
    if (nprod >= 10 && ((N + 31) & ~31) + K > 512) return KAGNN_EUNSUPPORTED;   // A and D share the 512 TMEM columns
    if (workspace_bytes < kagnn_tc_selftest_workspace(N, K)) return KAGNN_EWORKSPACE;
    if (!aligned16(workspace)) return KAGNN_EALIGN;
    DeviceProps props{};
    int rc = kagnn_get_props(&props);
    if (rc != KAGNN_OK) return rc;
    if (props.cc_major != 10) return KAGNN_EUNSUPPORTED;
    size_t smem = (size_t)(K / 8) * 2048 * 2 + (size_t)(K / 8) * N * 16 * 2 + 64;
    if (smem > (size_t)props.max_smem) return KAGNN_EUNSUPPORTED;
    int total = (K / 8) * N;
    tc_pack_b_kernel<<<(total + 127) / 128, 128, 0, stream>>>(B, N, K, static_cast<uint4*>(workspace));
    KAGNN_LAUNCH_CHECK();
    KAGNN_CUDA_TRY(cudaFuncSetAttribute(tc_selftest_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)props.max_smem));
    tc_selftest_kernel<<<1, 128, smem, stream>>>(A, static_cast<const uint4*>(workspace), N, K, D, nprod);
    KAGNN_LAUNCH_CHECK();
    return KAGNN_OK;
}
