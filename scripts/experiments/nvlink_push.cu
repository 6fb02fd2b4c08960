// Experiment (round 2): how should rows of a hidden matrix cross NVLink when the transfer has to overlap a compute kernel?
// One process, devices 0 and 1 with peer access.  Everything is launched on device 0; "remote" = memory of device 1.
//
//   push variants (device 0 writes rows INTO device 1; stores are posted, no round trip):
//     st_row     thread = row, 32-byte pieces per thread (the store pattern of the fused kernel's epilogue)
//     st_coal    16 or 32 lanes cover one row with 128-bit stores (what a shared-memory transposed epilogue would do)
//     tma_row    one bulk copy shared -> remote global per row
//     tma_tile   one bulk copy shared -> remote global per 128-row tile (rows contiguous in the destination)
//   pull variants (device 0 reads rows FROM device 1 into local memory by an id list):
//     ldg        warps, 16 rows in flight each (the kernel of kagnn_gather_rows_peer_ordered, simplified)
//     tma        bulk copies remote global -> shared -> local global, rows in flight bounded by shared memory
//   each with a grid of `ctas` blocks (whole SMs: large dynamic shared memory), to expose a per-SM cap.
// Output: one JSON object per line.  build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o nvlink_push nvlink_push.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <vector>
#include <algorithm>
#include <random>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "%s:%d %s\n", __FILE__, __LINE__, cudaGetErrorString(e_)); exit(2); } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* b, uint32_t n) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(n)); }
__device__ __forceinline__ void mbar_expect(uint64_t* b, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(b)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try(uint64_t* b, uint32_t parity) {
    uint32_t ok;
    asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}" : "=r"(ok) : "r"(smem_u32(b)), "r"(parity) : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* b, uint32_t parity) {
    for (long long i = 0; i < (1ll << 26); ++i) if (mbar_try(b, parity)) return;
    __trap();
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void bulk_s2g(void* dst, const void* src, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(smem_u32(src)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory"); }
template <int N> __device__ __forceinline__ void bulk_wait() { asm volatile("cp.async.bulk.wait_group %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ void fence_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

extern __shared__ __align__(128) unsigned char dyn_smem[];

// ---------------------------------------------------------------- push
// thread = row (512 threads: 4 warpgroups over the 128 rows of a tile, warpgroup wg writes the 8-column groups jb = wg, wg+4, ..)
__global__ void __launch_bounds__(512) push_st_row(float* __restrict__ dst, const int* __restrict__ pos, long long rows, int f, int ld) {
    const int wg = threadIdx.x >> 7, r = threadIdx.x & 127;
    const long long tiles = (rows + 127) / 128;
    for (long long t = blockIdx.x; t < tiles; t += gridDim.x) {
        const long long row = t * 128 + r;
        if (row >= rows) continue;
        float* y = dst + (pos ? (long long)pos[row] : row) * ld;
        const float v = (float)row;
        for (int jb = wg; jb < f / 8; jb += 4) {
            *reinterpret_cast<float4*>(y + 8 * jb) = make_float4(v, v + 1, v + 2, v + 3);
            *reinterpret_cast<float4*>(y + 8 * jb + 4) = make_float4(v, v + 1, v + 2, v + 3);
        }
    }
}
// f/4 lanes per row
__global__ void __launch_bounds__(512) push_st_coal(float* __restrict__ dst, const int* __restrict__ pos, long long rows, int f, int ld) {
    const int lpr = f / 4, rpi = 512 / lpr;                         // lanes per row, rows per block iteration
    const int sub = threadIdx.x / lpr, c = (threadIdx.x % lpr) * 4;
    const long long tiles = (rows + 127) / 128;
    for (long long t = blockIdx.x; t < tiles; t += gridDim.x) {
        for (int r0 = 0; r0 < 128; r0 += rpi) {
            const long long row = t * 128 + r0 + sub;
            if (row >= rows) continue;
            float* y = dst + (pos ? (long long)pos[row] : row) * ld;
            const float v = (float)row;
            *reinterpret_cast<float4*>(y + c) = make_float4(v, v + 1, v + 2, v + 3);
        }
    }
}
// tile staged in shared memory (written once: the experiment isolates the store path), one bulk copy per row or per tile
template <bool PER_TILE>
__global__ void __launch_bounds__(512) push_tma(float* __restrict__ dst, const int* __restrict__ pos, long long rows, int f, int ld) {
    float* stage = reinterpret_cast<float*>(dyn_smem);              // 128 x f
    for (int i = threadIdx.x; i < 128 * f; i += 512) stage[i] = (float)i;
    fence_async();
    __syncthreads();
    const long long tiles = (rows + 127) / 128;
    for (long long t = blockIdx.x; t < tiles; t += gridDim.x) {
        const int nr = (int)min(128ll, rows - t * 128);
        if (PER_TILE) {
            if (threadIdx.x == 0) {
                bulk_s2g(dst + t * 128 * ld, stage, (uint32_t)(nr * f * 4));
                bulk_commit();
                bulk_wait_read<0>();                                  // the staging tile may be overwritten by the next epilogue
            }
        } else if (threadIdx.x < nr) {
            const long long row = t * 128 + threadIdx.x;
            bulk_s2g(dst + (pos ? (long long)pos[row] : row) * ld, stage + threadIdx.x * f, (uint32_t)(f * 4));
            bulk_commit();
            bulk_wait_read<0>();
        }
        __syncthreads();
    }
    if (threadIdx.x < 128) bulk_wait<0>();
}

// ---------------------------------------------------------------- pull
constexpr int PULL_U = 16;
__global__ void __launch_bounds__(512) pull_ldg(const float* __restrict__ src, const int* __restrict__ ids, long long n, int f, float* __restrict__ out) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int vpr = f / 4;                                            // float4 per row; lanes >= vpr idle when f = 64
    const long long nb = (n + PULL_U - 1) / PULL_U;
    for (long long b = (long long)blockIdx.x * 16 + warp; b < nb; b += (long long)gridDim.x * 16) {
        float4 v[PULL_U];
        int id[PULL_U];
#pragma unroll
        for (int u = 0; u < PULL_U; ++u) { const long long i = b * PULL_U + u; id[u] = i < n ? __ldg(ids + i) : -1; }
#pragma unroll
        for (int u = 0; u < PULL_U; ++u)
            if (id[u] >= 0 && lane < vpr) v[u] = __ldg(reinterpret_cast<const float4*>(src + (long long)id[u] * f) + lane);
#pragma unroll
        for (int u = 0; u < PULL_U; ++u)
            if (id[u] >= 0 && lane < vpr) reinterpret_cast<float4*>(out + (b * PULL_U + u) * f)[lane] = v[u];
    }
}
// every warp: two batches of 32 rows; lane = row of the batch.  remote -> shared by bulk copy, shared -> local global by bulk copy
template <int NW>
__global__ void __launch_bounds__(NW * 32) pull_tma(const float* __restrict__ src, const int* __restrict__ ids, long long n, int f, float* __restrict__ out) {
    __shared__ uint64_t bars[NW][2];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rb = (uint32_t)f * 4;
    unsigned char* mine = dyn_smem + (size_t)warp * 2 * 32 * rb;
    if (lane == 0) { mbar_init(&bars[warp][0], 1); mbar_init(&bars[warp][1], 1); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    __syncwarp();
    const long long nb = (n + 31) / 32;
    const long long stride = (long long)gridDim.x * NW;
    long long b = (long long)blockIdx.x * NW + warp;
    uint32_t ph[2] = {0, 0};
    auto issue = [&](long long bb, int s) {
        const long long i = bb * 32 + lane;
        const int cnt = (int)min(32ll, n - bb * 32);
        if (lane == 0) mbar_expect(&bars[warp][s], (uint32_t)cnt * rb);
        __syncwarp();
        if (lane < cnt) bulk_g2s(mine + ((size_t)s * 32 + lane) * rb, src + (long long)__ldg(ids + i) * f, rb, &bars[warp][s]);
    };
    int s = 0;
    if (b < nb) issue(b, 0);
    for (; b < nb; b += stride, s ^= 1) {
        const long long nxt = b + stride;
        if (nxt < nb) {
            bulk_wait_read<0>();                                      // the other buffer's stores have read their rows
            issue(nxt, s ^ 1);
        }
        mbar_wait(&bars[warp][s], ph[s]);
        ph[s] ^= 1;
        const int cnt = (int)min(32ll, n - b * 32);
        if (lane < cnt) bulk_s2g(out + (b * 32 + lane) * f, mine + ((size_t)s * 32 + lane) * rb, rb);
        bulk_commit();
    }
    bulk_wait<0>();
}

template <class F>
static float time_ms(F&& launch, int reps = 5) {
    cudaEvent_t a, b;
    CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
    launch(); CK(cudaDeviceSynchronize());
    float best = 1e30f;
    for (int i = 0; i < reps; ++i) {
        CK(cudaEventRecord(a)); launch(); CK(cudaEventRecord(b)); CK(cudaEventSynchronize(b));
        float ms; CK(cudaEventElapsedTime(&ms, a, b)); best = std::min(best, ms);
    }
    CK(cudaGetLastError());
    return best;
}

int main() {
    int ndev = 0; CK(cudaGetDeviceCount(&ndev));
    const bool two = ndev >= 2;
    if (two) {
        int can = 0; CK(cudaDeviceCanAccessPeer(&can, 0, 1));
        if (!can) { printf("{\"error\": \"no peer access\"}\n"); return 0; }
        CK(cudaSetDevice(0)); CK(cudaDeviceEnablePeerAccess(1, 0));
    }
    const long long rows = 169343;
    std::mt19937 rng(1);
    for (int f : {64, 128}) {
        const size_t bytes = (size_t)rows * f * 4;
        float *remote, *local;
        CK(cudaSetDevice(two ? 1 : 0)); CK(cudaMalloc(&remote, bytes)); CK(cudaMemset(remote, 0, bytes)); CK(cudaDeviceSynchronize());
        CK(cudaSetDevice(0)); CK(cudaMalloc(&local, bytes));
        std::vector<int> perm(rows);
        for (long long i = 0; i < rows; ++i) perm[i] = (int)i;
        std::shuffle(perm.begin(), perm.end(), rng);
        std::vector<int> sorted_ids(perm.begin(), perm.begin() + (rows * 97) / 100);       // N=2: 97 % of the peer's rows are needed
        std::sort(sorted_ids.begin(), sorted_ids.end());
        int *d_perm, *d_ids, *d_ids_rand;
        CK(cudaMalloc(&d_perm, rows * 4)); CK(cudaMemcpy(d_perm, perm.data(), rows * 4, cudaMemcpyHostToDevice));
        CK(cudaMalloc(&d_ids, sorted_ids.size() * 4)); CK(cudaMemcpy(d_ids, sorted_ids.data(), sorted_ids.size() * 4, cudaMemcpyHostToDevice));
        CK(cudaMalloc(&d_ids_rand, sorted_ids.size() * 4)); CK(cudaMemcpy(d_ids_rand, perm.data(), sorted_ids.size() * 4, cudaMemcpyHostToDevice));
        const long long nid = (long long)sorted_ids.size();
        CK(cudaFuncSetAttribute(push_tma<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        CK(cudaFuncSetAttribute(push_tma<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        CK(cudaFuncSetAttribute(push_st_row, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        CK(cudaFuncSetAttribute(push_st_coal, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        CK(cudaFuncSetAttribute(pull_ldg, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        CK(cudaFuncSetAttribute(pull_tma<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        CK(cudaFuncSetAttribute(pull_tma<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        const int whole = 180 * 1024;                                  // claims the SM: one block per SM
        for (int ctas : {8, 16, 32, 148}) {
            auto rep = [&](const char* name, const char* target, float ms, double b) {
                printf("{\"kernel\": \"%s\", \"target\": \"%s\", \"f\": %d, \"ctas\": %d, \"ms\": %.4f, \"GBps\": %.1f, \"GBps_per_cta\": %.2f}\n", name, target, f, ctas, ms,
                       b / ms * 1e-6, b / ms * 1e-6 / ctas);
                fflush(stdout);
            };
            for (int tgt = 0; tgt < 2; ++tgt) {
                float* d = tgt ? remote : local;
                const char* tn = tgt ? (two ? "remote" : "local2") : "local";
                if (tgt == 0 && ctas != 148) continue;               // local numbers only as the full-grid reference
                rep("push_st_row", tn, time_ms([&] { push_st_row<<<ctas, 512, whole>>>(d, nullptr, rows, f, f); }), (double)bytes);
                rep("push_st_row_scatter", tn, time_ms([&] { push_st_row<<<ctas, 512, whole>>>(d, d_perm, rows, f, f); }), (double)bytes);
                rep("push_st_coal", tn, time_ms([&] { push_st_coal<<<ctas, 512, whole>>>(d, nullptr, rows, f, f); }), (double)bytes);
                rep("push_st_coal_scatter", tn, time_ms([&] { push_st_coal<<<ctas, 512, whole>>>(d, d_perm, rows, f, f); }), (double)bytes);
                rep("push_tma_row", tn, time_ms([&] { push_tma<false><<<ctas, 512, whole>>>(d, nullptr, rows, f, f); }), (double)bytes);
                rep("push_tma_row_scatter", tn, time_ms([&] { push_tma<false><<<ctas, 512, whole>>>(d, d_perm, rows, f, f); }), (double)bytes);
                rep("push_tma_tile", tn, time_ms([&] { push_tma<true><<<ctas, 512, whole>>>(d, nullptr, rows, f, f); }), (double)bytes);
            }
            for (int tgt = 0; tgt < 2; ++tgt) {
                const float* sp = tgt ? remote : local;
                float* op = tgt ? local : remote;                      // output always distinct from the source; remote source -> local output
                if (tgt == 0) continue;                                // pulls: remote source only
                const char* tn = two ? "remote" : "local2";
                const double pb = (double)nid * f * 4;
                rep("pull_ldg_sorted", tn, time_ms([&] { pull_ldg<<<ctas, 512, whole>>>(sp, d_ids, nid, f, op); }), pb);
                rep("pull_ldg_random", tn, time_ms([&] { pull_ldg<<<ctas, 512, whole>>>(sp, d_ids_rand, nid, f, op); }), pb);
                auto sm = [&](int nw) { return std::max(nw * 2 * 32 * f * 4, 120 * 1024); };
                rep("pull_tma4_sorted", tn, time_ms([&] { pull_tma<4><<<ctas, 128, sm(4)>>>(sp, d_ids, nid, f, op); }), pb);
                rep("pull_tma4_random", tn, time_ms([&] { pull_tma<4><<<ctas, 128, sm(4)>>>(sp, d_ids_rand, nid, f, op); }), pb);
                if (sm(8) <= 200 * 1024) {
                    rep("pull_tma8_sorted", tn, time_ms([&] { pull_tma<8><<<ctas, 256, sm(8)>>>(sp, d_ids, nid, f, op); }), pb);
                    rep("pull_tma8_random", tn, time_ms([&] { pull_tma<8><<<ctas, 256, sm(8)>>>(sp, d_ids_rand, nid, f, op); }), pb);
                }
            }
        }
        // the copy engine, for reference: the whole matrix in one peer copy
        if (two) {
            float ms = time_ms([&] { CK(cudaMemcpyPeerAsync(local, 0, remote, 1, bytes, 0)); });
            printf("{\"kernel\": \"copy_engine_pull_all\", \"f\": %d, \"ms\": %.4f, \"GBps\": %.1f}\n", f, ms, bytes / ms * 1e-6);
            ms = time_ms([&] { CK(cudaMemcpyPeerAsync(remote, 1, local, 0, bytes, 0)); });
            printf("{\"kernel\": \"copy_engine_push_all\", \"f\": %d, \"ms\": %.4f, \"GBps\": %.1f}\n", f, ms, bytes / ms * 1e-6);
        }
        // check the tma pull moved the right rows
        {
            CK(cudaMemset(local, 0, bytes));
            std::vector<float> h((size_t)rows * f);
            for (long long r = 0; r < rows; ++r) for (int c = 0; c < f; ++c) h[r * f + c] = (float)(r * 3 + c);
            CK(cudaMemcpy(remote, h.data(), bytes, cudaMemcpyDefault));
            pull_tma<4><<<16, 128, std::max(4 * 2 * 32 * f * 4, 120 * 1024)>>>(remote, d_ids_rand, nid, f, local);
            CK(cudaDeviceSynchronize());
            std::vector<float> o((size_t)nid * f);
            CK(cudaMemcpy(o.data(), local, o.size() * 4, cudaMemcpyDefault));
            long long bad = 0;
            for (long long i = 0; i < nid; ++i) for (int c = 0; c < f; ++c) bad += o[i * f + c] != (float)((long long)perm[i] * 3 + c);
            printf("{\"check\": \"pull_tma4_random\", \"f\": %d, \"mismatches\": %lld}\n", f, bad);
        }
        CK(cudaFree(local)); CK(cudaFree(d_perm)); CK(cudaFree(d_ids)); CK(cudaFree(d_ids_rand));
        CK(cudaSetDevice(two ? 1 : 0)); CK(cudaFree(remote)); CK(cudaSetDevice(0));
    }
    return 0;
}
