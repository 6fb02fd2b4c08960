// Recorded experiment (SURVEY.md section 7.3 item 5 / round-1 review): neighbour-row gather through the TMA
// (cp.async.bulk.tensor.2d ... tile::gather4) versus plain 128-bit loads, in isolation.
//   out[r, :] = sum_{k < DEG} x[idx[DEG r + k], :]          x: N x 128 fp32 (ogbn-arxiv-sized: 87 MB), DEG = 8
// Kernel A: warp per output row, DEG 128-bit loads in flight per lane, register accumulation (what gather_unit_fast does).
// Kernel B: persistent blocks; every warp owns a double-buffered 2 x DEG x 512 B staging area in shared memory; lane 0 issues
//           DEG / 4 gather4 copies per row onto the buffer's mbarrier, all lanes wait, LDS.128 + FADD, store.
// Build / run (GPU box):  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/tma_gather4 scripts/experiments/tma_gather4.cu -lcuda && /tmp/tma_gather4
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

constexpr int F = 128, DEG = 8;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__global__ void gather_ldg(const float* __restrict__ x, const int* __restrict__ idx, float* __restrict__ out, int rows) {
    const int r = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (r >= rows) return;
    float4 v[DEG];
#pragma unroll
    for (int k = 0; k < DEG; ++k) v[k] = __ldg(reinterpret_cast<const float4*>(x + (size_t)idx[DEG * r + k] * F) + lane);
    float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int k = 0; k < DEG; ++k) { a.x += v[k].x; a.y += v[k].y; a.z += v[k].z; a.w += v[k].w; }
    reinterpret_cast<float4*>(out + (size_t)r * F)[lane] = a;
}

template <int WARPS>
__global__ void __launch_bounds__(WARPS * 32) gather_tma(const __grid_constant__ CUtensorMap tmap, const int* __restrict__ idx,
                                                         float* __restrict__ out, int rows, int* __restrict__ fail) {
    extern __shared__ __align__(1024) uint8_t smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    uint8_t* stage = smem + (size_t)warp * 2 * DEG * F * 4;              // [2][DEG][512 B]
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (size_t)WARPS * 2 * DEG * F * 4) + 2 * warp;
    if (lane == 0) {
        for (int b = 0; b < 2; ++b) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bars[b])));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
    const int gwarp = blockIdx.x * WARPS + warp, nwarps = gridDim.x * WARPS;
    auto issue = [&](int r, int b) {
        if (lane == 0) {
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bars[b])), "r"(DEG * F * 4) : "memory");
#pragma unroll
            for (int q = 0; q < DEG / 4; ++q) {
                const int4 id = *reinterpret_cast<const int4*>(idx + DEG * r + 4 * q);
                asm volatile(
                    "cp.async.bulk.tensor.2d.shared::cta.global.tile::gather4.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5, %6}], [%7];"
                    ::"r"(smem_u32(stage + (size_t)(b * DEG + 4 * q) * F * 4)), "l"(&tmap), "r"(0), "r"(id.x), "r"(id.y), "r"(id.z), "r"(id.w),
                      "r"(smem_u32(&bars[b]))
                    : "memory");
            }
        }
    };
    int it = 0;
    if (gwarp < rows) issue(gwarp, 0);
    for (int r = gwarp; r < rows; r += nwarps, ++it) {
        const int b = it & 1;
        if (r + nwarps < rows) issue(r + nwarps, b ^ 1);
        uint32_t ok = 0, tries = 0;
        while (!ok) {
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                         : "=r"(ok) : "r"(smem_u32(&bars[b])), "r"((uint32_t)((it >> 1) & 1)) : "memory");
            if (!ok && ++tries > (1u << 22)) { if (lane == 0) atomicExch(fail, 1); return; }     // never hang the box
        }
        float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int k = 0; k < DEG; ++k) {
            const float4 v = reinterpret_cast<const float4*>(stage + (size_t)(b * DEG + k) * F * 4)[lane];
            a.x += v.x; a.y += v.y; a.z += v.z; a.w += v.w;
        }
        reinterpret_cast<float4*>(out + (size_t)r * F)[lane] = a;
        __syncwarp();                                                     // all lanes have read buffer b before it is refilled
    }
}

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("{\"error\": \"%s at line %d\"}\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)

int main() {
    const int N = 169343, rows = 166530;                                 // rows * DEG = 1.33 M gathered rows (arxiv: E + N)
    std::vector<int> hidx((size_t)rows * DEG);
    srand(1);
    for (auto& v : hidx) v = (int)(((unsigned)rand() * 32768u + (unsigned)rand()) % N);
    float *x, *oa, *ob; int *idx, *fail;
    CK(cudaMalloc(&x, (size_t)N * F * 4)); CK(cudaMalloc(&oa, (size_t)rows * F * 4)); CK(cudaMalloc(&ob, (size_t)rows * F * 4));
    CK(cudaMalloc(&idx, hidx.size() * 4)); CK(cudaMalloc(&fail, 4)); CK(cudaMemset(fail, 0, 4));
    CK(cudaMemcpy(idx, hidx.data(), hidx.size() * 4, cudaMemcpyHostToDevice));
    std::vector<float> hx((size_t)N * F);
    for (size_t i = 0; i < hx.size(); ++i) hx[i] = (float)((i * 2654435761u) % 1000) * 1e-3f;
    CK(cudaMemcpy(x, hx.data(), hx.size() * 4, cudaMemcpyHostToDevice));
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float ms_a = 1e9f;
    for (int i = 0; i < 8; ++i) {
        cudaEventRecord(e0);
        gather_ldg<<<(rows * 32 + 255) / 256, 256>>>(x, idx, oa, rows);
        cudaEventRecord(e1); CK(cudaEventSynchronize(e1));
        float t; cudaEventElapsedTime(&t, e0, e1); if (i >= 2 && t < ms_a) ms_a = t;
    }
    const double gb = (double)rows * DEG * F * 4 / 1e9;
    printf("{\"kernel\": \"ldg\", \"ms\": %.4f, \"GBs\": %.1f}\n", ms_a, gb / ms_a * 1e3);
    for (int box_rows = 1; box_rows <= 4; box_rows += 3) {
        CUtensorMap tmap;
        cuuint64_t gdim[2] = {(cuuint64_t)F, (cuuint64_t)N}, gstride[1] = {(cuuint64_t)F * 4};
        cuuint32_t box[2] = {(cuuint32_t)F, (cuuint32_t)box_rows}, estr[2] = {1, 1};
        CUresult cr = cuTensorMapEncodeTiled(&tmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, x, gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                             CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (cr != CUDA_SUCCESS) { printf("{\"kernel\": \"tma_gather4\", \"box_rows\": %d, \"error\": \"cuTensorMapEncodeTiled %d\"}\n", box_rows, (int)cr); continue; }
        constexpr int W = 8;
        const size_t smem = (size_t)W * 2 * DEG * F * 4 + W * 2 * 8;
        CK(cudaFuncSetAttribute(gather_tma<W>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        for (int blocks_per_sm = 1; blocks_per_sm <= 3; ++blocks_per_sm) {
            float ms_b = 1e9f; bool bad = false;
            for (int i = 0; i < 8 && !bad; ++i) {
                CK(cudaMemset(ob, 0, (size_t)rows * F * 4));
                cudaEventRecord(e0);
                gather_tma<W><<<148 * blocks_per_sm, W * 32, smem>>>(tmap, idx, ob, rows, fail);
                cudaEventRecord(e1);
                cudaError_t er = cudaEventSynchronize(e1);
                if (er != cudaSuccess) { printf("{\"kernel\": \"tma_gather4\", \"box_rows\": %d, \"error\": \"%s\"}\n", box_rows, cudaGetErrorString(er)); return 0; }
                float t; cudaEventElapsedTime(&t, e0, e1); if (i >= 2 && t < ms_b) ms_b = t;
                int hf = 0; cudaMemcpy(&hf, fail, 4, cudaMemcpyDeviceToHost);
                if (hf) { printf("{\"kernel\": \"tma_gather4\", \"box_rows\": %d, \"error\": \"mbarrier never completed\"}\n", box_rows); bad = true; cudaMemset(fail, 0, 4); }
            }
            if (bad) break;
            std::vector<float> ha(4096), hb(4096);
            cudaMemcpy(ha.data(), oa + (size_t)(rows - 32) * F, 4096 * 4, cudaMemcpyDeviceToHost);
            cudaMemcpy(hb.data(), ob + (size_t)(rows - 32) * F, 4096 * 4, cudaMemcpyDeviceToHost);
            double md = 0; for (int i = 0; i < 4096; ++i) { double d = ha[i] - hb[i]; if (d < 0) d = -d; if (d > md) md = d; }
            printf("{\"kernel\": \"tma_gather4\", \"box_rows\": %d, \"blocks_per_sm\": %d, \"warps_per_sm\": %d, \"ms\": %.4f, \"GBs\": %.1f, \"max_abs_diff_vs_ldg\": %.3g}\n",
                   box_rows, blocks_per_sm, blocks_per_sm * W, ms_b, gb / ms_b * 1e3, md);
        }
    }
    return 0;
}
