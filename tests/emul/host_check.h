// TEST INFRASTRUCTURE ONLY -- see kagnn_b200/csrc/launch.cuh.  A serial host stand-in for the handful of CUDA constructs that
// kagnn_b200/csrc/backward.cu uses, so that its kernels can be executed thread by thread under g++ to check their indexing.
#pragma once
#include <math.h>
#include <stdint.h>
#include <string.h>
#include <algorithm>
#include "../../include/kagnn_b200.h"

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __restrict__
#define __launch_bounds__(...)

struct dim3 {
    unsigned x, y, z;
    dim3(unsigned x_ = 1, unsigned y_ = 1, unsigned z_ = 1) : x(x_), y(y_), z(z_) {}
};
static dim3 threadIdx, blockIdx, blockDim, gridDim;
typedef void* cudaStream_t;
typedef int cudaError_t;
static const int cudaSuccess = 0;
static inline int cudaMemsetAsync(void* p, int v, size_t n, cudaStream_t) { memset(p, v, n); return 0; }
static inline int cudaGetLastError() { return 0; }
template <class T> static inline T atomicAdd(T* p, T v) { T o = *p; *p = o + v; return o; }
template <class T> static inline T __ldg(const T* p) { return *p; }
using std::max;
using std::min;

#define KAGNN_CUDA_TRY(expr) do { if ((expr) != cudaSuccess) return KAGNN_ECUDA; } while (0)
#define KAGNN_LAUNCH_CHECK() do { } while (0)
static inline int64_t ceil_div64(int64_t a, int64_t b) { return (a + b - 1) / b; }
static inline int pad4(int v) { return (v + 3) & ~3; }

template <class F> static inline void host_check_launch(dim3 grid, dim3 block, F&& body) {
    gridDim = grid; blockDim = block;
    for (unsigned bz = 0; bz < grid.z; ++bz) for (unsigned by = 0; by < grid.y; ++by) for (unsigned bx = 0; bx < grid.x; ++bx)
        for (unsigned ty = 0; ty < block.y; ++ty) for (unsigned tx = 0; tx < block.x; ++tx) {
            blockIdx = dim3(bx, by, bz);
            threadIdx = dim3(tx, ty, 0);
            body();
        }
}
#define KAGNN_LAUNCH(kernel, grid, block, stream, ...) host_check_launch(dim3(grid), dim3(block), [&] { kernel(__VA_ARGS__); })
