// Training-mode epilogue of a message-passing layer: x = dropout(bn(conv(x))) of
// node_classification_clean/models.py:196-198 (and graph_classification/models.py:113-116, graph_regression/models.py:113-116)
// as TWO launches instead of five framework passes: column statistics, then one pass that normalises with the batch statistics,
// applies the affine, updates the running estimates and applies the dropout mask.  The mask is never stored: it is a pure
// function of (seed, element index) -- Philox4x32-10 -- and the backward regenerates it.
//   forward   y = keep * bn(x) / (1 - p),   keep = [u(seed, r * cols + c) >= p]
//   backward  g = keep * dy / (1 - p);  dx = gamma * rstd * (g - mean(g) - xhat * mean(g * xhat)),  dgamma = sum g xhat,  dbeta = sum g
// Statistics in fp64 like kagnn_batchnorm_train_fwd / _bwd, which remain the p = 0 paths.
#include "common.cuh"

namespace {
constexpr int kTE = 256;

__device__ __forceinline__ uint2 mulhilo(uint32_t a, uint32_t b) {
    const unsigned long long p = (unsigned long long)a * b;
    return make_uint2((uint32_t)p, (uint32_t)(p >> 32));
}

// Philox4x32-10 (Salmon et al., SC'11): counter (c0..c3), key (k0, k1) -> four 32-bit words
__device__ __forceinline__ uint4 philox4x32_10(uint4 c, uint2 k) {
#pragma unroll
    for (int i = 0; i < 10; ++i) {
        const uint2 a = mulhilo(0xD2511F53u, c.x), b = mulhilo(0xCD9E8D57u, c.z);
        c = make_uint4(b.y ^ c.y ^ k.x, b.x, a.y ^ c.w ^ k.y, a.x);
        k.x += 0x9E3779B9u;
        k.y += 0xBB67AE85u;
    }
    return c;
}

// 1 / (1 - p) if element `idx` is kept, else 0
__device__ __forceinline__ float keep_scale(unsigned long long seed, long long idx, float p, float inv_keep) {
    const unsigned long long blk = (unsigned long long)idx >> 2;
    const uint4 r = philox4x32_10(make_uint4((uint32_t)blk, (uint32_t)(blk >> 32), 0u, 0u), make_uint2((uint32_t)seed, (uint32_t)(seed >> 32)));
    const int w = (int)(idx & 3);
    const uint32_t bits = w == 0 ? r.x : (w == 1 ? r.y : (w == 2 ? r.z : r.w));
    const float u = (float)(bits >> 8) * (1.0f / 16777216.0f);              // 24 uniform bits in [0, 1)
    return u >= p ? inv_keep : 0.f;
}

// sums[0..1][c] = sum x, sum x^2 (forward) -- each block owns a slab of rows, threads stride over columns
__global__ void te_stats_kernel(const float* __restrict__ x, long long ldx, long long rows, int cols, long long rows_per_block,
                                double* __restrict__ sums) {
    const long long r0 = (long long)blockIdx.x * rows_per_block, r1 = min(rows, r0 + rows_per_block);
    for (int c = threadIdx.x; c < cols; c += blockDim.x) {
        double s = 0.0, q = 0.0;
        for (long long r = r0; r < r1; ++r) {
            const double v = (double)x[r * ldx + c];
            s += v;
            q += v * v;
        }
        atomicAdd(&sums[c], s);
        atomicAdd(&sums[cols + c], q);
    }
}

__global__ void te_apply_kernel(const float* __restrict__ x, long long ldx, long long rows, int cols, const double* __restrict__ sums,
                                const float* __restrict__ weight, const float* __restrict__ bias, float eps, float momentum,
                                float* __restrict__ running_mean, float* __restrict__ running_var, float p, unsigned long long seed,
                                float* __restrict__ y, long long ldy) {
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= rows * cols) return;
    const int c = (int)(idx % cols);
    const long long r = idx / cols;
    const double mean = sums[c] / (double)rows;
    const double var = fmax(sums[cols + c] / (double)rows - mean * mean, 0.0);
    const float inv = (float)(1.0 / sqrt(var + (double)eps));
    float v = (x[r * ldx + c] - (float)mean) * inv;
    if (weight) v *= weight[c];
    if (bias) v += bias[c];
    y[r * ldy + c] = v * keep_scale(seed, idx, p, 1.0f / (1.0f - p));
    if (r == 0 && running_mean && running_var) {
        const double unbiased = rows > 1 ? var * (double)rows / (double)(rows - 1) : var;
        running_mean[c] = (1.0f - momentum) * running_mean[c] + momentum * (float)mean;
        running_var[c] = (1.0f - momentum) * running_var[c] + momentum * (float)unbiased;
    }
}

// sums[0..3][c] = sum x, sum x^2, sum g, sum g x   with g = keep * dy / (1 - p)
__global__ void te_bwd_sums_kernel(const float* __restrict__ x, long long ldx, const float* __restrict__ dy, long long ld_dy, long long rows,
                                   int cols, long long rows_per_block, float p, unsigned long long seed, double* __restrict__ sums) {
    const long long r0 = (long long)blockIdx.x * rows_per_block, r1 = min(rows, r0 + rows_per_block);
    const float inv_keep = 1.0f / (1.0f - p);
    for (int c = threadIdx.x; c < cols; c += blockDim.x) {
        double sx = 0.0, sxx = 0.0, sd = 0.0, sdx = 0.0;
        for (long long r = r0; r < r1; ++r) {
            const double v = (double)x[r * ldx + c];
            const double g = (double)(dy[r * ld_dy + c] * keep_scale(seed, r * cols + c, p, inv_keep));
            sx += v;
            sxx += v * v;
            sd += g;
            sdx += g * v;
        }
        atomicAdd(&sums[c], sx);
        atomicAdd(&sums[cols + c], sxx);
        atomicAdd(&sums[2 * cols + c], sd);
        atomicAdd(&sums[3 * cols + c], sdx);
    }
}

__global__ void te_bwd_apply_kernel(const float* __restrict__ x, long long ldx, const float* __restrict__ dy, long long ld_dy, long long rows,
                                    int cols, const double* __restrict__ sums, const float* __restrict__ weight, float eps, float p,
                                    unsigned long long seed, float* __restrict__ dx, long long ld_dx, float* __restrict__ d_weight,
                                    float* __restrict__ d_bias) {
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= rows * cols) return;
    const int c = (int)(idx % cols);
    const long long r = idx / cols;
    const double inv_n = 1.0 / (double)rows;
    const double mean = sums[c] * inv_n;
    const double var = fmax(sums[cols + c] * inv_n - mean * mean, 0.0);
    const double rstd = 1.0 / sqrt(var + (double)eps);
    const double sum_g = sums[2 * cols + c];
    const double sum_g_xhat = (sums[3 * cols + c] - mean * sum_g) * rstd;
    const double xhat = ((double)x[r * ldx + c] - mean) * rstd;
    const double gamma = weight ? (double)weight[c] : 1.0;
    const double g = (double)(dy[r * ld_dy + c] * keep_scale(seed, idx, p, 1.0f / (1.0f - p)));
    dx[r * ld_dx + c] = (float)(gamma * rstd * (g - sum_g * inv_n - xhat * sum_g_xhat * inv_n));
    if (r == 0) {
        if (d_weight) d_weight[c] = (float)sum_g_xhat;
        if (d_bias) d_bias[c] = (float)sum_g;
    }
}
}  // namespace

extern "C" size_t kagnn_bn_dropout_train_workspace(int32_t cols) { return cols > 0 ? (size_t)cols * 4 * sizeof(double) : 0; }

extern "C" int kagnn_bn_dropout_train_fwd(const float* x, int64_t ldx, int64_t rows, int32_t cols, const float* weight, const float* bias,
                                          float eps, float momentum, float* running_mean, float* running_var, float p_drop,
                                          uint64_t seed, float* y, int64_t ldy, void* workspace, size_t workspace_bytes, void* stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    if (rows <= 0 || cols <= 0 || !x || !y || ldx < cols || ldy < cols || !(p_drop >= 0.f) || !(p_drop < 1.f)) return KAGNN_EINVAL;
    if (!workspace || workspace_bytes < kagnn_bn_dropout_train_workspace(cols)) return KAGNN_EWORKSPACE;
    double* sums = static_cast<double*>(workspace);
    KAGNN_CUDA_TRY(cudaMemsetAsync(sums, 0, (size_t)cols * 2 * sizeof(double), stream));
    const long long rows_per_block = 256;
    te_stats_kernel<<<(unsigned)ceil_div64(rows, rows_per_block), 128, 0, stream>>>(x, ldx, rows, cols, rows_per_block, sums);
    KAGNN_LAUNCH_CHECK();
    te_apply_kernel<<<(unsigned)ceil_div64(rows * (int64_t)cols, kTE), kTE, 0, stream>>>(x, ldx, rows, cols, sums, weight, bias, eps, momentum,
                                                                                        running_mean, running_var, p_drop,
                                                                                        (unsigned long long)seed, y, ldy);
    KAGNN_LAUNCH_CHECK();
    return KAGNN_OK;
}

extern "C" int kagnn_bn_dropout_train_bwd(const float* x, int64_t ldx, const float* dy, int64_t ld_dy, int64_t rows, int32_t cols,
                                          const float* weight, float eps, float p_drop, uint64_t seed, float* dx, int64_t ld_dx,
                                          float* d_weight, float* d_bias, void* workspace, size_t workspace_bytes, void* stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    if (rows <= 0 || cols <= 0 || !x || !dy || !dx || ldx < cols || ld_dy < cols || ld_dx < cols || !(p_drop >= 0.f) || !(p_drop < 1.f))
        return KAGNN_EINVAL;
    if (!workspace || workspace_bytes < kagnn_bn_dropout_train_workspace(cols)) return KAGNN_EWORKSPACE;
    double* sums = static_cast<double*>(workspace);
    KAGNN_CUDA_TRY(cudaMemsetAsync(sums, 0, (size_t)cols * 4 * sizeof(double), stream));
    const long long rows_per_block = 256;
    te_bwd_sums_kernel<<<(unsigned)ceil_div64(rows, rows_per_block), 128, 0, stream>>>(x, ldx, dy, ld_dy, rows, cols, rows_per_block, p_drop,
                                                                                      (unsigned long long)seed, sums);
    KAGNN_LAUNCH_CHECK();
    te_bwd_apply_kernel<<<(unsigned)ceil_div64(rows * (int64_t)cols, kTE), kTE, 0, stream>>>(x, ldx, dy, ld_dy, rows, cols, sums, weight, eps,
                                                                                            p_drop, (unsigned long long)seed, dx, ld_dx,
                                                                                            d_weight, d_bias);
    KAGNN_LAUNCH_CHECK();
    return KAGNN_OK;
}
