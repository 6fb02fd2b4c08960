// Backward of the KAN layer and of the small model epilogues: the gradients that
// `loss.backward()` needs from the modules of kagnn_b200 so that the reference's training loops
// (node_classification_clean/utils.py:125-132, graph_classification/graph_classification_utils.py:44-54) run on them.
//
// FIRST CORRECT PATH, not a tuned one: plain fp32 CUDA-core kernels, no shared memory, no barriers, no warp intrinsics
// (which also makes them checkable thread by thread on the host, see launch.cuh).  The GEMM-shaped halves (dW = basis^T dY,
// dX = dY W^T .* basis') are the tcgen05 candidates of the next round.
//
//   forward (ekan.py:154-162)   y[n,o] = sum_i silu(x[n,i]) Wb[o,i] + sum_i sum_s B_s(x[n,i]) Ws[o,i,s] sc[o,i]
//   dX                          dx[n,i] = silu'(x) sum_o dy[n,o] Wb[o,i] + sum_s B_s'(x) sum_o dy[n,o] Ws[o,i,s] sc[o,i]
//   dW (packed layout)          dP[i][s][o] = sum_n dy[n,o] B_s(x[n,i]),   dP[i][S][o] = sum_n dy[n,o] silu(x[n,i])
//   chain through the packing   dWb = dP[.][S][.]^T, dWs = dP * sc, dsc = sum_s dP * Ws          (ekan.py:146-152)
//
// B_s' on uniform knots: B'_{j,k}(x) = (B_{j,k-1}(x) - B_{j+1,k-1}(x)) / h, evaluated on the same half-open interval the
// forward recursion selects (ekan.py:95-105), which is what autograd differentiates in the reference.
#include "launch.cuh"

namespace {
constexpr int kBwdThreads = 256;
constexpr int kMaxOrder = 4;
constexpr int kMaxSlots1 = 40;          // G + k + 1 <= 32 + 4 + 1

struct KanGeom {
    int in_f, out_f, out_pad, G, k, S;  // S = G + k spline slots; slot S = SiLU base column
    float t0, inv_h;
};

// Fractional position of x in its knot interval.  Returns false outside [t_0, t_last) and for NaN.
__device__ __forceinline__ bool locate(const KanGeom& g, float x, int* cell, float* fr) {
    const float u = (x - g.t0) * g.inv_h;
    const float fl = floorf(u);
    const bool valid = (u >= 0.0f) && (u < (float)(g.G + 2 * g.k));
    *cell = valid ? (int)fl : 0;
    *fr = u - fl;
    return valid;
}

// Local Cox-de Boor recursion: b[0..k] = the k+1 non-zero order-k bases B_{cell-k .. cell}, m[0..k-1] = the order-(k-1) ones
// B_{cell-k+1 .. cell} (needed for the derivative).
__device__ __forceinline__ void local_bases(int k, float fr, float* b, float* m) {
    for (int r = 0; r <= kMaxOrder; ++r) { b[r] = 0.f; m[r] = 0.f; }
    b[0] = 1.f;
    for (int d = 1; d <= k; ++d) {
        if (d == k)
            for (int r = 0; r < d; ++r) m[r] = b[r];
        const float inv_d = 1.0f / (float)d;
        float nb[kMaxOrder + 1];
        for (int r = 0; r <= d; ++r) {
            const float left = (r > 0) ? (fr + (float)(d - r)) * inv_d * b[r - 1] : 0.f;
            const float right = (r < d) ? ((float)(r + 1) - fr) * inv_d * b[r] : 0.f;
            nb[r] = left + right;
        }
        for (int r = 0; r <= d; ++r) b[r] = nb[r];
    }
}

__device__ __forceinline__ float sigmoid_f(float x) { return 1.0f / (1.0f + expf(-x)); }

// block = 256 rows x ONE input feature i (blockIdx.y), thread = row: the weight rows of feature i are shared by the whole block
// (a warp touches at most G + k distinct ones), every thread streams its own dy row
__global__ void kan_bwd_input_kernel(KanGeom g, const float* __restrict__ w, const float* __restrict__ x, long long ldx,
                                     const float* __restrict__ dy, long long ld_dy, long long n_rows, float* __restrict__ dx,
                                     long long ld_dx) {
    const long long n = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const int i = (int)blockIdx.y;
    if (n >= n_rows) return;
    const float xv = x[n * ldx + i];
    int cell;
    float fr;
    const bool valid = locate(g, xv, &cell, &fr);
    float b[kMaxOrder + 1], m[kMaxOrder + 1];
    local_bases(g.k, fr, b, m);
    // derivative of the r-th local basis and the weight row it meets; slots outside [0, S) or an x outside the knot range drop out
    float db[kMaxOrder + 1];
    const float* wr[kMaxOrder + 1];
    const float* wi = w + (long long)i * (g.S + 1) * g.out_pad;
    for (int r = 0; r <= kMaxOrder; ++r) {
        const int slot = cell - g.k + r;
        const bool on = valid && r <= g.k && slot >= 0 && slot < g.S;
        const float left = (r > 0) ? m[r - 1] : 0.f, right = (r < g.k) ? m[r] : 0.f;
        db[r] = on ? (left - right) * g.inv_h : 0.f;
        wr[r] = wi + (long long)(on ? slot : 0) * g.out_pad;
    }
    const float* wb = wi + (long long)g.S * g.out_pad;
    const float* dyr = dy + n * ld_dy;
    float gb = 0.f, gs[kMaxOrder + 1] = {0.f, 0.f, 0.f, 0.f, 0.f};
    for (int o = 0; o < g.out_f; ++o) {
        const float d = dyr[o];
        gb = fmaf(d, wb[o], gb);
        for (int r = 0; r <= kMaxOrder; ++r) gs[r] = fmaf(d, wr[r][o], gs[r]);
    }
    const float s = sigmoid_f(xv);
    float acc = gb * (s * (1.0f + xv * (1.0f - s)));            // d/dx [x sigmoid(x)]
    for (int r = 0; r <= kMaxOrder; ++r) acc = fmaf(db[r], gs[r], acc);
    dx[n * ld_dx + i] = acc;
}

// block = one input feature i x one slab of rows, thread = output column o; every thread keeps its own column of the
// (S+1) x out gradient block in local memory and adds it to HBM once at the end
__global__ void kan_bwd_weights_kernel(KanGeom g, const float* __restrict__ x, long long ldx, const float* __restrict__ dy,
                                       long long ld_dy, long long n_rows, long long rows_per_block, float* __restrict__ dP) {
    const int i = (int)(blockIdx.x % g.in_f);
    const int o = (int)(blockIdx.x / g.in_f) * (int)blockDim.x + (int)threadIdx.x;
    if (o >= g.out_f) return;
    const long long r0 = (long long)blockIdx.y * rows_per_block;
    const long long r1 = min(n_rows, r0 + rows_per_block);
    float acc[kMaxSlots1];
    for (int c = 0; c <= g.S; ++c) acc[c] = 0.f;
    for (long long n = r0; n < r1; ++n) {
        const float xv = x[n * ldx + i];
        const float d = dy[n * ld_dy + o];
        int cell;
        float fr;
        const bool valid = locate(g, xv, &cell, &fr);
        acc[g.S] = fmaf(d, xv * sigmoid_f(xv), acc[g.S]);
        if (valid) {
            float b[kMaxOrder + 1], m[kMaxOrder + 1];
            local_bases(g.k, fr, b, m);
            for (int r = 0; r <= g.k; ++r) {
                const int slot = cell - g.k + r;
                if (slot >= 0 && slot < g.S) acc[slot] = fmaf(d, b[r], acc[slot]);
            }
        }
    }
    float* out = dP + (long long)i * (g.S + 1) * g.out_pad + o;
    for (int c = 0; c <= g.S; ++c)
        if (acc[c] != 0.f) atomicAdd(out + (long long)c * g.out_pad, acc[c]);
}

// thread = (o, i)
__global__ void kan_unpack_grads_kernel(const float* __restrict__ dP, const float* __restrict__ spline_w,
                                        const float* __restrict__ scaler, int in_f, int out_f, int S, int out_pad,
                                        float* __restrict__ d_base, float* __restrict__ d_spline, float* __restrict__ d_scaler) {
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (long long)in_f * out_f) return;
    const int i = (int)(idx % in_f), o = (int)(idx / in_f);
    const float* p = dP + (long long)i * (S + 1) * out_pad + o;
    const long long oi = (long long)o * in_f + i;
    const float sc = scaler ? scaler[oi] : 1.0f;
    float ds = 0.f;
    for (int s = 0; s < S; ++s) {
        const float gsl = p[(long long)s * out_pad];
        d_spline[oi * S + s] = gsl * sc;
        ds = fmaf(gsl, spline_w[oi * S + s], ds);
    }
    if (d_scaler) d_scaler[oi] = ds;
    if (d_base) d_base[oi] = p[(long long)S * out_pad];
}

// ---- BatchNorm1d (training mode) ------------------------------------------------------------------------------------
// sums[0..3][c] = sum x, sum x^2, sum dy, sum dy*x over the rows, fp64
__global__ void bn_bwd_sums_kernel(const float* __restrict__ x, long long ldx, const float* __restrict__ dy, long long ld_dy,
                                   long long rows, int cols, long long rows_per_block, double* __restrict__ sums) {
    const long long r0 = (long long)blockIdx.x * rows_per_block, r1 = min(rows, r0 + rows_per_block);
    for (int c = (int)threadIdx.x; c < cols; c += (int)blockDim.x) {
        double sx = 0.0, sxx = 0.0, sd = 0.0, sdx = 0.0;
        for (long long r = r0; r < r1; ++r) {
            const double v = (double)x[r * ldx + c], d = (double)dy[r * ld_dy + c];
            sx += v;
            sxx += v * v;
            sd += d;
            sdx += d * v;
        }
        atomicAdd(&sums[c], sx);
        atomicAdd(&sums[cols + c], sxx);
        atomicAdd(&sums[2 * cols + c], sd);
        atomicAdd(&sums[3 * cols + c], sdx);
    }
}

// dx = gamma * rstd * (dy - mean(dy) - xhat * mean(dy * xhat)),  dgamma = sum dy * xhat,  dbeta = sum dy
__global__ void bn_bwd_apply_kernel(const float* __restrict__ x, long long ldx, const float* __restrict__ dy, long long ld_dy,
                                    long long rows, int cols, const double* __restrict__ sums, const float* __restrict__ weight,
                                    float eps, float* __restrict__ dx, long long ld_dx, float* __restrict__ d_weight,
                                    float* __restrict__ d_bias) {
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= rows * cols) return;
    const int c = (int)(idx % cols);
    const long long r = idx / cols;
    const double inv_n = 1.0 / (double)rows;
    const double mean = sums[c] * inv_n;
    const double var = fmax(sums[cols + c] * inv_n - mean * mean, 0.0);
    const double rstd = 1.0 / sqrt(var + (double)eps);
    const double sum_dy = sums[2 * cols + c];
    const double sum_dy_xhat = (sums[3 * cols + c] - mean * sum_dy) * rstd;
    const double xhat = ((double)x[r * ldx + c] - mean) * rstd;
    const double gamma = weight ? (double)weight[c] : 1.0;
    dx[r * ld_dx + c] = (float)(gamma * rstd * ((double)dy[r * ld_dy + c] - sum_dy * inv_n - xhat * sum_dy_xhat * inv_n));
    if (r == 0) {
        if (d_weight) d_weight[c] = (float)sum_dy_xhat;
        if (d_bias) d_bias[c] = (float)sum_dy;
    }
}

// out[c] += sum over a slab of rows (GCNConv bias gradient)
__global__ void column_sums_kernel(const float* __restrict__ x, long long ldx, long long rows, int cols, long long rows_per_block,
                                   float* __restrict__ out) {
    const long long r0 = (long long)blockIdx.x * rows_per_block, r1 = min(rows, r0 + rows_per_block);
    for (int c = (int)threadIdx.x; c < cols; c += (int)blockDim.x) {
        double s = 0.0;
        for (long long r = r0; r < r1; ++r) s += (double)x[r * ldx + c];
        atomicAdd(&out[c], (float)s);
    }
}

// y = log_softmax(x):  dx = dy - exp(y) * sum_c dy      (thread = row)
__global__ void log_softmax_bwd_kernel(const float* __restrict__ y, long long ldy, const float* __restrict__ dy, long long ld_dy,
                                       long long rows, int cols, float* __restrict__ dx, long long ld_dx) {
    const long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= rows) return;
    float s = 0.f;
    for (int c = 0; c < cols; ++c) s += dy[r * ld_dy + c];
    for (int c = 0; c < cols; ++c) dx[r * ld_dx + c] = dy[r * ld_dy + c] - expf(y[r * ldy + c]) * s;
}

// global_add_pool / global_mean_pool backward: dx[n,:] = d_pooled[batch[n],:] (/ segment length)
__global__ void segment_pool_bwd_kernel(const float* __restrict__ dp, long long ld_dp, const int32_t* __restrict__ ptr,
                                        const int64_t* __restrict__ batch, long long rows, int cols, int mean,
                                        float* __restrict__ dx, long long ld_dx) {
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= rows * cols) return;
    const int c = (int)(idx % cols);
    const long long r = idx / cols;
    const long long gph = batch[r];
    float v = dp[gph * ld_dp + c];
    if (mean) v /= (float)max(ptr[gph + 1] - ptr[gph], 1);
    dx[r * ld_dx + c] = v;
}

// y = silu(x):  dx = dy * sigmoid(x) * (1 + x * (1 - sigmoid(x)))
__global__ void silu_bwd_kernel(const float* __restrict__ x, long long ldx, const float* __restrict__ dy, long long ld_dy,
                                long long rows, int cols, float* __restrict__ dx, long long ld_dx) {
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= rows * cols) return;
    const int c = (int)(idx % cols);
    const long long r = idx / cols;
    const float xv = x[r * ldx + c];
    const float s = sigmoid_f(xv);
    dx[r * ld_dx + c] = dy[r * ld_dy + c] * (s * (1.0f + xv * (1.0f - s)));
}

__global__ void silu_fwd_kernel(const float* __restrict__ x, long long ldx, long long rows, int cols, float* __restrict__ y,
                                long long ldy) {
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= rows * cols) return;
    const int c = (int)(idx % cols);
    const long long r = idx / cols;
    const float xv = x[r * ldx + c];
    y[r * ldy + c] = xv * sigmoid_f(xv);
}

int geometry(const KagnnKanLayer* L, KanGeom* g) {
    if (!L || !L->packed_w) return KAGNN_EINVAL;
    if (L->basis != KAGNN_BASIS_BSPLINE) return KAGNN_EUNSUPPORTED;           // FastKAN backward: not built yet
    if (L->in_features <= 0 || L->out_features <= 0 || L->grid_size < 1 || L->grid_size > 32) return KAGNN_EINVAL;
    if (L->spline_order < 1 || L->spline_order > kMaxOrder) return KAGNN_EUNSUPPORTED;
    if (!(L->h > 0.f)) return KAGNN_EINVAL;
    g->in_f = L->in_features;
    g->out_f = L->out_features;
    g->out_pad = pad4(L->out_features);
    g->G = L->grid_size;
    g->k = L->spline_order;
    g->S = L->grid_size + L->spline_order;
    g->t0 = L->t0;
    g->inv_h = 1.0f / L->h;
    return KAGNN_OK;
}
}  // namespace

extern "C" int kagnn_kan_bwd_input(const KagnnKanLayer* layer, const float* x, int64_t ldx, const float* dy, int64_t ld_dy,
                                   int64_t num_rows, float* dx, int64_t ld_dx, void* stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    KanGeom g;
    const int rc = geometry(layer, &g);
    if (rc != KAGNN_OK) return rc;
    if (num_rows < 0 || (num_rows > 0 && (!x || !dy || !dx)) || ldx < g.in_f || ld_dy < g.out_f || ld_dx < g.in_f) return KAGNN_EINVAL;
    if (num_rows == 0) return KAGNN_OK;
    KAGNN_TRY_TILED(kagnn_kan_bwd_input_tc(layer, x, ldx, dy, ld_dy, num_rows, dx, ld_dx, stream));
    KAGNN_TRY_TILED(kagnn_kan_bwd_input_tiled(layer, x, ldx, dy, ld_dy, num_rows, dx, ld_dx, stream));
    if (g.in_f > 65535) return KAGNN_EUNSUPPORTED;                                  // gridDim.y
    KAGNN_LAUNCH(kan_bwd_input_kernel, dim3((unsigned)ceil_div64(num_rows, kBwdThreads), (unsigned)g.in_f, 1),
                 dim3((unsigned)kBwdThreads, 1, 1), stream, g, layer->packed_w, x, (long long)ldx, dy, (long long)ld_dy,
                 (long long)num_rows, dx, (long long)ld_dx);
    KAGNN_LAUNCH_CHECK();
    return KAGNN_OK;
}

extern "C" int kagnn_kan_bwd_weights(const KagnnKanLayer* layer, const float* x, int64_t ldx, const float* dy, int64_t ld_dy,
                                     int64_t num_rows, float* d_packed, void* stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    KanGeom g;
    const int rc = geometry(layer, &g);
    if (rc != KAGNN_OK) return rc;
    if (num_rows < 0 || !d_packed || (num_rows > 0 && (!x || !dy)) || ldx < g.in_f || ld_dy < g.out_f) return KAGNN_EINVAL;
    if (num_rows > 0) KAGNN_TRY_TILED(kagnn_kan_bwd_weights_tc(layer, x, ldx, dy, ld_dy, num_rows, d_packed, stream));
    if (num_rows > 0) KAGNN_TRY_TILED(kagnn_kan_bwd_weights_tiled(layer, x, ldx, dy, ld_dy, num_rows, d_packed, stream));
    KAGNN_CUDA_TRY(cudaMemsetAsync(d_packed, 0, sizeof(float) * (size_t)g.in_f * (size_t)(g.S + 1) * (size_t)g.out_pad, stream));
    if (num_rows == 0) return KAGNN_OK;
    const int threads = g.out_f >= 256 ? 256 : ((g.out_f + 31) / 32) * 32;
    const int o_blocks = (g.out_f + threads - 1) / threads;
    // enough row slabs to fill the machine a few times over, at most 64 atomics per gradient element
    int64_t slabs = ceil_div64(num_rows, 512);
    if (slabs > 64) slabs = 64;
    const int64_t rows_per_block = ceil_div64(num_rows, slabs);
    slabs = ceil_div64(num_rows, rows_per_block);
    KAGNN_LAUNCH(kan_bwd_weights_kernel, dim3((unsigned)(g.in_f * o_blocks), (unsigned)slabs, 1), dim3((unsigned)threads, 1, 1), stream,
                 g, x, (long long)ldx, dy, (long long)ld_dy, (long long)num_rows, (long long)rows_per_block, d_packed);
    KAGNN_LAUNCH_CHECK();
    return KAGNN_OK;
}

extern "C" int kagnn_kan_unpack_weight_grads(const float* d_packed, const float* spline_w, const float* scaler, int32_t in_f,
                                             int32_t out_f, int32_t slots, float* d_base_w, float* d_spline_w, float* d_scaler,
                                             void* stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    if (in_f <= 0 || out_f <= 0 || slots <= 0 || !d_packed || !spline_w || !d_spline_w) return KAGNN_EINVAL;
    if (d_scaler && !scaler) return KAGNN_EINVAL;
    const int64_t total = (int64_t)in_f * out_f;
    KAGNN_LAUNCH(kan_unpack_grads_kernel, (unsigned)ceil_div64(total, kBwdThreads), kBwdThreads, stream, d_packed, spline_w, scaler,
                 (int)in_f, (int)out_f, (int)slots, pad4(out_f), d_base_w, d_spline_w, d_scaler);
    KAGNN_LAUNCH_CHECK();
    return KAGNN_OK;
}

extern "C" size_t kagnn_batchnorm_bwd_workspace(int32_t num_cols) { return num_cols > 0 ? (size_t)num_cols * 4 * sizeof(double) : 0; }

extern "C" int kagnn_batchnorm_train_bwd(const float* x, int64_t ldx, const float* dy, int64_t ld_dy, int64_t num_rows,
                                         int32_t num_cols, const float* weight, float eps, float* dx, int64_t ld_dx,
                                         float* d_weight, float* d_bias, void* workspace, size_t workspace_bytes, void* stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    if (num_rows <= 0 || num_cols <= 0 || !x || !dy || !dx || ldx < num_cols || ld_dy < num_cols || ld_dx < num_cols) return KAGNN_EINVAL;
    if (!workspace || workspace_bytes < kagnn_batchnorm_bwd_workspace(num_cols)) return KAGNN_EWORKSPACE;
    double* sums = static_cast<double*>(workspace);
    KAGNN_CUDA_TRY(cudaMemsetAsync(sums, 0, (size_t)num_cols * 4 * sizeof(double), stream));
    KAGNN_TRY_TILED(kagnn_batchnorm_train_bwd_fast(x, ldx, dy, ld_dy, num_rows, num_cols, weight, eps, dx, ld_dx, d_weight, d_bias, sums, stream));
    const int64_t rows_per_block = 256;
    KAGNN_LAUNCH(bn_bwd_sums_kernel, (unsigned)ceil_div64(num_rows, rows_per_block), 128, stream, x, (long long)ldx, dy,
                 (long long)ld_dy, (long long)num_rows, (int)num_cols, (long long)rows_per_block, sums);
    KAGNN_LAUNCH_CHECK();
    KAGNN_LAUNCH(bn_bwd_apply_kernel, (unsigned)ceil_div64(num_rows * (int64_t)num_cols, kBwdThreads), kBwdThreads, stream, x,
                 (long long)ldx, dy, (long long)ld_dy, (long long)num_rows, (int)num_cols, (const double*)sums, weight, eps, dx,
                 (long long)ld_dx, d_weight, d_bias);
    KAGNN_LAUNCH_CHECK();
    return KAGNN_OK;
}

extern "C" int kagnn_column_sums(const float* x, int64_t ldx, int64_t num_rows, int32_t num_cols, float* out, void* stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    if (num_rows < 0 || num_cols <= 0 || !out || (num_rows > 0 && !x) || ldx < num_cols) return KAGNN_EINVAL;
    KAGNN_CUDA_TRY(cudaMemsetAsync(out, 0, (size_t)num_cols * sizeof(float), stream));
    if (num_rows == 0) return KAGNN_OK;
    const int64_t rows_per_block = 256;
    KAGNN_LAUNCH(column_sums_kernel, (unsigned)ceil_div64(num_rows, rows_per_block), 128, stream, x, (long long)ldx,
                 (long long)num_rows, (int)num_cols, (long long)rows_per_block, out);
    KAGNN_LAUNCH_CHECK();
    return KAGNN_OK;
}

extern "C" int kagnn_log_softmax_bwd(const float* y, int64_t ldy, const float* dy, int64_t ld_dy, int64_t rows, int32_t cols,
                                     float* dx, int64_t ld_dx, void* stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    if (rows < 0 || cols <= 0 || (rows > 0 && (!y || !dy || !dx)) || ldy < cols || ld_dy < cols || ld_dx < cols) return KAGNN_EINVAL;
    if (rows == 0) return KAGNN_OK;
    KAGNN_LAUNCH(log_softmax_bwd_kernel, (unsigned)ceil_div64(rows, 128), 128, stream, y, (long long)ldy, dy, (long long)ld_dy,
                 (long long)rows, (int)cols, dx, (long long)ld_dx);
    KAGNN_LAUNCH_CHECK();
    return KAGNN_OK;
}

extern "C" int kagnn_segment_pool_bwd(const float* d_pooled, int64_t ld_dp, const int32_t* segment_ptr, const int64_t* batch,
                                      int64_t num_rows, int32_t num_cols, int32_t mean, float* dx, int64_t ld_dx, void* stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    if (num_rows < 0 || num_cols <= 0 || ld_dp < num_cols || ld_dx < num_cols) return KAGNN_EINVAL;
    if (num_rows == 0) return KAGNN_OK;
    if (!d_pooled || !segment_ptr || !batch || !dx) return KAGNN_EINVAL;
    KAGNN_LAUNCH(segment_pool_bwd_kernel, (unsigned)ceil_div64(num_rows * (int64_t)num_cols, kBwdThreads), kBwdThreads, stream,
                 d_pooled, (long long)ld_dp, segment_ptr, batch, (long long)num_rows, (int)num_cols, (int)mean, dx, (long long)ld_dx);
    KAGNN_LAUNCH_CHECK();
    return KAGNN_OK;
}

extern "C" int kagnn_silu_bwd(const float* x, int64_t ldx, const float* dy, int64_t ld_dy, int64_t rows, int32_t cols, float* dx,
                              int64_t ld_dx, void* stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    if (rows < 0 || cols <= 0 || (rows > 0 && (!x || !dy || !dx)) || ldx < cols || ld_dy < cols || ld_dx < cols) return KAGNN_EINVAL;
    if (rows == 0) return KAGNN_OK;
    KAGNN_LAUNCH(silu_bwd_kernel, (unsigned)ceil_div64(rows * (int64_t)cols, kBwdThreads), kBwdThreads, stream, x, (long long)ldx, dy,
                 (long long)ld_dy, (long long)rows, (int)cols, dx, (long long)ld_dx);
    KAGNN_LAUNCH_CHECK();
    return KAGNN_OK;
}

extern "C" int kagnn_silu_fwd(const float* x, int64_t ldx, int64_t rows, int32_t cols, float* y, int64_t ldy, void* stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    if (rows < 0 || cols <= 0 || (rows > 0 && (!x || !y)) || ldx < cols || ldy < cols) return KAGNN_EINVAL;
    if (rows == 0) return KAGNN_OK;
    KAGNN_LAUNCH(silu_fwd_kernel, (unsigned)ceil_div64(rows * (int64_t)cols, kBwdThreads), kBwdThreads, stream, x, (long long)ldx,
                 (long long)rows, (int)cols, y, (long long)ldy);
    KAGNN_LAUNCH_CHECK();
    return KAGNN_OK;
}

// =====================================================================================================================
// FastKAN layer (fastkan.py:76-85):  z = LayerNorm(x),  y = sum_i sum_g phi_g(z_i) Ws[o, i*G+g] + sum_i silu(x_i) Wb[o,i] + bb[o],
// phi_g(z) = exp(-((z - c_g) / den)^2).  Gradients:
//   dz[n,i]  = sum_g phi_g'(z) sum_o dy[n,o] Ws[o,i,g],   phi_g' = -2 (z - c_g) / den^2 * phi_g
//   dxb[n,i] = silu'(x) sum_o dy[n,o] Wb[o,i]                                (the base branch sees the RAW x)
//   dP[i][g][o] = sum_n dy[n,o] phi_g(z[n,i]),  dP[i][G][o] = sum_n dy[n,o] silu(x[n,i])      (packed layout, no scaler)
//   LayerNorm: xhat = (x - mean) rstd, z = xhat gamma + beta;  dgamma = sum_n dz xhat, dbeta = sum_n dz,
//              dx = rstd (dz gamma - mean_i(dz gamma) - xhat mean_i(dz gamma xhat)) + dxb
// =====================================================================================================================
namespace {
constexpr int kMaxGrids = 32;

struct RbfGeom {
    int in_f, out_f, out_pad, G;
    float c0, step, inv_den;
    const float* ln_w;      // NULL = no LayerNorm
    const float* ln_b;
};

__device__ __forceinline__ float rbf_input(const RbfGeom& g, const float* __restrict__ stats, long long n, int i, float xv) {
    if (!stats) return xv;
    float z = (xv - stats[2 * n]) * stats[2 * n + 1];
    if (g.ln_w) z *= g.ln_w[i];
    if (g.ln_b) z += g.ln_b[i];
    return z;
}

// block = 256 rows x one input feature (blockIdx.y), thread = row
__global__ void rbf_bwd_input_kernel(RbfGeom g, const float* __restrict__ w, const float* __restrict__ x, long long ldx,
                                     const float* __restrict__ stats, const float* __restrict__ dy, long long ld_dy,
                                     long long n_rows, float* __restrict__ dz, long long ld_dz, float* __restrict__ dxb,
                                     long long ld_dxb) {
    const long long n = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const int i = (int)blockIdx.y;
    if (n >= n_rows) return;
    const float xv = x[n * ldx + i];
    const float z = rbf_input(g, stats, n, i, xv);
    const float* wi = w + (long long)i * (g.G + 1) * g.out_pad;
    const float* wb = wi + (long long)g.G * g.out_pad;
    const float* dyr = dy + n * ld_dy;
    float gs[kMaxGrids];
    for (int q = 0; q < g.G; ++q) gs[q] = 0.f;
    float gb = 0.f;
    for (int o = 0; o < g.out_f; ++o) {
        const float d = dyr[o];
        gb = fmaf(d, wb[o], gb);
        for (int q = 0; q < g.G; ++q) gs[q] = fmaf(d, wi[(long long)q * g.out_pad + o], gs[q]);
    }
    float acc = 0.f;
    for (int q = 0; q < g.G; ++q) {
        const float t = (z - (g.c0 + (float)q * g.step)) * g.inv_den;
        acc = fmaf(-2.0f * t * g.inv_den * expf(-t * t), gs[q], acc);
    }
    const float s = sigmoid_f(xv);
    const float base = gb * (s * (1.0f + xv * (1.0f - s)));
    if (dxb) {
        dz[n * ld_dz + i] = acc;
        dxb[n * ld_dxb + i] = base;
    } else {
        dz[n * ld_dz + i] = acc + base;          // no LayerNorm: this is the complete input gradient
    }
}

// block = one input feature x one slab of rows, thread = output column
__global__ void rbf_bwd_weights_kernel(RbfGeom g, const float* __restrict__ x, long long ldx, const float* __restrict__ stats,
                                       const float* __restrict__ dy, long long ld_dy, long long n_rows, long long rows_per_block,
                                       float* __restrict__ dP) {
    const int i = (int)(blockIdx.x % g.in_f);
    const int o = (int)(blockIdx.x / g.in_f) * (int)blockDim.x + (int)threadIdx.x;
    if (o >= g.out_f) return;
    const long long r0 = (long long)blockIdx.y * rows_per_block;
    const long long r1 = min(n_rows, r0 + rows_per_block);
    float acc[kMaxGrids + 1];
    for (int q = 0; q <= g.G; ++q) acc[q] = 0.f;
    for (long long n = r0; n < r1; ++n) {
        const float xv = x[n * ldx + i];
        const float z = rbf_input(g, stats, n, i, xv);
        const float d = dy[n * ld_dy + o];
        for (int q = 0; q < g.G; ++q) {
            const float t = (z - (g.c0 + (float)q * g.step)) * g.inv_den;
            acc[q] = fmaf(d, expf(-t * t), acc[q]);
        }
        acc[g.G] = fmaf(d, xv * sigmoid_f(xv), acc[g.G]);
    }
    float* out = dP + (long long)i * (g.G + 1) * g.out_pad + o;
    for (int q = 0; q <= g.G; ++q)
        if (acc[q] != 0.f) atomicAdd(out + (long long)q * g.out_pad, acc[q]);
}

// thread = row: dx = rstd (dz gamma - mean(dz gamma) - xhat mean(dz gamma xhat)) + dxb
__global__ void layernorm_bwd_rows_kernel(const float* __restrict__ x, long long ldx, const float* __restrict__ stats,
                                          const float* __restrict__ ln_w, const float* __restrict__ dz, long long ld_dz,
                                          const float* __restrict__ dxb, long long ld_dxb, long long n_rows, int cols,
                                          float* __restrict__ dx, long long ld_dx) {
    const long long n = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= n_rows) return;
    const float mean = stats[2 * n], rstd = stats[2 * n + 1];
    float s1 = 0.f, s2 = 0.f;
    for (int i = 0; i < cols; ++i) {
        const float gz = dz[n * ld_dz + i] * (ln_w ? ln_w[i] : 1.0f);
        const float xh = (x[n * ldx + i] - mean) * rstd;
        s1 += gz;
        s2 = fmaf(gz, xh, s2);
    }
    const float inv_f = 1.0f / (float)cols;
    for (int i = 0; i < cols; ++i) {
        const float gz = dz[n * ld_dz + i] * (ln_w ? ln_w[i] : 1.0f);
        const float xh = (x[n * ldx + i] - mean) * rstd;
        float v = rstd * (gz - s1 * inv_f - xh * s2 * inv_f);
        if (dxb) v += dxb[n * ld_dxb + i];
        dx[n * ld_dx + i] = v;
    }
}

// dgamma[i] += sum over a slab of rows dz xhat, dbeta[i] += sum dz     (thread = column)
__global__ void layernorm_bwd_params_kernel(const float* __restrict__ x, long long ldx, const float* __restrict__ stats,
                                            const float* __restrict__ dz, long long ld_dz, long long n_rows, int cols,
                                            long long rows_per_block, float* __restrict__ d_w, float* __restrict__ d_b) {
    const long long r0 = (long long)blockIdx.x * rows_per_block, r1 = min(n_rows, r0 + rows_per_block);
    for (int c = (int)threadIdx.x; c < cols; c += (int)blockDim.x) {
        double sw = 0.0, sb = 0.0;
        for (long long r = r0; r < r1; ++r) {
            const double d = (double)dz[r * ld_dz + c];
            sw += d * (double)((x[r * ldx + c] - stats[2 * r]) * stats[2 * r + 1]);
            sb += d;
        }
        if (d_w) atomicAdd(&d_w[c], (float)sw);
        if (d_b) atomicAdd(&d_b[c], (float)sb);
    }
}

int rbf_geometry(const KagnnKanLayer* L, const float* stats, RbfGeom* g) {
    if (!L || !L->packed_w) return KAGNN_EINVAL;
    if (L->basis != KAGNN_BASIS_RBF) return KAGNN_EINVAL;
    if (L->in_features <= 0 || L->out_features <= 0 || L->grid_size < 1) return KAGNN_EINVAL;
    if (L->grid_size > kMaxGrids) return KAGNN_EUNSUPPORTED;
    if ((L->ln_weight || L->ln_bias) && !stats) return KAGNN_EINVAL;          // a LayerNorm layer needs its row statistics
    g->in_f = L->in_features;
    g->out_f = L->out_features;
    g->out_pad = pad4(L->out_features);
    g->G = L->grid_size;
    g->c0 = L->t0;
    g->step = L->h;
    g->inv_den = L->inv_denominator;
    g->ln_w = stats ? L->ln_weight : nullptr;
    g->ln_b = stats ? L->ln_bias : nullptr;
    return KAGNN_OK;
}
}  // namespace

extern "C" int kagnn_rbf_bwd_input(const KagnnKanLayer* layer, const float* x, int64_t ldx, const float* ln_stats, const float* dy,
                                   int64_t ld_dy, int64_t num_rows, float* dz, int64_t ld_dz, float* dx_base, int64_t ld_dxb,
                                   void* stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    RbfGeom g;
    const int rc = rbf_geometry(layer, ln_stats, &g);
    if (rc != KAGNN_OK) return rc;
    if (num_rows < 0 || (num_rows > 0 && (!x || !dy || !dz)) || ldx < g.in_f || ld_dy < g.out_f || ld_dz < g.in_f) return KAGNN_EINVAL;
    if (ln_stats && (!dx_base || ld_dxb < g.in_f)) return KAGNN_EINVAL;       // with a LayerNorm the two branches stay separate
    if (num_rows == 0) return KAGNN_OK;
    KAGNN_TRY_TILED(kagnn_rbf_bwd_input_tc(layer, x, ldx, ln_stats, dy, ld_dy, num_rows, dz, ld_dz, dx_base, ld_dxb, stream));
    if (g.in_f > 65535) return KAGNN_EUNSUPPORTED;
    KAGNN_LAUNCH(rbf_bwd_input_kernel, dim3((unsigned)ceil_div64(num_rows, kBwdThreads), (unsigned)g.in_f, 1),
                 dim3((unsigned)kBwdThreads, 1, 1), stream, g, layer->packed_w, x, (long long)ldx, ln_stats, dy, (long long)ld_dy,
                 (long long)num_rows, dz, (long long)ld_dz, ln_stats ? dx_base : (float*)nullptr, (long long)ld_dxb);
    KAGNN_LAUNCH_CHECK();
    return KAGNN_OK;
}

extern "C" int kagnn_rbf_bwd_weights(const KagnnKanLayer* layer, const float* x, int64_t ldx, const float* ln_stats, const float* dy,
                                     int64_t ld_dy, int64_t num_rows, float* d_packed, void* stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    RbfGeom g;
    const int rc = rbf_geometry(layer, ln_stats, &g);
    if (rc != KAGNN_OK) return rc;
    if (num_rows < 0 || !d_packed || (num_rows > 0 && (!x || !dy)) || ldx < g.in_f || ld_dy < g.out_f) return KAGNN_EINVAL;
    if (num_rows > 0) KAGNN_TRY_TILED(kagnn_rbf_bwd_weights_tc(layer, x, ldx, ln_stats, dy, ld_dy, num_rows, d_packed, stream));
    KAGNN_CUDA_TRY(cudaMemsetAsync(d_packed, 0, sizeof(float) * (size_t)g.in_f * (size_t)(g.G + 1) * (size_t)g.out_pad, stream));
    if (num_rows == 0) return KAGNN_OK;
    const int threads = g.out_f >= 256 ? 256 : ((g.out_f + 31) / 32) * 32;
    const int o_blocks = (g.out_f + threads - 1) / threads;
    int64_t slabs = ceil_div64(num_rows, 512);
    if (slabs > 64) slabs = 64;
    const int64_t rows_per_block = ceil_div64(num_rows, slabs);
    slabs = ceil_div64(num_rows, rows_per_block);
    KAGNN_LAUNCH(rbf_bwd_weights_kernel, dim3((unsigned)(g.in_f * o_blocks), (unsigned)slabs, 1), dim3((unsigned)threads, 1, 1), stream,
                 g, x, (long long)ldx, ln_stats, dy, (long long)ld_dy, (long long)num_rows, (long long)rows_per_block, d_packed);
    KAGNN_LAUNCH_CHECK();
    return KAGNN_OK;
}

extern "C" int kagnn_layernorm_bwd(const float* x, int64_t ldx, const float* ln_stats, const float* ln_weight, const float* dz,
                                   int64_t ld_dz, const float* dx_base, int64_t ld_dxb, int64_t num_rows, int32_t num_cols,
                                   float* dx, int64_t ld_dx, float* d_weight, float* d_bias, void* stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    if (num_rows < 0 || num_cols <= 0 || ldx < num_cols || ld_dz < num_cols || ld_dx < num_cols) return KAGNN_EINVAL;
    if (dx_base && ld_dxb < num_cols) return KAGNN_EINVAL;
    if (d_weight) KAGNN_CUDA_TRY(cudaMemsetAsync(d_weight, 0, (size_t)num_cols * sizeof(float), stream));
    if (d_bias) KAGNN_CUDA_TRY(cudaMemsetAsync(d_bias, 0, (size_t)num_cols * sizeof(float), stream));
    if (num_rows == 0) return KAGNN_OK;
    if (!x || !ln_stats || !dz || !dx) return KAGNN_EINVAL;
    KAGNN_TRY_TILED(kagnn_layernorm_bwd_fast(x, ldx, ln_stats, ln_weight, dz, ld_dz, dx_base, ld_dxb, num_rows, num_cols, dx, ld_dx, d_weight,
                                             d_bias, stream));
    KAGNN_LAUNCH(layernorm_bwd_rows_kernel, (unsigned)ceil_div64(num_rows, 128), 128, stream, x, (long long)ldx, ln_stats, ln_weight,
                 dz, (long long)ld_dz, dx_base, (long long)ld_dxb, (long long)num_rows, (int)num_cols, dx, (long long)ld_dx);
    KAGNN_LAUNCH_CHECK();
    if (d_weight || d_bias) {
        const int64_t rows_per_block = 256;
        KAGNN_LAUNCH(layernorm_bwd_params_kernel, (unsigned)ceil_div64(num_rows, rows_per_block), 128, stream, x, (long long)ldx,
                     ln_stats, dz, (long long)ld_dz, (long long)num_rows, (int)num_cols, (long long)rows_per_block, d_weight, d_bias);
        KAGNN_LAUNCH_CHECK();
    }
    return KAGNN_OK;
}

// =====================================================================================================================
// GINE aggregation (PyG GINEConv as used at graph_regression/models.py:98):
//   a_i = self_scale x_i + sum_{e: dst_e = i} relu(x_{src_e} + ef_e)
//   dx_j = self_scale da_j + sum_{e: src_e = j} [x_j + ef_e > 0] da_{dst_e},      d ef_e = [x_{src_e} + ef_e > 0] da_{dst_e}
// Works on the COO edge list (edge features in COO order, one row per edge); float atomics into dx.
// =====================================================================================================================
namespace {
__global__ void scale_rows_kernel(const float* __restrict__ a, long long lda, long long rows, int cols, float s, float* __restrict__ y,
                                  long long ldy) {
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= rows * cols) return;
    const int c = (int)(idx % cols);
    const long long r = idx / cols;
    y[r * ldy + c] = s * a[r * lda + c];
}

// thread = (edge e, column c)
__global__ void gine_bwd_edges_kernel(const float* __restrict__ x, long long ldx, const float* __restrict__ ef, long long ld_e,
                                      const int64_t* __restrict__ src, const int64_t* __restrict__ dst, long long n_edges, int cols,
                                      const float* __restrict__ da, long long ld_da, float* __restrict__ dx, long long ld_dx,
                                      float* __restrict__ d_ef, long long ld_de, long long n_nodes) {
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= n_edges * cols) return;
    const int c = (int)(idx % cols);
    const long long e = idx / cols;
    const long long j = src[e], i = dst[e];
    if (j < 0 || j >= n_nodes || i < 0 || i >= n_nodes) {     // the forward's CSR build has flagged this edge_index (IndexError): no stray access
        if (d_ef) d_ef[e * ld_de + c] = 0.f;
        return;
    }
    const float m = x[j * ldx + c] + ef[e * ld_e + c];
    const float g = m > 0.f ? da[i * ld_da + c] : 0.f;
    if (d_ef) d_ef[e * ld_de + c] = g;
    if (g != 0.f) atomicAdd(&dx[j * ld_dx + c], g);
}
}  // namespace

extern "C" int kagnn_gine_bwd(const float* x, int64_t ldx, const float* edge_feat, int64_t ld_edge, const int64_t* edge_index,
                              int64_t num_edges, int64_t num_nodes, int32_t num_cols, const float* da, int64_t ld_da,
                              float self_scale, float* dx, int64_t ld_dx, float* d_edge_feat, int64_t ld_de, void* stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    if (num_nodes < 0 || num_edges < 0 || num_cols <= 0 || ldx < num_cols || ld_da < num_cols || ld_dx < num_cols) return KAGNN_EINVAL;
    if (num_nodes > 0 && (!x || !da || !dx)) return KAGNN_EINVAL;
    if (num_edges > 0 && (!edge_feat || !edge_index || ld_edge < num_cols || (d_edge_feat && ld_de < num_cols))) return KAGNN_EINVAL;
    if (num_nodes == 0) return KAGNN_OK;
    KAGNN_LAUNCH(scale_rows_kernel, (unsigned)ceil_div64(num_nodes * (int64_t)num_cols, kBwdThreads), kBwdThreads, stream, da,
                 (long long)ld_da, (long long)num_nodes, (int)num_cols, self_scale, dx, (long long)ld_dx);
    KAGNN_LAUNCH_CHECK();
    if (num_edges == 0) return KAGNN_OK;
    KAGNN_LAUNCH(gine_bwd_edges_kernel, (unsigned)ceil_div64(num_edges * (int64_t)num_cols, kBwdThreads), kBwdThreads, stream, x,
                 (long long)ldx, edge_feat, (long long)ld_edge, edge_index, edge_index + num_edges, (long long)num_edges, (int)num_cols,
                 da, (long long)ld_da, dx, (long long)ld_dx, d_edge_feat, (long long)ld_de, (long long)num_nodes);
    KAGNN_LAUNCH_CHECK();
    return KAGNN_OK;
}
