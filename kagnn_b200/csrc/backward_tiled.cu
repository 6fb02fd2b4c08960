// Shared-memory-tiled backward of the B-spline KAN layer: the two GEMM-shaped gradients
//
//   dW (packed layout)   dP[i][c][o] = sum_n E[n,i,c] dy[n,o]              E = (B_0..B_{S-1}, silu)(x[n,i])      ekan.py:154-162
//   dX                   dx[n,i]    = sum_c D[n,i,c] (sum_o dy[n,o] P[i][c][o])     D = d/dx of the same S + 1 functions
//
// as register-blocked fp32 FMA kernels.  backward.cu holds the first, barrier-free versions of both (one basis evaluation per
// output column in dW, an un-tiled stream of dy per input feature in dX: 148 ms for one training step of the arxiv-shaped
// model); those stay as the general fallback and as the host-checkable statement of the index arithmetic.  Here
//   * the S + 1 values of a (row, feature) pair are evaluated ONCE, as a dense vector (zeros in the slots outside the k + 1 live
//     ones), so every inner loop is a branch-free run of FMAs;
//   * dW: a block owns FG features x a slab of rows; per 32-row tile the dy rows and the basis vectors are staged in shared
//     memory, a thread keeps the (S + 1) x 4 gradient block of one feature and four output columns in registers
//     (36 FMAs per 4 shared loads) and adds it to HBM once per slab;
//   * dX: a block owns 32 rows x FG features; dy rows and the feature group's weights P[i][.][.] are staged in shared memory,
//     a thread keeps the S + 1 dot products of one feature for FOUR rows in registers (weights loaded once per four rows) and
//     contracts them with the derivative vector it evaluates itself.
// fp32 throughout (the reference's autograd is fp32), float atomics only between row slabs in dW (as before).
#include "common.cuh"

namespace {
constexpr int kT = 256;                 // threads per block
constexpr int kTR = 32;                 // rows per tile
constexpr int kS1Max = 12;              // S + 1 <= 12: G + k <= 11 (every configuration of the reference's drivers except the widest grids)
constexpr int kMaxOrderT = 4;

struct GeomT {
    int in_f, out_f, out_pad, G, k, S;
    float t0, inv_h;
};

__device__ __forceinline__ float sigmoid_t(float x) { return 1.0f / (1.0f + expf(-x)); }

// dense value vector E[0..S] (DERIV = false) or derivative vector D[0..S] (DERIV = true) of one input value; same interval
// selection and recursion as backward.cu's locate / local_bases (ekan.py:79-112 restricted to the k + 1 live bases)
template <bool DERIV>
__device__ __forceinline__ void dense_vector(const GeomT& g, float xv, float* __restrict__ out /* kS1Max */) {
#pragma unroll
    for (int c = 0; c < kS1Max; ++c) out[c] = 0.f;
    const float u = (xv - g.t0) * g.inv_h;
    const float fl = floorf(u);
    const bool valid = (u >= 0.0f) && (u < (float)(g.G + 2 * g.k));
    const float s = sigmoid_t(xv);
    const float base = DERIV ? s * (1.0f + xv * (1.0f - s)) : xv * s;
    if (valid) {
        const int cell = (int)fl;
        const float fr = u - fl;
        float b[kMaxOrderT + 1], m[kMaxOrderT + 1];
#pragma unroll
        for (int r = 0; r <= kMaxOrderT; ++r) { b[r] = 0.f; m[r] = 0.f; }
        b[0] = 1.f;
        for (int d = 1; d <= g.k; ++d) {
            if (d == g.k) {
#pragma unroll
                for (int r = 0; r < kMaxOrderT; ++r) m[r] = b[r];
            }
            const float inv_d = 1.0f / (float)d;
            float nb[kMaxOrderT + 1];
#pragma unroll
            for (int r = 0; r <= kMaxOrderT; ++r) {
                const float left = (r > 0 && r <= d) ? (fr + (float)(d - r)) * inv_d * b[r > 0 ? r - 1 : 0] : 0.f;
                const float right = (r < d) ? ((float)(r + 1) - fr) * inv_d * b[r] : 0.f;
                nb[r] = left + right;
            }
#pragma unroll
            for (int r = 0; r <= kMaxOrderT; ++r) b[r] = (r <= d) ? nb[r] : 0.f;
        }
#pragma unroll
        for (int r = 0; r <= kMaxOrderT; ++r) {
            const int slot = cell - g.k + r;
            float v;
            if (DERIV) {
                const float left = (r > 0) ? m[r > 0 ? r - 1 : 0] : 0.f, right = (r < g.k) ? m[r] : 0.f;
                v = (left - right) * g.inv_h;
            } else {
                v = b[r];
            }
            if (r <= g.k && slot >= 0 && slot < g.S) {
#pragma unroll
                for (int c = 0; c < kS1Max - 1; ++c)
                    if (c == slot) out[c] = v;
            }
        }
    }
#pragma unroll
    for (int c = 0; c < kS1Max; ++c)
        if (c == g.S) out[c] = base;
}

// ---------------------------------------------------------------------------------------------------------------------
// dW.  grid = (feature groups, row slabs); block = kT threads = FG features x OT threads, OT = ceil(out / 4).
// shared: dy tile [kTR][out4] | E tile [kTR][FG][kS1Max]
// ---------------------------------------------------------------------------------------------------------------------
template <int S1T>
__global__ void __launch_bounds__(kT) kan_bwd_weights_tiled_kernel(GeomT g, const float* __restrict__ x, long long ldx,
                                                                   const float* __restrict__ dy, long long ld_dy, long long n_rows,
                                                                   long long rows_per_slab, int FG, int OT, float* __restrict__ dP) {
    extern __shared__ __align__(16) float sm[];
    const int out4 = OT * 4;
    float* dys = sm;                                    // [kTR][out4]
    float* es = sm + kTR * out4;                        // [kTR][FG][kS1Max]
    const int tid = threadIdx.x;
    const int f_loc = tid / OT, ot = tid - f_loc * OT;  // this thread: feature f_loc of the group, output columns 4 ot .. 4 ot + 3
    const int f0 = blockIdx.x * FG;
    const bool active = f_loc < FG && (f0 + f_loc) < g.in_f;
    const long long r_beg = (long long)blockIdx.y * rows_per_slab, r_end = min(n_rows, r_beg + rows_per_slab);
    float acc[S1T][4];
#pragma unroll
    for (int c = 0; c < S1T; ++c) acc[c][0] = acc[c][1] = acc[c][2] = acc[c][3] = 0.f;

    for (long long rt = r_beg; rt < r_end; rt += kTR) {
        const int nr = (int)min((long long)kTR, r_end - rt);
        __syncthreads();                                // the previous tile has been consumed
        for (int idx = tid; idx < kTR * out4; idx += kT) {
            const int r = idx / out4, o = idx - r * out4;
            dys[idx] = (r < nr && o < g.out_f) ? dy[(rt + r) * ld_dy + o] : 0.f;
        }
        for (int idx = tid; idx < kTR * FG; idx += kT) {
            const int r = idx / FG, f = idx - r * FG;
            float e[kS1Max];
            if (r < nr && f0 + f < g.in_f) {
                dense_vector<false>(g, x[(rt + r) * ldx + f0 + f], e);
            } else {
#pragma unroll
                for (int c = 0; c < kS1Max; ++c) e[c] = 0.f;
            }
            float4* d = reinterpret_cast<float4*>(es + (size_t)idx * kS1Max);
            d[0] = make_float4(e[0], e[1], e[2], e[3]);
            d[1] = make_float4(e[4], e[5], e[6], e[7]);
            d[2] = make_float4(e[8], e[9], e[10], e[11]);
        }
        __syncthreads();
        if (active) {
#pragma unroll 4
            for (int r = 0; r < kTR; ++r) {
                const float4 d = *reinterpret_cast<const float4*>(dys + r * out4 + 4 * ot);
                const float4* ep = reinterpret_cast<const float4*>(es + ((size_t)r * FG + f_loc) * kS1Max);
                const float4 e0 = ep[0], e1 = ep[1], e2 = ep[2];
                const float ev[kS1Max] = {e0.x, e0.y, e0.z, e0.w, e1.x, e1.y, e1.z, e1.w, e2.x, e2.y, e2.z, e2.w};
#pragma unroll
                for (int c = 0; c < S1T; ++c) {
                    acc[c][0] = fmaf(ev[c], d.x, acc[c][0]);
                    acc[c][1] = fmaf(ev[c], d.y, acc[c][1]);
                    acc[c][2] = fmaf(ev[c], d.z, acc[c][2]);
                    acc[c][3] = fmaf(ev[c], d.w, acc[c][3]);
                }
            }
        }
    }
    if (active) {
        float* outp = dP + (long long)(f0 + f_loc) * (g.S + 1) * g.out_pad + 4 * ot;
#pragma unroll
        for (int c = 0; c < S1T; ++c) {
            if (c <= g.S) {
#pragma unroll
                for (int q = 0; q < 4; ++q)
                    if (4 * ot + q < g.out_f && acc[c][q] != 0.f) atomicAdd(outp + (long long)c * g.out_pad + q, acc[c][q]);
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// dX.  grid = (row tiles of kTR, feature groups); block = kT threads: warp w, quarter q (8 lanes) -> feature 4 w + q of the group,
// lane l8 of the quarter -> rows l8, l8 + 8, l8 + 16, l8 + 24 of the tile (adjacent lanes = adjacent rows: conflict-free 128-bit reads).  A block therefore covers 32 rows x 32 features per pass and loops
// over passes when the group is wider.
// shared: dy tile [kTR][out4 + 4] | P group [FG][S1][out4]
// ---------------------------------------------------------------------------------------------------------------------
template <int S1T>
__global__ void __launch_bounds__(kT) kan_bwd_input_tiled_kernel(GeomT g, const float* __restrict__ w, const float* __restrict__ x,
                                                                 long long ldx, const float* __restrict__ dy, long long ld_dy,
                                                                 long long n_rows, int FG, int out4, float* __restrict__ dx,
                                                                 long long ld_dx) {
    extern __shared__ __align__(16) float sm[];
    const int dld = out4 + 4;                           // (dld / 4) odd for out4 % 8 == 0: conflict-free float4 reads, lane = row group
    float* dys = sm;                                    // [kTR][dld]
    float* ps = sm + kTR * dld;                         // [FG][S1][out4]
    const int S1 = g.S + 1;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, q = lane >> 3, l8 = lane & 7;
    const long long rt = (long long)blockIdx.x * kTR;
    const int nr = (int)min((long long)kTR, n_rows - rt);
    const int f0 = blockIdx.y * FG;
    const int nf = min(FG, g.in_f - f0);
    for (int idx = tid; idx < kTR * out4; idx += kT) {
        const int r = idx / out4, o = idx - r * out4;
        dys[r * dld + o] = (r < nr && o < g.out_f) ? dy[(rt + r) * ld_dy + o] : 0.f;
    }
    for (int idx = tid; idx < nf * S1 * out4; idx += kT) {
        const int o = idx % out4, fc = idx / out4;      // fc = f * S1 + c
        ps[idx] = (o < g.out_f) ? w[((long long)(f0 * S1 + fc)) * g.out_pad + o] : 0.f;
    }
    __syncthreads();
    for (int fb = 0; fb < nf; fb += 32) {
        const int f = fb + 4 * warp + q;
        if (f >= nf) continue;
        float t[S1T][4];
#pragma unroll
        for (int c = 0; c < S1T; ++c) t[c][0] = t[c][1] = t[c][2] = t[c][3] = 0.f;
        const float* pf = ps + (size_t)f * S1 * out4;
        for (int o = 0; o < out4; o += 4) {
            float4 d[4];
#pragma unroll
            for (int rr = 0; rr < 4; ++rr) d[rr] = *reinterpret_cast<const float4*>(dys + (l8 + 8 * rr) * dld + o);
#pragma unroll
            for (int c = 0; c < S1T; ++c) {
                if (c < S1) {
                    const float4 p = *reinterpret_cast<const float4*>(pf + c * out4 + o);
#pragma unroll
                    for (int rr = 0; rr < 4; ++rr)
                        t[c][rr] = fmaf(d[rr].x, p.x, fmaf(d[rr].y, p.y, fmaf(d[rr].z, p.z, fmaf(d[rr].w, p.w, t[c][rr]))));
                }
            }
        }
#pragma unroll
        for (int rr = 0; rr < 4; ++rr) {
            const int r = l8 + 8 * rr;
            if (r < nr) {
                float dv[kS1Max];
                dense_vector<true>(g, x[(rt + r) * ldx + f0 + f], dv);
                float a = 0.f;
#pragma unroll
                for (int c = 0; c < S1T; ++c) a = fmaf(dv[c], t[c][rr], a);
                dx[(rt + r) * ld_dx + f0 + f] = a;
            }
        }
    }
}

int geometry_t(const KagnnKanLayer* L, GeomT* g) {
    if (!L || !L->packed_w || L->basis != KAGNN_BASIS_BSPLINE) return KAGNN_EUNSUPPORTED;
    if (L->in_features <= 0 || L->out_features <= 0 || L->grid_size < 1 || L->spline_order < 1 || L->spline_order > kMaxOrderT) return KAGNN_EUNSUPPORTED;
    if (L->grid_size + L->spline_order + 1 > kS1Max || !(L->h > 0.f)) return KAGNN_EUNSUPPORTED;
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

// returns KAGNN_EUNSUPPORTED for shapes the tiled kernels do not take (the caller then runs backward.cu's general kernels)
int kagnn_kan_bwd_weights_tiled(const KagnnKanLayer* layer, const float* x, int64_t ldx, const float* dy, int64_t ld_dy, int64_t num_rows,
                                float* d_packed, cudaStream_t stream) {
    GeomT g;
    int rc = geometry_t(layer, &g);
    if (rc != KAGNN_OK) return rc;
    if (g.out_f > 512 || num_rows < 1) return KAGNN_EUNSUPPORTED;
    const int OT = (g.out_f + 3) / 4;                   // threads per feature
    if (OT > kT) return KAGNN_EUNSUPPORTED;
    int FG = kT / OT;                                   // features per block
    if (FG > g.in_f) FG = g.in_f;
    const size_t smem_cap = 96 * 1024;
    while (FG > 1 && (size_t)(kTR * OT * 4 + kTR * FG * kS1Max) * sizeof(float) > smem_cap) --FG;
    const size_t smem = (size_t)(kTR * OT * 4 + kTR * FG * kS1Max) * sizeof(float);
    if (smem > smem_cap) return KAGNN_EUNSUPPORTED;
    KAGNN_CUDA_TRY(cudaMemsetAsync(d_packed, 0, sizeof(float) * (size_t)g.in_f * (size_t)(g.S + 1) * (size_t)g.out_pad, stream));
    const int fgroups = (g.in_f + FG - 1) / FG;
    // row slabs: enough blocks for a few waves of the machine, at most ~128 atomic adds per gradient element
    int64_t slabs = ceil_div64(num_rows, 1024);
    const int64_t want = ceil_div64(148 * 4, fgroups);
    if (slabs > want) slabs = want;
    if (slabs > 128) slabs = 128;
    if (slabs < 1) slabs = 1;
    int64_t rows_per_slab = ceil_div64(num_rows, slabs);
    rows_per_slab = ceil_div64(rows_per_slab, kTR) * kTR;
    slabs = ceil_div64(num_rows, rows_per_slab);
    auto kern = (g.S + 1 <= 9) ? kan_bwd_weights_tiled_kernel<9> : kan_bwd_weights_tiled_kernel<kS1Max>;
    KAGNN_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_cap));
    kern<<<dim3((unsigned)fgroups, (unsigned)slabs, 1), kT, smem, stream>>>(g, x, (long long)ldx, dy, (long long)ld_dy, (long long)num_rows,
                                                                           (long long)rows_per_slab, FG, OT, d_packed);
    KAGNN_LAUNCH_CHECK();
    return KAGNN_OK;
}

int kagnn_kan_bwd_input_tiled(const KagnnKanLayer* layer, const float* x, int64_t ldx, const float* dy, int64_t ld_dy, int64_t num_rows,
                              float* dx, int64_t ld_dx, cudaStream_t stream) {
    GeomT g;
    int rc = geometry_t(layer, &g);
    if (rc != KAGNN_OK) return rc;
    if (num_rows < 1) return KAGNN_EUNSUPPORTED;
    const int out4 = ((g.out_f + 7) / 8) * 8;           // multiple of 8: (out4 + 4) / 4 is odd
    const int S1 = g.S + 1;
    const size_t dy_bytes = (size_t)kTR * (out4 + 4) * sizeof(float);
    const size_t per_f = (size_t)S1 * out4 * sizeof(float);
    // two blocks per SM when a useful feature group fits in ~100 KB, else one block with up to 200 KB
    size_t smem_cap = 100 * 1024;
    if (dy_bytes + 16 * per_f > smem_cap) smem_cap = 200 * 1024;
    if (dy_bytes + 4 * per_f > smem_cap) return KAGNN_EUNSUPPORTED;
    int FG = (int)((smem_cap - dy_bytes) / per_f);
    if (FG > 32) FG = 32;                               // one pass of the block = 8 warps x 4 features
    FG = (FG / 4) * 4;
    if (FG > ((g.in_f + 3) / 4) * 4) FG = ((g.in_f + 3) / 4) * 4;
    const size_t smem = dy_bytes + (size_t)FG * per_f;
    const int fgroups = (g.in_f + FG - 1) / FG;
    if (fgroups > 65535) return KAGNN_EUNSUPPORTED;
    auto kern = (S1 <= 9) ? kan_bwd_input_tiled_kernel<9> : kan_bwd_input_tiled_kernel<kS1Max>;
    KAGNN_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_cap));
    kern<<<dim3((unsigned)ceil_div64(num_rows, kTR), (unsigned)fgroups, 1), kT, smem, stream>>>(
        g, layer->packed_w, x, (long long)ldx, dy, (long long)ld_dy, (long long)num_rows, FG, out4, dx, (long long)ld_dx);
    KAGNN_LAUNCH_CHECK();
    return KAGNN_OK;
}


// ---------------------------------------------------------------------------------------------------------------------
// LayerNorm backward (FastKANLayer's nn.LayerNorm, fastkan.py:66,78), product versions of backward.cu's thread-per-row kernels:
//   rows:    warp per row, coalesced column strides, the two row means by warp shuffles
//            dx = rstd (dz gamma - mean(dz gamma) - xhat mean(dz gamma xhat)) + dx_base
//   params:  block = a slab of rows x 32 columns, eight row lanes per column, partial sums joined through shared memory,
//            one float atomic per column and block:  dgamma = sum_n dz xhat, dbeta = sum_n dz
// ---------------------------------------------------------------------------------------------------------------------
namespace {
__global__ void layernorm_bwd_rows_warp_kernel(const float* __restrict__ x, long long ldx, const float* __restrict__ stats,
                                               const float* __restrict__ ln_w, const float* __restrict__ dz, long long ld_dz,
                                               const float* __restrict__ dxb, long long ld_dxb, long long n_rows, int cols,
                                               float* __restrict__ dx, long long ld_dx) {
    const long long n = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (n >= n_rows) return;
    const float mean = stats[2 * n], rstd = stats[2 * n + 1];
    const float *xr = x + n * ldx, *dzr = dz + n * ld_dz;
    float s1 = 0.f, s2 = 0.f;
    for (int i = lane; i < cols; i += 32) {
        const float gz = dzr[i] * (ln_w ? ln_w[i] : 1.0f);
        s1 += gz;
        s2 = fmaf(gz, (xr[i] - mean) * rstd, s2);
    }
    for (int o = 16; o; o >>= 1) {
        s1 += __shfl_xor_sync(0xffffffffu, s1, o);
        s2 += __shfl_xor_sync(0xffffffffu, s2, o);
    }
    const float inv_f = 1.0f / (float)cols;
    s1 *= inv_f;
    s2 *= inv_f;
    for (int i = lane; i < cols; i += 32) {
        const float gz = dzr[i] * (ln_w ? ln_w[i] : 1.0f);
        float v = rstd * (gz - s1 - (xr[i] - mean) * rstd * s2);
        if (dxb) v += dxb[n * ld_dxb + i];
        dx[n * ld_dx + i] = v;
    }
}

__global__ void __launch_bounds__(256) layernorm_bwd_params_tile_kernel(const float* __restrict__ x, long long ldx, const float* __restrict__ stats,
                                                                        const float* __restrict__ dz, long long ld_dz, long long n_rows, int cols,
                                                                        long long rows_per_block, float* __restrict__ d_w, float* __restrict__ d_b) {
    __shared__ float pw[8][33], pb[8][33];
    const int cl = threadIdx.x & 31, rl = threadIdx.x >> 5;
    const int c = blockIdx.y * 32 + cl;
    const long long r0 = (long long)blockIdx.x * rows_per_block, r1 = min(n_rows, r0 + rows_per_block);
    float sw = 0.f, sb = 0.f;
    if (c < cols) {
        for (long long r = r0 + rl; r < r1; r += 8) {
            const float d = dz[r * ld_dz + c];
            sw = fmaf(d, (x[r * ldx + c] - stats[2 * r]) * stats[2 * r + 1], sw);
            sb += d;
        }
    }
    pw[rl][cl] = sw;
    pb[rl][cl] = sb;
    __syncthreads();
    if (rl == 0 && c < cols) {
        float a = 0.f, b = 0.f;
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            a += pw[k][cl];
            b += pb[k][cl];
        }
        if (d_w) atomicAdd(&d_w[c], a);
        if (d_b) atomicAdd(&d_b[c], b);
    }
}
}  // namespace

// d_weight / d_bias are zeroed by the caller (kagnn_layernorm_bwd); always succeeds for num_rows > 0
int kagnn_layernorm_bwd_fast(const float* x, int64_t ldx, const float* ln_stats, const float* ln_weight, const float* dz, int64_t ld_dz,
                             const float* dx_base, int64_t ld_dxb, int64_t num_rows, int32_t num_cols, float* dx, int64_t ld_dx,
                             float* d_weight, float* d_bias, cudaStream_t stream) {
    if (num_rows < 1 || num_cols < 1) return KAGNN_EUNSUPPORTED;
    layernorm_bwd_rows_warp_kernel<<<(unsigned)ceil_div64(num_rows * 32, kT), kT, 0, stream>>>(
        x, (long long)ldx, ln_stats, ln_weight, dz, (long long)ld_dz, dx_base, (long long)ld_dxb, (long long)num_rows, (int)num_cols, dx,
        (long long)ld_dx);
    KAGNN_LAUNCH_CHECK();
    if (d_weight || d_bias) {
        const int64_t rows_per_block = 512;
        const dim3 grid((unsigned)ceil_div64(num_rows, rows_per_block), (unsigned)((num_cols + 31) / 32), 1);
        layernorm_bwd_params_tile_kernel<<<grid, 256, 0, stream>>>(x, (long long)ldx, ln_stats, dz, (long long)ld_dz, (long long)num_rows,
                                                                  (int)num_cols, (long long)rows_per_block, d_weight, d_bias);
        KAGNN_LAUNCH_CHECK();
    }
    return KAGNN_OK;
}


// ---------------------------------------------------------------------------------------------------------------------
// BatchNorm1d backward (batch statistics), product versions of backward.cu's kernels with the same tiling as the forward
// (graph.cu: bn_stats_tile_kernel): sums[0..3][c] = sum x, sum x^2, sum dy, sum dy x in fp64; then
//   dx = A dy + B (x - mean) + C,   A = gamma rstd,  B = -gamma rstd^2 m2,  C = -A mean(dy),  m2 = mean(dy xhat)
// with the per-column constants formed once per thread in fp64.
// ---------------------------------------------------------------------------------------------------------------------
namespace {
__global__ void __launch_bounds__(256) bn_bwd_sums_tile_kernel(const float* __restrict__ x, long long ldx, const float* __restrict__ dy, long long ld_dy,
                                                               long long rows, int cols, long long rows_per_block, double* __restrict__ sums) {
    __shared__ double p[4][8][33];
    const int cl = threadIdx.x & 31, rl = threadIdx.x >> 5;
    const int c = blockIdx.y * 32 + cl;
    const long long r0 = (long long)blockIdx.x * rows_per_block, r1 = min(rows, r0 + rows_per_block);
    double sx = 0.0, sxx = 0.0, sd = 0.0, sdx = 0.0;
    if (c < cols) {
        for (long long r = r0 + rl; r < r1; r += 8) {
            const double v = (double)x[r * ldx + c], d = (double)dy[r * ld_dy + c];
            sx += v;
            sxx += v * v;
            sd += d;
            sdx += d * v;
        }
    }
    p[0][rl][cl] = sx;
    p[1][rl][cl] = sxx;
    p[2][rl][cl] = sd;
    p[3][rl][cl] = sdx;
    __syncthreads();
    if (rl < 4 && c < cols) {                             // row lane q joins sum q of its column
        double a = 0.0;
#pragma unroll
        for (int k = 0; k < 8; ++k) a += p[rl][k][cl];
        atomicAdd(&sums[(long long)rl * cols + c], a);
    }
}

__global__ void __launch_bounds__(256) bn_bwd_apply_tile_kernel(const float* __restrict__ x, long long ldx, const float* __restrict__ dy, long long ld_dy,
                                                                long long rows, int cols, const double* __restrict__ sums,
                                                                const float* __restrict__ weight, float eps, long long rows_per_block,
                                                                float* __restrict__ dx, long long ld_dx, float* __restrict__ d_weight,
                                                                float* __restrict__ d_bias) {
    const int cl = threadIdx.x & 31, rl = threadIdx.x >> 5;
    const int c = blockIdx.y * 32 + cl;
    if (c >= cols) return;
    const double inv_n = 1.0 / (double)rows;
    const double mean = sums[c] * inv_n;
    const double var = fmax(sums[cols + c] * inv_n - mean * mean, 0.0);
    const double rstd = 1.0 / sqrt(var + (double)eps);
    const double sum_dy = sums[2 * cols + c];
    const double sum_dy_xhat = (sums[3 * cols + c] - mean * sum_dy) * rstd;
    const double gamma = weight ? (double)weight[c] : 1.0;
    const float A = (float)(gamma * rstd), B = (float)(-gamma * rstd * rstd * sum_dy_xhat * inv_n), C = (float)(-gamma * rstd * sum_dy * inv_n);
    const float mean_f = (float)mean;
    const long long r0 = (long long)blockIdx.x * rows_per_block, r1 = min(rows, r0 + rows_per_block);
    for (long long r = r0 + rl; r < r1; r += 8) dx[r * ld_dx + c] = fmaf(A, dy[r * ld_dy + c], fmaf(B, x[r * ldx + c] - mean_f, C));
    if (blockIdx.x == 0 && rl == 0) {
        if (d_weight) d_weight[c] = (float)sum_dy_xhat;
        if (d_bias) d_bias[c] = (float)sum_dy;
    }
}
}  // namespace

// `sums` (4 * num_cols doubles) is zeroed by the caller (kagnn_batchnorm_train_bwd)
int kagnn_batchnorm_train_bwd_fast(const float* x, int64_t ldx, const float* dy, int64_t ld_dy, int64_t num_rows, int32_t num_cols,
                                   const float* weight, float eps, float* dx, int64_t ld_dx, float* d_weight, float* d_bias, double* sums,
                                   cudaStream_t stream) {
    if (num_rows < 1 || num_cols < 1 || (num_cols + 31) / 32 > 65535) return KAGNN_EUNSUPPORTED;
    const int64_t rows_per_block = 512;
    const dim3 grid((unsigned)ceil_div64(num_rows, rows_per_block), (unsigned)((num_cols + 31) / 32), 1);
    bn_bwd_sums_tile_kernel<<<grid, 256, 0, stream>>>(x, (long long)ldx, dy, (long long)ld_dy, (long long)num_rows, (int)num_cols,
                                                     (long long)rows_per_block, sums);
    KAGNN_LAUNCH_CHECK();
    bn_bwd_apply_tile_kernel<<<grid, 256, 0, stream>>>(x, (long long)ldx, dy, (long long)ld_dy, (long long)num_rows, (int)num_cols, sums, weight,
                                                      eps, (long long)rows_per_block, dx, (long long)ld_dx, d_weight, d_bias);
    KAGNN_LAUNCH_CHECK();
    return KAGNN_OK;
}
