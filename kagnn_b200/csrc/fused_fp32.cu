// kagnn_fused_layer_fwd -- general fp32 path (any G in [1,32], k in [1,4], any widths).
//
// One persistent launch per GNN layer.  Each CTA owns tiles of 64 destination rows:
//   1. STAGE   : warp-per-row CSR gather-sum (GIN / GINE / GCN-weighted / segment pooling) with
//                coalesced 128-bit row reads, accumulated in registers, pre-affine (bias / eval-BN /
//                SiLU) applied, tile parked in shared memory (optionally also stored: GCN layer output);
//   2. KAN x n : for every KAN layer of the chain the basis expansion (closed-form local de Boor for
//                B-splines, Gaussian RBF after LayerNorm for FastKAN, plus the SiLU base column) is
//                generated on the fly into a [K-chunk x 64] shared tile and contracted against the
//                pre-packed weight block with a register-tiled (4 x TN) FMA loop; the (N, in, G+k)
//                tensor of the reference (node_classification_clean/ekan.py:79-112,158-161) never exists;
//   3. EPILOGUE: base bias, post-affine (eval BatchNorm) and store with a leading dimension, so a layer
//                writes straight into its column slice of the skip-concat buffer
//                (node_classification_clean/models.py:196-201).
// Intermediate activations of a KAN chain stay in shared memory.
#include "common.cuh"

namespace {

constexpr int BM = 64;        // rows per tile
constexpr int NT = 256;       // threads per CTA (16 x 16 thread grid, 4 rows x TN cols each)
constexpr int NWARPS = NT / 32;
constexpr int KC_MAX = 64;    // K elements (feature x slot) per chunk
constexpr int WT_MAX = 128;   // widest output tile

struct LayerDev {
    int basis, in_f, out_f, G, k, slots1, out_pad;
    float t0, h, inv_h, inv_den;
    const float *w, *bias, *lnw, *lnb;
};

struct FusedParams {
    KagnnAggregate agg;
    long long num_rows;
    KagnnAffine pre, post;
    int has_pre, has_post;
    float* agg_out;
    long long ld_agg_out;
    float* y;
    long long ldy;
    int n_layers, stage_input;
    int ld_a, ld_b;
    int n_tiles, vec;
    LayerDev layers[KAGNN_MAX_LAYERS];
};

__device__ __forceinline__ float silu_f(float x) { return x / (1.0f + expf(-x)); }

__device__ __forceinline__ float apply_affine(const KagnnAffine& a, int c, float v) {
    if (a.scale) v *= __ldg(a.scale + c);
    if (a.shift) v += __ldg(a.shift + c);
    if (a.act == KAGNN_ACT_SILU) v = silu_f(v);
    return v;
}

template <bool VEC>
__device__ __forceinline__ void ldw(const float* p, float (&v)[4]) {
    if (VEC) {
        float4 t = __ldg(reinterpret_cast<const float4*>(p));
        v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
    } else {
        v[0] = __ldg(p);
    }
}

// ---------------------------------------------------------------------------------------------------
// STAGE: one warp produces one aggregated row (all feature columns) into dst (shared or global).
// ---------------------------------------------------------------------------------------------------
template <bool VEC>
__device__ void stage_row(const FusedParams& p, long long i, float* __restrict__ dst, float* __restrict__ dst2, int lane) {
    constexpr int W = VEC ? 4 : 1;
    const KagnnAggregate& a = p.agg;
    const int F = a.num_cols;
    const int mode = a.mode;
    int beg = 0, end = 0;
    if (mode != KAGNN_AGG_NONE) {
        beg = __ldg(a.rowptr + i);
        end = __ldg(a.rowptr + i + 1);
    }
    const bool segment = (mode == KAGNN_AGG_SEGMENT_SUM) || (mode == KAGNN_AGG_SEGMENT_MEAN);
    float self_s = 1.0f;
    if (mode == KAGNN_AGG_GIN || mode == KAGNN_AGG_GINE) self_s = a.self_scale;
    if (mode == KAGNN_AGG_WEIGHTED) self_s = a.self_weight ? __ldg(a.self_weight + i) : a.self_scale;
    const float out_scale = (mode == KAGNN_AGG_SEGMENT_MEAN) ? 1.0f / (float)max(end - beg, 1) : 1.0f;
    const long long self_row = a.src_index ? (long long)__ldg(a.src_index + i) : i;

    for (int c0 = 0; c0 < F; c0 += 64 * W) {
        const int ca = c0 + lane * W, cb = ca + 32 * W;
        const bool va = ca < F, vb = cb < F;
        float acc_a[4] = {0.f, 0.f, 0.f, 0.f}, acc_b[4] = {0.f, 0.f, 0.f, 0.f};
        if (!segment) {
            const float* xr = a.x + self_row * a.ldx;
            float t[4];
            if (va) { ldw<VEC>(xr + ca, t);
#pragma unroll
                for (int q = 0; q < W; ++q) acc_a[q] = self_s * t[q]; }
            if (vb) { ldw<VEC>(xr + cb, t);
#pragma unroll
                for (int q = 0; q < W; ++q) acc_b[q] = self_s * t[q]; }
        }
        for (int e0 = beg; e0 < end; e0 += 32) {
            const int cnt = min(32, end - e0);
            int my_j = 0, my_er = 0;
            float my_w = 1.0f;
            if (lane < cnt) {
                my_j = a.col ? __ldg(a.col + e0 + lane) : (e0 + lane);
                if (a.src_index) my_j = __ldg(a.src_index + my_j);
                if (mode == KAGNN_AGG_WEIGHTED) my_w = __ldg(a.edge_weight + e0 + lane);
                if (mode == KAGNN_AGG_GINE) my_er = __ldg(a.edge_row + e0 + lane);
            }
            for (int t0 = 0; t0 < cnt; t0 += 4) {
                float va4[4][4], vb4[4][4], ea4[4][4], eb4[4][4], w4[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const int src_lane = min(t0 + u, cnt - 1);
                    const int j = __shfl_sync(0xffffffffu, my_j, src_lane);
                    w4[u] = __shfl_sync(0xffffffffu, my_w, src_lane);
                    const int er = __shfl_sync(0xffffffffu, my_er, src_lane);
                    const bool on = (t0 + u) < cnt;
                    const float* xr = a.x + (long long)j * a.ldx;
#pragma unroll
                    for (int q = 0; q < 4; ++q) { va4[u][q] = 0.f; vb4[u][q] = 0.f; ea4[u][q] = 0.f; eb4[u][q] = 0.f; }
                    if (on && va) ldw<VEC>(xr + ca, va4[u]);
                    if (on && vb) ldw<VEC>(xr + cb, vb4[u]);
                    if (mode == KAGNN_AGG_GINE) {
                        const float* er_p = a.edge_feat + (long long)er * a.ld_edge;
                        if (on && va) ldw<VEC>(er_p + ca, ea4[u]);
                        if (on && vb) ldw<VEC>(er_p + cb, eb4[u]);
                    }
                    if (!on) w4[u] = 0.f;
                }
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const bool on = (t0 + u) < cnt;
                    if (mode == KAGNN_AGG_GINE) {
                        if (on) {
#pragma unroll
                            for (int q = 0; q < W; ++q) {
                                acc_a[q] += fmaxf(va4[u][q] + ea4[u][q], 0.f);
                                acc_b[q] += fmaxf(vb4[u][q] + eb4[u][q], 0.f);
                            }
                        }
                    } else if (mode == KAGNN_AGG_WEIGHTED) {
#pragma unroll
                        for (int q = 0; q < W; ++q) {
                            acc_a[q] = fmaf(w4[u], va4[u][q], acc_a[q]);
                            acc_b[q] = fmaf(w4[u], vb4[u][q], acc_b[q]);
                        }
                    } else {
                        if (on) {
#pragma unroll
                            for (int q = 0; q < W; ++q) { acc_a[q] += va4[u][q]; acc_b[q] += vb4[u][q]; }
                        }
                    }
                }
            }
        }
#pragma unroll
        for (int q = 0; q < W; ++q) {
            if (va) {
                float v = acc_a[q] * out_scale;
                if (p.has_pre) v = apply_affine(p.pre, ca + q, v);
                dst[ca + q] = v;
                if (dst2) dst2[ca + q] = v;
            }
            if (vb) {
                float v = acc_b[q] * out_scale;
                if (p.has_pre) v = apply_affine(p.pre, cb + q, v);
                dst[cb + q] = v;
                if (dst2) dst2[cb + q] = v;
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------------
// Basis expansion of one (row, feature) pair into column r of the [slot][BM] block at `a`.
// ---------------------------------------------------------------------------------------------------
__device__ __forceinline__ void expand_bspline(const LayerDev& L, float x, float* __restrict__ a) {
    // Uniform knots t_j = t0 + j*h (node_classification_clean/ekan.py:28-37; update_grid is never called by
    // any reference driver).  Interval j = floor((x-t0)/h); the k+1 non-zero bases B_{j-k..j} follow from the
    // local Cox-de Boor recursion on the fractional position (same recursion as ekan.py:96-105, restricted to
    // the non-zero entries); outside [t_0, t_last) or for NaN every basis is 0 (half-open indicator, :95).
    const int S = L.slots1 - 1, k = L.k;
    const float u = (x - L.t0) * L.inv_h;
    const float fl = floorf(u);
    const float fr = u - fl;
    const bool valid = (u >= 0.0f) && (u < (float)(L.G + 2 * k));
    const int j = valid ? (int)fl : 0;
    float b[5] = {1.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int d = 1; d <= 4; ++d) {
        if (d <= k) {
            const float inv_d = 1.0f / (float)d;
            float nb[5];
#pragma unroll
            for (int r = 0; r <= d; ++r) {
                float left = (r > 0) ? (fr + (float)(d - r)) * inv_d * b[r - 1] : 0.f;
                float right = (r < d) ? ((float)(r + 1) - fr) * inv_d * b[r] : 0.f;
                nb[r] = left + right;
            }
#pragma unroll
            for (int r = 0; r <= d; ++r) b[r] = nb[r];
        }
    }
    // +-inf: the reference's recursion multiplies (x - t_j) = inf by a zero indicator -> every basis is NaN (ekan.py:96-105)
    const float fill = isinf(x) ? __int_as_float(0x7fc00000) : 0.f;
    for (int c = 0; c < S; ++c) a[c * BM] = fill;
    if (valid) {
#pragma unroll
        for (int r = 0; r <= 4; ++r) {
            const int slot = j - k + r;
            if (r <= k && slot >= 0 && slot < S) a[slot * BM] = b[r];
        }
    }
    a[S * BM] = silu_f(x);
}

__device__ __forceinline__ void expand_rbf(const LayerDev& L, float x, float z, float* __restrict__ a) {
    // exp(-((z - g)/den)^2), g = grid_min + i*h (node_classification_clean/fastkan.py:42-47); base column = silu(raw x) (:82-83)
    const int G = L.slots1 - 1;
    for (int g = 0; g < G; ++g) {
        const float d = (z - (L.t0 + (float)g * L.h)) * L.inv_den;
        a[g * BM] = __expf(-d * d);
    }
    a[G * BM] = silu_f(x);
}

// ---------------------------------------------------------------------------------------------------
// One KAN layer on a 64-row tile.  `in` is shared memory (staged / previous layer) or global (wide bare
// KANLinear); the result goes to shared `dst` or, for the last layer, through the epilogue to y.
// ---------------------------------------------------------------------------------------------------
template <int TN>
__device__ void kan_layer(const FusedParams& p, const LayerDev& L, const float* in, long long ld_in, int in_rows,
                          float* dst, int ld_dst, bool last, long long row0, int nrows, float* As, float* Ws,
                          float* stats) {
    constexpr int WT = 16 * TN;
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4, lane = tid & 31, warp = tid >> 5;
    const int slots1 = L.slots1;
    const int KF = KC_MAX / slots1;
    const bool rbf = (L.basis == KAGNN_BASIS_RBF);
    const bool use_ln = rbf && (L.lnw != nullptr);
    __syncthreads();  // input tile complete
    if (use_ln) {
        // LayerNorm row statistics (biased variance, eps 1e-5; node_classification_clean/fastkan.py:66,78)
        for (int r = warp; r < BM; r += NWARPS) {
            const float* xr = in + (long long)min(r, in_rows - 1) * ld_in;
            float s = 0.f;
            for (int c = lane; c < L.in_f; c += 32) s += xr[c];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
            const float mean = s / (float)L.in_f;
            float v = 0.f;
            for (int c = lane; c < L.in_f; c += 32) { float d = xr[c] - mean; v = fmaf(d, d, v); }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
            if (lane == 0) { stats[2 * r] = mean; stats[2 * r + 1] = 1.0f / sqrtf(v / (float)L.in_f + 1e-5f); }
        }
        __syncthreads();
    }
    for (int n0 = 0; n0 < L.out_f; n0 += WT) {
        float acc[4][TN];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int q = 0; q < TN; ++q) acc[i][q] = 0.f;
        for (int f0 = 0; f0 < L.in_f; f0 += KF) {
            const int kf = min(KF, L.in_f - f0);
            const int kc = kf * slots1;
            __syncthreads();  // previous chunk consumed
            const float* wsrc = L.w + (size_t)f0 * slots1 * L.out_pad;
            for (int idx = tid; idx < kc * (WT / 4); idx += NT) {
                const int kk = idx / (WT / 4), c4 = idx - kk * (WT / 4);
                const int colw = n0 + c4 * 4;
                float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                if (colw < L.out_pad) v = __ldg(reinterpret_cast<const float4*>(wsrc + (size_t)kk * L.out_pad + colw));
                *reinterpret_cast<float4*>(&Ws[kk * WT + c4 * 4]) = v;
            }
            for (int idx = tid; idx < kf * BM; idx += NT) {
                const int r = idx & (BM - 1), f = idx >> 6;
                const float x = in[(long long)min(r, in_rows - 1) * ld_in + f0 + f];
                float* a = As + (f * slots1) * BM + r;
                if (!rbf) {
                    expand_bspline(L, x, a);
                } else {
                    float z = x;
                    if (use_ln) z = (x - stats[2 * r]) * stats[2 * r + 1] * __ldg(L.lnw + f0 + f) + (L.lnb ? __ldg(L.lnb + f0 + f) : 0.f);
                    expand_rbf(L, x, z, a);
                }
            }
            __syncthreads();
#pragma unroll 4
            for (int kk = 0; kk < kc; ++kk) {
                const float4 a4 = *reinterpret_cast<const float4*>(&As[kk * BM + ty * 4]);
                float w[TN];
                if (TN == 1) {
                    w[0] = Ws[kk * WT + tx];
                } else if (TN == 2) {
                    const float2 t = *reinterpret_cast<const float2*>(&Ws[kk * WT + tx * 2]);
                    w[0] = t.x; w[1] = t.y;
                } else {
                    const float4 t = *reinterpret_cast<const float4*>(&Ws[kk * WT + tx * 4]);
                    w[0] = t.x; w[1] = t.y; w[2] = t.z; w[3] = t.w;
                    if (TN == 8) {
                        const float4 t2 = *reinterpret_cast<const float4*>(&Ws[kk * WT + 64 + tx * 4]);
                        w[TN - 4] = t2.x; w[TN - 3] = t2.y; w[TN - 2] = t2.z; w[TN - 1] = t2.w;
                    }
                }
                const float av[4] = {a4.x, a4.y, a4.z, a4.w};
#pragma unroll
                for (int i = 0; i < 4; ++i)
#pragma unroll
                    for (int q = 0; q < TN; ++q) acc[i][q] = fmaf(av[i], w[q], acc[i][q]);
            }
        }
        // epilogue of this output tile
#pragma unroll
        for (int q = 0; q < TN; ++q) {
            const int cl = (TN == 8) ? ((q < 4) ? tx * 4 + q : 64 + tx * 4 + (q - 4)) : tx * TN + q;
            const int colo = n0 + cl;
            if (colo >= L.out_f) continue;
            const float bias = L.bias ? __ldg(L.bias + colo) : 0.f;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int r = ty * 4 + i;
                float v = acc[i][q] + bias;
                if (!last) {
                    dst[r * ld_dst + colo] = v;
                } else if (r < nrows) {
                    if (p.has_post) v = apply_affine(p.post, colo, v);
                    p.y[(row0 + r) * p.ldy + colo] = v;
                }
            }
        }
    }
}

template <bool VEC>
__global__ void __launch_bounds__(NT) fused_layer_kernel(const __grid_constant__ FusedParams p) {
    extern __shared__ __align__(16) float smem[];
    float* bufA = smem;
    float* bufB = bufA + (size_t)BM * p.ld_a;
    float* As = bufB + (size_t)BM * p.ld_b;
    float* Ws = As + KC_MAX * BM;
    float* stats = Ws + KC_MAX * WT_MAX;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;

    for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x) {
        const long long row0 = (long long)tile * BM;
        const int nrows = (int)min((long long)BM, p.num_rows - row0);
        if (p.n_layers == 0) {  // pure aggregation straight to global
            for (int r = warp; r < nrows; r += NWARPS)
                stage_row<VEC>(p, row0 + r, p.agg_out + (row0 + r) * p.ld_agg_out, nullptr, lane);
            continue;
        }
        const float* cur;
        long long ld_cur;
        int cur_rows;
        __syncthreads();  // previous tile fully consumed before bufA is overwritten
        if (p.stage_input) {
            for (int r = warp; r < BM; r += NWARPS) {
                float* drow = bufA + (size_t)r * p.ld_a;
                if (r < nrows) {
                    stage_row<VEC>(p, row0 + r, drow, p.agg_out ? p.agg_out + (row0 + r) * p.ld_agg_out : nullptr, lane);
                } else {
                    for (int c = lane; c < p.agg.num_cols; c += 32) drow[c] = 0.f;
                }
            }
            cur = bufA; ld_cur = p.ld_a; cur_rows = BM;
        } else {
            cur = p.agg.x + row0 * p.agg.ldx; ld_cur = p.agg.ldx; cur_rows = nrows;
        }
        for (int l = 0; l < p.n_layers; ++l) {
            const LayerDev& L = p.layers[l];
            const bool last = (l == p.n_layers - 1);
            float* dst = (l & 1) ? bufA : bufB;
            const int ld_dst = (l & 1) ? p.ld_a : p.ld_b;
            if (L.out_f <= 16) kan_layer<1>(p, L, cur, ld_cur, cur_rows, dst, ld_dst, last, row0, nrows, As, Ws, stats);
            else if (L.out_f <= 32) kan_layer<2>(p, L, cur, ld_cur, cur_rows, dst, ld_dst, last, row0, nrows, As, Ws, stats);
            else if (L.out_f <= 64) kan_layer<4>(p, L, cur, ld_cur, cur_rows, dst, ld_dst, last, row0, nrows, As, Ws, stats);
            else kan_layer<8>(p, L, cur, ld_cur, cur_rows, dst, ld_dst, last, row0, nrows, As, Ws, stats);
            cur = dst; ld_cur = ld_dst; cur_rows = BM;
        }
    }
}

}  // namespace

extern "C" int kagnn_fused_layer_fwd(const KagnnAggregate* agg, int64_t num_rows, const KagnnAffine* pre,
                                     float* agg_out, int64_t ld_agg_out, int32_t n_layers,
                                     const KagnnKanLayer* layers, const KagnnAffine* post, float* y, int64_t ldy,
                                     void* stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    if (!agg || num_rows < 0 || n_layers < 0 || n_layers > KAGNN_MAX_LAYERS) return KAGNN_EINVAL;
    if (n_layers > 0 && (!layers || !y)) return KAGNN_EINVAL;
    if (n_layers == 0 && !agg_out) return KAGNN_EINVAL;
    if (agg->mode < KAGNN_AGG_NONE || agg->mode > KAGNN_AGG_SEGMENT_MEAN) return KAGNN_EINVAL;
    if (agg->num_cols <= 0 || !agg->x || agg->ldx < agg->num_cols) return KAGNN_EINVAL;
    if (agg->mode != KAGNN_AGG_NONE && !agg->rowptr) return KAGNN_EINVAL;
    const bool segment = agg->mode == KAGNN_AGG_SEGMENT_SUM || agg->mode == KAGNN_AGG_SEGMENT_MEAN;
    if (agg->mode != KAGNN_AGG_NONE && !segment && !agg->col) return KAGNN_EINVAL;
    if (agg->mode == KAGNN_AGG_WEIGHTED && !agg->edge_weight) return KAGNN_EINVAL;
    if (agg->mode == KAGNN_AGG_GINE && (!agg->edge_feat || !agg->edge_row || agg->ld_edge < agg->num_cols)) return KAGNN_EINVAL;
    if (agg_out && ld_agg_out < agg->num_cols) return KAGNN_EINVAL;
    if (num_rows == 0) return KAGNN_OK;
    if (num_rows > (int64_t)INT32_MAX * 32) return KAGNN_EUNSUPPORTED;

    DeviceProps props{};
    int rc = kagnn_get_props(&props);
    if (rc != KAGNN_OK) return rc;

    FusedParams p{};
    p.agg = *agg;
    p.num_rows = num_rows;
    p.has_pre = pre != nullptr;
    p.has_post = post != nullptr;
    if (pre) p.pre = *pre;
    if (post) p.post = *post;
    p.agg_out = agg_out;
    p.ld_agg_out = ld_agg_out;
    p.y = y;
    p.ldy = ldy;
    p.n_layers = n_layers;
    p.n_tiles = (int)ceil_div64(num_rows, BM);

    int width = agg->num_cols;
    for (int l = 0; l < n_layers; ++l) {
        const KagnnKanLayer& s = layers[l];
        LayerDev& d = p.layers[l];
        if (s.in_features != width || s.out_features <= 0 || !s.packed_w) return KAGNN_EINVAL;
        if (s.basis != KAGNN_BASIS_BSPLINE && s.basis != KAGNN_BASIS_RBF) return KAGNN_EINVAL;
        if (s.grid_size < 1) return KAGNN_EINVAL;
        if (s.basis == KAGNN_BASIS_BSPLINE && (s.spline_order < 1 || s.spline_order > 4)) return KAGNN_EUNSUPPORTED;
        if (!(s.h > 0.f) && !(s.basis == KAGNN_BASIS_RBF && s.grid_size == 1)) return KAGNN_EINVAL;
        d.basis = s.basis;
        d.in_f = s.in_features;
        d.out_f = s.out_features;
        d.G = s.grid_size;
        d.k = (s.basis == KAGNN_BASIS_BSPLINE) ? s.spline_order : 0;
        d.slots1 = ((s.basis == KAGNN_BASIS_BSPLINE) ? s.grid_size + s.spline_order : s.grid_size) + 1;
        if (d.slots1 > KC_MAX) return KAGNN_EUNSUPPORTED;
        d.out_pad = pad4(s.out_features);
        d.t0 = s.t0;
        d.h = s.h;
        d.inv_h = s.h > 0.f ? 1.0f / s.h : 0.f;
        d.inv_den = s.inv_denominator;
        d.w = s.packed_w;
        d.bias = s.base_bias;
        d.lnw = s.ln_weight;
        d.lnb = s.ln_bias;
        if (!aligned16(d.w)) return KAGNN_EALIGN;
        width = s.out_features;
    }
    if (n_layers > 0 && ldy < width) return KAGNN_EINVAL;

    // vectorised staging needs 16-byte aligned rows everywhere it touches
    bool vec = (agg->num_cols % 4 == 0) && aligned16(agg->x) && (agg->ldx % 4 == 0);
    if (agg->mode == KAGNN_AGG_GINE) vec = vec && aligned16(agg->edge_feat) && (agg->ld_edge % 4 == 0);
    p.vec = vec;

    size_t smem = 0;
    if (n_layers > 0) {
        auto plan = [&](int stage_input) {
            int wa = stage_input ? agg->num_cols : 0, wb = 0;
            for (int l = 0; l + 1 < n_layers; ++l) {  // last layer writes to global
                int w = layers[l].out_features;
                if (l & 1) wa = wa > w ? wa : w; else wb = wb > w ? wb : w;
            }
            p.ld_a = wa ? pad4(wa) + 4 : 0;
            p.ld_b = wb ? pad4(wb) + 4 : 0;
            p.stage_input = stage_input;
            return ((size_t)BM * p.ld_a + (size_t)BM * p.ld_b + (size_t)KC_MAX * BM + (size_t)KC_MAX * WT_MAX + 2 * BM) * sizeof(float);
        };
        smem = plan(1);
        if (smem > (size_t)props.max_smem) {
            const bool can_stream = agg->mode == KAGNN_AGG_NONE && !pre && !agg_out && !agg->src_index;
            if (!can_stream) return KAGNN_EUNSUPPORTED;  // caller splits: aggregate (n_layers=0) then bare KAN
            smem = plan(0);
            if (smem > (size_t)props.max_smem) return KAGNN_EUNSUPPORTED;
        }
    }

    auto kern = vec ? fused_layer_kernel<true> : fused_layer_kernel<false>;
    KAGNN_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)props.max_smem));
    int occ = 1;
    KAGNN_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, NT, smem));
    if (occ < 1) occ = 1;
    long long grid = (long long)props.num_sms * occ;
    if (grid > p.n_tiles) grid = p.n_tiles;
    kern<<<(unsigned)grid, NT, smem, stream>>>(p);
    KAGNN_LAUNCH_CHECK();
    return KAGNN_OK;
}
