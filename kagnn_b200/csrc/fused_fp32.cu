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
#include "stage.cuh"

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
    StageParams st;
    long long num_rows;
    KagnnAffine post;
    int has_post;
    float* agg_out;
    long long ld_agg_out;
    float* y;
    long long ldy;
    int n_layers, stage_input;
    int ld_a, ld_b;
    int n_tiles, vec;
    LayerDev layers[KAGNN_MAX_LAYERS];
};

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
                stage_row<VEC>(p.st, row0 + r, p.agg_out + (row0 + r) * p.ld_agg_out, nullptr, lane);
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
                    stage_row<VEC>(p.st, row0 + r, drow, p.agg_out ? p.agg_out + (row0 + r) * p.ld_agg_out : nullptr, lane);
                } else {
                    for (int c = lane; c < p.st.agg.num_cols; c += 32) drow[c] = 0.f;
                }
            }
            cur = bufA; ld_cur = p.ld_a; cur_rows = BM;
        } else {
            cur = p.st.agg.x + row0 * p.st.agg.ldx; ld_cur = p.st.agg.ldx; cur_rows = nrows;
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

int kagnn_fused_fwd_fp32(const KagnnAggregate* agg, int64_t num_rows, const KagnnAffine* pre, float* agg_out,
                         int64_t ld_agg_out, int32_t n_layers, const KagnnKanLayer* layers, const KagnnAffine* post,
                         float* y, int64_t ldy, cudaStream_t stream) {
    int vrc = kagnn_validate_fused_args(agg, num_rows, agg_out, ld_agg_out, n_layers, layers, y);
    if (vrc != KAGNN_OK) return vrc;
    if (num_rows == 0) return KAGNN_OK;

    DeviceProps props{};
    int rc = kagnn_get_props(&props);
    if (rc != KAGNN_OK) return rc;

    FusedParams p{};
    p.st.agg = *agg;
    p.st.has_pre = pre != nullptr;
    if (pre) p.st.pre = *pre;
    p.num_rows = num_rows;
    p.has_post = post != nullptr;
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
    if (agg->x_halo) vec = vec && aligned16(agg->x_halo) && (agg->ld_halo % 4 == 0);
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
