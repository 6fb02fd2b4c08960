// STAGE: the aggregation half of the fused layer, shared by the fp32 and the tensor-core kernels.
// One warp produces one aggregated destination row (all feature columns) from the CSR: coalesced 128-bit
// reads of neighbour rows, register accumulation in CSR order (deterministic, no atomics), pre-affine
// (GCNConv bias / eval BatchNorm / SiLU) applied before the row is parked in shared memory or stored.
//   GIN  : self_scale*x_i + sum_j x_j              (PyG GINConv,  node_classification_clean/models.py:48-56)
//   GINE : self_scale*x_i + sum_j relu(x_j + e_ji) (PyG GINEConv, graph_regression/models.py:98)
//   GCN  : self_w[i]*x_i + sum_j w_ij x_j          (PyG GCNConv,  node_classification_clean/models.py:31-37)
//   POOL : sum / mean of a contiguous row segment  (global_add_pool / global_mean_pool, graph_classification/models.py:117,192)
#pragma once
#include "common.cuh"

struct StageParams {
    KagnnAggregate agg;
    KagnnAffine pre;
    int has_pre;
    int _pad;
};

__device__ __forceinline__ float silu_f(float x) { return x / (1.0f + expf(-x)); }

__device__ __forceinline__ float apply_affine(const KagnnAffine& a, int c, float v) {
    if (a.scale) v *= __ldg(a.scale + c);
    if (a.shift) v += __ldg(a.shift + c);
    if (a.act == KAGNN_ACT_SILU) v = silu_f(v);
    return v;
}

// source row j of the (possibly node-sharded) feature matrix: owned rows in x, halo rows in x_halo
__device__ __forceinline__ const float* src_row(const KagnnAggregate& a, long long j) {
    if (a.x_halo != nullptr && j >= a.num_local_src) return a.x_halo + (j - a.num_local_src) * a.ld_halo;
    return a.x + j * a.ldx;
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
__device__ void stage_row(const StageParams& p, long long i, float* __restrict__ dst, float* __restrict__ dst2, int lane) {
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
                    const float* xr = src_row(a, j);
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

