// Shared helpers for libkagnn_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/kagnn_b200.h"

#define KAGNN_MAX_LAYERS 8
#define KAGNN_PULL_WARPS 16       // warps of a kagnn_gather_rows_peer_ordered block: arrivals per chunk counter and use (graph.cu, fused_tc2.cu)

#define KAGNN_CUDA_TRY(expr)                         \
    do {                                             \
        cudaError_t _e = (expr);                     \
        if (_e != cudaSuccess) return KAGNN_ECUDA;   \
    } while (0)

#define KAGNN_LAUNCH_CHECK()                                  \
    do {                                                      \
        if (cudaGetLastError() != cudaSuccess) return KAGNN_ECUDA; \
    } while (0)

static inline int64_t ceil_div64(int64_t a, int64_t b) { return (a + b - 1) / b; }
static inline int pad4(int v) { return (v + 3) & ~3; }
static inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

struct DeviceProps {
    int num_sms;
    int max_smem;
    int cc_major, cc_minor;
};
// cached per device; returns KAGNN_OK / KAGNN_ECUDA
int kagnn_get_props(DeviceProps* out);

// the two implementations behind kagnn_fused_layer_fwd (dispatch.cu)
int kagnn_fused_fwd_fp32(const KagnnAggregate* agg, int64_t num_rows, const KagnnAffine* pre, float* agg_out,
                         int64_t ld_agg_out, int32_t n_layers, const KagnnKanLayer* layers, const KagnnAffine* post,
                         float* y, int64_t ldy, cudaStream_t stream);
int kagnn_fused_fwd_tc(const KagnnAggregate* agg, int64_t num_rows, const KagnnAffine* pre, float* agg_out,
                       int64_t ld_agg_out, int32_t n_layers, const KagnnKanLayer* layers, const KagnnAffine* post, float* y,
                       int64_t ldy, cudaStream_t stream);
int kagnn_fused_fwd_tc2(const KagnnAggregate* agg, int64_t num_rows, const KagnnAffine* pre, float* agg_out,
                        int64_t ld_agg_out, int32_t n_layers, const KagnnKanLayer* layers, const KagnnAffine* post, float* y,
                        int64_t ldy, cudaStream_t stream);
int kagnn_fused_fwd_tc2_g16(const KagnnAggregate* agg, int64_t num_rows, const KagnnAffine* pre, float* agg_out,
                            int64_t ld_agg_out, int32_t n_layers, const KagnnKanLayer* layers, const KagnnAffine* post, float* y,
                            int64_t ldy, cudaStream_t stream);
int kagnn_aggregate_only_tc2(const KagnnAggregate* agg, int64_t num_rows, const KagnnAffine* pre, float* agg_out,
                             int64_t ld_agg_out, cudaStream_t stream);
// shared-memory-tiled B-spline backward (backward_tiled.cu); KAGNN_EUNSUPPORTED -> the general kernels of backward.cu
// tcgen05 versions (backward_tc.cu): tried first, KAGNN_EUNSUPPORTED outside their limits
int kagnn_kan_bwd_weights_tc(const KagnnKanLayer* layer, const float* x, int64_t ldx, const float* dy, int64_t ld_dy, int64_t num_rows,
                             float* d_packed, cudaStream_t stream);
int kagnn_kan_bwd_input_tc(const KagnnKanLayer* layer, const float* x, int64_t ldx, const float* dy, int64_t ld_dy, int64_t num_rows,
                           float* dx, int64_t ld_dx, cudaStream_t stream);
int kagnn_rbf_bwd_input_tc(const KagnnKanLayer* layer, const float* x, int64_t ldx, const float* ln_stats, const float* dy, int64_t ld_dy,
                           int64_t num_rows, float* dz, int64_t ld_dz, float* dx_base, int64_t ld_dxb, cudaStream_t stream);
int kagnn_rbf_bwd_weights_tc(const KagnnKanLayer* layer, const float* x, int64_t ldx, const float* ln_stats, const float* dy, int64_t ld_dy,
                             int64_t num_rows, float* d_packed, cudaStream_t stream);
int kagnn_batchnorm_train_bwd_fast(const float* x, int64_t ldx, const float* dy, int64_t ld_dy, int64_t num_rows, int32_t num_cols,
                                   const float* weight, float eps, float* dx, int64_t ld_dx, float* d_weight, float* d_bias, double* sums,
                                   cudaStream_t stream);
int kagnn_layernorm_bwd_fast(const float* x, int64_t ldx, const float* ln_stats, const float* ln_weight, const float* dz, int64_t ld_dz,
                             const float* dx_base, int64_t ld_dxb, int64_t num_rows, int32_t num_cols, float* dx, int64_t ld_dx,
                             float* d_weight, float* d_bias, cudaStream_t stream);
int kagnn_kan_bwd_weights_tiled(const KagnnKanLayer* layer, const float* x, int64_t ldx, const float* dy, int64_t ld_dy, int64_t num_rows,
                                float* d_packed, cudaStream_t stream);
int kagnn_kan_bwd_input_tiled(const KagnnKanLayer* layer, const float* x, int64_t ldx, const float* dy, int64_t ld_dy, int64_t num_rows,
                              float* dx, int64_t ld_dx, cudaStream_t stream);
int kagnn_validate_fused_args(const KagnnAggregate* agg, int64_t num_rows, const float* agg_out, int64_t ld_agg_out,
                              int32_t n_layers, const KagnnKanLayer* layers, const float* y);
