// kagnn_fused_layer_fwd: argument validation shared by both kernels happens in the implementations; this file
// only chooses between the tensor-core path (tcgen05, fused_tc.cu) and the general fp32 path (fused_fp32.cu).
// Both run on the GPU; there is no CPU path.
#include <atomic>

#include "common.cuh"

namespace {
std::atomic<int> g_path_mode{KAGNN_PATH_AUTO};
std::atomic<long long> g_count_tc{0}, g_count_fp32{0}, g_count_tc2{0}, g_count_agg{0};
std::atomic<int> g_precision{KAGNN_PREC_FP32};
std::atomic<int> g_tc_variant{0};   // 0 = auto, 1 = only the shared-memory-A kernel (fused_tc.cu); tests/benchmarks
}  // namespace

int kagnn_validate_fused_args(const KagnnAggregate* agg, int64_t num_rows, const float* agg_out, int64_t ld_agg_out,
                              int32_t n_layers, const KagnnKanLayer* layers, const float* y) {
    if (!agg || num_rows < 0 || n_layers < 0 || n_layers > KAGNN_MAX_LAYERS) return KAGNN_EINVAL;
    if (n_layers > 0 && (!layers || !y)) return KAGNN_EINVAL;
    if (n_layers == 0 && !agg_out) return KAGNN_EINVAL;
    if (agg->num_push && (n_layers == 0 || !agg->push_y || agg->num_push < 0 || agg->num_push > 8)) return KAGNN_EINVAL;   // the pushed rows are those of y
    if (agg->mode < KAGNN_AGG_NONE || agg->mode > KAGNN_AGG_SEGMENT_MEAN) return KAGNN_EINVAL;
    if (agg->num_cols <= 0 || !agg->x || agg->ldx < agg->num_cols - (agg->num_head_cols > 0 ? agg->num_head_cols : 0)) return KAGNN_EINVAL;
    if (agg->mode != KAGNN_AGG_NONE && !agg->rowptr) return KAGNN_EINVAL;
    const bool segment = agg->mode == KAGNN_AGG_SEGMENT_SUM || agg->mode == KAGNN_AGG_SEGMENT_MEAN;
    if (agg->mode != KAGNN_AGG_NONE && !segment && !agg->col) return KAGNN_EINVAL;
    if (agg->mode == KAGNN_AGG_WEIGHTED && !agg->edge_weight) return KAGNN_EINVAL;
    if (agg->mode == KAGNN_AGG_GINE && (!agg->edge_feat || !agg->edge_row || agg->ld_edge < agg->num_cols)) return KAGNN_EINVAL;
    if (agg_out && ld_agg_out < agg->num_cols) return KAGNN_EINVAL;
    if (agg->x_halo && (agg->ld_halo < agg->num_cols || agg->num_local_src < 0)) return KAGNN_EINVAL;
    if (agg->x_halo && (agg->mode == KAGNN_AGG_NONE || agg->src_index)) return KAGNN_EINVAL;
    if (agg->num_head_cols < 0 || (agg->num_head_cols > 0 && agg->num_head_cols >= agg->num_cols)) return KAGNN_EINVAL;
    if (agg->num_head_cols > 0 && (agg->mode != KAGNN_AGG_NONE || !agg->x_head || agg->ld_head < agg->num_head_cols || agg->src_index))
        return KAGNN_EINVAL;
    if (agg->peer_x && (agg->x_halo || agg->src_index || agg->rows_per_rank <= 0 || agg->num_ranks <= 0)) return KAGNN_EINVAL;
    if (agg->peer_x && agg->mode != KAGNN_AGG_GIN && agg->mode != KAGNN_AGG_WEIGHTED) return KAGNN_EINVAL;
    if (num_rows > (int64_t)INT32_MAX * 32) return KAGNN_EUNSUPPORTED;

    return KAGNN_OK;
}

extern "C" int kagnn_set_path(int mode) {
    if (mode != KAGNN_PATH_AUTO && mode != KAGNN_PATH_FP32 && mode != KAGNN_PATH_TC) return KAGNN_EINVAL;
    g_path_mode.store(mode);
    return KAGNN_OK;
}

extern "C" int kagnn_set_precision(int mode) {
    if (mode != KAGNN_PREC_FP32 && mode != KAGNN_PREC_BF16) return KAGNN_EINVAL;
    g_precision.store(mode);
    return KAGNN_OK;
}

extern "C" int kagnn_get_precision(void) { return g_precision.load(); }

extern "C" int kagnn_set_tc_variant(int variant) {
    if (variant != 0 && variant != 1) return KAGNN_EINVAL;
    g_tc_variant.store(variant);
    return KAGNN_OK;
}

extern "C" int64_t kagnn_get_tc2_launches(void) { return g_count_tc2.load(); }

extern "C" int kagnn_get_launch_counters(int64_t* tc_launches, int64_t* fp32_launches) {
    if (tc_launches) *tc_launches = g_count_tc.load();
    if (fp32_launches) *fp32_launches = g_count_fp32.load();
    return KAGNN_OK;
}

extern "C" int kagnn_fused_layer_fwd(const KagnnAggregate* agg, int64_t num_rows, const KagnnAffine* pre, float* agg_out,
                                     int64_t ld_agg_out, int32_t n_layers, const KagnnKanLayer* layers,
                                     const KagnnAffine* post, float* y, int64_t ldy, void* stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    int vrc = kagnn_validate_fused_args(agg, num_rows, agg_out, ld_agg_out, n_layers, layers, y);
    if (vrc != KAGNN_OK) return vrc;
    const int mode = g_path_mode.load();
    if (mode != KAGNN_PATH_FP32 && agg && n_layers >= 1 && layers && y && num_rows > 0) {
        bool have_tc = true;
        for (int l = 0; l < n_layers && l < KAGNN_MAX_LAYERS; ++l) have_tc = have_tc && layers[l].packed_w_tc != nullptr;
        if (have_tc) {
            // run the shared argument checks of the fp32 path first?  No: fused_tc validates what it touches and
            // returns KAGNN_EUNSUPPORTED for anything it cannot run, in which case the general kernel takes over.
            int rc = KAGNN_EUNSUPPORTED;
            if (g_tc_variant.load() != 1) {      // pipelined kernel (A operand in TMEM) first; it declines what it cannot run
                rc = kagnn_fused_fwd_tc2(agg, num_rows, pre, agg_out, ld_agg_out, n_layers, layers, post, y, ldy, stream);
                if (rc == KAGNN_OK) {
                    g_count_tc.fetch_add(1);
                    g_count_tc2.fetch_add(1);
                    return rc;
                }
                if (rc != KAGNN_EUNSUPPORTED) return rc;
            }
            if (agg->peer_x || agg->num_head_cols || agg->halo_flags || agg->num_push) return KAGNN_EUNSUPPORTED;   // peer gather / two-part rows / in-flight halo / pushed output: pipelined kernel only
            rc = kagnn_fused_fwd_tc(agg, num_rows, pre, agg_out, ld_agg_out, n_layers, layers, post, y, ldy, stream);
            if (rc == KAGNN_OK) {
                g_count_tc.fetch_add(1);
                return rc;
            }
            if (rc != KAGNN_EUNSUPPORTED || mode == KAGNN_PATH_TC) return rc;
        } else if (mode == KAGNN_PATH_TC) {
            return KAGNN_EUNSUPPORTED;
        }
    } else if (mode == KAGNN_PATH_TC && n_layers >= 1 && num_rows > 0) {
        return KAGNN_EUNSUPPORTED;
    }
    if (n_layers == 0 && num_rows > 0 && mode != KAGNN_PATH_FP32) {
        // aggregation only: the high-MLP flattened-list gather (fused_tc2.cu) when the shape allows, else the general kernel
        int rc = kagnn_aggregate_only_tc2(agg, num_rows, pre, agg_out, ld_agg_out, stream);
        if (rc == KAGNN_OK) {
            g_count_agg.fetch_add(1);
            return rc;
        }
        if (rc != KAGNN_EUNSUPPORTED) return rc;
    }
    if (agg->peer_x || agg->num_head_cols || agg->halo_flags || agg->num_push) return KAGNN_EUNSUPPORTED;
    int rc = kagnn_fused_fwd_fp32(agg, num_rows, pre, agg_out, ld_agg_out, n_layers, layers, post, y, ldy, stream);
    if (rc == KAGNN_OK && num_rows > 0) g_count_fp32.fetch_add(1);
    return rc;
}
