// Library-level entry points and the weight packer.
#include "common.cuh"

int kagnn_get_props(DeviceProps* out) {
    static DeviceProps cache[64];
    static bool have[64] = {false};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return KAGNN_ECUDA;
    if (!have[dev]) {
        DeviceProps p{};
        if (cudaDeviceGetAttribute(&p.num_sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) return KAGNN_ECUDA;
        if (cudaDeviceGetAttribute(&p.max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev) != cudaSuccess) return KAGNN_ECUDA;
        if (cudaDeviceGetAttribute(&p.cc_major, cudaDevAttrComputeCapabilityMajor, dev) != cudaSuccess) return KAGNN_ECUDA;
        if (cudaDeviceGetAttribute(&p.cc_minor, cudaDevAttrComputeCapabilityMinor, dev) != cudaSuccess) return KAGNN_ECUDA;
        cache[dev] = p;
        have[dev] = true;
    }
    *out = cache[dev];
    return KAGNN_OK;
}

extern "C" int kagnn_version(void) { return KAGNN_VERSION; }

extern "C" const char* kagnn_strerror(int code) {
    switch (code) {
        case KAGNN_OK: return "ok";
        case KAGNN_EINVAL: return "invalid argument (shape, null pointer or inconsistent sizes)";
        case KAGNN_EUNSUPPORTED: return "configuration not supported by the sm_100a kernels";
        case KAGNN_EALIGN: return "pointer or leading dimension misaligned";
        case KAGNN_EWORKSPACE: return "workspace missing or too small";
        case KAGNN_ECUDA: return "CUDA runtime error or kernel launch failure";
        case KAGNN_EINDEX: return "graph index out of range";
        default: return "unknown kagnn error code";
    }
}

extern "C" int kagnn_device_info(int32_t* num_sms, int32_t* max_smem, int32_t* cc_major, int32_t* cc_minor) {
    DeviceProps p{};
    int rc = kagnn_get_props(&p);
    if (rc != KAGNN_OK) return rc;
    if (num_sms) *num_sms = p.num_sms;
    if (max_smem) *max_smem = p.max_smem;
    if (cc_major) *cc_major = p.cc_major;
    if (cc_minor) *cc_minor = p.cc_minor;
    return KAGNN_OK;
}

// ---------------------------------------------------------------------------------------------------
// Weight packing.  Folds ekan.KANLinear.scaled_spline_weight (node_classification_clean/ekan.py:146-152)
// and lays the (out,in,S) spline tensor + (out,in) base matrix out as [in][S+1][out_pad4] so that one
// K-chunk of the fused kernel is a contiguous block.
// ---------------------------------------------------------------------------------------------------
namespace {
__global__ void pack_kernel(const float* __restrict__ base_w, const float* __restrict__ spline_w,
                            const float* __restrict__ scaler, int in_f, int out_f, int slots, int out_pad,
                            float* __restrict__ packed) {
    int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    int64_t total = (int64_t)in_f * (slots + 1) * out_pad;
    if (idx >= total) return;
    int o = (int)(idx % out_pad);
    int64_t t = idx / out_pad;
    int c = (int)(t % (slots + 1));
    int i = (int)(t / (slots + 1));
    float v = 0.f;
    if (o < out_f) {
        if (c < slots) {
            v = spline_w[((int64_t)o * in_f + i) * slots + c];
            if (scaler) v *= scaler[(int64_t)o * in_f + i];
        } else {
            v = base_w ? base_w[(int64_t)o * in_f + i] : 0.f;
        }
    }
    packed[idx] = v;
}
}  // namespace

extern "C" size_t kagnn_packed_weight_elems(int32_t in_f, int32_t out_f, int32_t slots) {
    if (in_f <= 0 || out_f <= 0 || slots <= 0) return 0;
    return (size_t)in_f * (size_t)(slots + 1) * (size_t)pad4(out_f);
}

extern "C" int kagnn_pack_kan_weights(const float* base_w, const float* spline_w, const float* scaler, int32_t in_f,
                                      int32_t out_f, int32_t slots, float* packed, void* stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    if (in_f <= 0 || out_f <= 0 || slots <= 0 || !spline_w || !packed) return KAGNN_EINVAL;
    int64_t total = (int64_t)kagnn_packed_weight_elems(in_f, out_f, slots);
    pack_kernel<<<(unsigned)ceil_div64(total, 256), 256, 0, stream>>>(base_w, spline_w, scaler, in_f, out_f, slots,
                                                                      pad4(out_f), packed);
    KAGNN_LAUNCH_CHECK();
    return KAGNN_OK;
}
