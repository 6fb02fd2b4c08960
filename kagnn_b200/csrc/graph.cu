// Graph ingestion for the fused layer: COO -> destination-sorted CSR, segment pointers, gcn_norm,
// row gather.  Replaces the per-forward index handling of PyG's MessagePassing.propagate /
// gcn_norm (call sites: node_classification_clean/models.py:31-37,48-56) with device kernels whose
// results are cached per graph by the host side.
#include <cub/device/device_radix_sort.cuh>

#include "common.cuh"

namespace {

constexpr int kThreads = 256;

__global__ void coo_keys_kernel(const int64_t* __restrict__ src, const int64_t* __restrict__ dst, int64_t E,
                                int64_t n_dst, int64_t n_src, int32_t* __restrict__ keys,
                                int32_t* __restrict__ vals, int32_t* __restrict__ err) {
    int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (e >= E) return;
    int64_t s = src[e], d = dst[e];
    bool bad = (d < 0) | (d >= n_dst) | (s < 0) | (s >= n_src);
    if (bad) {
        atomicOr(err, 1);
        d = 0;
    }
    keys[e] = (int32_t)d;
    vals[e] = (int32_t)e;
}

__global__ void csr_fill_kernel(const int64_t* __restrict__ src, const int32_t* __restrict__ sorted_eid, int64_t E,
                                int64_t n_src, int32_t* __restrict__ col, int32_t* __restrict__ perm) {
    int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (p >= E) return;
    int32_t e = sorted_eid[p];
    int64_t s = src[e];
    if (s < 0 || s >= n_src) s = 0;  // flagged by coo_keys_kernel; keep reads in bounds
    col[p] = (int32_t)s;
    perm[p] = e;
}

// ptr[i] = first position p with sorted[p] >= i  (i = 0..n)
template <typename T>
__global__ void lower_bound_ptr_kernel(const T* __restrict__ sorted, int64_t len, int64_t n, int32_t* __restrict__ ptr) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i > n) return;
    int64_t lo = 0, hi = len;
    while (lo < hi) {
        int64_t mid = (lo + hi) >> 1;
        if ((int64_t)sorted[mid] < i) lo = mid + 1; else hi = mid;
    }
    ptr[i] = (int32_t)lo;
}

// one warp per destination row: deg = loop_w + sum_{col != i} w ; dinv = deg^-1/2 (0 if deg == 0)
__global__ void gcn_degree_kernel(const int32_t* __restrict__ rowptr, const int32_t* __restrict__ col, int64_t n,
                                  const float* __restrict__ w_in, float* __restrict__ dinv,
                                  float* __restrict__ self_w) {
    int64_t row = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    int lane = threadIdx.x & 31;
    if (row >= n) return;
    int beg = rowptr[row], end = rowptr[row + 1];
    float deg = 0.f;
    int last_loop = -1;
    for (int e = beg + lane; e < end; e += 32) {
        if (col[e] == (int32_t)row) last_loop = e;           // e grows per lane: keeps the last one
        else deg += w_in ? w_in[e] : 1.f;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        deg += __shfl_xor_sync(0xffffffffu, deg, o);
        last_loop = max(last_loop, __shfl_xor_sync(0xffffffffu, last_loop, o));
    }
    if (lane == 0) {
        // add_remaining_self_loops: an existing loop donates its weight, otherwise fill value 1
        float loop_w = (last_loop >= 0 && w_in) ? w_in[last_loop] : 1.f;
        deg += loop_w;
        float di = deg > 0.f ? 1.0f / sqrtf(deg) : 0.f;
        if (!(deg > 0.f) || isinf(di)) di = 0.f;
        dinv[row] = di;
        self_w[row] = di * loop_w * di;
    }
}

__global__ void gcn_edge_weight_kernel(const int32_t* __restrict__ rowptr, const int32_t* __restrict__ col, int64_t n,
                                       const float* __restrict__ w_in, const float* __restrict__ dinv,
                                       const float* __restrict__ dinv_dst, float* __restrict__ w_out) {
    int64_t row = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    int lane = threadIdx.x & 31;
    if (row >= n) return;
    int beg = rowptr[row], end = rowptr[row + 1];
    float di = dinv_dst[row];
    for (int e = beg + lane; e < end; e += 32) {
        int c = col[e];
        float w = w_in ? w_in[e] : 1.f;
        w_out[e] = (c == (int32_t)row) ? 0.f : dinv[c] * w * di;
    }
}

template <bool VEC>
__global__ void gather_rows_kernel(const float* __restrict__ x, int64_t ldx, const int32_t* __restrict__ index,
                                   int64_t rows, int cols, float* __restrict__ out, int64_t ld_out) {
    int64_t r = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    int lane = threadIdx.x & 31;
    if (r >= rows) return;
    const float* src = x + (int64_t)(index ? index[r] : (int32_t)r) * ldx;
    float* dst = out + r * ld_out;
    if (VEC) {
        for (int c = lane * 4; c < cols; c += 128)
            *reinterpret_cast<float4*>(dst + c) = __ldg(reinterpret_cast<const float4*>(src + c));
    } else {
        for (int c = lane; c < cols; c += 32) dst[c] = __ldg(src + c);
    }
}

// Halo pull for node-sharded graphs: out[r, :] = row (ids[r] % rows_per_rank) of the matrix of rank ids[r] / rows_per_rank,
// read in place from the owner over NVLink through a table of peer-mapped base pointers (warp per row, 128-bit loads).
__global__ void gather_rows_peer_kernel(const float* const* __restrict__ peer_x, int64_t ldx, int64_t rows_per_rank,
                                        const int32_t* __restrict__ ids, int64_t rows, int cols, float* __restrict__ out,
                                        int64_t ld_out, int nseg) {
    // ids are sorted by owner: consecutive warps take rows `seg` apart (a transposed walk over `nseg` segments), so the
    // warps in flight at any moment read from ALL peers instead of queueing on one NVLink egress port at a time
    const int64_t v = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    int lane = threadIdx.x & 31;
    const int64_t seg = (rows + nseg - 1) / nseg;
    const int64_t r = (v % nseg) * seg + v / nseg;
    if (v >= seg * nseg || r >= rows) return;
    const int32_t gid = ids[r];
    const int owner = (int)(gid / rows_per_rank);
    const float* src = peer_x[owner] + (gid - owner * rows_per_rank) * ldx;
    float* dst = out + r * ld_out;
    for (int c = lane * 4; c < cols; c += 128)
        *reinterpret_cast<float4*>(dst + c) = __ldg(reinterpret_cast<const float4*>(src + c));
}

// The pull without an id list: candidate row id in [0, total) is copied from its owner into out[id] (a replica of the whole matrix)
// when need[id] != 0.  The plan of the sharded forward marks the rows a shard references with one scatter, so no sort / unique /
// compaction (and no host synchronisation) stands between an edge list and the transfer.
__global__ void gather_rows_peer_masked_kernel(const float* const* __restrict__ peer_x, int64_t ldx, int64_t rows_per_rank,
                                               const uint8_t* __restrict__ need, int64_t total, int cols, float* __restrict__ out,
                                               int64_t ld_out, int nseg) {
    const int64_t v = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    const int64_t seg = (total + nseg - 1) / nseg;      // consecutive warps walk `nseg` segments in turn: every peer's link stays busy
    const int64_t gid = (v % nseg) * seg + v / nseg;
    if (v >= seg * nseg || gid >= total || !need[gid]) return;
    const int owner = (int)(gid / rows_per_rank);
    const float* src = peer_x[owner] + (gid - owner * rows_per_rank) * ldx;
    float* dst = out + gid * ld_out;
    for (int c = lane * 4; c < cols; c += 128)
        *reinterpret_cast<float4*>(dst + c) = __ldg(reinterpret_cast<const float4*>(src + c));
}

// The same pull as a persistent kernel that publishes its progress: rows are copied in the given order, 256 at a time per block
// (block b: chunks b, b + gridDim.x, ...), and a finished chunk stores `epoch` into its flag with release semantics.  The fused
// layer that runs concurrently (fused_tc2.cu, KagnnAggregate.halo_flags) waits on the flags of the prefix its tile needs.
constexpr int kHaloChunk = 256;
constexpr int kPullU = 16;                              // rows in flight per warp
constexpr int kPullWarps = KAGNN_PULL_WARPS;            // 16 warps x 16 rows = one chunk; the block also claims a whole SM (see below)
constexpr int kPullSmem = 180 * 1024;                   // dynamic shared memory requested only to keep other blocks off the SM
// Why whole-SM blocks: the fused layer that runs next to this kernel needs an entire SM per block (register file and shared
// memory), and the hardware spreads the blocks of a small kernel one per SM.  Pull blocks that each claim an SM (180 KB of shared
// memory they never touch) occupy exactly gridDim.x SMs and leave the others to the layer (KagnnAggregate.reserve_sms); light
// blocks would be scattered over 4x as many SMs and push a third of the layer's blocks into a second wave.  16 warps x 16 rows
// x 512 B per block keep ~2 MB in flight on 16 SMs -- the NVLink latency-bandwidth product.  No block-wide barrier: every warp
// adds 1 to its chunk's counter (release) when its rows are stored, so warps run ahead into later chunks; a chunk is complete
// when its counter reaches KAGNN_PULL_WARPS * epoch.
__global__ void __launch_bounds__(kPullWarps * 32, 1) gather_rows_peer_ordered_kernel(const float* const* __restrict__ peer_x, int64_t ldx,
                                                                                      int64_t rows_per_rank, const int32_t* __restrict__ ids,
                                                                                      int64_t rows, int cols, float* __restrict__ out,
                                                                                      int64_t ld_out, int32_t* __restrict__ flags) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t n_chunks = (rows + kHaloChunk - 1) / kHaloChunk;
    // Software pipeline per warp: the row ids of the NEXT chunk are fetched while the current chunk's rows are in flight, and a
    // chunk is published (release add on its counter) one iteration late, just before the next batch of loads is issued -- by
    // then its stores have long been acknowledged, so an iteration costs about one NVLink round trip and nothing else.
    int32_t gid[kPullU];
    auto fetch_ids = [&](int64_t c) {
        const int64_t r0 = c * kHaloChunk, r1 = min(rows, r0 + (int64_t)kHaloChunk);
#pragma unroll
        for (int u = 0; u < kPullU; ++u) {
            const int64_t r = r0 + warp + kPullWarps * u;
            gid[u] = (r < r1) ? ids[r] : -1;
        }
    };
    int64_t c = blockIdx.x, prev = -1;
    if (c < n_chunks) fetch_ids(c);
    for (; c < n_chunks; c += gridDim.x) {
        const int64_t r0 = c * kHaloChunk;
        const float* src[kPullU];
#pragma unroll
        for (int u = 0; u < kPullU; ++u) {
            src[u] = nullptr;
            if (gid[u] >= 0) {
                const int owner = (int)(gid[u] / rows_per_rank);
                src[u] = peer_x[owner] + (gid[u] - owner * rows_per_rank) * ldx;
            }
        }
        if (prev >= 0 && lane == 0) asm volatile("red.release.gpu.global.add.s32 [%0], 1;" ::"l"(flags + prev) : "memory");
        for (int cc = lane * 4; cc < cols; cc += 128) {
            float4 v[kPullU];
#pragma unroll
            for (int u = 0; u < kPullU; ++u)
                if (src[u]) v[u] = __ldg(reinterpret_cast<const float4*>(src[u] + cc));
            if (cc == lane * 4 && c + gridDim.x < n_chunks) fetch_ids(c + gridDim.x);      // overlaps the row loads above
#pragma unroll
            for (int u = 0; u < kPullU; ++u)
                if (src[u]) *reinterpret_cast<float4*>(out + (r0 + warp + kPullWarps * u) * ld_out + cc) = v[u];
        }
        __syncwarp();                                   // the warp's stores are ordered before lane 0's release below / next iteration
        prev = c;
    }
    if (prev >= 0 && lane == 0) asm volatile("red.release.gpu.global.add.s32 [%0], 1;" ::"l"(flags + prev) : "memory");
}

inline int bits_for(int64_t n) {
    int b = 1;
    while (b < 31 && ((int64_t)1 << b) < n) ++b;
    return b;
}

struct CsrWorkspace {
    int32_t* err;
    int32_t *keys_in, *keys_out, *vals_in, *vals_out;
    void* cub_tmp;
    size_t cub_bytes;
    size_t total;
};

inline size_t align256(size_t v) { return (v + 255) & ~(size_t)255; }

CsrWorkspace carve(void* base, int64_t E) {
    CsrWorkspace w{};
    size_t off = 0;
    char* b = static_cast<char*>(base);
    auto take = [&](size_t bytes) { char* p = b ? b + off : nullptr; off += align256(bytes); return p; };
    w.err = reinterpret_cast<int32_t*>(take(256));
    size_t eb = (size_t)(E > 0 ? E : 1) * sizeof(int32_t);
    w.keys_in = reinterpret_cast<int32_t*>(take(eb));
    w.keys_out = reinterpret_cast<int32_t*>(take(eb));
    w.vals_in = reinterpret_cast<int32_t*>(take(eb));
    w.vals_out = reinterpret_cast<int32_t*>(take(eb));
    size_t cub_bytes = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, cub_bytes, (const int32_t*)nullptr, (int32_t*)nullptr,
                                    (const int32_t*)nullptr, (int32_t*)nullptr, (int)(E > 0 ? E : 1), 0, 31);
    w.cub_bytes = cub_bytes;
    w.cub_tmp = take(cub_bytes);
    w.total = off;
    return w;
}

}  // namespace

extern "C" size_t kagnn_csr_build_workspace(int64_t num_edges, int64_t num_nodes) {
    (void)num_nodes;
    if (num_edges < 0) return 0;
    return carve(nullptr, num_edges).total;
}

extern "C" int kagnn_csr_build(const int64_t* edge_index, int64_t E, int64_t N, int64_t N_src, int32_t* rowptr,
                               int32_t* col, int32_t* perm, void* workspace, size_t workspace_bytes, void* stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    if (E < 0 || N < 0 || N_src < 0 || !rowptr) return KAGNN_EINVAL;
    if (E > 0 && (!edge_index || !col || !perm)) return KAGNN_EINVAL;
    if (E >= INT32_MAX || N >= INT32_MAX || N_src >= INT32_MAX) return KAGNN_EUNSUPPORTED;
    if (!workspace) return KAGNN_EWORKSPACE;
    CsrWorkspace w = carve(workspace, E);
    if (workspace_bytes < w.total) return KAGNN_EWORKSPACE;
    KAGNN_CUDA_TRY(cudaMemsetAsync(w.err, 0, sizeof(int32_t), stream));
    if (E == 0) {
        KAGNN_CUDA_TRY(cudaMemsetAsync(rowptr, 0, (size_t)(N + 1) * sizeof(int32_t), stream));
        return KAGNN_OK;
    }
    const int64_t* src = edge_index;
    const int64_t* dst = edge_index + E;
    unsigned blocks = (unsigned)ceil_div64(E, kThreads);
    coo_keys_kernel<<<blocks, kThreads, 0, stream>>>(src, dst, E, N, N_src, w.keys_in, w.vals_in, w.err);
    KAGNN_LAUNCH_CHECK();
    size_t cub_bytes = w.cub_bytes;
    KAGNN_CUDA_TRY(cub::DeviceRadixSort::SortPairs(w.cub_tmp, cub_bytes, w.keys_in, w.keys_out, w.vals_in, w.vals_out,
                                                   (int)E, 0, bits_for(N), stream));
    csr_fill_kernel<<<blocks, kThreads, 0, stream>>>(src, w.vals_out, E, N_src, col, perm);
    KAGNN_LAUNCH_CHECK();
    lower_bound_ptr_kernel<int32_t><<<(unsigned)ceil_div64(N + 1, kThreads), kThreads, 0, stream>>>(w.keys_out, E, N, rowptr);
    KAGNN_LAUNCH_CHECK();
    return KAGNN_OK;
}

extern "C" int kagnn_segment_ptr(const int64_t* batch, int64_t N, int64_t B, int32_t* ptr, void* stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    if (N < 0 || B < 0 || !ptr || (N > 0 && !batch)) return KAGNN_EINVAL;
    if (N >= INT32_MAX || B >= INT32_MAX) return KAGNN_EUNSUPPORTED;
    lower_bound_ptr_kernel<int64_t><<<(unsigned)ceil_div64(B + 1, kThreads), kThreads, 0, stream>>>(batch, N, B, ptr);
    KAGNN_LAUNCH_CHECK();
    return KAGNN_OK;
}

extern "C" int kagnn_gcn_degree(const int32_t* rowptr, const int32_t* col, int64_t N, const float* w_in, float* self_w,
                                float* dinv, void* stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    if (N < 0 || !rowptr || !self_w || !dinv) return KAGNN_EINVAL;
    if (N == 0) return KAGNN_OK;
    unsigned blocks = (unsigned)ceil_div64(N * 32, kThreads);
    gcn_degree_kernel<<<blocks, kThreads, 0, stream>>>(rowptr, col, N, w_in, dinv, self_w);
    KAGNN_LAUNCH_CHECK();
    return KAGNN_OK;
}

extern "C" int kagnn_gcn_edge_weight(const int32_t* rowptr, const int32_t* col, int64_t N, const float* w_in,
                                     const float* dinv_src, const float* dinv_dst, float* w_out, void* stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    if (N < 0 || !rowptr || !dinv_src || !dinv_dst) return KAGNN_EINVAL;
    if (N == 0) return KAGNN_OK;
    unsigned blocks = (unsigned)ceil_div64(N * 32, kThreads);
    gcn_edge_weight_kernel<<<blocks, kThreads, 0, stream>>>(rowptr, col, N, w_in, dinv_src, dinv_dst, w_out);
    KAGNN_LAUNCH_CHECK();
    return KAGNN_OK;
}

extern "C" int kagnn_gcn_norm(const int32_t* rowptr, const int32_t* col, int64_t N, const float* w_in, float* w_out,
                              float* self_w, float* dinv, void* stream_) {
    int rc = kagnn_gcn_degree(rowptr, col, N, w_in, self_w, dinv, stream_);
    if (rc != KAGNN_OK) return rc;
    return kagnn_gcn_edge_weight(rowptr, col, N, w_in, dinv, dinv, w_out, stream_);
}

extern "C" int kagnn_gather_rows(const float* x, int64_t ldx, const int32_t* index, int64_t rows, int32_t cols,
                                 float* out, int64_t ld_out, void* stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    if (rows < 0 || cols < 0 || (rows > 0 && (!x || !out))) return KAGNN_EINVAL;
    if (rows == 0 || cols == 0) return KAGNN_OK;
    unsigned blocks = (unsigned)ceil_div64(rows * 32, kThreads);
    bool vec = aligned16(x) && aligned16(out) && (ldx % 4 == 0) && (ld_out % 4 == 0) && (cols % 4 == 0);
    if (vec) gather_rows_kernel<true><<<blocks, kThreads, 0, stream>>>(x, ldx, index, rows, cols, out, ld_out);
    else gather_rows_kernel<false><<<blocks, kThreads, 0, stream>>>(x, ldx, index, rows, cols, out, ld_out);
    KAGNN_LAUNCH_CHECK();
    return KAGNN_OK;
}

// out[r, w * cols + c] = x[r, c] - w * shift  (w = 0 .. windows-1): the input of a B-spline layer with more than eight slots per
// feature, laid out as `windows` virtual features of eight slots each (uniform B-splines are shift invariant:
// B_{8 w + j}(x) = B_j(x - 8 w h); kagnn_b200/ekan.py: KANLinear.kernel_spec)
__global__ void expand_windows_kernel(const float* __restrict__ x, long long ldx, long long rows, int cols, int windows, float shift,
                                      float* __restrict__ out, long long ld_out) {
    const long long total = rows * (long long)cols;
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
        const long long r = e / cols;
        const int c = (int)(e - r * cols);
        const float v = __ldg(x + r * ldx + c);
        float* o = out + r * ld_out + c;
        for (int w = 0; w < windows; ++w) o[(long long)w * cols] = v - (float)w * shift;
    }
}

extern "C" int kagnn_expand_windows(const float* x, int64_t ldx, int64_t rows, int32_t cols, int32_t windows, float shift, float* out,
                                    int64_t ld_out, void* stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    if (rows < 0 || cols < 0 || windows < 1 || (rows > 0 && cols > 0 && (!x || !out))) return KAGNN_EINVAL;
    if (rows == 0 || cols == 0) return KAGNN_OK;
    if (ldx < cols || ld_out < (int64_t)cols * windows) return KAGNN_EINVAL;
    int64_t blocks = ceil_div64(rows * cols, kThreads);
    if (blocks > 148 * 16) blocks = 148 * 16;
    expand_windows_kernel<<<(unsigned)blocks, kThreads, 0, stream>>>(x, (long long)ldx, (long long)rows, cols, windows, shift, out,
                                                                     (long long)ld_out);
    KAGNN_LAUNCH_CHECK();
    return KAGNN_OK;
}

extern "C" int kagnn_gather_rows_peer(const float* const* peer_x, int64_t ldx, int64_t rows_per_rank, const int32_t* ids,
                                      int64_t rows, int32_t cols, float* out, int64_t ld_out, void* stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    if (rows < 0 || cols < 0 || rows_per_rank <= 0 || (rows > 0 && (!peer_x || !ids || !out))) return KAGNN_EINVAL;
    if (rows == 0 || cols == 0) return KAGNN_OK;
    if (!aligned16(out) || (ldx % 4) || (ld_out % 4) || (cols % 4)) return KAGNN_EALIGN;
    const int nseg = 16;
    const int64_t seg = ceil_div64(rows, nseg);
    unsigned blocks = (unsigned)ceil_div64(seg * nseg * 32, kThreads);
    gather_rows_peer_kernel<<<blocks, kThreads, 0, stream>>>(peer_x, ldx, rows_per_rank, ids, rows, cols, out, ld_out, nseg);
    KAGNN_LAUNCH_CHECK();
    return KAGNN_OK;
}

extern "C" int kagnn_gather_rows_peer_masked(const float* const* peer_x, int64_t ldx, int64_t rows_per_rank, const uint8_t* need,
                                             int64_t total_rows, int32_t cols, float* out, int64_t ld_out, void* stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    if (total_rows < 0 || cols < 0 || rows_per_rank <= 0 || (total_rows > 0 && (!peer_x || !need || !out))) return KAGNN_EINVAL;
    if (total_rows == 0 || cols == 0) return KAGNN_OK;
    if (!aligned16(out) || (ldx % 4) || (ld_out % 4) || (cols % 4)) return KAGNN_EALIGN;
    const int nseg = 16;
    const int64_t seg = ceil_div64(total_rows, nseg);
    unsigned blocks = (unsigned)ceil_div64(seg * nseg * 32, kThreads);
    gather_rows_peer_masked_kernel<<<blocks, kThreads, 0, stream>>>(peer_x, ldx, rows_per_rank, need, total_rows, cols, out, ld_out, nseg);
    KAGNN_LAUNCH_CHECK();
    return KAGNN_OK;
}

extern "C" int kagnn_gather_rows_peer_ordered(const float* const* peer_x, int64_t ldx, int64_t rows_per_rank, const int32_t* ids,
                                              int64_t rows, int32_t cols, float* out, int64_t ld_out, int32_t* chunk_flags,
                                              int32_t epoch, int32_t num_ctas, void* stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    if (rows < 0 || cols < 0 || rows_per_rank <= 0 || num_ctas <= 0 || (rows > 0 && (!peer_x || !ids || !out || !chunk_flags)))
        return KAGNN_EINVAL;
    if (rows == 0 || cols == 0) return KAGNN_OK;
    if (!aligned16(out) || (ldx % 4) || (ld_out % 4) || (cols % 4)) return KAGNN_EALIGN;
    (void)epoch;                                        // the counters are cumulative: a chunk of use `epoch` is complete at KAGNN_PULL_WARPS * epoch
    const int64_t n_chunks = ceil_div64(rows, kHaloChunk);
    const unsigned blocks = (unsigned)(n_chunks < num_ctas ? n_chunks : num_ctas);
    KAGNN_CUDA_TRY(cudaFuncSetAttribute(gather_rows_peer_ordered_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kPullSmem));
    gather_rows_peer_ordered_kernel<<<blocks, kPullWarps * 32, kPullSmem, stream>>>(peer_x, ldx, rows_per_rank, ids, rows, cols, out,
                                                                                    ld_out, chunk_flags);
    KAGNN_LAUNCH_CHECK();
    return KAGNN_OK;
}

// ---------------------------------------------------------------------------------------------------
// GAT attention (PyG GATConv with the KAN as shared projection: node_classification_clean/models.py:39-46,
// graph_classification/models.py:165-172).  h = lin(x) has heads * C columns.
//   gat_scores:        a_src[n,hd] = sum_c h[n, hd C + c] att_src[hd C + c], a_dst likewise                 (warp per row)
//   gat_edge_softmax:  per destination row i and head: e = leaky_relu(a_src[j] + a_dst[i]) over the CSR entries j != i and the
//                      one self loop PyG appends (existing self loops are removed first), alpha = exp(e - max) / (sum + 1e-16);
//                      written as edge weights [head][nnz] (0 for a removed self loop) and self weights [head][N], i.e. exactly
//                      the operands of the WEIGHTED aggregation (kagnn_fused_layer_fwd), which then runs once per head on the
//                      head's column slice of h.
// ---------------------------------------------------------------------------------------------------
namespace {
__global__ void gat_scores_kernel(const float* __restrict__ h, int64_t ldh, int64_t rows, int heads, int C,
                                  const float* __restrict__ att_src, const float* __restrict__ att_dst, float* __restrict__ a_src,
                                  float* __restrict__ a_dst) {
    const int64_t r = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (r >= rows) return;
    const float* hr = h + r * ldh;
    for (int hd = 0; hd < heads; ++hd) {
        float s = 0.f, d = 0.f;
        for (int c = lane; c < C; c += 32) {
            const float v = hr[hd * C + c];
            s = fmaf(v, att_src[hd * C + c], s);
            d = fmaf(v, att_dst[hd * C + c], d);
        }
        for (int o = 16; o; o >>= 1) {
            s += __shfl_xor_sync(0xffffffffu, s, o);
            d += __shfl_xor_sync(0xffffffffu, d, o);
        }
        if (lane == 0) {
            a_src[r * heads + hd] = s;
            a_dst[r * heads + hd] = d;
        }
    }
}

__device__ __forceinline__ float lrelu(float v, float slope) { return v > 0.f ? v : v * slope; }

__global__ void gat_edge_softmax_kernel(const int32_t* __restrict__ rowptr, const int32_t* __restrict__ col, int64_t rows, int heads,
                                        const float* __restrict__ a_src, const float* __restrict__ a_dst, float slope,
                                        float* __restrict__ edge_w, int64_t nnz, float* __restrict__ self_w) {
    const int64_t i = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (i >= rows) return;
    const int beg = rowptr[i], end = rowptr[i + 1];
    for (int hd = 0; hd < heads; ++hd) {
        const float ad = a_dst[i * heads + hd];
        const float e_self = lrelu(a_src[i * heads + hd] + ad, slope);
        float m = e_self;
        for (int e = beg + lane; e < end; e += 32) {
            const int j = col[e];
            if (j != (int)i) m = fmaxf(m, lrelu(a_src[(int64_t)j * heads + hd] + ad, slope));
        }
        for (int o = 16; o; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
        float sum = 0.f;
        for (int e = beg + lane; e < end; e += 32) {
            const int j = col[e];
            if (j != (int)i) sum += expf(lrelu(a_src[(int64_t)j * heads + hd] + ad, slope) - m);
        }
        for (int o = 16; o; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
        const float x_self = expf(e_self - m);
        const float inv = 1.0f / (sum + x_self + 1e-16f);
        for (int e = beg + lane; e < end; e += 32) {
            const int j = col[e];
            edge_w[(int64_t)hd * nnz + e] = (j != (int)i) ? expf(lrelu(a_src[(int64_t)j * heads + hd] + ad, slope) - m) * inv : 0.f;
        }
        if (lane == 0) self_w[(int64_t)hd * rows + i] = x_self * inv;
    }
}
}  // namespace

extern "C" int kagnn_gat_scores(const float* h, int64_t ldh, int64_t num_rows, int32_t heads, int32_t channels, const float* att_src,
                                const float* att_dst, float* a_src, float* a_dst, void* stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    if (num_rows < 0 || heads <= 0 || channels <= 0 || ldh < (int64_t)heads * channels) return KAGNN_EINVAL;
    if (num_rows > 0 && (!h || !att_src || !att_dst || !a_src || !a_dst)) return KAGNN_EINVAL;
    if (num_rows == 0) return KAGNN_OK;
    gat_scores_kernel<<<(unsigned)ceil_div64(num_rows * 32, kThreads), kThreads, 0, stream>>>(h, ldh, num_rows, heads, channels, att_src,
                                                                                              att_dst, a_src, a_dst);
    KAGNN_LAUNCH_CHECK();
    return KAGNN_OK;
}

extern "C" int kagnn_gat_edge_softmax(const int32_t* rowptr, const int32_t* col, int64_t num_rows, int64_t nnz, int32_t heads,
                                      const float* a_src, const float* a_dst, float negative_slope, float* edge_weight,
                                      float* self_weight, void* stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    if (num_rows < 0 || nnz < 0 || heads <= 0) return KAGNN_EINVAL;
    if (num_rows > 0 && (!rowptr || !a_src || !a_dst || !self_weight || (nnz > 0 && (!col || !edge_weight)))) return KAGNN_EINVAL;
    if (num_rows == 0) return KAGNN_OK;
    gat_edge_softmax_kernel<<<(unsigned)ceil_div64(num_rows * 32, kThreads), kThreads, 0, stream>>>(
        rowptr, col, num_rows, heads, a_src, a_dst, negative_slope, edge_weight, nnz, self_weight);
    KAGNN_LAUNCH_CHECK();
    return KAGNN_OK;
}

// ---------------------------------------------------------------------------------------------------
// Backward of the GAT attention (what autograd derives from PyG GATConv.forward): with
//   out[i,hd,:] = sum_j alpha_ij h[j,hd,:] (+ bias),  alpha = softmax_j(leaky_relu(a_src[j] + a_dst[i])),  a_* = <h, att_*>
// and d out given, the caller has already put  dh = sum_i alpha_ij d out[i]  (the aggregation over the reversed edges) into dh.
//   gat_bwd_edges (warp per destination row i, per head):  g_e = <d out[i], h[j]>,  s = sum_e alpha_e g_e,
//       d e = alpha_e (g_e - s),  d pre = d e * leaky_relu'(a_src[j] + a_dst[i]);
//       d a_dst[i] = sum_e d pre_e (row sum),  d a_src[j] += d pre_e (float atomics)
//   gat_bwd_proj:  dh[n,hd,:] += d a_src[n,hd] att_src[hd,:] + d a_dst[n,hd] att_dst[hd,:];
//       d att_src[hd,c] = sum_n d a_src[n,hd] h[n,hd,c],  d att_dst likewise
// ---------------------------------------------------------------------------------------------------
namespace {
__global__ void gat_bwd_edges_kernel(const int32_t* __restrict__ rowptr, const int32_t* __restrict__ col, int64_t rows, int64_t nnz, int heads,
                                     int C, const float* __restrict__ h, int64_t ldh, const float* __restrict__ dout, int64_t ldo,
                                     const float* __restrict__ a_src, const float* __restrict__ a_dst, const float* __restrict__ edge_w,
                                     const float* __restrict__ self_w, float slope, float* __restrict__ g_tmp, float* __restrict__ da_src,
                                     float* __restrict__ da_dst) {
    const int64_t i = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (i >= rows) return;
    const int beg = rowptr[i], end = rowptr[i + 1];
    for (int hd = 0; hd < heads; ++hd) {
        const float* d = dout + i * ldo + hd * C;
        auto dot_with = [&](int64_t j) {
            const float* hj = h + j * ldh + hd * C;
            float t = 0.f;
            for (int c = lane; c < C; c += 32) t = fmaf(d[c], hj[c], t);
            for (int o = 16; o; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
            return t;
        };
        const float al_s = self_w[(int64_t)hd * rows + i];
        const float g_s = dot_with(i);
        float s = al_s * g_s;
        for (int e = beg; e < end; ++e) {
            const int j = col[e];
            const float al = edge_w[(int64_t)hd * nnz + e];
            float g = 0.f;
            if (j != (int)i) g = dot_with(j);            // a removed self loop carries weight 0
            if (lane == 0) g_tmp[(int64_t)hd * nnz + e] = g;
            s = fmaf(al, g, s);
        }
        __syncwarp();
        const float ad = a_dst[i * heads + hd];
        float row = 0.f;
        for (int e = beg + lane; e < end; e += 32) {
            const int j = col[e];
            if (j == (int)i) continue;
            const float pre = a_src[(int64_t)j * heads + hd] + ad;
            const float dpre = edge_w[(int64_t)hd * nnz + e] * (g_tmp[(int64_t)hd * nnz + e] - s) * (pre > 0.f ? 1.0f : slope);
            atomicAdd(da_src + (int64_t)j * heads + hd, dpre);
            row += dpre;
        }
        for (int o = 16; o; o >>= 1) row += __shfl_xor_sync(0xffffffffu, row, o);
        if (lane == 0) {
            const float pre = a_src[i * heads + hd] + ad;
            const float dpre = al_s * (g_s - s) * (pre > 0.f ? 1.0f : slope);
            atomicAdd(da_src + i * heads + hd, dpre);
            da_dst[i * heads + hd] = row + dpre;
        }
    }
}

// block = a slab of rows, threads stride over the heads * C columns (coalesced); partial column sums by one atomic per thread
__global__ void gat_bwd_proj_kernel(const float* __restrict__ h, int64_t ldh, int64_t rows, int heads, int C, const float* __restrict__ att_src,
                                    const float* __restrict__ att_dst, const float* __restrict__ da_src, const float* __restrict__ da_dst,
                                    int64_t rows_per_block, float* __restrict__ dh, int64_t ld_dh, float* __restrict__ d_att_src,
                                    float* __restrict__ d_att_dst) {
    const int64_t r0 = blockIdx.x * rows_per_block, r1 = min(rows, r0 + rows_per_block);
    for (int c = threadIdx.x; c < heads * C; c += blockDim.x) {
        const int hd = c / C;
        const float ws = att_src[c], wd = att_dst[c];
        float ps = 0.f, pd = 0.f;
        for (int64_t r = r0; r < r1; ++r) {
            const float as = da_src[r * heads + hd], ad = da_dst[r * heads + hd], hv = h[r * ldh + c];
            dh[r * ld_dh + c] += fmaf(as, ws, ad * wd);
            ps = fmaf(as, hv, ps);
            pd = fmaf(ad, hv, pd);
        }
        atomicAdd(d_att_src + c, ps);
        atomicAdd(d_att_dst + c, pd);
    }
}
}  // namespace

extern "C" size_t kagnn_gat_bwd_workspace(int64_t num_rows, int64_t nnz, int32_t heads) {
    if (num_rows < 0 || nnz < 0 || heads <= 0) return 0;
    return sizeof(float) * ((size_t)heads * (size_t)nnz + (size_t)4 * (size_t)num_rows * (size_t)heads) + 64;
}

extern "C" int kagnn_gat_bwd(const int32_t* rowptr, const int32_t* col, int64_t num_rows, int64_t nnz, int32_t heads, int32_t channels,
                             const float* h, int64_t ldh, const float* dout, int64_t ld_dout, const float* att_src, const float* att_dst,
                             const float* edge_weight, const float* self_weight, float negative_slope, void* workspace,
                             size_t workspace_bytes, float* dh, int64_t ld_dh, float* d_att_src, float* d_att_dst, void* stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    const int64_t hc = (int64_t)heads * channels;
    if (num_rows < 0 || nnz < 0 || heads <= 0 || channels <= 0 || ldh < hc || ld_dout < hc || ld_dh < hc) return KAGNN_EINVAL;
    if (!d_att_src || !d_att_dst) return KAGNN_EINVAL;
    KAGNN_CUDA_TRY(cudaMemsetAsync(d_att_src, 0, sizeof(float) * (size_t)hc, stream));
    KAGNN_CUDA_TRY(cudaMemsetAsync(d_att_dst, 0, sizeof(float) * (size_t)hc, stream));
    if (num_rows == 0) return KAGNN_OK;
    if (!rowptr || !h || !dout || !att_src || !att_dst || !self_weight || !dh || (nnz > 0 && (!col || !edge_weight))) return KAGNN_EINVAL;
    if (!workspace || workspace_bytes < kagnn_gat_bwd_workspace(num_rows, nnz, heads)) return KAGNN_EWORKSPACE;
    float* g_tmp = static_cast<float*>(workspace);
    float* a_src = g_tmp + (size_t)heads * (size_t)nnz;
    float* a_dst = a_src + (size_t)num_rows * heads;
    float* da_src = a_dst + (size_t)num_rows * heads;
    float* da_dst = da_src + (size_t)num_rows * heads;
    gat_scores_kernel<<<(unsigned)ceil_div64(num_rows * 32, kThreads), kThreads, 0, stream>>>(h, ldh, num_rows, heads, channels, att_src,
                                                                                              att_dst, a_src, a_dst);
    KAGNN_LAUNCH_CHECK();
    KAGNN_CUDA_TRY(cudaMemsetAsync(da_src, 0, sizeof(float) * (size_t)num_rows * heads, stream));
    gat_bwd_edges_kernel<<<(unsigned)ceil_div64(num_rows * 32, kThreads), kThreads, 0, stream>>>(
        rowptr, col, num_rows, nnz, heads, channels, h, ldh, dout, ld_dout, a_src, a_dst, edge_weight, self_weight, negative_slope, g_tmp,
        da_src, da_dst);
    KAGNN_LAUNCH_CHECK();
    const int64_t rows_per_block = 256;
    gat_bwd_proj_kernel<<<(unsigned)ceil_div64(num_rows, rows_per_block), 128, 0, stream>>>(h, ldh, num_rows, heads, channels, att_src, att_dst,
                                                                                            da_src, da_dst, rows_per_block, dh, ld_dh,
                                                                                            d_att_src, d_att_dst);
    KAGNN_LAUNCH_CHECK();
    return KAGNN_OK;
}

// ---------------------------------------------------------------------------------------------------
// Row-wise log_softmax of the class logits (graph_classification/models.py:119,194) and training-mode
// BatchNorm1d (node_classification_clean/models.py:197 with model.train(): batch statistics over all
// rows, biased variance for the normalisation, unbiased for the running estimate, momentum update).
// Small HBM-bound helpers so that no step of a product forward runs as framework math.
// ---------------------------------------------------------------------------------------------------
namespace {
__global__ void log_softmax_kernel(const float* __restrict__ x, int64_t ldx, int64_t rows, int cols, float* __restrict__ y,
                                   int64_t ldy) {
    const int64_t r = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (r >= rows) return;
    const float* xr = x + r * ldx;
    float m = -INFINITY;
    for (int c = lane; c < cols; c += 32) m = fmaxf(m, xr[c]);
    for (int o = 16; o; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    float s = 0.f;
    for (int c = lane; c < cols; c += 32) s += expf(xr[c] - m);
    for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    const float lse = m + logf(s);
    for (int c = lane; c < cols; c += 32) y[r * ldy + c] = xr[c] - lse;
}

// column sums and sums of squares: each block owns a slab of rows, threads stride over columns (coalesced), fp64 partials
__global__ void bn_stats_kernel(const float* __restrict__ x, int64_t ldx, int64_t rows, int cols, int64_t rows_per_block,
                                double* __restrict__ sums) {
    const int64_t r0 = blockIdx.x * rows_per_block, r1 = min(rows, r0 + rows_per_block);
    for (int c = threadIdx.x; c < cols; c += blockDim.x) {
        double s = 0.0, q = 0.0;
        for (int64_t r = r0; r < r1; ++r) {
            const double v = (double)x[r * ldx + c];
            s += v;
            q += v * v;
        }
        atomicAdd(&sums[c], s);
        atomicAdd(&sums[cols + c], q);
    }
}

__global__ void bn_apply_kernel(const float* __restrict__ x, int64_t ldx, int64_t rows, int cols, const double* __restrict__ sums,
                                const float* __restrict__ weight, const float* __restrict__ bias, float eps, float momentum,
                                float* __restrict__ running_mean, float* __restrict__ running_var, int act, float* __restrict__ y,
                                int64_t ldy) {
    const int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    const int c = (int)(idx % cols);
    const int64_t r = idx / cols;
    if (r >= rows) return;
    const double mean = sums[c] / (double)rows;
    const double var = fmax(sums[cols + c] / (double)rows - mean * mean, 0.0);
    const float inv = (float)(1.0 / sqrt(var + (double)eps));
    float v = (x[r * ldx + c] - (float)mean) * inv;
    if (weight) v *= weight[c];
    if (bias) v += bias[c];
    if (act == KAGNN_ACT_SILU) v = v / (1.0f + expf(-v));
    y[r * ldy + c] = v;
    if (r == 0 && running_mean && running_var) {
        const double unbiased = rows > 1 ? var * (double)rows / (double)(rows - 1) : var;
        running_mean[c] = (1.0f - momentum) * running_mean[c] + momentum * (float)mean;
        running_var[c] = (1.0f - momentum) * running_var[c] + momentum * (float)unbiased;
    }
}
}  // namespace

extern "C" int kagnn_log_softmax_rows(const float* x, int64_t ldx, int64_t rows, int32_t cols, float* y, int64_t ldy, void* stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    if (rows < 0 || cols <= 0 || (rows > 0 && (!x || !y)) || ldx < cols || ldy < cols) return KAGNN_EINVAL;
    if (rows == 0) return KAGNN_OK;
    log_softmax_kernel<<<(unsigned)ceil_div64(rows * 32, kThreads), kThreads, 0, stream>>>(x, ldx, rows, cols, y, ldy);
    KAGNN_LAUNCH_CHECK();
    return KAGNN_OK;
}

namespace {
// The same two passes, tiled: block = a slab of rows x 32 columns, eight row lanes per column (coalesced 128-byte rows, several
// loads in flight per thread); column sums in fp64, joined through shared memory, one fp64 atomic pair per column and block.
__global__ void __launch_bounds__(256) bn_stats_tile_kernel(const float* __restrict__ x, int64_t ldx, int64_t rows, int cols, int64_t rows_per_block,
                                                            double* __restrict__ sums) {
    __shared__ double ps[8][33], pq[8][33];
    const int cl = threadIdx.x & 31, rl = threadIdx.x >> 5;
    const int c = blockIdx.y * 32 + cl;
    const int64_t r0 = blockIdx.x * rows_per_block, r1 = min(rows, r0 + rows_per_block);
    double s = 0.0, q = 0.0;
    if (c < cols) {
        for (int64_t r = r0 + rl; r < r1; r += 8) {
            const double v = (double)x[r * ldx + c];
            s += v;
            q += v * v;
        }
    }
    ps[rl][cl] = s;
    pq[rl][cl] = q;
    __syncthreads();
    if (rl == 0 && c < cols) {
        double a = 0.0, b = 0.0;
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            a += ps[k][cl];
            b += pq[k][cl];
        }
        atomicAdd(&sums[c], a);
        atomicAdd(&sums[cols + c], b);
    }
}

// normalise + affine + activation: the per-column constants are formed once per thread (fp64), the rows stream in fp32
__global__ void __launch_bounds__(256) bn_apply_tile_kernel(const float* __restrict__ x, int64_t ldx, int64_t rows, int cols,
                                                            const double* __restrict__ sums, const float* __restrict__ weight,
                                                            const float* __restrict__ bias, float eps, float momentum,
                                                            float* __restrict__ running_mean, float* __restrict__ running_var, int act,
                                                            int64_t rows_per_block, float* __restrict__ y, int64_t ldy) {
    const int cl = threadIdx.x & 31, rl = threadIdx.x >> 5;
    const int c = blockIdx.y * 32 + cl;
    if (c >= cols) return;
    const double mean_d = sums[c] / (double)rows;
    const double var = fmax(sums[cols + c] / (double)rows - mean_d * mean_d, 0.0);
    const float mean = (float)mean_d, inv = (float)(1.0 / sqrt(var + (double)eps));
    const float w = weight ? weight[c] : 1.0f, b = bias ? bias[c] : 0.0f;
    const int64_t r0 = blockIdx.x * rows_per_block, r1 = min(rows, r0 + rows_per_block);
    for (int64_t r = r0 + rl; r < r1; r += 8) {
        float v = (x[r * ldx + c] - mean) * inv;
        if (weight) v *= w;
        if (bias) v += b;
        if (act == KAGNN_ACT_SILU) v = v / (1.0f + expf(-v));
        y[r * ldy + c] = v;
    }
    if (blockIdx.x == 0 && rl == 0 && running_mean && running_var) {
        const double unbiased = rows > 1 ? var * (double)rows / (double)(rows - 1) : var;
        running_mean[c] = (1.0f - momentum) * running_mean[c] + momentum * mean;
        running_var[c] = (1.0f - momentum) * running_var[c] + momentum * (float)unbiased;
    }
}
}  // namespace

extern "C" size_t kagnn_batchnorm_train_workspace(int32_t cols) { return cols > 0 ? (size_t)cols * 2 * sizeof(double) : 0; }

extern "C" int kagnn_batchnorm_train_fwd(const float* x, int64_t ldx, int64_t rows, int32_t cols, const float* weight,
                                         const float* bias, float eps, float momentum, float* running_mean, float* running_var,
                                         int32_t act, float* y, int64_t ldy, void* workspace, size_t workspace_bytes, void* stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    if (rows <= 0 || cols <= 0 || !x || !y || ldx < cols || ldy < cols) return KAGNN_EINVAL;
    if (!workspace || workspace_bytes < kagnn_batchnorm_train_workspace(cols)) return KAGNN_EWORKSPACE;
    double* sums = static_cast<double*>(workspace);
    KAGNN_CUDA_TRY(cudaMemsetAsync(sums, 0, (size_t)cols * 2 * sizeof(double), stream));
    const int64_t rows_per_block = 512;
    const dim3 grid((unsigned)ceil_div64(rows, rows_per_block), (unsigned)((cols + 31) / 32), 1);
    if (grid.x > 0 && grid.y <= 65535) {
        bn_stats_tile_kernel<<<grid, 256, 0, stream>>>(x, ldx, rows, cols, rows_per_block, sums);
        KAGNN_LAUNCH_CHECK();
        bn_apply_tile_kernel<<<grid, 256, 0, stream>>>(x, ldx, rows, cols, sums, weight, bias, eps, momentum, running_mean, running_var, act,
                                                       rows_per_block, y, ldy);
        KAGNN_LAUNCH_CHECK();
        return KAGNN_OK;
    }
    bn_stats_kernel<<<(unsigned)ceil_div64(rows, 256), 128, 0, stream>>>(x, ldx, rows, cols, 256, sums);
    KAGNN_LAUNCH_CHECK();
    bn_apply_kernel<<<(unsigned)ceil_div64(rows * (int64_t)cols, kThreads), kThreads, 0, stream>>>(
        x, ldx, rows, cols, sums, weight, bias, eps, momentum, running_mean, running_var, act, y, ldy);
    KAGNN_LAUNCH_CHECK();
    return KAGNN_OK;
}

namespace {
// warp per row, two passes (the second one hits L1): exact two-pass variance like torch.nn.functional.layer_norm
__global__ void layernorm_stats_kernel(const float* __restrict__ x, int64_t ldx, int cols, const float* __restrict__ xh, int64_t ldh,
                                       int hcols, int64_t rows, float eps, float* __restrict__ stats) {
    const int64_t r = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (r >= rows) return;
    const float* a = xh ? xh + r * ldh : nullptr;
    const float* b = x + r * ldx;
    const int total = hcols + cols;
    float s = 0.f;
    for (int c = lane; c < total; c += 32) s += c < hcols ? a[c] : b[c - hcols];
    for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    const float mean = s / (float)total;
    float q = 0.f;
    for (int c = lane; c < total; c += 32) {
        const float d = (c < hcols ? a[c] : b[c - hcols]) - mean;
        q = fmaf(d, d, q);
    }
    for (int o = 16; o; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
    if (lane == 0) {
        stats[2 * r] = mean;
        stats[2 * r + 1] = 1.0f / sqrtf(q / (float)total + eps);
    }
}
}  // namespace

extern "C" int kagnn_layernorm_stats(const float* x, int64_t ldx, int32_t cols, const float* x_head, int64_t ld_head, int32_t hcols,
                                     int64_t rows, float eps, float* stats, void* stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    if (rows < 0 || cols <= 0 || hcols < 0 || !x || !stats || ldx < cols || (hcols > 0 && (!x_head || ld_head < hcols))) return KAGNN_EINVAL;
    if (rows == 0) return KAGNN_OK;
    layernorm_stats_kernel<<<(unsigned)ceil_div64(rows * 32, kThreads), kThreads, 0, stream>>>(x, ldx, cols, hcols > 0 ? x_head : nullptr,
                                                                                                ld_head, hcols, rows, eps, stats);
    KAGNN_LAUNCH_CHECK();
    return KAGNN_OK;
}
