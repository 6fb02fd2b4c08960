// kagnn_fused_layer_fwd -- tensor-core path (tcgen05 + TMEM + bulk TMA), sm_100a.
//
// Same contract as the fp32 kernel (fused_fp32.cu): tile = aggregate(x) -> pre-affine -> KAN chain -> post-affine -> y,
// one persistent launch per GNN layer.  What changes is how the contraction of a KAN layer
//      y[r, :] = sum_i sum_c B_c(x[r,i]) * Ws[:, i, c]  +  sum_i silu(x[r,i]) * Wb[:, i]
// (node_classification_clean/ekan.py:154-162; fastkan.py:76-85 for the RBF family) is executed:
//
//   * a CTA owns 128 destination rows (UMMA M = 128); thread t of each producer warpgroup owns row t;
//   * K is ordered [feature][8 slots] for the spline part (one 16-byte "k-core" per (row, feature)) followed, per
//     64-feature group, by the SiLU base part (one k-core per 8 features).  The producer warpgroups evaluate the
//     basis in registers (closed-form local de Boor -> the 4 non-zero cubic values are shifted into their slots with
//     two 64-bit funnel shifts; 8 Gaussians for FastKAN), split every value into bf16 hi + lo and write the two
//     16-byte vectors straight into the UMMA canonical K-major layout (tc_common.cuh) -- no (N, in, G+k) tensor,
//     no A-operand round trip through HBM;
//   * B (pre-packed once per weight update in exactly that layout) is streamed chunk by chunk with bulk TMA
//     (cp.async.bulk -> mbarrier complete_tx) into a multi-stage ring;
//   * one elected thread issues tcgen05.mma (bf16 x bf16 -> fp32 in TMEM) three times per K-step:
//     hi*hi + hi*lo + lo*hi, which keeps the result within ~2^-17 of fp32 (BASELINE.json asks for 1e-4);
//   * tcgen05.commit hands ring slots back to the producers and signals "accumulator complete";
//   * the next KAN layer of the chain reads its input rows directly from TMEM (tcgen05.ld, thread = row) -- chained
//     activations never touch shared or global memory; the last layer's epilogue applies bias / eval-BatchNorm and
//     stores with a leading dimension (column slice of the skip-concat buffer).
//
// Warp roles (320 threads): warps 0-7 = two producer warpgroups (gather + basis + epilogue), warp 8 = MMA issuer
// (+ TMEM alloc/dealloc), warp 9 = B loader.  TMEM: two accumulator regions, ping-ponged between chained layers.
#include "common.cuh"
#include "stage.cuh"
#include "tc_common.cuh"

namespace {

constexpr int BM = 128;
constexpr int NWG = 2;                    // producer warpgroups
constexpr int NPROD = NWG * 128;
constexpr int NTHREADS = NPROD + 64;
constexpr int A_HALF = 8 * 2048;          // bytes of A_hi (or A_lo) per stage: 8 k-cores x 128 rows x 16 B
constexpr int MAX_STAGES = 4;
constexpr float kSqrtLog2e = 1.2011224087864498f;

enum { SRC_SMEM = 0, SRC_GLOBAL = 1, SRC_TMEM = 2 };

struct LayerTC {
    int basis, F, F_pad, N, N_pad, G, k, n_chunks;
    float t0, h, inv_h, inv_den_l2;
    const float *bias, *lnw, *lnb;
    const uint8_t* wtc;
};

struct TcParams {
    StageParams st;
    long long num_rows;
    KagnnAffine post;
    int has_post, n_layers;
    float* agg_out;
    long long ld_agg_out;
    float* y;
    long long ldy;
    int src0, ld_s, n_tiles, vec, n_stage, stage_bytes, tmem_cols, tmem_r1, y_vec, xs_bytes;
    LayerTC layers[KAGNN_MAX_LAYERS];
};

__device__ __forceinline__ uint64_t shl64(uint64_t v, int s) {   // PTX semantics: shift amounts >= 64 give 0
    uint64_t r;
    asm("shl.b64 %0, %1, %2;" : "=l"(r) : "l"(v), "r"(s));
    return r;
}
__device__ __forceinline__ uint64_t shr64(uint64_t v, int s) {
    uint64_t r;
    asm("shr.b64 %0, %1, %2;" : "=l"(r) : "l"(v), "r"(s));
    return r;
}
__device__ __forceinline__ float fast_silu(float x) { return __fdividef(x, 1.0f + __expf(-x)); }
__device__ __forceinline__ float ex2_approx(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

// B-spline: the k+1 non-zero bases of x shifted into the 8-slot vector of this feature, as bf16 hi / lo.
__device__ __forceinline__ void bspline_slots(const LayerTC& L, float x, uint4& hi, uint4& lo) {
    const int k = L.k;
    const float u = (x - L.t0) * L.inv_h;
    const float fl = floorf(u);
    const float fr = u - fl;
    const bool valid = (u >= 0.0f) && (u < (float)(L.G + 2 * k));   // half-open knot range; NaN -> false (ekan.py:95)
    float b0 = 1.0f - fr, b1 = fr, b2 = 0.f, b3 = 0.f;              // degree 1
    if (k >= 2) {
        const float n0 = 0.5f * (1.0f - fr) * b0;
        const float n1 = 0.5f * ((fr + 1.0f) * b0 + (2.0f - fr) * b1);
        const float n2 = 0.5f * fr * b1;
        b0 = n0; b1 = n1; b2 = n2;
    }
    if (k >= 3) {
        const float t = 1.0f / 3.0f;
        const float n0 = t * (1.0f - fr) * b0;
        const float n1 = t * ((fr + 2.0f) * b0 + (2.0f - fr) * b1);
        const float n2 = t * ((fr + 1.0f) * b1 + (3.0f - fr) * b2);
        const float n3 = t * fr * b2;
        b0 = n0; b1 = n1; b2 = n2; b3 = n3;
    }
    uint32_t h01, l01, h23, l23;
    tc::split2(b0, b1, h01, l01);
    tc::split2(b2, b3, h23, l23);
    uint64_t vh = ((uint64_t)h23 << 32) | h01, vl = ((uint64_t)l23 << 32) | l01;
    int sh = 0;
    if (valid) {
        sh = 16 * ((int)fl - k);                                    // slot of b0 = interval index - k, may be < 0 or > 4
    } else {
        vh = 0; vl = 0;
    }
    const uint64_t hlo = sh >= 0 ? shl64(vh, sh) : shr64(vh, -sh);
    const uint64_t hhi = sh <= 64 ? shr64(vh, 64 - sh) : shl64(vh, sh - 64);
    const uint64_t llo = sh >= 0 ? shl64(vl, sh) : shr64(vl, -sh);
    const uint64_t lhi = sh <= 64 ? shr64(vl, 64 - sh) : shl64(vl, sh - 64);
    hi = make_uint4((uint32_t)hlo, (uint32_t)(hlo >> 32), (uint32_t)hhi, (uint32_t)(hhi >> 32));
    lo = make_uint4((uint32_t)llo, (uint32_t)(llo >> 32), (uint32_t)lhi, (uint32_t)(lhi >> 32));
    if (isinf(x)) {   // the reference's recursion yields NaN for every basis of an infinite input (inf * 0)
        hi = make_uint4(0x7fc07fc0u, 0x7fc07fc0u, 0x7fc07fc0u, 0x7fc07fc0u);
        lo = make_uint4(0u, 0u, 0u, 0u);
    }
}

// FastKAN: 8 Gaussians exp(-((z - g_i)/den)^2) of the layer-normalised input (fastkan.py:46-47)
__device__ __forceinline__ void rbf_slots(const LayerTC& L, float z, uint4& hi, uint4& lo) {
    float v[8];
#pragma unroll
    for (int g = 0; g < 8; ++g) {
        const float d = (z - (L.t0 + (float)g * L.h)) * L.inv_den_l2;
        v[g] = ex2_approx(-d * d);
    }
    tc::split8(v, hi, lo);
}

struct RowSource {
    int kind;
    const float* ptr;        // SMEM: row base in the tile; GLOBAL: row base in x (row clamped)
    uint32_t taddr;          // TMEM: region base + lane base
    const float* prev_bias;  // TMEM: bias of the producing layer (FastKAN base_linear.bias)
    int F;                   // valid columns
    bool row_valid, vec;
};

// 8 consecutive input features f0..f0+7 of this thread's row
__device__ __forceinline__ void load8(const RowSource& s, int f0, float (&v)[8]) {
    if (s.kind == SRC_TMEM) {
        tc::tmem_ld8(s.taddr + (uint32_t)f0, v);
        if (s.prev_bias) {
#pragma unroll
            for (int i = 0; i < 8; ++i)
                if (f0 + i < s.F) v[i] += __ldg(s.prev_bias + f0 + i);
        }
    } else if (s.kind == SRC_SMEM) {
        const float4 a = *reinterpret_cast<const float4*>(s.ptr + f0);
        const float4 b = *reinterpret_cast<const float4*>(s.ptr + f0 + 4);
        v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
    } else {
        if (!s.row_valid) {
#pragma unroll
            for (int i = 0; i < 8; ++i) v[i] = 0.f;
        } else if (s.vec && f0 + 8 <= s.F) {
            const float4 a = __ldg(reinterpret_cast<const float4*>(s.ptr + f0));
            const float4 b = __ldg(reinterpret_cast<const float4*>(s.ptr + f0 + 4));
            v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
        } else {
#pragma unroll
            for (int i = 0; i < 8; ++i) v[i] = (f0 + i < s.F) ? __ldg(s.ptr + f0 + i) : 0.f;
        }
    }
}

struct ChunkInfo {
    int group, j, n_oct, nk;
    bool base;
    uint32_t b_off, b_bytes;
};
__device__ __forceinline__ ChunkInfo chunk_info(const LayerTC& L, int q) {
    ChunkInfo c;
    c.group = q / 9;
    c.j = q - 9 * c.group;
    c.n_oct = min(8, L.F_pad / 8 - 8 * c.group);
    c.base = (c.j >= c.n_oct);
    c.nk = c.base ? c.n_oct : 8;
    c.b_off = (uint32_t)(c.group * 9 + c.j) * 256u * (uint32_t)L.N_pad;
    c.b_bytes = 32u * (uint32_t)c.nk * (uint32_t)L.N_pad;
    return c;
}

__device__ __forceinline__ void producers_sync() { asm volatile("bar.sync 1, %0;" ::"n"(NPROD) : "memory"); }

template <bool VEC>
__global__ void __launch_bounds__(NTHREADS, 1) fused_tc_kernel(const __grid_constant__ TcParams p) {
    extern __shared__ __align__(128) uint8_t smem[];
    float* xs = reinterpret_cast<float*>(smem);
    uint8_t* stages = smem + p.xs_bytes;
    uint64_t* bars = reinterpret_cast<uint64_t*>(stages + (size_t)p.n_stage * p.stage_bytes);
    uint64_t* full_a = bars;
    uint64_t* full_b = bars + MAX_STAGES;
    uint64_t* empty = bars + 2 * MAX_STAGES;
    uint64_t* acc_full = bars + 3 * MAX_STAGES;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 3 * MAX_STAGES + 1);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

    if (warp == NPROD / 32) tc::tmem_alloc(tmem_slot, (uint32_t)p.tmem_cols);
    if (tid == NPROD + 32) {
        for (int s = 0; s < MAX_STAGES; ++s) {
            tc::mbar_init(&full_a[s], 128);
            tc::mbar_init(&full_b[s], 1);
            tc::mbar_init(&empty[s], 1);
        }
        tc::mbar_init(acc_full, 1);
        tc::mbar_fence_init();
    }
    tc::tc_fence_before_sync();
    __syncthreads();
    tc::tc_fence_after_sync();
    const uint32_t tmem_base = *tmem_slot;

    if (warp < NPROD / 32) {
        // =================================== PRODUCERS / EPILOGUE =====================================
        const int wg = warp >> 2;
        const int row = tid & 127;
        const uint32_t lane_base = (uint32_t)((warp & 3) * 32) << 16;
        uint32_t cq = 0, acc_use = 0;
        for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x) {
            const long long row0 = (long long)tile * BM;
            const int nrows = (int)min((long long)BM, p.num_rows - row0);
            producers_sync();   // previous tile: every epilogue read of TMEM / use of the x tile is over
            if (p.src0 == SRC_SMEM) {
                const int F = p.st.agg.num_cols, Fp = p.layers[0].F_pad;
                for (int r = warp; r < BM; r += NPROD / 32) {
                    float* drow = xs + (size_t)r * p.ld_s;
                    if (r < nrows) {
                        stage_row<VEC>(p.st, row0 + r, drow, p.agg_out ? p.agg_out + (row0 + r) * p.ld_agg_out : nullptr, lane);
                        for (int c = F + lane; c < Fp; c += 32) drow[c] = 0.f;
                    } else {
                        for (int c = lane; c < Fp; c += 32) drow[c] = 0.f;
                    }
                }
                producers_sync();
            }
            for (int l = 0; l < p.n_layers; ++l) {
                const LayerTC& L = p.layers[l];
                RowSource src;
                src.F = L.F;
                src.vec = p.vec != 0;
                src.row_valid = row < nrows;
                src.prev_bias = nullptr;
                src.taddr = 0;
                src.ptr = nullptr;
                if (l == 0) {
                    src.kind = p.src0;
                    if (p.src0 == SRC_SMEM) src.ptr = xs + (size_t)row * p.ld_s;
                    else src.ptr = p.st.agg.x + (row0 + (src.row_valid ? row : 0)) * p.st.agg.ldx;
                } else {
                    tc::mbar_wait(acc_full, acc_use & 1);
                    ++acc_use;
                    tc::tc_fence_after_sync();
                    src.kind = SRC_TMEM;
                    src.taddr = tmem_base + lane_base + (((l - 1) & 1) ? (uint32_t)p.tmem_r1 : 0u);
                    src.prev_bias = p.layers[l - 1].bias;
                }
                const bool rbf = (L.basis == KAGNN_BASIS_RBF);
                float mean = 0.f, rstd = 1.f;
                if (rbf && L.lnw) {   // LayerNorm statistics of this thread's row (two-pass, biased variance, eps 1e-5)
                    float s = 0.f;
                    for (int f0 = 0; f0 < L.F_pad; f0 += 8) {
                        float v[8];
                        load8(src, f0, v);
#pragma unroll
                        for (int i = 0; i < 8; ++i) s += (f0 + i < L.F) ? v[i] : 0.f;
                    }
                    mean = s / (float)L.F;
                    float ss = 0.f;
                    for (int f0 = 0; f0 < L.F_pad; f0 += 8) {
                        float v[8];
                        load8(src, f0, v);
#pragma unroll
                        for (int i = 0; i < 8; ++i) {
                            const float d = (f0 + i < L.F) ? v[i] - mean : 0.f;
                            ss = fmaf(d, d, ss);
                        }
                    }
                    rstd = 1.0f / sqrtf(ss / (float)L.F + 1e-5f);
                }
                for (int q = 0; q < L.n_chunks; ++q, ++cq) {
                    if ((q % NWG) != wg) continue;
                    const int s = (int)(cq % (uint32_t)p.n_stage);
                    const uint32_t use = cq / (uint32_t)p.n_stage;
                    tc::mbar_wait(&empty[s], (use & 1u) ^ 1u);
                    uint8_t* a_hi = stages + (size_t)s * p.stage_bytes + row * 16;
                    uint8_t* a_lo = a_hi + A_HALF;
                    const ChunkInfo c = chunk_info(L, q);
                    if (!c.base) {
                        const int f0 = 64 * c.group + 8 * c.j;
                        float v[8];
                        load8(src, f0, v);
#pragma unroll
                        for (int i = 0; i < 8; ++i) {
                            uint4 hi, lo;
                            if (!rbf) {
                                bspline_slots(L, v[i], hi, lo);
                            } else {
                                float z = v[i];
                                if (L.lnw) {
                                    const int f = min(f0 + i, L.F - 1);
                                    z = (z - mean) * rstd * __ldg(L.lnw + f) + (L.lnb ? __ldg(L.lnb + f) : 0.f);
                                }
                                rbf_slots(L, z, hi, lo);
                            }
                            *reinterpret_cast<uint4*>(a_hi + i * 2048) = hi;
                            *reinterpret_cast<uint4*>(a_lo + i * 2048) = lo;
                        }
                    } else {
                        for (int jj = 0; jj < c.n_oct; ++jj) {
                            float v[8];
                            load8(src, 64 * c.group + 8 * jj, v);
#pragma unroll
                            for (int i = 0; i < 8; ++i) v[i] = fast_silu(v[i]);
                            uint4 hi, lo;
                            tc::split8(v, hi, lo);
                            *reinterpret_cast<uint4*>(a_hi + jj * 2048) = hi;
                            *reinterpret_cast<uint4*>(a_lo + jj * 2048) = lo;
                        }
                    }
                    tc::fence_proxy_async_smem();
                    tc::mbar_arrive(&full_a[s]);
                }
            }
            // ---- epilogue of the last layer: TMEM -> registers -> bias / post-affine -> y ----------------
            {
                const LayerTC& L = p.layers[p.n_layers - 1];
                tc::mbar_wait(acc_full, acc_use & 1);
                ++acc_use;
                tc::tc_fence_after_sync();
                const uint32_t taddr = tmem_base + lane_base + (((p.n_layers - 1) & 1) ? (uint32_t)p.tmem_r1 : 0u);
                float* yrow = p.y + (row0 + row) * p.ldy;
                for (int jb = wg; jb < L.N_pad / 8; jb += NWG) {
                    float v[8];
                    tc::tmem_ld8(taddr + (uint32_t)(8 * jb), v);
                    if (row < nrows) {
#pragma unroll
                        for (int i = 0; i < 8; ++i) {
                            const int col = 8 * jb + i;
                            if (col < L.N) {
                                if (L.bias) v[i] += __ldg(L.bias + col);
                                if (p.has_post) v[i] = apply_affine(p.post, col, v[i]);
                            }
                        }
                        if (p.y_vec && 8 * jb + 8 <= L.N) {
                            *reinterpret_cast<float4*>(yrow + 8 * jb) = make_float4(v[0], v[1], v[2], v[3]);
                            *reinterpret_cast<float4*>(yrow + 8 * jb + 4) = make_float4(v[4], v[5], v[6], v[7]);
                        } else {
#pragma unroll
                            for (int i = 0; i < 8; ++i)
                                if (8 * jb + i < L.N) yrow[8 * jb + i] = v[i];
                        }
                    }
                }
                tc::tc_fence_before_sync();
            }
        }
    } else if (warp == NPROD / 32) {
        // ========================================= MMA ISSUER =========================================
        if (lane == 0) {
            uint32_t cq = 0;
            for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x) {
                for (int l = 0; l < p.n_layers; ++l) {
                    const LayerTC& L = p.layers[l];
                    const uint32_t idesc = tc::idesc_bf16_f32(BM, L.N_pad);
                    const uint32_t d_tmem = tmem_base + ((l & 1) ? (uint32_t)p.tmem_r1 : 0u);
                    const uint32_t lbo_b = (uint32_t)L.N_pad * 32u;     // k-core slab = hi rows + lo rows
                    for (int q = 0; q < L.n_chunks; ++q, ++cq) {
                        const int s = (int)(cq % (uint32_t)p.n_stage);
                        const uint32_t use = cq / (uint32_t)p.n_stage;
                        const ChunkInfo c = chunk_info(L, q);
                        tc::mbar_wait(&full_a[s], use & 1u);
                        tc::mbar_wait(&full_b[s], use & 1u);
                        tc::tc_fence_after_sync();
                        const uint32_t a_hi = tc::smem_u32(stages + (size_t)s * p.stage_bytes);
                        const uint32_t a_lo = a_hi + A_HALF;
                        const uint32_t b_hi = a_hi + 2 * A_HALF;
                        const uint32_t b_lo = b_hi + (uint32_t)L.N_pad * 16u;
                        for (int kk = 0; kk < c.nk / 2; ++kk) {
                            const uint64_t dah = tc::smem_desc(a_hi + kk * 4096, 2048, 128);
                            const uint64_t dal = tc::smem_desc(a_lo + kk * 4096, 2048, 128);
                            const uint64_t dbh = tc::smem_desc(b_hi + kk * 2 * lbo_b, lbo_b, 128);
                            const uint64_t dbl = tc::smem_desc(b_lo + kk * 2 * lbo_b, lbo_b, 128);
                            tc::umma_bf16(d_tmem, dah, dbh, idesc, (q | kk) != 0 ? 1u : 0u);
                            tc::umma_bf16(d_tmem, dah, dbl, idesc, 1u);
                            tc::umma_bf16(d_tmem, dal, dbh, idesc, 1u);
                        }
                        tc::umma_commit(&empty[s]);   // slot reusable once these MMAs have read it
                    }
                    tc::umma_commit(acc_full);        // accumulator of layer l complete
                }
            }
        }
    } else {
        // ========================================== B LOADER ==========================================
        if (lane == 0) {
            uint32_t cq = 0;
            for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x) {
                for (int l = 0; l < p.n_layers; ++l) {
                    const LayerTC& L = p.layers[l];
                    for (int q = 0; q < L.n_chunks; ++q, ++cq) {
                        const int s = (int)(cq % (uint32_t)p.n_stage);
                        const uint32_t use = cq / (uint32_t)p.n_stage;
                        const ChunkInfo c = chunk_info(L, q);
                        tc::mbar_wait(&empty[s], (use & 1u) ^ 1u);
                        tc::mbar_arrive_expect_tx(&full_b[s], c.b_bytes);
                        tc::bulk_g2s(stages + (size_t)s * p.stage_bytes + 2 * A_HALF, L.wtc + c.b_off, c.b_bytes, &full_b[s]);
                    }
                }
            }
        }
    }
    tc::tc_fence_before_sync();
    __syncthreads();
    if (warp == NPROD / 32) tc::tmem_dealloc(tmem_base, (uint32_t)p.tmem_cols);
}

// ---------------------------------------------------------------------------------------------------
// Weight packing for the tensor-core path: per chunk [hi k-cores | lo k-cores], k-core = N_pad rows x 8 bf16.
// Chunk order = the order the kernel consumes them: per 64-feature group, 8 spline chunks (one feature octet
// each, k-core = the 8 slots of one feature) then one base chunk (k-core = base weights of 8 features).
// Folds scaled_spline_weight (ekan.py:146-152).
// ---------------------------------------------------------------------------------------------------
__global__ void pack_tc_kernel(const float* __restrict__ base_w, const float* __restrict__ spline_w,
                               const float* __restrict__ scaler, int F, int N, int S, int F_pad, int N_pad,
                               uint8_t* __restrict__ out) {
    // one thread per (k-core, n): spline k-cores 0..F_pad-1 (one per feature), base k-cores F_pad..F_pad+F_pad/8-1
    const int total_kc = F_pad + F_pad / 8;
    const long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (idx >= (long long)total_kc * N_pad) return;
    const int n = (int)(idx % N_pad);
    const int kc = (int)(idx / N_pad);
    float v[8];
    int group, j_chunk, kc_in_chunk;
    if (kc < F_pad) {
        const int f = kc;
        group = f / 64;
        j_chunk = (f % 64) / 8;
        kc_in_chunk = f % 8;
#pragma unroll
        for (int c = 0; c < 8; ++c) {
            float w = 0.f;
            if (f < F && n < N && c < S) {
                w = spline_w[((size_t)n * F + f) * S + c];
                if (scaler) w *= scaler[(size_t)n * F + f];
            }
            v[c] = w;
        }
    } else {
        const int o = kc - F_pad;          // feature octet
        group = o / 8;
        const int n_oct = min(8, F_pad / 8 - 8 * group);
        j_chunk = n_oct;                   // the base chunk follows the group's spline chunks
        kc_in_chunk = o % 8;
#pragma unroll
        for (int e = 0; e < 8; ++e) {
            const int f = 8 * o + e;
            v[e] = (base_w && f < F && n < N) ? base_w[(size_t)n * F + f] : 0.f;
        }
    }
    uint4 hi, lo;
    tc::split8(v, hi, lo);
    uint8_t* chunk = out + (size_t)(group * 9 + j_chunk) * 256u * (size_t)N_pad;
    // k-core slab kc of a chunk = [hi rows 0..N_pad-1 | lo rows 0..N_pad-1] (2 * N_pad * 16 bytes): the hi and lo halves of
    // one slab are adjacent, so [W_hi | W_lo] is also ONE UMMA B operand of 2 * N_pad rows (fused_tc2.cu stacks them along N)
    *reinterpret_cast<uint4*>(chunk + ((size_t)(2 * kc_in_chunk) * N_pad + n) * 16) = hi;
    *reinterpret_cast<uint4*>(chunk + ((size_t)(2 * kc_in_chunk + 1) * N_pad + n) * 16) = lo;
}

inline int ceil16(int v) { return (v + 15) & ~15; }

}  // namespace

extern "C" size_t kagnn_packed_weight_tc_bytes(int32_t in_f, int32_t out_f) {
    if (in_f <= 0 || out_f <= 0) return 0;
    const size_t F_pad = ceil16(in_f), N_pad = ceil16(out_f);
    const size_t groups = (F_pad + 63) / 64;
    return (F_pad / 8 + groups) * 256 * N_pad;   // every chunk slot is sized for 8 k-cores
}

extern "C" int kagnn_pack_kan_weights_tc(const float* base_w, const float* spline_w, const float* scaler, int32_t in_f,
                                         int32_t out_f, int32_t slots, void* packed, void* stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    if (in_f <= 0 || out_f <= 0 || slots <= 0 || !spline_w || !packed) return KAGNN_EINVAL;
    if (slots > 8 || out_f > 256) return KAGNN_EUNSUPPORTED;
    if (!aligned16(packed)) return KAGNN_EALIGN;
    const int F_pad = ceil16(in_f), N_pad = ceil16(out_f);
    KAGNN_CUDA_TRY(cudaMemsetAsync(packed, 0, kagnn_packed_weight_tc_bytes(in_f, out_f), stream));
    const long long total = (long long)(F_pad + F_pad / 8) * N_pad;
    pack_tc_kernel<<<(unsigned)ceil_div64(total, 256), 256, 0, stream>>>(base_w, spline_w, scaler, in_f, out_f, slots, F_pad,
                                                                         N_pad, static_cast<uint8_t*>(packed));
    KAGNN_LAUNCH_CHECK();
    return KAGNN_OK;
}

int kagnn_fused_fwd_tc(const KagnnAggregate* agg, int64_t num_rows, const KagnnAffine* pre, float* agg_out,
                       int64_t ld_agg_out, int32_t n_layers, const KagnnKanLayer* layers, const KagnnAffine* post, float* y,
                       int64_t ldy, cudaStream_t stream) {
    if (n_layers < 1 || n_layers > KAGNN_MAX_LAYERS) return KAGNN_EUNSUPPORTED;
    DeviceProps props{};
    int rc = kagnn_get_props(&props);
    if (rc != KAGNN_OK) return rc;
    if (props.cc_major != 10) return KAGNN_EUNSUPPORTED;

    TcParams p{};
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

    int width = agg->num_cols, n_max = 0;
    for (int l = 0; l < n_layers; ++l) {
        const KagnnKanLayer& s = layers[l];
        LayerTC& d = p.layers[l];
        if (!s.packed_w_tc || s.in_features != width || s.out_features <= 0 || s.out_features > 256) return KAGNN_EUNSUPPORTED;
        if (!aligned16(s.packed_w_tc)) return KAGNN_EALIGN;
        if (s.basis == KAGNN_BASIS_BSPLINE) {
            if (s.spline_order < 1 || s.spline_order > 3 || s.grid_size < 1 || s.grid_size + s.spline_order > 8) return KAGNN_EUNSUPPORTED;
            if (!(s.h > 0.f)) return KAGNN_EINVAL;
        } else if (s.basis == KAGNN_BASIS_RBF) {
            if (s.grid_size < 1 || s.grid_size > 8) return KAGNN_EUNSUPPORTED;
        } else {
            return KAGNN_EINVAL;
        }
        d.basis = s.basis;
        d.F = s.in_features;
        d.F_pad = ceil16(s.in_features);
        d.N = s.out_features;
        d.N_pad = ceil16(s.out_features);
        d.G = s.grid_size;
        d.k = s.basis == KAGNN_BASIS_BSPLINE ? s.spline_order : 0;
        d.n_chunks = d.F_pad / 8 + (d.F_pad + 63) / 64;
        d.t0 = s.t0;
        d.h = s.h;
        d.inv_h = s.h > 0.f ? 1.0f / s.h : 0.f;
        d.inv_den_l2 = s.inv_denominator * kSqrtLog2e;
        d.bias = s.base_bias;
        d.lnw = s.ln_weight;
        d.lnb = s.ln_bias;
        d.wtc = static_cast<const uint8_t*>(s.packed_w_tc);
        if (d.N_pad > n_max) n_max = d.N_pad;
        width = s.out_features;
    }
    if (ldy < width) return KAGNN_EINVAL;

    bool vec = (agg->num_cols % 4 == 0) && aligned16(agg->x) && (agg->ldx % 4 == 0);
    if (agg->mode == KAGNN_AGG_GINE) vec = vec && aligned16(agg->edge_feat) && (agg->ld_edge % 4 == 0);
    if (agg->x_halo) vec = vec && aligned16(agg->x_halo) && (agg->ld_halo % 4 == 0);
    p.vec = vec;
    p.y_vec = aligned16(y) && (ldy % 4 == 0);

    const bool direct = agg->mode == KAGNN_AGG_NONE && !pre && !agg_out && !agg->src_index;
    p.src0 = direct ? SRC_GLOBAL : SRC_SMEM;
    p.ld_s = direct ? 0 : p.layers[0].F_pad + 4;                 // (ld/4) odd -> conflict-free float4 row reads
    p.xs_bytes = direct ? 0 : (int)(((size_t)BM * p.ld_s * sizeof(float) + 127) & ~(size_t)127);
    p.stage_bytes = 2 * A_HALF + 256 * n_max;
    const int tail = (3 * MAX_STAGES + 2) * 8;
    int n_stage = ((int)props.max_smem - p.xs_bytes - tail) / p.stage_bytes;
    if (n_stage > MAX_STAGES) n_stage = MAX_STAGES;
    if (n_stage < 2) return KAGNN_EUNSUPPORTED;                  // x tile too wide: the caller falls back / splits
    p.n_stage = n_stage;
    const size_t smem = (size_t)p.xs_bytes + (size_t)n_stage * p.stage_bytes + tail;
    p.tmem_r1 = n_layers > 1 ? n_max : 0;
    p.tmem_cols = (int)tc::tmem_cols_pow2((uint32_t)(n_layers > 1 ? 2 * n_max : n_max));
    if (p.tmem_cols > 512) return KAGNN_EUNSUPPORTED;

    auto kern = vec ? fused_tc_kernel<true> : fused_tc_kernel<false>;
    KAGNN_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)props.max_smem));
    int grid = p.n_tiles < props.num_sms ? p.n_tiles : props.num_sms;
    kern<<<(unsigned)grid, NTHREADS, smem, stream>>>(p);
    KAGNN_LAUNCH_CHECK();
    return KAGNN_OK;
}
