// tcgen05 backward of the B-spline KAN layer: the two GEMM-shaped gradients of
// KANLinear.forward (ekan.py:154-162) on the tensor cores, with the same bf16 hi/lo split (three products, fp32 accumulate
// in tensor memory) that keeps the forward inside the fp32 parity bound.  Packed layouts as in backward_tiled.cu:
// P[i][c][o], c = 0..S-1 spline slots, c = S the SiLU base column, o padded to a multiple of 4.
//
//   dX   dx[n,i] = sum_c D[n,i,c] T[n,(i,c)],   T = dY . P^T          D = d/dx of (B_0..B_{S-1}, silu)(x[n,i])
//        One GEMM per 128-row tile and pass of 16 features: M = 128 rows, K = out, N = 16 (S+1).  A = the dY tile, split and
//        written to tensor memory by its own rows (thread = row = TMEM lane, exactly the forward's A path); B = the 16 (S+1)
//        weight rows of the pass: the layer's tensor-core packing read as an MN-major operand (see PACKED below), or split on
//        the fly from the fp32 packed weights into the K-major canonical layout in shared memory; the epilogue reads the (S+1) dot products of a (row, feature) back from
//        tensor memory and contracts them with the derivative vector it evaluates -- the (N, in (S+1)) intermediate never
//        reaches HBM.
//   dW   dP[(i,c),o] = sum_n E[n,(i,c)] dY[n,o]                        E = the S + 1 function values
//        The reduction runs over ROWS, so both operands are MN-major: thread = row writes, for each feature, its eight slot
//        values as ONE 16-byte vector -- which is precisely the MN-major no-swizzle core-matrix layout ([unit of 8 M][row][8]),
//        conflict-free.  M = 128 = 14 features x 8 slots + 2 units of base values, N = out, K = 128 rows per step; the
//        accumulator stays in tensor memory across all row tiles of a slab and is added to HBM once (one float atomic per
//        element and slab).  Needs S <= 8 (every configuration the forward's pipelined kernel takes).
//
// Kernels in this file (kagnn_set_backward_path picks between the variants for the tests):
//   kan_bwd_input_tc_look_kernel   dX, layers up to 64 wide with the packed-weight operand (default there): passes of eight
//                                  features, two T buffers, 8 worker warps + a control warp (weight loads, MMA issue one pass
//                                  ahead) + a loader warp (next tile's dY rows by cp.async)
//   kan_bwd_input_tc_kernel        dX, every other shape: 256 threads, phases one after the other, two CTAs per SM overlap each
//                                  other; B from the packed weights (bulk copies) or split per tile from the fp32 weights
//   kan_bwd_weights_tcm_kernel     dW (default): one CTA per SM, 16 producer warps + control warp + 4 loader warps, as many feature
//                                  blocks per CTA as tensor memory has accumulators for, batches of 64 rows (<= 64 wide) or 32 rows
//   kan_bwd_weights_tc64_kernel    dW, <= 64 wide, one feature block per CTA, 64-row batches, three CTAs per SM
//   kan_bwd_weights_tc_kernel      dW, one feature block per CTA, 128-row batches (also: a wide layer with a single feature block)
// The K = 0 instantiations are the FastKAN (RBF) layer.  Shapes outside the limits return KAGNN_EUNSUPPORTED and the caller
// continues with backward_tiled.cu.
#include <atomic>
#include <cstdlib>

#include "common.cuh"
#include "tc_common.cuh"

namespace {
constexpr int kMaxOrderB = 3;           // closed forms for k = 1, 2, 3 (the orders the forward's tensor-core kernels take)
constexpr int kS1MaxB = 9;              // S + 1 <= 9: G + k <= 8
constexpr int kThreads = 256;           // two warpgroups: both cover the 128 rows (TMEM lanes) of a tile and split its features
constexpr float kMagicB = 12582912.0f;  // 1.5 * 2^23: u + kMagic (round down) = floor(u) + kMagic
constexpr float kLog2eB = 1.4426950408889634f;

struct GeomB {
    int in_f, out_f, out_pad, G, k, S;       // S = slots per feature (B-spline: G + k, RBF: G); slot S = the SiLU base column
    float t0, inv_h;
    // RBF layers (template K == 0; fastkan.py:76-85): centres c0 + g step, phi_g(z) = exp(-((z - c_g) inv_den)^2), z = LayerNorm(x)
    float c0, step, inv_den;
    const float *ln_w, *ln_b, *stats;        // stats = per-row (mean, rstd) or NULL: no LayerNorm, z = x
};

// RBF: the layer's input after its LayerNorm
__device__ __forceinline__ float rbf_z_b(const GeomB& g, float xv, float mean, float rstd, int f) {
    if (!g.stats) return xv;
    float z = (xv - mean) * rstd;
    if (g.ln_w) z *= __ldg(g.ln_w + f);
    if (g.ln_b) z += __ldg(g.ln_b + f);
    return z;
}

__device__ __forceinline__ float ex2_b(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ uint32_t prmt_b(uint32_t a, uint32_t b, uint32_t sel) {      // full PTX semantics (sign replication)
    uint32_t d;
    asm("prmt.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(sel));
    return d;
}
__device__ __forceinline__ uint32_t pack_trunc_b(float lo_elem, float hi_elem) { return prmt_b(__float_as_uint(lo_elem), __float_as_uint(hi_elem), 0x7632u); }
__device__ __forceinline__ float trunc_res_b(float v) { return v - __uint_as_float(__float_as_uint(v) & 0xffff0000u); }
__device__ __forceinline__ uint32_t pack_rn_b(float lo_elem, float hi_elem) {
    uint32_t r;
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi_elem), "f"(lo_elem));
    return r;
}

// knot interval of one input value: u = (x - t0) / h clamped into [-0.5, lim + 0.5] (out-of-range inputs and NaN land in the
// sentinel intervals -1 / lim, which carry no basis), idx = floor(u) + 1 in 0..lim+1, fr = fractional position in [0, 1)
__device__ __forceinline__ void locate_b(const GeomB& g, float x, int& idx, float& fr) {
    const float lim = (float)(g.G + 2 * g.k);
    const float u = fminf(fmaxf((x - g.t0) * g.inv_h, -0.5f), lim + 0.5f);
    const float t = __fadd_rd(u, kMagicB);
    fr = u - (t - kMagicB);
    idx = __float_as_int(t) - (0x4B400000 - 1);
}

// the k + 1 basis values of the interval (ekan.py:79-112 restricted to its live bases: local polynomials of fr)
template <int K>
__device__ __forceinline__ void local_values_b(float fr, float (&b)[4]) {
    b[2] = 0.f;
    b[3] = 0.f;
    if (K == 3) {
        const float omf = 1.0f - fr, f2 = fr * fr;
        b[0] = omf * omf * (omf * (1.0f / 6.0f));
        b[3] = f2 * (fr * (1.0f / 6.0f));
        b[1] = fmaf(f2, fmaf(fr, 0.5f, -1.0f), 2.0f / 3.0f);
        b[2] = fmaf(fr, fmaf(fr, fmaf(fr, -0.5f, 0.5f), 0.5f), 1.0f / 6.0f);
    } else if (K == 2) {
        const float omf = 1.0f - fr;
        b[0] = 0.5f * omf * omf;
        b[2] = 0.5f * fr * fr;
        b[1] = fmaf(fr, omf, 0.5f);
    } else {
        b[0] = 1.0f - fr;
        b[1] = fr;
    }
}
// their derivatives with respect to x: B'_{j,k} = (B_{j,k-1} - B_{j+1,k-1}) / h
template <int K>
__device__ __forceinline__ void local_derivs_b(float fr, float inv_h, float (&d)[4]) {
    d[2] = 0.f;
    d[3] = 0.f;
    if (K == 3) {
        const float omf = 1.0f - fr;
        const float q0 = 0.5f * omf * omf, q2 = 0.5f * fr * fr, q1 = fmaf(fr, omf, 0.5f);
        d[0] = -q0 * inv_h;
        d[1] = (q0 - q1) * inv_h;
        d[2] = (q1 - q2) * inv_h;
        d[3] = q2 * inv_h;
    } else if (K == 2) {
        const float omf = 1.0f - fr;
        d[0] = -omf * inv_h;
        d[1] = (omf - fr) * inv_h;
        d[2] = fr * inv_h;
    } else {
        d[0] = -inv_h;
        d[1] = inv_h;
    }
}

// selector row of the slot placement for interval row q (interval j = q - 1): output byte b of the 16-byte slot vector takes
// source byte b - 2 (j - K) of (v0 v1 | v2 v3) when that is in 0..7, else the replicated (zero) sign bit of source byte 1
__device__ __forceinline__ uint4 lut_row_b(int q, int K, int lim) {
    const int j = q - 1;
    const bool inside = j >= 0 && j < lim;
    uint32_t w[4];
    for (int m = 0; m < 4; ++m) {
        uint32_t sel = 0;
        for (int n = 0; n < 4; ++n) {
            const int src = 4 * m + n - 2 * (j - K);
            const uint32_t nib = (inside && src >= 0 && src <= 7) ? (uint32_t)src : 9u;
            sel |= nib << (4 * n);
        }
        w[m] = sel;
    }
    return make_uint4(w[0], w[1], w[2], w[3]);
}
constexpr int kLutRows = 13;

// MN-major no-swizzle operand: units of 8 consecutive M (or N) elements; element (unit u, k-row r) at u * 2048 + r * 16 bytes for
// a 128-row K extent, i.e. SBO (unit stride) = 2048, LBO (stride between groups of 8 k-rows) = 128
__host__ __device__ __forceinline__ uint32_t idesc_bf16_f32_mn(int m, int n) { return tc::idesc_bf16_f32(m, n) | (1u << 15) | (1u << 16); }

// 16- / 4-byte asynchronous copies into shared memory; src-size 0 zero-fills the destination and reads nothing
__device__ __forceinline__ void cp_async16(void* dst_smem, const void* src, bool on) {
    const uint32_t n = on ? 16u : 0u;                        // src-size 0: the destination is zero-filled, nothing is read
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(tc::smem_u32(dst_smem)), "l"(src), "r"(n) : "memory");
}
__device__ __forceinline__ void cp_async_f32(void* dst_smem, const float* src, bool on) {
    const uint32_t n = on ? 4u : 0u;
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(tc::smem_u32(dst_smem)), "l"(src), "r"(n) : "memory");
}
// the mbarrier gets one arrival from this thread once all its earlier cp.async copies have landed
__device__ __forceinline__ void cp_async_arrive_on(uint64_t* bar) {
    asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(tc::smem_u32(bar)) : "memory");
}
// named barriers of the warp-specialised kernels: producers ARRIVE and go on, the control warp SYNCs (ids 1 .. 3; 0 is __syncthreads)
template <int ID, int COUNT>
__device__ __forceinline__ void named_arrive() { asm volatile("bar.arrive %0, %1;" ::"n"(ID), "n"(COUNT) : "memory"); }
template <int ID, int COUNT>
__device__ __forceinline__ void named_sync() { asm volatile("bar.sync %0, %1;" ::"n"(ID), "n"(COUNT) : "memory"); }

// =====================================================================================================================
// dX
// =====================================================================================================================
constexpr int kFP = 16;                 // features per pass

// PACKED: B comes from the layer's tensor-core weight packing (kagnn_pack_kan_weights_tc, the forward's operand): a slab of that
// layout -- N_pad output rows x 16 bytes holding the eight slots of one feature -- read as an MN-major operand IS the unit of
// eight (feature, slot) rows x out that dX needs, so a pass's weights are one bulk copy (two spline chunks = 16 features) plus
// two slabs of the group's SiLU chunk, double-buffered under the previous pass; nothing is converted per tile.  T columns:
// feature i of the pass at [8 i, 8 i + 8), its base product at 128 + i.
template <int K, bool PACKED>
__global__ void __launch_bounds__(kThreads) kan_bwd_input_tc_kernel(GeomB g, const float* __restrict__ w, const uint8_t* __restrict__ wtc,
                                                                    int n_stage, const float* __restrict__ x,
                                                                    long long ldx, const float* __restrict__ dy, long long ld_dy,
                                                                    long long n_rows, int n_tiles, int KK, uint32_t tmem_cols,
                                                                    float* __restrict__ dx, long long ld_dx, float* __restrict__ dxb,
                                                                    long long ld_dxb) {
    extern __shared__ __align__(128) uint8_t smem[];
    const int S1 = PACKED ? 9 : g.S + 1;                 // columns per feature in T
    const int Np = kFP * S1;                             // N of one pass
    const int kcs = KK / 8;
    const uint32_t b_bytes = (uint32_t)kcs * (uint32_t)Np * 16u;     // one of hi / lo (on-the-fly conversion)
    const uint32_t stage_bytes = 576u * (uint32_t)KK;                 // PACKED: two spline chunks (512 N_pad) + two base slabs (64 N_pad)
    uint8_t* b_hi = smem;
    uint8_t* b_lo = smem + b_bytes;
    uint64_t* bar = reinterpret_cast<uint64_t*>(PACKED ? smem + (size_t)n_stage * stage_bytes : b_lo + b_bytes);
    uint64_t* full = bar + 1;                            // PACKED: weights of a stage have landed
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar + 3);
    const int tid = threadIdx.x, warp = tid >> 5, wg = warp >> 2, r128 = tid & 127;
    const uint32_t a_col = (uint32_t)((Np + 31) & ~31);

    if (warp == 0) tc::tmem_alloc(tmem_slot, tmem_cols);
    if (tid == 32) {
        tc::mbar_init(bar, 1);
        tc::mbar_init(&full[0], 1);
        tc::mbar_init(&full[1], 1);
        tc::mbar_fence_init();
    }
    tc::tc_fence_before_sync();
    __syncthreads();
    tc::tc_fence_after_sync();
    const uint32_t tmem_base = *tmem_slot;
    const uint32_t lane_base = (uint32_t)((warp & 3) * 32) << 16;
    const uint32_t idesc = tc::idesc_bf16_f32(128, Np);
    const uint32_t lbo_b = (uint32_t)Np * 16u;
    const int n_pass = (g.in_f + kFP - 1) / kFP;
    const bool w_vec = (g.out_pad % 4 == 0) && ((reinterpret_cast<uintptr_t>(w) & 15u) == 0);
    const bool dx_vec = (ld_dx % 4 == 0) && ((reinterpret_cast<uintptr_t>(dx) & 15u) == 0);
    const bool dy_vec = g.out_f % 4 == 0 && ld_dy % 4 == 0 && ((reinterpret_cast<uintptr_t>(dy) & 15u) == 0);
    uint32_t phase = 0;
    // PACKED: running pass counter over all tiles of this CTA -> stage and phase of the weight ring; one thread issues the copies
    uint32_t qi = 0;
    const int octs_all = ((g.in_f + 15) / 16) * 2;       // feature octets of the packed layout (F_pad / 8)
    auto load_pass = [&](int pass, uint32_t q) {
        const int f0 = pass * kFP, grp = f0 >> 6, j = (f0 & 63) >> 3;
        const int n_oct = min(8, octs_all - 8 * grp);
        const uint32_t st = (n_stage > 1) ? (q & 1u) : 0u;
        uint8_t* dst = smem + (size_t)st * stage_bytes;
        const uint8_t* chunk = wtc + (size_t)(grp * 9 + j) * 256u * (size_t)KK;
        const uint8_t* base = wtc + (size_t)(grp * 9 + n_oct) * 256u * (size_t)KK + (size_t)j * 32u * (size_t)KK;
        tc::mbar_arrive_expect_tx(&full[st], stage_bytes);
        tc::bulk_g2s(dst, chunk, 512u * (uint32_t)KK, &full[st]);
        tc::bulk_g2s(dst + 512u * (uint32_t)KK, base, 64u * (uint32_t)KK, &full[st]);
    };
    if (PACKED && tid == 0 && blockIdx.x < n_tiles) load_pass(0, 0);

    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const long long row = (long long)tile * 128 + r128;
        const bool row_ok = row < n_rows;
        float mean = 0.f, rstd = 1.f;
        if (K == 0 && g.stats && row_ok) {
            mean = __ldg(g.stats + 2 * row);
            rstd = __ldg(g.stats + 2 * row + 1);
        }
        // ---- A = this row of dY, split into bf16 hi / lo, into tensor memory (lane = row; hi at a_col, lo at a_col + KK/2);
        // the two warpgroups take alternate 8-column groups
        const float* dyr = dy + (row_ok ? row : 0) * ld_dy;
        for (int kb = wg; kb < kcs; kb += 8) {          // four 8-column groups per round: their loads are in flight together
            float4 q[4][2];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int kc = kb + 2 * j;
                const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
                if (dy_vec) {
                    q[j][0] = (row_ok && kc < kcs && kc * 8 + 4 <= g.out_f) ? __ldg(reinterpret_cast<const float4*>(dyr + kc * 8)) : z;
                    q[j][1] = (row_ok && kc < kcs && kc * 8 + 8 <= g.out_f) ? __ldg(reinterpret_cast<const float4*>(dyr + kc * 8 + 4)) : z;
                } else {
                    float v[8];
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const int o = kc * 8 + i;
                        v[i] = (row_ok && kc < kcs && o < g.out_f) ? __ldg(dyr + o) : 0.f;
                    }
                    q[j][0] = make_float4(v[0], v[1], v[2], v[3]);
                    q[j][1] = make_float4(v[4], v[5], v[6], v[7]);
                }
            }
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int kc = kb + 2 * j;
                if (kc < kcs) {
                    const float v[8] = {q[j][0].x, q[j][0].y, q[j][0].z, q[j][0].w, q[j][1].x, q[j][1].y, q[j][1].z, q[j][1].w};
                    uint4 hi, lo;
                    tc::split8(v, hi, lo);
                    tc::tmem_st4(tmem_base + lane_base + a_col + 4u * kc, hi.x, hi.y, hi.z, hi.w);
                    tc::tmem_st4(tmem_base + lane_base + a_col + (uint32_t)KK / 2 + 4u * kc, lo.x, lo.y, lo.z, lo.w);
                }
            }
        }
        tc::tmem_st_wait();
        for (int pass = 0; pass < n_pass; ++pass) {
            const int f0 = pass * kFP;
            // ---- B = the 16 (S+1) weight rows of the pass (row n = i (S+1) + c <-> P[f0+i][c][.]), split into the canonical
            // K-major layout: slab kc = Np rows x 8 bf16
            for (int base = tid; !PACKED && base < kcs * Np; base += 4 * kThreads) {     // four items per round: their loads are in flight together
                float4 q[4][2];
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const int idx = base + j * kThreads;
                    const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
                    q[j][0] = z;
                    q[j][1] = z;
                    if (idx < kcs * Np) {
                        const int kc = idx / Np, n = idx - kc * Np;
                        if (f0 + n / S1 < g.in_f) {
                            const float* wr = w + ((long long)f0 * S1 + n) * g.out_pad + kc * 8;
                            if (w_vec && kc * 8 + 8 <= g.out_pad) {
                                q[j][0] = __ldg(reinterpret_cast<const float4*>(wr));
                                q[j][1] = __ldg(reinterpret_cast<const float4*>(wr) + 1);
                            } else {
                                float v[8];
#pragma unroll
                                for (int i = 0; i < 8; ++i) v[i] = (kc * 8 + i < g.out_pad) ? __ldg(wr + i) : 0.f;
                                q[j][0] = make_float4(v[0], v[1], v[2], v[3]);
                                q[j][1] = make_float4(v[4], v[5], v[6], v[7]);
                            }
                        }
                    }
                }
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const int idx = base + j * kThreads;
                    if (idx < kcs * Np) {
                        const int kc = idx / Np;
                        float v[8] = {q[j][0].x, q[j][0].y, q[j][0].z, q[j][0].w, q[j][1].x, q[j][1].y, q[j][1].z, q[j][1].w};
#pragma unroll
                        for (int i = 0; i < 8; ++i)
                            if (kc * 8 + i >= g.out_f) v[i] = 0.f;
                        uint4 hi, lo;
                        tc::split8(v, hi, lo);
                        *reinterpret_cast<uint4*>(b_hi + (size_t)idx * 16) = hi;
                        *reinterpret_cast<uint4*>(b_lo + (size_t)idx * 16) = lo;
                    }
                }
            }
            tc::fence_proxy_async_smem();
            tc::tc_fence_before_sync();
            __syncthreads();
            if (tid == 0 && PACKED) {
                // the next pass's weights (same tile or the CTA's next tile) fly while this pass computes
                const bool more = (pass + 1 < n_pass) || (tile + (int)gridDim.x < n_tiles);
                if (n_stage > 1 && more) load_pass(pass + 1 < n_pass ? pass + 1 : 0, qi + 1);
                const uint32_t st = (n_stage > 1) ? (qi & 1u) : 0u;
                tc::mbar_wait(&full[st], (n_stage > 1) ? ((qi >> 1) & 1u) : (qi & 1u));
                tc::tc_fence_after_sync();
                const uint32_t sb = tc::smem_u32(smem) + st * stage_bytes;
                const uint32_t sbo = 32u * (uint32_t)KK, lo_off = 16u * (uint32_t)KK;
                const uint32_t idesc_s = tc::idesc_bf16_f32(128, 128) | (1u << 16), idesc_b = tc::idesc_bf16_f32(128, 16) | (1u << 16);
                uint32_t acc = 0;
                for (int ks = 0; ks < KK / 16; ++ks) {
                    const uint32_t off = (uint32_t)ks * 256u;
                    const uint64_t dsh = tc::smem_desc(sb + off, 128, sbo), dsl = tc::smem_desc(sb + lo_off + off, 128, sbo);
                    const uint64_t dbh = tc::smem_desc(sb + 512u * (uint32_t)KK + off, 128, sbo);
                    const uint64_t dbl = tc::smem_desc(sb + 512u * (uint32_t)KK + lo_off + off, 128, sbo);
                    const uint32_t tah = tmem_base + a_col + 8u * ks, tal = tah + (uint32_t)KK / 2;
                    tc::umma_bf16_ts(tmem_base, tah, dsh, idesc_s, acc);
                    tc::umma_bf16_ts(tmem_base, tah, dsl, idesc_s, 1);
                    tc::umma_bf16_ts(tmem_base, tal, dsh, idesc_s, 1);
                    tc::umma_bf16_ts(tmem_base + 128u, tah, dbh, idesc_b, acc);
                    tc::umma_bf16_ts(tmem_base + 128u, tah, dbl, idesc_b, 1);
                    tc::umma_bf16_ts(tmem_base + 128u, tal, dbh, idesc_b, 1);
                    acc = 1;
                }
                tc::umma_commit(bar);
                if (n_stage == 1 && more) {
                    // single stage (256-wide layers): the next copy may start only when these MMAs have read the stage
                    tc::mbar_wait(bar, phase);
                    load_pass(pass + 1 < n_pass ? pass + 1 : 0, qi + 1);
                }
            }
            if (tid == 0 && !PACKED) {
                tc::tc_fence_after_sync();
                uint32_t acc = 0;
                for (int ks = 0; ks < KK / 16; ++ks) {
                    const uint64_t dbh = tc::smem_desc(tc::smem_u32(b_hi) + ks * 2 * lbo_b, lbo_b, 128);
                    const uint64_t dbl = tc::smem_desc(tc::smem_u32(b_lo) + ks * 2 * lbo_b, lbo_b, 128);
                    const uint32_t tah = tmem_base + a_col + 8u * ks, tal = tah + (uint32_t)KK / 2;
                    tc::umma_bf16_ts(tmem_base, tah, dbh, idesc, acc);
                    tc::umma_bf16_ts(tmem_base, tah, dbl, idesc, 1);
                    tc::umma_bf16_ts(tmem_base, tal, dbh, idesc, 1);
                    acc = 1;
                }
                tc::umma_commit(bar);
            }
            ++qi;
            // this warpgroup's x values of the pass: requested before the wait, so the loads fly while the tensor pipe works
            const float* xr = x + (row_ok ? row : 0) * ldx;
            float xq[8];
#pragma unroll
            for (int ii = 0; ii < 8; ++ii) {
                const int f = f0 + 8 * wg + ii;
                xq[ii] = (row_ok && f < g.in_f) ? __ldg(xr + f) : 0.f;
            }
            tc::mbar_wait(bar, phase);
            phase ^= 1u;
            tc::tc_fence_after_sync();
            // ---- epilogue: warpgroup wg contracts features 8 wg .. 8 wg + 7 of the pass for its 128 rows
            float res[8], resb[8], tbase[8];
#pragma unroll
            for (int c = 0; c < 8; ++c) tbase[c] = 0.f;
            if (PACKED) tc::tmem_ld8(tmem_base + lane_base + 128u + 8u * (uint32_t)wg, tbase);
#pragma unroll
            for (int ii = 0; ii < 8; ++ii) {
                const int i = 8 * wg + ii, f = f0 + i;
                res[ii] = 0.f;
                resb[ii] = 0.f;
                if (f < g.in_f) {                       // uniform over the warpgroup
                    float t[16];
                    if (PACKED) {
                        float t0[8];
                        tc::tmem_ld8(tmem_base + lane_base + (uint32_t)(8 * i), t0);
#pragma unroll
                        for (int c = 0; c < 8; ++c) t[c] = t0[c];
#pragma unroll
                        for (int c = 8; c < 16; ++c) t[c] = 0.f;
                    } else {
                        float t0[8];
                        tc::tmem_ld8(tmem_base + lane_base + (uint32_t)(i * S1), t0);
#pragma unroll
                        for (int c = 0; c < 8; ++c) t[c] = t0[c];
                        tc::tmem_ld8(tmem_base + lane_base + (uint32_t)(i * S1 + 8), t0);   // only column S (<= 8) of these is used
#pragma unroll
                        for (int c = 0; c < 8; ++c) t[8 + c] = t0[c];
                    }
                    const float xv = xq[ii];
                    float a = 0.f;
                    if (K == 0) {
                        // RBF: every slot is live; phi_g'(z) = -2 (z - c_g) inv_den^2 phi_g
                        const float z = rbf_z_b(g, xv, mean, rstd, f);
#pragma unroll
                        for (int q = 0; q < 8; ++q) {
                            const float tt = (z - (g.c0 + (float)q * g.step)) * g.inv_den;
                            const float dq = -2.0f * tt * g.inv_den * ex2_b(-kLog2eB * tt * tt);
                            a = fmaf((q < g.S) ? dq : 0.f, t[q], a);
                        }
                    } else {
                        int idx;
                        float fr, d[4];
                        locate_b(g, xv, idx, fr);
                        local_derivs_b<(K == 0 ? 1 : K)>(fr, g.inv_h, d);
                        const int j = idx - 1;              // interval; its bases sit on slots j - K .. j
                        const bool live = j >= 0 && j < g.G + 2 * K;
                        // window of K + 1 consecutive slots starting at j - K (slots outside 0..S-1 count as zero): barrel shift of the
                        // slot array, padded with K zeros in front, by jj = j in 0..10
                        float tp[20];
#pragma unroll
                        for (int c = 0; c < 20; ++c) tp[c] = 0.f;
#pragma unroll
                        for (int c = 0; c < 8; ++c)
                            if (c < kS1MaxB - 1) tp[K + c] = (c < g.S) ? t[c] : 0.f;
                        const int jj = live ? j : 0;
                        float s0[12], s1[8], s2[6], s3[4];
#pragma unroll
                        for (int c = 0; c < 12; ++c) s0[c] = (jj & 8) ? tp[c + 8] : tp[c];
#pragma unroll
                        for (int c = 0; c < 8; ++c) s1[c] = (jj & 4) ? s0[c + 4] : s0[c];
#pragma unroll
                        for (int c = 0; c < 6; ++c) s2[c] = (jj & 2) ? s1[c + 2] : s1[c];
#pragma unroll
                        for (int c = 0; c < 4; ++c) s3[c] = (jj & 1) ? s2[c + 1] : s2[c];
#pragma unroll
                        for (int r = 0; r <= K; ++r) a = fmaf(d[r], s3[r], a);
                        if (!live) a = 0.f;
                    }
                    const float sg = __fdividef(1.0f, 1.0f + ex2_b(-kLog2eB * xv));
                    const float dbase = sg * fmaf(xv, 1.0f - sg, 1.0f);
                    float tb = 0.f;
                    if (PACKED) {
                        tb = tbase[ii];
                    } else {
#pragma unroll
                        for (int c = 0; c < 16; ++c)
                            if (c == g.S) tb = t[c];
                    }
                    if (K == 0 && dxb) {                    // with a LayerNorm the two branches stay separate (kagnn_layernorm_bwd joins them)
                        res[ii] = a;
                        resb[ii] = dbase * tb;
                        continue;
                    }
                    res[ii] = fmaf(dbase, tb, a);
                }
            }
            if (row_ok) {
                float* dxr = dx + row * ld_dx + f0 + 8 * wg;
                if (dx_vec && f0 + 8 * wg + 8 <= g.in_f) {
                    *reinterpret_cast<float4*>(dxr) = make_float4(res[0], res[1], res[2], res[3]);
                    *reinterpret_cast<float4*>(dxr + 4) = make_float4(res[4], res[5], res[6], res[7]);
                } else {
#pragma unroll
                    for (int ii = 0; ii < 8; ++ii)
                        if (f0 + 8 * wg + ii < g.in_f) dxr[ii] = res[ii];
                }
                if (K == 0 && dxb) {
                    float* br = dxb + row * ld_dxb + f0 + 8 * wg;
#pragma unroll
                    for (int ii = 0; ii < 8; ++ii)
                        if (f0 + 8 * wg + ii < g.in_f) br[ii] = resb[ii];
                }
            }
            tc::tc_fence_before_sync();
        }
        __syncthreads();                               // every thread is done with tensor memory before the next tile's A lands
    }
    tc::tc_fence_before_sync();
    __syncthreads();
    if (warp == 0) tc::tmem_dealloc(tmem_base, tmem_cols);
}

// ---------------------------------------------------------------------------------------------------------------------
// dX with look-ahead, for layers up to 64 outputs wide with the packed-weight operand: passes of EIGHT features and two T
// buffers in tensor memory (2 x (64 spline + 16 base columns), A behind them: 256 columns, two CTAs per SM as before), so the
// MMAs of pass p + 1 are issued before the epilogue of pass p and complete under it; the barrier in front of an issue only says
// "everybody is done with the epilogue of pass p - 1".  Weights arrive as before: stages of 16 features (= two passes), a ring
// of two, the load of chunk n + 2 issued when the last MMA of chunk n has completed.
// ---------------------------------------------------------------------------------------------------------------------
constexpr int kFPL = 8;
constexpr uint32_t kLookT = 96, kLookBase = 64, kLookA = 192;
constexpr int kLookSync = kThreads + 32;                     // workers + control warp: the named barriers of the T buffers
constexpr int kLookThreads = kThreads + 64;                  // + the loader warp (barrier 3 is theirs too)

template <int K>
__global__ void __launch_bounds__(kLookThreads, 2) kan_bwd_input_tc_look_kernel(GeomB g, const uint8_t* __restrict__ wtc, const float* __restrict__ x,
                                                                         long long ldx, const float* __restrict__ dy, long long ld_dy,
                                                                         long long n_rows, int n_tiles, int KK, uint32_t tmem_cols,
                                                                         float* __restrict__ dx, long long ld_dx, float* __restrict__ dxb,
                                                                         long long ld_dxb) {
    extern __shared__ __align__(128) uint8_t smem[];
    const int kcs = KK / 8;
    const uint32_t stage_bytes = 576u * (uint32_t)KK;    // two spline chunks (16 features) + two base slabs
    uint64_t* bar_t = reinterpret_cast<uint64_t*>(smem + 2 * (size_t)stage_bytes);     // [2]: the MMAs into T buffer b are done
    uint64_t* full = bar_t + 2;                                                         // [2]: weights of a stage have landed
    uint64_t* bar_dy = full + 2;                                                        // the staged dY tile has landed
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar_dy + 1);
    float* d_st = reinterpret_cast<float*>(smem + 2 * (size_t)stage_bytes + 128);       // [128][KK + 4]: the tile's dY rows (dy16 only)
    const int dld = KK + 4;
    const int tid = threadIdx.x, warp = tid >> 5, wg = warp >> 2, r128 = tid & 127;
    const bool control = warp == kThreads / 32;          // the ninth warp: weight loads and MMA issue
    const bool loader = warp == kThreads / 32 + 1;       // the tenth: stages the next tile's dY rows

    if (warp == 0) tc::tmem_alloc(tmem_slot, tmem_cols);
    if (tid == 32) {
        for (int i = 0; i < 4; ++i) tc::mbar_init(&bar_t[i], 1);
        tc::mbar_init(bar_dy, 32);
        tc::mbar_fence_init();
    }
    tc::tc_fence_before_sync();
    __syncthreads();
    tc::tc_fence_after_sync();
    const uint32_t tmem_base = *tmem_slot;
    const uint32_t lane_base = (uint32_t)((warp & 3) * 32) << 16;
    const int n_pass = (g.in_f + kFPL - 1) / kFPL, chunks = (g.in_f + 15) / 16;
    const bool dx_vec = (ld_dx % 4 == 0) && ((reinterpret_cast<uintptr_t>(dx) & 15u) == 0);
    const bool dy_vec = g.out_f % 4 == 0 && ld_dy % 4 == 0 && ((reinterpret_cast<uintptr_t>(dy) & 15u) == 0);
    const bool x_vec = g.in_f % 4 == 0 && ldx % 4 == 0 && ((reinterpret_cast<uintptr_t>(x) & 15u) == 0);
    const int my_tiles = (int)blockIdx.x < n_tiles ? (n_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;
    const uint32_t total_chunks = (uint32_t)my_tiles * (uint32_t)chunks;
    const int octs_all = chunks * 2;                     // feature octets of the packed layout (F_pad / 8)
    auto load_chunk = [&](uint32_t n) {                  // running chunk n -> stage n & 1
        const int f0 = (int)(n % (uint32_t)chunks) * 16, grp = f0 >> 6, j = (f0 & 63) >> 3;
        const int n_oct = min(8, octs_all - 8 * grp);
        const uint32_t st = n & 1u;
        uint8_t* dst = smem + (size_t)st * stage_bytes;
        const uint8_t* chunk = wtc + (size_t)(grp * 9 + j) * 256u * (size_t)KK;
        const uint8_t* base = wtc + (size_t)(grp * 9 + n_oct) * 256u * (size_t)KK + (size_t)j * 32u * (size_t)KK;
        tc::mbar_arrive_expect_tx(&full[st], stage_bytes);
        tc::bulk_g2s(dst, chunk, 512u * (uint32_t)KK, &full[st]);
        tc::bulk_g2s(dst + 512u * (uint32_t)KK, base, 64u * (uint32_t)KK, &full[st]);
    };
    if (control && (tid & 31) == 0) {
        if (total_chunks > 0) load_chunk(0);
        if (total_chunks > 1) load_chunk(1);
    }
    // one pass's MMAs (thread 0): half h of the stage's 16 features into T buffer qq & 1, its chunk's 16 base products next to them
    auto issue = [&](int p_local, uint32_t qq, uint32_t cc) {
        const uint32_t st = cc & 1u;
        if ((p_local & 1) == 0) tc::mbar_wait(&full[st], (cc >> 1) & 1u);
        tc::tc_fence_after_sync();
        const uint32_t sb = tc::smem_u32(smem) + st * stage_bytes + (uint32_t)(p_local & 1) * 256u * (uint32_t)KK;
        const uint32_t bb = tc::smem_u32(smem) + st * stage_bytes + 512u * (uint32_t)KK;
        const uint32_t sbo = 32u * (uint32_t)KK, lo_off = 16u * (uint32_t)KK;
        const uint32_t idesc_s = tc::idesc_bf16_f32(128, 64) | (1u << 16), idesc_b = tc::idesc_bf16_f32(128, 16) | (1u << 16);
        const uint32_t tb = tmem_base + (qq & 1u) * kLookT;
        uint32_t acc = 0;
        for (int ks = 0; ks < KK / 16; ++ks) {
            const uint32_t off = (uint32_t)ks * 256u;
            const uint64_t dsh = tc::smem_desc(sb + off, 128, sbo), dsl = tc::smem_desc(sb + lo_off + off, 128, sbo);
            const uint64_t dbh = tc::smem_desc(bb + off, 128, sbo), dbl = tc::smem_desc(bb + lo_off + off, 128, sbo);
            const uint32_t tah = tmem_base + kLookA + 8u * ks, tal = tah + (uint32_t)KK / 2;
            tc::umma_bf16_ts(tb, tah, dsh, idesc_s, acc);
            tc::umma_bf16_ts(tb, tah, dsl, idesc_s, 1);
            tc::umma_bf16_ts(tb, tal, dsh, idesc_s, 1);
            tc::umma_bf16_ts(tb + kLookBase, tah, dbh, idesc_b, acc);
            tc::umma_bf16_ts(tb + kLookBase, tah, dbl, idesc_b, 1);
            tc::umma_bf16_ts(tb + kLookBase, tal, dbh, idesc_b, 1);
            acc = 1;
        }
        tc::umma_commit(&bar_t[qq & 1u]);
    };
    uint32_t q = 0, cn = 0;                              // running pass / first chunk of the current tile

    if (loader) {
        // the tile's dY rows -> shared memory (16-byte copies, zero-filled outside the matrix), one tile ahead of the workers
        auto stage_dy = [&](int tile_) {
            if (dy_vec) {
                const int c4n = KK >> 2, lane = tid & 31;
                const int dq = 32 / c4n, dr = 32 - dq * c4n;
                int r = lane / c4n, c = lane - r * c4n;
                while (r < 128) {
                    const long long row = (long long)tile_ * 128 + r;
                    const bool on = row < n_rows && 4 * c < g.out_f;
                    cp_async16(d_st + r * dld + 4 * c, on ? dy + row * ld_dy + 4 * c : dy, on);
                    r += dq;
                    c += dr;
                    if (c >= c4n) {
                        c -= c4n;
                        ++r;
                    }
                }
            }
            cp_async_arrive_on(bar_dy);
        };
        if ((int)blockIdx.x < n_tiles) stage_dy((int)blockIdx.x);
        for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
            named_sync<3, kLookThreads>();                // the workers have read the staged tile
            if (tile + (int)gridDim.x < n_tiles) stage_dy(tile + (int)gridDim.x);
        }
    } else if (control) {
        // ---- control warp: "A is in tensor memory" (barrier 3) -> pass 0; "T buffer b is free" (barriers 1 + b, one arrival
        // generation per pass) -> the pass two ahead of the one that freed it; every generation is consumed, the last two of a
        // tile at its end
        const bool lead = (tid & 31) == 0;
        auto free_sync = [&](uint32_t b_) {
            if (b_) named_sync<2, kLookSync>();
            else named_sync<1, kLookSync>();
        };
        for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
            named_sync<3, kLookThreads>();                // A is in tensor memory; nobody reads the staged dY tile any more
            if (lead) issue(0, q, cn);
            __syncwarp();
            for (int pass = 0; pass < n_pass; ++pass, ++q) {
                const uint32_t cc = cn + (uint32_t)(pass >> 1);
                if (pass + 1 < n_pass) {
                    if (pass >= 1) free_sync((q + 1u) & 1u);
                    if (lead) issue(pass + 1, q + 1, cn + (uint32_t)((pass + 1) >> 1));
                    __syncwarp();
                }
                // the last MMA of a chunk has completed: its stage takes the chunk after next
                if (((pass & 1) == 1 || pass == n_pass - 1) && cc + 2 < total_chunks) {
                    if (lead) {
                        tc::mbar_wait(&bar_t[q & 1u], (q >> 1) & 1u);
                        load_chunk(cc + 2);
                    }
                    __syncwarp();
                }
            }
            if (n_pass >= 2) free_sync(q & 1u);          // pass n_pass - 2 (q is already one past the last pass)
            free_sync((q + 1u) & 1u);                    // pass n_pass - 1
            cn += (uint32_t)chunks;
        }
    } else {
    uint32_t tcount = 0;
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++tcount) {
        const long long row = (long long)tile * 128 + r128;
        const bool row_ok = row < n_rows;
        float mean = 0.f, rstd = 1.f;
        if (K == 0 && g.stats && row_ok) {
            mean = __ldg(g.stats + 2 * row);
            rstd = __ldg(g.stats + 2 * row + 1);
        }
        // ---- A = this row of dY, split into bf16 hi / lo, into tensor memory (lane = row); the two warpgroups take alternate
        // 8-column groups
        const float* dyr = dy + (row_ok ? row : 0) * ld_dy;
        tc::mbar_wait(bar_dy, tcount & 1u);
        {
            float4 qv[4][2];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int kc = wg + 2 * j;
                const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
                if (dy_vec) {                           // staged by the control warp
                    const float* ds = d_st + r128 * dld + kc * 8;
                    qv[j][0] = kc < kcs ? *reinterpret_cast<const float4*>(ds) : z;
                    qv[j][1] = kc < kcs ? *reinterpret_cast<const float4*>(ds + 4) : z;
                } else {
                    float v[8];
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const int o = kc * 8 + i;
                        v[i] = (row_ok && kc < kcs && o < g.out_f) ? __ldg(dyr + o) : 0.f;
                    }
                    qv[j][0] = make_float4(v[0], v[1], v[2], v[3]);
                    qv[j][1] = make_float4(v[4], v[5], v[6], v[7]);
                }
            }
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int kc = wg + 2 * j;
                if (kc < kcs) {
                    const float v[8] = {qv[j][0].x, qv[j][0].y, qv[j][0].z, qv[j][0].w, qv[j][1].x, qv[j][1].y, qv[j][1].z, qv[j][1].w};
                    uint4 hi, lo;
                    tc::split8(v, hi, lo);
                    tc::tmem_st4(tmem_base + lane_base + kLookA + 4u * kc, hi.x, hi.y, hi.z, hi.w);
                    tc::tmem_st4(tmem_base + lane_base + kLookA + (uint32_t)KK / 2 + 4u * kc, lo.x, lo.y, lo.z, lo.w);
                }
            }
        }
        tc::tmem_st_wait();
        tc::tc_fence_before_sync();
        named_arrive<3, kLookThreads>();
        const float* xr = x + (row_ok ? row : 0) * ldx;
        auto load_x4 = [&](int fs, float (&o)[4]) {     // this thread's four x values of a pass: one 16-byte load when the rows allow it
            if (x_vec) {
                const float4 v = (row_ok && fs < g.in_f) ? __ldg(reinterpret_cast<const float4*>(xr + fs)) : make_float4(0.f, 0.f, 0.f, 0.f);
                o[0] = v.x;
                o[1] = v.y;
                o[2] = v.z;
                o[3] = v.w;
            } else {
#pragma unroll
                for (int ii = 0; ii < 4; ++ii) o[ii] = (row_ok && fs + ii < g.in_f) ? __ldg(xr + fs + ii) : 0.f;
            }
        };
        float xq[4];
        load_x4(4 * wg, xq);
        for (int pass = 0; pass < n_pass; ++pass, ++q) {
            const int f0 = pass * kFPL;
            float xn[4] = {0.f, 0.f, 0.f, 0.f};
            if (pass + 1 < n_pass) load_x4(f0 + kFPL + 4 * wg, xn);
            tc::mbar_wait(&bar_t[q & 1u], (q >> 1) & 1u);
            tc::tc_fence_after_sync();
            // ---- epilogue: warpgroup wg contracts features 4 wg .. 4 wg + 3 of the pass for its 128 rows
            const uint32_t tb = tmem_base + lane_base + (q & 1u) * kLookT;
            float res[4], resb[4], tbase[8];
            tc::tmem_ld8(tb + kLookBase + 8u * (uint32_t)(pass & 1), tbase);
#pragma unroll
            for (int ii = 0; ii < 4; ++ii) {
                const int i = 4 * wg + ii, f = f0 + i;
                res[ii] = 0.f;
                resb[ii] = 0.f;
                if (f < g.in_f) {                       // uniform over the warpgroup
                    float t[8];
                    tc::tmem_ld8(tb + (uint32_t)(8 * i), t);
                    const float xv = xq[ii];
                    float a = 0.f;
                    if (K == 0) {
                        const float z = rbf_z_b(g, xv, mean, rstd, f);
#pragma unroll
                        for (int c = 0; c < 8; ++c) {
                            const float tt = (z - (g.c0 + (float)c * g.step)) * g.inv_den;
                            const float dq = -2.0f * tt * g.inv_den * ex2_b(-kLog2eB * tt * tt);
                            a = fmaf((c < g.S) ? dq : 0.f, t[c], a);
                        }
                    } else {
                        int idx;
                        float fr, d[4];
                        locate_b(g, xv, idx, fr);
                        local_derivs_b<(K == 0 ? 1 : K)>(fr, g.inv_h, d);
                        const int j = idx - 1;              // interval; its bases sit on slots j - K .. j
                        const bool live = j >= 0 && j < g.G + 2 * K;
                        float tp[20];
#pragma unroll
                        for (int c = 0; c < 20; ++c) tp[c] = 0.f;
#pragma unroll
                        for (int c = 0; c < 8; ++c)
                            if (c < kS1MaxB - 1) tp[K + c] = (c < g.S) ? t[c] : 0.f;
                        const int jj = live ? j : 0;
                        float s0[12], s1[8], s2[6], s3[4];
#pragma unroll
                        for (int c = 0; c < 12; ++c) s0[c] = (jj & 8) ? tp[c + 8] : tp[c];
#pragma unroll
                        for (int c = 0; c < 8; ++c) s1[c] = (jj & 4) ? s0[c + 4] : s0[c];
#pragma unroll
                        for (int c = 0; c < 6; ++c) s2[c] = (jj & 2) ? s1[c + 2] : s1[c];
#pragma unroll
                        for (int c = 0; c < 4; ++c) s3[c] = (jj & 1) ? s2[c + 1] : s2[c];
#pragma unroll
                        for (int r = 0; r <= K; ++r) a = fmaf(d[r], s3[r], a);
                        if (!live) a = 0.f;
                    }
                    const float sg = __fdividef(1.0f, 1.0f + ex2_b(-kLog2eB * xv));
                    const float dbase = sg * fmaf(xv, 1.0f - sg, 1.0f);
                    const float tbv = wg ? tbase[4 + ii] : tbase[ii];
                    if (K == 0 && dxb) {                    // with a LayerNorm the two branches stay separate (kagnn_layernorm_bwd joins them)
                        res[ii] = a;
                        resb[ii] = dbase * tbv;
                    } else {
                        res[ii] = fmaf(dbase, tbv, a);
                    }
                }
            }
            if (row_ok) {
                float* dxr = dx + row * ld_dx + f0 + 4 * wg;
                if (dx_vec && f0 + 4 * wg + 4 <= g.in_f) {
                    *reinterpret_cast<float4*>(dxr) = make_float4(res[0], res[1], res[2], res[3]);
                } else {
#pragma unroll
                    for (int ii = 0; ii < 4; ++ii)
                        if (f0 + 4 * wg + ii < g.in_f) dxr[ii] = res[ii];
                }
                if (K == 0 && dxb) {
                    float* br = dxb + row * ld_dxb + f0 + 4 * wg;
#pragma unroll
                    for (int ii = 0; ii < 4; ++ii)
                        if (f0 + 4 * wg + ii < g.in_f) br[ii] = resb[ii];
                }
            }
#pragma unroll
            for (int ii = 0; ii < 4; ++ii) xq[ii] = xn[ii];
            // this T buffer is free again (the control warp issues the pass after next into it)
            tc::tc_fence_before_sync();
            if (q & 1u) named_arrive<2, kLookSync>();
            else named_arrive<1, kLookSync>();
        }
        // (the next tile's A may land at once: this thread has seen every pass of the tile complete)
    }
    }
    tc::tc_fence_before_sync();
    __syncthreads();
    if (warp == 0) tc::tmem_dealloc(tmem_base, tmem_cols);
}

// =====================================================================================================================
// dW
// =====================================================================================================================
constexpr int kFB = 14;                 // features per block: 14 units of 8 slots + 2 units of base values = M 128

template <int K>
__global__ void __launch_bounds__(kThreads) kan_bwd_weights_tc_kernel(GeomB g, const float* __restrict__ x, long long ldx,
                                                                      const float* __restrict__ dy, long long ld_dy, long long n_rows,
                                                                      long long rows_per_slab, int N16, uint32_t tmem_cols, int swap_lbo,
                                                                      float* __restrict__ dP) {
    extern __shared__ __align__(128) uint8_t smem[];
    // A (E^T): 16 units x 128 rows x 16 B, hi then lo;  B (dY): N16/8 units x 128 rows x 16 B, hi then lo
    constexpr uint32_t kABytes = 16u * 2048u;
    const uint32_t b_bytes = (uint32_t)(N16 / 8) * 2048u;
    uint8_t* a_hi = smem;
    uint8_t* a_lo = a_hi + kABytes;
    uint8_t* b_hi = a_lo + kABytes;
    uint8_t* b_lo = b_hi + b_bytes;
    uint4* lut = reinterpret_cast<uint4*>(b_lo + b_bytes);
    uint64_t* bar = reinterpret_cast<uint64_t*>(lut + kLutRows);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar + 1);
    const int tid = threadIdx.x, warp = tid >> 5, wg = warp >> 2, r128 = tid & 127;
    const int f0 = blockIdx.x * kFB;
    const long long r_beg = (long long)blockIdx.y * rows_per_slab, r_end = min(n_rows, r_beg + rows_per_slab);

    if (warp == 0) tc::tmem_alloc(tmem_slot, tmem_cols);
    if (tid == 32) {
        tc::mbar_init(bar, 1);
        tc::mbar_fence_init();
    }
    if (K > 0 && tid >= 64 && tid < 64 + kLutRows) lut[tid - 64] = lut_row_b(tid - 64, K, g.G + 2 * K);
    tc::tc_fence_before_sync();
    __syncthreads();
    tc::tc_fence_after_sync();
    const uint32_t tmem_base = *tmem_slot;
    const uint32_t idesc = idesc_bf16_f32_mn(128, N16);
    const uint32_t lbo = swap_lbo ? 2048u : 128u, sbo = swap_lbo ? 128u : 2048u;
    uint32_t phase = 0, acc = 0;
    // warpgroup 0: features 0..7 (+ their base values, unit 14) and the first quarter of dY's units;
    // warpgroup 1: features 8..13 (+ base values, unit 15) and the rest of dY
    const int fi0 = wg ? 8 : 0, fi1 = wg ? kFB : 8;
    // share of dY's units per warpgroup: balanced against the features (8 vs 6, ~60 instructions each; a unit ~35)
    const int nu = N16 / 8, u_split = (35 * nu - 120) > 0 ? (35 * nu - 120) / 70 : 0;
    const int u0 = wg ? u_split : 0, u1 = wg ? nu : u_split;

    // operands of a tile are fetched into registers one tile ahead (while the tensor pipe works on the previous one): the x values
    // of this warpgroup's features and, when dY is at most 64 wide and 16-byte aligned, its dY units
    constexpr int kPU = 6;
    const bool dy_vec = g.out_f % 4 == 0 && ld_dy % 4 == 0 && ((reinterpret_cast<uintptr_t>(dy) & 15u) == 0);
    const bool dy_pf = nu <= 8 && dy_vec;
    float xq[8];
    float4 dq[kPU][2];
    auto load_tile = [&](long long rt_) {
        const long long row = rt_ + r128;
        const bool row_ok = row < r_end;
        const float* xr = x + (row_ok ? row : 0) * ldx;
#pragma unroll
        for (int ii = 0; ii < 8; ++ii) {
            const int i = fi0 + ii;
            xq[ii] = (row_ok && i < fi1 && (f0 + i) < g.in_f) ? __ldg(xr + f0 + i) : 0.f;
        }
        if (dy_pf) {
            const float* dyr = dy + (row_ok ? row : 0) * ld_dy;
#pragma unroll
            for (int q = 0; q < kPU; ++q) {
                const int u = u0 + q;
                const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
                dq[q][0] = (row_ok && u < u1 && 8 * u + 4 <= g.out_f) ? __ldg(reinterpret_cast<const float4*>(dyr + 8 * u)) : z;
                dq[q][1] = (row_ok && u < u1 && 8 * u + 8 <= g.out_f) ? __ldg(reinterpret_cast<const float4*>(dyr + 8 * u + 4)) : z;
            }
        }
    };
    if (r_beg < r_end) load_tile(r_beg);
    for (long long rt = r_beg; rt < r_end; rt += 128) {
        const long long row = rt + r128;
        const bool row_ok = row < r_end;
        float mean = 0.f, rstd = 1.f;
        if (K == 0 && g.stats && row_ok) {
            mean = __ldg(g.stats + 2 * row);
            rstd = __ldg(g.stats + 2 * row + 1);
        }
        // ---- E^T: this row's slot values, one 16-byte vector per feature (unit u = feature), base values in units 14 / 15
        float bb[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) bb[i] = 0.f;
#pragma unroll
        for (int ii = 0; ii < 8; ++ii) {
            const int i = fi0 + ii;
            if (i < fi1) {                              // uniform over the warpgroup
                const bool on = row_ok && (f0 + i) < g.in_f;
                const float xv = xq[ii];
                uint4 hi, lo;
                if (K == 0) {
                    // RBF: the G Gaussians of z = LayerNorm(x) fill the slots densely
                    const float z = rbf_z_b(g, xv, mean, rstd, f0 + i);
                    float e[8];
#pragma unroll
                    for (int q = 0; q < 8; ++q) {
                        const float tt = (z - (g.c0 + (float)q * g.step)) * g.inv_den;
                        e[q] = (on && q < g.S) ? ex2_b(-kLog2eB * tt * tt) : 0.f;
                    }
                    tc::split8(e, hi, lo);
                } else {
                    int idx;
                    float fr, b[4];
                    locate_b(g, xv, idx, fr);
                    local_values_b<(K == 0 ? 1 : K)>(fr, b);
                    // every value is >= 0, so hi = truncation to bf16 and lo = bf16(b - hi) are >= 0 too: their sign bits are 0, which
                    // lets the byte permute synthesise the empty slots by sign replication (selector nibble 9)
                    const uint32_t h01 = pack_trunc_b(b[0], b[1]), h23 = pack_trunc_b(b[2], b[3]);
                    const uint32_t l01 = pack_rn_b(trunc_res_b(b[0]), trunc_res_b(b[1])), l23 = pack_rn_b(trunc_res_b(b[2]), trunc_res_b(b[3]));
                    uint4 sel = lut[idx];
                    if (!on) sel = make_uint4(0x9999u, 0x9999u, 0x9999u, 0x9999u);
                    hi = make_uint4(prmt_b(h01, h23, sel.x), prmt_b(h01, h23, sel.y), prmt_b(h01, h23, sel.z), prmt_b(h01, h23, sel.w));
                    lo = make_uint4(prmt_b(l01, l23, sel.x), prmt_b(l01, l23, sel.y), prmt_b(l01, l23, sel.z), prmt_b(l01, l23, sel.w));
                }
                *reinterpret_cast<uint4*>(a_hi + (size_t)i * 2048 + r128 * 16) = hi;
                *reinterpret_cast<uint4*>(a_lo + (size_t)i * 2048 + r128 * 16) = lo;
                bb[ii] = on ? __fdividef(xv, 1.0f + ex2_b(-kLog2eB * xv)) : 0.f;
            }
        }
        {
            uint4 hi, lo;
            tc::split8(bb, hi, lo);
            *reinterpret_cast<uint4*>(a_hi + (size_t)(kFB + wg) * 2048 + r128 * 16) = hi;
            *reinterpret_cast<uint4*>(a_lo + (size_t)(kFB + wg) * 2048 + r128 * 16) = lo;
        }
        // ---- dY row: N16 / 8 units
        if (dy_pf) {
#pragma unroll
            for (int q = 0; q < kPU; ++q) {
                const int u = u0 + q;
                if (u < u1) {
                    const float v[8] = {dq[q][0].x, dq[q][0].y, dq[q][0].z, dq[q][0].w, dq[q][1].x, dq[q][1].y, dq[q][1].z, dq[q][1].w};
                    uint4 hi, lo;
                    tc::split8(v, hi, lo);
                    *reinterpret_cast<uint4*>(b_hi + (size_t)u * 2048 + r128 * 16) = hi;
                    *reinterpret_cast<uint4*>(b_lo + (size_t)u * 2048 + r128 * 16) = lo;
                }
            }
        } else {
            // wide dY (more than 8 units): no register prefetch, but four units' loads in flight per round
            const float* dyr = dy + (row_ok ? row : 0) * ld_dy;
            for (int ub = u0; ub < u1; ub += 4) {
                float4 q[4][2];
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const int u = ub + j;
                    const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (dy_vec) {
                        q[j][0] = (row_ok && u < u1 && 8 * u + 4 <= g.out_f) ? __ldg(reinterpret_cast<const float4*>(dyr + 8 * u)) : z;
                        q[j][1] = (row_ok && u < u1 && 8 * u + 8 <= g.out_f) ? __ldg(reinterpret_cast<const float4*>(dyr + 8 * u + 4)) : z;
                    } else {
                        float v[8];
#pragma unroll
                        for (int c = 0; c < 8; ++c) {
                            const int o = 8 * u + c;
                            v[c] = (row_ok && u < u1 && o < g.out_f) ? __ldg(dyr + o) : 0.f;
                        }
                        q[j][0] = make_float4(v[0], v[1], v[2], v[3]);
                        q[j][1] = make_float4(v[4], v[5], v[6], v[7]);
                    }
                }
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const int u = ub + j;
                    if (u < u1) {
                        const float v[8] = {q[j][0].x, q[j][0].y, q[j][0].z, q[j][0].w, q[j][1].x, q[j][1].y, q[j][1].z, q[j][1].w};
                        uint4 hi, lo;
                        tc::split8(v, hi, lo);
                        *reinterpret_cast<uint4*>(b_hi + (size_t)u * 2048 + r128 * 16) = hi;
                        *reinterpret_cast<uint4*>(b_lo + (size_t)u * 2048 + r128 * 16) = lo;
                    }
                }
            }
        }
        tc::fence_proxy_async_smem();
        tc::tc_fence_before_sync();
        __syncthreads();
        if (tid == 0) {
            tc::tc_fence_after_sync();
            for (int ks = 0; ks < 8; ++ks) {            // K = 16 rows per step: two groups of 8 k-rows, 256 bytes
                const uint32_t off = (uint32_t)ks * 256u;
                const uint64_t dah = tc::smem_desc(tc::smem_u32(a_hi) + off, lbo, sbo), dal = tc::smem_desc(tc::smem_u32(a_lo) + off, lbo, sbo);
                const uint64_t dbh = tc::smem_desc(tc::smem_u32(b_hi) + off, lbo, sbo), dbl = tc::smem_desc(tc::smem_u32(b_lo) + off, lbo, sbo);
                tc::umma_bf16(tmem_base, dah, dbh, idesc, acc);
                tc::umma_bf16(tmem_base, dah, dbl, idesc, 1);
                tc::umma_bf16(tmem_base, dal, dbh, idesc, 1);
                acc = 1;
            }
            tc::umma_commit(bar);
        }
        if (rt + 128 < r_end) load_tile(rt + 128);      // in flight while the tensor pipe works
        tc::mbar_wait(bar, phase);                      // the operands in shared memory may be overwritten
        phase ^= 1u;
    }
    // ---- the slab's sums: lane m of tensor memory = row m of the block's 128 gradient rows; the warpgroups split the columns
    tc::tc_fence_after_sync();
    if (r_beg < r_end) {
        const int m = r128, u = m >> 3, c = m & 7;
        long long dst = -1;
        if (u < kFB) {
            if (f0 + u < g.in_f && c < g.S) dst = ((long long)(f0 + u) * (g.S + 1) + c) * g.out_pad;
        } else {
            const int i = (u - kFB) * 8 + c;
            if (i < kFB && f0 + i < g.in_f) dst = ((long long)(f0 + i) * (g.S + 1) + g.S) * g.out_pad;
        }
        const uint32_t lane_base = (uint32_t)((warp & 3) * 32) << 16;
        for (int o0 = 8 * wg; o0 < N16; o0 += 16) {
            float v[8];
            tc::tmem_ld8(tmem_base + lane_base + (uint32_t)o0, v);   // warp-collective: every lane takes part
            if (dst >= 0) {
#pragma unroll
                for (int q = 0; q < 8; ++q)
                    if (o0 + q < g.out_f && v[q] != 0.f) atomicAdd(dP + dst + o0 + q, v[q]);
            }
        }
    }
    tc::tc_fence_before_sync();
    __syncthreads();
    if (warp == 0) tc::tmem_dealloc(tmem_base, tmem_cols);
}

// ---------------------------------------------------------------------------------------------------------------------
// dW for layers up to 64 outputs wide: the same GEMM on 64-row batches.  A batch needs 48 KB of shared memory instead of 96,
// so three CTAs share an SM and cover each other's phases (operand production / MMA / wait) better than two.  256 threads =
// 64 rows x 4 parts; part p expands features p, p + 4, p + 8, (p + 12) of the block's 14 and two or three of dY's 8-column
// units; the base values go into units 14 / 15 element by element.
// ---------------------------------------------------------------------------------------------------------------------
template <int K>
__global__ void __launch_bounds__(kThreads, 3) kan_bwd_weights_tc64_kernel(GeomB g, const float* __restrict__ x, long long ldx,
                                                                           const float* __restrict__ dy, long long ld_dy, long long n_rows,
                                                                           long long rows_per_slab, int N16, uint32_t tmem_cols,
                                                                           float* __restrict__ dP) {
    extern __shared__ __align__(128) uint8_t smem[];
    constexpr uint32_t kUnit = 64u * 16u;                    // one unit of 8 M (or N) elements x 64 rows
    constexpr uint32_t kABytes = 16u * kUnit;
    const int nu = N16 / 8;                                  // <= 8
    const uint32_t b_bytes = (uint32_t)nu * kUnit;
    uint8_t* a_hi = smem;
    uint8_t* a_lo = a_hi + kABytes;
    uint8_t* b_hi = a_lo + kABytes;
    uint8_t* b_lo = b_hi + b_bytes;
    uint4* lut = reinterpret_cast<uint4*>(b_lo + b_bytes);
    uint64_t* bar = reinterpret_cast<uint64_t*>(lut + kLutRows);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar + 1);
    const int tid = threadIdx.x, warp = tid >> 5, r64 = tid & 63, part = tid >> 6;
    const int f0 = blockIdx.x * kFB;
    const long long r_beg = (long long)blockIdx.y * rows_per_slab, r_end = min(n_rows, r_beg + rows_per_slab);

    if (warp == 0) tc::tmem_alloc(tmem_slot, tmem_cols);
    if (tid == 32) {
        tc::mbar_init(bar, 1);
        tc::mbar_fence_init();
    }
    if (K > 0 && tid >= 64 && tid < 64 + kLutRows) lut[tid - 64] = lut_row_b(tid - 64, K, g.G + 2 * K);
    tc::tc_fence_before_sync();
    __syncthreads();
    tc::tc_fence_after_sync();
    const uint32_t tmem_base = *tmem_slot;
    const uint32_t idesc = idesc_bf16_f32_mn(128, N16);
    uint32_t phase = 0, acc = 0;
    // dY units of this part: parts 2 and 3 expand three features and take three units each (of eight), parts 0 and 1 four features
    // and one unit each: unit u belongs to part kUnitPart[u]
    const int my_units = (part >= 2) ? 3 : 1;

    float xq[4];
    float4 dq[3][2];
    auto unit_of = [&](int q) { return part >= 2 ? (part - 2) + 2 * q : 6 + part; };      // parts 2,3: units 0..5 interleaved; parts 0,1: units 6,7
    auto load_tile = [&](long long rt_) {
        const long long row = rt_ + r64;
        const bool row_ok = row < r_end;
        const float* xr = x + (row_ok ? row : 0) * ldx;
#pragma unroll
        for (int ii = 0; ii < 4; ++ii) {
            const int i = part + 4 * ii;
            xq[ii] = (row_ok && i < kFB && (f0 + i) < g.in_f) ? __ldg(xr + f0 + i) : 0.f;
        }
        const float* dyr = dy + (row_ok ? row : 0) * ld_dy;
#pragma unroll
        for (int q = 0; q < 3; ++q) {
            const int u = unit_of(q);
            const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
            const bool on = row_ok && q < my_units && u < nu;
            dq[q][0] = (on && 8 * u + 4 <= g.out_f) ? __ldg(reinterpret_cast<const float4*>(dyr + 8 * u)) : z;
            dq[q][1] = (on && 8 * u + 8 <= g.out_f) ? __ldg(reinterpret_cast<const float4*>(dyr + 8 * u + 4)) : z;
        }
    };
    if (r_beg < r_end) load_tile(r_beg);
    for (long long rt = r_beg; rt < r_end; rt += 64) {
        const long long row = rt + r64;
        const bool row_ok = row < r_end;
        float mean = 0.f, rstd = 1.f;
        if (K == 0 && g.stats && row_ok) {
            mean = __ldg(g.stats + 2 * row);
            rstd = __ldg(g.stats + 2 * row + 1);
        }
#pragma unroll
        for (int ii = 0; ii < 4; ++ii) {
            const int i = part + 4 * ii;
            if (i < kFB) {                                  // uniform over the part (two warps)
                const bool on = row_ok && (f0 + i) < g.in_f;
                const float xv = xq[ii];
                uint4 hi, lo;
                if (K == 0) {
                    const float z = rbf_z_b(g, xv, mean, rstd, f0 + i);
                    float e[8];
#pragma unroll
                    for (int q = 0; q < 8; ++q) {
                        const float tt = (z - (g.c0 + (float)q * g.step)) * g.inv_den;
                        e[q] = (on && q < g.S) ? ex2_b(-kLog2eB * tt * tt) : 0.f;
                    }
                    tc::split8(e, hi, lo);
                } else {
                    int idx;
                    float fr, b[4];
                    locate_b(g, xv, idx, fr);
                    local_values_b<(K == 0 ? 1 : K)>(fr, b);
                    const uint32_t h01 = pack_trunc_b(b[0], b[1]), h23 = pack_trunc_b(b[2], b[3]);
                    const uint32_t l01 = pack_rn_b(trunc_res_b(b[0]), trunc_res_b(b[1])), l23 = pack_rn_b(trunc_res_b(b[2]), trunc_res_b(b[3]));
                    uint4 sel = lut[idx];
                    if (!on) sel = make_uint4(0x9999u, 0x9999u, 0x9999u, 0x9999u);
                    hi = make_uint4(prmt_b(h01, h23, sel.x), prmt_b(h01, h23, sel.y), prmt_b(h01, h23, sel.z), prmt_b(h01, h23, sel.w));
                    lo = make_uint4(prmt_b(l01, l23, sel.x), prmt_b(l01, l23, sel.y), prmt_b(l01, l23, sel.z), prmt_b(l01, l23, sel.w));
                }
                *reinterpret_cast<uint4*>(a_hi + (size_t)i * kUnit + r64 * 16) = hi;
                *reinterpret_cast<uint4*>(a_lo + (size_t)i * kUnit + r64 * 16) = lo;
                // base value: element (i & 7) of unit 14 + (i >> 3), as bf16 hi (truncated) / lo (rounded residual)
                const float sv = on ? __fdividef(xv, 1.0f + ex2_b(-kLog2eB * xv)) : 0.f;
                const uint32_t off = (uint32_t)(kFB + (i >> 3)) * kUnit + (uint32_t)r64 * 16u + (uint32_t)(i & 7) * 2u;
                *reinterpret_cast<uint16_t*>(a_hi + off) = (uint16_t)(__float_as_uint(sv) >> 16);
                *reinterpret_cast<uint16_t*>(a_lo + off) = (uint16_t)(pack_rn_b(trunc_res_b(sv), 0.f) & 0xffffu);
            }
        }
        if (part < 2) {                                      // the two pad elements (features 14, 15) of unit 15: zero, once per tile by one part each
            const uint32_t off = (uint32_t)(kFB + 1) * kUnit + (uint32_t)r64 * 16u + (uint32_t)(6 + part) * 2u;
            *reinterpret_cast<uint16_t*>(a_hi + off) = 0;
            *reinterpret_cast<uint16_t*>(a_lo + off) = 0;
        }
#pragma unroll
        for (int q = 0; q < 3; ++q) {
            const int u = unit_of(q);
            if (q < my_units && u < nu) {
                const float v[8] = {dq[q][0].x, dq[q][0].y, dq[q][0].z, dq[q][0].w, dq[q][1].x, dq[q][1].y, dq[q][1].z, dq[q][1].w};
                uint4 hi, lo;
                tc::split8(v, hi, lo);
                *reinterpret_cast<uint4*>(b_hi + (size_t)u * kUnit + r64 * 16) = hi;
                *reinterpret_cast<uint4*>(b_lo + (size_t)u * kUnit + r64 * 16) = lo;
            }
        }
        tc::fence_proxy_async_smem();
        tc::tc_fence_before_sync();
        __syncthreads();
        if (tid == 0) {
            tc::tc_fence_after_sync();
            for (int ks = 0; ks < 4; ++ks) {            // K = 16 rows per step: two groups of 8 k-rows, 256 bytes
                const uint32_t off = (uint32_t)ks * 256u;
                const uint64_t dah = tc::smem_desc(tc::smem_u32(a_hi) + off, 128, kUnit), dal = tc::smem_desc(tc::smem_u32(a_lo) + off, 128, kUnit);
                const uint64_t dbh = tc::smem_desc(tc::smem_u32(b_hi) + off, 128, kUnit), dbl = tc::smem_desc(tc::smem_u32(b_lo) + off, 128, kUnit);
                tc::umma_bf16(tmem_base, dah, dbh, idesc, acc);
                tc::umma_bf16(tmem_base, dah, dbl, idesc, 1);
                tc::umma_bf16(tmem_base, dal, dbh, idesc, 1);
                acc = 1;
            }
            tc::umma_commit(bar);
        }
        if (rt + 64 < r_end) load_tile(rt + 64);        // in flight while the tensor pipe works
        tc::mbar_wait(bar, phase);                      // the operands in shared memory may be overwritten
        phase ^= 1u;
    }
    // ---- the slab's sums: lane m of tensor memory = row m of the block's 128 gradient rows; the two halves of the CTA split the columns
    tc::tc_fence_after_sync();
    if (r_beg < r_end) {
        const int m = tid & 127, u = m >> 3, c = m & 7, half = tid >> 7;
        long long dst = -1;
        if (u < kFB) {
            if (f0 + u < g.in_f && c < g.S) dst = ((long long)(f0 + u) * (g.S + 1) + c) * g.out_pad;
        } else {
            const int i = (u - kFB) * 8 + c;
            if (i < kFB && f0 + i < g.in_f) dst = ((long long)(f0 + i) * (g.S + 1) + g.S) * g.out_pad;
        }
        const uint32_t lane_base = (uint32_t)((warp & 3) * 32) << 16;
        for (int o0 = 8 * half; o0 < N16; o0 += 16) {
            float v[8];
            tc::tmem_ld8(tmem_base + lane_base + (uint32_t)o0, v);   // warp-collective: every lane takes part
            if (dst >= 0) {
#pragma unroll
                for (int q = 0; q < 8; ++q)
                    if (o0 + q < g.out_f && v[q] != 0.f) atomicAdd(dP + dst + o0 + q, v[q]);
            }
        }
    }
    tc::tc_fence_before_sync();
    __syncthreads();
    if (warp == 0) tc::tmem_dealloc(tmem_base, tmem_cols);
}

// ---------------------------------------------------------------------------------------------------------------------
// dW for layers up to 64 outputs wide, several feature blocks per CTA.  One CTA per SM: 16 producer warps, 1 control warp,
// 4 loader warps.
//   * the hi / lo split of a dY batch (64 rows) is made ONCE and serves up to eight feature blocks, each with its own
//     accumulator in tensor memory (8 x 64 columns = all 512);
//   * producers: 512 threads = 64 rows x 8 parts; part p expands features p and p + 8 of a block's 14 (and one of dY's eight
//     units at the head of a batch) into a ring of three E^T operands; they only ARRIVE at the named barrier of the stage and
//     go on -- the mbarrier of the stage three blocks later is the only thing that can stop them;
//   * the control warp waits on that named barrier, issues the block's 12 MMAs and commits to the stage's mbarrier;
//   * the loader warps keep the next batch's x and dY rows coming: 16-byte cp.async copies (zero-filled outside the matrix) into
//     double-buffered, padded staging tiles, completion on an mbarrier: no producer ever waits on a global load;
//   * the dY operand is double-buffered too, so a batch does not wait for the previous batch's MMAs;
//   * the slab's sums leave as 16-byte vector reductions, each CTA starting at a different column.
// grid = (passes of up to 8 feature blocks, row slabs).
// ---------------------------------------------------------------------------------------------------------------------
constexpr int kMBlocks = 8;
constexpr int kMProducers = 512;
constexpr int kMSync = kMProducers + 32;                     // producers + the control warp: the named barriers of the operand ring
constexpr int kMLoaders = 128;
constexpr int kMThreads = kMSync + kMLoaders;
constexpr int kMStages = 3;
constexpr int kMXld = kMBlocks * kFB + 4;                    // staged x row: up to 112 values + 2 of alignment slack, 16-byte rows

__device__ __forceinline__ void red_add_v4(float* dst, float a, float b, float c, float d) {
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

struct DwSmem {                                              // carve-up of the dynamic shared memory (offsets in bytes)
    uint32_t a, b, xs, ds, lut, bars, total;
};
__host__ __device__ inline DwSmem dw_smem_layout(int N16, int rows) {
    DwSmem L;
    const uint32_t unit = (uint32_t)rows * 16u;
    L.a = 0;
    L.b = L.a + (uint32_t)kMStages * 2u * 16u * unit;        // [stage][hi | lo]
    L.xs = L.b + 2u * 2u * (uint32_t)(N16 / 8) * unit;       // [buffer][hi | lo]
    L.ds = L.xs + 2u * (uint32_t)rows * (uint32_t)kMXld * 4u;
    L.lut = L.ds + 2u * (uint32_t)rows * (uint32_t)(N16 + 4) * 4u;         // staged dY row: N16 values + 4 (16-byte rows, odd in 16-byte units)
    L.bars = L.lut + (uint32_t)kLutRows * 16u;
    L.total = L.bars + 160u;
    return L;
}

template <int K, int ROWS>
__global__ void __launch_bounds__(kMThreads, 1) kan_bwd_weights_tcm_kernel(GeomB g, const float* __restrict__ x, long long ldx,
                                                                             const float* __restrict__ dy, long long ld_dy, long long n_rows,
                                                                             long long rows_per_slab, int N16, int fblocks,
                                                                             int blocks_per_pass, uint32_t tmem_cols, int x16, int dy16,
                                                                             float* __restrict__ dP) {
    extern __shared__ __align__(128) uint8_t smem[];
    constexpr uint32_t kUnit = (uint32_t)ROWS * 16u;         // one unit of 8 M (or N) elements x ROWS rows
    constexpr uint32_t kAStage = 2u * 16u * kUnit;           // hi + lo of one E^T operand
    constexpr int kParts = kMProducers / ROWS;               // 8 (64-row batches) or 16 (32-row batches)
    constexpr int kFeatPerThread = 16 / kParts;              // 2 or 1
    const DwSmem L = dw_smem_layout(N16, ROWS);
    const int nu = N16 / 8, dld = N16 + 4;                   // units of dY (<= 8 with 64-row batches, <= 32 with 32-row ones)
    const uint32_t b_bytes = (uint32_t)nu * kUnit;
    uint8_t* a_st = smem + L.a;
    uint8_t* b_st = smem + L.b;
    float* x_st = reinterpret_cast<float*>(smem + L.xs);
    float* d_st = reinterpret_cast<float*>(smem + L.ds);
    uint4* lut = reinterpret_cast<uint4*>(smem + L.lut);
    uint64_t* bar_stage = reinterpret_cast<uint64_t*>(smem + L.bars);        // [3]: the MMAs that read A stage s are done
    uint64_t* bar_batch = bar_stage + kMStages;                              // [2]: all MMAs of a batch are done (its B buffer is free)
    uint64_t* bar_in = bar_batch + 2;                                        // [2]: x and dY rows of a batch have landed
    uint64_t* bar_free = bar_in + 2;                                         // [2]: every producer is done with a staging buffer
    uint64_t* bar_done = bar_free + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar_done + 1);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const bool control = warp == kMProducers / 32, loader = tid >= kMSync;
    const int fb0 = blockIdx.x * blocks_per_pass, nb = min(blocks_per_pass, fblocks - fb0);
    const long long r_beg = (long long)blockIdx.y * rows_per_slab, r_end = min(n_rows, r_beg + rows_per_slab);
    // staged x columns [xc0, xc0 + xcols): the pass's features; with 16-byte aligned rows (x16: in_f and the row stride are
    // multiples of 4) widened to 16-byte boundaries and copied 16 bytes at a time, else exactly those and 4 bytes at a time
    const int xc0 = x16 ? (fb0 * kFB) & ~3 : fb0 * kFB, xoff = fb0 * kFB - xc0;
    const int xcols = min(x16 ? (xoff + nb * kFB + 3) & ~3 : nb * kFB, g.in_f - xc0);

    if (control) {
        tc::tmem_alloc(tmem_slot, tmem_cols);
        if (lane == 0) {
            for (int i = 0; i < kMStages + 2; ++i) tc::mbar_init(&bar_stage[i], 1);
            tc::mbar_init(&bar_in[0], kMLoaders);
            tc::mbar_init(&bar_in[1], kMLoaders);
            tc::mbar_init(&bar_free[0], 1);
            tc::mbar_init(&bar_free[1], 1);
            tc::mbar_init(bar_done, 1);
            tc::mbar_fence_init();
        }
    }
    if (K > 0 && tid < kLutRows) lut[tid] = lut_row_b(tid, K, g.G + 2 * K);
    tc::tc_fence_before_sync();
    __syncthreads();
    tc::tc_fence_after_sync();
    const uint32_t tmem_base = *tmem_slot;

    if (control) {
        // ================================================== control warp ==================================================
        const uint32_t idesc = idesc_bf16_f32_mn(128, N16);
        // descriptors of stage 0 / buffer 0; the others are a constant 16-byte-unit offset away (the address field is the low 14 bits)
        const uint64_t dah0 = tc::smem_desc(tc::smem_u32(a_st), 128, kUnit), dal0 = tc::smem_desc(tc::smem_u32(a_st + 16u * kUnit), 128, kUnit);
        const uint64_t dbh0 = tc::smem_desc(tc::smem_u32(b_st), 128, kUnit), dbl0 = tc::smem_desc(tc::smem_u32(b_st + b_bytes), 128, kUnit);
        int st = 0, bbuf = 0;
        long long t = 0;
        for (long long rt = r_beg; rt < r_end; rt += ROWS, ++t, bbuf ^= 1) {
            for (int b = 0; b < nb; ++b) {
                if (st == 0) named_sync<1, kMSync>();
                else if (st == 1) named_sync<2, kMSync>();
                else named_sync<3, kMSync>();
                if (lane == 0) {
                    tc::tc_fence_after_sync();
                    const uint32_t d_col = tmem_base + (uint32_t)(b * N16);
                    const uint64_t a_off = (uint64_t)((uint32_t)st * (kAStage >> 4)), b_off = (uint64_t)((uint32_t)bbuf * ((2u * b_bytes) >> 4));
                    uint32_t acc = t == 0 ? 0u : 1u;
#pragma unroll
                    for (int ks = 0; ks < ROWS / 16; ++ks) { // K = 16 rows per step: two groups of 8 k-rows, 256 bytes
                        const uint64_t ko = (uint64_t)(ks * 16);
                        tc::umma_bf16(d_col, dah0 + a_off + ko, dbh0 + b_off + ko, idesc, acc);
                        tc::umma_bf16(d_col, dah0 + a_off + ko, dbl0 + b_off + ko, idesc, 1);
                        tc::umma_bf16(d_col, dal0 + a_off + ko, dbh0 + b_off + ko, idesc, 1);
                        acc = 1;
                    }
                    tc::umma_commit(&bar_stage[st]);
                    if (b == nb - 1) {
                        tc::umma_commit(&bar_batch[bbuf]);
                        tc::mbar_arrive(&bar_free[bbuf]);    // every producer is past this batch's last read of the staging tiles
                    }
                }
                __syncwarp();
                st = st == kMStages - 1 ? 0 : st + 1;
            }
        }
        if (lane == 0) tc::umma_commit(bar_done);
        __syncwarp();
    } else if (loader) {
        // ===================================================== loaders =====================================================
        const int lt = tid - kMSync;
        // (row, chunk) pairs, consecutive lanes along the row; e -> e + 128 advances (r, c) by (q, rem) with one carry.  The dY
        // tile is staged N16 wide: the columns from out_f on are zero-filled (they become zero columns of the operand).
        const int xsh = x16 ? 2 : 0, dsh = dy16 ? 2 : 0;
        const int xcn = (xcols + (1 << xsh) - 1) >> xsh, dcn = N16 >> dsh;
        const int xq = kMLoaders / xcn, xr = kMLoaders - xq * xcn, dq = kMLoaders / dcn, dr = kMLoaders - dq * dcn;
        long long t = 0;
        for (long long rt = r_beg; rt < r_end; rt += ROWS, ++t) {
            const int buf = (int)(t & 1);
            if (t >= 2) tc::mbar_wait(&bar_free[buf], (uint32_t)((t >> 1) - 1) & 1u);
            float* xs = x_st + (size_t)buf * ROWS * kMXld;
            float* ds = d_st + (size_t)buf * ROWS * dld;
            {
                int r = lt / xcn, c = lt - r * xcn;
                while (r < ROWS) {
                    const bool on = rt + r < r_end;
                    const float* src = on ? x + (rt + r) * ldx + xc0 + (c << xsh) : x;
                    if (x16) cp_async16(xs + r * kMXld + 4 * c, src, on);
                    else cp_async_f32(xs + r * kMXld + c, src, on);
                    r += xq;
                    c += xr;
                    if (c >= xcn) {
                        c -= xcn;
                        ++r;
                    }
                }
            }
            {
                int r = lt / dcn, c = lt - r * dcn;
                while (r < ROWS) {
                    const bool on = rt + r < r_end && (c << dsh) < g.out_f;
                    const float* src = on ? dy + (rt + r) * ld_dy + (c << dsh) : dy;
                    if (dy16) cp_async16(ds + r * dld + 4 * c, src, on);
                    else cp_async_f32(ds + r * dld + c, src, on);
                    r += dq;
                    c += dr;
                    if (c >= dcn) {
                        c -= dcn;
                        ++r;
                    }
                }
            }
            cp_async_arrive_on(&bar_in[buf]);
        }
    } else {
        // ==================================================== producers ====================================================
        const int r64 = tid & (ROWS - 1), part = tid / ROWS;       // row of the batch, part of the feature block
        int st = 0, bbuf = 0;
        uint32_t round = 0;                                  // uses of A stage `st` so far
        long long t = 0;
        for (long long rt = r_beg; rt < r_end; rt += ROWS, ++t, bbuf ^= 1) {
            const long long row = rt + r64;
            const bool row_ok = row < r_end;
            float mean = 0.f, rstd = 1.f;
            if (K == 0 && g.stats && row_ok) {
                mean = __ldg(g.stats + 2 * row);
                rstd = __ldg(g.stats + 2 * row + 1);
            }
            tc::mbar_wait(&bar_in[bbuf], (uint32_t)(t >> 1) & 1u);            // this batch's x and dY rows are in shared memory
            if (t >= 2) tc::mbar_wait(&bar_batch[bbuf], (uint32_t)((t >> 1) - 1) & 1u);   // the MMAs that read this B buffer are done
            for (int u = part; u < nu; u += kParts) {            // this thread's 8-column units of dY
                const float* dr = d_st + ((size_t)bbuf * ROWS + r64) * dld + 8 * u;
                const float4 d0 = *reinterpret_cast<const float4*>(dr), d1 = *reinterpret_cast<const float4*>(dr + 4);   // zero outside the matrix
                const float v[8] = {d0.x, d0.y, d0.z, d0.w, d1.x, d1.y, d1.z, d1.w};
                uint4 hi, lo;
                tc::split8(v, hi, lo);
                uint8_t* bh = b_st + (size_t)bbuf * 2u * b_bytes;
                *reinterpret_cast<uint4*>(bh + (size_t)u * kUnit + r64 * 16) = hi;
                *reinterpret_cast<uint4*>(bh + b_bytes + (size_t)u * kUnit + r64 * 16) = lo;
            }
            const float* xrow = x_st + ((size_t)bbuf * ROWS + r64) * kMXld + xoff;
            for (int b = 0; b < nb; ++b) {
                const int f0 = (fb0 + b) * kFB;
                if (round) tc::mbar_wait(&bar_stage[st], (round - 1u) & 1u);  // the MMAs of three blocks ago have read this stage
                uint8_t* a_hi = a_st + (size_t)st * kAStage;
                uint8_t* a_lo = a_hi + 16u * kUnit;
#pragma unroll
                for (int ii = 0; ii < kFeatPerThread; ++ii) {
                    const int i = part + kParts * ii;
                    if (i < kFB) {                          // uniform over the part (one or two warps)
                        const bool on = row_ok && (f0 + i) < g.in_f;
                        const float xv = xrow[b * kFB + i];
                        uint4 hi, lo;
                        if (K == 0) {
                            const float z = rbf_z_b(g, xv, mean, rstd, f0 + i);
                            float e[8];
#pragma unroll
                            for (int q = 0; q < 8; ++q) {
                                const float tt = (z - (g.c0 + (float)q * g.step)) * g.inv_den;
                                e[q] = (on && q < g.S) ? ex2_b(-kLog2eB * tt * tt) : 0.f;
                            }
                            tc::split8(e, hi, lo);
                        } else {
                            int idx;
                            float fr, bv[4];
                            locate_b(g, xv, idx, fr);
                            local_values_b<(K == 0 ? 1 : K)>(fr, bv);
                            const uint32_t h01 = pack_trunc_b(bv[0], bv[1]), h23 = pack_trunc_b(bv[2], bv[3]);
                            const uint32_t l01 = pack_rn_b(trunc_res_b(bv[0]), trunc_res_b(bv[1])), l23 = pack_rn_b(trunc_res_b(bv[2]), trunc_res_b(bv[3]));
                            uint4 sel = lut[idx];
                            if (!on) sel = make_uint4(0x9999u, 0x9999u, 0x9999u, 0x9999u);
                            hi = make_uint4(prmt_b(h01, h23, sel.x), prmt_b(h01, h23, sel.y), prmt_b(h01, h23, sel.z), prmt_b(h01, h23, sel.w));
                            lo = make_uint4(prmt_b(l01, l23, sel.x), prmt_b(l01, l23, sel.y), prmt_b(l01, l23, sel.z), prmt_b(l01, l23, sel.w));
                        }
                        *reinterpret_cast<uint4*>(a_hi + (size_t)i * kUnit + r64 * 16) = hi;
                        *reinterpret_cast<uint4*>(a_lo + (size_t)i * kUnit + r64 * 16) = lo;
                        const float sv = on ? __fdividef(xv, 1.0f + ex2_b(-kLog2eB * xv)) : 0.f;
                        const uint32_t off = (uint32_t)(kFB + (i >> 3)) * kUnit + (uint32_t)r64 * 16u + (uint32_t)(i & 7) * 2u;
                        *reinterpret_cast<uint16_t*>(a_hi + off) = (uint16_t)(__float_as_uint(sv) >> 16);
                        *reinterpret_cast<uint16_t*>(a_lo + off) = (uint16_t)(pack_rn_b(trunc_res_b(sv), 0.f) & 0xffffu);
                    } else {                                // the pad elements (features 14, 15) of unit 15
                        const uint32_t off = (uint32_t)(kFB + 1) * kUnit + (uint32_t)r64 * 16u + (uint32_t)(i & 7) * 2u;
                        *reinterpret_cast<uint16_t*>(a_hi + off) = 0;
                        *reinterpret_cast<uint16_t*>(a_lo + off) = 0;
                    }
                }
                tc::fence_proxy_async_smem();
                if (st == 0) named_arrive<1, kMSync>();
                else if (st == 1) named_arrive<2, kMSync>();
                else named_arrive<3, kMSync>();
                if (st == kMStages - 1) {
                    st = 0;
                    ++round;
                } else {
                    ++st;
                }
            }
        }
        // ---- the slab's sums: every MMA has completed
        if (r_beg < r_end) {
            tc::mbar_wait(bar_done, 0);
            tc::tc_fence_after_sync();
            // lane m of tensor memory = row m of a block's 128 gradient rows; the four warpgroups share the (block, 16-column pair) items
            const int m = tid & 127, u = m >> 3, c = m & 7, quarter = tid >> 7;
            const uint32_t lane_base = (uint32_t)((warp & 3) * 32) << 16;
            const int npair = N16 / 16;
            const int rot = (int)(blockIdx.y % (unsigned)npair) * 16;                // CTAs start at different columns: fewer collisions in L2
            for (int item = quarter; item < nb * npair; item += 4) {
                const int b = item / npair, oo = (item - b * npair) * 16;
                const int f0 = (fb0 + b) * kFB;
                long long dst = -1;
                if (u < kFB) {
                    if (f0 + u < g.in_f && c < g.S) dst = ((long long)(f0 + u) * (g.S + 1) + c) * g.out_pad;
                } else {
                    const int i = (u - kFB) * 8 + c;
                    if (i < kFB && f0 + i < g.in_f) dst = ((long long)(f0 + i) * (g.S + 1) + g.S) * g.out_pad;
                }
                int o0 = oo + rot;
                if (o0 >= N16) o0 -= N16;
                const int o1 = o0 + 8;
                float v[8], w[8];
                tc::tmem_ld8(tmem_base + lane_base + (uint32_t)(b * N16 + o0), v);   // warp-collective: every lane takes part
                tc::tmem_ld8(tmem_base + lane_base + (uint32_t)(b * N16 + o1), w);
                if (dst >= 0) {                             // a 4-vector that straddles out_f adds zeros to the row's pad columns (out_pad = pad4(out_f))
                    if (o0 < g.out_f) red_add_v4(dP + dst + o0, v[0], v[1], v[2], v[3]);
                    if (o0 + 4 < g.out_f) red_add_v4(dP + dst + o0 + 4, v[4], v[5], v[6], v[7]);
                    if (o1 < g.out_f) red_add_v4(dP + dst + o1, w[0], w[1], w[2], w[3]);
                    if (o1 + 4 < g.out_f) red_add_v4(dP + dst + o1 + 4, w[4], w[5], w[6], w[7]);
                }
            }
        }
    }
    tc::tc_fence_before_sync();
    __syncthreads();
    if (control) tc::tmem_dealloc(tmem_base, tmem_cols);
}

int geometry_b(const KagnnKanLayer* L, GeomB* g) {
    if (!L || !L->packed_w || L->basis != KAGNN_BASIS_BSPLINE) return KAGNN_EUNSUPPORTED;
    if (L->in_features <= 0 || L->out_features <= 0 || L->grid_size < 1 || L->spline_order < 1 || L->spline_order > 4) return KAGNN_EUNSUPPORTED;
    if (!(L->h > 0.f)) return KAGNN_EUNSUPPORTED;
    *g = GeomB{};
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

std::atomic<int> g_bwd_path{0};         // 0 = auto (tensor cores first), 1 = fp32 kernels only (tests / comparisons)
std::atomic<int> g_dw_rows64{1};        // dW of layers <= 64 wide on 64-row batches (1) or on the general 128-row kernel (0; mode 2 below)
std::atomic<int> g_dx_packed{1};        // dX reads the forward's packed weights (1) or splits the fp32 weights per tile (0; mode 2 below)
}  // namespace

extern "C" int kagnn_set_backward_path(int32_t mode) {
    if (mode < 0 || mode > 3) return KAGNN_EINVAL;      // 2, 3 = tensor cores with the alternative kernels (tests): see the header
    g_bwd_path.store(mode == 1 ? 1 : 0);
    g_dx_packed.store(mode == 2 ? 0 : (mode == 3 ? 2 : 1));      // 2 = packed operand, no look-ahead
    g_dw_rows64.store(mode == 2 ? 0 : (mode == 3 ? 2 : 1));
    return KAGNN_OK;
}

namespace {
int launch_bwd_input_tc(const GeomB& g, const float* w, const void* wtc, const float* x, int64_t ldx, const float* dy, int64_t ld_dy,
                        int64_t num_rows, float* dx, int64_t ld_dx, float* dxb, int64_t ld_dxb, cudaStream_t stream) {
    if (num_rows < 128 || g.S + 1 > kS1MaxB || g.k > kMaxOrderB || g.out_f > 256) return KAGNN_EUNSUPPORTED;   // small batches: not worth a tile
    DeviceProps props{};
    int rc = kagnn_get_props(&props);
    if (rc != KAGNN_OK) return rc;
    if (props.cc_major != 10) return KAGNN_EUNSUPPORTED;
    const int K = ((g.out_f + 15) / 16) * 16;
    // the layer's tensor-core packing (the forward's operand) doubles as dX's B operand; without it the weights are split per tile
    const bool packed = wtc != nullptr && aligned16(wtc) && g_dx_packed.load() != 0;
    const int S1 = packed ? 9 : g.S + 1;
    const int Np = kFP * S1;
    const uint32_t a_col = (uint32_t)((Np + 31) & ~31);
    if (a_col + (uint32_t)K > 512u) return KAGNN_EUNSUPPORTED;
    const uint32_t cols = tc::tmem_cols_pow2(a_col + (uint32_t)K);
    int n_stage = 2;
    if (packed && (size_t)2 * 576 * K + 128 > (size_t)props.max_smem) n_stage = 1;
    const size_t smem = packed ? (size_t)n_stage * 576 * K + 128 : (size_t)2 * (K / 8) * Np * 16 + 128;
    if (smem > (size_t)props.max_smem) return KAGNN_EUNSUPPORTED;
    const int n_tiles = (int)ceil_div64(num_rows, 128);
    // CTAs per SM: limited by tensor-memory columns and shared memory (the kernel overlaps its phases only across CTAs)
    int per_sm = (int)(512u / cols);
    const int by_smem = (int)((size_t)props.max_smem / (smem + 1024));
    if (per_sm > by_smem) per_sm = by_smem;
    if (per_sm > 4) per_sm = 4;
    if (per_sm < 1) per_sm = 1;
    const int grid = n_tiles < props.num_sms * per_sm ? n_tiles : props.num_sms * per_sm;
    if (packed && K <= 64 && g_dx_packed.load() == 1) {
        // narrow layers: passes of eight features with look-ahead (two T buffers), 256 columns of tensor memory, two CTAs per SM
        const size_t smem_l = (size_t)2 * 576 * K + 128 + (size_t)128 * (K + 4) * sizeof(float);
        const int grid_l = n_tiles < props.num_sms * 2 ? n_tiles : props.num_sms * 2;
        auto kl = g.k == 3 ? kan_bwd_input_tc_look_kernel<3> : (g.k == 2 ? kan_bwd_input_tc_look_kernel<2> : (g.k == 1 ? kan_bwd_input_tc_look_kernel<1> : kan_bwd_input_tc_look_kernel<0>));
        KAGNN_CUDA_TRY(cudaFuncSetAttribute(kl, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)props.max_smem));
        kl<<<(unsigned)grid_l, kLookThreads, smem_l, stream>>>(g, static_cast<const uint8_t*>(wtc), x, (long long)ldx, dy, (long long)ld_dy, (long long)num_rows,
                                                          n_tiles, K, 256u, dx, (long long)ld_dx, dxb, (long long)ld_dxb);
        KAGNN_LAUNCH_CHECK();
        return KAGNN_OK;
    }
    void (*kern)(GeomB, const float*, const uint8_t*, int, const float*, long long, const float*, long long, long long, int, int, uint32_t,
                 float*, long long, float*, long long);
    if (packed)
        kern = g.k == 3 ? kan_bwd_input_tc_kernel<3, true> : (g.k == 2 ? kan_bwd_input_tc_kernel<2, true> : (g.k == 1 ? kan_bwd_input_tc_kernel<1, true> : kan_bwd_input_tc_kernel<0, true>));
    else
        kern = g.k == 3 ? kan_bwd_input_tc_kernel<3, false> : (g.k == 2 ? kan_bwd_input_tc_kernel<2, false> : (g.k == 1 ? kan_bwd_input_tc_kernel<1, false> : kan_bwd_input_tc_kernel<0, false>));
    KAGNN_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)props.max_smem));
    kern<<<(unsigned)grid, kThreads, smem, stream>>>(g, w, static_cast<const uint8_t*>(wtc), n_stage, x, (long long)ldx, dy, (long long)ld_dy,
                                                    (long long)num_rows, n_tiles, K, cols, dx, (long long)ld_dx, dxb, (long long)ld_dxb);
    KAGNN_LAUNCH_CHECK();
    return KAGNN_OK;
}

int launch_bwd_weights_tc(const GeomB& g, const float* x, int64_t ldx, const float* dy, int64_t ld_dy, int64_t num_rows, float* d_packed,
                          cudaStream_t stream) {
    if (num_rows < 128 || g.S > 8 || g.k > kMaxOrderB || g.out_f > 256) return KAGNN_EUNSUPPORTED;
    DeviceProps props{};
    int rc = kagnn_get_props(&props);
    if (rc != KAGNN_OK) return rc;
    if (props.cc_major != 10) return KAGNN_EUNSUPPORTED;
    const int N16 = ((g.out_f + 15) / 16) * 16;
    const uint32_t cols = tc::tmem_cols_pow2((uint32_t)N16);
    const size_t smem = (size_t)2 * 16 * 2048 + (size_t)2 * (N16 / 8) * 2048 + kLutRows * 16 + 64;
    if (smem > (size_t)props.max_smem) return KAGNN_EUNSUPPORTED;
    KAGNN_CUDA_TRY(cudaMemsetAsync(d_packed, 0, sizeof(float) * (size_t)g.in_f * (size_t)(g.S + 1) * (size_t)g.out_pad, stream));
    const int fblocks = (g.in_f + kFB - 1) / kFB;
    // row slabs: about three CTAs per SM over the whole grid, at least 128 rows each
    int64_t slabs = ceil_div64((int64_t)props.num_sms * 3, fblocks);
    if (slabs > ceil_div64(num_rows, 128)) slabs = ceil_div64(num_rows, 128);
    if (slabs > 65535) slabs = 65535;
    if (slabs < 1) slabs = 1;
    int64_t rows_per_slab = ceil_div64(ceil_div64(num_rows, slabs), 128) * 128;
    slabs = ceil_div64(num_rows, rows_per_slab);
    int swap = 0;
#ifdef KAGNN_DEBUG_KNOBS
    if (const char* e = getenv("KAGNN_DEBUG_DW_SWAP")) swap = atoi(e);
#endif
    const bool dy_vec = g.out_f % 4 == 0 && ld_dy % 4 == 0 && aligned16(dy);
    const bool x_vec = g.in_f % 4 == 0 && ldx % 4 == 0 && aligned16(x);            // rows start on 16-byte boundaries
    const int dw_mode = g_dw_rows64.load();                   // 1 = default, 2 / 0 = the alternative kernels (kagnn_set_backward_path 3 / 2)
    if (dw_mode == 1 && g.out_pad % 4 == 0 && aligned16(d_packed)) {
        // several feature blocks per CTA (as many accumulators as tensor memory holds) share one split of the dY batch;
        // 64-row batches for layers up to 64 wide, 32-row batches above (the dY operand of a batch is N16 x rows x 4 bytes)
        const bool narrow = N16 <= 64;
        const int rows = narrow ? 64 : 32;
        int max_blocks = narrow ? kMBlocks : 512 / N16;
#ifdef KAGNN_DEBUG_KNOBS
        if (const char* e = getenv("KAGNN_DEBUG_DW_MBLOCKS")) max_blocks = atoi(e) < 1 ? 1 : (atoi(e) > max_blocks ? max_blocks : atoi(e));
#endif
        const int passes = (fblocks + max_blocks - 1) / max_blocks;
        const int bpp = (fblocks + passes - 1) / passes;                  // balanced: 10 blocks -> 5 + 5
        const uint32_t cols_m = tc::tmem_cols_pow2((uint32_t)(bpp * N16));
        const size_t smem_m = dw_smem_layout(N16, rows).total;
        int64_t slabs_m = (int64_t)props.num_sms / passes;              // one CTA (21 warps) per SM, one wave
        if (slabs_m > ceil_div64(num_rows, 64)) slabs_m = ceil_div64(num_rows, 64);
        if (slabs_m > 65535) slabs_m = 65535;
        if (slabs_m < 1) slabs_m = 1;
        int64_t rps_m = ceil_div64(ceil_div64(num_rows, slabs_m), 64) * 64;
        slabs_m = ceil_div64(num_rows, rps_m);
        // (a wide layer with a single feature block has nothing to share the dY split with: the 128-row kernel below is faster)
        if (smem_m <= (size_t)props.max_smem && cols_m <= 512 && (narrow || fblocks >= 2)) {
            auto km = narrow ? (g.k == 3 ? kan_bwd_weights_tcm_kernel<3, 64> : (g.k == 2 ? kan_bwd_weights_tcm_kernel<2, 64> : (g.k == 1 ? kan_bwd_weights_tcm_kernel<1, 64> : kan_bwd_weights_tcm_kernel<0, 64>)))
                             : (g.k == 3 ? kan_bwd_weights_tcm_kernel<3, 32> : (g.k == 2 ? kan_bwd_weights_tcm_kernel<2, 32> : (g.k == 1 ? kan_bwd_weights_tcm_kernel<1, 32> : kan_bwd_weights_tcm_kernel<0, 32>)));
            KAGNN_CUDA_TRY(cudaFuncSetAttribute(km, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)props.max_smem));
            km<<<dim3((unsigned)passes, (unsigned)slabs_m, 1), kMThreads, smem_m, stream>>>(g, x, (long long)ldx, dy, (long long)ld_dy,
                                                                                            (long long)num_rows, (long long)rps_m, N16, fblocks, bpp, cols_m, x_vec ? 1 : 0, dy_vec ? 1 : 0, d_packed);
            KAGNN_LAUNCH_CHECK();
            return KAGNN_OK;
        }
    }
    if (N16 <= 64 && dy_vec && dw_mode != 0) {
        // (mode 3 of kagnn_set_backward_path: one feature block per CTA) 64-row batches, three CTAs per SM
        const size_t smem64 = (size_t)2 * 16 * 1024 + (size_t)2 * (N16 / 8) * 1024 + kLutRows * 16 + 64;
        int64_t slabs64 = ceil_div64((int64_t)props.num_sms * 6, fblocks);
        if (slabs64 > ceil_div64(num_rows, 64)) slabs64 = ceil_div64(num_rows, 64);
        if (slabs64 > 65535) slabs64 = 65535;
        if (slabs64 < 1) slabs64 = 1;
        int64_t rps = ceil_div64(ceil_div64(num_rows, slabs64), 64) * 64;
        slabs64 = ceil_div64(num_rows, rps);
        auto k64 = g.k == 3 ? kan_bwd_weights_tc64_kernel<3> : (g.k == 2 ? kan_bwd_weights_tc64_kernel<2> : (g.k == 1 ? kan_bwd_weights_tc64_kernel<1> : kan_bwd_weights_tc64_kernel<0>));
        KAGNN_CUDA_TRY(cudaFuncSetAttribute(k64, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)props.max_smem));
        k64<<<dim3((unsigned)fblocks, (unsigned)slabs64, 1), kThreads, smem64, stream>>>(g, x, (long long)ldx, dy, (long long)ld_dy, (long long)num_rows,
                                                                                         (long long)rps, N16, cols, d_packed);
        KAGNN_LAUNCH_CHECK();
        return KAGNN_OK;
    }
    auto kern = g.k == 3 ? kan_bwd_weights_tc_kernel<3> : (g.k == 2 ? kan_bwd_weights_tc_kernel<2> : (g.k == 1 ? kan_bwd_weights_tc_kernel<1> : kan_bwd_weights_tc_kernel<0>));
    KAGNN_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)props.max_smem));
    kern<<<dim3((unsigned)fblocks, (unsigned)slabs, 1), kThreads, smem, stream>>>(
        g, x, (long long)ldx, dy, (long long)ld_dy, (long long)num_rows, (long long)rows_per_slab, N16, cols, swap, d_packed);
    KAGNN_LAUNCH_CHECK();
    return KAGNN_OK;
}

// FastKAN layer -> the kernels' geometry with k = 0
int geometry_rbf_b(const KagnnKanLayer* L, const float* stats, GeomB* g) {
    if (!L || !L->packed_w || L->basis != KAGNN_BASIS_RBF) return KAGNN_EUNSUPPORTED;
    if (L->in_features <= 0 || L->out_features <= 0 || L->grid_size < 1 || L->grid_size > 8) return KAGNN_EUNSUPPORTED;
    if ((L->ln_weight || L->ln_bias) && !stats) return KAGNN_EINVAL;
    *g = GeomB{};
    g->in_f = L->in_features;
    g->out_f = L->out_features;
    g->out_pad = pad4(L->out_features);
    g->G = L->grid_size;
    g->k = 0;
    g->S = L->grid_size;
    g->c0 = L->t0;
    g->step = L->h;
    g->inv_den = L->inv_denominator;
    g->stats = stats;
    g->ln_w = stats ? L->ln_weight : nullptr;
    g->ln_b = stats ? L->ln_bias : nullptr;
    return KAGNN_OK;
}
}  // namespace

int kagnn_kan_bwd_input_tc(const KagnnKanLayer* layer, const float* x, int64_t ldx, const float* dy, int64_t ld_dy, int64_t num_rows,
                           float* dx, int64_t ld_dx, cudaStream_t stream) {
    if (g_bwd_path.load() != 0) return KAGNN_EUNSUPPORTED;
    GeomB g;
    const int rc = geometry_b(layer, &g);
    if (rc != KAGNN_OK) return rc;
    return launch_bwd_input_tc(g, layer->packed_w, layer->packed_w_tc, x, ldx, dy, ld_dy, num_rows, dx, ld_dx, nullptr, 0, stream);
}

int kagnn_kan_bwd_weights_tc(const KagnnKanLayer* layer, const float* x, int64_t ldx, const float* dy, int64_t ld_dy, int64_t num_rows,
                             float* d_packed, cudaStream_t stream) {
    if (g_bwd_path.load() != 0) return KAGNN_EUNSUPPORTED;
    GeomB g;
    const int rc = geometry_b(layer, &g);
    if (rc != KAGNN_OK) return rc;
    return launch_bwd_weights_tc(g, x, ldx, dy, ld_dy, num_rows, d_packed, stream);
}

// FastKAN: dz (through the Gaussians) and dx_base (through the SiLU branch) -- one matrix when there is no LayerNorm
int kagnn_rbf_bwd_input_tc(const KagnnKanLayer* layer, const float* x, int64_t ldx, const float* ln_stats, const float* dy, int64_t ld_dy,
                           int64_t num_rows, float* dz, int64_t ld_dz, float* dx_base, int64_t ld_dxb, cudaStream_t stream) {
    if (g_bwd_path.load() != 0) return KAGNN_EUNSUPPORTED;
    GeomB g;
    const int rc = geometry_rbf_b(layer, ln_stats, &g);
    if (rc != KAGNN_OK) return rc;
    return launch_bwd_input_tc(g, layer->packed_w, layer->packed_w_tc, x, ldx, dy, ld_dy, num_rows, dz, ld_dz, ln_stats ? dx_base : nullptr, ld_dxb,
                               stream);
}

int kagnn_rbf_bwd_weights_tc(const KagnnKanLayer* layer, const float* x, int64_t ldx, const float* ln_stats, const float* dy, int64_t ld_dy,
                             int64_t num_rows, float* d_packed, cudaStream_t stream) {
    if (g_bwd_path.load() != 0) return KAGNN_EUNSUPPORTED;
    GeomB g;
    const int rc = geometry_rbf_b(layer, ln_stats, &g);
    if (rc != KAGNN_OK) return rc;
    return launch_bwd_weights_tc(g, x, ldx, dy, ld_dy, num_rows, d_packed, stream);
}
