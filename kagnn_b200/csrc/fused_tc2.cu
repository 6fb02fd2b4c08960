// kagnn_fused_layer_fwd -- pipelined tensor-core path for KAN chains (tcgen05, A operand in TMEM), sm_100a.
//
// Same contract as fused_fp32.cu / fused_tc.cu (tile = aggregate(x) -> pre-affine -> KAN chain -> post-affine -> y, one
// persistent launch per GNN layer), restructured so that the three resources of the layer run concurrently:
//
//   gather warps   (8 | 16)  CSR gather-sum of the NEXT 128-row tile, one "unit" (128 rows x 64/128 feature columns) at a
//                       time into a shared-memory ring: row pointers and column indices are fetched once per warp and the
//                       neighbour rows of its rows are streamed as ONE flattened list with 8 independent 128-bit loads in
//                       flight per warp (units of at most 64 columns: two source rows per load) -- L2/HBM latency is paid
//                       once per batch, not once per row, and is hidden behind the basis producers of the current tile;
//   producer warps (16 | 8)  thread = row, two teams taking alternate chunks.  For every 8 input features the closed-form
//                       uniform B-spline basis (node_classification_clean/ekan.py:79-112 restricted to its k+1
//                       non-zeros; FastKAN: the eight Gaussians of fastkan.py:46-47), split into bf16 hi + lo, is
//                       placed into the 8 coefficient slots of the feature with one byte-permute per 32-bit word
//                       (selectors from a 13-row shared LUT indexed by the knot interval) and written with tcgen05.st
//                       straight into TENSOR MEMORY, which tcgen05.mma reads as its A operand: the expanded
//                       (N, in, G+k) tensor of the reference exists neither in HBM nor in shared memory;
//   MMA warp       (1)  one elected lane issues, per K = 16 step, the products of the bf16 hi / lo split (fp32 accumulate in
//                       TMEM; error ~2^-17, inside BASELINE.json's 1e-4; bf16 mode: one product);  W chunks arrive by
//                       bulk TMA (loader warp);
//   push warps     (2, `PUSH` instantiation)  node-sharded graphs: relay finished output tiles to the peers' replicas;
//   chained KAN layers read their input rows back from the TMEM accumulator of the previous layer (two accumulator
//   regions, alternated per layer across tiles so the epilogue of tile t overlaps the first MMAs of tile t+1).
// The warp split is a compile-time choice: this file is built twice (fused_tc2_g16.cu: 8 producer + 16 gather warps for
// launches that gather over a CSR; here 16 + 8 for the others), see the note above KAGNN_TC2_ENTRY.
//
// Supported here: B-spline layers with one common spline_order k <= 3 and G + k <= 8, or FastKAN layers with at most 8
// centres; chains of layers up to 128 outputs wide, single layers up to 256; split-K over the input features for launches
// with few tiles and long rows.  Anything else returns KAGNN_EUNSUPPORTED and the dispatcher falls back to fused_tc.cu /
// fused_fp32.cu.
#include "common.cuh"
#include "stage.cuh"
#include "tc_common.cuh"
#include <cstdlib>

// Optional timeline trace of CTA 0 (development builds only: KAGNN_NVCC_EXTRA="-DKAGNN_TRACE=1"): clock stamps per role,
// event counter and event kind into a caller-provided buffer (scripts/trace_tc2.py).
#if defined(KAGNN_TRACE) && !defined(KAGNN_TC2_VARIANT_G16)
__device__ unsigned long long* g_trace = nullptr;
#define TR(role, k, evt)                                                                                           \
    do {                                                                                                             \
        if (blockIdx.x == 0 && g_trace && (k) < 512) g_trace[(((role) * 512) + (k)) * 8 + (evt)] = clock64();        \
    } while (0)
extern "C" int kagnn_debug_set_trace(unsigned long long* buf) {
    return cudaMemcpyToSymbol(g_trace, &buf, sizeof(buf)) == cudaSuccess ? 0 : -5;
}
#define TRL(role, k, evt) TR(role, k, evt)            // tile-level events (cheap: a few per tile)
#if KAGNN_TRACE >= 2
#define TRC(role, k, evt) TR(role, k, evt)            // per-chunk events (each costs ~100 cycles: perturbs the producers)
#else
#define TRC(role, k, evt) do { } while (0)
#endif
#else
#define TRL(role, k, evt) do { } while (0)
#define TRC(role, k, evt) do { } while (0)
#endif

namespace {

constexpr int BM = 128;
#ifndef KAGNN_TC2_GATHER_U
#define KAGNN_TC2_GATHER_U 8            // 128-bit row loads in flight per gather warp (register gather; the asynchronous gather is the default path)
#endif
#ifndef KAGNN_TC2_PREFETCH_DIST
#define KAGNN_TC2_PREFETCH_DIST 0        // tiles an L2 prefetch warp may run ahead of the gather; 0 = off (default: measured on the
                                         // bench, distance 3 makes the GIN layers 8 % SLOWER -- the extra requests cost more than the
                                         // DRAM latency they hide; through cp.async.bulk.prefetch they also delay the W loader, 1.7x)
#endif
#ifndef KAGNN_TC2_HUB
#define KAGNN_TC2_HUB 1
#endif
#ifndef KAGNN_TC2_PIPE_EPI
#define KAGNN_TC2_PIPE_EPI 1
#endif
#ifndef KAGNN_TC2_BAL
#define KAGNN_TC2_BAL 1
#endif
#ifndef KAGNN_TC2_NPW
#define KAGNN_TC2_NPW 16
#endif
constexpr int NPW = KAGNN_TC2_NPW;           // producer warps: 2 or 4 warpgroups, each expands 8 / NWG features of every chunk
constexpr int NWG = NPW / 4;
// Every warpgroup expands 8 / NWG features of EVERY chunk (lowest latency per chunk; the alternative -- whole chunks handed round
// robin to the warpgroups -- was measured 3-8 % slower on the GIN layers and needs ring depth >= NWG).
#ifndef KAGNN_TC2_CPR
#define KAGNN_TC2_CPR 2
#endif
// CPR = chunks produced concurrently: the NWG warpgroups form CPR teams, chunk number c (counted over the whole launch) belongs to
// team c % CPR, and the NWG / CPR warpgroups of a team split its 8 features.  CPR = 1: every warpgroup on every chunk (lowest
// latency per chunk, one barrier round per 8 / NWG features of a thread); CPR = 2: two chunks in flight, half as many barrier
// rounds per thread (each covers 4 features), needs two stages per team to keep the tensor pipe fed.
constexpr int CPR = (NWG >= 2 * KAGNN_TC2_CPR || KAGNN_TC2_CPR == 1) ? KAGNN_TC2_CPR : 1;
constexpr int WGT = NWG / CPR;               // warpgroups per team
constexpr int FPW = 8 / WGT;                 // features per warpgroup per spline chunk
#ifndef KAGNN_TC2_ELECT_ARRIVE
#define KAGNN_TC2_ELECT_ARRIVE 0
#endif
#ifndef KAGNN_TC2_WDIV
#define KAGNN_TC2_WDIV 1
#endif
#ifndef KAGNN_TC2_X2
#define KAGNN_TC2_X2 1                       // packed-pair (f32x2) polynomial evaluation in the basis producers
#endif
#ifndef KAGNN_TC2_NOMMA
#define KAGNN_TC2_NOMMA 0                    // development probe: the MMA warp only commits (results are wrong)
#endif
#ifndef KAGNN_TC2_STACK4
#define KAGNN_TC2_STACK4 1
#endif
#ifndef KAGNN_TC2_HALF
#define KAGNN_TC2_HALF 1                     // two rows per load for units of at most 64 columns (gather_unit_half)
#endif
#ifndef KAGNN_TC2_USE_G16
#define KAGNN_TC2_USE_G16 1                  // launches with a CSR gather run the 8-producer / 16-gather-warp build (fused_tc2_g16.cu)
#endif
#ifndef KAGNN_TC2_NOMATH
#define KAGNN_TC2_NOMATH 0                   // development probe: producers skip the basis expansion (results are wrong)
#endif
// arrivals on full[s]: every producer thread (or one elected lane per producer warp) + the W loader's expect_tx
constexpr int FULL_ARRIVALS = (KAGNN_TC2_ELECT_ARRIVE ? NPW / CPR : NPW / CPR * 32) + 1;
#ifndef KAGNN_TC2_NGW
#define KAGNN_TC2_NGW 8
#endif
constexpr int NGW = KAGNN_TC2_NGW;           // gather warps (8 or 16; NPW + NGW = 24: six warpgroups + the MMA / loader warpgroup)
static_assert(NPW + NGW == 24 && (NGW == 8 || NGW == 16), "role split");
// registers after setmaxnreg (launch: 72 per thread, 896 threads): 16 producer + 8 gather warps -> producers 72, gather 88;
// 8 producer + 16 gather warps -> producers 80, gather 80; the MMA / loader warpgroup gives up the difference either way
constexpr int REGS_PROD = NPW == 16 ? 72 : 80, REGS_GATHER = NGW == 8 ? 88 : 80, REGS_AUX = NPW == 16 ? 40 : 24;
constexpr int NTHREADS = (NPW + NGW + 4) * 32;   // + MMA, W loader and two idle warps (whole warpgroups for setmaxnreg)
constexpr int WARP_MMA = NPW + NGW;
constexpr int WARP_LOAD = NPW + NGW + 1;
constexpr int WARP_PUSH = NPW + NGW + 3;             // relays finished output tiles to the peers (KagnnAggregate.push_y)
constexpr int RELAY_PAD = 128 + 128;                 // pushed output: mbarrier / tile counter block + alignment slack in front of the relay slots
constexpr int RPW = BM / NGW;                // rows per gather warp
constexpr int MAX_STAGE = 4;                 // A stages in TMEM (64 columns each) / B stages in shared memory
constexpr int MAX_UNITS = 4;                 // x-tile ring slots
constexpr uint32_t TMEM_A0 = 256;            // columns [0,256): two accumulator regions of 128; [256,512): A stages
constexpr int LUT_ROWS = 13;                  // per layer: knot interval j = -1 .. 11 (row j + 1)
constexpr float kLog2e = 1.4426950408889634f;
constexpr float kMagic = 12582912.0f;        // 1.5 * 2^23: u + kMagic (round down) = floor(u) + kMagic

struct LayerT2 {
    int F, F_pad, N, N_pad, n_chunks;
    int stack;                               // 1: [W_hi | W_lo] is one B operand (N_pad <= 64): accumulator = [hi.hi + lo.hi | hi.lo]
    float c0, inv_h, lim;                    // B-spline: u = x * inv_h + c0 (= (x - t0)/h); valid iff 0 <= u < lim = G + 2k
    float rc0, rstep, rk;                    // RBF: centres rc0 + g * rstep, k = sqrt(log2 e) / denominator: phi = 2^-((z - c) k)^2
    const float *bias, *lnw, *lnb;           // RBF: base_linear.bias, LayerNorm weight / bias (NULL = none)
    const float* ln_stats;                   // RBF, layer 0: precomputed per-row (mean, rstd) or NULL (computed here)
    const uint8_t* wtc;
};

struct Tc2Params {
    KagnnAggregate agg;
    KagnnAffine pre, post;
    int has_pre, has_post;
    long long num_rows;
    float* agg_out;
    long long ld_agg_out;
    float* y;
    long long ldy;
    int n_layers, n_tiles, y_vec;
    int uw, uw_shift, xld, n_units, units_per_tile, unit_floats;   // x-tile ring geometry (uw = 1 << uw_shift = 64 or 128)
    int ns, bstage_bytes;                                 // A/B stage ring depth, bytes of one B stage
    int ag;                                               // 1: asynchronous (cp.async ring) gather, 64-column units
    int wide;                                             // 1: one layer up to 256 outputs wide: a single 256-column accumulator region
    int bf16;                                             // 1: single bf16 product (A_hi . W_hi), kagnn_set_precision(KAGNN_PREC_BF16)
    int n_items;                                          // work items of the persistent loops: n_tiles x n_split
    int n_split, units_per_item;                          // split-K (few tiles, long rows): an item = one tile x a window of x units; partial sums meet in y by float atomics
    int relay_bytes;                                      // pushed output: shared-memory relay of the push warps (16 KB if the rings keep their depth, else 8 KB)
    LayerT2 layers[KAGNN_MAX_LAYERS];
};

// Chunk order of one layer: per group of 64 input features, up to 8 spline chunks (8 features x 8 slots, K = 64) followed by
// one SiLU chunk (the group's n_oct feature octets, K = 8 n_oct).  Every role walks the same sequence with this cursor
// (no divisions in the per-chunk paths).
struct ChunkCursor {
    int group, j, n_oct, octs;
    __device__ __forceinline__ explicit ChunkCursor(int F_pad, int g0 = 0) : group(g0), j(0), n_oct(min(8, (F_pad >> 3) - 8 * g0)), octs(F_pad >> 3) {}
    __device__ __forceinline__ bool base() const { return j == n_oct; }
    __device__ __forceinline__ int nk() const { return j == n_oct ? n_oct : 8; }
    __device__ __forceinline__ uint32_t b_off(int N_pad) const { return (uint32_t)(group * 9 + j) * 256u * (uint32_t)N_pad; }
    __device__ __forceinline__ uint32_t b_bytes(int N_pad) const { return 32u * (uint32_t)nk() * (uint32_t)N_pad; }
    __device__ __forceinline__ void next() {
        if (j == n_oct) {
            ++group;
            j = 0;
            n_oct = min(8, octs - 8 * group);
        } else {
            ++j;
        }
    }
};

// Work item of the persistent loops: a 128-row tile, or (split-K: one KAN layer, no epilogue affine, fewer tiles than half the
// SMs) a tile x a window of x units.  ub0 / ub1 = the window in units, g0 = its first 64-feature group, n_chunks = its chunks.
struct WorkItem {
    int tile, ub0, ub1, g0, n_chunks;
};
#ifdef KAGNN_TC2_VARIANT_G16
constexpr bool kSplitK = false;            // the gather-heavy build never splits (its launches gather over a CSR) and has no registers to spare
#else
constexpr bool kSplitK = true;
#endif
__device__ __forceinline__ WorkItem work_item(const Tc2Params& p, int item) {
    WorkItem w;
    if (!kSplitK || p.n_split <= 1) {
        w.tile = item;
        w.ub0 = 0;
        w.ub1 = p.units_per_tile;
        w.g0 = 0;
        w.n_chunks = p.layers[0].n_chunks;
        return w;
    }
    w.tile = item / p.n_split;
    const int sp = item - w.tile * p.n_split;
    w.ub0 = sp * p.units_per_item;
    w.ub1 = min(p.units_per_tile, w.ub0 + p.units_per_item);
    const int gpu_ = p.uw >> 6;                            // 64-feature groups per unit
    const int octs = p.layers[0].F_pad >> 3;
    w.g0 = w.ub0 * gpu_;
    const int g1 = min((octs + 7) >> 3, w.ub1 * gpu_);
    w.n_chunks = (min(octs, 8 * g1) - 8 * w.g0) + (g1 - w.g0);
    return w;
}

__device__ __forceinline__ float ex2_approx(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
// silu(x) = x / (1 + exp(-x)); "+ (x - x)" turns +inf into NaN like the reference's spline branch does (inf * 0)
__device__ __forceinline__ float silu_nan(float x) { return __fdividef(x, 1.0f + ex2_approx(-kLog2e * x)) + (x - x); }

// prmt.b32 with the full PTX semantics (selector nibble bit 3 = replicate the sign of the selected byte); the
// __byte_perm intrinsic masks the selector with 0x7777 and cannot be used for that
__device__ __forceinline__ uint32_t prmt(uint32_t a, uint32_t b, uint32_t sel) {
    uint32_t d;
    asm("prmt.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(sel));
    return d;
}
__device__ __forceinline__ uint32_t pack_trunc(float lo_elem, float hi_elem) {   // two fp32 -> bf16x2 by truncation
    return prmt(__float_as_uint(lo_elem), __float_as_uint(hi_elem), 0x7632u);
}
__device__ __forceinline__ float trunc_residual(float v) { return v - __uint_as_float(__float_as_uint(v) & 0xffff0000u); }
__device__ __forceinline__ uint32_t pack_rn(float lo_elem, float hi_elem) {
    uint32_t r;
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi_elem), "f"(lo_elem));
    return r;
}

// One input value -> the 8 coefficient slots of its feature as bf16 hi (4 words) and lo (4 words).
// Uniform knots t_j = t0 + j h (ekan.py:28-37): interval j = floor(u), the k+1 non-zero bases B_{j-k..j} are the local
// polynomials of the fractional position; outside [t_0, t_last) and for NaN every basis is 0 (half-open indicator, :95).
template <int K>
__device__ __forceinline__ void bspline_slots(float inv_h, float c0, float limp, const uint4* __restrict__ lut, float x, uint32_t* hi,
                                              uint32_t* lo) {
    // u clamped to [-0.5, lim + 0.5]: out-of-range inputs and NaN (fmaxf drops it) land in interval -1 or lim, whose LUT rows
    // select nothing; the fractional position stays in [0, 1) so every basis value below is finite and >= 0
    const float u = fminf(fmaxf(fmaf(x, inv_h, c0), -0.5f), limp);
    const float t = __fadd_rd(u, kMagic);
    const float fr = u - (t - kMagic);
    const int idx = __float_as_int(t) - (0x4B400000 - 1);
    float b0, b1, b2 = 0.f, b3 = 0.f;
    if (K == 3) {
        const float omf = 1.0f - fr, f2 = fr * fr;
        b0 = omf * omf * (omf * (1.0f / 6.0f));
        b3 = f2 * (fr * (1.0f / 6.0f));
        b1 = fmaf(f2, fmaf(fr, 0.5f, -1.0f), 2.0f / 3.0f);
        b2 = fmaf(fr, fmaf(fr, fmaf(fr, -0.5f, 0.5f), 0.5f), 1.0f / 6.0f);
    } else if (K == 2) {
        const float omf = 1.0f - fr;
        b0 = 0.5f * omf * omf;
        b2 = 0.5f * fr * fr;
        b1 = fmaf(fr, omf, 0.5f);            // (-2 fr^2 + 2 fr + 1) / 2
    } else {
        b0 = 1.0f - fr;
        b1 = fr;
    }
    // every value is >= 0, so hi = truncation to bf16 and lo = bf16(b - hi) are >= 0 too: their sign bits are 0, which
    // lets the byte-permute synthesise the zero slots by sign replication (selector nibble 9)
    const uint32_t h01 = pack_trunc(b0, b1), h23 = pack_trunc(b2, b3);
#if KAGNN_TC2_BAL
    // residuals from the packed hi words: the even element's hi is a left shift (fma pipe), the odd one's a mask (alu pipe);
    // the LUT row address is one multiply-add on the float's bit pattern -- both move work off the busier alu pipe
    const uint32_t l01 = pack_rn(b0 - __uint_as_float(h01 << 16), b1 - __uint_as_float(h01 & 0xffff0000u));
    const uint32_t l23 = (K >= 2) ? pack_rn(b2 - __uint_as_float(h23 << 16), b3 - __uint_as_float(h23 & 0xffff0000u)) : 0u;
    uint4 sel;
    {
        const uint32_t addr = __float_as_uint(t) * 16u + (tc::smem_u32(lut) - (uint32_t)((0x4B400000u - 1u) * 16u));
        asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(sel.x), "=r"(sel.y), "=r"(sel.z), "=r"(sel.w) : "r"(addr));
        (void)idx;
    }
#else
    const uint32_t l01 = pack_rn(trunc_residual(b0), trunc_residual(b1));
    const uint32_t l23 = (K >= 2) ? pack_rn(trunc_residual(b2), trunc_residual(b3)) : 0u;
    const uint4 sel = lut[idx];
#endif
    hi[0] = prmt(h01, h23, sel.x);
    hi[1] = prmt(h01, h23, sel.y);
    hi[2] = prmt(h01, h23, sel.z);
    hi[3] = prmt(h01, h23, sel.w);
    lo[0] = prmt(l01, l23, sel.x);
    lo[1] = prmt(l01, l23, sel.y);
    lo[2] = prmt(l01, l23, sel.z);
    lo[3] = prmt(l01, l23, sel.w);
}

// ---- packed fp32 pairs (FFMA2 / FMUL2 / FADD2 of sm_100): the same IEEE operations on two values per issue slot ----------------
__device__ __forceinline__ unsigned long long f2_bits(float2 v) { return *reinterpret_cast<unsigned long long*>(&v); }
__device__ __forceinline__ float2 bits_f2(unsigned long long b) { return *reinterpret_cast<float2*>(&b); }
__device__ __forceinline__ float2 fma2(float2 a, float2 b, float2 c) {
    unsigned long long d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(f2_bits(a)), "l"(f2_bits(b)), "l"(f2_bits(c)));
    return bits_f2(d);
}
__device__ __forceinline__ float2 mul2(float2 a, float2 b) {
    unsigned long long d;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(f2_bits(a)), "l"(f2_bits(b)));
    return bits_f2(d);
}
__device__ __forceinline__ float2 add2(float2 a, float2 b) {
    unsigned long long d;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(f2_bits(a)), "l"(f2_bits(b)));
    return bits_f2(d);
}
__device__ __forceinline__ float2 sub2(float2 a, float2 b) {
    unsigned long long d;
    asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(f2_bits(a)), "l"(f2_bits(b)));
    return bits_f2(d);
}
__device__ __forceinline__ float2 add2_rm(float2 a, float2 b) {
    unsigned long long d;
    asm("add.rm.f32x2 %0, %1, %2;" : "=l"(d) : "l"(f2_bits(a)), "l"(f2_bits(b)));
    return bits_f2(d);
}
__device__ __forceinline__ float2 splat(float v) { return make_float2(v, v); }

// Two input values (two features of one row) -> their slot words, bit-identical to two bspline_slots calls: the polynomial part
// runs as packed pairs (element .x = first feature, .y = second), which halves its issue slots.
template <int K, bool BF16 = false>
__device__ __forceinline__ void bspline_slots2(float inv_h, float c0, float limp, const uint4* __restrict__ lut, float xa, float xb,
                                               uint32_t* hi, uint32_t* lo) {
    float2 u = fma2(make_float2(xa, xb), splat(inv_h), splat(c0));
    u.x = fminf(fmaxf(u.x, -0.5f), limp);
    u.y = fminf(fmaxf(u.y, -0.5f), limp);
    const float2 t = add2_rm(u, splat(kMagic));
    const float2 fr = sub2(u, sub2(t, splat(kMagic)));
    float2 b0, b1, b2 = splat(0.f), b3 = splat(0.f);
    if (K == 3) {
        const float2 omf = sub2(splat(1.0f), fr), f2 = mul2(fr, fr);
        b0 = mul2(mul2(omf, omf), mul2(omf, splat(1.0f / 6.0f)));
        b3 = mul2(f2, mul2(fr, splat(1.0f / 6.0f)));
        b1 = fma2(f2, fma2(fr, splat(0.5f), splat(-1.0f)), splat(2.0f / 3.0f));
        b2 = fma2(fr, fma2(fr, fma2(fr, splat(-0.5f), splat(0.5f)), splat(0.5f)), splat(1.0f / 6.0f));
    } else if (K == 2) {
        const float2 omf = sub2(splat(1.0f), fr);
        b0 = mul2(mul2(splat(0.5f), omf), omf);
        b2 = mul2(mul2(splat(0.5f), fr), fr);
        b1 = fma2(fr, omf, splat(0.5f));
    } else {
        b0 = sub2(splat(1.0f), fr);
        b1 = fr;
    }
    if (BF16) {
        // single-product precision: round the bases to bf16 (nearest), no residual words
        const uint32_t g01a = pack_rn(b0.x, b1.x), g23a = pack_rn(b2.x, b3.x), g01b = pack_rn(b0.y, b1.y), g23b = pack_rn(b2.y, b3.y);
        const uint32_t lbase = tc::smem_u32(lut) - (uint32_t)((0x4B400000u - 1u) * 16u);
        const uint32_t aa = __float_as_uint(t.x) * 16u + lbase, ab = __float_as_uint(t.y) * 16u + lbase;
        uint4 sa, sb;
        asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(sa.x), "=r"(sa.y), "=r"(sa.z), "=r"(sa.w) : "r"(aa));
        asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(sb.x), "=r"(sb.y), "=r"(sb.z), "=r"(sb.w) : "r"(ab));
        hi[0] = prmt(g01a, g23a, sa.x); hi[1] = prmt(g01a, g23a, sa.y); hi[2] = prmt(g01a, g23a, sa.z); hi[3] = prmt(g01a, g23a, sa.w);
        hi[4] = prmt(g01b, g23b, sb.x); hi[5] = prmt(g01b, g23b, sb.y); hi[6] = prmt(g01b, g23b, sb.z); hi[7] = prmt(g01b, g23b, sb.w);
        return;
    }
    const uint32_t h01a = pack_trunc(b0.x, b1.x), h23a = pack_trunc(b2.x, b3.x);
    const uint32_t h01b = pack_trunc(b0.y, b1.y), h23b = pack_trunc(b2.y, b3.y);
    const float2 r0 = sub2(b0, make_float2(__uint_as_float(h01a << 16), __uint_as_float(h01b << 16)));
    const float2 r1 = sub2(b1, make_float2(__uint_as_float(h01a & 0xffff0000u), __uint_as_float(h01b & 0xffff0000u)));
    uint32_t l23a = 0u, l23b = 0u;
    if (K >= 2) {
        const float2 r2 = sub2(b2, make_float2(__uint_as_float(h23a << 16), __uint_as_float(h23b << 16)));
        const float2 r3 = sub2(b3, make_float2(__uint_as_float(h23a & 0xffff0000u), __uint_as_float(h23b & 0xffff0000u)));
        l23a = pack_rn(r2.x, r3.x);
        l23b = pack_rn(r2.y, r3.y);
    }
    const uint32_t l01a = pack_rn(r0.x, r1.x), l01b = pack_rn(r0.y, r1.y);
    uint4 sa, sb;
    {
        const uint32_t lbase = tc::smem_u32(lut) - (uint32_t)((0x4B400000u - 1u) * 16u);
        const uint32_t aa = __float_as_uint(t.x) * 16u + lbase, ab = __float_as_uint(t.y) * 16u + lbase;
        asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(sa.x), "=r"(sa.y), "=r"(sa.z), "=r"(sa.w) : "r"(aa));
        asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(sb.x), "=r"(sb.y), "=r"(sb.z), "=r"(sb.w) : "r"(ab));
    }
    hi[0] = prmt(h01a, h23a, sa.x); hi[1] = prmt(h01a, h23a, sa.y); hi[2] = prmt(h01a, h23a, sa.z); hi[3] = prmt(h01a, h23a, sa.w);
    lo[0] = prmt(l01a, l23a, sa.x); lo[1] = prmt(l01a, l23a, sa.y); lo[2] = prmt(l01a, l23a, sa.z); lo[3] = prmt(l01a, l23a, sa.w);
    hi[4] = prmt(h01b, h23b, sb.x); hi[5] = prmt(h01b, h23b, sb.y); hi[6] = prmt(h01b, h23b, sb.z); hi[7] = prmt(h01b, h23b, sb.w);
    lo[4] = prmt(l01b, l23b, sb.x); lo[5] = prmt(l01b, l23b, sb.y); lo[6] = prmt(l01b, l23b, sb.z); lo[7] = prmt(l01b, l23b, sb.w);
}

// FastKAN: the 8 Gaussians exp(-((z - c_g)/den)^2) of one layer-normalised input (fastkan.py:46-47) as bf16 hi / lo slot words.
// All 8 slots are dense (no placement); slots past num_grids meet zero weights.
template <bool BF16 = false>
__device__ __forceinline__ void rbf_slots(float rc0, float rstep, float rk, float z, uint32_t* hi, uint32_t* lo) {
    float v[8];
#pragma unroll
    for (int g = 0; g < 8; ++g) {
        const float d = (z - (rc0 + (float)g * rstep)) * rk;
        v[g] = ex2_approx(-d * d);
    }
#pragma unroll
    for (int g = 0; g < 8; g += 2) {
        if (BF16) {
            hi[g / 2] = pack_rn(v[g], v[g + 1]);
        } else {
            hi[g / 2] = pack_trunc(v[g], v[g + 1]);
            lo[g / 2] = pack_rn(trunc_residual(v[g]), trunc_residual(v[g + 1]));
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// GATHER: one warp aggregates RPW destination rows of one unit (column block [c0, c0 + ucols)) into the ring slot.
// ---------------------------------------------------------------------------------------------------------------------
template <bool VEC>
__device__ __forceinline__ void ldp4(const float* __restrict__ ptr, bool on, float (&v)[4], const bool (&cv)[4]) {
    if (VEC) {
        if (on && cv[0]) {
            const float4 t = __ldg(reinterpret_cast<const float4*>(ptr));
            v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
        } else {
            v[0] = v[1] = v[2] = v[3] = 0.f;
        }
    } else {
#pragma unroll
        for (int i = 0; i < 4; ++i) v[i] = (on && cv[i]) ? __ldg(ptr + 32 * i) : 0.f;
    }
}
__device__ __forceinline__ const float* shfl_ptr(const float* p, int src_lane) {
    const unsigned long long v = reinterpret_cast<unsigned long long>(p);
    const uint32_t lo = __shfl_sync(0xffffffffu, (uint32_t)v, src_lane), hi = __shfl_sync(0xffffffffu, (uint32_t)(v >> 32), src_lane);
    return reinterpret_cast<const float*>(((unsigned long long)hi << 32) | lo);
}

// One gather warp, one unit: RPW destination rows x the unit's column block.  The self term of a row is treated as one
// more entry in front of the row's CSR entries ("virtual" entries), so the warp walks ONE flattened list
//     [self_0, nbrs of row 0..., self_1, nbrs of row 1..., ...]
// in batches of 32 (lane-parallel: row lookup, column index, source-row pointer, weight) and sub-batches of U
// independent 128-bit row loads; sums are kept in registers in CSR order (deterministic, no atomics) and a finished row
// gets the mean scale / pre-affine (GCNConv bias, eval BatchNorm, SiLU) and is parked in the ring slot (and agg_out).
template <bool VEC, bool GINE>
__device__ __forceinline__ void gather_unit(const Tc2Params& p, long long row0, int c0, int ucols, float* __restrict__ xsu, int gw,
                                            int lane) {
    constexpr int U = GINE ? 4 : 8;
    const KagnnAggregate& a = p.agg;
    const int F = a.num_cols, mode = a.mode, xld = p.xld;
    const bool segment = (mode == KAGNN_AGG_SEGMENT_SUM) || (mode == KAGNN_AGG_SEGMENT_MEAN);
    const int rl0 = gw * RPW;
    const int cl = VEC ? 4 * lane : lane;                  // lane -> 4 columns: VEC c0+4*lane+i, scalar c0+lane+32*i
    const int cbase = c0 + cl;
    bool cv[4], cin[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int lc = cl + (VEC ? i : 32 * i);
        cin[i] = lc < ucols;
        cv[i] = cin[i] && (c0 + lc) < F;
    }
    if (mode == KAGNN_AGG_NONE && !p.has_pre && !p.agg_out) {
        // plain row tile (bare KANLinear / KAN chain input): coalesced copy, rows past the end zero-filled
#pragma unroll 1
        for (int rb = 0; rb < RPW; rb += 8) {
            float v[8][4];
            // two-part rows: units below num_head_cols come from x_head, the others from x (unit-aligned split)
            const bool head = c0 < a.num_head_cols;
            const float* src = head ? a.x_head : a.x;
            const long long ld = head ? a.ld_head : a.ldx;
            const int cb = head ? cbase : cbase - a.num_head_cols;
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const long long r = row0 + rl0 + rb + u;
                const bool on = r < p.num_rows;
                ldp4<VEC>(src + (on ? r : 0) * ld + cb, on, v[u], cv);
            }
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                float* d = xsu + (rl0 + rb + u) * xld + cl;
                if (VEC) {
                    if (cin[0]) *reinterpret_cast<float4*>(d) = make_float4(v[u][0], v[u][1], v[u][2], v[u][3]);
                } else {
#pragma unroll
                    for (int i = 0; i < 4; ++i)
                        if (cin[i]) d[32 * i] = v[u][i];
                }
            }
        }
        return;
    }
    float sc[4] = {1.f, 1.f, 1.f, 1.f}, sh[4] = {0.f, 0.f, 0.f, 0.f};
    if (p.has_pre) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int c = cbase + (VEC ? i : 32 * i);
            if (cv[i] && p.pre.scale) sc[i] = __ldg(p.pre.scale + c);
            if (cv[i] && p.pre.shift) sh[i] = __ldg(p.pre.shift + c);
        }
    }
    const bool pre_silu = p.has_pre && p.pre.act == KAGNN_ACT_SILU;
    const int self1 = segment ? 0 : 1;                     // pooling has no self term
    // lane l <= RPW: CSR pointer of row l and the virtual start of row l (CSR pointer + one self entry per earlier row)
    int rp = 0;
    if (mode != KAGNN_AGG_NONE) {
        long long r = row0 + rl0 + min(lane, RPW);
        if (r > p.num_rows) r = p.num_rows;
        rp = __ldg(a.rowptr + r);
    }
    const int vs = rp + self1 * min(lane, RPW);
    const int v_beg = __shfl_sync(0xffffffffu, vs, 0), v_end = __shfl_sync(0xffffffffu, vs, RPW);
    int cur = 0, cur_vend = __shfl_sync(0xffffffffu, vs, 1);
    float acc[4] = {0.f, 0.f, 0.f, 0.f};

    auto finish_row = [&]() {
        const long long r = row0 + rl0 + cur;
        const bool rv = r < p.num_rows;
        float os = 1.0f;
        if (mode == KAGNN_AGG_SEGMENT_MEAN)
            os = 1.0f / (float)max(__shfl_sync(0xffffffffu, rp, cur + 1) - __shfl_sync(0xffffffffu, rp, cur), 1);
        float o[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            float t = acc[i] * os;
            if (p.has_pre) {
                t = fmaf(t, sc[i], sh[i]);
                if (pre_silu) t = __fdividef(t, 1.0f + ex2_approx(-kLog2e * t));
            }
            o[i] = (cv[i] && rv) ? t : 0.f;
            acc[i] = 0.f;
        }
        float* d = xsu + (rl0 + cur) * xld + cl;
        if (VEC) {
            if (cin[0]) *reinterpret_cast<float4*>(d) = make_float4(o[0], o[1], o[2], o[3]);
        } else {
#pragma unroll
            for (int i = 0; i < 4; ++i)
                if (cin[i]) d[32 * i] = o[i];
        }
        if (p.agg_out && rv) {
            float* g = p.agg_out + r * p.ld_agg_out + cbase;
#pragma unroll
            for (int i = 0; i < 4; ++i)
                if (cv[i]) g[VEC ? i : 32 * i] = o[i];
        }
        ++cur;
        cur_vend = __shfl_sync(0xffffffffu, vs, min(cur + 1, RPW));
    };

#pragma unroll 1
    for (int vbase = v_beg; vbase < v_end; vbase += 32) {
        const int cnt = min(32, v_end - vbase);
        // ---- lane-parallel: virtual entry vbase+lane -> (source-row pointer, weight, self flag) ----------------------
        const int ve = vbase + lane;
        int r = 0;
#pragma unroll
        for (int l = 1; l <= RPW; ++l) r += (__shfl_sync(0xffffffffu, vs, l) <= ve) ? 1 : 0;
        r = min(r, RPW - 1);
        const int vs_r = __shfl_sync(0xffffffffu, vs, r), rp_r = __shfl_sync(0xffffffffu, rp, r);
        const bool is_self = (self1 != 0) && (ve == vs_r);
        const float* my_row = a.x;
        const float* my_erow = a.edge_feat;
        float my_w = 0.0f;
        int my_flag = 0;                                   // bit 0: load the row, bit 1: self entry
        if (lane < cnt) {
            if (is_self) {
                const long long rg = row0 + rl0 + r;
                if (rg < p.num_rows) {
                    const long long sr = a.src_index ? (long long)__ldg(a.src_index + rg) : rg;
                    my_row = a.x + sr * a.ldx;
                    my_w = a.self_scale;
                    if (mode == KAGNN_AGG_NONE) my_w = 1.0f;
                    if (mode == KAGNN_AGG_WEIGHTED && a.self_weight) my_w = __ldg(a.self_weight + rg);
                    my_flag = 3;
                }
            } else {
                const int e = rp_r + (ve - vs_r) - self1;
                int j = a.col ? __ldg(a.col + e) : e;
                if (a.src_index) j = __ldg(a.src_index + j);
                my_row = src_row(a, j);
                my_w = (mode == KAGNN_AGG_WEIGHTED) ? __ldg(a.edge_weight + e) : 1.0f;
                if (GINE) my_erow = a.edge_feat + (long long)__ldg(a.edge_row + e) * a.ld_edge;
                my_flag = 1;
            }
        }
#pragma unroll 1
        for (int t0 = 0; t0 < cnt; t0 += U) {
            float v[U][4], ev[GINE ? U : 1][4], w[U];
            int fl[U];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                fl[u] = __shfl_sync(0xffffffffu, my_flag, t0 + u);
                w[u] = __shfl_sync(0xffffffffu, my_w, t0 + u);
                const bool on = (fl[u] & 1) != 0;
                ldp4<VEC>(shfl_ptr(my_row, t0 + u) + cbase, on, v[u], cv);
                if (GINE) ldp4<VEC>(shfl_ptr(my_erow, t0 + u) + cbase, on && !(fl[u] & 2), ev[u], cv);
            }
            // consume: entries [done, lim) of the sub-batch belong to the current row; a row boundary inside it
            // finishes the row (possibly several empty ones) and continues -- all warp-uniform
            const int n_sub = min(U, cnt - t0), e0 = vbase + t0;
            int done = 0;
            for (;;) {
                const int lim = min(n_sub, cur_vend - e0);
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    if (u >= done && u < lim) {
#pragma unroll
                        for (int i = 0; i < 4; ++i) {
                            if (GINE) acc[i] += (fl[u] & 2) ? w[u] * v[u][i] : fmaxf(v[u][i] + ev[u][i], 0.f);
                            else acc[i] = fmaf(w[u], v[u][i], acc[i]);      // w = 1 for plain sums
                        }
                    }
                }
                if (lim >= n_sub) break;
                done = max(done, lim);
                finish_row();
            }
        }
    }
    while (cur < RPW) finish_row();
}


// A row of zeros in device memory: list entries that do not exist (tail of the last batch, rows past the end of the graph)
// point here, so every load of a sub-batch is unconditional and the U loads are all in flight together.
__device__ float g_zero_row[4096];

// Degree skew (R-MAT hubs): a destination row with more than HUB_T incoming entries would serialise on the one warp that owns
// it.  Such rows are left out of the owner's flattened list (only the self term is parked) and are then summed COOPERATIVELY:
// each of the NGW gather warps of the tile takes a contiguous 1/NGW slice of the row's CSR entries, the partial sums meet in
// this shared scratch, and the owner adds them in warp order (deterministic, no atomics).
constexpr int HUB_T = 512;
struct HubScratch {
    float part[NGW][128];
};
__device__ __forceinline__ void gather_warps_barrier() { asm volatile("bar.sync 2, %0;" ::"n"(NGW * 32) : "memory"); }

// Fast gather (128-bit path, no edge features): same flattened list as gather_unit, restructured for memory-level
// parallelism and a short instruction stream:
//   * every entry of a sub-batch is loaded unconditionally from a valid address (missing entries -> g_zero_row with
//     weight 0), so ptxas keeps the U 128-bit loads independent instead of chaining predicated load + move pairs;
//   * the column index / edge weight of the NEXT batch of 32 entries are fetched before the current batch is consumed;
//   * a finished row is parked raw (one STS.128); mean scale, pre-affine and the agg_out store run in a post-pass over
//     the warp's 16 rows only when the layer has any of them (a plain GIN layer has none).
template <bool WEIGHTED, bool GOUT = false, int U_ = KAGNN_TC2_GATHER_U>      // GOUT: rows are parked in GLOBAL memory (aggregation-only launch): bounds-checked
__device__ __forceinline__ void gather_unit_fast(const Tc2Params& p, long long row0, int c0, int ucols, float* __restrict__ xsu,
                                                 int gw, int lane, HubScratch* hub = nullptr) {
    constexpr int U = U_;
    const KagnnAggregate& a = p.agg;
    const int F = a.num_cols, mode = a.mode, xld = p.xld;
    const bool segment = (mode == KAGNN_AGG_SEGMENT_SUM) || (mode == KAGNN_AGG_SEGMENT_MEAN);
    const int rl0 = gw * RPW;
    const int cl = 4 * lane;
    const bool cin0 = cl < ucols;
    const bool cv0 = cin0 && (c0 + cl) < F;
    const int cload = cv0 ? (c0 + cl) : 0;                 // lanes past the last column re-read column 0 and discard it
    const int self1 = segment ? 0 : 1;
    int rp = 0;
    if (mode != KAGNN_AGG_NONE) {
        long long r = row0 + rl0 + min(lane, RPW);
        if (r > p.num_rows) r = p.num_rows;
        rp = __ldg(a.rowptr + r);
    }
    // hub rows of the whole 128-row tile (every gather warp computes the same four ballots, so no communication is needed to
    // agree on them); my own hub rows contribute only their self entry to the flattened list
    uint32_t tile_hub[4] = {0u, 0u, 0u, 0u};
    bool any_hub = false;
    if (hub != nullptr && !segment && mode != KAGNN_AGG_NONE) {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const long long r = row0 + 32 * k + lane;
            int d = 0;
            if (r < p.num_rows) d = __ldg(a.rowptr + r + 1) - __ldg(a.rowptr + r);
            tile_hub[k] = __ballot_sync(0xffffffffu, d > HUB_T);
            any_hub = any_hub || tile_hub[k] != 0u;
        }
    }
    const uint32_t my_hub = any_hub ? ((tile_hub[rl0 >> 5] >> (rl0 & 31)) & ((1u << RPW) - 1u)) : 0u;
    // virtual start of every row = exclusive prefix sum of the rows' list lengths (self entry + the CSR entries it keeps)
    int vs = rp + self1 * min(lane, RPW);                  // no hub in the tile: list positions follow the CSR directly
    if (any_hub) {
        const int rp_next = __shfl_down_sync(0xffffffffu, rp, 1);
        int len = 0;
        if (lane < RPW) len = self1 + (((my_hub >> lane) & 1u) ? 0 : (rp_next - rp));
        int incl = len;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int t = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += t;
        }
        vs = incl - len;                                   // lane RPW holds the total
    }
    const int v_beg = __shfl_sync(0xffffffffu, vs, 0), v_end = __shfl_sync(0xffffffffu, vs, RPW);
    int cur = 0, cur_vend = __shfl_sync(0xffffffffu, vs, 1);
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    float* const dst0 = xsu + rl0 * xld + cl;
    // pointer of source row j (owned / halo matrix, or the owner's memory over NVLink)
    auto entry_row = [&](int j) -> const float* {
        if (a.src_index) j = __ldg(a.src_index + j);
        if (a.peer_x) {
            const int owner = j / (int)a.rows_per_rank;
            return reinterpret_cast<const float*>(__ldg(reinterpret_cast<const unsigned long long*>(a.peer_x) + owner)) +
                   (long long)(j - owner * (int)a.rows_per_rank) * a.ldx;
        }
        return src_row(a, j);
    };

    auto finish_row = [&]() {
        if (GOUT ? (cv0 && row0 + rl0 + cur < p.num_rows) : cin0) {
            const float4 o = cv0 ? make_float4(acc[0], acc[1], acc[2], acc[3]) : make_float4(0.f, 0.f, 0.f, 0.f);
            *reinterpret_cast<float4*>(dst0 + (long long)cur * xld) = o;
        }
        acc[0] = acc[1] = acc[2] = acc[3] = 0.f;
        ++cur;
        cur_vend = (cur < RPW) ? __shfl_sync(0xffffffffu, vs, min(cur + 1, RPW)) : 0x7fffffff;
    };
    // lane-parallel description of list entry vbase + lane: row lookup, self flag, CSR entry id; the index / weight loads
    // are issued here and consumed one batch later
    struct Meta {
        int r, e, col;
        float w;
        bool self_, on;
    };
    auto prep = [&](int vbase) {
        Meta m;
        const int ve = vbase + lane;
        int r = 0;
#pragma unroll
        for (int step = RPW / 2; step >= 1; step >>= 1) {
            const int t = __shfl_sync(0xffffffffu, vs, r + step);
            r += (t <= ve) ? step : 0;
        }
        const int vs_r = __shfl_sync(0xffffffffu, vs, r), rp_r = __shfl_sync(0xffffffffu, rp, r);
        m.r = r;
        m.on = ve < v_end;
        m.self_ = (self1 != 0) && (ve == vs_r);
        m.e = rp_r + (ve - vs_r) - self1;
        m.col = m.e;
        m.w = 1.0f;
        if (m.on && !m.self_) {
            if (a.col) m.col = __ldg(a.col + m.e);
            if (WEIGHTED) m.w = __ldg(a.edge_weight + m.e);
        }
        return m;
    };

    Meta nxt = prep(v_beg);
#ifdef KAGNN_TRACE
    int trk = 0;
#endif
#pragma unroll 1
    for (int vbase = v_beg; vbase < v_end; vbase += 32) {
        const Meta m = nxt;
        if (vbase + 32 < v_end) nxt = prep(vbase + 32);
        const float* my_row = g_zero_row;
        float my_w = 0.0f;
        if (m.on) {
            if (m.self_) {
                const long long rg = row0 + rl0 + m.r;
                if (rg < p.num_rows) {
                    const long long sr = a.src_index ? (long long)__ldg(a.src_index + rg) : rg;
                    my_row = a.x + sr * a.ldx;
                    my_w = (mode == KAGNN_AGG_NONE) ? 1.0f : a.self_scale;
                    if (WEIGHTED && a.self_weight) my_w = __ldg(a.self_weight + rg);
                }
            } else {
                my_row = entry_row(m.col);
                my_w = m.w;
            }
        }
        const int cnt = min(32, v_end - vbase);
#pragma unroll 1
        for (int t0 = 0; t0 < cnt; t0 += U) {
#ifdef KAGNN_TRACE
            if (lane == 0 && gw == 0 && row0 >= 148 * BM && row0 < 149 * BM) { TRL(7, trk, 0); }
#endif
            float4 v[U];
#pragma unroll
            for (int u = 0; u < U; ++u) v[u] = __ldg(reinterpret_cast<const float4*>(shfl_ptr(my_row, t0 + u) + cload));
#ifdef KAGNN_TRACE
            if (lane == 0 && gw == 0 && row0 >= 148 * BM && row0 < 149 * BM) { TRL(7, trk, 1); }    // CTA 0's second tile
#endif
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int e = vbase + t0 + u;
                const float w = __shfl_sync(0xffffffffu, my_w, t0 + u);
                while (e >= cur_vend) finish_row();        // warp-uniform; also steps over empty rows
                acc[0] = fmaf(w, v[u].x, acc[0]);
                acc[1] = fmaf(w, v[u].y, acc[1]);
                acc[2] = fmaf(w, v[u].z, acc[2]);
                acc[3] = fmaf(w, v[u].w, acc[3]);
#ifdef KAGNN_TRACE
                if (u == 0 && lane == 0 && gw == 0 && row0 >= 148 * BM && row0 < 149 * BM) { TRL(7, trk, 2); }
#endif
            }
#ifdef KAGNN_TRACE
            if (lane == 0 && gw == 0 && row0 >= 148 * BM && row0 < 149 * BM) { TRL(7, trk, 3); ++trk; }
#endif
        }
    }
    while (cur < RPW) finish_row();

    if (any_hub) {
        // ---- cooperative pass over the tile's hub rows (uniform across the NGW gather warps) ------------------------------
#pragma unroll 1
        for (int k = 0; k < 4; ++k) {
            uint32_t bits = tile_hub[k];
#pragma unroll 1
            while (bits) {
                const int h = 32 * k + (__ffs(bits) - 1);           // row of the tile
                bits &= bits - 1;
                const int beg = __ldg(a.rowptr + row0 + h), end = __ldg(a.rowptr + row0 + h + 1);
                const int per = (end - beg + NGW - 1) / NGW;
                const int s_beg = min(end, beg + gw * per), s_end = min(end, s_beg + per);
                float ph[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll 1
                for (int eb = s_beg; eb < s_end; eb += 32) {
                    const int e = eb + lane;
                    const float* my_row = g_zero_row;
                    float my_w = 0.0f;
                    if (e < s_end) {
                        my_row = entry_row(a.col ? __ldg(a.col + e) : e);
                        my_w = WEIGHTED ? __ldg(a.edge_weight + e) : 1.0f;
                    }
                    const int cnt = min(32, s_end - eb);
#pragma unroll 1
                    for (int t0 = 0; t0 < cnt; t0 += U) {
                        float4 v[U];
#pragma unroll
                        for (int u = 0; u < U; ++u) v[u] = __ldg(reinterpret_cast<const float4*>(shfl_ptr(my_row, t0 + u) + cload));
#pragma unroll
                        for (int u = 0; u < U; ++u) {
                            const float w = __shfl_sync(0xffffffffu, my_w, t0 + u);
                            ph[0] = fmaf(w, v[u].x, ph[0]);
                            ph[1] = fmaf(w, v[u].y, ph[1]);
                            ph[2] = fmaf(w, v[u].z, ph[2]);
                            ph[3] = fmaf(w, v[u].w, ph[3]);
                        }
                    }
                }
                if (cin0) *reinterpret_cast<float4*>(&hub->part[gw][cl]) = make_float4(ph[0], ph[1], ph[2], ph[3]);
                gather_warps_barrier();
                if (h >= rl0 && h < rl0 + RPW && (GOUT ? (cv0 && row0 + h < p.num_rows) : cin0)) {
                    float* d = dst0 + (long long)(h - rl0) * xld;
                    float4 t = *reinterpret_cast<const float4*>(d);
#pragma unroll
                    for (int w8 = 0; w8 < NGW; ++w8) {
                        const float4 q = *reinterpret_cast<const float4*>(&hub->part[w8][cl]);
                        t.x += q.x; t.y += q.y; t.z += q.z; t.w += q.w;
                    }
                    if (!cv0) t = make_float4(0.f, 0.f, 0.f, 0.f);
                    *reinterpret_cast<float4*>(d) = t;
                }
                gather_warps_barrier();
            }
        }
    }

    const bool pre_silu = p.has_pre && p.pre.act == KAGNN_ACT_SILU;
    if (p.has_pre || (!GOUT && p.agg_out) || mode == KAGNN_AGG_SEGMENT_MEAN) {
        // post-pass: every lane revisits the values it parked itself (no synchronisation needed)
        float4 sc = make_float4(1.f, 1.f, 1.f, 1.f), sh = make_float4(0.f, 0.f, 0.f, 0.f);
        if (p.has_pre && cv0) {
            if (p.pre.scale) sc = __ldg(reinterpret_cast<const float4*>(p.pre.scale + c0 + cl));
            if (p.pre.shift) sh = __ldg(reinterpret_cast<const float4*>(p.pre.shift + c0 + cl));
        }
        const bool pre_vec = (!p.pre.scale || (reinterpret_cast<uintptr_t>(p.pre.scale) & 15u) == 0) &&
                             (!p.pre.shift || (reinterpret_cast<uintptr_t>(p.pre.shift) & 15u) == 0);
        if (p.has_pre && cv0 && !pre_vec) {
            const int c = c0 + cl;
            if (p.pre.scale) sc = make_float4(__ldg(p.pre.scale + c), __ldg(p.pre.scale + c + 1), __ldg(p.pre.scale + c + 2), __ldg(p.pre.scale + c + 3));
            if (p.pre.shift) sh = make_float4(__ldg(p.pre.shift + c), __ldg(p.pre.shift + c + 1), __ldg(p.pre.shift + c + 2), __ldg(p.pre.shift + c + 3));
        }
        const bool out_vec = p.agg_out && ((reinterpret_cast<uintptr_t>(p.agg_out) & 15u) == 0) && (p.ld_agg_out % 4 == 0);
#pragma unroll 1
        for (int rr = 0; rr < RPW; ++rr) {
            const long long r = row0 + rl0 + rr;
            const bool rv = r < p.num_rows;
            float os = 1.0f;
            if (mode == KAGNN_AGG_SEGMENT_MEAN)
                os = 1.0f / (float)max(__shfl_sync(0xffffffffu, rp, rr + 1) - __shfl_sync(0xffffffffu, rp, rr), 1);
            if (GOUT ? !(cv0 && rv) : !cin0) continue;
            float4 t = *reinterpret_cast<const float4*>(dst0 + (long long)rr * xld);
            float o[4] = {t.x * os, t.y * os, t.z * os, t.w * os};
            const float scv[4] = {sc.x, sc.y, sc.z, sc.w}, shv[4] = {sh.x, sh.y, sh.z, sh.w};
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                if (p.has_pre) {
                    o[i] = fmaf(o[i], scv[i], shv[i]);
                    if (pre_silu) o[i] = __fdividef(o[i], 1.0f + ex2_approx(-kLog2e * o[i]));
                }
                if (!(cv0 && rv)) o[i] = 0.f;
            }
            *reinterpret_cast<float4*>(dst0 + (long long)rr * xld) = make_float4(o[0], o[1], o[2], o[3]);
            if (!GOUT && p.agg_out && rv && cv0) {
                float* g = p.agg_out + r * p.ld_agg_out + c0 + cl;
                if (out_vec) *reinterpret_cast<float4*>(g) = make_float4(o[0], o[1], o[2], o[3]);
                else { g[0] = o[0]; g[1] = o[1]; g[2] = o[2]; g[3] = o[3]; }
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// Two rows per load (units of at most 64 columns).  gather_unit_fast gives every lane 4 columns of ONE source row, so a 64-column
// unit (hidden width 64: two of the three GIN layers of the arxiv-shaped model) keeps half the warp idle in every 128-bit load.
// Here the flattened list is a list of entry PAIRS: lanes 0..15 read the pair's first row, lanes 16..31 its second, the two halves
// accumulate separately and meet (one shuffle-add per register) when the destination row is finished.  A row's list
// [self, nbr_0, nbr_1, ...] is padded to an even length with a zero row, so a pair never straddles two destination rows and the
// bookkeeping of the flattened list carries over with "entry" read as "pair".  Metadata: lane L describes slot (L & 1) of pair
// (L >> 1) of the current batch of 16 pairs, so one pointer shuffle per lane fetches "my half's" row of pair j from lane
// 2 j + (lane >> 4).  Tiles with hub rows (cooperative pass) stay on gather_unit_fast.
// ---------------------------------------------------------------------------------------------------------------------
template <bool WEIGHTED, bool GOUT = false, int U_ = KAGNN_TC2_GATHER_U>
__device__ __forceinline__ void gather_unit_half(const Tc2Params& p, long long row0, int c0, int ucols, float* __restrict__ xsu, int gw,
                                                 int lane) {
    constexpr int U = U_;
    constexpr unsigned FULL = 0xffffffffu;
    const KagnnAggregate& a = p.agg;
    const int F = a.num_cols, mode = a.mode, xld = p.xld;
    const bool segment = (mode == KAGNN_AGG_SEGMENT_SUM) || (mode == KAGNN_AGG_SEGMENT_MEAN);
    const int rl0 = gw * RPW;
    const int half = lane >> 4, cl = 4 * (lane & 15);
    const bool cin0 = cl < ucols;
    const bool cv0 = cin0 && (c0 + cl) < F;
    const int cload = cv0 ? (c0 + cl) : 0;
    const bool storer = half == 0;                         // after the halves have met, lanes 0..15 hold (and park) the row
    const int self1 = segment ? 0 : 1;
    int rp = 0;
    if (mode != KAGNN_AGG_NONE) {
        long long r = row0 + rl0 + min(lane, RPW);
        if (r > p.num_rows) r = p.num_rows;
        rp = __ldg(a.rowptr + r);
    }
    // pairs per row and their exclusive prefix sum (lane RPW holds the total)
    const int rp_next = __shfl_down_sync(FULL, rp, 1);
    int len = 0;
    if (lane < RPW) len = (self1 + (rp_next - rp) + 1) >> 1;
    int incl = len;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(FULL, incl, o);
        if (lane >= o) incl += t;
    }
    const int vs = incl - len;
    const int v_end = __shfl_sync(FULL, vs, RPW);
    int cur = 0, cur_vend = __shfl_sync(FULL, vs, 1);
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    float* const dst0 = xsu + rl0 * xld + cl;
    auto entry_row = [&](int j) -> const float* {
        if (a.src_index) j = __ldg(a.src_index + j);
        if (a.peer_x) {
            const int owner = j / (int)a.rows_per_rank;
            return reinterpret_cast<const float*>(__ldg(reinterpret_cast<const unsigned long long*>(a.peer_x) + owner)) +
                   (long long)(j - owner * (int)a.rows_per_rank) * a.ldx;
        }
        return src_row(a, j);
    };
    auto finish_row = [&]() {
        acc[0] += __shfl_xor_sync(FULL, acc[0], 16);
        acc[1] += __shfl_xor_sync(FULL, acc[1], 16);
        acc[2] += __shfl_xor_sync(FULL, acc[2], 16);
        acc[3] += __shfl_xor_sync(FULL, acc[3], 16);
        if (storer && (GOUT ? (cv0 && row0 + rl0 + cur < p.num_rows) : cin0)) {
            const float4 o = cv0 ? make_float4(acc[0], acc[1], acc[2], acc[3]) : make_float4(0.f, 0.f, 0.f, 0.f);
            *reinterpret_cast<float4*>(dst0 + (long long)cur * xld) = o;
        }
        acc[0] = acc[1] = acc[2] = acc[3] = 0.f;
        ++cur;
        cur_vend = (cur < RPW) ? __shfl_sync(FULL, vs, min(cur + 1, RPW)) : 0x7fffffff;
    };
    // lane L <-> slot (L & 1) of pair vbase + (L >> 1): source-row pointer and weight of that list entry (zero row, weight 0 for
    // the pad slot of an odd list and past the end); index / weight loads are issued one batch ahead
    struct Meta {
        int r, k, col;
        float w;
        bool on, self_;
    };
    auto prep = [&](int vbase) {
        Meta m;
        const int ve = vbase + (lane >> 1);
        int r = 0;
#pragma unroll
        for (int step = RPW / 2; step >= 1; step >>= 1) {
            const int t = __shfl_sync(FULL, vs, r + step);
            r += (t <= ve) ? step : 0;
        }
        const int vs_r = __shfl_sync(FULL, vs, r), rp_r = __shfl_sync(FULL, rp, r), rp_n = __shfl_sync(FULL, rp, r + 1);
        m.r = r;
        m.k = 2 * (ve - vs_r) + (lane & 1);                // position in the row's list [self, nbr_0, ...]
        m.on = ve < v_end && m.k < self1 + (rp_n - rp_r);
        m.self_ = (self1 != 0) && m.k == 0;
        m.col = rp_r + m.k - self1;                        // CSR entry id (also the row id in the SEGMENT modes)
        m.w = 1.0f;
        if (m.on && !m.self_) {
            const int e = m.col;
            if (a.col) m.col = __ldg(a.col + e);
            if (WEIGHTED) m.w = __ldg(a.edge_weight + e);
        }
        return m;
    };

    Meta nxt = prep(0);
#pragma unroll 1
    for (int vbase = 0; vbase < v_end; vbase += 16) {
        const Meta m = nxt;
        if (vbase + 16 < v_end) nxt = prep(vbase + 16);
        const float* my_row = g_zero_row;
        float my_w = 0.0f;
        if (m.on) {
            if (m.self_) {
                const long long rg = row0 + rl0 + m.r;
                if (rg < p.num_rows) {
                    const long long sr = a.src_index ? (long long)__ldg(a.src_index + rg) : rg;
                    my_row = a.x + sr * a.ldx;
                    my_w = (mode == KAGNN_AGG_NONE) ? 1.0f : a.self_scale;
                    if (WEIGHTED && a.self_weight) my_w = __ldg(a.self_weight + rg);
                }
            } else {
                my_row = entry_row(m.col);
                my_w = m.w;
            }
        }
        const int cnt = min(16, v_end - vbase);
#pragma unroll 1
        for (int t0 = 0; t0 < cnt; t0 += U) {
            float4 v[U];
#pragma unroll
            for (int u = 0; u < U; ++u) v[u] = __ldg(reinterpret_cast<const float4*>(shfl_ptr(my_row, min(2 * (t0 + u), 30) + half) + cload));
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int e = vbase + t0 + u;
                float w = __shfl_sync(FULL, my_w, min(2 * (t0 + u), 30) + half);
                if (t0 + u >= cnt) w = 0.f;                // pairs past the end of the list (their rows were loaded from valid lanes)
                while (e >= cur_vend && cur < RPW) finish_row();
                acc[0] = fmaf(w, v[u].x, acc[0]);
                acc[1] = fmaf(w, v[u].y, acc[1]);
                acc[2] = fmaf(w, v[u].z, acc[2]);
                acc[3] = fmaf(w, v[u].w, acc[3]);
            }
        }
    }
    while (cur < RPW) finish_row();

    const bool pre_silu = p.has_pre && p.pre.act == KAGNN_ACT_SILU;
    if (storer && (p.has_pre || (!GOUT && p.agg_out) || mode == KAGNN_AGG_SEGMENT_MEAN)) {
        // post-pass: lanes 0..15 revisit the values they parked themselves (no synchronisation needed)
        float4 sc = make_float4(1.f, 1.f, 1.f, 1.f), sh = make_float4(0.f, 0.f, 0.f, 0.f);
        if (p.has_pre && cv0) {
            const int c = c0 + cl;
            if (p.pre.scale) sc = make_float4(__ldg(p.pre.scale + c), __ldg(p.pre.scale + c + 1), __ldg(p.pre.scale + c + 2), __ldg(p.pre.scale + c + 3));
            if (p.pre.shift) sh = make_float4(__ldg(p.pre.shift + c), __ldg(p.pre.shift + c + 1), __ldg(p.pre.shift + c + 2), __ldg(p.pre.shift + c + 3));
        }
        const bool out_vec = p.agg_out && ((reinterpret_cast<uintptr_t>(p.agg_out) & 15u) == 0) && (p.ld_agg_out % 4 == 0);
#pragma unroll 1
        for (int rr = 0; rr < RPW; ++rr) {
            const long long r = row0 + rl0 + rr;
            const bool rv = r < p.num_rows;
            float os = 1.0f;
            if (mode == KAGNN_AGG_SEGMENT_MEAN && rv) os = 1.0f / (float)max(__ldg(a.rowptr + r + 1) - __ldg(a.rowptr + r), 1);
            if (GOUT ? !(cv0 && rv) : !cin0) continue;
            const float4 t = *reinterpret_cast<const float4*>(dst0 + (long long)rr * xld);
            float o[4] = {t.x * os, t.y * os, t.z * os, t.w * os};
            const float scv[4] = {sc.x, sc.y, sc.z, sc.w}, shv[4] = {sh.x, sh.y, sh.z, sh.w};
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                if (p.has_pre) {
                    o[i] = fmaf(o[i], scv[i], shv[i]);
                    if (pre_silu) o[i] = __fdividef(o[i], 1.0f + ex2_approx(-kLog2e * o[i]));
                }
                if (!(cv0 && rv)) o[i] = 0.f;
            }
            *reinterpret_cast<float4*>(dst0 + (long long)rr * xld) = make_float4(o[0], o[1], o[2], o[3]);
            if (!GOUT && p.agg_out && rv && cv0) {
                float* g = p.agg_out + r * p.ld_agg_out + c0 + cl;
                if (out_vec) *reinterpret_cast<float4*>(g) = make_float4(o[0], o[1], o[2], o[3]);
                else { g[0] = o[0]; g[1] = o[1]; g[2] = o[2]; g[3] = o[3]; }
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// Asynchronous gather (the default for GIN / GCN layers whose width is a multiple of 64): neighbour rows travel
// global -> shared memory with cp.async (LDGSTS, 16 bytes per lane, L2 -> shared without a register or L1 stop-over) into a
// small ring of row slots, and are summed from there.  What this buys over the register gathers above:
//   * memory-level parallelism no longer costs registers or warps: every half-warp keeps AG_D row copies in flight all the time
//     (issue of entry t and summation of entry t - AG_D alternate), instead of batches of loads that drain before the next batch;
//   * a lane only ever reads the 16 bytes it copied itself, so the only synchronisation is cp.async.wait_group -- no barrier, no
//     fence, no shuffled 64-bit pointers;
//   * a unit is 64 columns wide, so ONE instruction moves two rows (one per half-warp): half h owns destination rows 8h .. 8h+7
//     of the warp's 16 and walks their entries as one flattened list [self_0, nbrs_0 ..., self_1, ...] (balanced inside the half).
// Per list entry: ~1 shuffle + address + LDGSTS to issue, 1 shuffle + LDS.128 + 4 FMA + row-end test to sum.
// ---------------------------------------------------------------------------------------------------------------------
#ifndef KAGNN_TC2_AG
#define KAGNN_TC2_AG 0
#endif
#ifndef KAGNN_TC2_AG_DEPTH
#define KAGNN_TC2_AG_DEPTH 8
#endif
constexpr int AG_D = KAGNN_TC2_AG_DEPTH;          // row copies in flight per half-warp (<= 16: two metadata batches are kept)
constexpr int AG_R = AG_D + 2;                    // ring slots per half-warp
constexpr int AG_ROW_BYTES = 256;                 // one row of a 64-column unit
constexpr int AG_WARP_BYTES = 2 * AG_R * AG_ROW_BYTES;
constexpr int AG_BYTES = NGW * AG_WARP_BYTES;
static_assert(AG_D >= 1 && AG_D <= 16, "AG_DEPTH");

__device__ __forceinline__ void cp_async16(uint32_t dst_smem, const void* src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst_smem), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ float4 lds128(uint32_t addr) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
    return v;
}

template <bool WEIGHTED>
__device__ __forceinline__ void gather_unit_ag(const Tc2Params& p, long long row0, int c0, float* __restrict__ xsu, int gw, int lane,
                                               uint32_t stage_warp) {
    constexpr unsigned FULL = 0xffffffffu;
    const KagnnAggregate& a = p.agg;
    const int mode = a.mode, xld = p.xld;
    const int half = lane >> 4, l16 = lane & 15;
    const int rh0 = gw * RPW + 8 * half;                   // first tile-local row of this half-warp
    const long long grow0 = row0 + rh0;
    const int nvalid = (int)max(0LL, min(8LL, p.num_rows - grow0));      // rows of this half that exist
    const int li = min(min(l16, 8), nvalid);
    int rp = 0;
    if (mode != KAGNN_AGG_NONE) rp = __ldg(a.rowptr + min(grow0 + li, p.num_rows));
    const int rp0 = __shfl_sync(FULL, rp, 0, 16);
    const int vs = (rp - rp0) + li;                        // list position where row l16 starts (one self entry per existing row)
    const int L = __shfl_sync(FULL, vs, 8, 16);            // length of this half's list
    const int T = max(L, __shfl_xor_sync(FULL, L, 16));    // steps of the warp (both halves run in lockstep)
    const float* const xcol = a.x + c0 + 4 * l16;          // column offset folded into the base pointers
    const uint32_t stage_half = stage_warp + (uint32_t)(half * AG_R * AG_ROW_BYTES + 16 * l16);

    // source row of list code j (bit 30 = self entry, always an owned row)
    auto src_ptr = [&](int code) -> const float* {
        int j = code & 0x3fffffff;
        if (code & 0x40000000) {
            if (a.src_index) j = __ldg(a.src_index + j);
            return xcol + (long long)j * a.ldx;
        }
        if (a.src_index) j = __ldg(a.src_index + j);
        if (a.peer_x) {
            const int owner = j / (int)a.rows_per_rank;
            return reinterpret_cast<const float*>(__ldg(reinterpret_cast<const unsigned long long*>(a.peer_x) + owner)) +
                   (long long)(j - owner * (int)a.rows_per_rank) * a.ldx + (c0 + 4 * l16);
        }
        if (a.x_halo != nullptr && j >= a.num_local_src) return a.x_halo + (long long)(j - a.num_local_src) * a.ld_halo + (c0 + 4 * l16);
        return xcol + (long long)j * a.ldx;
    };
    // lane-parallel description of list position pb + l16: code = source id | self << 30 | last-entry-of-its-row << 31
    struct Meta {
        int code;
        float w;
    };
    auto prep = [&](int pb) {
        Meta m;
        m.code = 0;
        m.w = 0.f;
        const int pos = pb + l16;
        int r = 0;
#pragma unroll
        for (int step = 4; step >= 1; step >>= 1) {
            const int t = __shfl_sync(FULL, vs, r + step, 16);
            r += (t <= pos) ? step : 0;
        }
        const int vs_r = __shfl_sync(FULL, vs, r, 16), vs_n = __shfl_sync(FULL, vs, r + 1, 16), rp_r = __shfl_sync(FULL, rp, r, 16);
        if (pos < L) {
            const int lastbit = (pos + 1 == vs_n) ? (int)0x80000000 : 0;
            if (pos == vs_r) {
                const long long rg = grow0 + r;
                m.code = (int)rg | 0x40000000 | lastbit;
                m.w = (mode == KAGNN_AGG_NONE) ? 1.0f : a.self_scale;
                if (mode == KAGNN_AGG_WEIGHTED && a.self_weight) m.w = __ldg(a.self_weight + rg);
            } else {
                const int e = rp_r + (pos - vs_r) - 1;
                m.code = (a.col ? __ldg(a.col + e) : e) | lastbit;
                m.w = (mode == KAGNN_AGG_WEIGHTED) ? __ldg(a.edge_weight + e) : 1.0f;
            }
        }
        return m;
    };

    Meta prv, cur, nxt = prep(0);
    prv.code = cur.code = 0;
    prv.w = cur.w = 0.f;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    float* drow = xsu + (long long)rh0 * xld + 4 * l16;    // destination of the row being summed
    uint32_t wslot = 0, rslot = 0;
#pragma unroll 1
    for (int t = 0; t < T + AG_D; ++t) {
        if ((t & 15) == 0) {
            prv = cur;
            cur = nxt;
            if (t + 16 < T) nxt = prep(t + 16);
        }
        // ---- issue: the copy of list entry t ------------------------------------------------------------------------------
        {
            const int code = __shfl_sync(FULL, cur.code, t & 15, 16);
            if (t < L) cp_async16(stage_half + wslot * AG_ROW_BYTES, src_ptr(code));
            cp_async_commit();
            wslot = (wslot + 1 == AG_R) ? 0u : wslot + 1;
        }
        // ---- sum: list entry u = t - AG_D, whose copy has landed once at most AG_D newer groups are pending --------------------
        const int u = t - AG_D;
        if (u >= 0) {
            cp_async_wait<AG_D>();
            const bool in_cur = (u >> 4) == (t >> 4);
            const int code = __shfl_sync(FULL, in_cur ? cur.code : prv.code, u & 15, 16);
            const float w = WEIGHTED ? __shfl_sync(FULL, in_cur ? cur.w : prv.w, u & 15, 16) : 1.0f;
            if (u < L) {
                const float4 v = lds128(stage_half + rslot * AG_ROW_BYTES);
                acc.x = fmaf(w, v.x, acc.x);
                acc.y = fmaf(w, v.y, acc.y);
                acc.z = fmaf(w, v.z, acc.z);
                acc.w = fmaf(w, v.w, acc.w);
                if (code < 0) {                            // last entry of its destination row: park the sum
                    *reinterpret_cast<float4*>(drow) = acc;
                    acc = make_float4(0.f, 0.f, 0.f, 0.f);
                    drow += xld;
                }
            }
            rslot = (rslot + 1 == AG_R) ? 0u : rslot + 1;
        }
    }
    // rows past the end of the graph (last tile): zeros for the basis producers
    for (int r = nvalid; r < 8; ++r) *reinterpret_cast<float4*>(xsu + (long long)(rh0 + r) * xld + 4 * l16) = make_float4(0.f, 0.f, 0.f, 0.f);

    if (p.has_pre || p.agg_out) {
        // post-pass over the values this lane parked itself: pre-affine (GCNConv bias / eval BatchNorm / SiLU) and the agg_out copy
        const bool pre_silu = p.has_pre && p.pre.act == KAGNN_ACT_SILU;
        const int c = c0 + 4 * l16;
        float4 sc = make_float4(1.f, 1.f, 1.f, 1.f), sh = make_float4(0.f, 0.f, 0.f, 0.f);
        if (p.has_pre) {
            if (p.pre.scale) sc = make_float4(__ldg(p.pre.scale + c), __ldg(p.pre.scale + c + 1), __ldg(p.pre.scale + c + 2), __ldg(p.pre.scale + c + 3));
            if (p.pre.shift) sh = make_float4(__ldg(p.pre.shift + c), __ldg(p.pre.shift + c + 1), __ldg(p.pre.shift + c + 2), __ldg(p.pre.shift + c + 3));
        }
        const bool out_vec = p.agg_out && ((reinterpret_cast<uintptr_t>(p.agg_out) & 15u) == 0) && (p.ld_agg_out % 4 == 0);
#pragma unroll 1
        for (int r = 0; r < nvalid; ++r) {
            float* d = xsu + (long long)(rh0 + r) * xld + 4 * l16;
            const float4 t = *reinterpret_cast<const float4*>(d);
            float o[4] = {t.x, t.y, t.z, t.w};
            if (p.has_pre) {
                o[0] = fmaf(o[0], sc.x, sh.x); o[1] = fmaf(o[1], sc.y, sh.y); o[2] = fmaf(o[2], sc.z, sh.z); o[3] = fmaf(o[3], sc.w, sh.w);
                if (pre_silu) {
#pragma unroll
                    for (int q = 0; q < 4; ++q) o[q] = __fdividef(o[q], 1.0f + ex2_approx(-kLog2e * o[q]));
                }
                *reinterpret_cast<float4*>(d) = make_float4(o[0], o[1], o[2], o[3]);
            }
            if (p.agg_out) {
                float* g = p.agg_out + (grow0 + r) * p.ld_agg_out + c;
                if (out_vec) *reinterpret_cast<float4*>(g) = make_float4(o[0], o[1], o[2], o[3]);
                else { g[0] = o[0]; g[1] = o[1]; g[2] = o[2]; g[3] = o[3]; }
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------------
template <int K, bool BF16, bool PUSH>
__global__ void __launch_bounds__(NTHREADS, 1) fused_tc2_kernel(const __grid_constant__ Tc2Params p) {
    extern __shared__ __align__(128) uint8_t smem[];
    float* xs = reinterpret_cast<float*>(smem);
    uint8_t* bst = smem + (size_t)p.n_units * p.unit_floats * sizeof(float);
    uint4* lut = reinterpret_cast<uint4*>(bst + (size_t)p.ns * p.bstage_bytes);
    uint64_t* bars = reinterpret_cast<uint64_t*>(lut + KAGNN_MAX_LAYERS * LUT_ROWS);
    uint64_t* xs_full = bars;
    uint64_t* xs_empty = bars + MAX_UNITS;
    uint64_t* full = bars + 2 * MAX_UNITS;        // A stage written by all producers AND W chunk landed (one wait for the MMA)
    uint64_t* empty = full + MAX_STAGE;
    uint64_t* acc_full = empty + MAX_STAGE;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_full + 2);       // acc_full[region]: one barrier per accumulator region
    float* post_sc = reinterpret_cast<float*>(tmem_slot + 4);     // post-affine of the last layer (<= 256 columns each; 16-byte aligned)
    float* post_sh = post_sc + 256;
    volatile int* gather_progress = reinterpret_cast<volatile int*>(post_sh + 256);   // tiles the gather warps have started
    HubScratch* hub_scratch = reinterpret_cast<HubScratch*>(post_sh + 256 + 4);
    float2* ln_part = reinterpret_cast<float2*>(hub_scratch + 1);       // [2][NWG][128]: FastKAN LayerNorm partial sums
    uint8_t* ag_stage = reinterpret_cast<uint8_t*>(ln_part + 2 * NWG * 128);   // [NGW][2][AG_R][256 B]: row slots of the asynchronous gather
    // pushed output (KagnnAggregate.push_y): [mbarrier 0 | tiles stored so far | mbarrier 1 | pad to 128 B | relay_bytes of row slots]
    // (addresses are derived where they are used: keeping them live across the role dispatch costs every role registers)
    auto relay_block = [&]() -> uint8_t* {
        return reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(ag_stage + ((KAGNN_TC2_AG && p.ag) ? AG_BYTES : 0)) + 127u) & ~(uintptr_t)127u);
    };
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

    if (warp == WARP_MMA) tc::tmem_alloc(tmem_slot, 512);
    if (tid == WARP_LOAD * 32) {
        for (int s = 0; s < MAX_UNITS; ++s) {
            tc::mbar_init(&xs_full[s], NGW);
            tc::mbar_init(&xs_empty[s], NPW);
        }
        for (int s = 0; s < MAX_STAGE; ++s) {
            tc::mbar_init(&full[s], FULL_ARRIVALS);
            tc::mbar_init(&empty[s], 1);
        }
        tc::mbar_init(&acc_full[0], 1);
        tc::mbar_init(&acc_full[1], 1);
        if (PUSH) {
            tc::mbar_init(reinterpret_cast<uint64_t*>(relay_block()), 1);
            tc::mbar_init(reinterpret_cast<uint64_t*>(relay_block() + 16), 1);
            *reinterpret_cast<uint32_t*>(relay_block() + 8) = 0u;
        }
        tc::mbar_fence_init();
    }
    if (tid == 0) *gather_progress = 0;
    if (tid >= 128 && tid < 384) {
        const int c = tid - 128;
        const LayerT2& LL = p.layers[p.n_layers - 1];
        const bool on = p.has_post && c < LL.N;
        const float sc = (on && p.post.scale) ? __ldg(p.post.scale + c) : 1.0f;
        const float sh = (on && p.post.shift) ? __ldg(p.post.shift + c) : 0.0f;
        const float bb = (K == 0 && LL.bias && c < LL.N) ? __ldg(LL.bias + c) : 0.0f;     // FastKAN base_linear.bias
        post_sc[c] = sc;
        post_sh[c] = fmaf(bb, sc, sh);
    }
    if (tid < p.n_layers * LUT_ROWS) {
        // per layer, row q <-> knot interval j = q - 1.  A valid interval (0 <= j < G + 2k) puts its first non-zero basis
        // into slot s = j - K: output byte b of the 16-byte slot vector takes source byte b - 2s of (b0 b1 | b2 b3) when that
        // is in 0..7, else the replicated (zero) sign bit of source byte 1.  Rows of intervals outside the knot range select
        // nothing (every slot zero).
        const int l = tid / LUT_ROWS, j = tid - l * LUT_ROWS - 1;
        const bool inside = j >= 0 && j < (int)p.layers[l].lim;
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
        lut[tid] = make_uint4(w[0], w[1], w[2], w[3]);
    }
    tc::tc_fence_before_sync();
    __syncthreads();
    tc::tc_fence_after_sync();
    const uint32_t tmem_base = *tmem_slot;

    if (warp < NPW) {
        // register budget (launch: 72 per thread): see REGS_PROD / REGS_GATHER / REGS_AUX
        if (REGS_PROD > 72) tc::reg_inc<REGS_PROD>();
        // ============================== BASIS PRODUCERS / EPILOGUE =============================================
        // Every warpgroup works on EVERY chunk: warpgroup wg expands FPW features of a spline chunk (every NWG-th octet of a
        // SiLU chunk), so the parts of a chunk are produced concurrently and the groups stay balanced.
        const int wg = warp >> 2;
        const int row = tid & 127;
        const uint32_t lane_base = (uint32_t)((warp & 3) * 32) << 16;
        uint32_t cq = 0, lc = 0;                           // running chunk / layer counters
        int u_slot = 0;                                    // x-tile ring: slot and phase parity of the unit being consumed
        uint32_t u_par = 0;
        // ---- epilogue of a tile's last layer: TMEM -> registers -> post-affine -> y.  Software-pipelined: it runs AFTER the
        // basis production of the NEXT tile's first layer (which writes the other accumulator region), so the wait for the
        // last MMAs, the TMEM reads and the stores overlap the tensor-core work already queued for the next tile.
        auto epilogue = [&](long long e_row0, uint32_t e_lc) {
            const LayerT2& L = p.layers[p.n_layers - 1];
            const int nrows = (int)min((long long)BM, p.num_rows - e_row0);
            if (lane == 0 && (warp & 3) == 0 && wg < 2) TRL(4 + wg, e_lc, 2);
            tc::mbar_wait(&acc_full[(e_lc - 1) & 1], ((e_lc - 1) >> 1) & 1);
            if (lane == 0 && (warp & 3) == 0 && wg < 2) TRL(4 + wg, e_lc, 3);
            tc::tc_fence_after_sync();
            const uint32_t taddr = tmem_base + lane_base + ((((e_lc - 1) & 1) && !p.wide) ? 128u : 0u);
            float* yrow = p.y + (e_row0 + row) * p.ldy;
            for (int jb = wg; jb < L.N_pad / 8; jb += NWG) {
                float v[8];
                tc::tmem_ld8(taddr + (uint32_t)(8 * jb), v);
                if (L.stack) {
                    float v2[8];
                    tc::tmem_ld8(taddr + (uint32_t)(L.N_pad + 8 * jb), v2);
#pragma unroll
                    for (int i = 0; i < 8; ++i) v[i] += v2[i];
                }
                if (row < nrows) {
                    if (p.has_post || (K == 0 && L.bias)) {
                        const float4 s0 = *reinterpret_cast<const float4*>(post_sc + 8 * jb), s1 = *reinterpret_cast<const float4*>(post_sc + 8 * jb + 4);
                        const float4 h0 = *reinterpret_cast<const float4*>(post_sh + 8 * jb), h1 = *reinterpret_cast<const float4*>(post_sh + 8 * jb + 4);
                        v[0] = fmaf(v[0], s0.x, h0.x); v[1] = fmaf(v[1], s0.y, h0.y); v[2] = fmaf(v[2], s0.z, h0.z); v[3] = fmaf(v[3], s0.w, h0.w);
                        v[4] = fmaf(v[4], s1.x, h1.x); v[5] = fmaf(v[5], s1.y, h1.y); v[6] = fmaf(v[6], s1.z, h1.z); v[7] = fmaf(v[7], s1.w, h1.w);
                        if (p.post.act == KAGNN_ACT_SILU) {
#pragma unroll
                            for (int i = 0; i < 8; ++i) v[i] = silu_f(v[i]);
                        }
                    }
                    if (kSplitK && p.n_split > 1) {         // split-K: this item's partial sums join the others' in y (zeroed by the launcher)
#pragma unroll
                        for (int i = 0; i < 8; ++i)
                            if (8 * jb + i < L.N) atomicAdd(yrow + 8 * jb + i, v[i]);
                    } else if (p.y_vec && 8 * jb + 8 <= L.N) {
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
            if (PUSH) tc::fence_proxy_async_global();   // the push warp reads these rows back with bulk copies
            if (lane == 0 && (warp & 3) == 0 && wg < 2) TRL(4 + wg, e_lc, 4);
        };
        // After an epilogue.  (1) The accumulator region just read is the one this tile's NEXT layer (or, for one-layer chains,
        // the next tile) overwrites: with every warpgroup on every chunk that MMA cannot start before all warps have left the
        // epilogue; with teams (CPR > 1) a team could hand over its chunk while the other team still reads -> barrier.
        // (2) Pushed output: every row of the tile is stored -> publish the count to the push warp.
        auto tile_stored = [&]() {
            if (CPR > 1 || PUSH) asm volatile("bar.sync 4, %0;" ::"n"(NPW * 32) : "memory");
            if (PUSH && tid == 0) {                     // the only writer: the count lives in shared memory, not in a register
                const uint32_t a = tc::smem_u32(relay_block() + 8);
                uint32_t v;
                asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a) : "memory");
                asm volatile("st.release.cta.shared.u32 [%0], %1;" ::"r"(a), "r"(v + 1u) : "memory");
            }
        };
        long long pend_row0 = 0;
        uint32_t pend_lc = 0;
        bool have_pend = false;
        for (int item = blockIdx.x; item < p.n_items; item += gridDim.x) {
            const WorkItem wi = work_item(p, item);
            const long long row0 = (long long)wi.tile * BM;
            for (int l = 0; l < p.n_layers; ++l, ++lc) {
                const LayerT2& L = p.layers[l];
                const float inv_h = L.inv_h, c0f = L.c0, limp = L.lim + 0.5f;
                const uint4* lutL = lut + l * LUT_ROWS;
                const int n_chunks = l == 0 ? wi.n_chunks : L.n_chunks;
                uint32_t src_t = 0, src_lo = 0;               // src_lo != 0: previous layer's accumulator is stacked
                int cur_unit = -1;
                const float* xrow = nullptr;
                if (l > 0) {
                    if (lane == 0 && (warp & 3) == 0 && wg < 2) TRL(4 + wg, lc, 0);
                    tc::mbar_wait(&acc_full[(lc - 1) & 1], ((lc - 1) >> 1) & 1);
                    if (lane == 0 && (warp & 3) == 0 && wg < 2) TRL(4 + wg, lc, 1);
                    tc::tc_fence_after_sync();
                    src_t = tmem_base + lane_base + (((lc - 1) & 1) ? 128u : 0u);
                    src_lo = p.layers[l - 1].stack ? (uint32_t)p.layers[l - 1].N_pad : 0u;
                }
                int s = (int)(cq % (uint32_t)p.ns);
                uint32_t par = ((cq / (uint32_t)p.ns) & 1u) ^ 1u;
                // ---- FastKAN (K == 0): previous layer's base bias rides on the values read back from TMEM; LayerNorm statistics
                // of this thread's row (fastkan.py:66,78: biased variance, eps 1e-5), shifted one-pass sums split over the warpgroups
                const float* pbias = (K == 0 && l > 0) ? p.layers[l - 1].bias : nullptr;
                const float rc0 = L.rc0, rstep = L.rstep, rk = L.rk;
                float ln_mean = 0.f, ln_rstd = 1.f;
                if (K == 0 && L.lnw && l == 0 && L.ln_stats) {
                    // statistics from the pre-pass (rows wider than one x unit)
                    const long long r = min(row0 + row, p.num_rows - 1);
                    const float2 st = __ldg(reinterpret_cast<const float2*>(L.ln_stats) + r);
                    ln_mean = st.x;
                    ln_rstd = st.y;
                } else if (K == 0) {
                    if (l == 0 && L.lnw) {                            // the launcher guarantees one x unit per tile in this case
                        cur_unit = 0;
                        xrow = xs + (size_t)u_slot * p.unit_floats + (size_t)row * p.xld;
                        tc::mbar_wait_relaxed(&xs_full[u_slot], u_par);
                    }
                    if (L.lnw) {
                        auto load8 = [&](int f0, float (&v)[8]) {
                            if (l == 0) {
                                const float4 t0 = *reinterpret_cast<const float4*>(xrow + f0);
                                const float4 t1 = *reinterpret_cast<const float4*>(xrow + f0 + 4);
                                v[0] = t0.x; v[1] = t0.y; v[2] = t0.z; v[3] = t0.w; v[4] = t1.x; v[5] = t1.y; v[6] = t1.z; v[7] = t1.w;
                            } else {
                                tc::tmem_ld8(src_t + (uint32_t)f0, v);
                                if (src_lo) {
                                    float v2[8];
                                    tc::tmem_ld8(src_t + src_lo + (uint32_t)f0, v2);
#pragma unroll
                                    for (int i = 0; i < 8; ++i) v[i] += v2[i];
                                }
                                if (pbias) {
#pragma unroll
                                    for (int i = 0; i < 8; ++i) v[i] += __ldg(pbias + min(f0 + i, L.F - 1));
                                }
                            }
                        };
                        float v[8];
                        load8(0, v);
                        const float x0 = v[0];
                        float sd = 0.f, sq = 0.f;
                        for (int f0 = 8 * wg; f0 < L.F_pad; f0 += 8 * NWG) {
                            load8(f0, v);
#pragma unroll
                            for (int i = 0; i < 8; ++i) {
                                const float d = (f0 + i < L.F) ? v[i] - x0 : 0.f;
                                sd += d;
                                sq = fmaf(d, d, sq);
                            }
                        }
                        float2* part = ln_part + (size_t)(lc & 1u) * NWG * 128;
                        part[wg * 128 + row] = make_float2(sd, sq);
                        asm volatile("bar.sync 3, %0;" ::"n"(NPW * 32) : "memory");
                        sd = 0.f;
                        sq = 0.f;
#pragma unroll
                        for (int w = 0; w < NWG; ++w) {
                            const float2 t = part[w * 128 + row];
                            sd += t.x;
                            sq += t.y;
                        }
                        const float inv_f = 1.0f / (float)L.F, md = sd * inv_f;
                        ln_mean = x0 + md;
                        ln_rstd = rsqrtf(fmaxf(fmaf(sq, inv_f, -md * md), 0.f) + 1e-5f);
                    }
                }
                ChunkCursor c(L.F_pad, l == 0 ? wi.g0 : 0);
                for (int q = 0; q < n_chunks; ++q, c.next()) {
                    if (l == 0 && c.j == 0) {
                        // x-tile ring: entering a new unit releases the previous one and waits for the gather warps
                        const int ul = (64 * c.group) >> p.uw_shift;
                        if (ul != cur_unit) {
                            if (cur_unit >= 0) {
                                __syncwarp();
                                if (lane == 0) tc::mbar_arrive(&xs_empty[u_slot]);
                                if (++u_slot == p.n_units) { u_slot = 0; u_par ^= 1u; }
                            }
                            cur_unit = ul;
                            if (lane == 0 && (warp & 3) == 0 && wg < 2) TRL(wg, cq + q, 0);
                            tc::mbar_wait_relaxed(&xs_full[u_slot], u_par);
                            if (lane == 0 && (warp & 3) == 0 && wg < 2) TRL(wg, cq + q, 1);
                            xrow = xs + (size_t)u_slot * p.unit_floats + (size_t)row * p.xld - (size_t)ul * p.uw;
                        }
                    }
                    if (CPR == 1 || ((cq + (uint32_t)q) % CPR) == (uint32_t)(wg / WGT)) {
                        const int wsub = wg % WGT;                            // this warpgroup's share inside its team
                        if (lane == 0 && (warp & 3) == 0 && wg < 2) TRC(wg, cq + q, 2);
                        tc::mbar_wait_relaxed(&empty[s], par);
                        if (lane == 0 && (warp & 3) == 0 && wg < 2) TRC(wg, cq + q, 3);
                        tc::tc_fence_after_sync();
                        const uint32_t a_t = tmem_base + lane_base + TMEM_A0 + 64u * s;
                        if (KAGNN_TC2_NOMATH) {
                        } else if (!c.base()) {
                            const int fsh = FPW * wsub;                       // first feature of this warpgroup inside the chunk
                            const int f0 = 64 * c.group + 8 * c.j + fsh;
                            float v[FPW];
                            if (l == 0) {
                                if (FPW == 8) {
                                    const float4 t0 = *reinterpret_cast<const float4*>(xrow + f0);
                                    const float4 t1 = *reinterpret_cast<const float4*>(xrow + f0 + 4);
                                    v[0] = t0.x; v[1] = t0.y; v[2 % FPW] = t0.z; v[3 % FPW] = t0.w;
                                    v[4 % FPW] = t1.x; v[5 % FPW] = t1.y; v[6 % FPW] = t1.z; v[7 % FPW] = t1.w;
                                } else if (FPW == 4) {
                                    const float4 t = *reinterpret_cast<const float4*>(xrow + f0);
                                    v[0] = t.x; v[1] = t.y; v[2 % FPW] = t.z; v[3 % FPW] = t.w;
                                } else {
                                    const float2 t = *reinterpret_cast<const float2*>(xrow + f0);
                                    v[0] = t.x; v[1] = t.y;
                                }
                            } else {
                                tc::tmem_ldn<FPW>(src_t + (uint32_t)f0, v);
                                if (src_lo) {
                                    float v2[FPW];
                                    tc::tmem_ldn<FPW>(src_t + src_lo + (uint32_t)f0, v2);
#pragma unroll
                                    for (int i = 0; i < FPW; ++i) v[i] += v2[i];
                                }
                                if (pbias) {
#pragma unroll
                                    for (int i = 0; i < FPW; ++i) v[i] += __ldg(pbias + min(f0 + i, L.F - 1));
                                }
                            }
                            if (K == 0 && L.lnw) {
#pragma unroll
                                for (int i = 0; i < FPW; ++i) {
                                    const int f = min(f0 + i, L.F - 1);
                                    v[i] = fmaf((v[i] - ln_mean) * ln_rstd, __ldg(L.lnw + f), L.lnb ? __ldg(L.lnb + f) : 0.f);
                                }
                            }
#pragma unroll
                            for (int i = 0; i < FPW; i += 2) {
                                uint32_t hi[8], lo[8];
                                if (K == 0) {
                                    rbf_slots<BF16>(rc0, rstep, rk, v[i], hi, lo);
                                    rbf_slots<BF16>(rc0, rstep, rk, v[i + 1], hi + 4, lo + 4);
                                } else if (KAGNN_TC2_X2 || BF16) {
                                    bspline_slots2<(K == 0 ? 1 : K), BF16>(inv_h, c0f, limp, lutL, v[i], v[i + 1], hi, lo);
                                } else {
                                    bspline_slots<(K == 0 ? 1 : K)>(inv_h, c0f, limp, lutL, v[i], hi, lo);
                                    bspline_slots<(K == 0 ? 1 : K)>(inv_h, c0f, limp, lutL, v[i + 1], hi + 4, lo + 4);
                                }
                                tc::tmem_st8(a_t + 4u * (uint32_t)(fsh + i), hi);
                                if (!BF16) tc::tmem_st8(a_t + 32u + 4u * (uint32_t)(fsh + i), lo);
                            }
                        } else {
#pragma unroll 1
                            for (int jj = wsub; jj < c.n_oct; jj += WGT) {
                                const int f0 = 64 * c.group + 8 * jj;
                                float v[8];
                                if (l == 0) {
                                    const float4 t0 = *reinterpret_cast<const float4*>(xrow + f0);
                                    const float4 t1 = *reinterpret_cast<const float4*>(xrow + f0 + 4);
                                    v[0] = t0.x; v[1] = t0.y; v[2] = t0.z; v[3] = t0.w; v[4] = t1.x; v[5] = t1.y; v[6] = t1.z; v[7] = t1.w;
                                } else {
                                    tc::tmem_ld8(src_t + (uint32_t)f0, v);
                                    if (src_lo) {
                                        float v2[8];
                                        tc::tmem_ld8(src_t + src_lo + (uint32_t)f0, v2);
#pragma unroll
                                        for (int i = 0; i < 8; ++i) v[i] += v2[i];
                                    }
                                    if (pbias) {
#pragma unroll
                                        for (int i = 0; i < 8; ++i) v[i] += __ldg(pbias + min(f0 + i, L.F - 1));
                                    }
                                }
                                float r[8];
#pragma unroll
                                for (int i = 0; i < 8; ++i) {
                                    v[i] = (K == 0) ? __fdividef(v[i], 1.0f + ex2_approx(-kLog2e * v[i])) : silu_nan(v[i]);
                                    r[i] = trunc_residual(v[i]);
                                }
                                if (BF16) {
                                    tc::tmem_st4(a_t + 4u * jj, pack_rn(v[0], v[1]), pack_rn(v[2], v[3]), pack_rn(v[4], v[5]), pack_rn(v[6], v[7]));
                                } else {
                                    tc::tmem_st4(a_t + 4u * jj, pack_trunc(v[0], v[1]), pack_trunc(v[2], v[3]), pack_trunc(v[4], v[5]),
                                                 pack_trunc(v[6], v[7]));
                                    tc::tmem_st4(a_t + 32u + 4u * jj, pack_rn(r[0], r[1]), pack_rn(r[2], r[3]), pack_rn(r[4], r[5]),
                                                 pack_rn(r[6], r[7]));
                                }
                            }
                        }
                        tc::tmem_st_wait();
                        tc::tc_fence_before_sync();
#if KAGNN_TC2_ELECT_ARRIVE
                        __syncwarp();
                        if (lane == 0) tc::mbar_arrive(&full[s]);
#else
                        tc::mbar_arrive(&full[s]);
#endif
                        if (lane == 0 && (warp & 3) == 0 && wg < 2) TRC(wg, cq + q, 4);
                    }
                    if (++s == p.ns) { s = 0; par ^= 1u; }
                }
                cq += (uint32_t)n_chunks;
                if (l == 0) {
                    __syncwarp();
                    if (lane == 0) tc::mbar_arrive(&xs_empty[u_slot]);
                    if (++u_slot == p.n_units) { u_slot = 0; u_par ^= 1u; }
                    if (have_pend) {                       // previous tile's epilogue, behind this tile's first layer
                        epilogue(pend_row0, pend_lc);
                        have_pend = false;
                        tile_stored();
                    }
                }
            }
            pend_row0 = row0;
            pend_lc = lc;
            have_pend = true;
            if (!KAGNN_TC2_PIPE_EPI || p.wide) {
                // one accumulator region only (wide layers): read it out before the next tile's first chunk can be handed over
                epilogue(pend_row0, pend_lc);
                have_pend = false;
                tile_stored();
            }
        }
        if (have_pend) {
            epilogue(pend_row0, pend_lc);
            if (PUSH) tile_stored();
        }
    } else if (warp < NPW + NGW) {
        // ========================================= GATHER ======================================================
        if (REGS_GATHER > 72) tc::reg_inc<REGS_GATHER>();
        const int gw = warp - NPW;
        const bool vec = (p.agg.num_cols % 4 == 0) && ((reinterpret_cast<uintptr_t>(p.agg.x) & 15u) == 0) && (p.agg.ldx % 4 == 0) &&
                         (!p.agg.x_halo || (((reinterpret_cast<uintptr_t>(p.agg.x_halo) & 15u) == 0) && (p.agg.ld_halo % 4 == 0))) &&
                         (p.agg.mode != KAGNN_AGG_GINE ||
                          (((reinterpret_cast<uintptr_t>(p.agg.edge_feat) & 15u) == 0) && (p.agg.ld_edge % 4 == 0)));
        const bool gine = p.agg.mode == KAGNN_AGG_GINE;
        const bool head_ok = !p.agg.num_head_cols || (((reinterpret_cast<uintptr_t>(p.agg.x_head) & 15u) == 0) && (p.agg.ld_head % 4 == 0));
        const bool plain_copy = p.agg.mode == KAGNN_AGG_NONE && !p.has_pre && !p.agg_out && !p.agg.src_index;
        const int F_pad = p.layers[0].F_pad;
        uint32_t uc = 0;
        int it = 0;
        int halo_done = 0;                                            // leading chunks of 256 halo rows known to have landed
        for (int item = blockIdx.x; item < p.n_items; item += gridDim.x, ++it) {
            const WorkItem wi = work_item(p, item);
            const int tile = wi.tile;
            const long long row0 = (long long)tile * BM;
            if (gw == 0 && lane == 0) *gather_progress = it;          // paces the L2 prefetch warp
            if (p.agg.halo_flags) {
                // halo rows are being pulled over NVLink while this kernel runs (kagnn_gather_rows_peer_ordered): wait until
                // the prefix that tiles 0..tile reference has landed
                const int nchunk = (__ldg(p.agg.halo_need + tile) + 255) >> 8;
                uint32_t tries = 0;
                while (halo_done < nchunk) {
                    int v;
                    asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p.agg.halo_flags + halo_done) : "memory");
                    if (v >= KAGNN_PULL_WARPS * p.agg.halo_epoch) {   // every warp of the pull block has stored its rows of this chunk
                        ++halo_done;
                    } else {
                        __nanosleep(200);
                        if (++tries > (1u << 24)) __trap();          // the pull never came: fail the launch instead of hanging
                    }
                }
            }
            for (int ub = wi.ub0; ub < wi.ub1; ++ub, ++uc) {
                const int u = (int)(uc % (uint32_t)p.n_units);
                if (lane == 0 && gw == 0) TRL(6, uc, 0);
                tc::mbar_wait_relaxed(&xs_empty[u], ((uc / (uint32_t)p.n_units) & 1u) ^ 1u);
                if (lane == 0 && gw == 0) TRL(6, uc, 1);
                float* xsu = xs + (size_t)u * p.unit_floats;
                const int c0 = ub * p.uw, ucols = min(p.uw, F_pad - c0);
                // does the tile hold a hub row (more than HUB_T entries)?  Those are summed cooperatively by gather_unit_fast; the
                // decision is tile-uniform (every gather warp evaluates the same four ballots)
                const bool csr_mode = p.agg.mode == KAGNN_AGG_GIN || p.agg.mode == KAGNN_AGG_WEIGHTED;
                const bool half_ok = KAGNN_TC2_HALF && vec && head_ok && !gine && !plain_copy && ucols <= 64;
                bool tile_has_hub = false;
                if (KAGNN_TC2_HUB && csr_mode && ((KAGNN_TC2_AG && p.ag) || half_ok)) {
#pragma unroll
                    for (int k4 = 0; k4 < 4; ++k4) {
                        const long long rr = row0 + 32 * k4 + lane;
                        int d = 0;
                        if (rr < p.num_rows) d = __ldg(p.agg.rowptr + rr + 1) - __ldg(p.agg.rowptr + rr);
                        if (__any_sync(0xffffffffu, d > HUB_T)) tile_has_hub = true;
                    }
                }
                bool gpr = false;
                if (KAGNN_TC2_AG && p.ag && !tile_has_hub) {
                    // asynchronous (cp.async ring) gather: 64-column units of a GIN / GCN layer (compiled out by default)
                    gpr = true;
                    const uint32_t st = tc::smem_u32(ag_stage) + (uint32_t)(gw * AG_WARP_BYTES);
                    if (p.agg.mode == KAGNN_AGG_WEIGHTED || p.agg.self_scale != 1.0f) gather_unit_ag<true>(p, row0, c0, xsu, gw, lane, st);
                    else gather_unit_ag<false>(p, row0, c0, xsu, gw, lane, st);
                } else if (half_ok && !tile_has_hub) {
                    // units of at most 64 columns: two source rows per 128-bit load instruction
                    gpr = true;
                    if (p.agg.mode == KAGNN_AGG_WEIGHTED) gather_unit_half<true>(p, row0, c0, ucols, xsu, gw, lane);
                    else gather_unit_half<false>(p, row0, c0, ucols, xsu, gw, lane);
                }
                if (gpr) {
                } else if (vec && head_ok) {
                    if (gine) gather_unit<true, true>(p, row0, c0, ucols, xsu, gw, lane);
                    else if (plain_copy) gather_unit<true, false>(p, row0, c0, ucols, xsu, gw, lane);
                    else if (p.agg.mode == KAGNN_AGG_WEIGHTED) gather_unit_fast<true>(p, row0, c0, ucols, xsu, gw, lane, KAGNN_TC2_HUB ? hub_scratch : nullptr);
                    else gather_unit_fast<false>(p, row0, c0, ucols, xsu, gw, lane, KAGNN_TC2_HUB ? hub_scratch : nullptr);
                } else {
                    if (gine) gather_unit<false, true>(p, row0, c0, ucols, xsu, gw, lane);
                    else gather_unit<false, false>(p, row0, c0, ucols, xsu, gw, lane);
                }
                __syncwarp();
                if (lane == 0) tc::mbar_arrive(&xs_full[u]);
                if (lane == 0 && gw == 0) TRL(6, uc, 2);
            }
        }
    } else if (warp == WARP_MMA) {
        tc::reg_dec<REGS_AUX>();
        // ========================================= MMA ISSUER ==================================================
        // The whole warp walks the loops (uniform control flow, descriptor arithmetic on the uniform datapath); one elected
        // lane issues.  Per K = 16 step: stacked layers (N_pad <= 64) issue A_hi.[W_hi | W_lo] (N = 2 N_pad) + A_lo.W_hi,
        // wider layers the three products separately.
        uint32_t cq = 0, lc = 0;
        int s = 0;
        uint32_t par = 0;
        for (int item = blockIdx.x; item < p.n_items; item += gridDim.x) {
            const WorkItem wi = work_item(p, item);
            for (int l = 0; l < p.n_layers; ++l, ++lc) {
                const LayerT2& L = p.layers[l];
                const int n_chunks = l == 0 ? wi.n_chunks : L.n_chunks, stack = L.stack;
                const uint32_t idesc_n = tc::idesc_bf16_f32(BM, L.N_pad);
                const uint32_t idesc_2n = tc::idesc_bf16_f32(BM, 2 * L.N_pad);
                const uint32_t d_tmem = tmem_base + (((lc & 1) && !p.wide) ? 128u : 0u);
                const uint32_t lbo_b = (uint32_t)L.N_pad * (BF16 ? 16u : 32u);   // k-core slab = hi rows + lo rows (bf16 mode: hi rows only)
                const uint32_t lo_off = (uint32_t)L.N_pad;               // (N_pad * 16 bytes) >> 4: hi -> lo rows of a slab
                const uint32_t kk_step = (2u * lbo_b) >> 4;              // two slabs per MMA
                ChunkCursor c(L.F_pad, l == 0 ? wi.g0 : 0);
                for (int q = 0; q < n_chunks; ++q, ++cq, c.next()) {
                    if (lane == 0) TRC(2, cq, 0);
                    tc::mbar_wait(&full[s], par);
                    if (lane == 0) TRC(2, cq, 2);
                    tc::tc_fence_after_sync();
                    if (tc::elect_one()) {
                        const uint32_t a_hi = tmem_base + TMEM_A0 + 64u * s, a_lo = a_hi + 32u;
                        const uint64_t d0 = tc::smem_desc(tc::smem_u32(bst + (size_t)s * p.bstage_bytes), lbo_b, 128);
                        const int nkk = c.nk() >> 1;
#pragma unroll
                        for (int kk = 0; kk < 4; ++kk) {
                            if (kk < nkk && !KAGNN_TC2_NOMMA) {
                                const uint64_t dbh = d0 + (uint64_t)(kk * kk_step);
                                const uint32_t acc = (q | kk) != 0 ? 1u : 0u;
                                if (BF16) {
                                    tc::umma_bf16_ts(d_tmem, a_hi + 8u * kk, dbh, idesc_n, acc);
                                } else if (stack) {
                                    // [W_hi | W_lo] is one B operand: A_hi gives hi.hi | hi.lo; A_lo gives lo.hi (and, with
                                    // KAGNN_TC2_STACK4, lo.lo as well: the full product, relative error ~2^-24 instead of ~2^-17,
                                    // for 128 instead of 96 tensor cycles per K step on layers whose tensor pipe is ~30 % busy)
                                    tc::umma_bf16_ts(d_tmem, a_hi + 8u * kk, dbh, idesc_2n, acc);
                                    tc::umma_bf16_ts(d_tmem, a_lo + 8u * kk, dbh, KAGNN_TC2_STACK4 ? idesc_2n : idesc_n, 1u);
                                } else {
                                    tc::umma_bf16_ts(d_tmem, a_hi + 8u * kk, dbh, idesc_n, acc);
                                    tc::umma_bf16_ts(d_tmem, a_hi + 8u * kk, dbh + lo_off, idesc_n, 1u);
                                    tc::umma_bf16_ts(d_tmem, a_lo + 8u * kk, dbh, idesc_n, 1u);
                                }
                            }
                        }
                        tc::umma_commit(&empty[s]);
                        if (q == n_chunks - 1) tc::umma_commit(&acc_full[lc & 1]);
                    }
                    __syncwarp();
                    if (lane == 0) TRC(2, cq, 3);
                    if (++s == p.ns) { s = 0; par ^= 1u; }
                }
            }
        }
    } else {
        // ========================================== W LOADER (+ two idle warps) ================================
        tc::reg_dec<REGS_AUX>();
        if (warp == WARP_LOAD + 1 && KAGNN_TC2_PREFETCH_DIST > 0) {
            // ------------------------------------ L2 PREFETCHER ------------------------------------------------
            // One otherwise idle warp walks the CSR a few tiles ahead of the gather warps and asks L2 for the self and
            // neighbour rows (cp.async.bulk.prefetch.L2, one instruction per row), so that the gather's loads of a cold
            // feature matrix (first touch after the previous layer / an L2 flush) hit L2 instead of paying DRAM latency.
            const KagnnAggregate& a = p.agg;
            const bool csr = a.mode == KAGNN_AGG_GIN || a.mode == KAGNN_AGG_WEIGHTED || a.mode == KAGNN_AGG_GINE;
            const uint32_t row_bytes = (uint32_t)a.num_cols * 4u;
            const bool ok = csr && !a.peer_x && !a.src_index && (a.num_cols % 4 == 0) && ((reinterpret_cast<uintptr_t>(a.x) & 15u) == 0) &&
                            (a.ldx % 4 == 0) && (!a.x_halo || (((reinterpret_cast<uintptr_t>(a.x_halo) & 15u) == 0) && (a.ld_halo % 4 == 0)));
            if (ok) {
                int it = 0;
                for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x, ++it) {
                    uint32_t naps = 0;
                    while (it > *gather_progress + KAGNN_TC2_PREFETCH_DIST && ++naps < (1u << 20)) __nanosleep(256);
                    const long long row0 = (long long)tile * BM;
                    const long long row1 = min(row0 + (long long)BM, p.num_rows);
                    for (long long r = row0 + lane; r < row1; r += 32) tc::prefetch_l2(a.x + r * a.ldx, row_bytes);
                    const int beg = __ldg(a.rowptr + row0), end = __ldg(a.rowptr + row1);
                    for (int e = beg + lane; e < end; e += 32) tc::prefetch_l2(src_row(a, __ldg(a.col + e)), row_bytes);
                }
            }
        }
        if (PUSH && (warp == WARP_PUSH || (warp == WARP_PUSH - 1 && KAGNN_TC2_PREFETCH_DIST == 0))) {
            // ------------------------------------ OUTPUT PUSH ---------------------------------------------------
            // Node-sharded graphs: the rows of every finished tile also go to the other ranks' replicas of this layer's output
            // (KagnnAggregate.push_y).  NVLink WRITES are posted, so two warps relaying rows through shared memory with bulk
            // copies (global -> shared -> peer global) keep up with the tile rate; measured in isolation 16 such CTAs already
            // saturate the link (profiles/r2_nvlink_push.jsonl), where peer LOADS need every SM of the GPU.  The two warps take
            // alternate batches of rows, each with its own half of the relay buffer and its own mbarrier.
            const int pw = (KAGNN_TC2_PREFETCH_DIST == 0) ? warp - (WARP_PUSH - 1) : 0;
            const int npw = (KAGNN_TC2_PREFETCH_DIST == 0) ? 2 : 1;
            const uint32_t rb = (uint32_t)p.layers[p.n_layers - 1].N * 4u;          // bytes of one output row (multiple of 16)
            const int half = p.relay_bytes / npw;
            const int S = min(32, half / (int)rb);                                  // rows per relay batch
            uint64_t* relay_bar = reinterpret_cast<uint64_t*>(relay_block() + 16 * pw);
            const uint32_t tiles_stored = tc::smem_u32(relay_block() + 8);
            uint8_t* slot = relay_block() + 128 + (size_t)pw * half + (size_t)lane * rb;
            float* const* dsts = p.agg.push_y;
            uint32_t ph = 0, want = 0, bi = 0;
            for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x) {
                ++want;
                uint32_t have, tries = 0;
                for (;;) {
                    asm volatile("ld.acquire.cta.shared.u32 %0, [%1];" : "=r"(have) : "r"(tiles_stored) : "memory");
                    if (have >= want) break;
                    __nanosleep(128);
                    if (++tries > (1u << 24)) __trap();
                }
                const long long row0 = (long long)tile * BM;
                const int nrows = (int)min((long long)BM, p.num_rows - row0);
                for (int r0 = 0; r0 < nrows; r0 += S, ++bi) {
                    if ((int)(bi % (uint32_t)npw) != pw) continue;
                    const int cnt = min(S, nrows - r0);
                    tc::bulk_wait_read_all();                                       // the previous batch left the slots
                    __syncwarp();
                    if (lane == 0) tc::mbar_arrive_expect_tx(relay_bar, (uint32_t)cnt * rb);
                    __syncwarp();
                    const long long row = row0 + r0 + lane;
                    if (lane < cnt) tc::bulk_g2s(slot, p.y + row * p.ldy, rb, relay_bar);
                    tc::mbar_wait(relay_bar, ph);
                    ph ^= 1u;
                    if (lane < cnt) {
                        const uint32_t m = p.agg.push_mask ? (uint32_t)__ldg(p.agg.push_mask + row) : 0xffu;
                        for (int i = 0; i < p.agg.num_push; ++i)
                            if ((m >> i) & 1u) tc::bulk_s2g(dsts[i] + row * p.agg.ld_push, slot, rb);
                        tc::bulk_commit();
                    }
                }
            }
            tc::bulk_wait_all();
            __threadfence_system();
        }
        if (warp == WARP_LOAD && lane == 0) {
            uint32_t cq = 0, par = 1;
            int s = 0;
            for (int item = blockIdx.x; item < p.n_items; item += gridDim.x) {
                const WorkItem wi = work_item(p, item);
                for (int l = 0; l < p.n_layers; ++l) {
                    const LayerT2& L = p.layers[l];
                    const int n_chunks = l == 0 ? wi.n_chunks : L.n_chunks;
                    ChunkCursor c(L.F_pad, l == 0 ? wi.g0 : 0);
                    for (int q = 0; q < n_chunks; ++q, ++cq, c.next()) {
                        TRC(3, cq, 0);
                        tc::mbar_wait_relaxed(&empty[s], par);
                        TRC(3, cq, 1);
                        if (BF16) {
                            // hi rows only: one bulk copy per k-core slab (the packed layout interleaves hi and lo rows), dense in the stage
                            const uint32_t slab = 16u * (uint32_t)L.N_pad;
                            const int nk = c.nk();
                            tc::mbar_arrive_expect_tx(&full[s], slab * (uint32_t)nk);
                            for (int i = 0; i < nk; ++i)
                                tc::bulk_g2s(bst + (size_t)s * p.bstage_bytes + (size_t)i * slab, L.wtc + c.b_off(L.N_pad) + (size_t)i * 2u * slab, slab,
                                             &full[s]);
                        } else {
                            const uint32_t bytes = c.b_bytes(L.N_pad) / KAGNN_TC2_WDIV;   // WDIV > 1: development probe (results are wrong)
                            tc::mbar_arrive_expect_tx(&full[s], bytes);
                            tc::bulk_g2s(bst + (size_t)s * p.bstage_bytes, L.wtc + c.b_off(L.N_pad), bytes, &full[s]);
                        }
                        if (++s == p.ns) { s = 0; par ^= 1u; }
                    }
                }
            }
        }
    }
    tc::tc_fence_before_sync();
    __syncthreads();
    if (warp == WARP_MMA) tc::tmem_dealloc(tmem_base, 512);
}

// ---------------------------------------------------------------------------------------------------------------------
// Aggregation-only launch (n_layers == 0: the last layer of a GCN-flavour model, a stand-alone GCNConv after its KAN, pooling
// without a readout): the same flattened-list gather, one warp per 16 destination rows, result (after mean scale / pre-affine)
// written straight to agg_out.  HBM-bound: 16 independent 128-bit row loads in flight per warp, grid = all row groups.
// ---------------------------------------------------------------------------------------------------------------------
#ifndef KAGNN_TC2_VARIANT_G16
constexpr int AGG_WARPS = NGW;
// Occupancy of the aggregation-only kernel: its limiter on a graph that does not fit L2 (R-MAT 10 M nodes) is memory-level
// parallelism, and skewed degrees leave warps of a block idle behind its heaviest one, so more resident blocks with fewer loads
// each win (measured on the 10 M-node / 100 M-edge R-MAT KAGCN layer: 11.5 ms at 3 blocks x 8 loads -> 10.0 ms at 5 x 4).
#ifndef KAGNN_AGG_MINB
#define KAGNN_AGG_MINB 5
#endif
#ifndef KAGNN_AGG_U
#define KAGNN_AGG_U 4
#endif
__global__ void __launch_bounds__(AGG_WARPS * 32, AGG_WARPS == 8 ? KAGNN_AGG_MINB : 2) aggregate_only_kernel(const __grid_constant__ Tc2Params p) {
    __shared__ HubScratch hub;
    const int lane = threadIdx.x & 31, gw = threadIdx.x >> 5;
    const int F = p.agg.num_cols;
    for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x) {     // block = one 128-row tile at a time, warp gw = rows 16 gw ..
        const long long row0 = (long long)tile * BM;
        for (int c0 = 0; c0 < F; c0 += 128) {
            float* out = p.agg_out + row0 * p.ld_agg_out + c0;
            const int ucols = min(128, F - c0);
            bool half = KAGNN_TC2_HALF && ucols <= 64;
            if (half && KAGNN_TC2_HUB && (p.agg.mode == KAGNN_AGG_GIN || p.agg.mode == KAGNN_AGG_WEIGHTED)) {
#pragma unroll
                for (int k4 = 0; k4 < 4; ++k4) {
                    const long long rr = row0 + 32 * k4 + lane;
                    int d = 0;
                    if (rr < p.num_rows) d = __ldg(p.agg.rowptr + rr + 1) - __ldg(p.agg.rowptr + rr);
                    if (__any_sync(0xffffffffu, d > HUB_T)) half = false;       // block-uniform: the cooperative hub pass needs every warp
                }
            }
            if (half) {
                if (p.agg.mode == KAGNN_AGG_WEIGHTED) gather_unit_half<true, true, KAGNN_AGG_U>(p, row0, c0, ucols, out, gw, lane);
                else gather_unit_half<false, true, KAGNN_AGG_U>(p, row0, c0, ucols, out, gw, lane);
            } else if (p.agg.mode == KAGNN_AGG_WEIGHTED) gather_unit_fast<true, true, KAGNN_AGG_U>(p, row0, c0, ucols, out, gw, lane, &hub);
            else gather_unit_fast<false, true, KAGNN_AGG_U>(p, row0, c0, ucols, out, gw, lane, &hub);
        }
    }
}

#endif  // !KAGNN_TC2_VARIANT_G16

inline int ceil16(int v) { return (v + 15) & ~15; }

}  // namespace

#ifndef KAGNN_TC2_VARIANT_G16
int kagnn_aggregate_only_tc2(const KagnnAggregate* agg, int64_t num_rows, const KagnnAffine* pre, float* agg_out,
                             int64_t ld_agg_out, cudaStream_t stream) {
    // shapes the 128-bit flattened-list gather serves; everything else stays on the general kernel
    if (!agg_out || agg->mode == KAGNN_AGG_GINE || agg->num_head_cols || agg->num_cols > 4096) return KAGNN_EUNSUPPORTED;
    if (agg->num_cols % 4 != 0 || !aligned16(agg->x) || agg->ldx % 4 != 0 || !aligned16(agg_out) || ld_agg_out % 4 != 0)
        return KAGNN_EUNSUPPORTED;
    if (agg->x_halo && (!aligned16(agg->x_halo) || agg->ld_halo % 4 != 0)) return KAGNN_EUNSUPPORTED;
    if (agg->peer_x && agg->rows_per_rank * (int64_t)agg->num_ranks > (int64_t)INT32_MAX) return KAGNN_EUNSUPPORTED;
    if (agg->mode == KAGNN_AGG_NONE && !pre && !agg->src_index) return KAGNN_EUNSUPPORTED;      // a plain copy: not worth a gather
    Tc2Params p{};
    p.agg = *agg;
    p.has_pre = pre != nullptr;
    if (pre) p.pre = *pre;
    p.num_rows = num_rows;
    p.agg_out = agg_out;
    p.ld_agg_out = ld_agg_out;
    p.xld = (int)ld_agg_out;
    if (ld_agg_out > (int64_t)INT32_MAX) return KAGNN_EUNSUPPORTED;
    static_assert(AGG_WARPS == NGW && AGG_WARPS * RPW == BM, "aggregate_only_kernel: one block = one 128-row tile of NGW warps");
    unsigned blocks = (unsigned)ceil_div64(num_rows, BM);
    p.n_tiles = (int)blocks;
    size_t dbg_smem = 0;
#ifdef KAGNN_DEBUG_KNOBS
    if (const char* e = getenv("KAGNN_DEBUG_AGG_GRID")) blocks = (unsigned)atoi(e);
    if (const char* e = getenv("KAGNN_DEBUG_AGG_SMEM")) {
        dbg_smem = (size_t)atoi(e);
        cudaFuncSetAttribute(aggregate_only_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dbg_smem);
    }
#endif
    aggregate_only_kernel<<<blocks, AGG_WARPS * 32, dbg_smem, stream>>>(p);
    KAGNN_LAUNCH_CHECK();
    return KAGNN_OK;
}

#endif  // !KAGNN_TC2_VARIANT_G16

// This file is compiled twice (fused_tc2_g16.cu includes it with the other role split): 16 basis-producer + 8 gather warps for
// launches without a neighbour gather (bare KAN chains, the read-out: producer-bound), 8 producer + 16 gather warps for the
// launches that gather over a CSR (GIN / GCN / GINE: bound by the memory-level parallelism of the gather).  Measured on the
// arxiv-shaped layers: gather layers 0.249 -> 0.226 ms (128 wide) and 0.185 -> 0.167 ms (64 wide) with 16 gather warps, plain
// layers 0.157 -> 0.179 ms and the read-out 0.226 -> 0.267 ms -- hence one variant each.
#ifdef KAGNN_TC2_VARIANT_G16
#define KAGNN_TC2_ENTRY kagnn_fused_fwd_tc2_g16
#else
#define KAGNN_TC2_ENTRY kagnn_fused_fwd_tc2
#endif
int KAGNN_TC2_ENTRY(const KagnnAggregate* agg, int64_t num_rows, const KagnnAffine* pre, float* agg_out,
                    int64_t ld_agg_out, int32_t n_layers, const KagnnKanLayer* layers, const KagnnAffine* post, float* y,
                    int64_t ldy, cudaStream_t stream) {
    if (n_layers < 1 || n_layers > KAGNN_MAX_LAYERS) return KAGNN_EUNSUPPORTED;
#if !defined(KAGNN_TC2_VARIANT_G16) && KAGNN_TC2_USE_G16
    if (agg->mode == KAGNN_AGG_GIN || agg->mode == KAGNN_AGG_WEIGHTED || agg->mode == KAGNN_AGG_GINE)
        return kagnn_fused_fwd_tc2_g16(agg, num_rows, pre, agg_out, ld_agg_out, n_layers, layers, post, y, ldy, stream);
#endif
    DeviceProps props{};
    int rc = kagnn_get_props(&props);
    if (rc != KAGNN_OK) return rc;
    if (props.cc_major != 10) return KAGNN_EUNSUPPORTED;

    Tc2Params p{};
    p.agg = *agg;
    p.has_pre = pre != nullptr;
    if (pre) p.pre = *pre;
    p.has_post = post != nullptr;
    if (post) p.post = *post;
    p.num_rows = num_rows;
    p.agg_out = agg_out;
    p.ld_agg_out = ld_agg_out;
    p.y = y;
    p.ldy = ldy;
    p.n_layers = n_layers;
    p.n_tiles = (int)ceil_div64(num_rows, BM);

    const bool rbf = layers[0].basis == KAGNN_BASIS_RBF;
    const int k = rbf ? 0 : layers[0].spline_order;
    const bool bf16 = kagnn_get_precision() == KAGNN_PREC_BF16;
    int width = agg->num_cols, n_max = 0;
    for (int l = 0; l < n_layers; ++l) {
        const KagnnKanLayer& s = layers[l];
        LayerT2& d = p.layers[l];
        if (s.basis != layers[0].basis || !s.packed_w_tc || s.in_features != width) return KAGNN_EUNSUPPORTED;
        if (rbf) {
            if (s.grid_size < 1 || s.grid_size > 8) return KAGNN_EUNSUPPORTED;
        } else {
            if (s.basis != KAGNN_BASIS_BSPLINE) return KAGNN_EINVAL;
            if (s.spline_order != k || k < 1 || k > 3 || s.grid_size < 1 || s.grid_size + k > 8) return KAGNN_EUNSUPPORTED;
            if (!(s.h > 0.f)) return KAGNN_EINVAL;
        }
        // up to 128 outputs per layer in a chain (two accumulator regions of 128 columns ping-pong between chained layers);
        // a single layer may be up to 256 wide (one region of 256 columns, epilogue not overlapped with the next tile)
        if (s.out_features <= 0 || s.out_features > (n_layers == 1 ? 256 : 128)) return KAGNN_EUNSUPPORTED;
        if (!aligned16(s.packed_w_tc)) return KAGNN_EALIGN;
        d.F = s.in_features;
        d.F_pad = ceil16(s.in_features);
        d.N = s.out_features;
        d.N_pad = ceil16(s.out_features);
        d.n_chunks = d.F_pad / 8 + (d.F_pad + 63) / 64;
        d.stack = (d.N_pad <= 64 && !bf16) ? 1 : 0;
        if (rbf) {
            d.rc0 = s.t0;
            d.rstep = s.h;
            d.rk = s.inv_denominator * 1.2011224087864498f;          // sqrt(log2 e): exp(-d^2) = 2^-(d sqrt(log2 e))^2
            d.bias = s.base_bias;
            d.lnw = s.ln_weight;
            d.lnb = s.ln_bias;
            d.ln_stats = (l == 0 && agg->mode == KAGNN_AGG_NONE && !pre) ? s.ln_stats : nullptr;
        } else {
            d.inv_h = 1.0f / s.h;
            d.c0 = -s.t0 * d.inv_h;
            d.lim = (float)(s.grid_size + 2 * k);
        }
        d.wtc = static_cast<const uint8_t*>(s.packed_w_tc);
        if (d.N_pad > n_max) n_max = d.N_pad;
        width = s.out_features;
    }
    if (ldy < width) return KAGNN_EINVAL;
    if (agg->num_cols > 4096) return KAGNN_EUNSUPPORTED;   // g_zero_row covers 4096 columns
    if (agg->num_head_cols) {                              // two-part rows: plain-copy path, split on a unit boundary
        if (agg->mode != KAGNN_AGG_NONE || pre || agg_out || agg->src_index) return KAGNN_EUNSUPPORTED;
        if (agg->num_head_cols % 128 != 0 || agg->num_cols - agg->num_head_cols <= 0) return KAGNN_EUNSUPPORTED;
    }
    if (agg->peer_x) {                                     // served by gather_unit_fast only (128-bit loads)
        if (agg->num_cols % 4 != 0 || !aligned16(agg->x) || agg->ldx % 4 != 0) return KAGNN_EUNSUPPORTED;
        if (agg->rows_per_rank * (int64_t)agg->num_ranks > (int64_t)INT32_MAX) return KAGNN_EUNSUPPORTED;
    }
    p.y_vec = aligned16(y) && (ldy % 4 == 0);
    if (agg->num_push > 0) {                               // pushed output: whole rows move as bulk copies (16-byte granules)
        if (!agg->push_y || agg->num_push > 8) return KAGNN_EINVAL;
        if (!p.y_vec || width % 4 != 0 || width > 128 || agg->ld_push % 4 != 0 || agg->ld_push < width) return KAGNN_EUNSUPPORTED;
    }

    const int F_pad0 = p.layers[0].F_pad;
    // asynchronous gather: GIN / GCN aggregation of rows whose width is a multiple of 64 columns, 128-bit aligned operands
    p.ag = KAGNN_TC2_AG && (agg->mode == KAGNN_AGG_GIN || agg->mode == KAGNN_AGG_WEIGHTED) && agg->num_cols % 64 == 0 &&
           aligned16(agg->x) && agg->ldx % 4 == 0 && (!agg->x_halo || (aligned16(agg->x_halo) && agg->ld_halo % 4 == 0)) &&
           num_rows < (1LL << 30) && (!agg->peer_x || agg->rows_per_rank * (int64_t)agg->num_ranks < (1LL << 30)) &&
           (!agg->x_halo || agg->num_local_src < (1LL << 29)) &&
           !(rbf && layers[0].ln_weight && F_pad0 > 64);          // in-kernel LayerNorm statistics need the row in ONE unit
    p.wide = n_max > 128;
    p.bf16 = bf16 ? 1 : 0;
    p.bstage_bytes = 256 * n_max;
    // LUTs | mbarriers (x ring full/empty, stage full/empty, 2 accumulator) | tmem slot (16 B) | post scale/shift | progress word
    const int tail0 = KAGNN_MAX_LAYERS * LUT_ROWS * 16 + (2 * MAX_UNITS + 2 * MAX_STAGE + 2) * 8 + 16 + 2 * 128 * 4 + 16 + (int)sizeof(HubScratch) + 2 * NWG * 128 * 8 + 2 * 128 * 4 +
                     (p.ag ? AG_BYTES : 0);
    // x-ring geometry: units of 128 or 64 columns.  128 (one unit per tile up to 128 inputs) is the default; 64-column units are
    // forced by the asynchronous gather and by wide layers (64 KB W stages), and preferred for plain row tiles (no gather) when
    // they buy a deeper W / A stage ring.  Ring depths: at least 2 units and 2 stages; deeper stages first, then more units.
    // (pushed output: the relay of the push warps comes out of the same budget -- 16 KB when the stage ring keeps its full depth
    // with it, else 8 KB)
    int best_units = 0, best_ns = 0, best_uw = 0;
    auto choose = [&](int relay) {
        best_units = best_ns = best_uw = 0;
        p.relay_bytes = relay;
        const int tail = tail0 + (relay ? relay + RELAY_PAD : 0);
        for (int uw = 128; uw >= 64; uw >>= 1) {
            if (uw == 128 && (F_pad0 <= 64 || p.ag || p.wide)) continue;
            if (uw == 64 && best_units != 0 && !(agg->mode == KAGNN_AGG_NONE && !pre && !agg->src_index)) break;
            // in-kernel LayerNorm statistics need the whole input row in ONE unit
            if (uw == 64 && best_units != 0 && rbf && layers[0].ln_weight && !p.layers[0].ln_stats) break;
            const int unit_bytes_c = BM * (uw + 4) * (int)sizeof(float);
            for (int nu = 2; nu <= MAX_UNITS; ++nu) {
                const int left = (int)props.max_smem - tail - nu * unit_bytes_c;
                int ns = left / p.bstage_bytes;
                if (ns > MAX_STAGE) ns = MAX_STAGE;
                if (ns < 2) break;
                if (best_units == 0 || ns > best_ns || (ns == best_ns && uw == best_uw)) { best_units = nu; best_ns = ns; best_uw = uw; }
                else break;
            }
        }
    };
    if (agg->num_push > 0) {
        choose(16384);
        if (best_units == 0 || best_ns < MAX_STAGE) choose(8192);
    } else {
        choose(0);
    }
    if (best_units == 0) return KAGNN_EUNSUPPORTED;
    p.uw = best_uw;
    p.uw_shift = p.uw == 128 ? 7 : 6;
    p.xld = p.uw + 4;                                   // (xld / 4) odd: conflict-free float4 reads with thread = row
    p.unit_floats = BM * p.xld;
    p.units_per_tile = (F_pad0 + p.uw - 1) / p.uw;
    const int unit_bytes = p.unit_floats * (int)sizeof(float);
    p.n_units = best_units;
    p.ns = best_ns;
    const size_t smem = (size_t)p.n_units * unit_bytes + (size_t)p.ns * p.bstage_bytes + tail0 + (p.relay_bytes ? p.relay_bytes + RELAY_PAD : 0);

    // in-kernel LayerNorm statistics need the whole input row in one x unit; wider rows need the pre-pass (ln_stats)
    if (rbf && p.units_per_tile != 1 && p.layers[0].lnw && !p.layers[0].ln_stats) return KAGNN_EUNSUPPORTED;
    void (*kern)(Tc2Params);
    // the pushed-output variant is a separate instantiation: its relay warp and the extra barrier cost the other launches
    // registers in the producer loop (measured 3 % of the step when it was a run-time switch)
    if (agg->num_push > 0 && bf16) return KAGNN_EUNSUPPORTED;
    if (agg->num_push > 0)
        kern = k == 3 ? fused_tc2_kernel<3, false, true> : (k == 2 ? fused_tc2_kernel<2, false, true> : (k == 1 ? fused_tc2_kernel<1, false, true> : fused_tc2_kernel<0, false, true>));
    else if (bf16)
        kern = k == 3 ? fused_tc2_kernel<3, true, false> : (k == 2 ? fused_tc2_kernel<2, true, false> : (k == 1 ? fused_tc2_kernel<1, true, false> : fused_tc2_kernel<0, true, false>));
    else
        kern = k == 3 ? fused_tc2_kernel<3, false, false> : (k == 2 ? fused_tc2_kernel<2, false, false> : (k == 1 ? fused_tc2_kernel<1, false, false> : fused_tc2_kernel<0, false, false>));
    KAGNN_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)props.max_smem));
    if (agg->halo_flags && (!agg->x_halo || !agg->halo_need)) return KAGNN_EINVAL;
    int sms = props.num_sms - (agg->reserve_sms > 0 ? agg->reserve_sms : 0);
    if (sms < 1) sms = 1;
    // Split-K for launches with few tiles and long rows (BASELINE config C1: 2 708 nodes x 1 433 features = 22 tiles, 203 chunks
    // each): a tile's x units are dealt to several CTAs, each runs the whole pipeline over its window of input features and
    // adds its partial sums into y with float atomics.  Only where nothing follows the contraction inside the launch: one
    // B-spline layer, rows as they are, no epilogue affine, no pushed output.
    p.n_split = 1;
    p.units_per_item = p.units_per_tile;
    const bool can_split = n_layers == 1 && !rbf && agg->mode == KAGNN_AGG_NONE && !pre && !agg_out && !agg->src_index && !post &&
                           agg->num_push == 0 && !agg->halo_flags;
    if (kSplitK && can_split && p.units_per_tile >= 4 && p.n_tiles * 2 <= sms) {
        const int want = sms / p.n_tiles;                  // items per tile that still fit one wave
        int upi = (p.units_per_tile + want - 1) / want;
        if (upi < 2) upi = 2;
        p.units_per_item = upi;
        p.n_split = (p.units_per_tile + upi - 1) / upi;
    }
    p.n_items = p.n_tiles * p.n_split;
    if (p.n_split > 1)
        KAGNN_CUDA_TRY(cudaMemset2DAsync(y, (size_t)ldy * sizeof(float), 0, (size_t)width * sizeof(float), (size_t)num_rows, stream));
    const int grid = p.n_items < sms ? p.n_items : sms;
    kern<<<(unsigned)grid, NTHREADS, smem, stream>>>(p);
    KAGNN_LAUNCH_CHECK();
    return KAGNN_OK;
}
