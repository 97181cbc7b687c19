// tn_mlp.cuh -- the fused interpolate -> MLP kernel on tcgen05 (sm_100a).
//
// Replaces, for one pass over n_active*S samples, the chain
//   interpolate_values (src/tetrahedra_tracer.cu:195-221)  ->  mlp_base (3x Linear+ReLU, model.py:433-438)
//   -> DensityFieldHead (Linear+Softplus, :455) [-> mlp_head (Linear+ReLU, :447-452) -> RGBFieldHead
//   (Linear+Sigmoid, :454)]                                               (call sites model.py:569-621)
// which the reference runs as ~10 torch kernels with [R*S,128] fp32 activations round-tripping HBM.
//
// One persistent CTA per SM, 24 warps in six warpgroups with per-role register budgets (setmaxnreg):
//   warps 0..15  : epilogue workers, two "slots" of 8 warps.  A slot owns one 128-sample tile at a time; warp (q,h)
//                  of a slot owns sample rows 32q..32q+31 (its TMEM lane quarter, q = warp_id % 4) and the column
//                  half h.  For each layer it waits for the accumulator, applies bias + ReLU to its 64 columns, converts
//                  the result (bf16 hi/lo pair, or one fp16 value) and writes the next A operand back into TMEM (activations
//                  never touch shared or global memory); the last layer's epilogue also forms the head dot products.
//   warp 16      : TMEM allocation, TMA bulk staging of the resident weight image, and (FINE) the producer of the
//                  2-stage (f16w2: 3-stage) TMA ring that streams the 4th layer's weights in 16 KB (128 output rows x 64 K) chunks
//   warp 17      : issues every tcgen05.mma (one elected lane), strictly alternating between the two slots, which run
//                  half a tile apart: one slot's epilogue overlaps the other slot's MMAs
//   warps 20..23 : gather warps.  They run AHEAD of the slots: tile after tile they fetch the four vertex rows of
//                  every sample from the [V,64] field shadow (16-byte loads, 16 in flight per lane), form the
//                  barycentric interpolation with the reference's FMA order, split it into bf16 hi/lo and store it
//                  as the layer-0 A operand in shared memory (K-major, 128-byte swizzle), so the L2 latency of the
//                  gather hides under the MMAs and epilogues of the previous tiles.  Layer 0 is an SS MMA (A and B
//                  from shared memory), layers 1.. are TS MMAs (A from TMEM).
// Products: two operand precisions, template parameter PREC (see mlp_ring_stages below for the full note).
//   PREC 3 "bf16x3": a*w ~= a_hi*w_hi + a_lo*w_hi + a_hi*w_lo, bf16 halves, fp32 accumulation in TMEM (measured 5e-6 relative on B200,
//           tests/test_gpu_mlp.py) -- the reference computes in fp32 and the parity bar is 1e-4 absolute on colour/density, which
//           single-pass bf16/tf32 cannot hold.  Training forward.
//   PREC 2 "f16w2":  a*(w_hi + w_lo) with ONE fp16 activation value and fp16 hi/lo weights (2.8e-5 absolute per sample on unit-scale
//           outputs, tests/test_gpu_render.py) -- the inference default.
// TMEM (512 columns): slot s uses [256s, 256s+128) for D, [+128,+192) A (hi), [+192,+256) A_lo (bf16x3 only).
#pragma once
#include "tn_common.cuh"
#include "tn_tc.cuh"

#ifndef TN_MLP_STATS
#define TN_MLP_STATS 0   // 1: the gather warps also accumulate busy / wait cycles into the debug timeline (costs registers)
#endif
#ifndef TN_MLP_F16_WIDE
#define TN_MLP_F16_WIDE 1   // f16w2 epilogue with two 32-column TMEM loads per warp instead of four 16-column ones
#endif
#ifndef TN_MLP_GATHER_ON
#define TN_MLP_GATHER_ON true
#endif
namespace tn {

constexpr uint32_t MLP_THREADS = 768;
constexpr uint32_t MLP_GATHER_WARP0 = 20, MLP_GATHER_WARPS = 4;
// setmaxnreg only moves registers inside the CTA's launch allocation (768 threads x 80): 512*88 + 128*40 + 128*88 = 61440
constexpr uint32_t MLP_REGS_WORKER = 88, MLP_REGS_CTRL = 40, MLP_REGS_GATHER = 88;
static_assert(512 * MLP_REGS_WORKER + 128 * MLP_REGS_CTRL + 128 * MLP_REGS_GATHER <= MLP_THREADS * 80, "register budget exceeds the launch allocation: setmaxnreg.inc would block forever");
constexpr uint32_t MLP_TL_CAP = 96;                        // debug timeline records buffered in shared memory
constexpr uint32_t MLP_W_RESIDENT = 163840;                // L1 32K + L2 64K + L3 64K
constexpr uint32_t MLP_RING_STAGES = 2, MLP_RING_CHUNK = 16384;   // (stages of the bf16x3 mode; the f16w2 mode has three, see mlp_ring_stages)
// Operand precision of the tensor-core products (template parameter PREC of k_mlp):
//   3 = "bf16x3": a*w ~= a_hi*w_hi + a_lo*w_hi + a_hi*w_lo, bf16 halves, 3 MMAs per K step -- 5e-7 absolute on density / colour of
//       unit scale, the training forward and the reference mode;
//   2 = "f16w2": the activations are ONE fp16 value (11-bit significand), the weights an fp16 hi/lo pair: a*w ~= a*(w_hi + w_lo),
//       2 MMAs per K step, no A_lo operand (half the epilogue's conversion work and TMEM stores, a 16 KB layer-0 operand, and the
//       freed 16 KB become a third stage of the layer-4 weight ring).  Error ~2^-12 relative per activation: 2.6e-5 absolute on
//       density / colour of unit scale (tools/split_accuracy.py), inside the 1e-4 per-sample bar.
__host__ __device__ constexpr uint32_t mlp_ring_stages(int prec) { return prec == 2 ? 3u : 2u; }
constexpr uint32_t MLP_OFF_RING = MLP_W_RESIDENT;                                // 2 x 16K
constexpr uint32_t MLP_OFF_A0 = MLP_OFF_RING + MLP_RING_STAGES * MLP_RING_CHUNK; // layer-0 A operand: hi 16K | lo 16K
constexpr uint32_t MLP_OFF_HEAD = MLP_OFF_A0 + 32768;                            // wd[128] wc[3][128] bd bc[3] (+pad)
constexpr uint32_t MLP_OFF_BARS = MLP_OFF_HEAD + 520 * 4;
constexpr uint32_t MLP_OFF_TL = MLP_OFF_BARS + 160;  // debug timeline: counter, enable flag, MLP_TL_CAP records
// (the hidden-layer biases are read through L1 and the head partial sums are exchanged through TMEM: with 160 KB of
//  resident weights, the 32 KB ring and the 32 KB layer-0 operand there is no shared memory left for them)
constexpr uint32_t MLP_SMEM_BYTES = MLP_OFF_TL + 8 * (MLP_TL_CAP + 2);
static_assert(MLP_OFF_A0 % 1024 == 0 && MLP_OFF_RING % 1024 == 0, "swizzled operands need 1024-byte alignment");
static_assert(MLP_SMEM_BYTES <= 232448, "k_mlp shared memory exceeds 227 KB");

struct MlpParams {
    const uint32_t *n_active;  // device scalar: number of non-empty rays
    uint32_t S;                // samples per ray in this pass
    const uint4 *vi;           // [n_active*S] matched vertex ids (E = unmatched)
    const float *bary;         // [n_active*S,3]
    const float *fshadow;      // [V,64] row-major field
    const uint8_t *wimg;       // weight image: L1 | L2 | L3 | L4(base part), see tn_mlp_pack.cuh
    const float *bias;         // b1,b2,b3 [3][128]
    const float *head;         // wd[128], wc[3][128], bd, bc[3]
    const float *dirbias;      // FINE: [n_active,128]  = b4 + W4[:, :27] . enc(dir)
    float *out;                // COARSE: density [rows] ; FINE: (sigma,r,g,b) [rows,4]
    uint32_t *tile_ctr;        // device counter (zeroed before the launch): dynamic tile scheduler
    unsigned long long *timeline;  // debug: [0] = count, then (tag << 40 | clock) records of CTA 0; nullptr in production
};

// debug timeline: tag = warp(5) | event(4) << 5 | tile(8) << 9 | layer(3) << 17 ; records are buffered in shared
// memory (cheap) and copied to global by CTA 0 at kernel end
__device__ __forceinline__ void tl_mark(unsigned long long *buf, int lane, uint32_t warp, uint32_t ev, uint32_t it, uint32_t layer);

__device__ __forceinline__ float softplus_f(float x) { return x > 20.0f ? x : log1pf(expf(x)); }  // torch Softplus(beta=1, threshold=20)
__device__ __forceinline__ float sigmoid_f(float x) { return 1.0f / (1.0f + expf(-x)); }

// packed fp32x2 arithmetic (sm_100 FADD2 / FFMA2): halves the issue slots of the epilogue
__device__ __forceinline__ float2 add2(float2 a, float2 b) {
    float2 d;
    asm("{.reg .b64 ra, rb, rd;\n\tmov.b64 ra, {%2, %3};\n\tmov.b64 rb, {%4, %5};\n\tadd.rn.f32x2 rd, ra, rb;\n\tmov.b64 {%0, %1}, rd;}"
        : "=f"(d.x), "=f"(d.y) : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y));
    return d;
}
__device__ __forceinline__ float2 fma2(float2 a, float2 b, float2 c) {
    float2 d;
    asm("{.reg .b64 ra, rb, rc, rd;\n\tmov.b64 ra, {%2, %3};\n\tmov.b64 rb, {%4, %5};\n\tmov.b64 rc, {%6, %7};\n\t"
        "fma.rn.f32x2 rd, ra, rb, rc;\n\tmov.b64 {%0, %1}, rd;}"
        : "=f"(d.x), "=f"(d.y) : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y), "f"(c.x), "f"(c.y));
    return d;
}

// ReLU'd pair -> bf16 hi/lo, packed (element 2c in the low half).  r must already be >= 0.
__device__ __forceinline__ void split2(float2 r, uint32_t &hi, uint32_t &lo) {
    uint32_t h;
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(h) : "f"(r.y), "f"(r.x));  // upper <- r.y, lower <- r.x
    const float2 hf = make_float2(-__uint_as_float(h << 16), -__uint_as_float(h & 0xFFFF0000u));
    const float2 d = add2(r, hf);
    uint32_t l;
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(l) : "f"(d.y), "f"(d.x));
    hi = h;
    lo = l;
}

// epilogue of one layer for one warp: its 32 rows x 64 columns [64h, 64h+64) of the accumulator.
//   KIND 0: hidden layer  -> bias + ReLU, next A operand; with `dens` (3rd layer of the FINE pass) also the density-head partial dot
//           product (acc.x)
//   KIND 2: 3rd layer, last (COARSE) -> bias + ReLU + density partial, no next operand
//   KIND 3: 4th layer (FINE, last)   -> per-ray bias + ReLU + colour-head partial (acc.y/z/w)
// (one body per KIND and a NON-unrolled layer loop around them: the 16 worker warps of the two slots run different layers at the same
//  time, and with one copy of this code per layer the epilogues did not fit the 32 KB instruction cache -- ncu: 15 % of all stall
//  samples "no instruction".)
// The bias of the first 16 columns is loaded by the caller BEFORE it waits for the accumulator (bpre), and every bias register is
// refilled with the next chunk's value as soon as it has been consumed: the loads (L1 / L2, SM-dependent latency) never sit between
// the TMEM load and the first add any more.
template <int KIND, int PREC>
__device__ __forceinline__ void layer_epilogue(uint32_t d_t, uint32_t ahi, uint32_t alo, uint32_t h, const float *__restrict__ bias128, float2 (&bpre)[8],
                                               bool dens, const float *__restrict__ wd, const float *__restrict__ wc, float4 &acc) {
    using namespace tc;
    float2 dsum = make_float2(0.f, 0.f), c0 = dsum, c1 = dsum, c2 = dsum;
    // 4 chunks of 16 accumulator columns (small register footprint); the TMEM load of chunk ch+1 is in flight while
    // chunk ch is processed
    uint32_t rbuf[2][16];
    tmem_ld16(d_t + h * 64u, rbuf[0]);
#pragma unroll
    for (int ch = 0; ch < 4; ++ch) {
        const uint32_t col0 = h * 64u + ch * 16u;
        tmem_ld_wait();
        if (ch + 1 < 4) tmem_ld16(d_t + col0 + 16u, rbuf[(ch + 1) & 1]);
        const uint32_t *r = rbuf[ch & 1];
        uint32_t ph[8], pl[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const float2 b = bpre[i];
            if (ch + 1 < 4) bpre[i] = __ldg(reinterpret_cast<const float2 *>(bias128 + col0 + 16u + 2 * i));  // next chunk's bias, a chunk ahead
            float2 x = add2(make_float2(__uint_as_float(r[2 * i]), __uint_as_float(r[2 * i + 1])), b);
            x.x = fmaxf(x.x, 0.f);
            x.y = fmaxf(x.y, 0.f);
            if (KIND == 0) { if (PREC == 2) ph[i] = tc::pack2_f16(x.x, x.y); else split2(x, ph[i], pl[i]); }
            if (KIND == 2 || (KIND == 0 && dens)) dsum = fma2(x, *reinterpret_cast<const float2 *>(wd + col0 + 2 * i), dsum);
            if (KIND == 3) {
                c0 = fma2(x, *reinterpret_cast<const float2 *>(wc + col0 + 2 * i), c0);
                c1 = fma2(x, *reinterpret_cast<const float2 *>(wc + 128 + col0 + 2 * i), c1);
                c2 = fma2(x, *reinterpret_cast<const float2 *>(wc + 256 + col0 + 2 * i), c2);
            }
        }
        if (KIND == 0) {
            tmem_st8(ahi + (col0 >> 1), ph);
            if (PREC != 2) tmem_st8(alo + (col0 >> 1), pl);
        }
    }
    if (KIND == 2 || (KIND == 0 && dens)) acc.x = dsum.x + dsum.y;
    if (KIND == 3) { acc.y = c0.x + c0.y; acc.z = c1.x + c1.y; acc.w = c2.x + c2.y; }
}
// f16w2 form of the epilogue: the conversion is four instructions per pair, so what the epilogue waits for is TMEM -- the in-kernel
// timeline put an epilogue at 1.7-3.3k cycles, mostly the four load round trips (~230 cycles each under the other slot's MMAs).  Here the
// warp's 64 columns come in TWO loads of 32 columns, the second in flight while the first is processed; the biases are read on the
// fly (L1).  64 data registers: possible because this mode keeps no lo halves.
template <int KIND>
__device__ __forceinline__ void layer_epilogue_f16(uint32_t d_t, uint32_t ahi, uint32_t h, const float *__restrict__ bias128, bool dens,
                                                   const float *__restrict__ wd, const float *__restrict__ wc, float4 &acc) {
    using namespace tc;
    float2 dsum = make_float2(0.f, 0.f), c0 = dsum, c1 = dsum, c2 = dsum;
    uint32_t ra[32], rb[32];
    tmem_ld32(d_t + h * 64u, ra);
    tmem_ld_wait();
    tmem_ld32(d_t + h * 64u + 32u, rb);
#pragma unroll
    for (int half = 0; half < 2; ++half) {
        if (half == 1) tmem_ld_wait();
        const uint32_t *r = half ? rb : ra;
#pragma unroll
        for (int g = 0; g < 2; ++g) {  // 16 columns -> 8 packed words -> one tcgen05.st
            const uint32_t col0 = h * 64u + half * 32u + g * 16u;
            uint32_t ph[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const float2 b = __ldg(reinterpret_cast<const float2 *>(bias128 + col0 + 2 * i));
                float2 x = add2(make_float2(__uint_as_float(r[g * 16 + 2 * i]), __uint_as_float(r[g * 16 + 2 * i + 1])), b);
                x.x = fmaxf(x.x, 0.f);
                x.y = fmaxf(x.y, 0.f);
                if (KIND == 0) ph[i] = pack2_f16(x.x, x.y);
                if (KIND == 2 || (KIND == 0 && dens)) dsum = fma2(x, *reinterpret_cast<const float2 *>(wd + col0 + 2 * i), dsum);
                if (KIND == 3) {
                    c0 = fma2(x, *reinterpret_cast<const float2 *>(wc + col0 + 2 * i), c0);
                    c1 = fma2(x, *reinterpret_cast<const float2 *>(wc + 128 + col0 + 2 * i), c1);
                    c2 = fma2(x, *reinterpret_cast<const float2 *>(wc + 256 + col0 + 2 * i), c2);
                }
            }
            if (KIND == 0) tmem_st8(ahi + (col0 >> 1), ph);
        }
    }
    if (KIND == 2 || (KIND == 0 && dens)) acc.x = dsum.x + dsum.y;
    if (KIND == 3) { acc.y = c0.x + c0.y; acc.z = c1.x + c1.y; acc.w = c2.x + c2.y; }
}
// the first chunk's bias of a layer (16 columns from 64 h): issued before the wait for that layer's accumulator
__device__ __forceinline__ void bias_prefetch(const float *__restrict__ bias128, uint32_t h, float2 (&bpre)[8]) {
#pragma unroll
    for (int i = 0; i < 8; ++i) bpre[i] = __ldg(reinterpret_cast<const float2 *>(bias128 + h * 64u + 2 * i));
}

extern __shared__ __align__(1024) uint8_t tn_mlp_smem[];
__device__ __forceinline__ void tl_mark(unsigned long long *buf, int lane, uint32_t warp, uint32_t ev, uint32_t it, uint32_t layer) {
    // records warps 0, 8 (one worker per slot), 17 (issuer), 20 (gather) of CTA 0 once the enable flag is set (steady state)
    if (buf != nullptr && blockIdx.x == 0 && lane == 0 && ((0x00120101u >> warp) & 1u)) {
        volatile unsigned long long *tl = reinterpret_cast<volatile unsigned long long *>(tn_mlp_smem + MLP_OFF_TL);
        if (tl[1] != 0ull) {
            const unsigned long long i = atomicAdd(const_cast<unsigned long long *>(tl), 1ull);
            if (i < MLP_TL_CAP) tl[2 + i] = ((unsigned long long)(warp | (ev << 5) | ((it & 255u) << 9) | (layer << 17)) << 40) | ((unsigned long long)clock64() & 0xFFFFFFFFFFull);
        }
    }
}

// 16-byte read-only load that does not allocate in L1 (the gathered field rows stream through; L1 is left to the biases)
__device__ __forceinline__ float4 ldg_stream(const float4 *p) {
    float4 v;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
    return v;
}

template <uint32_t N> __device__ __forceinline__ void reg_inc() { asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(N)); }
template <uint32_t N> __device__ __forceinline__ void reg_dec() { asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(N)); }

// state of the MMA issuer: scalars and 32-bit shared-memory addresses only (it runs with a 40-register budget and nothing
// here may end up in local memory); barrier k of the kernel's barrier block is at bars + 8 k
struct MlpIssue {
    uint32_t tbase, bars;
    uint32_t dw, dr, da0;  // low descriptor words (address >> 4) of the resident image, the layer-4 ring and the layer-0 A operand
    uint32_t rst, rpar;    // ring stage / parity
    uint32_t nseq, done;   // sequence number of the CTA's next tile; set once the tile scheduler has run dry
};
constexpr uint32_t MLP_BAR_A_READY = 0, MLP_BAR_D_READY = 2, MLP_BAR_W = 4, MLP_BAR_RING_FULL = 5, MLP_BAR_RING_EMPTY = 8, MLP_BAR_A0_FULL = 11,
                   MLP_BAR_A0_EMPTY = 13;
constexpr uint32_t MLP_TILE_IDS_OFF = 128;        // byte offset of tile_ids[8] inside the barrier block (after 15 barriers + tmem ptr + stop flag)
constexpr uint32_t MLP_NO_TILE = 0xFFFFFFFFu;     // sentinel: the scheduler has run dry

// 12 MMAs of one 64-wide K block KB with A in TMEM: (A_hi,W_hi) (A_lo,W_hi) (A_hi,W_lo).  a = TMEM address of A_hi
// (A_lo is 64 columns further); WH / WL = low-descriptor-word offsets of the hi / lo weight blocks in the resident image
template <bool FIRST, uint32_t KB, uint32_t WH, uint32_t WL, int PREC>
__device__ __forceinline__ void kblock_ts(uint32_t d_t, uint32_t a, uint32_t dw, uint32_t idesc) {
    using namespace tc;
    mma_ts_o<!FIRST, KB * 32u + 0u, WH + 0u>(d_t, a, dw, idesc);
    mma_ts_o<true, KB * 32u + 8u, WH + 2u>(d_t, a, dw, idesc);
    mma_ts_o<true, KB * 32u + 16u, WH + 4u>(d_t, a, dw, idesc);
    mma_ts_o<true, KB * 32u + 24u, WH + 6u>(d_t, a, dw, idesc);
    if (PREC != 2) {  // (A_lo, W_hi)
        mma_ts_o<true, 64u + KB * 32u + 0u, WH + 0u>(d_t, a, dw, idesc);
        mma_ts_o<true, 64u + KB * 32u + 8u, WH + 2u>(d_t, a, dw, idesc);
        mma_ts_o<true, 64u + KB * 32u + 16u, WH + 4u>(d_t, a, dw, idesc);
        mma_ts_o<true, 64u + KB * 32u + 24u, WH + 6u>(d_t, a, dw, idesc);
    }
    mma_ts_o<true, KB * 32u + 0u, WL + 0u>(d_t, a, dw, idesc);
    mma_ts_o<true, KB * 32u + 8u, WL + 2u>(d_t, a, dw, idesc);
    mma_ts_o<true, KB * 32u + 16u, WL + 4u>(d_t, a, dw, idesc);
    mma_ts_o<true, KB * 32u + 24u, WL + 6u>(d_t, a, dw, idesc);
}

// one ring chunk of layer 4: BLK 0: hi kb0, 1: lo kb0, 2: hi kb1, 3: lo kb1 (128 output rows x 64 K); hi blocks take A_hi and A_lo
template <int BLK, int PREC>
__device__ __forceinline__ void ring_chunk(uint32_t d_t, uint32_t a, uint32_t b, uint32_t idesc) {
    using namespace tc;
    constexpr uint32_t KB = (uint32_t)(BLK >> 1) * 32u;
    mma_ts_o<(BLK != 0), KB + 0u, 0u>(d_t, a, b, idesc);
    mma_ts_o<true, KB + 8u, 2u>(d_t, a, b, idesc);
    mma_ts_o<true, KB + 16u, 4u>(d_t, a, b, idesc);
    mma_ts_o<true, KB + 24u, 6u>(d_t, a, b, idesc);
    if ((BLK & 1) == 0 && PREC != 2) {
        mma_ts_o<true, 64u + KB + 0u, 0u>(d_t, a, b, idesc);
        mma_ts_o<true, 64u + KB + 8u, 2u>(d_t, a, b, idesc);
        mma_ts_o<true, 64u + KB + 16u, 4u>(d_t, a, b, idesc);
        mma_ts_o<true, 64u + KB + 24u, 6u>(d_t, a, b, idesc);
    }
}

// one layer of one slot; returns false when the slot has no further tile (its workers have been released)
template <bool FINE, int SLOT, int PREC>
__device__ __forceinline__ bool mlp_serve(MlpIssue &s, uint32_t &par, uint32_t &layer, const MlpParams &p) {
    using namespace tc;
    constexpr uint32_t L = FINE ? 4 : 3;
    constexpr uint32_t NBUF = FINE ? 1 : 2;
    constexpr uint32_t idesc = PREC == 2 ? make_idesc_f16(128, 128) : make_idesc_bf16(128, 128);
    mbar_wait_a(s.bars + 8u * (MLP_BAR_A_READY + SLOT), par);
    par ^= 1u;
    const uint32_t l = layer;
    layer = l + 1u == L ? 0u : l + 1u;
    const uint32_t d_t = s.tbase + 256u * SLOT, a = d_t + 128u;
    if (l == 0) {
        // layer 0 of the CTA's next tile (sequence number nseq): A = interpolated features staged in shared memory by the
        // gather warps (hi block, lo block 16 KB further); resident image: L1 hi at 0, lo at 16 KB
        uint32_t tile = MLP_NO_TILE;
        const uint32_t b = NBUF == 2 ? (s.nseq & 1u) : 0u;
        if (!s.done) {
            mbar_wait_a(s.bars + 8u * (MLP_BAR_A0_FULL + b), (s.nseq / NBUF) & 1u);
            asm volatile("ld.volatile.shared.u32 %0, [%1];" : "=r"(tile) : "r"(s.bars + MLP_TILE_IDS_OFF + 4u * (s.nseq & 7u)) : "memory");
            s.nseq++;
        }
        if (tile == MLP_NO_TILE) {  // out of tiles: release the slot's workers (they read the same sentinel) and retire the slot
            s.done = 1u;
            asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(s.bars + 8u * (MLP_BAR_D_READY + SLOT)) : "memory");
            return false;
        }
        fence_after_sync();
        tl_mark(p.timeline, 0, 17, 1 + SLOT, 0, l);
        const uint32_t da = b ? s.dr : s.da0;  // COARSE: the second operand buffer lives in the (unused) ring area
        mma_ss_o<false, 0u, 0u>(d_t, da, s.dw, idesc);
        mma_ss_o<true, 2u, 2u>(d_t, da, s.dw, idesc);
        mma_ss_o<true, 4u, 4u>(d_t, da, s.dw, idesc);
        mma_ss_o<true, 6u, 6u>(d_t, da, s.dw, idesc);
        if (PREC != 2) {  // (A0_lo, W_hi): the lo block of the operand is 16 KB further
            mma_ss_o<true, 1024u + 0u, 0u>(d_t, da, s.dw, idesc);
            mma_ss_o<true, 1024u + 2u, 2u>(d_t, da, s.dw, idesc);
            mma_ss_o<true, 1024u + 4u, 4u>(d_t, da, s.dw, idesc);
            mma_ss_o<true, 1024u + 6u, 6u>(d_t, da, s.dw, idesc);
        }
        mma_ss_o<true, 0u, 1024u + 0u>(d_t, da, s.dw, idesc);
        mma_ss_o<true, 2u, 1024u + 2u>(d_t, da, s.dw, idesc);
        mma_ss_o<true, 4u, 1024u + 4u>(d_t, da, s.dw, idesc);
        mma_ss_o<true, 6u, 1024u + 6u>(d_t, da, s.dw, idesc);
        mma_commit_a(s.bars + 8u * (MLP_BAR_A0_EMPTY + b));
    } else {
        fence_after_sync();
        tl_mark(p.timeline, 0, 17, 1 + SLOT, 0, l);
        if (l == 1) {
            kblock_ts<true, 0u, 2048u, 3072u, PREC>(d_t, a, s.dw, idesc);
            kblock_ts<false, 1u, 4096u, 5120u, PREC>(d_t, a, s.dw, idesc);
        } else if (l == 2) {
            kblock_ts<true, 0u, 6144u, 7168u, PREC>(d_t, a, s.dw, idesc);
            kblock_ts<false, 1u, 8192u, 9216u, PREC>(d_t, a, s.dw, idesc);
        } else {
            // layer 4: the chunks hi(kb0) lo(kb0) hi(kb1) lo(kb1) stream through the ring
#define TN_RING_CHUNK(J)                                                                              \
    {                                                                                                 \
        mbar_wait_a(s.bars + 8u * (MLP_BAR_RING_FULL + s.rst), s.rpar);                               \
        ring_chunk<(J), PREC>(d_t, a, s.dr + s.rst * (MLP_RING_CHUNK >> 4), idesc);                   \
        mma_commit_a(s.bars + 8u * (MLP_BAR_RING_EMPTY + s.rst));                                     \
        if (++s.rst == mlp_ring_stages(PREC)) { s.rst = 0; s.rpar ^= 1u; }                            \
    }
            TN_RING_CHUNK(0) TN_RING_CHUNK(1) TN_RING_CHUNK(2) TN_RING_CHUNK(3)
#undef TN_RING_CHUNK
        }
    }
    mma_commit_a(s.bars + 8u * (MLP_BAR_D_READY + SLOT));
    tl_mark(p.timeline, 0, 17, 3 + SLOT, 0, l);
    return true;
}

template <bool FINE, int PREC>
__global__ void __launch_bounds__(MLP_THREADS, 1) k_mlp(const MlpParams p) {
    constexpr uint32_t RS = mlp_ring_stages(PREC);
    constexpr uint32_t A0_OFF = MLP_OFF_A0 + (PREC == 2 ? 16384u : 0u);  // f16w2: the first 16 KB of the operand area are ring stage 2
    using namespace tc;
    uint8_t *smem = tn_mlp_smem;
    uint8_t *w_s = smem;
    uint8_t *ring_s = smem + MLP_OFF_RING;
    float *head_s = reinterpret_cast<float *>(smem + MLP_OFF_HEAD);
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem + MLP_OFF_BARS);
    uint64_t *a_ready = bars + MLP_BAR_A_READY;        // [2] count 8: the slot's warps are done with D / have written the next A
    uint64_t *d_ready = bars + MLP_BAR_D_READY;        // [2] count 1 (tcgen05.commit; plain arrive when the slot is retired)
    uint64_t *w_bar = bars + MLP_BAR_W;                // resident weights landed
    uint64_t *ring_full = bars + MLP_BAR_RING_FULL;    // [3]
    uint64_t *ring_empty = bars + MLP_BAR_RING_EMPTY;  // [3]
    uint64_t *a0_full = bars + MLP_BAR_A0_FULL;        // [2] count 4: every gather warp has stored its rows of the layer-0 operand
    uint64_t *a0_empty = bars + MLP_BAR_A0_EMPTY;      // [2] count 1 (tcgen05.commit after the layer-0 MMAs)
    uint32_t *tmem_ptr = reinterpret_cast<uint32_t *>(bars + 15);
    volatile uint32_t *stop_flag = tmem_ptr + 1;       // set by the issuer when both slots are retired (stops the ring producer)
    volatile uint32_t *tile_ids = reinterpret_cast<volatile uint32_t *>(smem + MLP_OFF_BARS + MLP_TILE_IDS_OFF);  // [8] tile of sequence number n at n & 7

    constexpr int L = FINE ? 4 : 3;
    constexpr uint32_t NBUF = FINE ? 1 : 2;  // layer-0 operand buffers (COARSE has the ring area to spare)
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t n_active = *p.n_active;
    const uint64_t total_rows = (uint64_t)n_active * p.S;
    const uint32_t ntiles = (uint32_t)((total_rows + 127) / 128);
    // Tiles are handed out dynamically (the SMs do not run at the same speed: L2 distance shows up in every latency-bound
    // phase): a CTA's first tile is blockIdx.x, the following ones come from a global counter.
    const bool has_work = blockIdx.x < ntiles;

    if (warp == 16) {
        if (lane == 0) {
            mbar_init(&a_ready[0], 8); mbar_init(&a_ready[1], 8);
            mbar_init(&d_ready[0], 1); mbar_init(&d_ready[1], 1);
            mbar_init(w_bar, 1);
            for (int i = 0; i < 3; ++i) { mbar_init(&ring_full[i], 1); mbar_init(&ring_empty[i], 1); }
            for (int i = 0; i < 2; ++i) { mbar_init(&a0_full[i], MLP_GATHER_WARPS); mbar_init(&a0_empty[i], 1); }
            fence_barrier_init();
            *stop_flag = 0u;
            tile_ids[0] = has_work ? blockIdx.x : MLP_NO_TILE;
        }
        __syncwarp();
        tmem_alloc(tmem_ptr, 512);
    }
    if (threadIdx.x < 2) reinterpret_cast<unsigned long long *>(smem + MLP_OFF_TL)[threadIdx.x] = 0ull;
    if (p.timeline != nullptr && threadIdx.x == 0) {  // debug: per-CTA start time (ns)
        unsigned long long t;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
        p.timeline[1000 + 8 * blockIdx.x] = t;
        uint32_t smid;
        asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
        p.timeline[1000 + 8 * blockIdx.x + 4] = smid;
        p.timeline[1000 + 8 * blockIdx.x + 7] = (unsigned long long)clock64();
    }
    for (uint32_t i = threadIdx.x; i < 516; i += MLP_THREADS) head_s[i] = p.head[i];
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
    const uint32_t tbase = *tmem_ptr;

    if (warp >= 16 && warp < 20) {
        reg_dec<MLP_REGS_CTRL>();
        if (warp == 16) {
            // ================= TMA: resident weights, then the L4 ring =================
            if (lane == 0 && has_work) {
                mbar_arrive_expect_tx(w_bar, MLP_W_RESIDENT);
                for (uint32_t off = 0; off < MLP_W_RESIDENT; off += 16384) tma_bulk_g2s(w_s + off, p.wimg + off, 16384, w_bar);
                if (FINE) {
                    // keeps both stages filled (chunk i of the endless sequence hi0 lo0 hi1 lo1 ...) until the issuer says stop;
                    // the one or two chunks fetched beyond the last tile are awaited before the CTA exits
                    const uint8_t *w4 = p.wimg + MLP_W_RESIDENT;
                    uint32_t st = 0, par = 1, i = 0;  // first round: the stages are empty (wait on the preceding phase passes)
                    for (;; ++i) {
                        bool stop = false;
                        while (!mbar_test(&ring_empty[st], par)) {
                            if (*stop_flag != 0u) { stop = true; break; }
                            __nanosleep(64);
                        }
                        if (stop) break;
                        mbar_arrive_expect_tx(&ring_full[st], MLP_RING_CHUNK);
                        tma_bulk_g2s(ring_s + st * MLP_RING_CHUNK, w4 + (i & 3u) * MLP_RING_CHUNK, MLP_RING_CHUNK, &ring_full[st]);
                        if (++st == RS) { st = 0; par ^= 1u; }
                    }
                    for (uint32_t q = 0; q < RS; ++q) {  // fills of stage q among chunks 0..i-1; phase f of ring_full = fill f
                        const uint32_t fills = (i + RS - 1u - q) / RS;
                        if (fills) mbar_wait_backoff(&ring_full[q], (fills - 1u) & 1u, 64);
                    }
                }
            }
        } else if (warp == 17) {
            // ================= MMA issuer =================
            if (lane == 0 && has_work) {
                MlpIssue s;
                s.tbase = tbase;
                s.bars = smem_u32(bars);
                s.dw = desc_lo(smem_u32(w_s));
                s.dr = desc_lo(smem_u32(ring_s));
                s.da0 = desc_lo(smem_u32(smem + A0_OFF));
                s.rst = 0; s.rpar = 0; s.nseq = 0; s.done = 0;
                uint32_t par0 = 0, par1 = 0, layer0 = 0, layer1 = 0;
                mbar_wait(w_bar, 0);
                // strict alternation between the slots; slot 1 starts one round late so that the layer-0 passes (the consumers
                // of the layer-0 operand buffer) are evenly spaced and each slot's epilogue runs under the other slot's MMAs
                bool alive0 = mlp_serve<FINE, 0, PREC>(s, par0, layer0, p), alive1 = true;
                while (alive0 | alive1) {
                    if (alive0) alive0 = mlp_serve<FINE, 0, PREC>(s, par0, layer0, p);
                    if (alive1) alive1 = mlp_serve<FINE, 1, PREC>(s, par1, layer1, p);
                }
                *stop_flag = 1u;
            }
        }
    } else if (warp >= (int)MLP_GATHER_WARP0) {
        // ================= gather warps: interpolated features -> layer-0 A operand in shared memory =================
        // warp g owns tile rows 32g..32g+31: 16 steps of two rows (one per half-warp, lane l16 -> features 4*l16..+3);
        // lane i keeps the matched vertex ids / weights of row i; the vertex-row loads run 4 steps (16 x 16 B per lane)
        // ahead of the interpolation, across tile boundaries.  Lane 0 of the first gather warp is the CTA's tile
        // scheduler: it draws tile numbers from the global counter two tiles ahead and publishes them in tile_ids.
        reg_inc<MLP_REGS_GATHER>();
        if (has_work) {
        const uint32_t g = (uint32_t)warp - MLP_GATHER_WARP0;
        const bool sched = warp == (int)MLP_GATHER_WARP0 && lane == 0;
        const uint32_t hw = (uint32_t)lane >> 4, l16 = (uint32_t)lane & 15u;
        const float4 *fs = reinterpret_cast<const float4 *>(p.fshadow) + l16;  // a field row is 16 float4
        uint4 cv, nv;
        float cb0, cb1, cb2, nb0, nb1, nb2;
#define TN_LOAD_IDS(TILE, V, B0, B1, B2)                                                                            \
    {                                                                                                                \
        V = make_uint4(TN_EMPTY, TN_EMPTY, TN_EMPTY, TN_EMPTY);                                                      \
        B0 = 0.f; B1 = 0.f; B2 = 0.f;                                                                                \
        if ((TILE) != MLP_NO_TILE) {                                                                                 \
            const uint64_t gr = (uint64_t)(TILE) * 128u + g * 32u + (uint32_t)lane;                                  \
            if (gr < total_rows) {                                                                                   \
                V = __ldg(p.vi + gr);                                                                                \
                B0 = __ldg(p.bary + 3 * gr); B1 = __ldg(p.bary + 3 * gr + 1); B2 = __ldg(p.bary + 3 * gr + 2);       \
            }                                                                                                        \
        }                                                                                                            \
    }
        float4 f[4][4];  // [pipeline buffer][vertex]
#define TN_ISSUE(BUF, V, C)                                                                                          \
    {                                                                                                                \
        const int r_ = 2 * (C) + (int)hw;                                                                            \
        const uint32_t v0 = __shfl_sync(0xffffffffu, V.x, r_), v1 = __shfl_sync(0xffffffffu, V.y, r_);               \
        const uint32_t v2 = __shfl_sync(0xffffffffu, V.z, r_), v3 = __shfl_sync(0xffffffffu, V.w, r_);               \
        const bool m_ = TN_MLP_GATHER_ON && v0 != TN_EMPTY;                                                          \
        const float4 z_ = make_float4(0.f, 0.f, 0.f, 0.f);                                                           \
        f[BUF][0] = m_ ? ldg_stream(fs + (size_t)v0 * 16) : z_;                                                      \
        f[BUF][1] = m_ ? ldg_stream(fs + (size_t)v1 * 16) : z_;                                                      \
        f[BUF][2] = m_ ? ldg_stream(fs + (size_t)v2 * 16) : z_;                                                      \
        f[BUF][3] = m_ ? ldg_stream(fs + (size_t)v3 * 16) : z_;                                                      \
    }
        uint32_t cur = blockIdx.x, nxt = MLP_NO_TILE, pending = MLP_NO_TILE;
        if (sched) pending = gridDim.x + atomicAdd(p.tile_ctr, 1u);  // tile of sequence number 1 (read one tile later)
        TN_LOAD_IDS(cur, cv, cb0, cb1, cb2)
        TN_ISSUE(0, cv, 0) TN_ISSUE(1, cv, 1) TN_ISSUE(2, cv, 2) TN_ISSUE(3, cv, 3)
#if TN_MLP_STATS
        long long dbg_wait = 0, dbg_busy = 0, dbg_t = 0;
#endif
        uint32_t n = 0;
        for (; cur != MLP_NO_TILE; ++n) {
            if (sched) {  // publish the tile of sequence number n+1, draw the one of n+2
                const uint32_t v = pending < ntiles ? pending : MLP_NO_TILE;
                tile_ids[(n + 1u) & 7u] = v;
                if (v == MLP_NO_TILE) tile_ids[(n + 2u) & 7u] = MLP_NO_TILE;  // both slots must see the end
                pending = v != MLP_NO_TILE ? gridDim.x + atomicAdd(p.tile_ctr, 1u) : MLP_NO_TILE;
            }
            asm volatile("bar.sync 3, 128;" ::: "memory");
            nxt = tile_ids[(n + 1u) & 7u];
            TN_LOAD_IDS(nxt, nv, nb0, nb1, nb2)
            const uint32_t b = NBUF == 2 ? (n & 1u) : 0u;
            const uint32_t a0 = smem_u32(smem + (b ? MLP_OFF_RING : A0_OFF));
#if TN_MLP_STATS
            if (p.timeline != nullptr) dbg_t = clock64();
#endif
            if (n >= NBUF) mbar_wait_backoff(&a0_empty[b], (n / NBUF - 1u) & 1u, 64);  // the layer-0 MMAs of the tile that used this buffer are done
#if TN_MLP_STATS
            if (p.timeline != nullptr) { const long long t = clock64(); dbg_wait += t - dbg_t; dbg_t = t; }
#endif
            if (p.timeline != nullptr && blockIdx.x == 0 && sched && n == 8)
                reinterpret_cast<volatile unsigned long long *>(smem + MLP_OFF_TL)[1] = 1ull;  // steady state reached: start recording
            tl_mark(p.timeline, lane, warp, 5, n, 0);  // ev 5: gather stores start
            {
#pragma unroll
                for (int cc = 0; cc < 16; cc += 2) {
#pragma unroll
                    for (int c = cc; c < cc + 2; ++c) {
                        const int r = 2 * c + (int)hw;
                        const float b0 = __shfl_sync(0xffffffffu, cb0, r), b1 = __shfl_sync(0xffffffffu, cb1, r), b2 = __shfl_sync(0xffffffffu, cb2, r);
                        const float4 *fc = f[c & 3];
                        const float w0 = __fsub_rn(1.0f, __fadd_rn(__fadd_rn(b0, b1), b2));
                        float4 o;  // tetrahedra_tracer.cu:203-220: v1*b0, + v2*b1, + v3*b2, + v0*w0 (fused multiply-adds)
                        o.x = __fmaf_rn(b0, fc[1].x, 0.f); o.y = __fmaf_rn(b0, fc[1].y, 0.f); o.z = __fmaf_rn(b0, fc[1].z, 0.f); o.w = __fmaf_rn(b0, fc[1].w, 0.f);
                        o.x = __fmaf_rn(b1, fc[2].x, o.x); o.y = __fmaf_rn(b1, fc[2].y, o.y); o.z = __fmaf_rn(b1, fc[2].z, o.z); o.w = __fmaf_rn(b1, fc[2].w, o.w);
                        o.x = __fmaf_rn(b2, fc[3].x, o.x); o.y = __fmaf_rn(b2, fc[3].y, o.y); o.z = __fmaf_rn(b2, fc[3].z, o.z); o.w = __fmaf_rn(b2, fc[3].w, o.w);
                        o.x = __fmaf_rn(w0, fc[0].x, o.x); o.y = __fmaf_rn(w0, fc[0].y, o.y); o.z = __fmaf_rn(w0, fc[0].z, o.z); o.w = __fmaf_rn(w0, fc[0].w, o.w);
                        uint32_t h0, l0 = 0, h1, l1 = 0;
                        if (PREC == 2) { h0 = pack2_f16(o.x, o.y); h1 = pack2_f16(o.z, o.w); }
                        else { split_pack2(o.x, o.y, h0, l0); split_pack2(o.z, o.w, h1, l1); }
                        // tile row R = 32 g + 2 c + hw inside the swizzled [128][64] bf16 block
                        const uint32_t R = g * 32u + (uint32_t)r, r7 = R & 7u;
                        const uint32_t addr = a0 + (R >> 3) * 1024u + r7 * 128u + (((l16 >> 1) ^ r7) << 4) + (l16 & 1u) * 8u;
                        asm volatile("st.shared.v2.b32 [%0], {%1, %2};" ::"r"(addr), "r"(h0), "r"(h1) : "memory");
                        if (PREC != 2) asm volatile("st.shared.v2.b32 [%0], {%1, %2};" ::"r"(addr + 16384u), "r"(l0), "r"(l1) : "memory");
                    }
                    // refill the two pipeline buffers just consumed: 4 steps ahead, running into the next tile at the end of this one
                    if (cc + 4 < 16) { TN_ISSUE(cc & 3, cv, cc + 4) TN_ISSUE((cc + 1) & 3, cv, cc + 5) }
                    else { TN_ISSUE(cc & 3, nv, cc + 4 - 16) TN_ISSUE((cc + 1) & 3, nv, cc + 5 - 16) }
                }
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy stores -> visible to the MMA (async proxy)
            __syncwarp();
            if (lane == 0) mbar_arrive(&a0_full[b]);
            tl_mark(p.timeline, lane, warp, 6, n, 0);  // ev 6: operand stored, arrived
#if TN_MLP_STATS
            if (p.timeline != nullptr) dbg_busy += clock64() - dbg_t;
#endif
            cur = nxt; cv = nv; cb0 = nb0; cb1 = nb1; cb2 = nb2;
        }
        {   // sequence number n is the first one without a tile: complete its "operand ready" phase (after the buffer's
            // previous consumer, like a real tile) so that the issuer wakes up, reads the sentinel and retires the slots
            const uint32_t b = NBUF == 2 ? (n & 1u) : 0u;
            if (n >= NBUF) mbar_wait_backoff(&a0_empty[b], (n / NBUF - 1u) & 1u, 64);
            __syncwarp();
            if (lane == 0) mbar_arrive(&a0_full[b]);
        }
#if TN_MLP_STATS
        if (p.timeline != nullptr && sched) {
            p.timeline[1000 + 8 * blockIdx.x + 5] = (unsigned long long)dbg_busy;
            p.timeline[1000 + 8 * blockIdx.x + 6] = (unsigned long long)dbg_wait;
        }
#endif
#undef TN_ISSUE
#undef TN_LOAD_IDS
        }
    } else {
        // ================= slot warps: per-layer epilogues, heads =================
        reg_inc<MLP_REGS_WORKER>();
        const int slot = warp >> 3;
        const uint32_t h = (uint32_t)(warp >> 2) & 1u;  // column half
        const uint32_t q = (uint32_t)warp & 3u;         // TMEM lane quarter this warp may access
        const uint32_t lane_base = (q * 32u) << 16;
        const uint32_t d_t = tbase + 256u * slot + lane_base, ahi = d_t + 128u, alo = d_t + 192u;
        const float *wd = head_s, *wc = head_s + 128;
        uint32_t dpar = 0, ntl = 0;
        if (has_work) {
        if (lane == 0) mbar_arrive(&a_ready[slot]);  // the slot's accumulator is free for the first tile
        for (uint32_t n = slot;; n += 2) {            // the slot handles the CTA's tiles of sequence number n = slot, slot + 2, ...
            uint32_t tile = MLP_NO_TILE;
            uint64_t my_row = 0;
            const float *bias4 = nullptr;
            float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
            float2 bpre[8];
            if (!(PREC == 2 && TN_MLP_F16_WIDE)) bias_prefetch(p.bias, h, bpre);  // layer 0: known before the tile is
#pragma unroll 1
            for (int l = 0; l < L; ++l) {
                tl_mark(p.timeline, lane, warp, 7, n, l);  // ev 7: start waiting for D
                mbar_wait_backoff(&d_ready[slot], dpar, 32);  // (0 / 8 / 32 / 128 ns measured: no difference)
                dpar ^= 1u;
                fence_after_sync();
                tl_mark(p.timeline, lane, warp, 8, n, l);  // ev 8: D ready seen
                if (l == 0) {
                    tile = tile_ids[n & 7u];
                    if (tile == MLP_NO_TILE) break;  // retired by the issuer
                    ++ntl;
                    my_row = (uint64_t)tile * 128u + q * 32u + (uint32_t)lane;
                }
                if (l < L - 1) {
                    if (PREC == 2 && TN_MLP_F16_WIDE) layer_epilogue_f16<0>(d_t, ahi, h, p.bias + l * 128, FINE && l == 2, wd, wc, acc);
                    else {
                        layer_epilogue<0, PREC>(d_t, ahi, alo, h, p.bias + l * 128, bpre, FINE && l == 2, wd, wc, acc);
                        // the next layer's first bias chunk: in flight while the next GEMM runs
                        if (FINE && l == 2) bias_prefetch(bias4, h, bpre); else bias_prefetch(p.bias + (l + 1) * 128, h, bpre);
                    }
                } else if (FINE) {
                    if (PREC == 2 && TN_MLP_F16_WIDE) layer_epilogue_f16<3>(d_t, ahi, h, bias4, false, wd, wc, acc);
                    else layer_epilogue<3, PREC>(d_t, ahi, alo, h, bias4, bpre, false, wd, wc, acc);
                } else {
                    if (PREC == 2 && TN_MLP_F16_WIDE) layer_epilogue_f16<2>(d_t, ahi, h, p.bias + 256, false, wd, wc, acc);
                    else layer_epilogue<2, PREC>(d_t, ahi, alo, h, p.bias + 256, bpre, false, wd, wc, acc);
                }
                // next A operand written (l < L-1) / accumulator read out and free for the next tile (l == L-1)
                if (l < L - 1) tmem_st_wait();
                fence_before_sync();
                __syncwarp();
                if (lane == 0) mbar_arrive(&a_ready[slot]);
                // per-ray direction bias of this thread's row (first needed after layer 3): the division is kept off the layer-1 epilogue,
                // which sits on the slot's critical chain (the launcher guarantees rows < 2^32)
                if (FINE && l == 0) bias4 = p.dirbias + (size_t)((uint32_t)min(my_row, total_rows - 1) / p.S) * 128;
                tl_mark(p.timeline, lane, warp, 9, n, l);  // ev 9: epilogue of layer l done (+arrive)
            }
            if (tile == MLP_NO_TILE) break;
            // ---- heads: combine the two column halves, activation, store ----
            // The partial sums of the upper column half travel through TMEM: the first four columns of the slot's A region
            // (dead once the last layer's MMAs are done; next written by this lane quarter's h == 0 warp itself, in the
            // next tile's first epilogue).
            if (h == 1) {
                tmem_st4(ahi, reinterpret_cast<const uint32_t *>(&acc));
                tmem_st_wait();
                fence_before_sync();
            }
            asm volatile("bar.sync %0, 256;" ::"r"(1 + slot) : "memory");
            if (h == 0) {
                fence_after_sync();
                float4 o;
                tmem_ld4(ahi, reinterpret_cast<uint32_t *>(&o));
                tmem_ld_wait();
                if (my_row < total_rows) {
                    const float sigma = softplus_f(acc.x + o.x + head_s[512]);
                    if (FINE) {
                        reinterpret_cast<float4 *>(p.out)[my_row] = make_float4(sigma, sigmoid_f(acc.y + o.y + head_s[513]),
                                                                               sigmoid_f(acc.z + o.z + head_s[514]), sigmoid_f(acc.w + o.w + head_s[515]));
                    } else {
                        p.out[my_row] = sigma;
                    }
                }
            }
        }
        }
        if (p.timeline != nullptr && threadIdx.x == 0) p.timeline[1000 + 8 * blockIdx.x + 3] = ntl;
    }
    fence_before_sync();
    __syncthreads();
    if (warp == 16) tmem_dealloc(tbase, 512);
    if (p.timeline != nullptr && threadIdx.x == 0) {
        unsigned long long t;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
        p.timeline[1000 + 8 * blockIdx.x + 1] = t;
        p.timeline[1000 + 8 * blockIdx.x + 7] = (unsigned long long)clock64() - p.timeline[1000 + 8 * blockIdx.x + 7];
    }
    if (p.timeline != nullptr && blockIdx.x == 0) {
        const unsigned long long *tl = reinterpret_cast<const unsigned long long *>(smem + MLP_OFF_TL);
        const uint32_t n = (uint32_t)min(tl[0], (unsigned long long)MLP_TL_CAP);
        for (uint32_t i = threadIdx.x; i <= n; i += MLP_THREADS) p.timeline[i] = i == 0 ? (unsigned long long)n : tl[i + 1];
    }
}

}  // namespace tn
