// tn_mlp.cuh -- the fused interpolate -> MLP kernel on tcgen05 (sm_100a).
//
// Replaces, for one pass over n_active*S samples, the chain
//   interpolate_values (src/tetrahedra_tracer.cu:195-221)  ->  mlp_base (3x Linear+ReLU, model.py:433-438)
//   -> DensityFieldHead (Linear+Softplus, :455) [-> mlp_head (Linear+ReLU, :447-452) -> RGBFieldHead
//   (Linear+Sigmoid, :454)]                                               (call sites model.py:569-621)
// which the reference runs as ~10 torch kernels with [R*S,128] fp32 activations round-tripping HBM.
//
// One persistent CTA per SM, 18 warps:
//   warp 17     : issues every tcgen05.mma (one elected lane), strictly alternating between the two slots; it has the
//                 highest warp id because the SM's warp arbiter is highest-id-first and the issuer must never starve
//   warp 16     : TMEM allocation, TMA bulk staging of the resident weight image, and (FINE) the
//                 producer of the 2-stage TMA ring that streams the 4th layer's weights
//   warps 0..15 : two "slots" of 8 warps.  A slot owns one 128-sample tile at a time; warp (q,h) of a
//                 slot owns sample rows 32q..32q+31 (its TMEM lane quarter, q = warp_id % 4) and the
//                 column half h.  It gathers the four vertex rows of its samples from the [V,64] field
//                 shadow with coalesced 128-byte reads, forms the barycentric interpolation with the
//                 reference's FMA order, splits it into bf16 hi/lo and writes it into TMEM as the A
//                 operand; then for each layer it waits for the accumulator, applies bias + ReLU to its
//                 64 columns, splits again and writes the next A operand back into TMEM (activations
//                 never touch shared or global memory).  The two slots run half a tile apart, so one
//                 slot's epilogue overlaps the other slot's MMAs.
// Products are "bf16x3": a*w ~= a_hi*w_hi + a_lo*w_hi + a_hi*w_lo with fp32 accumulation in TMEM
// (measured 5e-6 relative on B200, tests/test_gpu_mlp.py) -- the reference computes in fp32 and the
// parity bar is 1e-4 absolute on colour/density, which single-pass bf16/tf32 cannot hold.
// TMEM (512 columns): slot s uses [256s, 256s+128) for D, [+128,+192) A_hi, [+192,+256) A_lo.
#pragma once
#include "tn_common.cuh"
#include "tn_tc.cuh"

namespace tn {

constexpr uint32_t MLP_THREADS = 576;
constexpr uint32_t MLP_TL_CAP = 960;  // debug timeline records buffered in shared memory
constexpr uint32_t MLP_W_RESIDENT = 163840;               // L1 32K + L2 64K + L3 64K
constexpr uint32_t MLP_OFF_RING = MLP_W_RESIDENT;          // 2 x 16K
constexpr uint32_t MLP_OFF_STAGE = MLP_OFF_RING + 32768;   // 16 warps x 4 rows x 36 floats
constexpr uint32_t MLP_STAGE_STRIDE = 36;                  // floats per staged sample row (32 features + pad)
constexpr uint32_t MLP_OFF_BIAS = MLP_OFF_STAGE + 16 * 4 * MLP_STAGE_STRIDE * 4;  // 3 x 128 floats (b1,b2,b3)
constexpr uint32_t MLP_OFF_HEAD = MLP_OFF_BIAS + 3 * 128 * 4;                     // wd[128] wc[3][128] bd bc[3] (+pad)
constexpr uint32_t MLP_OFF_DIRB = MLP_OFF_HEAD + 520 * 4;                         // 2 slots x 4 rays x 128 floats
constexpr uint32_t MLP_OFF_RED = MLP_OFF_DIRB + 2 * 4 * 128 * 4;                  // 2 slots x 128 rows x float4 (head partials)
constexpr uint32_t MLP_OFF_BARS = MLP_OFF_RED + 2 * 128 * 16;
constexpr uint32_t MLP_OFF_TL = MLP_OFF_BARS + 128;  // debug timeline: counter + MLP_TL_CAP records
constexpr uint32_t MLP_SMEM_BYTES = MLP_OFF_TL + 8 * (MLP_TL_CAP + 1);
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
//   KIND 0: hidden layer  -> bias + ReLU, next A operand
//   KIND 1: 3rd layer     -> as 0, plus the density-head partial dot product (acc.x)
//   KIND 2: 3rd layer, last (COARSE) -> bias + ReLU + density partial, no next operand
//   KIND 3: 4th layer (FINE, last)   -> per-ray bias + ReLU + colour-head partial (acc.y/z/w)
template <int KIND>
__device__ __forceinline__ void layer_epilogue(uint32_t d_t, uint32_t ahi, uint32_t alo, uint32_t h, const float *__restrict__ bias128,
                                               const float *__restrict__ wd, const float *__restrict__ wc, float4 &acc) {
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
            const float2 b = *reinterpret_cast<const float2 *>(bias128 + col0 + 2 * i);
            float2 x = add2(make_float2(__uint_as_float(r[2 * i]), __uint_as_float(r[2 * i + 1])), b);
            x.x = fmaxf(x.x, 0.f);
            x.y = fmaxf(x.y, 0.f);
            if (KIND == 0 || KIND == 1) split2(x, ph[i], pl[i]);
            if (KIND == 1 || KIND == 2) dsum = fma2(x, *reinterpret_cast<const float2 *>(wd + col0 + 2 * i), dsum);
            if (KIND == 3) {
                c0 = fma2(x, *reinterpret_cast<const float2 *>(wc + col0 + 2 * i), c0);
                c1 = fma2(x, *reinterpret_cast<const float2 *>(wc + 128 + col0 + 2 * i), c1);
                c2 = fma2(x, *reinterpret_cast<const float2 *>(wc + 256 + col0 + 2 * i), c2);
            }
        }
        if (KIND == 0 || KIND == 1) {
            tmem_st8(ahi + (col0 >> 1), ph);
            tmem_st8(alo + (col0 >> 1), pl);
        }
    }
    if (KIND == 1 || KIND == 2) acc.x = dsum.x + dsum.y;
    if (KIND == 3) { acc.y = c0.x + c0.y; acc.z = c1.x + c1.y; acc.w = c2.x + c2.y; }
}

extern __shared__ __align__(1024) uint8_t tn_mlp_smem[];
__device__ __forceinline__ void tl_mark(unsigned long long *buf, int lane, uint32_t warp, uint32_t ev, uint32_t it, uint32_t layer) {
    if (buf != nullptr && blockIdx.x == 0 && lane == 0) {
        unsigned long long *tl = reinterpret_cast<unsigned long long *>(tn_mlp_smem + MLP_OFF_TL);
        const unsigned long long i = atomicAdd(tl, 1ull);
        if (i < MLP_TL_CAP) tl[1 + i] = ((unsigned long long)(warp | (ev << 5) | ((it & 255u) << 9) | (layer << 17)) << 40) | ((unsigned long long)clock64() & 0xFFFFFFFFFFull);
    }
}

template <bool FINE>
__global__ void __launch_bounds__(MLP_THREADS, 1) k_mlp(const MlpParams p) {
    using namespace tc;
    uint8_t *smem = tn_mlp_smem;
    uint8_t *w_s = smem;
    uint8_t *ring_s = smem + MLP_OFF_RING;
    float *bias_s = reinterpret_cast<float *>(smem + MLP_OFF_BIAS);
    float *head_s = reinterpret_cast<float *>(smem + MLP_OFF_HEAD);
    float *dirb_s = reinterpret_cast<float *>(smem + MLP_OFF_DIRB);
    float4 *red_s = reinterpret_cast<float4 *>(smem + MLP_OFF_RED);
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem + MLP_OFF_BARS);
    uint64_t *a_ready = bars;          // [2] count 8 (one arrive per slot warp)
    uint64_t *d_ready = bars + 2;      // [2] count 1 (tcgen05.commit)
    uint64_t *w_bar = bars + 4;        // resident weights landed
    uint64_t *ring_full = bars + 5;    // [2]
    uint64_t *ring_empty = bars + 7;   // [2]
    uint32_t *tmem_ptr = reinterpret_cast<uint32_t *>(bars + 10);

    constexpr int L = FINE ? 4 : 3;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t n_active = *p.n_active;
    const uint64_t total_rows = (uint64_t)n_active * p.S;
    const uint32_t ntiles = (uint32_t)((total_rows + 127) / 128);
    const uint32_t my_tiles = blockIdx.x < ntiles ? (ntiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;

    if (warp == 16) {
        if (lane == 0) {
            mbar_init(&a_ready[0], 8); mbar_init(&a_ready[1], 8);
            mbar_init(&d_ready[0], 1); mbar_init(&d_ready[1], 1);
            mbar_init(w_bar, 1);
            mbar_init(&ring_full[0], 1); mbar_init(&ring_full[1], 1);
            mbar_init(&ring_empty[0], 1); mbar_init(&ring_empty[1], 1);
            fence_barrier_init();
        }
        __syncwarp();
        tmem_alloc(tmem_ptr, 512);
    }
    if (threadIdx.x == 0) *reinterpret_cast<unsigned long long *>(smem + MLP_OFF_TL) = 0ull;
    for (uint32_t i = threadIdx.x; i < 3 * 128; i += MLP_THREADS) bias_s[i] = p.bias[i];
    for (uint32_t i = threadIdx.x; i < 516; i += MLP_THREADS) head_s[i] = p.head[i];
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
    const uint32_t tbase = *tmem_ptr;

    if (warp == 16) {
        // ================= TMA: resident weights, then the L4 ring =================
        if (lane == 0 && my_tiles > 0) {
            mbar_arrive_expect_tx(w_bar, MLP_W_RESIDENT);
            for (uint32_t off = 0; off < MLP_W_RESIDENT; off += 16384) tma_bulk_g2s(w_s + off, p.wimg + off, 16384, w_bar);
            if (FINE) {
                const uint8_t *w4 = p.wimg + MLP_W_RESIDENT;
                const uint32_t nchunks = my_tiles * 4;
                for (uint32_t i = 0; i < nchunks; ++i) {
                    const uint32_t st = i & 1u;
                    mbar_wait_backoff(&ring_empty[st], ((i >> 1) & 1u) ^ 1u, 128);
                    mbar_arrive_expect_tx(&ring_full[st], 16384);
                    tma_bulk_g2s(ring_s + st * 16384, w4 + (i & 3u) * 16384, 16384, &ring_full[st]);
                }
            }
        }
    } else if (warp == 17) {
        // ================= MMA issuer (highest warp id: the arbiter favours it over the epilogue warps) =================
        if (lane == 0 && my_tiles > 0) {
            const uint32_t idesc = make_idesc_bf16(128, 128);
            // scalar per-slot state (no dynamically indexed arrays: they would live in local memory and every
            // MMA issue would pay a local-memory round trip)
            uint32_t left0 = ((my_tiles + 1) / 2) * (uint32_t)L, left1 = (my_tiles / 2) * (uint32_t)L;
            uint32_t par0 = 0, par1 = 0, layer0 = 0, layer1 = 0, ring_par = 0;
            mbar_wait(w_bar, 0);
            // descriptor of byte offset `off` in the resident image = dw + (off >> 4); ring stage st = dr + st * 1024
            const uint64_t dw = make_desc_sw128(smem_u32(w_s)), dr = make_desc_sw128(smem_u32(ring_s));
            // one 64-wide K block: (A_hi,W_hi) (A_lo,W_hi) (A_hi,W_lo); descriptors differ by compile-time constants
#define TN_KBLOCK(FIRST, DHI, DLO, KB)                                                                      \
    mma_ts_c<!(FIRST)>(d_t, ahi + (KB) * 32u + 0u, (DHI) + 0ull, idesc);                                     \
    mma_ts_c<true>(d_t, ahi + (KB) * 32u + 8u, (DHI) + 2ull, idesc);                                         \
    mma_ts_c<true>(d_t, ahi + (KB) * 32u + 16u, (DHI) + 4ull, idesc);                                        \
    mma_ts_c<true>(d_t, ahi + (KB) * 32u + 24u, (DHI) + 6ull, idesc);                                        \
    mma_ts_c<true>(d_t, alo + (KB) * 32u + 0u, (DHI) + 0ull, idesc);                                         \
    mma_ts_c<true>(d_t, alo + (KB) * 32u + 8u, (DHI) + 2ull, idesc);                                         \
    mma_ts_c<true>(d_t, alo + (KB) * 32u + 16u, (DHI) + 4ull, idesc);                                        \
    mma_ts_c<true>(d_t, alo + (KB) * 32u + 24u, (DHI) + 6ull, idesc);                                        \
    mma_ts_c<true>(d_t, ahi + (KB) * 32u + 0u, (DLO) + 0ull, idesc);                                         \
    mma_ts_c<true>(d_t, ahi + (KB) * 32u + 8u, (DLO) + 2ull, idesc);                                         \
    mma_ts_c<true>(d_t, ahi + (KB) * 32u + 16u, (DLO) + 4ull, idesc);                                        \
    mma_ts_c<true>(d_t, ahi + (KB) * 32u + 24u, (DLO) + 6ull, idesc);
#define TN_SERVE(SLOT, LEFT, PAR, LAYER)                                                                     \
    if (LEFT) {                                                                                              \
        mbar_wait_backoff(&a_ready[SLOT], PAR, 32);                                                          \
        fence_after_sync();                                                                                  \
        const uint32_t l = LAYER;                                                                            \
        const uint32_t d_t = tbase + 256u * (SLOT), ahi = d_t + 128u, alo = d_t + 192u;                      \
        tl_mark(p.timeline, 0, 17, 1 + (SLOT), 0, l);                                                        \
        if (l == 0) {                                                                                        \
            TN_KBLOCK(true, dw, dw + 1024ull, 0u)                                                            \
        } else if (l == 1) {                                                                                 \
            TN_KBLOCK(true, dw + 2048ull, dw + 3072ull, 0u)                                                  \
            TN_KBLOCK(false, dw + 4096ull, dw + 5120ull, 1u)                                                 \
        } else if (l == 2) {                                                                                 \
            TN_KBLOCK(true, dw + 6144ull, dw + 7168ull, 0u)                                                  \
            TN_KBLOCK(false, dw + 8192ull, dw + 9216ull, 1u)                                                 \
        } else {                                                                                             \
            /* layer 4: chunks hi(kb0) lo(kb0) hi(kb1) lo(kb1) stream through the ring (hi -> stage 0, lo -> stage 1) */ \
            mbar_wait(&ring_full[0], ring_par);                                                              \
            mma_ts_c<false>(d_t, ahi + 0u, dr + 0ull, idesc); mma_ts_c<true>(d_t, ahi + 8u, dr + 2ull, idesc);   \
            mma_ts_c<true>(d_t, ahi + 16u, dr + 4ull, idesc); mma_ts_c<true>(d_t, ahi + 24u, dr + 6ull, idesc);  \
            mma_ts_c<true>(d_t, alo + 0u, dr + 0ull, idesc); mma_ts_c<true>(d_t, alo + 8u, dr + 2ull, idesc);    \
            mma_ts_c<true>(d_t, alo + 16u, dr + 4ull, idesc); mma_ts_c<true>(d_t, alo + 24u, dr + 6ull, idesc);  \
            mma_commit(&ring_empty[0]);                                                                      \
            mbar_wait(&ring_full[1], ring_par);                                                              \
            mma_ts_c<true>(d_t, ahi + 0u, dr + 1024ull, idesc); mma_ts_c<true>(d_t, ahi + 8u, dr + 1026ull, idesc);   \
            mma_ts_c<true>(d_t, ahi + 16u, dr + 1028ull, idesc); mma_ts_c<true>(d_t, ahi + 24u, dr + 1030ull, idesc); \
            mma_commit(&ring_empty[1]);                                                                      \
            ring_par ^= 1u;                                                                                  \
            mbar_wait(&ring_full[0], ring_par);                                                              \
            mma_ts_c<true>(d_t, ahi + 32u, dr + 0ull, idesc); mma_ts_c<true>(d_t, ahi + 40u, dr + 2ull, idesc);  \
            mma_ts_c<true>(d_t, ahi + 48u, dr + 4ull, idesc); mma_ts_c<true>(d_t, ahi + 56u, dr + 6ull, idesc);  \
            mma_ts_c<true>(d_t, alo + 32u, dr + 0ull, idesc); mma_ts_c<true>(d_t, alo + 40u, dr + 2ull, idesc);  \
            mma_ts_c<true>(d_t, alo + 48u, dr + 4ull, idesc); mma_ts_c<true>(d_t, alo + 56u, dr + 6ull, idesc);  \
            mma_commit(&ring_empty[0]);                                                                      \
            mbar_wait(&ring_full[1], ring_par);                                                              \
            mma_ts_c<true>(d_t, ahi + 32u, dr + 1024ull, idesc); mma_ts_c<true>(d_t, ahi + 40u, dr + 1026ull, idesc); \
            mma_ts_c<true>(d_t, ahi + 48u, dr + 1028ull, idesc); mma_ts_c<true>(d_t, ahi + 56u, dr + 1030ull, idesc); \
            mma_commit(&ring_empty[1]);                                                                      \
            ring_par ^= 1u;                                                                                  \
        }                                                                                                    \
        mma_commit(&d_ready[SLOT]);                                                                          \
        tl_mark(p.timeline, 0, 17, 3 + (SLOT), 0, l);                                                        \
        PAR ^= 1u;                                                                                           \
        LAYER = (l + 1u) % (uint32_t)L;                                                                      \
        LEFT--;                                                                                              \
    }
            while (left0 | left1) {  // strict alternation: the two slots run in lockstep, one slot's epilogue under the other's MMAs
                TN_SERVE(0, left0, par0, layer0)
                TN_SERVE(1, left1, par1, layer1)
            }
#undef TN_SERVE
#undef TN_KBLOCK
        }
    } else {
        // ================= slot warps: gather -> A0, per-layer epilogues, heads =================
        const int slot = warp >> 3;
        const uint32_t h = (uint32_t)(warp >> 2) & 1u;  // column half
        const uint32_t q = (uint32_t)warp & 3u;               // TMEM lane quarter this warp may access
        const uint32_t lane_base = (q * 32u) << 16;
        const uint32_t d_t = tbase + 256u * slot + lane_base, ahi = d_t + 128u, alo = d_t + 192u;
        float *stage = reinterpret_cast<float *>(smem + MLP_OFF_STAGE) + (size_t)warp * 4 * MLP_STAGE_STRIDE;
        float *dirb = dirb_s + slot * 4 * 128;
        float4 *red = red_s + slot * 128;
        const float *wd = head_s, *wc = head_s + 128;
        const uint32_t hw = (uint32_t)lane >> 4, l16 = (uint32_t)lane & 15u;
        uint32_t dpar = 0;
        for (uint32_t it = slot; it < my_tiles; it += 2) {
            const uint64_t tile_row0 = (uint64_t)(blockIdx.x + (uint64_t)it * gridDim.x) * 128u;
            const uint64_t warp_row0 = tile_row0 + q * 32u;
            // ---- gather + barycentric interpolation (tetrahedra_tracer.cu:203-220) ----
            // lane i first fetches the matched vertex ids / weights of row i (one coalesced load per warp), then the
            // rows are processed 4 at a time (2 per half-warp, lanes over the 32 features of this column half) with the
            // vertex-row loads of the next step in flight while the current step is interpolated and transposed.
            uint4 myv = make_uint4(TN_EMPTY, TN_EMPTY, TN_EMPTY, TN_EMPTY);
            float myb0 = 0.f, myb1 = 0.f, myb2 = 0.f;
            {
                const uint64_t g = warp_row0 + lane;
                if (g < total_rows) {
                    myv = __ldg(p.vi + g);
                    myb0 = __ldg(p.bary + 3 * g); myb1 = __ldg(p.bary + 3 * g + 1); myb2 = __ldg(p.bary + 3 * g + 2);
                }
            }
            tl_mark(p.timeline, lane, warp, 5, it, 0);    // ev 5: gather start
            uint32_t hi[16], lo[16];
            float2 f[2][2][4];  // [buffer][sample of this half-warp][vertex]
            const float2 *fs = reinterpret_cast<const float2 *>(p.fshadow + h * 32) + l16;
#define TN_ISSUE(BUF, C)                                                                                              \
    _Pragma("unroll") for (int j = 0; j < 2; ++j) {                                                                   \
        const int r = 4 * (C) + 2 * j + (int)hw;                                                                      \
        const uint32_t v0 = __shfl_sync(0xffffffffu, myv.x, r), v1 = __shfl_sync(0xffffffffu, myv.y, r);              \
        const uint32_t v2 = __shfl_sync(0xffffffffu, myv.z, r), v3 = __shfl_sync(0xffffffffu, myv.w, r);              \
        const bool m = v0 != TN_EMPTY;                                                                                \
        f[BUF][j][0] = m ? __ldg(fs + (size_t)v0 * 32) : make_float2(0.f, 0.f);                                       \
        f[BUF][j][1] = m ? __ldg(fs + (size_t)v1 * 32) : make_float2(0.f, 0.f);                                       \
        f[BUF][j][2] = m ? __ldg(fs + (size_t)v2 * 32) : make_float2(0.f, 0.f);                                       \
        f[BUF][j][3] = m ? __ldg(fs + (size_t)v3 * 32) : make_float2(0.f, 0.f);                                       \
    }
            TN_ISSUE(0, 0)
#pragma unroll
            for (int c = 0; c < 8; ++c) {
                if (c + 1 < 8) {
                    if ((c & 1) == 0) { TN_ISSUE(1, c + 1) } else { TN_ISSUE(0, c + 1) }
                }
                __syncwarp();  // previous step's readers are done with the staging rows
#pragma unroll
                for (int j = 0; j < 2; ++j) {
                    const int r = 4 * c + 2 * j + (int)hw;
                    const float b0 = __shfl_sync(0xffffffffu, myb0, r), b1 = __shfl_sync(0xffffffffu, myb1, r), b2 = __shfl_sync(0xffffffffu, myb2, r);
                    const float2 *fc = f[c & 1][j];
                    const float w0 = __fsub_rn(1.0f, __fadd_rn(__fadd_rn(b0, b1), b2));
                    float2 o;
                    o.x = __fmaf_rn(b0, fc[1].x, 0.f); o.y = __fmaf_rn(b0, fc[1].y, 0.f);
                    o.x = __fmaf_rn(b1, fc[2].x, o.x); o.y = __fmaf_rn(b1, fc[2].y, o.y);
                    o.x = __fmaf_rn(b2, fc[3].x, o.x); o.y = __fmaf_rn(b2, fc[3].y, o.y);
                    o.x = __fmaf_rn(w0, fc[0].x, o.x); o.y = __fmaf_rn(w0, fc[0].y, o.y);
                    reinterpret_cast<float2 *>(stage + (2 * j + hw) * MLP_STAGE_STRIDE)[l16] = o;
                }
                __syncwarp();
                if ((lane >> 2) == c) {  // lanes 4c..4c+3 own rows 4c..4c+3 of this warp's 32 rows
                    const float4 *rowp = reinterpret_cast<const float4 *>(stage + (lane & 3) * MLP_STAGE_STRIDE);
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const float4 x = rowp[i];
                        split_pack2(x.x, x.y, hi[2 * i], lo[2 * i]);
                        split_pack2(x.z, x.w, hi[2 * i + 1], lo[2 * i + 1]);
                    }
                }
            }
#undef TN_ISSUE
            tmem_st16(ahi + 16 * h, hi);
            tmem_st16(alo + 16 * h, lo);
            tmem_st_wait();
            fence_before_sync();
            __syncwarp();
            if (lane == 0) mbar_arrive(&a_ready[slot]);
            tl_mark(p.timeline, lane, warp, 6, it, 0);    // ev 6: A0 written, arrived

            // ---- FINE: stage the per-ray direction bias of the (few) rays this tile touches ----
            const uint64_t my_row = warp_row0 + lane;
            const float *bias4 = nullptr;
            if (FINE) {
                const uint32_t ray0 = (uint32_t)(tile_row0 / p.S);
                const uint64_t last_row = min(tile_row0 + 127, total_rows - 1);
                const uint32_t nr = (uint32_t)(last_row / p.S) - ray0 + 1;
                const uint32_t my_ray_off = (uint32_t)(min(my_row, total_rows - 1) / p.S) - ray0;
                if (nr <= 4) {
                    for (uint32_t i = (warp & 7) * 32 + lane; i < nr * 128; i += 256) dirb[i] = __ldg(p.dirbias + (size_t)ray0 * 128 + i);
                    asm volatile("bar.sync %0, 256;" ::"r"(1 + slot) : "memory");
                    bias4 = dirb + my_ray_off * 128;
                } else {
                    bias4 = p.dirbias + (size_t)(ray0 + my_ray_off) * 128;  // many short rays per tile: read the bias from L2
                }
            }

            float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
            for (int l = 0; l < L; ++l) {
                tl_mark(p.timeline, lane, warp, 7, it, l);  // ev 7: start waiting for D
                mbar_wait_backoff(&d_ready[slot], dpar, 32);
                dpar ^= 1u;
                fence_after_sync();
                tl_mark(p.timeline, lane, warp, 8, it, l);  // ev 8: D ready seen
                if (l < 2) layer_epilogue<0>(d_t, ahi, alo, h, bias_s + l * 128, wd, wc, acc);
                else if (l == 2 && FINE) layer_epilogue<1>(d_t, ahi, alo, h, bias_s + 256, wd, wc, acc);
                else if (l == 2) layer_epilogue<2>(d_t, ahi, alo, h, bias_s + 256, wd, wc, acc);
                else layer_epilogue<3>(d_t, ahi, alo, h, bias4, wd, wc, acc);
                if (l < L - 1) {
                    tmem_st_wait();
                    fence_before_sync();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&a_ready[slot]);
                }
                tl_mark(p.timeline, lane, warp, 9, it, l);  // ev 9: epilogue of layer l done (+arrive)
            }
            // ---- heads: combine the two column halves, activation, store ----
            if (h == 1) red[q * 32 + lane] = acc;
            asm volatile("bar.sync %0, 256;" ::"r"(1 + slot) : "memory");
            if (h == 0 && my_row < total_rows) {
                const float4 o = red[q * 32 + lane];
                const float sigma = softplus_f(acc.x + o.x + head_s[512]);
                if (FINE) {
                    reinterpret_cast<float4 *>(p.out)[my_row] = make_float4(sigma, sigmoid_f(acc.y + o.y + head_s[513]),
                                                                           sigmoid_f(acc.z + o.z + head_s[514]), sigmoid_f(acc.w + o.w + head_s[515]));
                } else {
                    p.out[my_row] = sigma;
                }
            }
            asm volatile("bar.sync %0, 256;" ::"r"(1 + slot) : "memory");  // red / dirb reuse by the next tile
        }
    }
    fence_before_sync();
    __syncthreads();
    if (warp == 16) tmem_dealloc(tbase, 512);
    if (p.timeline != nullptr && blockIdx.x == 0) {
        const unsigned long long *tl = reinterpret_cast<const unsigned long long *>(smem + MLP_OFF_TL);
        const uint32_t n = (uint32_t)min(tl[0], (unsigned long long)MLP_TL_CAP);
        for (uint32_t i = threadIdx.x; i <= n; i += MLP_THREADS) p.timeline[i] = i == 0 ? (unsigned long long)n : tl[i];
    }
}

}  // namespace tn
