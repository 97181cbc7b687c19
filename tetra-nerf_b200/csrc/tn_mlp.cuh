// tn_mlp.cuh -- the fused interpolate -> MLP kernel on tcgen05 (sm_100a).
//
// Replaces, for one pass over n_active*S samples, the chain
//   interpolate_values (src/tetrahedra_tracer.cu:195-221)  ->  mlp_base (3x Linear+ReLU, model.py:433-438)
//   -> DensityFieldHead (Linear+Softplus, :455) [-> mlp_head (Linear+ReLU, :447-452) -> RGBFieldHead
//   (Linear+Sigmoid, :454)]                                               (call sites model.py:569-621)
// which the reference runs as ~10 torch kernels with [R*S,128] fp32 activations round-tripping HBM.
//
// One persistent CTA per SM, 10 warps:
//   warp 0      : issues every tcgen05.mma (one elected lane)
//   warp 1      : TMEM allocation, TMA bulk staging of the resident weight image, and (FINE) the
//                 producer of the 2-stage TMA ring that streams the 4th layer's weights
//   warps 2..9  : two "slots" of 4 warps (thread = sample row; TMEM lane quarter = warp_id % 4).
//                 A slot owns one 128-sample tile at a time: it gathers the four vertex rows of every
//                 sample from the [V,64] field shadow with coalesced 256-byte reads, forms the
//                 barycentric interpolation with the reference's FMA order, splits it into bf16 hi/lo
//                 and writes it into TMEM as the A operand; then for each layer it waits for the
//                 accumulator, applies bias + ReLU, splits again and writes the next A operand back
//                 into TMEM (activations never touch shared or global memory).  The two slots run
//                 half a tile apart, so one slot's epilogue overlaps the other slot's MMAs.
// Products are "bf16x3": a*w ~= a_hi*w_hi + a_lo*w_hi + a_hi*w_lo with fp32 accumulation in TMEM
// (measured 5e-6 relative on B200, tests/test_gpu_mlp.py) -- the reference computes in fp32 and the
// parity bar is 1e-4 absolute on colour/density, which single-pass bf16/tf32 cannot hold.
// TMEM (512 columns): slot s uses [256s, 256s+128) for D, [+128,+192) A_hi, [+192,+256) A_lo.
#pragma once
#include "tn_common.cuh"
#include "tn_tc.cuh"

namespace tn {

constexpr uint32_t MLP_THREADS = 320;
constexpr uint32_t MLP_W_RESIDENT = 163840;               // L1 32K + L2 64K + L3 64K
constexpr uint32_t MLP_OFF_RING = MLP_W_RESIDENT;          // 2 x 16K
constexpr uint32_t MLP_OFF_STAGE = MLP_OFF_RING + 32768;   // 8 warps x 2176
constexpr uint32_t MLP_STAGE_STRIDE = 68;                  // floats per staged sample row
constexpr uint32_t MLP_OFF_BIAS = MLP_OFF_STAGE + 8 * 8 * MLP_STAGE_STRIDE * 4;  // 3 x 128 floats (b1,b2,b3)
constexpr uint32_t MLP_OFF_HEAD = MLP_OFF_BIAS + 3 * 128 * 4;                    // wd[128] wc[3][128] bd bc[3] (+pad)
constexpr uint32_t MLP_OFF_DIRB = MLP_OFF_HEAD + 520 * 4;                        // 2 slots x 4 rays x 128 floats
constexpr uint32_t MLP_OFF_BARS = MLP_OFF_DIRB + 2 * 4 * 128 * 4;
constexpr uint32_t MLP_SMEM_BYTES = MLP_OFF_BARS + 128;

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
};

__device__ __forceinline__ float softplus_f(float x) { return x > 20.0f ? x : log1pf(expf(x)); }  // torch Softplus(beta=1, threshold=20)
__device__ __forceinline__ float sigmoid_f(float x) { return 1.0f / (1.0f + expf(-x)); }

// bias + ReLU on 2 accumulator values, split to bf16 hi/lo, packed (element 2c in the low half)
__device__ __forceinline__ void act_split2(float a0, float a1, uint32_t &hi, uint32_t &lo, float &r0, float &r1) {
    r0 = fmaxf(a0, 0.0f);
    r1 = fmaxf(a1, 0.0f);
    uint32_t h;
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(h) : "f"(r1), "f"(r0));  // upper <- r1, lower <- r0
    const float h0 = __uint_as_float(h << 16), h1 = __uint_as_float(h & 0xFFFF0000u);
    uint32_t l;
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(l) : "f"(r1 - h1), "f"(r0 - h0));
    hi = h;
    lo = l;
}

template <bool FINE>
__global__ void __launch_bounds__(MLP_THREADS, 1) k_mlp(const MlpParams p) {
    using namespace tc;
    extern __shared__ __align__(1024) uint8_t smem[];
    uint8_t *w_s = smem;
    uint8_t *ring_s = smem + MLP_OFF_RING;
    float *bias_s = reinterpret_cast<float *>(smem + MLP_OFF_BIAS);
    float *head_s = reinterpret_cast<float *>(smem + MLP_OFF_HEAD);
    float *dirb_s = reinterpret_cast<float *>(smem + MLP_OFF_DIRB);
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem + MLP_OFF_BARS);
    uint64_t *a_ready = bars;          // [2] count 4 (one arrive per slot warp)
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

    if (warp == 1) {
        if (lane == 0) {
            mbar_init(&a_ready[0], 4); mbar_init(&a_ready[1], 4);
            mbar_init(&d_ready[0], 1); mbar_init(&d_ready[1], 1);
            mbar_init(w_bar, 1);
            mbar_init(&ring_full[0], 1); mbar_init(&ring_full[1], 1);
            mbar_init(&ring_empty[0], 1); mbar_init(&ring_empty[1], 1);
            fence_barrier_init();
        }
        __syncwarp();
        tmem_alloc(tmem_ptr, 512);
    }
    for (uint32_t i = threadIdx.x; i < 3 * 128; i += MLP_THREADS) bias_s[i] = p.bias[i];
    for (uint32_t i = threadIdx.x; i < 516; i += MLP_THREADS) head_s[i] = p.head[i];
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
    const uint32_t tbase = *tmem_ptr;

    if (warp == 1) {
        // ================= TMA: resident weights, then the L4 ring =================
        if (lane == 0 && my_tiles > 0) {
            mbar_arrive_expect_tx(w_bar, MLP_W_RESIDENT);
            for (uint32_t off = 0; off < MLP_W_RESIDENT; off += 16384) tma_bulk_g2s(w_s + off, p.wimg + off, 16384, w_bar);
            if (FINE) {
                const uint8_t *w4 = p.wimg + MLP_W_RESIDENT;
                const uint32_t nchunks = my_tiles * 4;
                for (uint32_t i = 0; i < nchunks; ++i) {
                    const uint32_t st = i & 1u;
                    mbar_wait(&ring_empty[st], ((i >> 1) & 1u) ^ 1u);
                    mbar_arrive_expect_tx(&ring_full[st], 16384);
                    tma_bulk_g2s(ring_s + st * 16384, w4 + (i & 3u) * 16384, 16384, &ring_full[st]);
                }
            }
        }
    } else if (warp == 0) {
        // ================= MMA issuer =================
        if (lane == 0 && my_tiles > 0) {
            const uint32_t idesc = make_idesc_bf16(128, 128);
            uint32_t left[2] = {((my_tiles + 1) / 2) * (uint32_t)L, (my_tiles / 2) * (uint32_t)L};
            uint32_t par[2] = {0, 0}, layer[2] = {0, 0};
            uint32_t ring_i = 0;
            mbar_wait(w_bar, 0);
            const uint32_t w_base = smem_u32(w_s);
            while (left[0] | left[1]) {
#pragma unroll
                for (int s = 0; s < 2; ++s) {
                    if (!left[s] || !mbar_test(&a_ready[s], par[s])) continue;
                    fence_after_sync();
                    const uint32_t l = layer[s];
                    const uint32_t d_t = tbase + 256u * s, ahi = d_t + 128u, alo = d_t + 192u;
                    if (l < 3) {
                        const uint32_t wl = w_base + (l == 0 ? 0u : (l == 1 ? 32768u : 98304u));
                        const uint32_t nkb = l == 0 ? 1u : 2u;
                        uint32_t acc = 0;
                        for (uint32_t kb = 0; kb < nkb; ++kb) {
                            const uint32_t whi = wl + kb * 32768u, wlo = whi + 16384u;
#pragma unroll
                            for (uint32_t k = 0; k < 4; ++k) { mma_ts(d_t, ahi + kb * 32u + k * 8u, make_desc_sw128(whi + k * 32u), idesc, acc); acc = 1; }
#pragma unroll
                            for (uint32_t k = 0; k < 4; ++k) mma_ts(d_t, alo + kb * 32u + k * 8u, make_desc_sw128(whi + k * 32u), idesc, 1);
#pragma unroll
                            for (uint32_t k = 0; k < 4; ++k) mma_ts(d_t, ahi + kb * 32u + k * 8u, make_desc_sw128(wlo + k * 32u), idesc, 1);
                        }
                    } else {
                        // layer 4: chunks hi(kb0) lo(kb0) hi(kb1) lo(kb1) arrive through the ring
                        uint32_t acc = 0;
                        for (uint32_t c = 0; c < 4; ++c, ++ring_i) {
                            const uint32_t st = ring_i & 1u, kb = c >> 1;
                            mbar_wait(&ring_full[st], (ring_i >> 1) & 1u);
                            const uint32_t wc = smem_u32(ring_s + st * 16384u);
                            if ((c & 1u) == 0) {
#pragma unroll
                                for (uint32_t k = 0; k < 4; ++k) { mma_ts(d_t, ahi + kb * 32u + k * 8u, make_desc_sw128(wc + k * 32u), idesc, acc); acc = 1; }
#pragma unroll
                                for (uint32_t k = 0; k < 4; ++k) mma_ts(d_t, alo + kb * 32u + k * 8u, make_desc_sw128(wc + k * 32u), idesc, 1);
                            } else {
#pragma unroll
                                for (uint32_t k = 0; k < 4; ++k) mma_ts(d_t, ahi + kb * 32u + k * 8u, make_desc_sw128(wc + k * 32u), idesc, 1);
                            }
                            mma_commit(&ring_empty[st]);
                        }
                    }
                    mma_commit(&d_ready[s]);
                    par[s] ^= 1u;
                    layer[s] = (l + 1u) % (uint32_t)L;
                    left[s]--;
                }
            }
        }
    } else {
        // ================= slot warps: gather -> A0, per-layer epilogues, heads =================
        const int slot = (warp - 2) >> 2;
        const uint32_t q = (uint32_t)warp & 3u;  // TMEM lane quarter this warp may access
        const uint32_t lane_base = (q * 32u) << 16;
        const uint32_t d_t = tbase + 256u * slot + lane_base, ahi = d_t + 128u, alo = d_t + 192u;
        float *stage = reinterpret_cast<float *>(smem + MLP_OFF_STAGE) + (size_t)(warp - 2) * 8 * MLP_STAGE_STRIDE;
        float *dirb = dirb_s + slot * 4 * 128;
        const float *wd = head_s, *wc = head_s + 128;
        const float bd = head_s[512], bc0 = head_s[513], bc1 = head_s[514], bc2 = head_s[515];
        uint32_t dpar = 0;
        for (uint32_t it = slot; it < my_tiles; it += 2) {
            const uint64_t tile_row0 = (uint64_t)(blockIdx.x + (uint64_t)it * gridDim.x) * 128u;
            const uint64_t warp_row0 = tile_row0 + q * 32u;
            // ---- gather + barycentric interpolation (tetrahedra_tracer.cu:203-220), 8 samples per pass ----
            uint32_t hi[32], lo[32];
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                uint4 v[8];
                float b0[8], b1[8], b2[8];
#pragma unroll
                for (int s8 = 0; s8 < 8; ++s8) {
                    const uint64_t g = warp_row0 + c * 8 + s8;
                    if (g < total_rows) {
                        v[s8] = __ldg(p.vi + g);
                        b0[s8] = __ldg(p.bary + 3 * g); b1[s8] = __ldg(p.bary + 3 * g + 1); b2[s8] = __ldg(p.bary + 3 * g + 2);
                    } else {
                        v[s8] = make_uint4(TN_EMPTY, TN_EMPTY, TN_EMPTY, TN_EMPTY);
                        b0[s8] = b1[s8] = b2[s8] = 0.f;
                    }
                }
                float2 f[8][4];
#pragma unroll
                for (int s8 = 0; s8 < 8; ++s8) {
                    const bool m = v[s8].x != TN_EMPTY;
                    const uint32_t vv[4] = {v[s8].x, v[s8].y, v[s8].z, v[s8].w};
#pragma unroll
                    for (int k = 0; k < 4; ++k)
                        f[s8][k] = m ? __ldg(reinterpret_cast<const float2 *>(p.fshadow + (size_t)vv[k] * 64) + lane) : make_float2(0.f, 0.f);
                }
                __syncwarp();  // previous pass's readers are done with the staging rows
#pragma unroll
                for (int s8 = 0; s8 < 8; ++s8) {
                    const float w0 = __fsub_rn(1.0f, __fadd_rn(__fadd_rn(b0[s8], b1[s8]), b2[s8]));
                    float2 o;
                    o.x = __fmaf_rn(b0[s8], f[s8][1].x, 0.f); o.y = __fmaf_rn(b0[s8], f[s8][1].y, 0.f);
                    o.x = __fmaf_rn(b1[s8], f[s8][2].x, o.x); o.y = __fmaf_rn(b1[s8], f[s8][2].y, o.y);
                    o.x = __fmaf_rn(b2[s8], f[s8][3].x, o.x); o.y = __fmaf_rn(b2[s8], f[s8][3].y, o.y);
                    o.x = __fmaf_rn(w0, f[s8][0].x, o.x); o.y = __fmaf_rn(w0, f[s8][0].y, o.y);
                    reinterpret_cast<float2 *>(stage + s8 * MLP_STAGE_STRIDE)[lane] = o;
                }
                __syncwarp();
                if ((lane >> 3) == c) {  // lanes 8c..8c+7 own rows 8c..8c+7 of this warp's 32 rows
                    const float4 *rowp = reinterpret_cast<const float4 *>(stage + (lane & 7) * MLP_STAGE_STRIDE);
#pragma unroll
                    for (int i = 0; i < 16; ++i) {
                        const float4 x = rowp[i];
                        split_pack2(x.x, x.y, hi[2 * i], lo[2 * i]);
                        split_pack2(x.z, x.w, hi[2 * i + 1], lo[2 * i + 1]);
                    }
                }
            }
            tmem_st16(ahi, hi); tmem_st16(ahi + 16, hi + 16);
            tmem_st16(alo, lo); tmem_st16(alo + 16, lo + 16);
            tmem_st_wait();
            fence_before_sync();
            __syncwarp();
            if (lane == 0) mbar_arrive(&a_ready[slot]);

            // ---- FINE: stage the per-ray direction bias of the (few) rays this tile touches ----
            const uint64_t my_row = warp_row0 + lane;
            uint32_t ray0 = 0, my_ray_off = 0;
            bool dirb_smem = true;
            if (FINE) {
                ray0 = (uint32_t)(tile_row0 / p.S);
                const uint64_t last_row = min(tile_row0 + 127, total_rows - 1);
                const uint32_t nr = (uint32_t)(last_row / p.S) - ray0 + 1;
                dirb_smem = nr <= 4;
                my_ray_off = (uint32_t)(min(my_row, total_rows - 1) / p.S) - ray0;
                if (dirb_smem) {
                    // 4 warps of the slot cooperatively copy nr*128 floats; named barrier 1+slot syncs the slot
                    for (uint32_t i = q * 32 + lane; i < nr * 128; i += 128) dirb[i] = __ldg(p.dirbias + (size_t)ray0 * 128 + i);
                    asm volatile("bar.sync %0, 128;" ::"r"(1 + slot) : "memory");
                }
            }

            float dens = 0.f, cr = 0.f, cg = 0.f, cb = 0.f;
#pragma unroll 1
            for (int l = 0; l < L; ++l) {
                mbar_wait(&d_ready[slot], dpar);
                dpar ^= 1u;
                fence_after_sync();
#pragma unroll
                for (int ch = 0; ch < 4; ++ch) {
                    uint32_t r[32];
                    tmem_ld32(d_t + ch * 32, r);
                    tmem_ld_wait();
                    uint32_t ph[16], pl[16];
                    const float *bl = (l < 3) ? (bias_s + l * 128 + ch * 32)
                                              : (dirb_smem ? (dirb + my_ray_off * 128 + ch * 32) : nullptr);
#pragma unroll
                    for (int i = 0; i < 16; ++i) {
                        float x0 = __uint_as_float(r[2 * i]), x1 = __uint_as_float(r[2 * i + 1]);
                        if (bl) { x0 += bl[2 * i]; x1 += bl[2 * i + 1]; }
                        else {
                            const float *gb = p.dirbias + (size_t)(ray0 + my_ray_off) * 128 + ch * 32;
                            x0 += __ldg(gb + 2 * i); x1 += __ldg(gb + 2 * i + 1);
                        }
                        float r0, r1;
                        act_split2(x0, x1, ph[i], pl[i], r0, r1);
                        if (l == 2) { dens = fmaf(r0, wd[ch * 32 + 2 * i], dens); dens = fmaf(r1, wd[ch * 32 + 2 * i + 1], dens); }
                        if (FINE && l == 3) {
                            const int c0 = ch * 32 + 2 * i;
                            cr = fmaf(r0, wc[c0], cr); cr = fmaf(r1, wc[c0 + 1], cr);
                            cg = fmaf(r0, wc[128 + c0], cg); cg = fmaf(r1, wc[128 + c0 + 1], cg);
                            cb = fmaf(r0, wc[256 + c0], cb); cb = fmaf(r1, wc[256 + c0 + 1], cb);
                        }
                    }
                    if (l < L - 1) { tmem_st16(ahi + ch * 16, ph); tmem_st16(alo + ch * 16, pl); }
                }
                if (l < L - 1) {
                    tmem_st_wait();
                    fence_before_sync();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&a_ready[slot]);
                }
            }
            if (my_row < total_rows) {
                const float sigma = softplus_f(dens + bd);
                if (FINE) {
                    reinterpret_cast<float4 *>(p.out)[my_row] = make_float4(sigma, sigmoid_f(cr + bc0), sigmoid_f(cg + bc1), sigmoid_f(cb + bc2));
                } else {
                    p.out[my_row] = sigma;
                }
            }
            if (FINE) asm volatile("bar.sync %0, 128;" ::"r"(1 + slot) : "memory");  // dirb reuse by the next tile
        }
    }
    fence_before_sync();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tbase, 512);
}

}  // namespace tn
