// tn_mlp_bwd.cuh -- fused backward of the interpolate -> MLP fine pass on tcgen05 (sm_100a): the training half of the hot path.
//
// Replaces, for the n_active*S2 samples of the fine pass of one training step, what autograd runs in the reference as ~40 torch
// kernels over [R*S,128] fp32 activations in HBM:
//   backward of RGBFieldHead / DensityFieldHead / mlp_head / mlp_base (tetranerf/nerfstudio/model.py:596-621, nerfstudio MLP)
//   and interpolate_values_backward (src/tetrahedra_tracer.cu:223-248, tetranerf/utils/extension/__init__.py:36-42).
// Nothing of size [samples,128] ever reaches HBM: per 128-sample tile the forward activations are RECOMPUTED (4 GEMMs), the
// input-gradient chain (4 GEMMs) and the weight-gradient GEMMs (4) run on the tensor cores with bf16x3 products (fp32-level
// accuracy), the weight gradients of the three 128x128 layers stay in TMEM for the whole kernel, and the feature gradient
// goes to the [V,64] field-gradient shadow with 16-byte vector reductions.
//
// Per tile (rows = samples; every operand is bf16 hi|lo in ONE shared-memory layout: [row][64-column block], 128-byte rows,
// 128-byte swizzle; the same bytes are read K-major (K along the columns) or MN-major (K along the ROWS) through the
// instruction descriptor's major bits, so no transposed copy of anything is ever made):
//   forward   L1: D = X  W1^T     L2: D = H1 W2^T     L3: D = H2 W3^T     L4: D = H3 W4b^T         (A, B K-major)
//   dX chain  dH3 = dA4 W4b       dH2 = dA3 W3        dH1 = dA2 W2        dX = dA1 W1              (A K-major, B = W MN-major)
//   dW        dW4b += dA4^T H3    dW3 += dA3^T H2     dW2 += dA2^T H1     dW1 += dA1^T X           (A, B MN-major, K = samples)
// X = interpolated features (gather warps, as the forward kernel), H_l = relu(.) activations, dA_l = dH_l * (H_l > 0).
// H1, H2 (and X) are needed again long after their buffers have been overwritten: they are parked in an L2-resident per-CTA
// scratch (TMA bulk store / load of the exact shared-memory image), 160 KB per CTA.
// Shared memory: X 32 KB | H 64 KB | dA 64 KB | weight ring 2 x 32 KB | heads, barriers.   TMEM: D 128 | dW4b 128 | dW3 128 | dW2 128.
// Roles (12 warps): 0..7 epilogue workers (lane quarter q = warp & 3, column half h = warp >> 2), 8 weight-ring producer (TMA),
// 9 MMA issuer (+ the park / restore copies), 10..11 gather warps (64 tile rows each; + tile scheduler).  X is parked as well (and
// restored into the H buffer for dW1), so the gather of the next tile runs under the whole current tile.
#pragma once
#include "tn_common.cuh"
#include "tn_mlp.cuh"
#include "tn_tc.cuh"

namespace tn {

constexpr uint32_t BWD_THREADS = 384;   // 12 warps = 3 per SM sub-partition: up to 168 registers per thread
constexpr uint32_t BWD_GATHER_WARP0 = 10;
constexpr uint32_t BWD_OFF_X = 0;                     // hi 16 KB | lo 16 KB
constexpr uint32_t BWD_OFF_H = 32768;                 // hi blk0 | hi blk1 | lo blk0 | lo blk1 (16 KB each)
constexpr uint32_t BWD_OFF_DA = 98304;                // same layout
constexpr uint32_t BWD_OFF_RING = 163840;             // 2 stages x 32 KB
constexpr uint32_t BWD_STAGE = 32768;
constexpr uint32_t BWD_OFF_HEAD = 229376;             // wd[128] wc[3][128]
constexpr uint32_t BWD_OFF_BARS = BWD_OFF_HEAD + 2048;
constexpr uint32_t BWD_SMEM_BYTES = BWD_OFF_BARS + 256;
static_assert(BWD_SMEM_BYTES <= 232448, "k_mlp_bwd shared memory exceeds 227 KB");
constexpr uint32_t BWD_WIMG_BYTES = 7 * BWD_STAGE;    // [L1 hi|lo][L2 HI][L2 LO][L3 HI][L3 LO][L4 HI][L4 LO]
constexpr uint32_t BWD_SCRATCH_PER_CTA = 2 * 65536 + 32768;   // parked H1, H2, X
constexpr uint32_t BWD_NSEQ = 14;                     // weight stages consumed per tile
// barriers
constexpr uint32_t BB_X_FULL = 0, BB_X_EMPTY = 1, BB_RING_FULL = 2, BB_RING_EMPTY = 4, BB_D_READY = 6, BB_OPS_READY = 7, BB_DW_DONE = 8, BB_H_LOADED = 9;
constexpr uint32_t BWD_TILE_IDS_OFF = 96;             // byte offset of tile_ids[8] inside the barrier block; tmem ptr at 128, stop flag at 132

struct MlpBwdParams {
    const uint32_t *n_active;
    uint32_t S;                 // samples per ray of the fine pass (S2)
    const uint4 *vi;            // [rows] matched vertex ids
    const float *bary;          // [rows,3]
    const float *fshadow;       // [V,64]
    const uint8_t *wimg;        // backward weight image (7 stages of 32 KB, see pack_weights_bwd)
    const float *bias;          // b1 b2 b3 [3][128]
    const float *head;          // wd[128] wc[3][128] ...
    const float *dirbias;       // [n_active,128]
    const float4 *dout;         // [rows] (d sigma_pre, d z_r, d z_g, d z_b): gradients at the head pre-activations
    uint8_t *scratch;           // [grid] x 128 KB
    float *gshadow;             // [V,64] field gradient accumulator (zeroed by the caller)
    float *gw;                  // packed MLP gradient accumulators (zeroed by the caller), see GW_* offsets
    float *g_dirbias;           // [n_active,128] gradient at the per-ray direction bias (zeroed by the caller)
    uint32_t *tile_ctr;
};
// offsets (floats) inside gw
constexpr uint32_t GW_W1 = 0, GW_W2 = 8192, GW_W3 = 24576, GW_W4B = 40960, GW_B1 = 57344, GW_B2 = 57472, GW_B3 = 57600, GW_WD = 57728, GW_WC = 57856,
                   GW_SUMS = 58240 /* sum ds, sum dz_r, dz_g, dz_b */, GW_W4DIR = 58244 /* [128][27] */, GW_B4 = 61700, GW_TOTAL = 61828;

// low descriptor word of a shared-memory operand (start address >> 4 | leading-dimension byte offset >> 4 at bit 16)
__device__ __forceinline__ uint32_t dlo(uint32_t smem_addr, uint32_t lbo_bytes) { return ((smem_addr & 0x3FFFFu) >> 4) | ((lbo_bytes >> 4) << 16); }
// D[tmem] (+)= A[smem] * B[smem]; descriptors given by their low words (high word: SBO 1024 B, version 1, 128-byte swizzle)
__device__ __forceinline__ void mma_ss_lo(uint32_t d_tmem, uint32_t a_lo, uint32_t b_lo, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t.reg .b32 hi;\n\t.reg .b64 da, db;\n\t"
        "setp.ne.b32 p, %4, 0;\n\tmov.u32 hi, " TN_DESC_HI_SW128 ";\n\tmov.b64 da, {%1, hi};\n\tmov.b64 db, {%2, hi};\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %3, p;\n\t}" ::"r"(d_tmem), "r"(a_lo), "r"(b_lo), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void tma_bulk_s2g(void *dst_gmem, uint32_t src_smem, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst_gmem), "r"(src_smem), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait0() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void red_add_v4(float *dst, float a, float b, float c, float d) {
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

// column sums over the 32 rows (lanes) of a warp for 16 columns held one per register: four halving exchange steps leave lane l
// with the sum over the lanes of its 16-lane half for column l & 15, a last exchange adds the other half: 16 shuffles, result
// replicated in lanes l and l ^ 16.  Every index is a compile-time constant after unrolling, so v[] stays in registers.
__device__ __forceinline__ float warp_colsum16(float (&v)[16], int lane) {
#pragma unroll
    for (int w = 8; w >= 1; w >>= 1) {
        const bool up = (lane & w) != 0;
#pragma unroll
        for (int i = 0; i < w; ++i) {
            const float send = up ? v[i] : v[i + w];
            const float keep = up ? v[i + w] : v[i];
            v[i] = keep + __shfl_xor_sync(0xffffffffu, send, w);
        }
    }
    return v[0] + __shfl_xor_sync(0xffffffffu, v[0], 16);
}

// 16 fp32 values (columns col0 .. col0+15 of this thread's row, col0 a multiple of 16) -> bf16 hi/lo pairs, then stored into a
// [row][64-col block] operand buffer (hi at row_addr + block / chunk offsets, lo `lo_off` bytes further).
// row_addr = buffer + (R >> 3) * 1024 + (R & 7) * 128.
__device__ __forceinline__ void pack16(const float (&x)[16], uint32_t (&h)[8], uint32_t (&l)[8]) {
#pragma unroll
    for (int i = 0; i < 8; ++i) tc::split_pack2(x[2 * i], x[2 * i + 1], h[i], l[i]);
}
__device__ __forceinline__ void store_packed16(uint32_t row_addr, uint32_t r7, uint32_t col0, uint32_t lo_off, const uint32_t (&h)[8], const uint32_t (&l)[8]) {
    const uint32_t blk = col0 >> 6, c16 = (col0 & 63u) >> 3;  // 16-byte chunk index of the first column inside its 128-byte row
#pragma unroll
    for (int q = 0; q < 2; ++q) {
        const uint32_t a = row_addr + blk * 16384u + (((c16 + (uint32_t)q) ^ r7) << 4);
        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(a), "r"(h[4 * q]), "r"(h[4 * q + 1]), "r"(h[4 * q + 2]), "r"(h[4 * q + 3]) : "memory");
        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(a + lo_off), "r"(l[4 * q]), "r"(l[4 * q + 1]), "r"(l[4 * q + 2]), "r"(l[4 * q + 3]) : "memory");
    }
}
__device__ __forceinline__ void store_operand16(uint32_t row_addr, uint32_t r7, uint32_t col0, uint32_t lo_off, const float (&x)[16]) {
    uint32_t h[8], l[8];
    pack16(x, h, l);
    store_packed16(row_addr, r7, col0, lo_off, h, l);
}

extern __shared__ __align__(1024) uint8_t tn_bwd_smem[];

__global__ void __launch_bounds__(BWD_THREADS, 1) k_mlp_bwd(const MlpBwdParams p) {
    using namespace tc;
    uint8_t *smem = tn_bwd_smem;
    float *head_s = reinterpret_cast<float *>(smem + BWD_OFF_HEAD);  // wd[128], wc[3][128]
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem + BWD_OFF_BARS);
    volatile uint32_t *tile_ids = reinterpret_cast<volatile uint32_t *>(smem + BWD_OFF_BARS + BWD_TILE_IDS_OFF);
    uint32_t *tmem_ptr = reinterpret_cast<uint32_t *>(smem + BWD_OFF_BARS + 128);
    volatile uint32_t *stop_flag = reinterpret_cast<volatile uint32_t *>(smem + BWD_OFF_BARS + 132);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t n_active = *p.n_active;
    const uint64_t total_rows = (uint64_t)n_active * p.S;
    const uint32_t ntiles = (uint32_t)((total_rows + 127) / 128);
    const bool has_work = blockIdx.x < ntiles;

    if (warp == 8) {
        if (lane == 0) {
            mbar_init(&bars[BB_X_FULL], 2); mbar_init(&bars[BB_X_EMPTY], 1);
            mbar_init(&bars[BB_RING_FULL], 1); mbar_init(&bars[BB_RING_FULL + 1], 1);
            mbar_init(&bars[BB_RING_EMPTY], 1); mbar_init(&bars[BB_RING_EMPTY + 1], 1);
            mbar_init(&bars[BB_D_READY], 1); mbar_init(&bars[BB_OPS_READY], 8); mbar_init(&bars[BB_DW_DONE], 1); mbar_init(&bars[BB_H_LOADED], 1);
            fence_barrier_init();
            *stop_flag = 0u;
            tile_ids[0] = has_work ? blockIdx.x : MLP_NO_TILE;
        }
        __syncwarp();
        tmem_alloc(tmem_ptr, 512);
    }
    for (uint32_t i = threadIdx.x; i < 512; i += BWD_THREADS) head_s[i] = p.head[i];
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
    const uint32_t tbase = *tmem_ptr;
    const uint32_t sX = smem_u32(smem + BWD_OFF_X), sH = smem_u32(smem + BWD_OFF_H), sDA = smem_u32(smem + BWD_OFF_DA), sR = smem_u32(smem + BWD_OFF_RING);

    if (warp == 8) {
        // ================= weight ring producer: the endless stage sequence of the tiles, two stages in flight =================
        if (lane == 0 && has_work) {
            uint32_t st = 0, par = 1, i = 0;
            for (;; ++i) {
                bool stop = false;
                while (!mbar_test(&bars[BB_RING_EMPTY + st], par)) {
                    if (*stop_flag != 0u) { stop = true; break; }
                    __nanosleep(64);
                }
                if (stop) break;
                const uint32_t q = i % BWD_NSEQ;
                // forward L1, L2 HI, L2 LO, L3 HI, L3 LO, L4 HI, L4 LO; backward W4 HI, W4 LO, W3 HI, W3 LO, W2 HI, W2 LO, W1
                const uint32_t img = q < 7u ? q : (q == 13u ? 0u : (q & 1u ? 12u - q : 14u - q));
                mbar_arrive_expect_tx(&bars[BB_RING_FULL + st], BWD_STAGE);
                tma_bulk_g2s(smem + BWD_OFF_RING + st * BWD_STAGE, p.wimg + img * BWD_STAGE, 16384, &bars[BB_RING_FULL + st]);
                tma_bulk_g2s(smem + BWD_OFF_RING + st * BWD_STAGE + 16384, p.wimg + img * BWD_STAGE + 16384, 16384, &bars[BB_RING_FULL + st]);
                if (++st == 2) { st = 0; par ^= 1u; }
            }
            const uint32_t fills0 = (i + 1u) / 2u, fills1 = i / 2u;  // fills of stage 0 / 1 issued so far: wait for the last of each to land
            if (fills0) mbar_wait_backoff(&bars[BB_RING_FULL], (fills0 - 1u) & 1u, 64);
            if (fills1) mbar_wait_backoff(&bars[BB_RING_FULL + 1], (fills1 - 1u) & 1u, 64);
        }
    } else if (warp == 9) {
        // ================= MMA issuer (one thread) =================
        if (lane == 0 && has_work) {
            constexpr uint32_t ID_FWD = make_idesc_bf16(128, 128);
            constexpr uint32_t ID_DX = make_idesc_bf16(128, 128) | (1u << 16);
            constexpr uint32_t ID_DX1 = make_idesc_bf16(128, 64) | (1u << 16);
            constexpr uint32_t ID_DW = make_idesc_bf16(128, 128) | (1u << 15) | (1u << 16);
            constexpr uint32_t ID_DW1 = make_idesc_bf16(128, 64) | (1u << 15) | (1u << 16);
            const uint32_t tD = tbase, tW4 = tbase + 128u, tW3 = tbase + 256u, tW2 = tbase + 384u;
            uint8_t *scr = p.scratch + (size_t)blockIdx.x * BWD_SCRATCH_PER_CTA;
            uint32_t rst = 0, rpar = 0;              // ring stage / parity
            uint32_t p_ops = 0, p_dw = 0, p_hl = 0;  // parities of ops_ready, dw_done, h_loaded
            uint32_t nseq = 0;
            auto ring_wait = [&]() -> uint32_t {
                mbar_wait(&bars[BB_RING_FULL + rst], rpar);
                return sR + rst * BWD_STAGE;
            };
            auto ring_release = [&]() {
                mma_commit(&bars[BB_RING_EMPTY + rst]);
                if (++rst == 2) { rst = 0; rpar ^= 1u; }
            };
            auto ops_wait = [&]() {
                mbar_wait(&bars[BB_OPS_READY], p_ops);
                p_ops ^= 1u;
                fence_after_sync();
            };
            // forward layer with a 128-wide input: A = H (K-major), B = weight stage pair (K-major)
            auto fwd128 = [&]() {
                uint32_t w = ring_wait();
#pragma unroll 1
                for (uint32_t t = 0; t < 2; ++t)  // (A_hi, W_hi), (A_lo, W_hi)
#pragma unroll 1
                    for (uint32_t j = 0; j < 8; ++j)
                        mma_ss_lo(tD, dlo(sH + t * 32768u + (j >> 2) * 16384u + (j & 3u) * 32u, 0), dlo(w + (j >> 2) * 16384u + (j & 3u) * 32u, 0), ID_FWD, (t | j) != 0u);
                ring_release();
                w = ring_wait();
#pragma unroll 1
                for (uint32_t j = 0; j < 8; ++j)  // (A_hi, W_lo)
                    mma_ss_lo(tD, dlo(sH + (j >> 2) * 16384u + (j & 3u) * 32u, 0), dlo(w + (j >> 2) * 16384u + (j & 3u) * 32u, 0), ID_FWD, 1u);
                ring_release();
            };
            // dH = dA W (A = dA K-major over its 128 columns, B = W MN-major: K runs over the stage's 128 rows, N over its two 64-col blocks)
            auto dx128 = [&]() {
                uint32_t w = ring_wait();
#pragma unroll 1
                for (uint32_t t = 0; t < 2; ++t)
#pragma unroll 1
                    for (uint32_t j = 0; j < 8; ++j)
                        mma_ss_lo(tD, dlo(sDA + t * 32768u + (j >> 2) * 16384u + (j & 3u) * 32u, 0), dlo(w + j * 2048u, 16384u), ID_DX, (t | j) != 0u);
                ring_release();
                w = ring_wait();
#pragma unroll 1
                for (uint32_t j = 0; j < 8; ++j) mma_ss_lo(tD, dlo(sDA + (j >> 2) * 16384u + (j & 3u) * 32u, 0), dlo(w + j * 2048u, 16384u), ID_DX, 1u);
                ring_release();
            };
            // dW += dA^T H (both MN-major, K = the tile's 128 samples = the buffers' rows)
            auto dw128 = [&](uint32_t tW, bool first_tile) {
#pragma unroll 1
                for (uint32_t t = 0; t < 3; ++t) {  // (dA_hi, H_hi), (dA_lo, H_hi), (dA_hi, H_lo)
                    const uint32_t ao = t == 1u ? 32768u : 0u, bo = t == 2u ? 32768u : 0u;
#pragma unroll 1
                    for (uint32_t j = 0; j < 8; ++j)
                        mma_ss_lo(tW, dlo(sDA + ao + j * 2048u, 16384u), dlo(sH + bo + j * 2048u, 16384u), ID_DW, (first_tile && (t | j) == 0u) ? 0u : 1u);
                }
                mma_commit(&bars[BB_DW_DONE]);
            };
            auto restore_h = [&](uint32_t which) {  // parked H -> H buffer, once the dW MMAs that still read it are done
                mbar_wait(&bars[BB_DW_DONE], p_dw);
                p_dw ^= 1u;
                bulk_wait0();  // the park stores have been written
                asm volatile("fence.proxy.async;" ::: "memory");
                mbar_arrive_expect_tx(&bars[BB_H_LOADED], 65536);
                for (uint32_t c = 0; c < 4; ++c) tma_bulk_g2s(smem + BWD_OFF_H + c * 16384u, scr + which * 65536u + c * 16384u, 16384, &bars[BB_H_LOADED]);
            };
            for (;;) {
                // ---- next tile ----
                mbar_wait(&bars[BB_X_FULL], nseq & 1u);
                uint32_t tile;
                asm volatile("ld.volatile.shared.u32 %0, [%1];" : "=r"(tile) : "r"(smem_u32(smem + BWD_OFF_BARS + BWD_TILE_IDS_OFF) + 4u * (nseq & 7u)) : "memory");
                if (tile == MLP_NO_TILE) {  // release the epilogue workers (they read the same sentinel) and stop
                    mbar_arrive(&bars[BB_D_READY]);
                    break;
                }
                const bool first_tile = nseq == 0;
                nseq++;
                fence_after_sync();
                // ---- forward L1: D = X W1^T (K = 64; stage = [W1 hi | W1 lo]); X is parked for dW1 and its buffer handed back to the gather ----
                {
                    tma_bulk_s2g(scr + 131072u, sX, 16384);
                    tma_bulk_s2g(scr + 131072u + 16384u, sX + 16384u, 16384);
                    bulk_commit();
                    const uint32_t w = ring_wait();
#pragma unroll 1
                    for (uint32_t t = 0; t < 3; ++t) {
                        const uint32_t ao = t == 1u ? 16384u : 0u, bo = t == 2u ? 16384u : 0u;
#pragma unroll 1
                        for (uint32_t k = 0; k < 4; ++k) mma_ss_lo(tD, dlo(sX + ao + k * 32u, 0), dlo(w + bo + k * 32u, 0), ID_FWD, (t | k) != 0u);
                    }
                    ring_release();
                    mma_commit(&bars[BB_D_READY]);
                    bulk_wait_read0();
                    mma_commit(&bars[BB_X_EMPTY]);  // the L1 MMAs and the park copy have read X
                }
                // ---- forward L2 (H = H1: park it), L3 (H = H2: park it), L4 (H = H3 stays) ----
                for (uint32_t l = 0; l < 3; ++l) {
                    ops_wait();
                    if (l < 2) {
                        for (uint32_t c = 0; c < 4; ++c) tma_bulk_s2g(scr + l * 65536u + c * 16384u, sH + c * 16384u, 16384);
                        bulk_commit();
                    }
                    fwd128();
                    if (l < 2) bulk_wait_read0();  // the park copy has read H before the next epilogue may overwrite it
                    mma_commit(&bars[BB_D_READY]);
                }
                // ---- backward layer 4: dH3 = dA4 W4b ; dW4b += dA4^T H3 ----
                ops_wait();
                dx128();
                mma_commit(&bars[BB_D_READY]);
                dw128(tW4, first_tile);
                restore_h(1);  // H2
                // ---- layer 3 ----
                ops_wait();
                dx128();
                mma_commit(&bars[BB_D_READY]);
                mbar_wait(&bars[BB_H_LOADED], p_hl); p_hl ^= 1u;
                dw128(tW3, first_tile);
                restore_h(0);  // H1
                // ---- layer 2 ----
                ops_wait();
                dx128();
                mma_commit(&bars[BB_D_READY]);
                mbar_wait(&bars[BB_H_LOADED], p_hl); p_hl ^= 1u;
                dw128(tW2, first_tile);
                {   // parked X -> first 32 KB of the H buffer, once dW2 is done with H1
                    mbar_wait(&bars[BB_DW_DONE], p_dw);
                    p_dw ^= 1u;
                    bulk_wait0();
                    asm volatile("fence.proxy.async;" ::: "memory");
                    mbar_arrive_expect_tx(&bars[BB_H_LOADED], 32768);
                    tma_bulk_g2s(smem + BWD_OFF_H, scr + 131072u, 16384, &bars[BB_H_LOADED]);
                    tma_bulk_g2s(smem + BWD_OFF_H + 16384u, scr + 131072u + 16384u, 16384, &bars[BB_H_LOADED]);
                }
                // ---- layer 1: dX = dA1 W1 (N = 64; stage = [W1 hi | W1 lo], K over its 128 rows) ----
                ops_wait();
                {
                    const uint32_t w = ring_wait();
#pragma unroll 1
                    for (uint32_t t = 0; t < 3; ++t) {
                        const uint32_t ao = t == 1u ? 32768u : 0u, bo = t == 2u ? 16384u : 0u;
#pragma unroll 1
                        for (uint32_t j = 0; j < 8; ++j)
                            mma_ss_lo(tD, dlo(sDA + ao + (j >> 2) * 16384u + (j & 3u) * 32u, 0), dlo(w + bo + j * 2048u, 0), ID_DX1, (t | j) != 0u);
                    }
                    ring_release();
                    mma_commit(&bars[BB_D_READY]);
                }
                // ---- dW1 (this tile's part) = dA1^T X into D once the dX epilogue has read it; X sits in the H buffer now ----
                ops_wait();
                mbar_wait(&bars[BB_H_LOADED], p_hl); p_hl ^= 1u;
#pragma unroll 1
                for (uint32_t t = 0; t < 3; ++t) {
                    const uint32_t ao = t == 1u ? 32768u : 0u, bo = t == 2u ? 16384u : 0u;
#pragma unroll 1
                    for (uint32_t j = 0; j < 8; ++j) mma_ss_lo(tD, dlo(sDA + ao + j * 2048u, 16384u), dlo(sH + bo + j * 2048u, 0), ID_DW1, (t | j) != 0u);
                }
                mma_commit(&bars[BB_D_READY]);
                ops_wait();  // D is free again; H is free as well (the workers have seen the dW1 result)
            }
            *stop_flag = 1u;
        }
    } else if (warp >= (int)BWD_GATHER_WARP0) {
        // ================= gather warps: interpolated features -> X operand (bf16 hi | lo, K-major, 128-byte swizzle) =================
        // Two warps, 64 tile rows each, as eight passes of 8 rows: per pass a half-warp owns one row at a time (lane l16 -> features
        // 4 l16 .. 4 l16 + 3), 4 steps of two rows, the vertex-row loads of all 4 steps (16 x 16 bytes per lane) in flight together.  The X buffer is handed
        // back right after the tile's first GEMM (it is parked for dW1), so this runs under the rest of the previous tile.
        if (has_work) {
            const uint32_t gwp = (uint32_t)warp - BWD_GATHER_WARP0;
            const bool sched = gwp == 0 && lane == 0;
            const uint32_t hw = (uint32_t)lane >> 4, l16 = (uint32_t)lane & 15u;
            const float4 *fs = reinterpret_cast<const float4 *>(p.fshadow) + l16;
            uint32_t cur = blockIdx.x, pending = MLP_NO_TILE;
            if (sched) pending = gridDim.x + atomicAdd(p.tile_ctr, 1u);
            uint32_t n = 0;
            for (; cur != MLP_NO_TILE; ++n) {
                if (sched) {  // publish the tile of sequence number n+1, draw the one of n+2
                    const uint32_t v = pending < ntiles ? pending : MLP_NO_TILE;
                    tile_ids[(n + 1u) & 7u] = v;
                    pending = v != MLP_NO_TILE ? gridDim.x + atomicAdd(p.tile_ctr, 1u) : MLP_NO_TILE;
                }
                if (n >= 1) mbar_wait_backoff(&bars[BB_X_EMPTY], (n - 1u) & 1u, 64);  // the first GEMM and the park copy of the previous tile have read X
#pragma unroll 1
                for (uint32_t pass = 0; pass < 8; ++pass) {
                    const uint32_t row0 = gwp * 64u + pass * 8u;  // rows row0 .. row0 + 7
                    // lane i < 8 keeps the matched vertex ids / weights of row row0 + i
                    uint4 cv = make_uint4(TN_EMPTY, TN_EMPTY, TN_EMPTY, TN_EMPTY);
                    float cb0 = 0.f, cb1 = 0.f, cb2 = 0.f;
                    {
                        const uint64_t gr = (uint64_t)cur * 128u + row0 + ((uint32_t)lane & 7u);
                        if (gr < total_rows) {
                            cv = __ldg(p.vi + gr);
                            cb0 = __ldg(p.bary + 3 * gr); cb1 = __ldg(p.bary + 3 * gr + 1); cb2 = __ldg(p.bary + 3 * gr + 2);
                        }
                    }
                    float4 f[4][4];
#pragma unroll
                    for (int c = 0; c < 4; ++c) {
                        const int r_ = 2 * c + (int)hw;
                        const uint32_t v0 = __shfl_sync(0xffffffffu, cv.x, r_), v1 = __shfl_sync(0xffffffffu, cv.y, r_);
                        const uint32_t v2 = __shfl_sync(0xffffffffu, cv.z, r_), v3 = __shfl_sync(0xffffffffu, cv.w, r_);
                        const bool m_ = v0 != TN_EMPTY;
                        const float4 z_ = make_float4(0.f, 0.f, 0.f, 0.f);
                        f[c][0] = m_ ? ldg_stream(fs + (size_t)v0 * 16) : z_;
                        f[c][1] = m_ ? ldg_stream(fs + (size_t)v1 * 16) : z_;
                        f[c][2] = m_ ? ldg_stream(fs + (size_t)v2 * 16) : z_;
                        f[c][3] = m_ ? ldg_stream(fs + (size_t)v3 * 16) : z_;
                    }
#pragma unroll
                    for (int c = 0; c < 4; ++c) {
                        const int r = 2 * c + (int)hw;
                        const float b0 = __shfl_sync(0xffffffffu, cb0, r), b1 = __shfl_sync(0xffffffffu, cb1, r), b2 = __shfl_sync(0xffffffffu, cb2, r);
                        const float4 *fc = f[c];
                        const float w0 = __fsub_rn(1.0f, __fadd_rn(__fadd_rn(b0, b1), b2));
                        float4 o;  // tetrahedra_tracer.cu:203-220 (the forward's FMA order)
                        o.x = __fmaf_rn(b0, fc[1].x, 0.f); o.y = __fmaf_rn(b0, fc[1].y, 0.f); o.z = __fmaf_rn(b0, fc[1].z, 0.f); o.w = __fmaf_rn(b0, fc[1].w, 0.f);
                        o.x = __fmaf_rn(b1, fc[2].x, o.x); o.y = __fmaf_rn(b1, fc[2].y, o.y); o.z = __fmaf_rn(b1, fc[2].z, o.z); o.w = __fmaf_rn(b1, fc[2].w, o.w);
                        o.x = __fmaf_rn(b2, fc[3].x, o.x); o.y = __fmaf_rn(b2, fc[3].y, o.y); o.z = __fmaf_rn(b2, fc[3].z, o.z); o.w = __fmaf_rn(b2, fc[3].w, o.w);
                        o.x = __fmaf_rn(w0, fc[0].x, o.x); o.y = __fmaf_rn(w0, fc[0].y, o.y); o.z = __fmaf_rn(w0, fc[0].z, o.z); o.w = __fmaf_rn(w0, fc[0].w, o.w);
                        uint32_t h0, l0, h1, l1;
                        split_pack2(o.x, o.y, h0, l0);
                        split_pack2(o.z, o.w, h1, l1);
                        const uint32_t R = row0 + (uint32_t)r, r7 = R & 7u;
                        const uint32_t addr = sX + (R >> 3) * 1024u + r7 * 128u + (((l16 >> 1) ^ r7) << 4) + (l16 & 1u) * 8u;
                        asm volatile("st.shared.v2.b32 [%0], {%1, %2};" ::"r"(addr), "r"(h0), "r"(h1) : "memory");
                        asm volatile("st.shared.v2.b32 [%0], {%1, %2};" ::"r"(addr + 16384u), "r"(l0), "r"(l1) : "memory");
                    }
                }
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                __syncwarp();
                if (lane == 0) mbar_arrive(&bars[BB_X_FULL]);
                asm volatile("bar.sync 3, 64;" ::: "memory");  // the scheduler's tile_ids[n + 1] is visible to both gather warps
                cur = tile_ids[(n + 1u) & 7u];
            }
            // sequence number n has no tile: complete its "X ready" phase so that the issuer wakes up, reads the sentinel and stops
            if (n >= 1) mbar_wait_backoff(&bars[BB_X_EMPTY], (n - 1u) & 1u, 64);
            __syncwarp();
            if (lane == 0) mbar_arrive(&bars[BB_X_FULL]);
        }
    } else {
        // ================= epilogue workers =================
        // warp (q, h): rows 32 q .. 32 q + 31 of the tile (its TMEM lane quarter), columns 64 h .. 64 h + 63, in four chunks of 16
        const uint32_t q = (uint32_t)warp & 3u, h = (uint32_t)warp >> 2;
        const uint32_t lane_base = (q * 32u) << 16;
        const uint32_t tD = tbase + lane_base;
        const uint32_t Rt = q * 32u + (uint32_t)lane, r7 = Rt & 7u;            // this thread's row inside the tile
        const uint32_t rowoff = (Rt >> 3) * 1024u + r7 * 128u;
        const float *wd = head_s, *wc = head_s + 128;
        // per-lane column sums accumulated over the CTA's tiles: lanes l and l ^ 16 hold column 64 h + 16 c + (l & 15) of chunk c
        float acc_b[3][4], acc_wd[4], acc_wc[3][4];
#pragma unroll
        for (int c = 0; c < 4; ++c) { acc_b[0][c] = acc_b[1][c] = acc_b[2][c] = 0.f; acc_wd[c] = 0.f; acc_wc[0][c] = acc_wc[1][c] = acc_wc[2][c] = 0.f; }
        uint32_t p_d = 0, p_dw = 0, ntl = 0;
        auto d_wait = [&]() {
            mbar_wait_backoff(&bars[BB_D_READY], p_d, 32);
            p_d ^= 1u;
            fence_after_sync();
        };
        auto ops_arrive = [&](bool wrote_smem) {
            if (wrote_smem) asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            fence_before_sync();
            __syncwarp();
            if (lane == 0) mbar_arrive(&bars[BB_OPS_READY]);
        };
        float bq[16];  // bias of the chunk that is processed next (loaded a chunk ahead: never between the TMEM load and its use)
#pragma unroll
        for (int i = 0; i < 16; ++i) bq[i] = __ldg(p.bias + h * 64u + i);
        if (has_work) {
            for (uint32_t n = 0;; ++n) {
                d_wait();  // forward layer 1 of the CTA's n-th tile (or the issuer's wake-up call when there is none)
                const uint32_t tile = tile_ids[n & 7u];
                if (tile == MLP_NO_TILE) break;
                ++ntl;
                const uint64_t my_row = (uint64_t)tile * 128u + Rt;
                const bool valid = my_row < total_rows;
                const float4 g4 = valid ? __ldg(p.dout + my_row) : make_float4(0.f, 0.f, 0.f, 0.f);  // (d sigma_pre, d z_rgb) of this thread's row
                // ---------- E1..E3: forward layers 1..3: H_l = relu(D + b_l) -> H buffer; keep the sign masks ----------
                unsigned long long mask[3];
#pragma unroll
                for (int l = 0; l < 3; ++l) {
                    if (l > 0) d_wait();
                    const float *bias = p.bias + l * 128;
                    unsigned long long mk = 0ull;
#pragma unroll
                    for (int c = 0; c < 4; ++c) {
                        const uint32_t col0 = h * 64u + (uint32_t)c * 16u;
                        uint32_t r[16];
                        tmem_ld16(tD + col0, r);
                        tmem_ld_wait();
                        float x[16];
                        uint32_t m = 0;
#pragma unroll
                        for (int i = 0; i < 16; ++i) {
                            const float v = __uint_as_float(r[i]) + bq[i];
                            // refill with the bias of the next chunk / the next layer's first chunk: a chunk ahead of its use
                            if (c < 3) bq[i] = __ldg(bias + col0 + 16u + i);
                            else bq[i] = __ldg(p.bias + (l < 2 ? (l + 1) * 128 : 0) + h * 64u + i);  // (after layer 3: layer 1 of the next tile)
                            m |= (v > 0.f ? 1u : 0u) << i;
                            x[i] = fmaxf(v, 0.f);
                        }
                        mk |= (unsigned long long)m << (16 * c);
                        store_operand16(sH + rowoff, r7, col0, 32768u, x);
                        if (l == 2) {  // d wd[k] = sum_s d sigma_pre[s] H3[s,k]
#pragma unroll
                            for (int i = 0; i < 16; ++i) x[i] *= g4.x;
                            acc_wd[c] += warp_colsum16(x, lane);
                        }
                    }
                    mask[l] = mk;
                    ops_arrive(true);
                }
                // ---------- E4: forward layer 4 (H4 = relu(D + per-ray direction bias)) -> head weight gradients, dA4 ----------
                {
                    d_wait();
                    const uint64_t rr = valid ? my_row : total_rows - 1;
                    const uint32_t slot = (uint32_t)(rr / p.S);
                    const float *db = p.dirbias + (size_t)slot * 128;
                    // the warp's rows belong to slots slot_lo .. slot_hi (usually one, two at a ray boundary)
                    const uint32_t slot_lo = __shfl_sync(0xffffffffu, slot, 0), slot_hi = __shfl_sync(0xffffffffu, slot, 31);
#pragma unroll
                    for (int c = 0; c < 4; ++c) {
                        const uint32_t col0 = h * 64u + (uint32_t)c * 16u;
                        uint32_t r[16];
                        tmem_ld16(tD + col0, r);
                        tmem_ld_wait();
                        float x[16], y[16];
                        uint32_t m = 0;
#pragma unroll
                        for (int i = 0; i < 16; ++i) {
                            const float v = __uint_as_float(r[i]) + __ldg(db + col0 + i);
                            m |= (v > 0.f ? 1u : 0u) << i;
                            x[i] = fmaxf(v, 0.f);  // H4
                        }
                        // d wc[ch][k] = sum_s d z_ch[s] H4[s,k]
#pragma unroll
                        for (int i = 0; i < 16; ++i) y[i] = x[i] * g4.y;
                        acc_wc[0][c] += warp_colsum16(y, lane);
#pragma unroll
                        for (int i = 0; i < 16; ++i) y[i] = x[i] * g4.z;
                        acc_wc[1][c] += warp_colsum16(y, lane);
#pragma unroll
                        for (int i = 0; i < 16; ++i) y[i] = x[i] * g4.w;
                        acc_wc[2][c] += warp_colsum16(y, lane);
                        // dA4 = (d z . wc) * (H4 > 0)
#pragma unroll
                        for (int i = 0; i < 16; ++i) {
                            const float dh = fmaf(g4.w, wc[256 + col0 + i], fmaf(g4.z, wc[128 + col0 + i], g4.y * wc[col0 + i]));
                            x[i] = ((m >> i) & 1u) ? dh : 0.f;
                        }
                        store_operand16(sDA + rowoff, r7, col0, 32768u, x);
                        // gradient at the per-ray direction bias (-> b4 and W4[:, :27] in k_dirbias_grads): column sums per ray
                        for (uint32_t sl = slot_lo; sl <= slot_hi; ++sl) {
#pragma unroll
                            for (int i = 0; i < 16; ++i) y[i] = (slot == sl && valid) ? x[i] : 0.f;
                            const float cs = warp_colsum16(y, lane);
                            if (lane < 16) atomicAdd(p.g_dirbias + (size_t)sl * 128 + col0 + (uint32_t)lane, cs);
                        }
                    }
                    ops_arrive(true);
                }
                // ---------- E5..E7: dA3, dA2, dA1 = dH * (H > 0); bias gradients ----------
                // the arithmetic runs as soon as dH is there, under the dW MMAs of the layer above; only the stores wait for them
#pragma unroll
                for (int l = 2; l >= 0; --l) {
                    d_wait();
                    const unsigned long long mk = mask[l];
                    uint32_t ph[4][8], pl[4][8];
#pragma unroll
                    for (int c = 0; c < 4; ++c) {
                        const uint32_t col0 = h * 64u + (uint32_t)c * 16u;
                        uint32_t r[16];
                        tmem_ld16(tD + col0, r);
                        tmem_ld_wait();
                        float x[16];
                        const uint32_t m = (uint32_t)(mk >> (16 * c)) & 0xFFFFu;
#pragma unroll
                        for (int i = 0; i < 16; ++i) {
                            float dh = __uint_as_float(r[i]);
                            if (l == 2) dh = fmaf(g4.x, wd[col0 + i], dh);  // the density head reads H3 as well
                            x[i] = ((m >> i) & 1u) ? dh : 0.f;
                        }
                        pack16(x, ph[c], pl[c]);
                        acc_b[l][c] += warp_colsum16(x, lane);  // (consumes x)
                    }
                    mbar_wait_backoff(&bars[BB_DW_DONE], p_dw, 32);  // the dW MMAs of the layer above have read dA (and H)
                    p_dw ^= 1u;
#pragma unroll
                    for (int c = 0; c < 4; ++c) store_packed16(sDA + rowoff, r7, h * 64u + (uint32_t)c * 16u, 32768u, ph[c], pl[c]);
                    ops_arrive(true);
                }
                // ---------- E8: dX -> field gradient (interpolate_values_backward, tetrahedra_tracer.cu:231-247) ----------
                {
                    d_wait();
                    uint32_t r[32];
                    tmem_ld32(tD + h * 32u, r);
                    tmem_ld_wait();
                    ops_arrive(false);  // D has been read: the dW1 MMAs may overwrite it
                    if (valid) {
                        const uint4 v = __ldg(p.vi + my_row);
                        if (v.x != TN_EMPTY) {
                            const float b0 = __ldg(p.bary + 3 * my_row), b1 = __ldg(p.bary + 3 * my_row + 1), b2 = __ldg(p.bary + 3 * my_row + 2);
                            const float w0 = 1.0f - ((b0 + b1) + b2);
                            const uint32_t vs[4] = {v.x, v.y, v.z, v.w};
                            const float ws[4] = {w0, b0, b1, b2};
#pragma unroll
                            for (int k = 0; k < 4; ++k) {
                                float *dst = p.gshadow + (size_t)vs[k] * 64 + h * 32u;
#pragma unroll
                                for (int j = 0; j < 8; ++j)
                                    red_add_v4(dst + 4 * j, ws[k] * __uint_as_float(r[4 * j]), ws[k] * __uint_as_float(r[4 * j + 1]),
                                               ws[k] * __uint_as_float(r[4 * j + 2]), ws[k] * __uint_as_float(r[4 * j + 3]));
                            }
                        }
                    }
                }
                // ---------- E9: this tile's dW1 (TMEM lanes = output features) ----------
                {
                    d_wait();
                    uint32_t r[32];
                    tmem_ld32(tD + h * 32u, r);
                    tmem_ld_wait();
                    ops_arrive(false);
                    float *dst = p.gw + GW_W1 + (size_t)Rt * 64 + h * 32u;
#pragma unroll
                    for (int j = 0; j < 8; ++j)
                        red_add_v4(dst + 4 * j, __uint_as_float(r[4 * j]), __uint_as_float(r[4 * j + 1]), __uint_as_float(r[4 * j + 2]), __uint_as_float(r[4 * j + 3]));
                }
            }
            // ---------- flush: the TMEM-resident weight gradients and the per-lane column sums ----------
            if (ntl != 0) {
                fence_after_sync();
#pragma unroll 1
                for (uint32_t L = 0; L < 3; ++L) {  // TMEM regions dW4b, dW3, dW2
                    float *gdst = p.gw + (L == 0 ? GW_W4B : (L == 1 ? GW_W3 : GW_W2)) + (size_t)Rt * 128 + h * 64u;
#pragma unroll 1
                    for (uint32_t c = 0; c < 2; ++c) {
                        uint32_t r[32];
                        tmem_ld32(tD + 128u * (L + 1u) + h * 64u + c * 32u, r);
                        tmem_ld_wait();
#pragma unroll
                        for (int j = 0; j < 8; ++j)
                            red_add_v4(gdst + c * 32u + 4 * j, __uint_as_float(r[4 * j]), __uint_as_float(r[4 * j + 1]), __uint_as_float(r[4 * j + 2]), __uint_as_float(r[4 * j + 3]));
                    }
                }
                if (lane < 16) {
#pragma unroll
                    for (int c = 0; c < 4; ++c) {
                        const uint32_t col = h * 64u + (uint32_t)c * 16u + (uint32_t)lane;
                        atomicAdd(p.gw + GW_B1 + col, acc_b[0][c]);
                        atomicAdd(p.gw + GW_B2 + col, acc_b[1][c]);
                        atomicAdd(p.gw + GW_B3 + col, acc_b[2][c]);
                        atomicAdd(p.gw + GW_WD + col, acc_wd[c]);
                        atomicAdd(p.gw + GW_WC + col, acc_wc[0][c]);
                        atomicAdd(p.gw + GW_WC + 128 + col, acc_wc[1][c]);
                        atomicAdd(p.gw + GW_WC + 256 + col, acc_wc[2][c]);
                    }
                }
            }
        }
    }
    fence_before_sync();
    __syncthreads();
    if (warp == 8) tmem_dealloc(tbase, 512);
}

}  // namespace tn
