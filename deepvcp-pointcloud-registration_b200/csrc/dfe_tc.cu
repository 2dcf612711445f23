// Target-side gather + deep-feature embedding + max-pool on the 5th-generation
// tensor cores (tcgen05 / TMEM), sm_100a only.
//
// Reference: get_cat_feat_tgt.py:53-96 + deep_feat_embedding.py:46-60. The three
// Linear layers have no activation between them (deep_feat_embedding.py:48-50), so
// they are one affine map 35 -> 32; the host collapses them in float64
// (Wc = W3 W2 W1, bc = W3 (W2 b1 + b2) + b3) and this kernel evaluates
//     Y[128 x 32] = X[128 x 40] * Wc^T[40 x 32]
// per tile of 4 candidates x 32 neighbours, K = 32 weighted features + 3 local
// coordinates + 1 (bias) + 4 zero columns, on the tensor cores with a split that keeps FP32-level
// accuracy: X = Xh + Xl with Xh the TF32 truncation and Xl the (13-bit) remainder, W = Wh + Wl;
//     Y = Xh*Wh + Xh*Wl   (kind::tf32)   +   Xl*W   (kind::f16, Xl and W rounded to BF16)
// all accumulated in FP32 in the same TMEM columns. Error: 2^-19 of |x||w| per product (BF16 rounding of Xl and
// of W inside the small term Xl*W, itself 2^-11 of the product); measured
// 4e-6 of the feature scale at the KITTI shape (bar: 1e-3).
//
// The kernel is bound by shared-memory bandwidth (the producers' operand stores, the gathered lines
// coming through L1, and the tensor core's operand reads share one 128 B/clk pipe), so the MMA shapes and
// operand formats are chosen for the fewest operand bytes per tile: the GATHERED ROWS are the M-side operand
// (M = 128, read once per MMA) and the N-side operand is the 64-row image [Wh ; Wl]: ONE N = 64 MMA forms
// Xh*Wh (TMEM columns 0..31) and Xh*Wl (columns 32..63) from a single read of Xh; Xl is stored as BF16 (2
// bytes per element on the store AND on the read side) and Xl*W takes three K = 16 MMAs (N = 32) into columns
// 0..31. 8 MMAs and 45 KB of operand reads per tile (round 1's channel-major form: 15 MMAs, 120 KB, tensor pipe
// 61 % busy on 4x the useful work; Xl as TF32: 10 MMAs, 55 KB, 0.83 ms against 0.77 ms now).
// The accumulator is row-major (TMEM lane = neighbour, column = channel): an epilogue warp per candidate
// adds the two column halves in registers and takes the max over its 32 lanes with a transposing butterfly
// (16 shuffles per 16 channels).
//
// Roles in a CTA (18 warps, one CTA per SM at a time -- 8 CTAs per SM slot for large clouds, see the launcher --,
// 3 operand stages + 3 TMEM accumulators of 64 columns):
//   warps 0-3   epilogue: warp w owns TMEM lanes 32w..32w+31 = candidate w of the tile: tcgen05.ld of
//               16 + 16 columns at a time, add, max over the lanes, one coalesced 128-byte store;
//   warps 4-15  producers, 3 groups of 4 (group g builds tiles g, g+3, ... into stage g): gather the
//               neighbours' feature rows into registers, scale by the distance weights (float64 in the
//               reference, carried as float pairs), split hi/lo, store into shared memory in the UMMA
//               K-major SWIZZLE_128B layout;
//   warp 16     allocates TMEM, waits for the operand, issues the 8 tcgen05.mma of a tile from
//               one lane and commits to the mbarriers;
//   warp 17     loader: the KNN indices / distances of the next run of 8 tiles (two contiguous 4 KB pieces)
//               come in by cp.async.bulk (TMA) into a double buffer, completion counted on an mbarrier.
// TMA feeds what is contiguous: the index / distance stream. The feature rows are gathered (index-driven) AND
// scaled by a per-candidate, per-feature weight before the contraction (quirk Q7), so they pass through
// registers: a bulk copy (UBLKCP takes its addresses from uniform registers: one 128-byte row per
// instruction, 32 per candidate) or a tile::gather4 tensor copy could only land them in a staging buffer that
// the same warps would have to read again -- more traffic on the shared-memory / L1 data pipe, which is the
// binding resource of this kernel (ncu: LSU + tensor-core wavefronts fill 91 % of its cycles).
// The operands reach the tensor core through shared-memory matrix descriptors.
#include <cuda_bf16.h>

#include "common.cuh"

namespace dvcp {

constexpr unsigned TC_M = 128;                   // UMMA M = rows of a tile
constexpr int TC_K = 40;                       // padded reduction length
constexpr int TC_ROWS = 128;                   // rows per tile = 4 candidates x 32 neighbours
constexpr int TC_A_SW_BYTES = TC_ROWS * 128;    // A plane, columns 0..31: 128-byte rows, SWIZZLE_128B (K-major)
constexpr int TC_A_TAIL_BYTES = TC_ROWS * 32;   // A plane, columns 32..39: no-swizzle core matrices, K = 8
constexpr int TC_A_BYTES = TC_A_SW_BYTES + TC_A_TAIL_BYTES;   // one A plane (hi or lo)
constexpr int TC_B_BYTES = 32 * TC_K * 4;       // one weight plane as the host lays it out (32 channels)
constexpr int TC_BW_BYTES = TC_B_BYTES;         // the N-side operand is [hi plane ; lo plane]: a 64-row image
constexpr int TC_ACC_COLS = 64;                 // TMEM columns per accumulator: Xh*Wh + Xl*Wh | Xh*Wl
// The low parts of the rows (X - Xh: 13 significant bits) are kept as BF16 (8 bits: the product's error stays at
// 2^-19 of |x||w|, far inside the feature bar): 2 bytes per element instead of 4 on the operand stores and on the
// tensor core's operand reads -- the resource that binds this kernel. Xl * W is then a kind::f16 MMA (K = 16 per
// instruction) into the same accumulator. Layout: no-swizzle K-major core matrices (8 rows x 16 B) with the K
// chunks 160 bytes apart (not 128): the four chunks a half-warp's 8-byte stores touch then start 8 banks apart and
// the store is conflict-free (at 128 all four fall on the same banks, at 192 they collide in pairs: measured).
constexpr int TC_L16_LBO = 160, TC_L16_KCHUNKS = 6;          // K = 48 = 32 features + 3 coordinates + 1 (bias) + zeros
constexpr int TC_L16_SBO = TC_L16_KCHUNKS * TC_L16_LBO;      // 8-row groups 960 bytes apart
constexpr int TC_L16_BYTES = (TC_ROWS / 8) * TC_L16_SBO;     // 15 KB per stage
constexpr int TC_W16_BYTES = 4 * TC_L16_SBO;                 // the 32 weight rows in the same layout
__host__ __device__ constexpr int tc_l16_off(int r, int k) {
    return (r >> 3) * TC_L16_SBO + (k >> 3) * TC_L16_LBO + (r & 7) * 16 + (k & 7) * 2;
}
static_assert(TC_A_SW_BYTES + TC_L16_BYTES + TC_A_TAIL_BYTES <= 2 * TC_A_BYTES, "stage layout");
#ifndef DVCP_DFE_GROUPS
#define DVCP_DFE_GROUPS 3
#endif
constexpr int TC_STAGES = DVCP_DFE_GROUPS;     // shared-memory A stages == TMEM accumulators
constexpr int TC_GROUPS = DVCP_DFE_GROUPS;     // producer groups of 4 warps (one candidate per warp); group g owns stage g
constexpr unsigned TC_TMEM_COLS = TC_STAGES * 64 <= 256 ? 256u : 512u;   // allocation: a power of two
constexpr int TC_EPI_WARPS = 4;                // warp w reads TMEM lanes 32w..32w+31 = the neighbours of candidate w
constexpr int TC_PROD_WARPS = 4 * TC_GROUPS;   // warps 4 .. 4 + TC_PROD_WARPS - 1
constexpr int TC_MMA_WARP = TC_EPI_WARPS + TC_PROD_WARPS;
constexpr int TC_LOAD_WARP = TC_MMA_WARP + 1;   // issues the bulk (TMA) copies of the neighbour index / distance stream
constexpr int TC_THREADS = (TC_LOAD_WARP + 1) * 32;
constexpr int TC_W_BYTES = TC_PROD_WARPS * 32 * 8;   // per producer warp: (hi, lo) distance weight of each feature
constexpr int TC_RUN = 8;                      // a CTA takes its tiles in runs of 8 consecutive tiles = 32 consecutive candidates
constexpr int TC_T_BYTES = 32 * 33 * 4;        // epilogue transposition tile (feature-major output)
constexpr int TC_RUN_BYTES = TC_RUN * 4 * 32 * 4;   // the idx (or dist) rows of one run of tiles: 32 candidates x 32 neighbours
constexpr int TC_SMEM = TC_STAGES * 2 * TC_A_BYTES + 2 * TC_BW_BYTES + TC_W_BYTES + 256 + TC_T_BYTES + 4 * TC_RUN_BYTES + TC_W16_BYTES + 1024;   // + alignment slack

// first candidate of this CTA's i-th tile: runs of TC_RUN consecutive tiles, the runs dealt round-robin to the CTAs
__device__ __forceinline__ int64_t tc_tile_cand(int i) {   // (tile counts fit 32 bits: checked by the launcher)
    const unsigned t = ((unsigned)i / TC_RUN * gridDim.x + blockIdx.x) * TC_RUN + ((unsigned)i % TC_RUN);
    return (int64_t)t * 4;
}

// byte offset of element (row r, column k) in the K-major, no-swizzle canonical layout:
// 8x(16 B) core matrices; core (r/8, k/4) at ((r/8) * (K/4) + k/4) * 128 B.
__host__ __device__ constexpr int tc_off(int r, int k) {
    return ((r >> 3) * (TC_K / 4) + (k >> 2)) * 128 + (r & 7) * 16 + (k & 3) * 4;
}
constexpr unsigned TC_LBO = 128;                  // next core matrix along K
constexpr unsigned TC_SBO = (TC_K / 4) * 128;     // next core matrix along M / N

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ uint64_t make_desc(uint32_t saddr) {
    return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)(TC_LBO >> 4) << 16) | ((uint64_t)(TC_SBO >> 4) << 32) |
           (1ull << 46);   // version 1 (Blackwell), base offset 0, SWIZZLE_NONE
}
// A, columns 0..31: rows of 128 bytes, the 16-byte chunk c of row r stored at chunk c ^ (r & 7)
// (SWIZZLE_128B, 8-row atoms of 1024 bytes); a K = 8 step advances the start address by 32 bytes.
__device__ __forceinline__ uint64_t make_desc_sw128(uint32_t saddr) {
    return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)1 << 16) | ((uint64_t)(1024u >> 4) << 32) | (1ull << 46) |
           (2ull << 61);
}
// A, columns 32..39: K = 8 in the no-swizzle layout (two core matrices per 8 rows)
__host__ __device__ constexpr int tc_tail_off(int r, int k) {
    return (r >> 3) * 256 + (k >> 2) * 128 + (r & 7) * 16 + (k & 3) * 4;
}
__device__ __forceinline__ uint64_t make_desc_tail(uint32_t saddr) {
    return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)(128u >> 4) << 16) | ((uint64_t)(256u >> 4) << 32) |
           (1ull << 46);
}
__device__ __forceinline__ uint64_t make_desc_l16(uint32_t saddr) {
    return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)((unsigned)TC_L16_LBO >> 4) << 16) |
           ((uint64_t)((unsigned)TC_L16_SBO >> 4) << 32) | (1ull << 46);
}
// kind::f16 with BF16 operands, FP32 accumulate, K-major, M = 128, N = 32
constexpr uint32_t TC_IDESC_BF16 = (1u << 4) | (1u << 7) | (1u << 10) | ((32u >> 3) << 17) | ((TC_M >> 4) << 24);
// kind::tf32, FP32 accumulate, A and B K-major, M = 128 (rows of the tile), N = 64 ([Wh ; Wl]) or 32 (Wh)
__host__ __device__ constexpr uint32_t tc_idesc(unsigned n) { return (1u << 4) | (2u << 7) | (2u << 10) | ((n >> 3) << 17) | ((TC_M >> 4) << 24); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, unsigned parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_LOOP:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra WAIT_DONE;\n\t"
        "bra WAIT_LOOP;\n\t"
        "WAIT_DONE:\n\t}" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
template <unsigned NCOLS>
__device__ __forceinline__ void umma_tf32(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
        "l"(adesc), "l"(bdesc), "r"(tc_idesc(NCOLS)), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void umma_bf16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, 1, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
        "l"(adesc), "l"(bdesc), "r"(TC_IDESC_BF16)
        : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t *bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
                 : "memory");
}

template <bool PF>   // PF: weights indexed by feature (quirk Q7), else by neighbour
__global__ void __launch_bounds__(TC_THREADS, 1)
dfe_tgt_tc_kernel(const float *__restrict__ cand, Cloud txyz, const float *__restrict__ tfeat,
                  const float *__restrict__ kdist, const int32_t *__restrict__ kidx, int N, int64_t total_cand,
                  int64_t Q, const float *__restrict__ Bhi, const float *__restrict__ Blo, float inv_Q,
                  int fm_C, float *__restrict__ out) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    unsigned char *smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);   // SWIZZLE_128B atoms: 1024-byte aligned
    unsigned char *sA = smem;                                      // [stage][hi sw | lo sw | hi tail | lo tail]
    unsigned char *sB = smem + TC_STAGES * 2 * TC_A_BYTES;         // [hi|lo][TC_BW_BYTES]
    float2 *sW = reinterpret_cast<float2 *>(sB + 2 * TC_BW_BYTES);  // [producer warp][32]
    uint64_t *bars = reinterpret_cast<uint64_t *>(sB + 2 * TC_BW_BYTES + TC_W_BYTES);
    uint64_t *full = bars, *empty = bars + TC_STAGES, *tfull = bars + 2 * TC_STAGES, *tempty = bars + 3 * TC_STAGES;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bars + 4 * TC_STAGES);
    float *sT = reinterpret_cast<float *>(reinterpret_cast<unsigned char *>(bars) + 256);   // [32][33]
    // neighbour indices / distances of a run of tiles, double-buffered, filled by cp.async.bulk (16-byte aligned)
    unsigned char *sRun = reinterpret_cast<unsigned char *>(sT) + TC_T_BYTES + ((16u - ((TC_T_BYTES) & 15u)) & 15u);
    uint64_t *rfull = bars + 4 * TC_STAGES + 1, *rempty = rfull + 2;
    unsigned char *sW16 = sRun + 4 * TC_RUN_BYTES;   // the weights as BF16, N-side operand of the low-part MMAs

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // the layout is row-block major, so the host's two 32-row images back to back are the 64-row image [Wh ; Wl]
    for (int i = threadIdx.x; i < TC_BW_BYTES / 4; i += blockDim.x) {
        reinterpret_cast<float *>(sB)[i] = Bhi[i];
        reinterpret_cast<float *>(sB + TC_BW_BYTES)[i] = Blo[i];
    }
    for (int i = threadIdx.x; i < TC_W16_BYTES / 4; i += blockDim.x) reinterpret_cast<uint32_t *>(sW16)[i] = 0u;
    __syncthreads();
    for (int i = threadIdx.x; i < 32 * TC_K; i += blockDim.x) {
        const int n = i / TC_K, k = i - n * TC_K;
        const float w = reinterpret_cast<const float *>(sB)[tc_off(n, k) / 4] +
                        reinterpret_cast<const float *>(sB + TC_BW_BYTES)[tc_off(n, k) / 4];   // the copies made above
        *reinterpret_cast<__nv_bfloat16 *>(sW16 + tc_l16_off(n, k)) = __float2bfloat16_rn(w);
    }
    if (threadIdx.x == 0) {
        for (int s = 0; s < TC_STAGES; ++s) {
            mbar_init(&full[s], 4);     // one arrive per warp of the producing group
            mbar_init(&empty[s], 1);    // tcgen05.commit
            mbar_init(&tfull[s], 1);    // tcgen05.commit
            mbar_init(&tempty[s], TC_EPI_WARPS);
        }
        for (int r = 0; r < 2; ++r) {
            mbar_init(&rfull[r], 1);               // the loader's arrive.expect_tx; the copies complete the bytes
            mbar_init(&rempty[r], TC_PROD_WARPS);  // every producer warp has read its rows of the run
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == TC_MMA_WARP) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                     "r"(TC_TMEM_COLS)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    // B was written with ordinary stores: make it visible to the tensor core's (async) proxy
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_slot;

    const int64_t ntiles = (total_cand + 3) / 4;
    const int64_t nruns = (ntiles + TC_RUN - 1) / TC_RUN;
    // (the tiles of a last, partial run beyond ntiles are dead: every role skips candidates >= total_cand)
    const int my_tiles = blockIdx.x < nruns ? (int)((nruns - blockIdx.x + gridDim.x - 1) / gridDim.x * TC_RUN) : 0;

    if (warp == TC_MMA_WARP) {
        // ------------------------------ MMA issuer ------------------------------
        const uint32_t a_base = smem_u32(sA), b_w = smem_u32(sB), w16 = smem_u32(sW16);   // [Wh ; Wl], 64 rows; W as BF16
        for (int i = 0; i < my_tiles; ++i) {
            const int s = (int)(i % TC_STAGES);
            const unsigned ph = (unsigned)((i / TC_STAGES) & 1);
            mbar_wait(&full[s], ph);            // A(i) is in shared memory
            mbar_wait(&tempty[s], ph ^ 1);      // accumulator s drained (tile i - TC_STAGES)
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            if (lane == 0) {
                const uint32_t d = tmem_base + (uint32_t)s * TC_ACC_COLS;   // 64 FP32 columns per accumulator
                // stage: [Xh features, SWIZZLE_128B | Xl, BF16 | Xh coordinates + bias]
                const uint32_t a_hi = a_base + (uint32_t)s * 2 * TC_A_BYTES, a_l16 = a_hi + TC_A_SW_BYTES;
                const uint32_t t_hi = a_l16 + TC_L16_BYTES;
#pragma unroll
                for (int ks = 0; ks < 4; ++ks) {
                    const uint32_t ka = (uint32_t)ks * 32;          // 8 floats inside the 128-byte swizzled row
                    const uint32_t kb = (uint32_t)ks * 2 * 128;     // two 16-byte K chunks of W per MMA (K = 8)
                    // M side: the gathered rows; N side: [Wh ; Wl] (columns 0..31 | 32..63)
                    umma_tf32<64>(d, make_desc_sw128(a_hi + ka), make_desc(b_w + kb), ks > 0);
                }
                umma_tf32<64>(d, make_desc_tail(t_hi), make_desc(b_w + 4 * 2 * 128), 1u);
                // Xl (BF16) * W (BF16) into columns 0..31: three K = 16 steps
#pragma unroll
                for (int j = 0; j < 3; ++j)
                    umma_bf16(d, make_desc_l16(a_l16 + (uint32_t)j * 2 * TC_L16_LBO), make_desc_l16(w16 + (uint32_t)j * 2 * TC_L16_LBO));
                umma_commit(&empty[s]);   // shared-memory stage may be rewritten
                umma_commit(&tfull[s]);   // accumulator is complete
            }
            __syncwarp();
        }
    } else if (warp == TC_LOAD_WARP) {
        // ---------- loader: the neighbour indices / distances of a run of 8 tiles (32 candidates) are two
        // contiguous 4 KB pieces of the KNN result: one thread brings them in with cp.async.bulk (TMA, completion
        // counted in bytes on an mbarrier), one run ahead of the producers ----------
        if (lane == 0) {
            const int my_runs = my_tiles / TC_RUN;
            for (int r = 0; r < my_runs; ++r) {
                const int buf = r & 1, use = r >> 1;
                if (use > 0) mbar_wait(&rempty[buf], (unsigned)((use - 1) & 1));   // the producers are done with the buffer
                const int64_t c0 = tc_tile_cand(r * TC_RUN);
                const int64_t left = total_cand - c0;
                const uint32_t bytes = left <= 0 ? 0u : (uint32_t)(left < TC_RUN * 4 ? left : TC_RUN * 4) * 128u;
                const uint32_t bar = smem_u32(&rfull[buf]);
                asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(2u * bytes) : "memory");
                if (bytes) {
                    const uint32_t dst = smem_u32(sRun + (size_t)buf * 2 * TC_RUN_BYTES);
                    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                                 ::"r"(dst), "l"(kidx + c0 * 32), "r"(bytes), "r"(bar) : "memory");
                    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                                 ::"r"(dst + TC_RUN_BYTES), "l"(kdist + c0 * 32), "r"(bytes), "r"(bar) : "memory");
                }
            }
        }
    } else if (warp < TC_EPI_WARPS) {
        // ------------------ epilogue: TMEM -> max over the 32 neighbours -> global ------------------
        // warp = candidate of the tile, lane = neighbour (TMEM lane 32 * warp + lane); columns c and 32 + c are
        // the two partial sums of channel c. After the butterfly lane o holds channel o's max.
        const uint32_t lane_base = ((uint32_t)warp * 32u) << 16;
        for (int i = 0; i < my_tiles; ++i) {
            const int s = (int)(i % TC_STAGES);
            const unsigned ph = (unsigned)((i / TC_STAGES) & 1);
            const int64_t gq = tc_tile_cand(i) + warp;
            mbar_wait(&tfull[s], ph);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            float best = 0.f;
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                uint32_t u[16], v[16];
                const uint32_t taddr = tmem_base + lane_base + (uint32_t)s * TC_ACC_COLS + (uint32_t)h * 16u;
                asm volatile(
                    "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
                    "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                    : "=r"(u[0]), "=r"(u[1]), "=r"(u[2]), "=r"(u[3]), "=r"(u[4]), "=r"(u[5]), "=r"(u[6]), "=r"(u[7]),
                      "=r"(u[8]), "=r"(u[9]), "=r"(u[10]), "=r"(u[11]), "=r"(u[12]), "=r"(u[13]), "=r"(u[14]),
                      "=r"(u[15])
                    : "r"(taddr));
                asm volatile(
                    "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
                    "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                    : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                      "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]),
                      "=r"(v[15])
                    : "r"(taddr + 32u));
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                float y[16];
#pragma unroll
                for (int o = 0; o < 16; ++o) y[o] = __uint_as_float(u[o]) + __uint_as_float(v[o]);
                // transposing butterfly: 16 values x 32 lanes -> lane l holds the max of value l & 15
#pragma unroll
                for (int st = 8; st >= 1; st >>= 1) {
                    const bool up = (lane & st) != 0;
#pragma unroll
                    for (int o = 0; o < st; ++o) {
                        const float send = up ? y[o] : y[o + st];
                        const float keep = up ? y[o + st] : y[o];
                        y[o] = fmaxf(keep, __shfl_xor_sync(0xffffffffu, send, st));
                    }
                }
                const float m = fmaxf(y[0], __shfl_xor_sync(0xffffffffu, y[0], 16));
                if ((lane >> 4) == h) best = m;   // channels 16h .. 16h + 15 live in lanes of the same half
                // (one CREDUX.MAX.F32 per channel instead of the butterfly: fewer instructions, measured slower,
                //  0.852 against 0.824 ms)
            }
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncwarp();
            if (lane == 0) mbar_arrive(&tempty[s]);
            if (fm_C == 0) {
                if (gq < total_cand) out[gq * 32 + lane] = best;
            } else {
                // feature-major inside every block of fm_C candidates (one key-point): the logical [32, C] order
                // cpg.py:34 re-reads (quirk Q4), so that the CPG kernel finds a voxel's 32 values contiguous.
                // The 8 tiles of a run are 32 consecutive candidates: transposed through shared memory, every
                // store instruction then writes one feature of 32 consecutive candidates (coalesced).
                sT[lane * 33 + (int)(i % TC_RUN) * 4 + warp] = best;
                if (i % TC_RUN == TC_RUN - 1) {
                    asm volatile("bar.sync 2, %0;" ::"r"(TC_EPI_WARPS * 32) : "memory");
                    const int64_t gl = tc_tile_cand(i) - (TC_RUN - 1) * 4 + lane;   // lane = candidate of the run
                    if (gl < total_cand) {
                        const int64_t kp = gl / fm_C;
                        float *o = out + kp * 32 * fm_C + (gl - kp * fm_C);
#pragma unroll
                        for (int f = warp * 8; f < warp * 8 + 8; ++f) o[(int64_t)f * fm_C] = sT[f * 33 + lane];
                    }
                    asm volatile("bar.sync 2, %0;" ::"r"(TC_EPI_WARPS * 32) : "memory");
                }
            }
        }
    } else {
        // ---------------------- A producers: group g builds tiles g, g + TC_GROUPS, ... ----------------------
        // A warp builds the 32 rows (neighbours) of one candidate. Feature rows are fetched
        // cooperatively: lane = (rsub = lane >> 3, chunk = lane & 7) loads the 16-byte chunk of row
        // 4g + rsub, so a quarter-warp reads one whole 128-byte line (4 wavefronts per LDG.128) and
        // writes the 8 chunks of one swizzled 128-byte row (conflict-free STS.128).
        const int pw = warp - TC_EPI_WARPS, group = pw >> 2, cw = pw & 3;   // cw: candidate of the tile
        float *wsh = reinterpret_cast<float *>(sW + pw * 32), *wsl = wsh + 32;
        const int rsub = lane >> 3, chunk = lane & 7;
        const bool xyz4 = txyz.ps == 4 && txyz.cs == 1 && txyz.bs % 4 == 0 && (reinterpret_cast<uintptr_t>(txyz.p) & 15) == 0;
        // stage s == group: a group always rebuilds its own stage
        unsigned char *const ahi = sA + (size_t)group * 2 * TC_A_BYTES;
        unsigned char *const al16 = ahi + TC_A_SW_BYTES;                                   // Xl, BF16
        unsigned char *const thi = al16 + TC_L16_BYTES + tc_tail_off(cw * 32 + lane, 0);     // Xh, coordinates + bias
        unsigned char *const tl16 = al16 + tc_l16_off(cw * 32 + lane, 32);                   // Xl, coordinates (K chunk 4)
        // zero padding, written once, never touched again: columns 36..39 of both parts, columns 40..47 of Xl
        *reinterpret_cast<float4 *>(thi + 128) = make_float4(0.f, 0.f, 0.f, 0.f);            // tc_tail_off(r, 4)
        *reinterpret_cast<uint2 *>(tl16 + 8) = make_uint2(0u, 0u);
        *reinterpret_cast<uint4 *>(tl16 + TC_L16_LBO) = make_uint4(0u, 0u, 0u, 0u);
        // Xl feature rows: row r = cw * 32 + 4 g + rsub, my four columns 4 * chunk .. + 3 (8 bytes)
        unsigned char *const lrow = al16 + (cw * 4) * TC_L16_SBO + (chunk >> 1) * TC_L16_LBO + rsub * 16 + (chunk & 1) * 8;
        // row 4 g + rsub of candidate cw: 128-byte rows, the 16-byte chunk c of row r sits at chunk c ^ (r & 7)
        // (SWIZZLE_128B); r & 7 alternates between rsub and rsub + 4, so two bases + immediates address all rows
        unsigned char *const row0 = ahi + (cw * 32 + rsub) * 128 + ((chunk ^ rsub) << 4);
        unsigned char *const row1 = ahi + (cw * 32 + rsub) * 128 + ((chunk ^ (rsub + 4)) << 4);
        int i = group;
        for (unsigned it = 0; i < my_tiles; i += TC_GROUPS, ++it) {
            const int64_t gq = tc_tile_cand(i) + cw;   // this warp's candidate
            const bool live = gq < total_cand;
            // my neighbour's index and distance: from the run buffer the loader filled (a dead candidate reads row 0)
            const int run = i / TC_RUN, buf = run & 1;
            mbar_wait(&rfull[buf], (unsigned)((run >> 1) & 1));
            const unsigned char *rb = sRun + (size_t)buf * 2 * TC_RUN_BYTES + ((i % TC_RUN) * 4 + cw) * 128 + lane * 4;
            const int id = live ? *reinterpret_cast<const int *>(rb) : 0;
            const float djf = live ? *reinterpret_cast<const float *>(rb + TC_RUN_BYTES) : 0.f;
            if ((i + TC_GROUPS) / TC_RUN != run || i + TC_GROUPS >= my_tiles) {   // my last tile of this run
                __syncwarp();
                if (lane == 0) mbar_arrive(&rempty[buf]);
            }
            // cloud of the candidate: gq / Q from the float quotient, corrected by at most one either way
            int b = 0;
            if (live) {
                b = (int)((float)gq * inv_Q);
                const int64_t lo = (int64_t)b * Q;
                b += (gq - lo >= Q) - (gq < lo);
            }
            // ---- gather first (long latency). Every load is unconditional: a dead candidate (beyond total_cand)
            //      has id = 0, b = 0 and reads row 0 -- its rows are built from finite garbage and never stored.
            //      (Predicated loads made the compiler copy every result out of a temporary before the next load
            //      was issued: the eight gathers of a candidate were serialised on their latency.) ----
            float4 f[8];
            const float4 *fbase = reinterpret_cast<const float4 *>(tfeat + (int64_t)b * N * 32) + chunk;
#pragma unroll
            for (int g = 0; g < 8; ++g) {
                const int rid = __shfl_sync(0xffffffffu, id, 4 * g + rsub);
                f[g] = __ldg(fbase + rid * 8);
            }
            // raw coordinate loads are issued here, consumed after the weight arithmetic below
            float px, py, pz;
            if (xyz4) {
                const float4 pp = __ldg(reinterpret_cast<const float4 *>(txyz.p + (int64_t)b * txyz.bs) + id);
                px = pp.x; py = pp.y; pz = pp.z;
            } else {
                px = txyz.at(b, id, 0); py = txyz.at(b, id, 1); pz = txyz.at(b, id, 2);
            }
            const float *cp = cand + (live ? gq : 0) * 3;
            const float cx = __ldg(cp), cy = __ldg(cp + 1), cz = __ldg(cp + 2);
            // ---- w = dist / sum(dist), float64 in the reference (get_cat_feat_tgt.py:57-58), carried as a float
            //      pair (whi + wlo = w to ~2^-43): the product with a float32 feature is then
            //      float32(double(f) * w) up to one rounding far below float32. Formed with error-free float32
            //      transformations: the FP64 unit of this part is a slow shared pipe (ncu: 'math pipe throttle' on
            //      every DADD / DFMA of the float64 form, a fifth of the producers' time) ----
#ifdef DVCP_DFE_W64
            const double dj = (double)djf;
            double sum = dj;
#pragma unroll
            for (int o = 16; o; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
            const double wl = live ? dj / sum : 0.0;
            const float whi = (float)wl, wlo = (float)(wl - (double)whi);
#else
            float sh = djf, sl = 0.f;   // sum of the 32 distances as sh + sl (TwoSum per butterfly round: symmetric, so
#pragma unroll                      // every lane ends with the same pair)
            for (int o = 16; o; o >>= 1) {
                const float oh = __shfl_xor_sync(0xffffffffu, sh, o), ol = __shfl_xor_sync(0xffffffffu, sl, o);
                const float t = __fadd_rn(sh, oh), bb = __fsub_rn(t, sh);
                const float e = __fadd_rn(__fsub_rn(sh, __fsub_rn(t, bb)), __fsub_rn(oh, bb));
                sl = __fadd_rn(__fadd_rn(sl, ol), e);
                sh = t;
            }
            const float S = __fadd_rn(sh, sl), Sl = __fsub_rn(sl, __fsub_rn(S, sh));
            const float rcp = __fdividef(1.0f, S);
            const float qh = __fmul_rn(djf, rcp);
            const float rem = __fmaf_rn(-qh, Sl, __fmaf_rn(-qh, S, djf));   // d - qh (S + Sl)
            const float ql = __fmul_rn(rem, rcp);
            float whi = __fadd_rn(qh, ql), wlo = __fsub_rn(ql, __fsub_rn(whi, qh));
            if (!live) whi = wlo = 0.f;
#endif
            // per-feature mode: the weights of my 4 channels 4 * chunk .. 4 * chunk + 3, high parts and low parts as
            // two float4 (pairs of adjacent channels feed the packed FMUL2 / FFMA2 below)
            float4 wqh = make_float4(0.f, 0.f, 0.f, 0.f), wql = wqh;
            if (PF) {
                __syncwarp();
                wsh[lane] = whi;
                wsl[lane] = wlo;
                __syncwarp();
                wqh = *reinterpret_cast<const float4 *>(wsh + chunk * 4);
                wql = *reinterpret_cast<const float4 *>(wsl + chunk * 4);
            }
            // my neighbour's local coordinates + bias column
            const float4 loc = make_float4(px - cx, py - cy, pz - cz, 1.0f);
            mbar_wait(&empty[group], (it & 1) ^ 1);   // the MMAs of this group's previous tile have read the stage
#pragma unroll
            for (int g = 0; g < 8; ++g) {
                float2 h01, h23, l01, l23;   // weight pairs
                if (PF) {
                    h01 = make_float2(wqh.x, wqh.y); h23 = make_float2(wqh.z, wqh.w);
                    l01 = make_float2(wql.x, wql.y); l23 = make_float2(wql.z, wql.w);
                } else {
                    const float rhi = __shfl_sync(0xffffffffu, whi, 4 * g + rsub), rlo = __shfl_sync(0xffffffffu, wlo, 4 * g + rsub);
                    h01 = h23 = make_float2(rhi, rhi);
                    l01 = l23 = make_float2(rlo, rlo);
                }
                // v = fma(f, w_hi, f * w_lo); hi = v truncated to TF32 (exact); lo = v - hi (exact) -- two channels per
                // packed instruction (FMUL2 / FFMA2 of sm_100: the same IEEE operations, half the issue slots)
                const float2 f01 = make_float2(f[g].x, f[g].y), f23 = make_float2(f[g].z, f[g].w);
                const float2 v01 = __ffma2_rn(f01, h01, __fmul2_rn(f01, l01));
                const float2 v23 = __ffma2_rn(f23, h23, __fmul2_rn(f23, l23));
                float4 hv;
                hv.x = __uint_as_float(__float_as_uint(v01.x) & 0xffffe000u);
                hv.y = __uint_as_float(__float_as_uint(v01.y) & 0xffffe000u);
                hv.z = __uint_as_float(__float_as_uint(v23.x) & 0xffffe000u);
                hv.w = __uint_as_float(__float_as_uint(v23.y) & 0xffffe000u);
                const float2 m1 = make_float2(-1.f, -1.f);
                const float2 d01 = __ffma2_rn(make_float2(hv.x, hv.y), m1, v01);   // v - hi, exact
                const float2 d23 = __ffma2_rn(make_float2(hv.z, hv.w), m1, v23);
                unsigned char *const dst = ((g & 1) ? row1 : row0) + g * 512;
                *reinterpret_cast<float4 *>(dst) = hv;
                const __nv_bfloat162 p01 = __floats2bfloat162_rn(d01.x, d01.y), p23 = __floats2bfloat162_rn(d23.x, d23.y);
                *reinterpret_cast<uint2 *>(lrow + (g >> 1) * TC_L16_SBO + (g & 1) * 64) =
                    make_uint2(*reinterpret_cast<const unsigned *>(&p01), *reinterpret_cast<const unsigned *>(&p23));
            }
            {
                float4 hv, lv;
                const float le[4] = {loc.x, loc.y, loc.z, loc.w};
                float *hp = &hv.x, *lp = &lv.x;
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const float t = __uint_as_float(__float_as_uint(le[e]) & 0xffffe000u);
                    hp[e] = t;
                    lp[e] = le[e] - t;
                }
                *reinterpret_cast<float4 *>(thi) = hv;
                const __nv_bfloat162 q01 = __floats2bfloat162_rn(lv.x, lv.y), q23 = __floats2bfloat162_rn(lv.z, lv.w);
                *reinterpret_cast<uint2 *>(tl16) =
                    make_uint2(*reinterpret_cast<const unsigned *>(&q01), *reinterpret_cast<const unsigned *>(&q23));
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            __syncwarp();
            if (lane == 0) mbar_arrive(&full[group]);
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == TC_MMA_WARP) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base),
                     "r"(TC_TMEM_COLS)
                     : "memory");
    }
}

}  // namespace dvcp

using namespace dvcp;

extern "C" int dvcp_dfe_tc_b_floats(void) { return 32 * TC_K; }

// Host helper: position (in floats) of B[n][k] inside the 32 x 40 operand image.
extern "C" int dvcp_dfe_tc_b_offset(int n, int k) {
    if (n < 0 || n >= 32 || k < 0 || k >= TC_K) return DVCP_E_ARG;
    return tc_off(n, k) / 4;
}

extern "C" int dvcp_dfe_tgt_tc(const float *cand, dvcp_cloud_t tgt_xyz, const float *tgt_feat, const float *knn_dist,
                               const int32_t *knn_idx, int B, int N, int64_t Q, const float *b_hi,
                               const float *b_lo, int quirks, int feature_major_c, float *out, dvcp_stream_t stream) {
    if (!cand || !tgt_xyz.base || !tgt_feat || !knn_dist || !knn_idx || !b_hi || !b_lo || !out || B <= 0 || N <= 0 ||
        Q <= 0 || feature_major_c < 0 || (feature_major_c > 0 && Q % feature_major_c != 0))
        return DVCP_E_ARG;
    const int64_t total = (int64_t)B * Q;
    const int64_t ntiles = (total + 3) / 4;
    if (ntiles + TC_RUN * DVCP_NUM_SMS >= (1ll << 29) || (int64_t)N * 8 >= (1ll << 31)) return DVCP_E_UNSUPPORTED;   // 32-bit tile / row arithmetic
    if (((uintptr_t)knn_idx | (uintptr_t)knn_dist) & 15) return DVCP_E_UNSUPPORTED;   // bulk copies: 16-byte aligned sources
    DVCP_CUDA(cudaFuncSetAttribute(dfe_tgt_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM));
    DVCP_CUDA(cudaFuncSetAttribute(dfe_tgt_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM));
    // several CTAs per SM slot (common.cuh) when the clouds are large enough for one-CTA-per-cloud sampling kernels to
    // hold SMs for milliseconds beside this kernel; small clouds (N <= 2048: 0.1 ms samplings): one persistent CTA per SM
    int64_t grid = (int64_t)DVCP_NUM_SMS * (N > 2048 ? DVCP_DFE_WAVES : 1);
    const int64_t nruns = (ntiles + TC_RUN - 1) / TC_RUN;
    if (grid > nruns) grid = nruns;
    // one cloud stride for the whole batch: tgt_xyz is addressed with b = candidate / Q
    const float inv_Q = 1.0f / (float)Q;
    if ((quirks >> 1) & 1)
        dfe_tgt_tc_kernel<true><<<(unsigned)grid, TC_THREADS, TC_SMEM, (cudaStream_t)stream>>>(
            cand, as_cloud(tgt_xyz), tgt_feat, knn_dist, knn_idx, N, total, Q, b_hi, b_lo, inv_Q, feature_major_c, out);
    else
        dfe_tgt_tc_kernel<false><<<(unsigned)grid, TC_THREADS, TC_SMEM, (cudaStream_t)stream>>>(
            cand, as_cloud(tgt_xyz), tgt_feat, knn_dist, knn_idx, N, total, Q, b_hi, b_lo, inv_Q, feature_major_c, out);
    DVCP_CHECK_LAUNCH();
    return 0;
}
