// Tensor-core input projection, second generation:  C[M,N] = A[M,K] * W[N,K]^T + bias[N]  with fp32-grade accuracy from
// the FP16 tensor-core path ("3xFP16"), twice the TF32 rate and half the operand bytes of gemm_tc.cu.
//
// Replaces the hoisted input projection W_ih x_t + b_ih + b_hh of nn.LSTM (mobileposer/models/rnn.py:27) for M = B*T >= 2048.
//
// Split: every fp32 operand x is carried as two halves,  hi = fp16(x)  and  lo = fp16((x - hi) * 2^11)  -- x - hi is exact in
// fp32, it is at most half an fp16 ulp of x, and the 2^11 scale keeps it out of the fp16 subnormal range for every |x| >= 2^-14
// (below that both halves are subnormal with absolute error <= 2^-36).  hi + lo * 2^-11 carries 22 significant bits, the same as
// the TF32 hi/lo pair of gemm_tc.cu.  Three products per K step,
//     corr += A_lo W_hi + A_hi W_lo        (scaled by 2^11)            main += A_hi W_hi
// in two fp32 accumulators in tensor memory (fp16 x fp16 products are exact in fp32); the epilogue forms
// main + corr * 2^-11 + bias.  The dropped A_lo W_lo term is <= 2^-22 relative.
// Range: |x| must stay below 65504 (fp16 max).  Weights are checked when a head is packed (api.cu); the activations on this
// path are ReLU(linear1) of O(1) inputs and LSTM outputs in (-1, 1).
//
// Both operands arrive ALREADY split (A: [2][M, K] halves from the producing kernel's epilogue or launch_split_f16; W: [2N, K]
// halves made once at pack time), so there are no splitter warps: the kernel is the plain TMA -> tcgen05 pipeline.
//   warp 0      TMA producer: cp.async.bulk.tensor (64B swizzle), 4 boxes per stage (A_hi, A_lo, W_hi, W_lo), K = 32 per stage;
//               the ring runs on across the CTA's tiles (persistent kernel)
//   warp 1      TMEM allocation + single-thread tcgen05.mma.kind::f16 issue: 3 products x 2 K-steps (UMMA 128x256x16) per stage;
//               tcgen05.commit releases the stage; a final commit per tile hands the accumulators to the epilogue
//   warps 2-9   epilogue (two per TMEM lane quarter, half of the columns each): tcgen05.ld both accumulators, main + corr * 2^-11 +
//               bias, 32 x 32 boxes through shared memory, TMA store
// Shared-memory traffic per stage (what bounds gemm_tc.cu): 48 KB landed + 72 KB operand fetch against 768 clk of MMA.
#include "mp_common.cuh"

#include <cuda.h>
#include <cuda_fp16.h>

#include <cstdlib>
#include <cstring>

namespace mp {

namespace {

constexpr int HB_M = 128, HB_N = 256, HB_K = 32, HB_STAGES = 4, HP_STAGES = 6;
constexpr int HB_EPI_WARPS = 8;                // two per TMEM lane quarter, each owns half of the tile's columns
constexpr int HB_THREADS = 64 + 32 * HB_EPI_WARPS;
constexpr uint32_t HA_TILE = HB_M * HB_K * 2;   // 8 KiB
constexpr uint32_t HW_TILE = HB_N * HB_K * 2;   // 16 KiB
constexpr uint32_t H_STAGE = 2 * HA_TILE + 2 * HW_TILE;
static_assert(HP_STAGES * (2 * HA_TILE + HW_TILE) == HB_STAGES * H_STAGE, "both variants use the same ring size");
constexpr uint32_t H_SMEM = HB_STAGES * H_STAGE + HB_EPI_WARPS * 4096 /*epilogue boxes*/ + 1024 /*align*/ + 256 /*barriers*/;
constexpr uint32_t H_TMEM_COLS = 512;   // [0,256) main, [256,512) correction
constexpr float kLoScale = 2048.0f, kLoInv = 1.0f / 2048.0f;

__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, int c0, int c1, uint32_t bar) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(dst),
        "l"(reinterpret_cast<uint64_t>(map)), "r"(c0), "r"(c1), "r"(bar)
        : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
        "}" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// K-major operand slab with 64-byte rows (32 halves), SWIZZLE_64B: 8-row groups are 512 B apart (SBO), LBO unused (1)
__device__ __forceinline__ uint64_t umma_desc_sw64(uint32_t smem_addr) {
    return (uint64_t)((smem_addr & 0x3FFFF) >> 4) | (1ull << 16) | ((uint64_t)(512 >> 4) << 32) | (1ull << 46) | (4ull << 61);
}
// K-major operand slab with 128-byte rows, SWIZZLE_128B: 8-row groups are 1024 B apart (SBO); the start address may point 32 / 64 / 96
// bytes into the row (the K sub-block of the swizzle atom), the hardware applies the XOR on the absolute address bits
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t smem_addr) {
    return (uint64_t)((smem_addr & 0x3FFFF) >> 4) | (1ull << 16) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
          "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]),
          "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]),
          "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr)
        : "memory");
}

__device__ __forceinline__ void tma_store_2d(const CUtensorMap* map, uint32_t src, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(reinterpret_cast<uint64_t>(map)),
                 "r"(src), "r"(c0), "r"(c1)
                 : "memory");
}

// ---- CTA-pair helpers (cta_group::2: the two SMs of a TPC compute one 256 x 256 tile) -----------------------------------------------
__device__ __forceinline__ void tma_load_2d_pair(uint32_t dst, const CUtensorMap* map, int c0, int c1, uint32_t leader_bar) {
    // lands in THIS CTA's shared memory, completes its bytes on the LEADER CTA's mbarrier (shared::cluster address)
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(dst),
        "l"(reinterpret_cast<uint64_t>(map)), "r"(c0), "r"(c1), "r"(leader_bar)
        : "memory");
}
__device__ __forceinline__ void tc_commit_pair(uint32_t bar) {      // arrives on the barrier at this offset in BOTH CTAs of the pair
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar),
                 "h"((unsigned short)3)
                 : "memory");
}
__device__ __forceinline__ void umma_f16_pair(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t"
        "}" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_bar) {     // release at cluster scope, any CTA of the cluster
    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_bar) : "memory");
}
// ... with the default semantics (.release at CTA scope, what CUTLASS's umma_arrive_2x1SM_sm0 uses for the same hand-shake): orders
// this thread's earlier reads (the tile id, the accumulators behind tcgen05.wait::ld + tcgen05.fence) before the arrive without the
// MEMBAR.ALL.GPU the .release.cluster form puts in front of it (see lstm_rec_f16.cu)
__device__ __forceinline__ void mbar_arrive_cluster_cta(uint32_t cluster_bar) {
    asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(cluster_bar) : "memory");
}
__device__ __forceinline__ void mbar_wait_cluster(uint32_t bar, uint32_t parity) {       // acquire at cluster scope (remote arrivals)
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "MP_WAITC_%=:\n\t"
        "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra MP_DONEC_%=;\n\t"
        "bra MP_WAITC_%=;\n\t"
        "MP_DONEC_%=:\n\t"
        "}" ::"r"(bar), "r"(parity)
        : "memory");
}

// Persistent with a DYNAMIC tile queue: the CTAs draw output tiles from a global counter (column tile fastest, so the CTAs running at the
// same time share A rows in L2).  Dynamic because this kernel rarely has the GPU to itself: the cluster recurrences of other heads /
// batches hold SMs for a millisecond at a time, a CTA that gets its SM late must simply find less work left, not a statically assigned
// share.  The producer warp draws the tile ids and hands them to the MMA and epilogue warps through a two-entry shared-memory queue.
// The TMA producer's stage ring runs on across tile boundaries, so the first K stages of the next tile land while the epilogue drains
// the accumulators; the epilogue stores through shared memory (32 x 32 fp32 boxes, 128B swizzle) with TMA so every store is whole
// 128-byte rows.
//
// PAIR = true: two CTAs on the SMs of one TPC (cluster of 2, tcgen05 cta_group::2) compute a 256 x 256 tile.  CTA r holds rows
// [128 r, 128 r + 128) of A and of the accumulators and HALF of the W tile (rows 128 r .. of the 256): the MMA reads the B operand from
// both SMs, so a CTA fetches 32 KB per K stage instead of 48 KB (the L2 -> SM operand stream is what starves the single-CTA kernel) and
// the ring holds 6 stages.  The leader (rank 0) draws the tile, publishes it to both CTAs, collects the bytes of both CTAs' loads on
// its `full` barriers and issues every MMA; tcgen05.commit multicasts the `free` / `accum` arrivals to both CTAs; the peer's consumers
// arrive remotely on the leader's `drained` / `tile_free` barriers.
template <bool PAIR>
__global__ void __launch_bounds__(HB_THREADS, 1)
gemm_f16x3_kernel(const __grid_constant__ CUtensorMap map_a_hi, const __grid_constant__ CUtensorMap map_a_lo,
                  const __grid_constant__ CUtensorMap map_w, const __grid_constant__ CUtensorMap map_c,
                  const __grid_constant__ CUtensorMap map_c_lo, const float* __restrict__ bias, int M, int N, int K, int N_out,
                  int flags, unsigned int* __restrict__ sched, unsigned long long* __restrict__ dbg) {
    constexpr int STAGES = PAIR ? HP_STAGES : HB_STAGES;
    constexpr uint32_t W_TILE = PAIR ? HW_TILE / 2 : HW_TILE;          // W rows per CTA: 128 (pair) / 256
    constexpr uint32_t STAGE = 2 * HA_TILE + 2 * W_TILE;               // bytes landing in THIS CTA per stage
    constexpr int TILE_M = PAIR ? 2 * HB_M : HB_M;
    constexpr uint32_t CONSUMERS = (PAIR ? 2u : 1u) * (1 + HB_EPI_WARPS);
    extern __shared__ unsigned char smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t s_out = base + STAGES * STAGE;                      // one 32 x 32 fp32 box (4 KiB) per epilogue warp
    const uint32_t bars = s_out + HB_EPI_WARPS * 4096;
    auto bar_full = [&](int s) { return bars + 8u * s; };
    auto bar_free = [&](int s) { return bars + 8u * (STAGES + s); };
    const uint32_t bar_accum = bars + 8u * (2 * STAGES);               // MMAs of a tile committed
    const uint32_t bar_drained = bars + 8u * (2 * STAGES + 1);         // the epilogue warps hold the accumulators in registers
    auto bar_tile_full = [&](int q) { return bars + 8u * (2 * STAGES + 2 + q); };      // tile id q published
    auto bar_tile_free = [&](int q) { return bars + 8u * (2 * STAGES + 4 + q); };      // ... and read by every consumer
    const uint32_t tmem_slot = bars + 8u * (2 * STAGES + 6);
    const uint32_t tile_q_addr = tmem_slot + 8;
    unsigned char* gen = smem_raw + (base - smem_u32(smem_raw));
    volatile int* tile_q = reinterpret_cast<volatile int*>(gen + (tile_q_addr - base));

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t rank = PAIR ? cluster_ctarank() : 0u;               // 0 = leader
    const int KB = K / HB_K;
    const int tiles_n = N / HB_N, tiles = tiles_n * ((M + TILE_M - 1) / TILE_M);

    if (tid == 0) {
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(bar_full(s), 1);
            mbar_init(bar_free(s), 1);
        }
        mbar_init(bar_accum, 1);
        mbar_init(bar_drained, (PAIR ? 2 : 1) * HB_EPI_WARPS);
        for (int q = 0; q < 2; ++q) {
            mbar_init(bar_tile_full(q), 1);
            mbar_init(bar_tile_free(q), CONSUMERS);
        }
        mbar_fence_init_cluster();
    }
    if (warp == 1) {
        if (PAIR) {
            asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(H_TMEM_COLS) : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
        } else {
            asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(H_TMEM_COLS) : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
        }
    }
    tc_fence_before();
    __syncthreads();
    if (PAIR) cluster_sync_all();          // both CTAs' barriers exist before any remote arrival / multicast commit
    tc_fence_after();
    const uint32_t tmem = *reinterpret_cast<volatile uint32_t*>(gen + (tmem_slot - base));
    // the leader's barriers the peer's threads arrive on
    const uint32_t lead_drained = PAIR ? mapa_u32(bar_drained, 0) : bar_drained;

    if (warp == 0) {
        if (lane == 0) {
            int it = 0;
            for (int j = 0;; ++j) {
                const int q = j & 1;
                int t;
                if (!PAIR || rank == 0) {
                    if (j >= 2) {
                        if (PAIR) mbar_wait_cluster(bar_tile_free(q), ((j >> 1) - 1) & 1);
                        else mbar_wait(bar_tile_free(q), ((j >> 1) - 1) & 1);
                    }
                    t = (int)atomicAdd(sched, 1u);
                    if (t >= tiles) t = -1;
                    tile_q[q] = t;
                    if (PAIR) {
                        asm volatile("st.shared::cluster.u32 [%0], %1;" ::"r"(mapa_u32(tile_q_addr + 4u * q, 1)), "r"(t) : "memory");
                        mbar_arrive_cluster(mapa_u32(bar_tile_full(q), 1));
                    }
                    asm volatile("mbarrier.arrive.release.cta.shared::cta.b64 _, [%0];" ::"r"(bar_tile_full(q)) : "memory");
                } else {
                    mbar_wait_cluster(bar_tile_full(q), (j >> 1) & 1);
                    t = tile_q[q];
                    mbar_arrive_cluster_cta(mapa_u32(bar_tile_free(q), 0));
                }
                if (t < 0) break;
                const int n0 = (t % tiles_n) * HB_N, m0 = (t / tiles_n) * TILE_M + (int)rank * HB_M;
                const int wrow = n0 + (PAIR ? (int)rank * (HB_N / 2) : 0);       // this CTA's rows of the W tile
                for (int kb = 0; kb < KB; ++kb, ++it) {
                    const int s = it % STAGES;
                    if (it >= STAGES) mbar_wait(bar_free(s), ((it / STAGES) - 1) & 1);
                    const uint32_t st = base + s * STAGE;
                    if (PAIR) {
                        if (rank == 0) mbar_arrive_expect_tx(bar_full(s), 2 * STAGE);
                        const uint32_t fb = mapa_u32(bar_full(s), 0);
                        if (flags & 4) {
                            tma_load_2d_pair(st, &map_a_hi, kb * 2 * HB_K, m0, fb);
                            tma_load_2d_pair(st + 2 * HA_TILE, &map_w, kb * 2 * HB_K, wrow, fb);
                        } else {
                            tma_load_2d_pair(st, &map_a_hi, kb * HB_K, m0, fb);
                            tma_load_2d_pair(st + HA_TILE, &map_a_lo, kb * HB_K, m0, fb);
                            tma_load_2d_pair(st + 2 * HA_TILE, &map_w, kb * HB_K, wrow, fb);
                            tma_load_2d_pair(st + 2 * HA_TILE + W_TILE, &map_w, kb * HB_K, N + wrow, fb);
                        }
                    } else {
                        mbar_arrive_expect_tx(bar_full(s), STAGE);
                        if (flags & 4) {   // interleaved operands: one box per operand, 128-byte rows = (hi[32], lo[32]) of a K block
                            tma_load_2d(st, &map_a_hi, kb * 2 * HB_K, m0, bar_full(s));
                            tma_load_2d(st + 2 * HA_TILE, &map_w, kb * 2 * HB_K, n0, bar_full(s));
                        } else {
                            tma_load_2d(st, &map_a_hi, kb * HB_K, m0, bar_full(s));
                            tma_load_2d(st + HA_TILE, &map_a_lo, kb * HB_K, m0, bar_full(s));
                            tma_load_2d(st + 2 * HA_TILE, &map_w, kb * HB_K, n0, bar_full(s));
                            tma_load_2d(st + 2 * HA_TILE + W_TILE, &map_w, kb * HB_K, N + n0, bar_full(s));
                        }
                    }
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0 && rank == 0) {
            // instruction descriptor: D = f32, A = B = f16, both K-major, N = 256, M = 128 (256 across the pair)
            const uint32_t idesc = (1u << 4) | ((uint32_t)(HB_N >> 3) << 17) | ((uint32_t)(TILE_M >> 4) << 24);
            auto issue = [&](int kb, int s) {
                const uint32_t a_hi = base + s * STAGE, a_lo = a_hi + HA_TILE;
                const uint32_t w_hi = a_hi + 2 * HA_TILE, w_lo = w_hi + W_TILE;
#pragma unroll
                for (int p = 0; p < 3; ++p) {
                    const uint32_t d = (p == 2) ? tmem : tmem + HB_N;
                    const uint32_t acc = (kb | (p == 1 ? 1 : 0)) != 0 ? 1u : 0u;
#pragma unroll
                    for (int k2 = 0; k2 < HB_K / 16; ++k2) {
                        uint64_t da, dw;
                        if (flags & 4) {
                            da = umma_desc_sw128(a_hi + ((p == 0) ? 64u : 0u) + k2 * 32);
                            dw = umma_desc_sw128(w_hi + ((p == 1) ? 64u : 0u) + k2 * 32);
                        } else {
                            da = umma_desc_sw64(((p == 0) ? a_lo : a_hi) + k2 * 32);
                            dw = umma_desc_sw64(((p == 1) ? w_lo : w_hi) + k2 * 32);
                        }
                        if (PAIR) umma_f16_pair(d, da, dw, idesc, (acc | k2) != 0 ? 1u : 0u);
                        else umma_f16(d, da, dw, idesc, (acc | k2) != 0 ? 1u : 0u);
                    }
                }
            };
            int it = 0;
            // MP_GEMM_DBG: where the issuing thread's time goes (cycles waiting for operands / for the epilogue), CTA 0 only
            long long t_full = 0, t_drain = 0, t_tile = 0, c0 = 0;
            unsigned long long ns0 = 0;
            const bool stamp = dbg != nullptr && blockIdx.x == 0;
            if (stamp) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(ns0));
            const long long t_begin = clock64();
            int j_done = 0;
            for (int j = 0;; ++j) {
                if (stamp) c0 = clock64();
                mbar_wait(bar_tile_full(j & 1), (j >> 1) & 1);
                if (stamp) t_tile += clock64() - c0;
                const int t = tile_q[j & 1];
                asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar_tile_free(j & 1)) : "memory");
                if (t < 0) break;
                j_done = j + 1;
                if (j > 0) {                               // the epilogue warps hold the previous tile's accumulators in registers
                    if (stamp) c0 = clock64();
                    if (PAIR) mbar_wait_cluster(bar_drained, (j - 1) & 1);
                    else mbar_wait(bar_drained, (j - 1) & 1);
                    if (stamp) t_drain += clock64() - c0;
                    tc_fence_after();
                }
                for (int kb = 0; kb < KB; ++kb, ++it) {
                    const int s = it % STAGES;
                    if (stamp) c0 = clock64();
                    mbar_wait(bar_full(s), (it / STAGES) & 1);
                    if (stamp) t_full += clock64() - c0;
                    tc_fence_after();
                    issue(kb, s);
                    if (PAIR) tc_commit_pair(bar_free(s)); else tc_commit(bar_free(s));
                }
                if (PAIR) tc_commit_pair(bar_accum); else tc_commit(bar_accum);
            }
            if (stamp) {
                unsigned long long ns1;
                asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(ns1));
                dbg[0] = (unsigned long long)(clock64() - t_begin);
                dbg[1] = (unsigned long long)t_full;
                dbg[2] = (unsigned long long)t_drain;
                dbg[3] = ns1 - ns0;
                dbg[4] = (unsigned long long)j_done;
                dbg[5] = (unsigned long long)t_tile;
            }
        }
    } else {
        // ---- epilogue: TMEM -> registers -> main + corr * 2^-11 + bias -> swizzled shared-memory box -> TMA store -------
        const int wq = warp & 3;                         // TMEM lane quarter this warp may read
        const int ch = (warp - 2) >> 2;                  // which half of the tile's columns
        const uint32_t box = s_out + (uint32_t)(warp - 2) * 4096u;
        unsigned char* gbox = gen + (box - base);
        const uint32_t lead_tile_free[2] = {PAIR ? mapa_u32(bar_tile_free(0), 0) : bar_tile_free(0),
                                            PAIR ? mapa_u32(bar_tile_free(1), 0) : bar_tile_free(1)};
        for (int j = 0;; ++j) {
            if (PAIR && rank != 0) mbar_wait_cluster(bar_tile_full(j & 1), (j >> 1) & 1);
            else mbar_wait(bar_tile_full(j & 1), (j >> 1) & 1);
            const int t = tile_q[j & 1];
            __syncwarp();
            if (lane == 0) {
                if (PAIR) mbar_arrive_cluster_cta(lead_tile_free[j & 1]);
                else asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar_tile_free(j & 1)) : "memory");
            }
            if (t < 0) break;
            const int n0 = (t % tiles_n) * HB_N, m0 = (t / tiles_n) * TILE_M + (int)rank * HB_M;
            mbar_wait(bar_accum, j & 1);
            tc_fence_after();
            // narrow outputs (linear2: N_out = 72 / 96 columns of a zero-padded 256-row weight tile): only the boxes that hold real
            // columns are read and stored; the store's tensor map clips the last one
            const int c_all = min(HB_N, ((N_out - n0 + 31) >> 5) << 5);
            // (1) read this warp's share of both accumulators (4 chunks of 32 columns), combining main + corr * 2^-11 on the way, and
            // hand the accumulators back as soon as they are in registers: the next tile's MMAs run under the rest of the epilogue (bias,
            // activation, shared-memory boxes, TMA stores).  What remains exposed is the read itself: tensor memory delivers ~120 B/clk,
            // 256 KB of accumulators per tile = 2.2 k clk (MP_GEMM_DBG stamps, profiles/r02_gemm_f16_variants.txt)
            float acc[4][32];
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                const int c0 = ch * (HB_N / 2) + c * 32;
                if (c0 < c_all) {
                    uint32_t v[32], u[32];
                    const uint32_t taddr = tmem + ((uint32_t)(wq * 32) << 16) + (uint32_t)c0;
                    tmem_ld32(taddr, v);
                    tmem_ld32(taddr + (uint32_t)HB_N, u);
                    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
                    for (int i = 0; i < 32; ++i) acc[c][i] = fmaf(__uint_as_float(u[i]), kLoInv, __uint_as_float(v[i]));
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) {
                if (PAIR) mbar_arrive_cluster_cta(lead_drained);
                else asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar_drained) : "memory");
            }
            // (2) bias / activation / store, chunk by chunk through the warp's box
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                const int c0 = ch * (HB_N / 2) + c * 32;
                if (c0 >= c_all) continue;
                // the box is free once the previous TMA store has read it
                if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
                __syncwarp();
                const float4* b4 = reinterpret_cast<const float4*>(bias + n0 + c0);
                if (flags & 2) {
                    // output as the NEXT projection's operand (linear1 -> layer-0 W_ih): two [M, N_out] planes of halves, hi = fp16(x),
                    // lo = fp16((x - hi) * 2^11).  Two 32 x 32 boxes of halves per warp (64-byte rows, 64B swizzle: chunk ^ (row / 2) % 4)
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        float o[8];
#pragma unroll
                        for (int h = 0; h < 2; ++h) {
                            const float4 b = __ldg(b4 + 2 * q + h);
                            const int i = 8 * q + 4 * h;
                            o[4 * h + 0] = acc[c][i + 0] + b.x;
                            o[4 * h + 1] = acc[c][i + 1] + b.y;
                            o[4 * h + 2] = acc[c][i + 2] + b.z;
                            o[4 * h + 3] = acc[c][i + 3] + b.w;
                        }
                        uint4 ph, pl;
                        uint32_t* php = reinterpret_cast<uint32_t*>(&ph);
                        uint32_t* plp = reinterpret_cast<uint32_t*>(&pl);
#pragma unroll
                        for (int e = 0; e < 4; ++e) {
                            float x0 = o[2 * e], x1 = o[2 * e + 1];
                            if (flags & 1) { x0 = fmaxf(x0, 0.f); x1 = fmaxf(x1, 0.f); }
                            const __half2 hh = __floats2half2_rn(x0, x1);
                            const float2 hf = __half22float2(hh);
                            const __half2 ll = __floats2half2_rn((x0 - hf.x) * kLoScale, (x1 - hf.y) * kLoScale);
                            php[e] = *reinterpret_cast<const uint32_t*>(&hh);
                            plp[e] = *reinterpret_cast<const uint32_t*>(&ll);
                        }
                        const int pos = lane * 64 + ((q ^ ((lane >> 1) & 3)) << 4);
                        *reinterpret_cast<uint4*>(gbox + pos) = ph;
                        *reinterpret_cast<uint4*>(gbox + 2048 + pos) = pl;
                    }
                    fence_proxy_async_smem();
                    __syncwarp();
                    if (lane == 0) {
                        tma_store_2d(&map_c, box, n0 + c0, m0 + wq * 32);
                        tma_store_2d(&map_c_lo, box + 2048u, n0 + c0, m0 + wq * 32);
                        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                    }
                    continue;
                }
#pragma unroll
                for (int q = 0; q < 8; ++q) {
                    const float4 b = __ldg(b4 + q);
                    float4 o;
                    o.x = acc[c][4 * q + 0] + b.x;
                    o.y = acc[c][4 * q + 1] + b.y;
                    o.z = acc[c][4 * q + 2] + b.z;
                    o.w = acc[c][4 * q + 3] + b.w;
                    if (flags & 1) { o.x = fmaxf(o.x, 0.f); o.y = fmaxf(o.y, 0.f); o.z = fmaxf(o.z, 0.f); o.w = fmaxf(o.w, 0.f); }
                    // row = lane (128 B per row), 16-byte chunk q at position q ^ (row % 8): the 128B swizzle of the store's tensor map
                    *reinterpret_cast<float4*>(gbox + lane * 128 + ((q ^ (lane & 7)) << 4)) = o;
                }
                fence_proxy_async_smem();
                __syncwarp();
                if (lane == 0) {
                    tma_store_2d(&map_c, box, n0 + c0, m0 + wq * 32);
                    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                }
            }
        }
        if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");      // all stores of this CTA are complete
    }
    tc_fence_before();
    __syncthreads();
    if (PAIR) cluster_sync_all();          // the peer's shared memory and barriers stay alive until the leader is done with them
    if (warp == 1) {
        if (PAIR) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(H_TMEM_COLS) : "memory");
        else asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(H_TMEM_COLS) : "memory");
    }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_fn() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
    }
    return fn;
}

// interleaved operand [rows, K/32, {hi[32], lo[32]}] = [rows, 2K] halves -> (64 x box_rows) box with 128-byte rows, 128-byte swizzle
int make_map_f16_il(CUtensorMap* map, const __half* ptr, int rows, int K, int box_rows) {
    EncodeTiledFn fn = encode_fn();
    if (!fn) {
        set_error("gemm_f16: cuTensorMapEncodeTiled is not available from this driver");
        return MP_ERR_CUDA;
    }
    const cuuint64_t dims[2] = {(cuuint64_t)2 * K, (cuuint64_t)rows};
    const cuuint64_t strides[1] = {(cuuint64_t)K * 4};
    const cuuint32_t box[2] = {(cuuint32_t)2 * HB_K, (cuuint32_t)box_rows};
    const cuuint32_t estr[2] = {1, 1};
    const CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<__half*>(ptr), dims, strides, box, estr,
                          CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("gemm_f16: cuTensorMapEncodeTiled (interleaved) failed with CUresult %d (rows=%d K=%d)", (int)r, rows, K);
        return MP_ERR_CUDA;
    }
    return MP_OK;
}

// [rows, K] fp16 row-major -> 2-D tensor map with a (32 x box_rows) box, 64-byte swizzle; rows past the end read as zero
int make_map_f16(CUtensorMap* map, const __half* ptr, int rows, int K, int box_rows) {
    EncodeTiledFn fn = encode_fn();
    if (!fn) {
        set_error("gemm_f16: cuTensorMapEncodeTiled is not available from this driver");
        return MP_ERR_CUDA;
    }
    const cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)rows};
    const cuuint64_t strides[1] = {(cuuint64_t)K * 2};
    const cuuint32_t box[2] = {(cuuint32_t)HB_K, (cuuint32_t)box_rows};
    const cuuint32_t estr[2] = {1, 1};
    const CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<__half*>(ptr), dims, strides, box, estr,
                          CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("gemm_f16: cuTensorMapEncodeTiled failed with CUresult %d (rows=%d K=%d)", (int)r, rows, K);
        return MP_ERR_CUDA;
    }
    return MP_OK;
}

// C [M, N] fp32 row-major -> 2-D tensor map with a 32 x 32 box (128-byte rows), 128-byte swizzle, for the epilogue's TMA stores
int make_map_c(CUtensorMap* map, float* ptr, int M, int N) {
    EncodeTiledFn fn = encode_fn();
    if (!fn) {
        set_error("gemm_f16: cuTensorMapEncodeTiled is not available from this driver");
        return MP_ERR_CUDA;
    }
    const cuuint64_t dims[2] = {(cuuint64_t)N, (cuuint64_t)M};
    const cuuint64_t strides[1] = {(cuuint64_t)N * 4};
    const cuuint32_t box[2] = {32, 32};
    const cuuint32_t estr[2] = {1, 1};
    const CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, ptr, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                          CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("gemm_f16: cuTensorMapEncodeTiled (output) failed with CUresult %d (M=%d N=%d)", (int)r, M, N);
        return MP_ERR_CUDA;
    }
    return MP_OK;
}

// one [M, N] plane of halves -> 2-D tensor map with a 32 x 32 box (64-byte rows), 64-byte swizzle, for the plane-output epilogue
int make_map_c16(CUtensorMap* map, __half* ptr, int M, int N) {
    EncodeTiledFn fn = encode_fn();
    if (!fn) {
        set_error("gemm_f16: cuTensorMapEncodeTiled is not available from this driver");
        return MP_ERR_CUDA;
    }
    const cuuint64_t dims[2] = {(cuuint64_t)N, (cuuint64_t)M};
    const cuuint64_t strides[1] = {(cuuint64_t)N * 2};
    const cuuint32_t box[2] = {32, 32};
    const cuuint32_t estr[2] = {1, 1};
    const CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, ptr, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                          CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("gemm_f16: cuTensorMapEncodeTiled (plane output) failed with CUresult %d (M=%d N=%d)", (int)r, M, N);
        return MP_ERR_CUDA;
    }
    return MP_OK;
}

// x -> (hi, lo): out[i] = fp16(x), out[n + i] = fp16((x - hi) * 2^11).  Streaming: 4 B in, 4 B out per element.
__global__ void split_f16_kernel(const float4* __restrict__ x, size_t n4, __half2* __restrict__ hi, __half2* __restrict__ lo) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) {
        const float4 v = __ldg(x + i);
        const __half2 h0 = __floats2half2_rn(v.x, v.y), h1 = __floats2half2_rn(v.z, v.w);
        const float2 f0 = __half22float2(h0), f1 = __half22float2(h1);
        const __half2 l0 = __floats2half2_rn((v.x - f0.x) * kLoScale, (v.y - f0.y) * kLoScale);
        const __half2 l1 = __floats2half2_rn((v.z - f1.x) * kLoScale, (v.w - f1.y) * kLoScale);
        hi[2 * i] = h0; hi[2 * i + 1] = h1;
        lo[2 * i] = l0; lo[2 * i + 1] = l1;
    }
}

// cat(A1[M,K1], A2[M,K2]) zero-padded to Kp columns -> (hi, lo) planes [2][M, Kp] of halves: the activation operand of linear1
// (rnn.py:22; net.py:106,113 form the concatenation) for the tensor-core kernel.  One thread per 4 columns; K1, K2, Kp % 4 == 0.
__global__ void pack_cat_f16_kernel(const float* __restrict__ A1, int K1, const float* __restrict__ A2, int K2, size_t M, int Kp,
                                    __half* __restrict__ hi, __half* __restrict__ lo) {
    const int g = Kp >> 2;
    const size_t total = M * (size_t)g;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const size_t m = i / g;
        const int k = (int)(i - m * g) << 2;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (k < K1) v = __ldg(reinterpret_cast<const float4*>(A1 + m * K1 + k));
        else if (k < K1 + K2) v = __ldg(reinterpret_cast<const float4*>(A2 + m * K2 + (k - K1)));
        const __half2 h0 = __floats2half2_rn(v.x, v.y), h1 = __floats2half2_rn(v.z, v.w);
        const float2 f0 = __half22float2(h0), f1 = __half22float2(h1);
        const __half2 l0 = __floats2half2_rn((v.x - f0.x) * kLoScale, (v.y - f0.y) * kLoScale);
        const __half2 l1 = __floats2half2_rn((v.z - f1.x) * kLoScale, (v.w - f1.y) * kLoScale);
        uint2 ph, pl;
        ph.x = *reinterpret_cast<const uint32_t*>(&h0); ph.y = *reinterpret_cast<const uint32_t*>(&h1);
        pl.x = *reinterpret_cast<const uint32_t*>(&l0); pl.y = *reinterpret_cast<const uint32_t*>(&l1);
        *reinterpret_cast<uint2*>(hi + m * Kp + k) = ph;
        *reinterpret_cast<uint2*>(lo + m * Kp + k) = pl;
    }
}

// x [rows, K] (K % 32 == 0) -> interleaved [rows, K/32, {hi[32], lo[32]}]: element e lands at 2 * (e & ~31) + (e & 31), its lo 32 further
__global__ void split_f16_il_kernel(const float4* __restrict__ x, size_t n4, __half* __restrict__ out) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) {
        const float4 v = __ldg(x + i);
        const size_t e = 4 * i, o = 2 * (e & ~(size_t)31) + (e & 31);
        const __half2 h0 = __floats2half2_rn(v.x, v.y), h1 = __floats2half2_rn(v.z, v.w);
        const float2 f0 = __half22float2(h0), f1 = __half22float2(h1);
        const __half2 l0 = __floats2half2_rn((v.x - f0.x) * kLoScale, (v.y - f0.y) * kLoScale);
        const __half2 l1 = __floats2half2_rn((v.z - f1.x) * kLoScale, (v.w - f1.y) * kLoScale);
        uint2 ph, pl;
        ph.x = *reinterpret_cast<const uint32_t*>(&h0); ph.y = *reinterpret_cast<const uint32_t*>(&h1);
        pl.x = *reinterpret_cast<const uint32_t*>(&l0); pl.y = *reinterpret_cast<const uint32_t*>(&l1);
        *reinterpret_cast<uint2*>(out + o) = ph;
        *reinterpret_cast<uint2*>(out + o + 32) = pl;
    }
}

}  // namespace

int launch_split_f16_il(const float* x, size_t n, void* out, cudaStream_t stream) {
    MP_REQUIRE(x && out && n % 32 == 0 && ((uintptr_t)x & 15) == 0 && ((uintptr_t)out & 15) == 0,
               "split_f16_il: n must be a multiple of 32 and the pointers 16-byte aligned");
    ProfileScope prof("split_f16", 8.0 * (double)n, stream);
    const int blocks = (int)std::min<size_t>((n / 4 + 255) / 256, (size_t)148 * 8);
    split_f16_il_kernel<<<blocks, 256, 0, stream>>>(reinterpret_cast<const float4*>(x), n / 4, reinterpret_cast<__half*>(out));
    MP_CUDA_TRY(cudaGetLastError());
    count_launch();
    return MP_OK;
}

int launch_pack_cat_f16(const float* A1, int K1, const float* A2, int K2, size_t M, int Kp, void* out_hi_lo, cudaStream_t stream) {
    MP_REQUIRE(A1 && out_hi_lo && M > 0 && K1 > 0 && K2 >= 0 && (K2 == 0 || A2), "pack_cat_f16: bad arguments");
    MP_REQUIRE((K1 & 3) == 0 && (K2 & 3) == 0 && (Kp & 3) == 0 && Kp >= K1 + K2, "pack_cat_f16: K1=%d K2=%d Kp=%d must be multiples of 4, Kp >= K1 + K2", K1, K2, Kp);
    MP_REQUIRE(((uintptr_t)A1 & 15) == 0 && ((uintptr_t)A2 & 15) == 0 && ((uintptr_t)out_hi_lo & 15) == 0, "pack_cat_f16: pointers must be 16-byte aligned");
    __half* hi = reinterpret_cast<__half*>(out_hi_lo);
    ProfileScope prof("pack_cat_f16", 4.0 * (double)M * (K1 + K2) + 4.0 * (double)M * Kp, stream);
    const size_t total = M * (size_t)(Kp >> 2);
    const int blocks = (int)std::min<size_t>((total + 255) / 256, (size_t)148 * 16);
    pack_cat_f16_kernel<<<blocks, 256, 0, stream>>>(A1, K1, A2, K2, M, Kp, hi, hi + M * (size_t)Kp);
    MP_CUDA_TRY(cudaGetLastError());
    count_launch();
    return MP_OK;
}

unsigned long long* g_gemm_dbg = nullptr;      // MP_GEMM_DBG (test entry): 6 words of clock stamps from CTA 0's issuing thread

bool gemm_f16_eligible(int M, int N, int K) {
    const char* v = getenv("MP_GEMM");
    if (v && (strcmp(v, "ffma") == 0 || strcmp(v, "tf32") == 0)) return false;
    const int min_m = (v && strcmp(v, "tc") == 0) ? 1 : 2048;
    return M >= min_m && N % HB_N == 0 && K % HB_K == 0 && K >= HB_K;
}

int launch_split_f16(const float* x, size_t n, void* out_hi_lo, cudaStream_t stream) {
    MP_REQUIRE(x && out_hi_lo && n % 4 == 0 && ((uintptr_t)x & 15) == 0 && ((uintptr_t)out_hi_lo & 15) == 0,
               "split_f16: n must be a multiple of 4 and the pointers 16-byte aligned");
    __half2* hi = reinterpret_cast<__half2*>(out_hi_lo);
    ProfileScope prof("split_f16", 8.0 * (double)n, stream);
    const int blocks = (int)std::min<size_t>((n / 4 + 255) / 256, (size_t)148 * 8);
    split_f16_kernel<<<blocks, 256, 0, stream>>>(reinterpret_cast<const float4*>(x), n / 4, hi, hi + n / 2);
    MP_CUDA_TRY(cudaGetLastError());
    count_launch();
    return MP_OK;
}

// A_split: [2][M, K] halves (hi plane, then lo plane); W_split: [2N, K] halves (hi rows, then lo rows)
// sched: two zeroed 32-bit words owned by this launch (tile counter, CTAs done), e.g. a slice of the caller's workspace cleared on the
// same stream; the kernel leaves them zero again.  N_out <= N: C is [M, N_out] and only its columns are stored (W_split and bias are
// padded to N rows / entries with zeros); N_out % 4 == 0.
int launch_gemm_f16x3(const void* A_split, const void* W_split, const float* bias, float* C, int M, int N, int K, int N_out,
                      unsigned int* sched, cudaStream_t stream) {
    return launch_gemm_f16x3_act(A_split, W_split, bias, C, M, N, K, N_out, 0, sched, stream);
}

// flags bit 0: ReLU; bit 1: C is written as two [M, N_out] planes of halves (fp16 hi, scaled lo) -- the A_split of the next launch
int launch_gemm_f16x3_act(const void* A_split, const void* W_split, const float* bias, void* C_, int M, int N, int K, int N_out,
                          int flags, unsigned int* sched, cudaStream_t stream) {
    float* C = reinterpret_cast<float*>(C_);
    MP_REQUIRE(A_split && W_split && bias && C && sched && M > 0, "gemm_f16: bad arguments");
    MP_REQUIRE(N_out > 0 && N_out <= N && (N_out & 3) == 0, "gemm_f16: N_out=%d must be a positive multiple of 4, at most N=%d", N_out, N);
    MP_REQUIRE(N % HB_N == 0 && K % HB_K == 0, "gemm_f16: N=%d must be a multiple of %d and K=%d of %d", N, HB_N, K, HB_K);
    MP_REQUIRE(((uintptr_t)A_split & 15) == 0 && ((uintptr_t)W_split & 15) == 0 && ((uintptr_t)C & 15) == 0 && ((uintptr_t)bias & 15) == 0,
               "gemm_f16: pointers must be 16-byte aligned");
    const __half* a = reinterpret_cast<const __half*>(A_split);
    alignas(64) CUtensorMap map_a_hi, map_a_lo, map_w, map_c, map_c_lo;
    if (flags & 2) {
        MP_REQUIRE((N_out & 7) == 0, "gemm_f16: plane output needs N_out %% 8 == 0 (got %d)", N_out);
        MP_TRY(make_map_c16(&map_c, reinterpret_cast<__half*>(C), M, N_out));
        MP_TRY(make_map_c16(&map_c_lo, reinterpret_cast<__half*>(C) + (size_t)M * N_out, M, N_out));
    } else {
        MP_TRY(make_map_c(&map_c, C, M, N_out));
        map_c_lo = map_c;
    }
    // the CTA-pair kernel wins 10-12 % where the K loop is long enough to amortise the pair hand-shake (K >= 256: 0.372 vs 0.421 ms at
    // N = 2048, K = 512; slower at K = 64): profiles/r02_gemm_f16_variants.txt.  MP_GEMM_PAIR=0 / 1 forces either.
    static const int pair_env = getenv("MP_GEMM_PAIR") ? (atoi(getenv("MP_GEMM_PAIR")) != 0 ? 1 : 0) : -1;
    const bool pair_default = pair_env >= 0 ? pair_env == 1 : (K >= 256 && M >= 4 * HB_M);
    const bool pair = (flags & 16) || (pair_default && !(flags & 32));
    const int w_box = pair ? HB_N / 2 : HB_N;
    if (flags & 4) {
        MP_TRY(make_map_f16_il(&map_a_hi, a, M, K, HB_M));
        map_a_lo = map_a_hi;
        MP_TRY(make_map_f16_il(&map_w, reinterpret_cast<const __half*>(W_split), N, K, w_box));
    } else {
        MP_TRY(make_map_f16(&map_a_hi, a, M, K, HB_M));
        MP_TRY(make_map_f16(&map_a_lo, a + (size_t)M * K, M, K, HB_M));
        MP_TRY(make_map_f16(&map_w, reinterpret_cast<const __half*>(W_split), 2 * N, K, w_box));
    }
    static bool configured = false;
    if (!configured) {
        MP_CUDA_TRY(cudaFuncSetAttribute(gemm_f16x3_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)H_SMEM));
        MP_CUDA_TRY(cudaFuncSetAttribute(gemm_f16x3_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)H_SMEM));
        configured = true;
    }
    // algorithmic bytes: the operands as the caller holds them (fp32-equivalent: 4 B per element either way) + the output
    ProfileScope prof((flags & 2) ? "gemm_f16x3_linear1" : N_out < N ? "gemm_f16x3_linear2" : "gemm_f16x3", 4.0 * ((double)N_out * K + N_out + (double)M * K + (double)M * N_out), stream);
    static int n_sm = 0;
    if (!n_sm) {
        int dev = 0;
        MP_CUDA_TRY(cudaGetDevice(&dev));
        MP_CUDA_TRY(cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev));
    }
    const int kflags = flags & 7;
    if (pair) {
        const int tiles = (N / HB_N) * ((M + 2 * HB_M - 1) / (2 * HB_M));
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(2 * std::min(tiles, n_sm / 2));
        cfg.blockDim = dim3(HB_THREADS);
        cfg.dynamicSmemBytes = H_SMEM;
        cfg.stream = stream;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr;
        cfg.numAttrs = 1;
        MP_CUDA_TRY(cudaLaunchKernelEx(&cfg, gemm_f16x3_kernel<true>, map_a_hi, map_a_lo, map_w, map_c, map_c_lo, bias, M, N, K, N_out, kflags, sched, g_gemm_dbg));
    } else {
        const int tiles = (N / HB_N) * ((M + HB_M - 1) / HB_M);
        gemm_f16x3_kernel<false><<<std::min(tiles, n_sm), HB_THREADS, H_SMEM, stream>>>(map_a_hi, map_a_lo, map_w, map_c, map_c_lo, bias, M, N, K, N_out, kflags, sched, g_gemm_dbg);
    }
    MP_CUDA_TRY(cudaGetLastError());
    count_launch();
    return MP_OK;
}

}  // namespace mp
