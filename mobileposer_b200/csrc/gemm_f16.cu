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
//   warps 2-5   epilogue: tcgen05.ld both accumulators, main + corr * 2^-11 + bias, 32 x 32 boxes through shared memory, TMA store
// Shared-memory traffic per stage (what bounds gemm_tc.cu): 48 KB landed + 72 KB operand fetch against 768 clk of MMA.
#include "mp_common.cuh"

#include <cuda.h>
#include <cuda_fp16.h>

#include <cstdlib>
#include <cstring>

namespace mp {

namespace {

constexpr int HB_M = 128, HB_N = 256, HB_K = 32, HB_STAGES = 4;
constexpr int HB_THREADS = 192;
constexpr uint32_t HA_TILE = HB_M * HB_K * 2;   // 8 KiB
constexpr uint32_t HW_TILE = HB_N * HB_K * 2;   // 16 KiB
constexpr uint32_t H_STAGE = 2 * HA_TILE + 2 * HW_TILE;
constexpr uint32_t H_SMEM = HB_STAGES * H_STAGE + 4 * 4096 /*epilogue boxes*/ + 1024 /*align*/ + 256 /*barriers*/;
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

// Persistent with a DYNAMIC tile queue: gridDim.x CTAs draw 128 x 256 output tiles from a global counter (column tile fastest, so
// the CTAs running at the same time share A rows in L2).  Dynamic because this kernel rarely has the GPU to itself: the cluster
// recurrences of other heads / batches hold SMs for a millisecond at a time, a CTA that gets its SM late must simply find less work
// left, not a statically assigned share.  The producer warp draws the tile ids and hands them to the MMA and epilogue warps through
// a two-entry shared-memory queue.  The TMA producer's stage ring runs on across tile boundaries, so the first K stages of the next tile land
// while the epilogue drains the accumulators; the epilogue stores through shared memory (32 x 32 fp32 boxes, 128B swizzle) with TMA
// so every store is whole 128-byte rows (thread-per-row fp32 stores cost one L1 wavefront per 16 bytes: 8 k clk per tile, more than
// the K = 256 main loop itself -- ncu on the first version: tensor pipe 29 % active).
__global__ void __launch_bounds__(HB_THREADS, 1)
gemm_f16x3_kernel(const __grid_constant__ CUtensorMap map_a_hi, const __grid_constant__ CUtensorMap map_a_lo,
                  const __grid_constant__ CUtensorMap map_w, const __grid_constant__ CUtensorMap map_c,
                  const float* __restrict__ bias, int M, int N, int K, int N_out, unsigned int* __restrict__ sched) {
    extern __shared__ unsigned char smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t s_out = base + HB_STAGES * H_STAGE;                 // 4 epilogue warps x one 32 x 32 fp32 box (4 KiB)
    const uint32_t bars = s_out + 4 * 4096;
    auto bar_full = [&](int s) { return bars + 8u * s; };
    auto bar_free = [&](int s) { return bars + 8u * (HB_STAGES + s); };
    const uint32_t bar_accum = bars + 8u * (2 * HB_STAGES);            // MMAs of a tile committed
    const uint32_t bar_drained = bars + 8u * (2 * HB_STAGES + 1);      // epilogue has read the accumulators
    auto bar_tile_full = [&](int q) { return bars + 8u * (2 * HB_STAGES + 2 + q); };      // tile id q published
    auto bar_tile_free = [&](int q) { return bars + 8u * (2 * HB_STAGES + 4 + q); };      // ... and read by the MMA thread + 4 epilogue warps
    const uint32_t tmem_slot = bars + 8u * (2 * HB_STAGES + 6);
    unsigned char* gen = smem_raw + (base - smem_u32(smem_raw));
    volatile int* tile_q = reinterpret_cast<volatile int*>(gen + (tmem_slot + 8 - base));

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int KB = K / HB_K;
    const int tiles_n = N / HB_N, tiles = tiles_n * ((M + HB_M - 1) / HB_M);

    if (tid == 0) {
        for (int s = 0; s < HB_STAGES; ++s) {
            mbar_init(bar_full(s), 1);
            mbar_init(bar_free(s), 1);
        }
        mbar_init(bar_accum, 1);
        mbar_init(bar_drained, 4);
        for (int q = 0; q < 2; ++q) {
            mbar_init(bar_tile_full(q), 1);
            mbar_init(bar_tile_free(q), 5);
        }
        mbar_fence_init_cluster();
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(H_TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *reinterpret_cast<volatile uint32_t*>(gen + (tmem_slot - base));

    if (warp == 0) {
        if (lane == 0) {
            int it = 0;
            for (int j = 0;; ++j) {
                const int q = j & 1;
                if (j >= 2) mbar_wait(bar_tile_free(q), ((j >> 1) - 1) & 1);
                const int t = (int)atomicAdd(sched, 1u);
                tile_q[q] = t < tiles ? t : -1;
                asm volatile("mbarrier.arrive.release.cta.shared::cta.b64 _, [%0];" ::"r"(bar_tile_full(q)) : "memory");
                if (t >= tiles) break;
                const int n0 = (t % tiles_n) * HB_N, m0 = (t / tiles_n) * HB_M;
                for (int kb = 0; kb < KB; ++kb, ++it) {
                    const int s = it % HB_STAGES;
                    if (it >= HB_STAGES) mbar_wait(bar_free(s), ((it / HB_STAGES) - 1) & 1);
                    mbar_arrive_expect_tx(bar_full(s), H_STAGE);
                    const uint32_t st = base + s * H_STAGE;
                    tma_load_2d(st, &map_a_hi, kb * HB_K, m0, bar_full(s));
                    tma_load_2d(st + HA_TILE, &map_a_lo, kb * HB_K, m0, bar_full(s));
                    tma_load_2d(st + 2 * HA_TILE, &map_w, kb * HB_K, n0, bar_full(s));
                    tma_load_2d(st + 2 * HA_TILE + HW_TILE, &map_w, kb * HB_K, N + n0, bar_full(s));
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            // instruction descriptor: D = f32, A = B = f16, both K-major, N = 256, M = 128
            const uint32_t idesc = (1u << 4) | ((uint32_t)(HB_N >> 3) << 17) | ((uint32_t)(HB_M >> 4) << 24);
            int it = 0;
            for (int j = 0;; ++j) {
                mbar_wait(bar_tile_full(j & 1), (j >> 1) & 1);
                const int t = tile_q[j & 1];
                asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar_tile_free(j & 1)) : "memory");
                if (t < 0) break;
                if (j > 0) {                               // the epilogue must have drained the previous tile's accumulators
                    mbar_wait(bar_drained, (j - 1) & 1);
                    tc_fence_after();
                }
                for (int kb = 0; kb < KB; ++kb, ++it) {
                    const int s = it % HB_STAGES;
                    mbar_wait(bar_full(s), (it / HB_STAGES) & 1);
                    tc_fence_after();
                    const uint32_t a_hi = base + s * H_STAGE, a_lo = a_hi + HA_TILE;
                    const uint32_t w_hi = a_hi + 2 * HA_TILE, w_lo = w_hi + HW_TILE;
#pragma unroll
                    for (int p = 0; p < 3; ++p) {
                        const uint32_t a = (p == 0) ? a_lo : a_hi;
                        const uint32_t w = (p == 1) ? w_lo : w_hi;
                        const uint32_t d = (p == 2) ? tmem : tmem + HB_N;
#pragma unroll
                        for (int k2 = 0; k2 < HB_K / 16; ++k2)
                            umma_f16(d, umma_desc_sw64(a + k2 * 32), umma_desc_sw64(w + k2 * 32), idesc,
                                     (kb | (p == 1 ? 1 : 0) | k2) != 0 ? 1u : 0u);
                    }
                    tc_commit(bar_free(s));
                }
                tc_commit(bar_accum);
            }
        }
    } else {
        // ---- epilogue: TMEM -> registers -> main + corr * 2^-11 + bias -> swizzled shared-memory box -> TMA store -------
        const int wq = warp & 3;                         // TMEM lane quarter this warp may read
        const uint32_t box = s_out + (uint32_t)wq * 4096u;
        unsigned char* gbox = gen + (box - base);
        for (int j = 0;; ++j) {
            mbar_wait(bar_tile_full(j & 1), (j >> 1) & 1);
            const int t = tile_q[j & 1];
            __syncwarp();
            if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar_tile_free(j & 1)) : "memory");
            if (t < 0) break;
            const int n0 = (t % tiles_n) * HB_N, m0 = (t / tiles_n) * HB_M;
            mbar_wait(bar_accum, j & 1);
            tc_fence_after();
            // narrow outputs (linear2: N_out = 72 / 96 columns of a zero-padded 256-row weight tile): only the boxes that hold real
            // columns are read and stored; the store's tensor map clips the last one
            const int c_end = min(HB_N, ((N_out - n0 + 31) >> 5) << 5);
            for (int c0 = 0; c0 < c_end; c0 += 32) {
                uint32_t v[32], u[32];
                const uint32_t taddr = tmem + ((uint32_t)(wq * 32) << 16) + (uint32_t)c0;
                tmem_ld32(taddr, v);
                tmem_ld32(taddr + (uint32_t)HB_N, u);
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                if (c0 + 32 == c_end) {                  // last read of this tile's accumulators: the MMA warp may start the next tile
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar_drained) : "memory");
                }
                // the box is free once the previous TMA store has read it
                if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
                __syncwarp();
                const float4* b4 = reinterpret_cast<const float4*>(bias + n0 + c0);
#pragma unroll
                for (int q = 0; q < 8; ++q) {
                    const float4 b = __ldg(b4 + q);
                    float4 o;
                    o.x = fmaf(__uint_as_float(u[4 * q + 0]), kLoInv, __uint_as_float(v[4 * q + 0])) + b.x;
                    o.y = fmaf(__uint_as_float(u[4 * q + 1]), kLoInv, __uint_as_float(v[4 * q + 1])) + b.y;
                    o.z = fmaf(__uint_as_float(u[4 * q + 2]), kLoInv, __uint_as_float(v[4 * q + 2])) + b.z;
                    o.w = fmaf(__uint_as_float(u[4 * q + 3]), kLoInv, __uint_as_float(v[4 * q + 3])) + b.w;
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
    if (warp == 1) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(H_TMEM_COLS) : "memory");
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

}  // namespace

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
    MP_REQUIRE(A_split && W_split && bias && C && sched && M > 0, "gemm_f16: bad arguments");
    MP_REQUIRE(N_out > 0 && N_out <= N && (N_out & 3) == 0, "gemm_f16: N_out=%d must be a positive multiple of 4, at most N=%d", N_out, N);
    MP_REQUIRE(N % HB_N == 0 && K % HB_K == 0, "gemm_f16: N=%d must be a multiple of %d and K=%d of %d", N, HB_N, K, HB_K);
    MP_REQUIRE(((uintptr_t)A_split & 15) == 0 && ((uintptr_t)W_split & 15) == 0 && ((uintptr_t)C & 15) == 0 && ((uintptr_t)bias & 15) == 0,
               "gemm_f16: pointers must be 16-byte aligned");
    const __half* a = reinterpret_cast<const __half*>(A_split);
    alignas(64) CUtensorMap map_a_hi, map_a_lo, map_w, map_c;
    MP_TRY(make_map_c(&map_c, C, M, N_out));
    MP_TRY(make_map_f16(&map_a_hi, a, M, K, HB_M));
    MP_TRY(make_map_f16(&map_a_lo, a + (size_t)M * K, M, K, HB_M));
    MP_TRY(make_map_f16(&map_w, reinterpret_cast<const __half*>(W_split), 2 * N, K, HB_N));
    static bool configured = false;
    if (!configured) {
        MP_CUDA_TRY(cudaFuncSetAttribute(gemm_f16x3_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)H_SMEM));
        configured = true;
    }
    // algorithmic bytes: the operands as the caller holds them (fp32-equivalent: 4 B per element either way) + the output
    ProfileScope prof(N_out < N ? "gemm_f16x3_linear2" : "gemm_f16x3", 4.0 * ((double)N_out * K + N_out + (double)M * K + (double)M * N_out), stream);
    static int n_sm = 0;
    if (!n_sm) {
        int dev = 0;
        MP_CUDA_TRY(cudaGetDevice(&dev));
        MP_CUDA_TRY(cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev));
    }
    const int tiles = (N / HB_N) * ((M + HB_M - 1) / HB_M);
    gemm_f16x3_kernel<<<std::min(tiles, n_sm), HB_THREADS, H_SMEM, stream>>>(map_a_hi, map_a_lo, map_w, map_c, bias, M, N, K, N_out, sched);
    MP_CUDA_TRY(cudaGetLastError());
    count_launch();
    return MP_OK;
}

}  // namespace mp
