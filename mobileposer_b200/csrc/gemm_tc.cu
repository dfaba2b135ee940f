// Tensor-core input projection for large batches:  C[M,N] = A[M,K] * W[N,K]^T + bias[N]  with fp32-grade
// accuracy from TF32 tensor cores ("3xTF32").
//
// Replaces the hoisted input projection W_ih x_t + b_ih + b_hh of nn.LSTM (mobileposer/models/rnn.py:27) when
// M = B*T is large -- 57 % of the path's FLOPs and the only place where hidden_dim makes
// a real dense contraction.
//
// Why 3xTF32: the parity bar (1e-4 rad / 1e-4 m through 4 chained LSTM stacks) needs fp32-accurate gate
// pre-activations; a single TF32 pass (10-bit mantissa) is ~1e-3 relative.  Each fp32 operand x is split
// EXACTLY into x = hi + lo with hi = x & 0xFFFFE000 (representable in TF32) and lo = x - hi (13 significant
// bits, of which TF32 keeps 11), and  A*W ~= A_lo*W_hi + A_hi*W_lo + A_hi*W_hi  is accumulated in fp32 in
// tensor memory: the dropped terms are <= 2^-21 relative.
//
// Structure (one 128 x 256 output tile per CTA, K walked in 16-float slabs, 4-stage ring):
//   warp 0      TMA producer: cp.async.bulk.tensor (64B swizzle) of the fp32 A and W slabs into shared memory
//   warps 2-5   splitter: rewrite each landed slab in place as `hi` and write `lo` next to it (element-wise, so the
//               TMA swizzle is preserved), fence.proxy.async, arrive on the stage's "split" mbarrier; after the
//               K loop the same warps are the epilogue: tcgen05.ld the accumulator, add the bias, store fp32 rows
//   warp 1      TMEM allocation + single-thread tcgen05.mma issue: 3 products x 2 K-steps (UMMA 128x256x8, kind::tf32)
//               per slab, tcgen05.commit releases the stage to the producer; a final commit hands the accumulator
//               to the epilogue
#include "mp_common.cuh"

#include <cuda.h>

#include <cstdlib>
#include <cstring>

namespace mp {

namespace {

constexpr int TC_BM = 128, TC_BN = 256, TC_BK = 16, TC_STAGES = 4;
constexpr int TC_THREADS = 192;
constexpr uint32_t A_TILE = TC_BM * TC_BK * 4;   // 8 KiB
constexpr uint32_t W_TILE = TC_BN * TC_BK * 4;   // 16 KiB
constexpr uint32_t STAGE_BYTES = 2 * A_TILE + 2 * W_TILE;
constexpr uint32_t TC_SMEM = TC_STAGES * STAGE_BYTES + 1024 /*align*/ + 256 /*barriers*/;
constexpr uint32_t TMEM_COLS = 512;   // two fp32 accumulators: [0,256) main (hi*hi), [256,512) correction (lo*hi + hi*lo)

__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, int c0, int c1, uint32_t bar) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(dst),
        "l"(reinterpret_cast<uint64_t>(map)), "r"(c0), "r"(c1), "r"(bar)
        : "memory");
}
__device__ __forceinline__ void tcgen05_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tcgen05_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tcgen05_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
        "}" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ uint32_t tf32_rna(float x) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
    return r;
}
// K-major operand slab with 64-byte rows, SWIZZLE_64B: 8-row groups are 512 B apart (SBO), LBO unused (1)
__device__ __forceinline__ uint64_t umma_desc_sw64(uint32_t smem_addr) {
    return (uint64_t)((smem_addr & 0x3FFFF) >> 4) | (1ull << 16) | ((uint64_t)(512 >> 4) << 32) | (1ull << 46) | (4ull << 61);
}

// NARROW = false: full 256-column tiles, everything about the tile is a compile-time constant (the runtime-width variant
// costs the wide projections 45%: 12.9 ms against 8.8 ms per cfg3 step)
// PRESPLIT = true: W arrives already split (rows [0, N) = hi, rows [N, 2N) = lo of a [2N, K] array made once when the head
// is packed): the producer fetches both halves and the splitter warps only touch A -- a third of their shared-memory traffic,
// which is what bounds this kernel (per slab: 24 KB landed + 72 KB split traffic + 72 KB operand fetch against 768 clk of MMA)
template <bool NARROW, bool PRESPLIT>
__global__ void __launch_bounds__(TC_THREADS, 1)
gemm_tf32x3_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_w,
                   const float* __restrict__ bias, float* __restrict__ C, int M, int N, int K) {
    extern __shared__ unsigned char smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t bars = base + TC_STAGES * STAGE_BYTES;
    // barrier map (8 B each): [0,S) tma_full, [S,2S) split_done, [2S,3S) mma_done, [3S] accum_full, then tmem slot
    auto bar_full = [&](int s) { return bars + 8u * s; };
    auto bar_split = [&](int s) { return bars + 8u * (TC_STAGES + s); };
    auto bar_free = [&](int s) { return bars + 8u * (2 * TC_STAGES + s); };
    const uint32_t bar_accum = bars + 8u * (3 * TC_STAGES);
    const uint32_t tmem_slot = bars + 8u * (3 * TC_STAGES + 1);
    unsigned char* gen = smem_raw + (base - smem_u32(smem_raw));   // generic pointer to the aligned region

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int n0 = blockIdx.x * TC_BN, m0 = blockIdx.y * TC_BM;
    const int KB = K / TC_BK;
    // narrow outputs (linear2: N = 72 / 96, one column tile): the W box has n_mma = N rounded up to 16 rows (TMA zero-fills
    // the few rows past the end, the transaction counts the whole box), the MMAs span n_mma columns, the epilogue masks its
    // stores.  (A full 256-row box over a 72-row tensor works too but is 6x slower: 1.8 ms against 0.3 ms for the FFMA kernel.)
    const int n_valid = NARROW ? min(TC_BN, N - n0) : TC_BN;
    const int n_mma = NARROW ? (n_valid + 15) & ~15 : TC_BN;

    if (tid == 0) {
        for (int s = 0; s < TC_STAGES; ++s) {
            mbar_init(bar_full(s), 1);
            mbar_init(bar_split(s), 128);
            mbar_init(bar_free(s), 1);
        }
        mbar_init(bar_accum, 1);
        mbar_fence_init_cluster();
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem = *reinterpret_cast<volatile uint32_t*>(gen + (tmem_slot - base));

    if (warp == 0) {
        if (lane == 0) {
            for (int kb = 0; kb < KB; ++kb) {
                const int s = kb % TC_STAGES;
                if (kb >= TC_STAGES) mbar_wait(bar_free(s), ((kb / TC_STAGES) - 1) & 1);
                mbar_arrive_expect_tx(bar_full(s), A_TILE + (PRESPLIT ? 2u : 1u) * (uint32_t)n_mma * (TC_BK * 4));
                const uint32_t st = base + s * STAGE_BYTES;
                tma_load_2d(st, &map_a, kb * TC_BK, m0, bar_full(s));
                tma_load_2d(st + 2 * A_TILE, &map_w, kb * TC_BK, n0, bar_full(s));
                if (PRESPLIT) tma_load_2d(st + 2 * A_TILE + W_TILE, &map_w, kb * TC_BK, N + n0, bar_full(s));
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            // instruction descriptor: D = f32, A = B = tf32, both K-major, N = n_mma (256 for full tiles), M = 128
            const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(n_mma >> 3) << 17) | ((uint32_t)(TC_BM >> 4) << 24);
            for (int kb = 0; kb < KB; ++kb) {
                const int s = kb % TC_STAGES;
                mbar_wait(bar_split(s), (kb / TC_STAGES) & 1);
                tcgen05_fence_after();
                const uint32_t a_hi = base + s * STAGE_BYTES, a_lo = a_hi + A_TILE;
                const uint32_t w_hi = a_hi + 2 * A_TILE, w_lo = w_hi + W_TILE;
                // The tensor core truncates its fp32 accumulator on every add; keeping the 2^-11-sized correction
                // products in their own accumulator means those truncations are 2^-11 smaller too, and the main
                // chain sees K/8 adds instead of 3K/8 (measured: 4e-6 -> see test_tensor_core_gemm_matches_fp32).
#pragma unroll
                for (int p = 0; p < 3; ++p) {
                    const uint32_t a = (p == 0) ? a_lo : a_hi;
                    const uint32_t w = (p == 1) ? w_lo : w_hi;
                    const uint32_t d = (p == 2) ? tmem : tmem + TC_BN;
#pragma unroll
                    for (int k2 = 0; k2 < TC_BK / 8; ++k2)
                        umma_tf32(d, umma_desc_sw64(a + k2 * 32), umma_desc_sw64(w + k2 * 32), idesc,
                                  (kb | (p == 1 ? 1 : 0) | k2) != 0 ? 1u : 0u);
                }
                tcgen05_commit(bar_free(s));             // arrives when the MMAs above have finished reading the stage
            }
            tcgen05_commit(bar_accum);
        }
    } else {
        // ---- splitter ----------------------------------------------------------------------------------------
        const int t = tid - 64;
        for (int kb = 0; kb < KB; ++kb) {
            const int s = kb % TC_STAGES;
            mbar_wait(bar_full(s), (kb / TC_STAGES) & 1);
            uint4* a_hi = reinterpret_cast<uint4*>(gen + s * STAGE_BYTES);
            uint4* a_lo = a_hi + A_TILE / 16;
            uint4* w_hi = a_hi + 2 * A_TILE / 16;
            uint4* w_lo = w_hi + W_TILE / 16;
            constexpr bool WRITE_HI = false;
            auto split = [](uint4* hi_p, uint4* lo_p, int i) {
                const uint4 v = hi_p[i];
                uint4 h, l;
                h.x = v.x & 0xFFFFE000u; h.y = v.y & 0xFFFFE000u; h.z = v.z & 0xFFFFE000u; h.w = v.w & 0xFFFFE000u;
                // lo = x - hi is exact; round it to TF32 ourselves (to nearest) so the tensor core's truncation of the
                // operand cannot add a one-sided 2^-22 error per product (measured 3e-6 absolute at K = 512 without this)
                l.x = tf32_rna(__uint_as_float(v.x) - __uint_as_float(h.x));
                l.y = tf32_rna(__uint_as_float(v.y) - __uint_as_float(h.y));
                l.z = tf32_rna(__uint_as_float(v.z) - __uint_as_float(h.z));
                l.w = tf32_rna(__uint_as_float(v.w) - __uint_as_float(h.w));
                // `hi` is not written back: kind::tf32 reads fp32 words and drops the low 13 mantissa bits itself, which is exactly
                // hi = x & 0xFFFFE000 -- the raw slab IS the hi operand (test_tensor_core_gemm_matches_fp32 would see 1e-4 instead
                // of 1e-6 if the tensor core rounded instead of truncating)
                if (WRITE_HI) hi_p[i] = h;
                lo_p[i] = l;
            };
#pragma unroll
            for (int i = 0; i < (int)(A_TILE / 16 / 128); ++i) split(a_hi, a_lo, t + i * 128);
#pragma unroll
            for (int i = 0; i < (int)(PRESPLIT ? 0 : W_TILE / 16 / 128); ++i)
                if (!NARROW || (t + i * 128) * 16 < n_mma * (TC_BK * 4)) split(w_hi, w_lo, t + i * 128);      // 64-byte rows: only the rows the MMAs read
            fence_proxy_async_smem();
            mbar_arrive(bar_split(s));
        }
        // ---- epilogue: TMEM -> registers -> +bias -> global ---------------------------------------------------
        mbar_wait(bar_accum, 0);
        tcgen05_fence_after();
        const int wq = warp & 3;                         // TMEM lane quarter this warp may read
        const int row = m0 + wq * 32 + lane;
        for (int c0 = 0; c0 < n_valid; c0 += 32) {
            uint32_t v[32];
            const uint32_t taddr = tmem + ((uint32_t)(wq * 32) << 16) + (uint32_t)c0;
            asm volatile(
                "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
                "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                  "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
                  "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
                  "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
                : "r"(taddr)
                : "memory");
            uint32_t u[32];
            asm volatile(
                "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
                "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                : "=r"(u[0]), "=r"(u[1]), "=r"(u[2]), "=r"(u[3]), "=r"(u[4]), "=r"(u[5]), "=r"(u[6]), "=r"(u[7]),
                  "=r"(u[8]), "=r"(u[9]), "=r"(u[10]), "=r"(u[11]), "=r"(u[12]), "=r"(u[13]), "=r"(u[14]), "=r"(u[15]),
                  "=r"(u[16]), "=r"(u[17]), "=r"(u[18]), "=r"(u[19]), "=r"(u[20]), "=r"(u[21]), "=r"(u[22]), "=r"(u[23]),
                  "=r"(u[24]), "=r"(u[25]), "=r"(u[26]), "=r"(u[27]), "=r"(u[28]), "=r"(u[29]), "=r"(u[30]), "=r"(u[31])
                : "r"(taddr + (uint32_t)TC_BN)
                : "memory");
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = __float_as_uint(__uint_as_float(v[j]) + __uint_as_float(u[j]));
            if (row < M) {
                float* dst = C + (size_t)row * N + n0 + c0;
                const float4* b4 = reinterpret_cast<const float4*>(bias + n0 + c0);
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    if (NARROW && c0 + 4 * j >= n_valid) break;          // N is a multiple of 4: whole float4s are valid or not
                    const float4 b = __ldg(b4 + j);
                    float4 o;
                    o.x = __uint_as_float(v[4 * j + 0]) + b.x;
                    o.y = __uint_as_float(v[4 * j + 1]) + b.y;
                    o.z = __uint_as_float(v[4 * j + 2]) + b.z;
                    o.w = __uint_as_float(v[4 * j + 3]) + b.w;
                    reinterpret_cast<float4*>(dst)[j] = o;
                }
            }
        }
    }
    tcgen05_fence_before();
    __syncthreads();
    if (warp == 1) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(TMEM_COLS) : "memory");
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

// [rows, K] fp32 row-major -> 2-D tensor map with a (16 x box_rows) box, 64-byte swizzle
int make_map(CUtensorMap* map, const float* ptr, int rows, int K, int box_rows) {
    EncodeTiledFn fn = encode_fn();
    if (!fn) {
        set_error("gemm_tc: cuTensorMapEncodeTiled is not available from this driver");
        return MP_ERR_CUDA;
    }
    const cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)rows};
    const cuuint64_t strides[1] = {(cuuint64_t)K * 4};
    const cuuint32_t box[2] = {(cuuint32_t)TC_BK, (cuuint32_t)box_rows};
    const cuuint32_t estr[2] = {1, 1};
    const CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(ptr), dims, strides, box, estr,
                          CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("gemm_tc: cuTensorMapEncodeTiled failed with CUresult %d (rows=%d K=%d)", (int)r, rows, K);
        return MP_ERR_CUDA;
    }
    return MP_OK;
}

}  // namespace

bool gemm_tc_eligible(int M, int N, int K) {
    const char* v = getenv("MP_GEMM");
    if (v && strcmp(v, "ffma") == 0) return false;
    const int min_m = (v && strcmp(v, "tc") == 0) ? 1 : 2048;
    // Narrow outputs (linear2, N = 72 / 96) run 2.5x faster here in isolation (0.12 ms against 0.31 ms for the FFMA kernel at
    // M = 76800, K = 512) but make the whole step 1.6 ms SLOWER: a CTA of this kernel owns all 512 TMEM columns and 195 KB of
    // shared memory, so it cannot share an SM with the recurrence clusters of the other heads the way the FFMA kernel does
    // (and it fragments the GPCs those clusters need).  They stay on the FFMA kernel unless MP_GEMM=tc asks otherwise.
    const bool narrow_ok = v && strcmp(v, "tc") == 0 && N < TC_BN && N % 4 == 0 && N >= 16;
    return M >= min_m && (N % TC_BN == 0 || narrow_ok) && K % TC_BK == 0 && K >= TC_BK;
}

static int launch_gemm_tc_impl(const float* A, const float* W, const float* bias, float* C, int M, int N, int K, bool presplit, cudaStream_t stream) {
    MP_REQUIRE(A && W && bias && C && M > 0, "gemm_tc: bad arguments");
    MP_REQUIRE((N % TC_BN == 0 || (N < TC_BN && N % 4 == 0 && N >= 16)) && K % TC_BK == 0,
               "gemm_tc: N=%d must be a multiple of %d (or of 4, in [16, %d)) and K=%d of %d", N, TC_BN, TC_BN, K, TC_BK);
    MP_REQUIRE(((uintptr_t)A & 15) == 0 && ((uintptr_t)W & 15) == 0 && ((uintptr_t)C & 15) == 0 && ((uintptr_t)bias & 15) == 0,
               "gemm_tc: pointers must be 16-byte aligned");
    alignas(64) CUtensorMap map_a, map_w;
    MP_TRY(make_map(&map_a, A, M, K, TC_BM));
    MP_REQUIRE(!presplit || N % TC_BN == 0, "gemm_tc: pre-split weights need N=%d to be a multiple of %d", N, TC_BN);
    MP_TRY(make_map(&map_w, W, presplit ? 2 * N : N, K, N % TC_BN == 0 ? TC_BN : ((N + 15) & ~15)));
    static bool configured = false;
    if (!configured) {
        MP_CUDA_TRY(cudaFuncSetAttribute(gemm_tf32x3_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TC_SMEM));
        MP_CUDA_TRY(cudaFuncSetAttribute(gemm_tf32x3_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TC_SMEM));
        MP_CUDA_TRY(cudaFuncSetAttribute(gemm_tf32x3_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TC_SMEM));
        configured = true;
    }
    ProfileScope prof("gemm_tf32x3", 4.0 * ((double)N * K + N + (double)M * K + (double)M * N), stream);
    dim3 grid((N + TC_BN - 1) / TC_BN, (M + TC_BM - 1) / TC_BM);
    if (presplit)
        gemm_tf32x3_kernel<false, true><<<grid, TC_THREADS, TC_SMEM, stream>>>(map_a, map_w, bias, C, M, N, K);
    else if (N % TC_BN == 0)
        gemm_tf32x3_kernel<false, false><<<grid, TC_THREADS, TC_SMEM, stream>>>(map_a, map_w, bias, C, M, N, K);
    else
        gemm_tf32x3_kernel<true, false><<<grid, TC_THREADS, TC_SMEM, stream>>>(map_a, map_w, bias, C, M, N, K);
    MP_CUDA_TRY(cudaGetLastError());
    count_launch();
    return MP_OK;
}

int launch_gemm_tf32x3(const float* A, const float* W, const float* bias, float* C, int M, int N, int K, cudaStream_t stream) {
    return launch_gemm_tc_impl(A, W, bias, C, M, N, K, false, stream);
}

// W_split [2N, K]: rows [0, N) = TF32 hi parts, rows [N, 2N) = TF32-rounded lo parts of W (launch_split_weights)
int launch_gemm_tf32x3_presplit(const float* A, const float* W_split, const float* bias, float* C, int M, int N, int K, cudaStream_t stream) {
    return launch_gemm_tc_impl(A, W_split, bias, C, M, N, K, true, stream);
}

namespace {
__global__ void split_weights_kernel(const float* __restrict__ w, size_t n, float* __restrict__ out) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const float x = w[i];
        const uint32_t hi = __float_as_uint(x) & 0xFFFFE000u;
        out[i] = __uint_as_float(hi);
        out[n + i] = __uint_as_float(tf32_rna(x - __uint_as_float(hi)));      // same rounding as the in-kernel splitter
    }
}
}  // namespace

int launch_split_weights(const float* W, size_t n, float* W_split, cudaStream_t stream) {
    split_weights_kernel<<<296, 256, 0, stream>>>(W, n, W_split);
    MP_CUDA_TRY(cudaGetLastError());
    return MP_OK;
}

}  // namespace mp
