// fp32 GEMM with fused concat / bias / ReLU:  C = act([A1 | A2] * W^T + bias)
//
// Replaces the torch `addmm` call sites of the hot path (mobileposer/models/rnn.py:22,32 and the
// hoisted input projection W_ih x_t + b_ih + b_hh of nn.LSTM, rnn.py:27) plus the torch.cat of
// mobileposer/models/net.py:106,113 (two-source A operand).
//
// fp32 FFMA on purpose: the parity bar is 1e-4 rad / 1e-4 m through up-to-3000-step recurrences
// against an fp32 reference, so operands are not rounded to TF32/BF16 here ("precision").
// Classic register-tiled SGEMM: BMxBNx16 smem tiles stored k-major so the inner loop is LDS.128 +
// FFMA, double-buffered smem with a register prefetch of the next k-slab, one barrier per slab.
#include "mp_common.cuh"

#include <cuda_fp16.h>

#include <algorithm>
#include <cstdlib>

namespace mp {

namespace {

constexpr int BK = 16;

// Blackwell packed fp32 (SASS `FFMA2 Rd, Ra.F32x2, Rb.F32, Rc.F32x2`, the scalar operand is broadcast): one issue slot,
// two independent round-to-nearest FMAs -- bit-identical to two FFMA.  The k-major A tile makes the pairs free: a
// float4 of As holds 4 consecutive rows at one k, so (x, y) and (z, w) are already adjacent registers.
typedef unsigned long long f32x2;
__device__ __forceinline__ f32x2 pack2(float lo, float hi) {
    f32x2 r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ void unpack2(f32x2 v, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ f32x2 ffma2(f32x2 a, float b, f32x2 c) {
    f32x2 d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(pack2(b, b)), "l"(c));
    return d;
}

// Tile shapes (BM x BN per CTA, TM x TN per thread, THREADS = (BM/TM)*(BN/TN)):
//   128 x 128, 8 x 8, 256 threads   wide outputs (linear1: N = 256; small-batch projections)
//   128 x  96, 4 x 12, 256 threads  linear2 of the pose head (N = 96)
//   128 x  72, 4 x 12, 192 threads  linear2 of the joints / velocity heads (N = 72): no padded columns
//   128 x  64, 4 x 8, 256 threads   linear1 of the foot-contact head (N = 64)
//    64 x  64, 4 x 4, 256 threads   grids that would not fill the SMs otherwise
// (The 128-wide tile spent 44 % of its FMAs on padding at N = 72: 0.31 ms against a 0.08 ms FFMA floor.)
template <int BM, int BN, int TM, int TN, int THREADS, int MINB, bool PACKED>
__global__ void __launch_bounds__(THREADS, MINB)
gemm_bias_act_kernel(const float* __restrict__ A1, int K1, const float* __restrict__ A2, int K2,
                     const float* __restrict__ W, const float* __restrict__ bias, float* __restrict__ C,
                     int M, int N, int relu) {
    static_assert((BM / TM) * (BN / TN) == THREADS, "thread tiling");
    static_assert(TM % 4 == 0 && TN % 4 == 0, "float4 groups");
    constexpr int GEMM_THREADS = THREADS;
    constexpr int RM = TM / 4, RN = TN / 4;      // float4 groups per thread along M / N
    constexpr int LDA = BM + 4, LDB = BN + 4;    // +4 keeps rows 16B aligned, 2-way store conflicts only
    constexpr int A_TOT = BM * BK / 4, B_TOT = BN * BK / 4;   // float4 per k-slab
    constexpr int A_F4 = (A_TOT + GEMM_THREADS - 1) / GEMM_THREADS;
    constexpr int B_F4 = (B_TOT + GEMM_THREADS - 1) / GEMM_THREADS;
    constexpr bool A_EVEN = A_TOT % GEMM_THREADS == 0, B_EVEN = B_TOT % GEMM_THREADS == 0;

    __shared__ __align__(16) float As[2][BK][LDA];
    __shared__ __align__(16) float Bs[2][BK][LDB];

    const int K = K1 + K2;
    const int tid = threadIdx.x;
    const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
    const int tx = tid % (BN / TN), ty = tid / (BN / TN);

    float4 ra[A_F4], rb[B_F4];

    auto load_a = [&](int k0) {
#pragma unroll
        for (int i = 0; i < A_F4; ++i) {
            const int f = tid + i * GEMM_THREADS;
            const int row = f / (BK / 4), k = k0 + (f % (BK / 4)) * 4;
            const int m = m0 + row;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if ((A_EVEN || f < A_TOT) && m < M && k < K) {
                const float* src = (k < K1) ? (A1 + (size_t)m * K1 + k) : (A2 + (size_t)m * K2 + (k - K1));
                v = __ldg(reinterpret_cast<const float4*>(src));
            }
            ra[i] = v;
        }
    };
    auto load_b = [&](int k0) {
#pragma unroll
        for (int i = 0; i < B_F4; ++i) {
            const int f = tid + i * GEMM_THREADS;
            const int row = f / (BK / 4), k = k0 + (f % (BK / 4)) * 4;
            const int n = n0 + row;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if ((B_EVEN || f < B_TOT) && n < N && k < K) v = __ldg(reinterpret_cast<const float4*>(W + (size_t)n * K + k));
            rb[i] = v;
        }
    };
    auto store_ab = [&](int buf) {
#pragma unroll
        for (int i = 0; i < A_F4; ++i) {
            const int f = tid + i * GEMM_THREADS;
            if (!A_EVEN && f >= A_TOT) break;
            const int row = f / (BK / 4), kk = (f % (BK / 4)) * 4;
            As[buf][kk + 0][row] = ra[i].x;
            As[buf][kk + 1][row] = ra[i].y;
            As[buf][kk + 2][row] = ra[i].z;
            As[buf][kk + 3][row] = ra[i].w;
        }
#pragma unroll
        for (int i = 0; i < B_F4; ++i) {
            const int f = tid + i * GEMM_THREADS;
            if (!B_EVEN && f >= B_TOT) break;
            const int row = f / (BK / 4), kk = (f % (BK / 4)) * 4;
            Bs[buf][kk + 0][row] = rb[i].x;
            Bs[buf][kk + 1][row] = rb[i].y;
            Bs[buf][kk + 2][row] = rb[i].z;
            Bs[buf][kk + 3][row] = rb[i].w;
        }
    };

    float acc[TM][TN];
    f32x2 acc2[PACKED ? TM / 2 : 1][PACKED ? TN : 1];   // PACKED: rows (2p, 2p+1) of column j
#pragma unroll
    for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;
    if constexpr (PACKED) {
#pragma unroll
        for (int i = 0; i < TM / 2; ++i)
#pragma unroll
            for (int j = 0; j < TN; ++j) acc2[i][j] = 0ull;
    }

    const int nslab = (K + BK - 1) / BK;
    load_a(0);
    load_b(0);
    store_ab(0);
    __syncthreads();

    for (int s = 0; s < nslab; ++s) {
        const int buf = s & 1;
        if (s + 1 < nslab) {
            load_a((s + 1) * BK);
            load_b((s + 1) * BK);
        }
#pragma unroll
        for (int kk = 0; kk < BK; ++kk) {
            float a[TM], b[TN];
#pragma unroll
            for (int g = 0; g < RM; ++g) {
                const float4 v = *reinterpret_cast<const float4*>(&As[buf][kk][g * (BM / RM) + ty * 4]);
                a[g * 4 + 0] = v.x; a[g * 4 + 1] = v.y; a[g * 4 + 2] = v.z; a[g * 4 + 3] = v.w;
            }
#pragma unroll
            for (int g = 0; g < RN; ++g) {
                const float4 v = *reinterpret_cast<const float4*>(&Bs[buf][kk][g * (BN / RN) + tx * 4]);
                b[g * 4 + 0] = v.x; b[g * 4 + 1] = v.y; b[g * 4 + 2] = v.z; b[g * 4 + 3] = v.w;
            }
            if constexpr (PACKED) {
#pragma unroll
                for (int i = 0; i < TM / 2; ++i) {
                    const f32x2 ap = pack2(a[2 * i], a[2 * i + 1]);
#pragma unroll
                    for (int j = 0; j < TN; ++j) acc2[i][j] = ffma2(ap, b[j], acc2[i][j]);
                }
            } else {
#pragma unroll
                for (int i = 0; i < TM; ++i)
#pragma unroll
                    for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
            }
        }
        if (s + 1 < nslab) store_ab(buf ^ 1);
        __syncthreads();
    }
    if constexpr (PACKED) {
#pragma unroll
        for (int i = 0; i < TM / 2; ++i)
#pragma unroll
            for (int j = 0; j < TN; ++j) unpack2(acc2[i][j], acc[2 * i][j], acc[2 * i + 1][j]);
    }

    const bool vec = (N & 3) == 0;
#pragma unroll
    for (int gi = 0; gi < RM; ++gi)
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int m = m0 + gi * (BM / RM) + ty * 4 + i;
            if (m >= M) continue;
#pragma unroll
            for (int gj = 0; gj < RN; ++gj) {
                const int n = n0 + gj * (BN / RN) + tx * 4;
                if (n >= N) continue;
                float v[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    float x = acc[gi * 4 + i][gj * 4 + j] + ((n + j < N) ? __ldg(bias + n + j) : 0.f);
                    v[j] = (relu & 1) ? fmaxf(x, 0.f) : x;
                }
                float* dst = C + (size_t)m * N + n;
                if (relu & 2) {
                    // output for the fp16-split tensor-core projection (gemm_f16.cu): C holds two [M, N] planes of halves, hi = fp16(x)
                    // and lo = fp16((x - hi) * 2^11) -- the same bytes as the fp32 array, no separate split pass (N % 4 == 0)
                    __half* hi = reinterpret_cast<__half*>(C) + (size_t)m * N + n;
                    __half* lo = hi + (size_t)M * N;
                    const __half2 h0 = __floats2half2_rn(v[0], v[1]), h1 = __floats2half2_rn(v[2], v[3]);
                    const float2 f0 = __half22float2(h0), f1 = __half22float2(h1);
                    const __half2 l0 = __floats2half2_rn((v[0] - f0.x) * 2048.0f, (v[1] - f0.y) * 2048.0f);
                    const __half2 l1 = __floats2half2_rn((v[2] - f1.x) * 2048.0f, (v[3] - f1.y) * 2048.0f);
                    uint2 ph, pl;
                    ph.x = *reinterpret_cast<const uint32_t*>(&h0); ph.y = *reinterpret_cast<const uint32_t*>(&h1);
                    pl.x = *reinterpret_cast<const uint32_t*>(&l0); pl.y = *reinterpret_cast<const uint32_t*>(&l1);
                    *reinterpret_cast<uint2*>(hi) = ph;
                    *reinterpret_cast<uint2*>(lo) = pl;
                } else if (vec) {
                    *reinterpret_cast<float4*>(dst) = make_float4(v[0], v[1], v[2], v[3]);
                } else {
#pragma unroll
                    for (int j = 0; j < 4; ++j)
                        if (n + j < N) dst[j] = v[j];
                }
            }
        }
}

// C[m, 0:N] = act(A[m, :] . W[n, :] + bias[n]) for N <= 8 (linear2 of the foot-contact head: N = 2, K = 128).  One warp per
// row, 4 rows in flight per warp; a lane owns the float4 chunks lane, lane + 32, ... of the row, so a row is read with
// fully coalesced 512-byte requests and A crosses HBM exactly once (39 MB at cfg3: an HBM-bound kernel, where the
// 128 x 128 tile spent 79 us computing 126 padded columns).  W (<= 16 KB) comes from L1.
constexpr int ROWDOT_NMAX = 8;
__global__ void __launch_bounds__(256)
gemm_rowdot_kernel(const float* __restrict__ A, int K, const float* __restrict__ W, const float* __restrict__ bias,
                   float* __restrict__ C, int M, int N, int relu) {
    constexpr int RPW = 4;
    const int lane = threadIdx.x & 31;
    const int warp = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int nwarps = gridDim.x * (blockDim.x >> 5);
    const int kq = K >> 2;
    for (int m0 = warp * RPW; m0 < M; m0 += nwarps * RPW) {
        float acc[RPW][ROWDOT_NMAX];
#pragma unroll
        for (int r = 0; r < RPW; ++r)
#pragma unroll
            for (int n = 0; n < ROWDOT_NMAX; ++n) acc[r][n] = 0.f;
        for (int q = lane; q < kq; q += 32) {
            float4 a[RPW];
#pragma unroll
            for (int r = 0; r < RPW; ++r)
                a[r] = (m0 + r < M) ? __ldg(reinterpret_cast<const float4*>(A + (size_t)(m0 + r) * K) + q) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
            for (int n = 0; n < ROWDOT_NMAX; ++n) {
                if (n < N) {
                    const float4 w = __ldg(reinterpret_cast<const float4*>(W + (size_t)n * K) + q);
#pragma unroll
                    for (int r = 0; r < RPW; ++r)
                        acc[r][n] = fmaf(a[r].w, w.w, fmaf(a[r].z, w.z, fmaf(a[r].y, w.y, fmaf(a[r].x, w.x, acc[r][n]))));
                }
            }
        }
#pragma unroll
        for (int n = 0; n < ROWDOT_NMAX; ++n) {
            if (n < N) {
#pragma unroll
                for (int r = 0; r < RPW; ++r) {
                    float v = acc[r][n];
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
                    acc[r][n] = v;
                }
            }
        }
        if (lane < RPW && m0 + lane < M) {
#pragma unroll
            for (int n = 0; n < ROWDOT_NMAX; ++n) {
                if (n < N) {
                    float v = 0.f;
#pragma unroll
                    for (int r = 0; r < RPW; ++r) v = (lane == r) ? acc[r][n] : v;
                    v += __ldg(bias + n);
                    C[(size_t)(m0 + lane) * N + n] = (relu & 1) ? fmaxf(v, 0.f) : v;
                }
            }
        }
    }
}

}  // namespace

int launch_gemm_bias_act(const float* A1, int K1, const float* A2, int K2, const float* W,
                         const float* bias, float* C, int M, int N, int relu, cudaStream_t stream) {
    // large single-source projections without activation go to the tensor cores (gemm_tc.cu)
    if (!relu && K2 == 0 && gemm_tc_eligible(M, N, K1)) return launch_gemm_tf32x3(A1, W, bias, C, M, N, K1, stream);
    return launch_gemm_ffma(A1, K1, A2, K2, W, bias, C, M, N, relu, stream);
}

int launch_gemm_ffma(const float* A1, int K1, const float* A2, int K2, const float* W,
                     const float* bias, float* C, int M, int N, int relu, cudaStream_t stream) {
    if (M <= 0 || N <= 0) return MP_OK;
    MP_REQUIRE(A1 && W && bias && C, "gemm: null pointer");
    MP_REQUIRE(K1 > 0 && (K1 & 3) == 0 && K2 >= 0 && (K2 & 3) == 0, "gemm: K1=%d K2=%d must be multiples of 4", K1, K2);
    MP_REQUIRE(K2 == 0 || A2, "gemm: second operand missing");
    MP_REQUIRE(!(relu & 2) || ((N & 3) == 0 && N > ROWDOT_NMAX), "gemm: the fp16-split output needs N %% 4 == 0 (N=%d)", N);
    MP_REQUIRE(((uintptr_t)A1 & 15) == 0 && ((uintptr_t)A2 & 15) == 0 && ((uintptr_t)W & 15) == 0 &&
                   ((uintptr_t)C & 15) == 0, "gemm: pointers must be 16-byte aligned");
    ProfileScope prof(N >= 512 ? "gemm_input_proj" : "gemm_linear",
                      4.0 * ((double)N * (K1 + K2) + N + (double)M * (K1 + K2) + (double)M * N), stream);
    const int mt = (M + 127) / 128;
    if (N <= ROWDOT_NMAX && K2 == 0 && M >= 1024) {
        // a handful of outputs per row (foot-contact logits): stream A once, one warp per row
        const int warps = 8;
        const int grid = std::min((M + warps * 4 - 1) / (warps * 4), 148 * 8);
        gemm_rowdot_kernel<<<grid, warps * 32, 0, stream>>>(A1, K1, W, bias, C, M, N, relu);
    } else if ((long)mt * ((N + 127) / 128) >= 148) {
        static const bool packed = !(getenv("MP_GEMM_FFMA2") && atoi(getenv("MP_GEMM_FFMA2")) == 0);
        static const bool wide_only = getenv("MP_GEMM_WIDE") != nullptr;   // A/B switch: the 128-wide tile for every N
#define MP_GEMM_LAUNCH(BM, BN, TM, TN, TH, MINB)                                                                             \
        do {                                                                                                                 \
            dim3 grid((N + BN - 1) / BN, (M + BM - 1) / BM);                                                                 \
            if (packed) gemm_bias_act_kernel<BM, BN, TM, TN, TH, MINB, true><<<grid, TH, 0, stream>>>(A1, K1, A2, K2, W, bias, C, M, N, relu); \
            else gemm_bias_act_kernel<BM, BN, TM, TN, TH, MINB, false><<<grid, TH, 0, stream>>>(A1, K1, A2, K2, W, bias, C, M, N, relu);      \
        } while (0)
        if (wide_only || N > 96) MP_GEMM_LAUNCH(128, 128, 8, 8, 256, 2);
        else if (N > 72) MP_GEMM_LAUNCH(128, 96, 4, 12, 256, 2);
        else if (N > 64) MP_GEMM_LAUNCH(128, 72, 4, 12, 192, 3);
        else MP_GEMM_LAUNCH(128, 64, 4, 8, 256, 2);
#undef MP_GEMM_LAUNCH
    } else {
        dim3 grid((N + 63) / 64, (M + 63) / 64);
        gemm_bias_act_kernel<64, 64, 4, 4, 256, 2, false><<<grid, 256, 0, stream>>>(A1, K1, A2, K2, W, bias, C, M, N, relu);
    }
    MP_CUDA_TRY(cudaGetLastError());
    count_launch();
    return MP_OK;
}

}  // namespace mp
