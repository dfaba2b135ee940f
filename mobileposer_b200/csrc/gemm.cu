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

namespace mp {

namespace {

constexpr int BK = 16;
constexpr int GEMM_THREADS = 256;

template <int BM, int BN, int TM, int TN>
__global__ void __launch_bounds__(GEMM_THREADS, 2)
gemm_bias_act_kernel(const float* __restrict__ A1, int K1, const float* __restrict__ A2, int K2,
                     const float* __restrict__ W, const float* __restrict__ bias, float* __restrict__ C,
                     int M, int N, int relu) {
    static_assert((BM / TM) * (BN / TN) == GEMM_THREADS, "thread tiling");
    constexpr int RM = TM / 4, RN = TN / 4;      // float4 groups per thread along M / N
    constexpr int LDA = BM + 4, LDB = BN + 4;    // +4 keeps rows 16B aligned, 2-way store conflicts only
    constexpr int A_F4 = BM * BK / 4 / GEMM_THREADS;
    constexpr int B_F4 = BN * BK / 4 / GEMM_THREADS;
    static_assert(A_F4 >= 1 && B_F4 >= 1, "tile too small");

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
            if (m < M && k < K) {
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
            if (n < N && k < K) v = __ldg(reinterpret_cast<const float4*>(W + (size_t)n * K + k));
            rb[i] = v;
        }
    };
    auto store_ab = [&](int buf) {
#pragma unroll
        for (int i = 0; i < A_F4; ++i) {
            const int f = tid + i * GEMM_THREADS;
            const int row = f / (BK / 4), kk = (f % (BK / 4)) * 4;
            As[buf][kk + 0][row] = ra[i].x;
            As[buf][kk + 1][row] = ra[i].y;
            As[buf][kk + 2][row] = ra[i].z;
            As[buf][kk + 3][row] = ra[i].w;
        }
#pragma unroll
        for (int i = 0; i < B_F4; ++i) {
            const int f = tid + i * GEMM_THREADS;
            const int row = f / (BK / 4), kk = (f % (BK / 4)) * 4;
            Bs[buf][kk + 0][row] = rb[i].x;
            Bs[buf][kk + 1][row] = rb[i].y;
            Bs[buf][kk + 2][row] = rb[i].z;
            Bs[buf][kk + 3][row] = rb[i].w;
        }
    };

    float acc[TM][TN];
#pragma unroll
    for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

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
#pragma unroll
            for (int i = 0; i < TM; ++i)
#pragma unroll
                for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
        }
        if (s + 1 < nslab) store_ab(buf ^ 1);
        __syncthreads();
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
                    v[j] = relu ? fmaxf(x, 0.f) : x;
                }
                float* dst = C + (size_t)m * N + n;
                if (vec) {
                    *reinterpret_cast<float4*>(dst) = make_float4(v[0], v[1], v[2], v[3]);
                } else {
#pragma unroll
                    for (int j = 0; j < 4; ++j)
                        if (n + j < N) dst[j] = v[j];
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
    MP_REQUIRE(((uintptr_t)A1 & 15) == 0 && ((uintptr_t)A2 & 15) == 0 && ((uintptr_t)W & 15) == 0 &&
                   ((uintptr_t)C & 15) == 0, "gemm: pointers must be 16-byte aligned");
    ProfileScope prof(N >= 512 ? "gemm_input_proj" : "gemm_linear",
                      4.0 * ((double)N * (K1 + K2) + N + (double)M * (K1 + K2) + (double)M * N), stream);
    const long big_ctas = (long)((M + 127) / 128) * ((N + 127) / 128);
    if (big_ctas >= 148) {
        dim3 grid((N + 127) / 128, (M + 127) / 128);
        gemm_bias_act_kernel<128, 128, 8, 8><<<grid, GEMM_THREADS, 0, stream>>>(A1, K1, A2, K2, W, bias, C, M, N, relu);
    } else {
        dim3 grid((N + 63) / 64, (M + 63) / 64);
        gemm_bias_act_kernel<64, 64, 4, 4><<<grid, GEMM_THREADS, 0, stream>>>(A1, K1, A2, K2, W, bias, C, M, N, relu);
    }
    MP_CUDA_TRY(cudaGetLastError());
    count_launch();
    return MP_OK;
}

}  // namespace mp
