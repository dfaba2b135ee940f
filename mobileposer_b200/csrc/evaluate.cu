// Translation-error windows of the reference evaluator on the device.
//
// Replaces the Python loops of mobileposer/evaluate.py:66-92 (`evaluate_pose(..., evaluate_tran=True)`): for every sequence,
// the distance the ground-truth root has travelled up to each frame (evaluate.py:68-71), and for each window of w = 1..7 m a
// two-pointer sweep over (start, end) frame pairs across which the ground truth moves at least w metres (evaluate.py:73-83),
// the relative drift `|dt - dp| / moved * w` of each pair and its mean (evaluate.py:85-92).  The reference runs ~2T Python
// iterations with tensor scalars per window and sequence; here one CTA serves a sequence and one warp a window.
//
// Exactness: which pairs exist depends on fp32 comparisons of the running distance, so the distance is accumulated the way
// the reference does it -- sequentially, in fp32, one frame after the other (a parallel scan would round differently and
// move pair boundaries).  That chain is T dependent adds by one thread (T = 3000: ~50 us); everything around it is parallel:
// per-frame step lengths by the whole CTA, the seven windows by seven warps, the sequences by the grid.
#include "mp_common.cuh"

namespace mp {

namespace {

constexpr int TW_THREADS = 256;
constexpr int TW_WINDOWS = 7;

__device__ __forceinline__ float norm3_rn(float x, float y, float z) {
    // ((x^2 + y^2) + z^2) without FMA contraction: the order of torch's 3-element reduction
    return sqrtf(__fadd_rn(__fadd_rn(__fmul_rn(x, x), __fmul_rn(y, y)), __fmul_rn(z, z)));
}

__global__ void __launch_bounds__(TW_THREADS)
eval_tran_windows_kernel(const float* __restrict__ tran_p, const float* __restrict__ tran_t, const int32_t* __restrict__ lengths,
                         int T, float* __restrict__ err, int32_t* __restrict__ count) {
    extern __shared__ float move[];                        // [T] distance travelled by the ground truth up to frame j
    const int s = blockIdx.x, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int n = lengths ? min(max(lengths[s], 0), T) : T;
    const float* tp = tran_p + (size_t)s * T * 3;
    const float* tt = tran_t + (size_t)s * T * 3;

    // step lengths v[j] = |tran_t[j+1] - tran_t[j]| -> move[j+1]
    for (int j = tid; j + 1 < n; j += TW_THREADS) {
        const float dx = __fsub_rn(tt[3 * (j + 1) + 0], tt[3 * j + 0]);
        const float dy = __fsub_rn(tt[3 * (j + 1) + 1], tt[3 * j + 1]);
        const float dz = __fsub_rn(tt[3 * (j + 1) + 2], tt[3 * j + 2]);
        move[j + 1] = norm3_rn(dx, dy, dz);
    }
    __syncthreads();
    if (tid == 0 && n > 0) {                               // the reference's sequential fp32 accumulation
        float acc = 0.f;
        move[0] = 0.f;
        for (int j = 1; j < n; ++j) {
            acc = __fadd_rn(acc, move[j]);
            move[j] = acc;
        }
    }
    __syncthreads();

    if (warp < TW_WINDOWS && lane == 0) {
        const float w = (float)(warp + 1);
        int start = 0, end = 1, last_end = -1, pairs = 0;
        float total = 0.f;
        while (end < n) {
            const float moved = __fsub_rn(move[end], move[start]);
            if (moved < w) {
                ++end;
            } else {
                if (last_end != end) {                     // only the first start that reaches this end counts
                    const float dx = __fsub_rn(__fsub_rn(tt[3 * end + 0], tt[3 * start + 0]), __fsub_rn(tp[3 * end + 0], tp[3 * start + 0]));
                    const float dy = __fsub_rn(__fsub_rn(tt[3 * end + 1], tt[3 * start + 1]), __fsub_rn(tp[3 * end + 1], tp[3 * start + 1]));
                    const float dz = __fsub_rn(__fsub_rn(tt[3 * end + 2], tt[3 * start + 2]), __fsub_rn(tp[3 * end + 2], tp[3 * start + 2]));
                    const float e = __fmul_rn(__fdiv_rn(norm3_rn(dx, dy, dz), moved), w);
                    total = __fadd_rn(total, e);
                    last_end = end;
                    ++pairs;
                }
                ++start;
            }
        }
        err[s * TW_WINDOWS + warp] = pairs > 0 ? __fdiv_rn(total, (float)pairs) : __int_as_float(0x7fc00000);
        count[s * TW_WINDOWS + warp] = pairs;
    }
}

}  // namespace

int launch_eval_tran_windows(const float* tran_p, const float* tran_t, const int32_t* lengths, int S, int T, float* err,
                             int32_t* count, cudaStream_t stream) {
    MP_REQUIRE(tran_p && tran_t && err && count && S > 0 && T > 0, "eval_tran_windows: bad arguments");
    const size_t smem = (size_t)T * sizeof(float);
    MP_REQUIRE(smem <= 200 * 1024, "eval_tran_windows: T=%d frames exceed the %d the running distance fits in shared memory for", T,
               200 * 1024 / 4);
    if (smem > 48 * 1024)
        MP_CUDA_TRY(cudaFuncSetAttribute(eval_tran_windows_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    eval_tran_windows_kernel<<<S, TW_THREADS, smem, stream>>>(tran_p, tran_t, lengths, T, err, count);
    MP_CUDA_TRY(cudaGetLastError());
    count_launch();
    return MP_OK;
}

}  // namespace mp
