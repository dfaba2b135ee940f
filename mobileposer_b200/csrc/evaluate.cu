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

// ---- the [10, 2] mean / std rows of FullMotionEvaluator.__call__ (articulate/evaluator.py:326-343) from the per-frame errors ------
// One CTA per sequence.  Every row is `x.mean()` and `x.std(dim=0).mean()` of a [frames, columns] array: per column the sum and the
// sum of squares over the frames are all that is needed, accumulated in double (lane = joint column, warps stride over frames).
//   0 joint position error [n,24]   2 local angle error   3 global angle error   4 / 5 jitter of the predicted / true joints
//   ((p[t+3] - 3 p[t+2] + 3 p[t+1] - p[t]) f^3, norm)   6 root translation error over one second, cm   7-9 rows 0, 2, 3 on the masked joints
// Row 1 (mesh) is left NaN: mp_eval_vertex_errors fills it when a template is given.  Replaces ~40 small torch launches per sequence.
constexpr int ER_THREADS = 256, ER_WARPS = ER_THREADS / 32, ER_ACC = 12;

__global__ void __launch_bounds__(ER_THREADS)
eval_motion_rows_kernel(const float* __restrict__ jp, const float* __restrict__ jt, const float* __restrict__ je, const float* __restrict__ lae,
                        const float* __restrict__ gae, int n, int fps, unsigned mask_bits, float* __restrict__ rows,
                        const long long* __restrict__ offsets) {
    if (offsets) {
        // batched form: CTA b reduces the frames [offsets[b], offsets[b + 1]) of the concatenated arrays into rows[b] -- the same
        // accumulation order per sequence as a launch of its own, so the rows are bit-identical; the sequences run side by side
        // instead of one latency-bound CTA after the other (50 x 0.9 ms per cfg4 pass)
        const long long o = offsets[blockIdx.x];
        n = (int)(offsets[blockIdx.x + 1] - o);
        jp += o * 72; jt += o * 72; je += o * 24; lae += o * 24; gae += o * 24;
        rows += (size_t)blockIdx.x * 20;
    }
    __shared__ double red[ER_WARPS][32][ER_ACC];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    double acc[ER_ACC];
#pragma unroll
    for (int i = 0; i < ER_ACC; ++i) acc[i] = 0.0;
    const float f3 = (float)fps * (float)fps * (float)fps;
    for (int t = warp; t < n; t += ER_WARPS) {
        if (lane < 24) {
            const float a = je[(size_t)t * 24 + lane], b = lae[(size_t)t * 24 + lane], c = gae[(size_t)t * 24 + lane];
            acc[0] += a; acc[1] += (double)a * a;
            acc[2] += b; acc[3] += (double)b * b;
            acc[4] += c; acc[5] += (double)c * c;
            if (t + 3 < n) {
#pragma unroll
                for (int w = 0; w < 2; ++w) {
                    const float* q = (w ? jt : jp) + ((size_t)t * 24 + lane) * 3;
                    float v[3];
#pragma unroll
                    for (int k = 0; k < 3; ++k) {
                        // ((p3 - 3 p2) + 3 p1) - p0, then * f^3: the order of the torch expression
                        const float d = __fsub_rn(__fadd_rn(__fsub_rn(q[216 + k], __fmul_rn(3.f, q[144 + k])), __fmul_rn(3.f, q[72 + k])), q[k]);
                        v[k] = __fmul_rn(d, f3);
                    }
                    const float jn = norm3_rn(v[0], v[1], v[2]);
                    acc[6 + 2 * w] += jn; acc[7 + 2 * w] += (double)jn * jn;
                }
            }
        } else if (lane == 24 && t + fps < n) {
            float v[3];
#pragma unroll
            for (int k = 0; k < 3; ++k)
                v[k] = __fsub_rn(__fsub_rn(jp[(size_t)(t + fps) * 72 + k], jp[(size_t)t * 72 + k]),
                                 __fsub_rn(jt[(size_t)(t + fps) * 72 + k], jt[(size_t)t * 72 + k]));
            const float e = __fmul_rn(norm3_rn(v[0], v[1], v[2]), 100.f);
            acc[10] += e; acc[11] += (double)e * e;
        }
    }
#pragma unroll
    for (int i = 0; i < ER_ACC; ++i) red[warp][lane][i] = acc[i];
    __syncthreads();
    if (warp == 0) {
#pragma unroll
        for (int i = 0; i < ER_ACC; ++i) {
            double v = 0.0;
            for (int w = 0; w < ER_WARPS; ++w) v += red[w][lane][i];
            red[0][lane][i] = v;
        }
    }
    __syncthreads();
    if (tid < 10) {
        // row -> (accumulator pair, frames, column set)
        const int r = tid;
        const float nan = __int_as_float(0x7fc00000);
        float mean = nan, sd = nan;
        if (r != 1) {
            const int a = (r == 0 || r == 7) ? 0 : (r == 2 || r == 8) ? 2 : (r == 3 || r == 9) ? 4 : (r == 4) ? 6 : (r == 5) ? 8 : 10;
            const long long cnt = (r == 4 || r == 5) ? (long long)n - 3 : (r == 6) ? (long long)n - fps : (long long)n;
            const unsigned cols = (r == 6) ? (1u << 24) : (r >= 7) ? mask_bits : 0x00FFFFFFu;
            if (cnt > 0 && cols) {
                double sum = 0.0, sds = 0.0;
                int nc = 0;
                for (int c = 0; c < 25; ++c)
                    if ((cols >> c) & 1u) {
                        const double s1 = red[0][c][a], s2 = red[0][c][a + 1];
                        sum += s1;
                        if (cnt > 1) sds += sqrt(fmax((s2 - s1 * s1 / (double)cnt) / (double)(cnt - 1), 0.0));
                        ++nc;
                    }
                mean = (float)(sum / ((double)cnt * nc));
                sd = (float)(sds / nc);
            }
        }
        rows[2 * r] = mean;
        rows[2 * r + 1] = sd;
    }
}

}  // namespace

int launch_eval_motion_rows(const float* jp, const float* jt, const float* je, const float* lae, const float* gae, int64_t n, int fps,
                            unsigned mask_bits, float* rows, cudaStream_t stream) {
    MP_REQUIRE(jp && jt && je && lae && gae && rows && n > 0 && n < (int64_t)1 << 30 && fps > 0, "eval_motion_rows: bad arguments");
    MP_REQUIRE((mask_bits & ~0x00FFFFFFu) == 0, "eval_motion_rows: the joint mask has 24 bits");
    eval_motion_rows_kernel<<<1, ER_THREADS, 0, stream>>>(jp, jt, je, lae, gae, (int)n, fps, mask_bits, rows, nullptr);
    MP_CUDA_TRY(cudaGetLastError());
    count_launch();
    return MP_OK;
}

int launch_eval_motion_rows_batch(const float* jp, const float* jt, const float* je, const float* lae, const float* gae, const long long* offsets,
                                  int n_sequences, int fps, unsigned mask_bits, float* rows, cudaStream_t stream) {
    MP_REQUIRE(jp && jt && je && lae && gae && offsets && rows && n_sequences > 0 && fps > 0, "eval_motion_rows_batch: bad arguments");
    MP_REQUIRE((mask_bits & ~0x00FFFFFFu) == 0, "eval_motion_rows_batch: the joint mask has 24 bits");
    eval_motion_rows_kernel<<<n_sequences, ER_THREADS, 0, stream>>>(jp, jt, je, lae, gae, 0, fps, mask_bits, rows, offsets);
    MP_CUDA_TRY(cudaGetLastError());
    count_launch();
    return MP_OK;
}

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
