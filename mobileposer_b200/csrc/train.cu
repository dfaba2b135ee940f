// Training step of one RNN head -- first slice of SURVEY.md 8f row N4: forward with saved activations and the full backward pass of
//   RNN.forward            mobileposer/models/rnn.py:20-33   (linear1 -> ReLU -> dropout mask -> 2-layer (bi)LSTM over packed sequences -> linear2)
//   Joints.shared_step     mobileposer/models/joints.py:54-75 (MSE + 1e-5 x temporal L1 loss over the padded prediction)
// in plain fp32 on the CUDA cores.  Correctness first: this slice exists so that the gradient path has a pinned, tested definition
// (tests/golden/train_joints_step.npz comes from the live reference's own shared_step + backward); the recurrences are simple
// multi-sequence kernels that stream W_hh from L2 every step, not the cluster / tensor-core kernels of the inference path, and the
// weight-gradient contractions are an atomics-based split-M kernel.  Everything is in torch's layouts (gate rows i, f, g, o), so the
// results compare with autograd tensor by tensor.
#include "mp_common.cuh"
#include "mp_constants.cuh"

#include <algorithm>

namespace mp {

namespace {

constexpr int TR_NB = 8;      // sequences per CTA of the training recurrences (W_hh is re-read from L2 once per step and CTA)

__device__ __forceinline__ float sigm(float x) { return 1.0f / (1.0f + expf(-x)); }

// out[c][r] = in[r][c]      in [R, C] row-major
__global__ void transpose_kernel(const float* __restrict__ in, float* __restrict__ out, int R, int C) {
    __shared__ float tile[32][33];
    const int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
    for (int i = threadIdx.y; i < 32; i += blockDim.y) {
        const int r = r0 + i, c = c0 + threadIdx.x;
        tile[i][threadIdx.x] = (r < R && c < C) ? in[(size_t)r * C + c] : 0.f;
    }
    __syncthreads();
    for (int i = threadIdx.y; i < 32; i += blockDim.y) {
        const int c = c0 + i, r = r0 + threadIdx.x;
        if (r < R && c < C) out[(size_t)c * R + r] = tile[threadIdx.x][i];
    }
}

__global__ void mul_kernel(const float* __restrict__ a, const float* __restrict__ b, float* __restrict__ o, size_t n) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) o[i] = a[i] * b[i];
}

// out[m, n] = sum_k a[m, k] * w[k, n]   for a handful of k (the foot-contact head's two logits): one thread per output element
__global__ void gemm_small_k_kernel(const float* __restrict__ a, const float* __restrict__ w, float* __restrict__ out, size_t M, int K, int N) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < M * N; i += (size_t)gridDim.x * blockDim.x) {
        const size_t m = i / N;
        const int n = (int)(i % N);
        float acc = 0.f;
        for (int k = 0; k < K; ++k) acc = fmaf(a[m * K + k], __ldg(w + (size_t)k * N + n), acc);
        out[i] = acc;
    }
}

// dz = g * mask * (relu_out > 0)
__global__ void relu_mask_bwd_kernel(const float* __restrict__ g, const float* __restrict__ relu_out, const float* __restrict__ mask,
                                     float* __restrict__ dz, size_t n) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
        dz[i] = relu_out[i] > 0.f ? g[i] * (mask ? mask[i] : 1.f) : 0.f;
}

// Forward recurrence of one layer, all directions, TR_NB sequences per CTA, thread = hidden unit.  gin = W_ih x + b_ih (torch column
// order gate*H + unit); b_hh is added here.  Saves the post-activation gates and the cell states for the backward pass.
template <int H>
__global__ void __launch_bounds__(H) lstm_train_fwd_kernel(const float* __restrict__ gin, const float* __restrict__ bhh,
                                                          const float* __restrict__ wT, float* __restrict__ y, float* __restrict__ gates,
                                                          float* __restrict__ cs, const int32_t* __restrict__ lengths, int B, int T, int dirs) {
    __shared__ float h[TR_NB][H];
    const int j = threadIdx.x, dir = blockIdx.y, b0 = blockIdx.x * TR_NB;
    const float* wt = wT + (size_t)dir * H * 4 * H;          // [k][gate*H + unit]
    int len[TR_NB], maxlen = 0;
    float c[TR_NB];
#pragma unroll
    for (int n = 0; n < TR_NB; ++n) {
        len[n] = (b0 + n < B) ? (lengths ? min(max(lengths[b0 + n], 0), T) : T) : 0;
        maxlen = max(maxlen, len[n]);
        c[n] = 0.f;
        h[n][j] = 0.f;
    }
    float bh[4];
#pragma unroll
    for (int g = 0; g < 4; ++g) bh[g] = bhh[(size_t)dir * 4 * H + g * H + j];
    __syncthreads();
    const size_t G = (size_t)dirs * 4 * H, Y = (size_t)dirs * H;
    for (int s = 0; s < maxlen; ++s) {
        float acc[4][TR_NB];
#pragma unroll
        for (int g = 0; g < 4; ++g)
#pragma unroll
            for (int n = 0; n < TR_NB; ++n) acc[g][n] = 0.f;
        for (int k = 0; k < H; ++k) {
            float w[4];
#pragma unroll
            for (int g = 0; g < 4; ++g) w[g] = __ldg(wt + (size_t)k * 4 * H + g * H + j);
#pragma unroll
            for (int n = 0; n < TR_NB; ++n) {
                const float hk = h[n][k];
#pragma unroll
                for (int g = 0; g < 4; ++g) acc[g][n] = fmaf(w[g], hk, acc[g][n]);
            }
        }
        __syncthreads();
#pragma unroll
        for (int n = 0; n < TR_NB; ++n) {
            if (s >= len[n]) continue;
            const int t = dir ? len[n] - 1 - s : s;
            const size_t row = (size_t)(b0 + n) * T + t;
            const float* gi = gin + row * G + (size_t)dir * 4 * H;
            const float iv = sigm(acc[0][n] + gi[j] + bh[0]), fv = sigm(acc[1][n] + gi[H + j] + bh[1]);
            const float gv = tanhf(acc[2][n] + gi[2 * H + j] + bh[2]), ov = sigm(acc[3][n] + gi[3 * H + j] + bh[3]);
            c[n] = fmaf(fv, c[n], iv * gv);
            const float hv = ov * tanhf(c[n]);
            float* go = gates + row * G + (size_t)dir * 4 * H;
            go[j] = iv; go[H + j] = fv; go[2 * H + j] = gv; go[3 * H + j] = ov;
            cs[row * Y + (size_t)dir * H + j] = c[n];
            y[row * Y + (size_t)dir * H + j] = hv;
            h[n][j] = hv;
        }
        __syncthreads();
    }
    // frames >= len: zero output (pad_packed_sequence, rnn.py:31)
#pragma unroll
    for (int n = 0; n < TR_NB; ++n)
        if (b0 + n < B)
            for (int t = len[n]; t < T; ++t) y[((size_t)(b0 + n) * T + t) * Y + (size_t)dir * H + j] = 0.f;
}

// Backward recurrence of one layer: dy (gradient w.r.t. the layer output) -> dG (gradient w.r.t. the gate PRE-activations, torch
// column order), walking every sequence against its processing order.  dh_{t-1} += W_hh^T dG_t is the per-step product
// (thread = input unit k, W_hh [4H, H] read row by row: coalesced).  dG must be zero-initialised (padded frames stay zero).
template <int H>
__global__ void __launch_bounds__(H) lstm_train_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ gates,
                                                          const float* __restrict__ cs, const float* __restrict__ whh,
                                                          float* __restrict__ dG, const int32_t* __restrict__ lengths, int B, int T, int dirs) {
    __shared__ float dg[TR_NB][4 * H];
    const int j = threadIdx.x, dir = blockIdx.y, b0 = blockIdx.x * TR_NB;
    const float* w = whh + (size_t)dir * 4 * H * H;           // torch layout [4H, H]
    int len[TR_NB], maxlen = 0;
    float dhrec[TR_NB], dc[TR_NB];
#pragma unroll
    for (int n = 0; n < TR_NB; ++n) {
        len[n] = (b0 + n < B) ? (lengths ? min(max(lengths[b0 + n], 0), T) : T) : 0;
        maxlen = max(maxlen, len[n]);
        dhrec[n] = 0.f;
        dc[n] = 0.f;
    }
    const size_t G = (size_t)dirs * 4 * H, Y = (size_t)dirs * H;
    for (int s = maxlen - 1; s >= 0; --s) {
#pragma unroll
        for (int n = 0; n < TR_NB; ++n) {
            float d[4] = {0.f, 0.f, 0.f, 0.f};
            if (s < len[n]) {
                const int t = dir ? len[n] - 1 - s : s;
                const size_t row = (size_t)(b0 + n) * T + t;
                const float* ga = gates + row * G + (size_t)dir * 4 * H;
                const float iv = ga[j], fv = ga[H + j], gv = ga[2 * H + j], ov = ga[3 * H + j];
                const float cv = cs[row * Y + (size_t)dir * H + j];
                const float cprev = s > 0 ? cs[((size_t)(b0 + n) * T + (dir ? t + 1 : t - 1)) * Y + (size_t)dir * H + j] : 0.f;
                const float dh = dy[row * Y + (size_t)dir * H + j] + dhrec[n];
                const float tc = tanhf(cv);
                const float dcv = fmaf(dh * ov, 1.f - tc * tc, dc[n]);
                d[0] = dcv * gv * iv * (1.f - iv);
                d[1] = dcv * cprev * fv * (1.f - fv);
                d[2] = dcv * iv * (1.f - gv * gv);
                d[3] = dh * tc * ov * (1.f - ov);
                dc[n] = dcv * fv;
                float* go = dG + row * G + (size_t)dir * 4 * H;
                go[j] = d[0]; go[H + j] = d[1]; go[2 * H + j] = d[2]; go[3 * H + j] = d[3];
            }
#pragma unroll
            for (int g = 0; g < 4; ++g) dg[n][g * H + j] = d[g];
        }
        __syncthreads();
        float acc[TR_NB];
#pragma unroll
        for (int n = 0; n < TR_NB; ++n) acc[n] = 0.f;
        for (int r = 0; r < 4 * H; ++r) {
            const float wv = __ldg(w + (size_t)r * H + j);
#pragma unroll
            for (int n = 0; n < TR_NB; ++n) acc[n] = fmaf(wv, dg[n][r], acc[n]);
        }
#pragma unroll
        for (int n = 0; n < TR_NB; ++n) dhrec[n] = acc[n];
        __syncthreads();
    }
}

// hprev[b, t, dir, :] = the hidden state the step at frame t started from: y[b, t-1] (forward), y[b, t+1] (reverse), zero at the start
__global__ void shift_prev_kernel(const float* __restrict__ y, const int32_t* __restrict__ lengths, float* __restrict__ hprev, int B, int T,
                                  int dirs, int H) {
    const size_t n = (size_t)B * T * dirs * H;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const int j = (int)(i % H), dir = (int)((i / H) % dirs);
        const size_t bt = i / ((size_t)dirs * H);
        const int t = (int)(bt % T), b = (int)(bt / T);
        const int len = lengths ? min(max(lengths[b], 0), T) : T;
        float v = 0.f;
        if (t < len) {
            const int tp = dir ? t + 1 : t - 1;
            if (tp >= 0 && tp < len) v = y[((size_t)b * T + tp) * dirs * H + (size_t)dir * H + j];
        }
        hprev[i] = v;
    }
}

// C[n1, n2] += sum_m A[m, n1] * B[m, n2]   (A, B row-major with leading dimensions lda / ldb; C zero-initialised, ldc); the M range is
// split over gridDim.z and combined with atomics
__global__ void __launch_bounds__(256) gemm_tn_kernel(const float* __restrict__ A, int lda, const float* __restrict__ Bm, int ldb,
                                                     float* __restrict__ C, int ldc, int M, int N1, int N2, int rows_per_split) {
    __shared__ float As[16][64], Bs[16][64];
    const int tid = threadIdx.x, ty = tid >> 4, tx = tid & 15;
    const int n1_0 = blockIdx.y * 64, n2_0 = blockIdx.x * 64;
    const int m_begin = blockIdx.z * rows_per_split, m_end = min(M, m_begin + rows_per_split);
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int k = 0; k < 4; ++k) acc[i][k] = 0.f;
    for (int m0 = m_begin; m0 < m_end; m0 += 16) {
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const int idx = tid + e * 256, mm = idx >> 6, cc = idx & 63;
            const int m = m0 + mm;
            As[mm][cc] = (m < m_end && n1_0 + cc < N1) ? A[(size_t)m * lda + n1_0 + cc] : 0.f;
            Bs[mm][cc] = (m < m_end && n2_0 + cc < N2) ? Bm[(size_t)m * ldb + n2_0 + cc] : 0.f;
        }
        __syncthreads();
#pragma unroll
        for (int mm = 0; mm < 16; ++mm) {
            float a[4], b[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) a[i] = As[mm][ty * 4 + i];
#pragma unroll
            for (int k = 0; k < 4; ++k) b[k] = Bs[mm][tx * 4 + k];
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int k = 0; k < 4; ++k) acc[i][k] = fmaf(a[i], b[k], acc[i][k]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int n1 = n1_0 + ty * 4 + i, n2 = n2_0 + tx * 4 + k;
            if (n1 < N1 && n2 < N2) atomicAdd(C + (size_t)n1 * ldc + n2, acc[i][k]);
        }
}

// out[n] += sum_m A[m, n]
__global__ void colsum_kernel(const float* __restrict__ A, int lda, float* __restrict__ out, int M, int N, int rows_per_split) {
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= N) return;
    const int m_begin = blockIdx.y * rows_per_split, m_end = min(M, m_begin + rows_per_split);
    float acc = 0.f;
    for (int m = m_begin; m < m_end; ++m) acc += A[(size_t)m * lda + n];
    atomicAdd(out + n, acc);
}

// block + grid reduction of a per-thread double into *out
__device__ __forceinline__ void reduce_add_double(double local, double* out) {
    __shared__ double red[32];
    for (int o = 16; o > 0; o >>= 1) local += __shfl_xor_sync(0xffffffffu, local, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = local;
    __syncthreads();
    if (threadIdx.x < 32) {
        double v = threadIdx.x < (blockDim.x >> 5) ? red[threadIdx.x] : 0.0;
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if (threadIdx.x == 0) atomicAdd(out, v);
    }
}

// joints.py:54-75 on a padded prediction [B, T, D]:  loss = mean((p - y)^2) + tw * mean_b sum_t |p[t+2] + p[t] - 2 p[t+1]|_1 ;
// dpred = 2 (p - y) / (B T D) + tw / B * (sign(acc[t-2]) + sign(acc[t]) - 2 sign(acc[t-1]))   (terms that exist)
__global__ void joints_loss_kernel(const float* __restrict__ pred, const float* __restrict__ target, int B, int T, int D, float tw,
                                   double* __restrict__ loss, float* __restrict__ dpred) {
    const size_t n = (size_t)B * T * D;
    double local = 0.0;
    auto sgn = [](float v) { return v > 0.f ? 1.f : (v < 0.f ? -1.f : 0.f); };
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const int t = (int)((i / D) % T);
        const float p = pred[i], e = p - target[i];
        float g = 2.f * e / (float)n;
        local += (double)e * e / (double)n;
        auto acc_at = [&](int t0) { return pred[i + (size_t)(t0 + 2 - t) * D] + pred[i + (size_t)(t0 - t) * D] - 2.f * pred[i + (size_t)(t0 + 1 - t) * D]; };
        float s = 0.f;
        if (t + 2 < T) {                       // acc[t] = p[t+2] + p[t] - 2 p[t+1]: this element is its `p[t]` term
            const float a = acc_at(t);
            s += sgn(a);
            local += (double)tw * fabsf(a) / (double)B;
        }
        if (t >= 2) s += sgn(acc_at(t - 2));                  // ... the `p[t+2]` term of acc[t-2]
        if (t >= 1 && t + 1 < T) s -= 2.f * sgn(acc_at(t - 1));   // ... the `-2 p[t+1]` term of acc[t-1]
        dpred[i] = g + tw / (float)B * s;
    }
    reduce_add_double(local, loss);
}

// footcontact.py:31,63: nn.BCEWithLogitsLoss() over every element of the padded [B, T, 2] logits:
//   loss = mean(max(p, 0) - p y + log(1 + exp(-|p|))),  dpred = (sigmoid(p) - y) / n
__global__ void bce_logits_loss_kernel(const float* __restrict__ pred, const float* __restrict__ target, size_t n, double* __restrict__ loss,
                                       float* __restrict__ dpred) {
    double local = 0.0;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const float p = pred[i], y = target[i];
        local += (double)(fmaxf(p, 0.f) - p * y + log1pf(expf(-fabsf(p)))) / (double)n;
        dpred[i] = (1.0f / (1.0f + expf(-p)) - y) / (float)n;
    }
    reduce_add_double(local, loss);
}

// velocity.py:72-86: sum over n in {1, 3, 9} of the MSE losses of the T // n windows of n frames; a window's MSE is a mean over
// B * n * D elements, so frame t carries the weight c_t = sum_n [t < n (T // n)] / (B n D):  loss = sum c_t e^2,  dpred = 2 c_t e
__global__ void velocity_loss_kernel(const float* __restrict__ pred, const float* __restrict__ target, int B, int T, int D,
                                     double* __restrict__ loss, float* __restrict__ dpred) {
    const size_t n = (size_t)B * T * D;
    double local = 0.0;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const int t = (int)((i / D) % T);
        float c = 0.f;
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            const int w = k == 0 ? 1 : (k == 1 ? 3 : 9);
            if (t < w * (T / w)) c += 1.0f / ((float)B * (float)w * (float)D);
        }
        const float e = pred[i] - target[i];
        local += (double)c * e * e;
        dpred[i] = 2.f * c * e;
    }
    reduce_add_double(local, loss);
}

// poser.py:65-98 on a padded prediction p [B, T, 96] (r6d of the 16 reduced joints, global rotations):
//   loss = mean((p - pose_t)^2) + tw * mean_b sum_t |p[t+3] - 3 p[t+2] + 3 p[t+1] - p[t]|_1
//          + mean((FK(reduced_global_to_full(p)) - joints_t)^2)          (use_pos_loss: poser.py:93-96)
// One thread per frame.  The position term: every reduced joint's global rotation is the Gram-Schmidt of its 6 numbers
// (angular.py:167-182), the other joints inherit their parent's (their local rotation is the identity, poser.py:55), joint positions
// follow the zero-pose bones down the tree (model.py:208-232); backward = subtree sums of d loss / d position, outer products with the
// bones into d loss / d rotation, then the Gram-Schmidt's adjoint.  (Composing local = parent^T child and recomposing it in the forward
// kinematics cancels identically because Gram-Schmidt always returns orthonormal matrices, so the derivative is that of the
// telescoped form.)
__global__ void poser_loss_kernel(const float* __restrict__ pred, const float* __restrict__ pose_t, const float* __restrict__ joints_t, int B,
                                  int T, float tw, double* __restrict__ loss, float* __restrict__ dpred) {
    constexpr int kParent[24] = MP_SMPL_PARENT_INIT, kSlot[24] = MP_REDUCED_SLOT_INIT;
    constexpr float kJ[24][3] = MP_SMPL_J_ZERO_INIT;
    const long long n_frames = (long long)B * T;
    double local = 0.0;
    auto sgn = [](float v) { return v > 0.f ? 1.f : (v < 0.f ? -1.f : 0.f); };
    for (long long f = (long long)blockIdx.x * blockDim.x + threadIdx.x; f < n_frames; f += (long long)gridDim.x * blockDim.x) {
        const int t = (int)(f % T);
        const float* p = pred + f * 96;
        float g[96];
        // (1) + (2): element-wise terms
        const float inv_n = 1.0f / ((float)n_frames * 96.f);
        for (int k = 0; k < 96; ++k) {
            const float e = p[k] - pose_t[f * 96 + k];
            local += (double)e * e * inv_n;
            float gk = 2.f * e * inv_n;
            auto jerk_at = [&](int t0) {
                const float* q = p + (long long)(t0 - t) * 96 + k;
                return q[3 * 96] - 3.f * q[2 * 96] + 3.f * q[96] - q[0];
            };
            float sj = 0.f;
            if (t + 3 < T) {
                const float j = jerk_at(t);
                sj -= sgn(j);
                local += (double)tw * fabsf(j) / (double)B;
            }
            if (t >= 1 && t + 2 < T) sj += 3.f * sgn(jerk_at(t - 1));
            if (t >= 2 && t + 1 < T) sj -= 3.f * sgn(jerk_at(t - 2));
            if (t >= 3) sj += sgn(jerk_at(t - 3));
            g[k] = gk + tw / (float)B * sj;
        }
        // (3) position term
        float G[24][9], a_n[16], u_n[16], c0b[16];
        for (int j = 0; j < 24; ++j) {
            const int sl = kSlot[j];
            if (sl < 0) {
                for (int i = 0; i < 9; ++i) G[j][i] = G[kParent[j]][i];
                continue;
            }
            const float ax = p[sl * 6], ay = p[sl * 6 + 1], az = p[sl * 6 + 2], bx = p[sl * 6 + 3], by = p[sl * 6 + 4], bz = p[sl * 6 + 5];
            const float na = sqrtf(ax * ax + ay * ay + az * az);
            const float c0x = ax / na, c0y = ay / na, c0z = az / na;
            const float d = c0x * bx + c0y * by + c0z * bz;
            const float ux = bx - d * c0x, uy = by - d * c0y, uz = bz - d * c0z;
            const float nu = sqrtf(ux * ux + uy * uy + uz * uz);
            const float c1x = ux / nu, c1y = uy / nu, c1z = uz / nu;
            const float c2x = c0y * c1z - c0z * c1y, c2y = c0z * c1x - c0x * c1z, c2z = c0x * c1y - c0y * c1x;
            // columns c0, c1, c2; row major R[r][c]
            G[j][0] = c0x; G[j][1] = c1x; G[j][2] = c2x;
            G[j][3] = c0y; G[j][4] = c1y; G[j][5] = c2y;
            G[j][6] = c0z; G[j][7] = c1z; G[j][8] = c2z;
            a_n[sl] = na; u_n[sl] = nu; c0b[sl] = d;
        }
        float P[24][3], S[24][3];
        P[0][0] = P[0][1] = P[0][2] = 0.f;
        const float inv_p = 1.0f / ((float)n_frames * 72.f);
        for (int j = 1; j < 24; ++j) {
            const int pa = kParent[j];
            const float b0 = kJ[j][0] - kJ[pa][0], b1 = kJ[j][1] - kJ[pa][1], b2 = kJ[j][2] - kJ[pa][2];
            for (int r = 0; r < 3; ++r) P[j][r] = P[pa][r] + G[pa][3 * r] * b0 + G[pa][3 * r + 1] * b1 + G[pa][3 * r + 2] * b2;
        }
        for (int j = 0; j < 24; ++j)
            for (int r = 0; r < 3; ++r) {
                const float e = P[j][r] - joints_t[f * 72 + j * 3 + r];
                local += (double)e * e * inv_p;
                S[j][r] = 2.f * e * inv_p;
            }
        for (int j = 23; j >= 1; --j)
            for (int r = 0; r < 3; ++r) S[kParent[j]][r] += S[j][r];          // subtree sums of d loss / d position
        float dG[24][9];
        for (int j = 0; j < 24; ++j)
            for (int i = 0; i < 9; ++i) dG[j][i] = 0.f;
        for (int j = 1; j < 24; ++j) {
            const int pa = kParent[j];
            const float b[3] = {kJ[j][0] - kJ[pa][0], kJ[j][1] - kJ[pa][1], kJ[j][2] - kJ[pa][2]};
            for (int r = 0; r < 3; ++r)
                for (int c = 0; c < 3; ++c) dG[pa][3 * r + c] += S[j][r] * b[c];
        }
        for (int j = 23; j >= 1; --j)                                          // a non-reduced joint's rotation IS its parent's
            if (kSlot[j] < 0)
                for (int i = 0; i < 9; ++i) dG[kParent[j]][i] += dG[j][i];
        for (int j = 0; j < 24; ++j) {
            const int sl = kSlot[j];
            if (sl < 0) continue;
            const float c0[3] = {G[j][0], G[j][3], G[j][6]}, c1[3] = {G[j][1], G[j][4], G[j][7]};
            float d0[3] = {dG[j][0], dG[j][3], dG[j][6]}, d1[3] = {dG[j][1], dG[j][4], dG[j][7]};
            const float d2[3] = {dG[j][2], dG[j][5], dG[j][8]};
            // c2 = c0 x c1:  d0 += c1 x d2,  d1 += d2 x c0
            d0[0] += c1[1] * d2[2] - c1[2] * d2[1]; d0[1] += c1[2] * d2[0] - c1[0] * d2[2]; d0[2] += c1[0] * d2[1] - c1[1] * d2[0];
            d1[0] += d2[1] * c0[2] - d2[2] * c0[1]; d1[1] += d2[2] * c0[0] - d2[0] * c0[2]; d1[2] += d2[0] * c0[1] - d2[1] * c0[0];
            // c1 = u / |u|
            const float c1d1 = c1[0] * d1[0] + c1[1] * d1[1] + c1[2] * d1[2];
            float du[3];
            for (int r = 0; r < 3; ++r) du[r] = (d1[r] - c1[r] * c1d1) / u_n[sl];
            // u = b - (c0 . b) c0
            const float c0du = c0[0] * du[0] + c0[1] * du[1] + c0[2] * du[2];
            const float bvec[3] = {p[sl * 6 + 3], p[sl * 6 + 4], p[sl * 6 + 5]};
            float db[3];
            for (int r = 0; r < 3; ++r) {
                db[r] = du[r] - c0[r] * c0du;
                d0[r] += -c0b[sl] * du[r] - c0du * bvec[r];
            }
            // c0 = a / |a|
            const float c0d0 = c0[0] * d0[0] + c0[1] * d0[1] + c0[2] * d0[2];
            for (int r = 0; r < 3; ++r) {
                g[sl * 6 + r] += (d0[r] - c0[r] * c0d0) / a_n[sl];
                g[sl * 6 + 3 + r] += db[r];
            }
        }
        for (int k = 0; k < 96; ++k) dpred[f * 96 + k] = g[k];
    }
    reduce_add_double(local, loss);
}

inline size_t align_up(size_t x, size_t a = 256) { return (x + a - 1) / a * a; }

struct TrainLayout {
    size_t x1, x1m, gates[2], cs[2], y[2], gin, dG, dY, dX1, hprev, wT[2], wcat[2], bcat[2], bhh[2], wcatT[2], w2T, zeros, total;
};

TrainLayout train_layout(const mp_rnn_weights_t* w, size_t M) {
    TrainLayout L;
    const size_t H = w->n_hidden, D = w->bidirectional ? 2 : 1;
    size_t off = 0;
    auto take = [&](size_t floats) { size_t o = off; off = align_up(off + floats * sizeof(float)); return o; };
    L.x1 = take(M * H); L.x1m = take(M * H);
    for (int l = 0; l < 2; ++l) { L.gates[l] = take(M * D * 4 * H); L.cs[l] = take(M * D * H); L.y[l] = take(M * D * H); }
    L.gin = take(M * D * 4 * H); L.dG = take(M * D * 4 * H); L.dY = take(M * D * H); L.dX1 = take(M * H); L.hprev = take(M * D * H);
    for (int l = 0; l < 2; ++l) {
        const size_t in = l == 0 ? H : D * H;
        L.wT[l] = take(D * H * 4 * H); L.wcat[l] = take(D * 4 * H * in); L.bcat[l] = take(D * 4 * H); L.bhh[l] = take(D * 4 * H);
        L.wcatT[l] = take(in * D * 4 * H);
    }
    L.w2T = take(D * H * (size_t)w->n_output);
    L.zeros = take(std::max<size_t>(D * 4 * H, 1024));
    L.total = off;
    return L;
}

int check_train_args(const mp_rnn_weights_t* w, const float* x, int B, int T, void* ws, size_t ws_bytes) {
    MP_REQUIRE(w && x && ws, "rnn_train: null argument");
    MP_REQUIRE(w->n_layers == 2 && (w->n_hidden == 256 || w->n_hidden == 64), "rnn_train: 2 layers, hidden 64 or 256");
    MP_REQUIRE((w->n_input & 3) == 0 && w->n_output > 0, "rnn_train: n_input must be a multiple of 4");
    MP_REQUIRE(B > 0 && T > 0, "rnn_train: empty batch");
    MP_REQUIRE(((uintptr_t)ws & 255) == 0, "rnn_train: workspace must be 256-byte aligned");
    const size_t need = train_layout(w, (size_t)B * T).total;
    if (ws_bytes < need) {
        set_error("rnn_train: workspace %zu < required %zu", ws_bytes, need);
        return MP_ERR_WORKSPACE;
    }
    return MP_OK;
}

int transpose(const float* in, float* out, int R, int C, cudaStream_t s) {
    transpose_kernel<<<dim3((C + 31) / 32, (R + 31) / 32), dim3(32, 8), 0, s>>>(in, out, R, C);
    MP_CUDA_TRY(cudaGetLastError());
    return MP_OK;
}

int gemm_tn(const float* A, int lda, const float* Bm, int ldb, float* C, int ldc, int M, int N1, int N2, cudaStream_t s) {
    MP_CUDA_TRY(cudaMemsetAsync(C, 0, (size_t)N1 * ldc * sizeof(float), s));
    const int splits = std::max(1, std::min(64, M / 512));
    const int rows = ((M + splits - 1) / splits + 15) / 16 * 16;
    gemm_tn_kernel<<<dim3((N2 + 63) / 64, (N1 + 63) / 64, (M + rows - 1) / rows), 256, 0, s>>>(A, lda, Bm, ldb, C, ldc, M, N1, N2, rows);
    MP_CUDA_TRY(cudaGetLastError());
    return MP_OK;
}

int colsum(const float* A, int lda, float* out, int M, int N, cudaStream_t s) {
    MP_CUDA_TRY(cudaMemsetAsync(out, 0, (size_t)N * sizeof(float), s));
    const int splits = std::max(1, std::min(64, M / 256));
    const int rows = (M + splits - 1) / splits;
    colsum_kernel<<<dim3((N + 127) / 128, (M + rows - 1) / rows), 128, 0, s>>>(A, lda, out, M, N, rows);
    MP_CUDA_TRY(cudaGetLastError());
    return MP_OK;
}

template <typename F>
int per_hidden(int H, F&& f) {
    if (H == 256) return f(std::integral_constant<int, 256>());
    return f(std::integral_constant<int, 64>());
}

}  // namespace

size_t rnn_train_workspace_bytes(const mp_rnn_weights_t* w, int B, int T) {
    if (!w || B <= 0 || T <= 0) return 0;
    return train_layout(w, (size_t)B * T).total;
}

int rnn_train_forward(const mp_rnn_weights_t* w, const float* x, int B, int T, const int32_t* lengths, const float* mask, float* yout,
                      void* ws_, size_t ws_bytes, cudaStream_t s) {
    MP_TRY(check_train_args(w, x, B, T, ws_, ws_bytes));
    MP_REQUIRE(yout, "rnn_train_forward: null output");
    const int H = w->n_hidden, D = w->bidirectional ? 2 : 1, M = B * T;
    const TrainLayout L = train_layout(w, (size_t)M);
    char* ws = (char*)ws_;
    auto F = [&](size_t o) { return (float*)(ws + o); };
    MP_CUDA_TRY(cudaMemsetAsync(F(L.zeros), 0, std::max<size_t>((size_t)D * 4 * H, 1024) * sizeof(float), s));
    // linear1 + ReLU, then the dropout mask (rnn.py:22)
    MP_TRY(launch_gemm_ffma(x, w->n_input, nullptr, 0, w->linear1_w, w->linear1_b, F(L.x1), M, H, 1, s));
    const float* in = F(L.x1);
    if (mask) {
        mul_kernel<<<296, 256, 0, s>>>(F(L.x1), mask, F(L.x1m), (size_t)M * H);
        MP_CUDA_TRY(cudaGetLastError());
        in = F(L.x1m);
    }
    int in_w = H;
    for (int l = 0; l < 2; ++l) {
        // stacked W_ih / b_ih of the directions, transposed W_hh, b_hh (prepared per call: training changes the weights every step)
        for (int d = 0; d < D; ++d) {
            MP_CUDA_TRY(cudaMemcpyAsync(F(L.wcat[l]) + (size_t)d * 4 * H * in_w, w->w_ih[l][d], (size_t)4 * H * in_w * 4, cudaMemcpyDeviceToDevice, s));
            MP_CUDA_TRY(cudaMemcpyAsync(F(L.bcat[l]) + (size_t)d * 4 * H, w->b_ih[l][d], (size_t)4 * H * 4, cudaMemcpyDeviceToDevice, s));
            MP_CUDA_TRY(cudaMemcpyAsync(F(L.bhh[l]) + (size_t)d * 4 * H, w->b_hh[l][d], (size_t)4 * H * 4, cudaMemcpyDeviceToDevice, s));
            MP_TRY(transpose(w->w_hh[l][d], F(L.wT[l]) + (size_t)d * H * 4 * H, 4 * H, H, s));
        }
        MP_TRY(launch_gemm_ffma(in, in_w, nullptr, 0, F(L.wcat[l]), F(L.bcat[l]), F(L.gin), M, D * 4 * H, 0, s));
        const dim3 grid((B + TR_NB - 1) / TR_NB, D);
        MP_TRY(per_hidden(H, [&](auto hc) {
            lstm_train_fwd_kernel<decltype(hc)::value><<<grid, decltype(hc)::value, 0, s>>>(F(L.gin), F(L.bhh[l]), F(L.wT[l]), F(L.y[l]), F(L.gates[l]),
                                                                                     F(L.cs[l]), lengths, B, T, D);
            return cudaGetLastError() == cudaSuccess ? MP_OK : MP_ERR_CUDA;
        }));
        count_launch();
        in = F(L.y[l]);
        in_w = D * H;
    }
    MP_TRY(launch_gemm_ffma(in, in_w, nullptr, 0, w->linear2_w, w->linear2_b, yout, M, w->n_output, 0, s));
    return MP_OK;
}

int rnn_train_backward(const mp_rnn_weights_t* w, const float* x, int B, int T, const int32_t* lengths, const float* mask, const float* dy,
                       const mp_rnn_grads_t* g, void* ws_, size_t ws_bytes, cudaStream_t s) {
    MP_TRY(check_train_args(w, x, B, T, ws_, ws_bytes));
    MP_REQUIRE(dy && g && g->linear1_w && g->linear1_b && g->linear2_w && g->linear2_b, "rnn_train_backward: null argument");
    const int H = w->n_hidden, D = w->bidirectional ? 2 : 1, M = B * T, NO = w->n_output;
    const TrainLayout L = train_layout(w, (size_t)M);
    char* ws = (char*)ws_;
    auto F = [&](size_t o) { return (float*)(ws + o); };
    const float* x1in = mask ? F(L.x1m) : F(L.x1);
    // linear2: dW2 = dy^T y1, db2 = colsum(dy), dY1 = dy W2
    MP_TRY(gemm_tn(dy, NO, F(L.y[1]), D * H, g->linear2_w, D * H, M, NO, D * H, s));
    MP_TRY(colsum(dy, NO, g->linear2_b, M, NO, s));
    if ((NO & 3) == 0) {
        MP_TRY(transpose(w->linear2_w, F(L.w2T), NO, D * H, s));                  // [DH, NO]
        MP_TRY(launch_gemm_ffma(dy, NO, nullptr, 0, F(L.w2T), F(L.zeros), F(L.dY), M, D * H, 0, s));
    } else {                                                                      // a handful of outputs (foot-contact logits)
        gemm_small_k_kernel<<<296, 256, 0, s>>>(dy, w->linear2_w, F(L.dY), (size_t)M, NO, D * H);
        MP_CUDA_TRY(cudaGetLastError());
    }
    for (int l = 1; l >= 0; --l) {
        const int in_w = l == 0 ? H : D * H;
        const float* lin = l == 0 ? x1in : F(L.y[0]);
        MP_CUDA_TRY(cudaMemsetAsync(F(L.dG), 0, (size_t)M * D * 4 * H * sizeof(float), s));
        const dim3 grid((B + TR_NB - 1) / TR_NB, D);
        // the kernel wants the raw (torch-layout) W_hh of both directions contiguous: stacked into the region that held the transposed
        // copy of the forward pass (same size; the next forward rebuilds it)
        float* whh_cat = F(L.wT[l]);
        for (int d = 0; d < D; ++d)
            MP_CUDA_TRY(cudaMemcpyAsync(whh_cat + (size_t)d * 4 * H * H, w->w_hh[l][d], (size_t)4 * H * H * 4, cudaMemcpyDeviceToDevice, s));
        MP_TRY(per_hidden(H, [&](auto hc) {
            lstm_train_bwd_kernel<decltype(hc)::value><<<grid, decltype(hc)::value, 0, s>>>(F(L.dY), F(L.gates[l]), F(L.cs[l]), whh_cat, F(L.dG), lengths, B, T, D);
            return cudaGetLastError() == cudaSuccess ? MP_OK : MP_ERR_CUDA;
        }));
        count_launch();
        shift_prev_kernel<<<296, 256, 0, s>>>(F(L.y[l]), lengths, F(L.hprev), B, T, D, H);
        MP_CUDA_TRY(cudaGetLastError());
        for (int d = 0; d < D; ++d) {
            MP_REQUIRE(g->w_ih[l][d] && g->w_hh[l][d] && g->b_ih[l][d] && g->b_hh[l][d], "rnn_train_backward: null LSTM gradient (layer %d dir %d)", l, d);
            const float* dGd = F(L.dG) + (size_t)d * 4 * H;
            MP_TRY(gemm_tn(dGd, D * 4 * H, lin, in_w, g->w_ih[l][d], in_w, M, 4 * H, in_w, s));
            MP_TRY(gemm_tn(dGd, D * 4 * H, F(L.hprev) + (size_t)d * H, D * H, g->w_hh[l][d], H, M, 4 * H, H, s));
            MP_TRY(colsum(dGd, D * 4 * H, g->b_ih[l][d], M, 4 * H, s));
            MP_CUDA_TRY(cudaMemcpyAsync(g->b_hh[l][d], g->b_ih[l][d], (size_t)4 * H * 4, cudaMemcpyDeviceToDevice, s));
        }
        // gradient w.r.t. the layer input: dIn = dG [M, D4H] . Wcat [D4H, in]
        MP_TRY(transpose(F(L.wcat[l]), F(L.wcatT[l]), D * 4 * H, in_w, s));       // [in, D4H]
        float* din = l == 0 ? F(L.dX1) : F(L.dY);
        MP_TRY(launch_gemm_ffma(F(L.dG), D * 4 * H, nullptr, 0, F(L.wcatT[l]), F(L.zeros), din, M, in_w, 0, s));
    }
    // dropout mask and ReLU (rnn.py:22), then linear1: dW1 = dz^T x, db1 = colsum(dz)
    relu_mask_bwd_kernel<<<296, 256, 0, s>>>(F(L.dX1), F(L.x1), mask, F(L.dX1), (size_t)M * H);
    MP_CUDA_TRY(cudaGetLastError());
    MP_TRY(gemm_tn(F(L.dX1), H, x, w->n_input, g->linear1_w, w->n_input, M, H, w->n_input, s));
    MP_TRY(colsum(F(L.dX1), H, g->linear1_b, M, H, s));
    return MP_OK;
}

// ---- optimizer: what Lightning runs between shared_step's backward and the next step (overfit.py:41-50: gradient_clip_val = 1;
// joints.py:113-114: torch.optim.AdamW(lr = 1e-3) with torch's defaults) ----------------------------------------------------------
// sum of squares of a flat gradient buffer, accumulated into a double (torch.nn.utils.clip_grad_norm_'s total norm, squared)
__global__ void __launch_bounds__(256) grad_sq_norm_kernel(const float* __restrict__ g, size_t n, double* __restrict__ out) {
    double acc = 0.0;
    const size_t n4 = n >> 2;
    const float4* g4 = reinterpret_cast<const float4*>(g);
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) {
        const float4 v = __ldg(g4 + i);
        acc += (double)v.x * v.x + (double)v.y * v.y + (double)v.z * v.z + (double)v.w * v.w;
    }
    if (blockIdx.x == 0 && threadIdx.x < (n & 3)) {
        const float v = g[(n4 << 2) + threadIdx.x];
        acc += (double)v * v;
    }
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    __shared__ double part[8];
    if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int w = 0; w < 8; ++w) t += part[w];
        atomicAdd(out, t);
    }
}

// One fused pass over the flat parameter buffer: gradient scaling (mean over ranks x clip coefficient), decoupled weight decay, both
// moment updates and the parameter step, in torch.optim.AdamW's operation order (single-tensor path, amsgrad off):
//   p *= 1 - lr wd;  m += (g - m)(1 - b1);  v = v b2 + (1 - b2) g g;  p -= (lr / bc1) m / (sqrt(v) / sqrt(bc2) + eps)
// 20 B read + 12 B written per parameter: HBM-bound streaming (float4, grid-stride).
__global__ void __launch_bounds__(256) adamw_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                                                    float* __restrict__ v, size_t n, float lr, float b1, float b2, float eps, float wd,
                                                    float bc1, float sqrt_bc2, const double* __restrict__ sq_norm, float max_norm,
                                                    float grad_scale) {
    float gs = grad_scale;
    if (sq_norm) {            // clip_grad_norm_: coefficient max_norm / (total_norm + 1e-6), clamped to 1
        const float total = (float)sqrt(*sq_norm) * grad_scale;
        gs *= fminf(max_norm / (total + 1e-6f), 1.0f);
    }
    const float decay = 1.0f - lr * wd, step = lr / bc1;
    auto upd = [&](float& pp, float gg, float& mm, float& vv) {
        gg *= gs;
        pp *= decay;
        mm = mm + (gg - mm) * (1.0f - b1);
        vv = vv * b2 + (1.0f - b2) * gg * gg;
        pp = pp - step * (mm / (sqrtf(vv) / sqrt_bc2 + eps));
    };
    const size_t n4 = n >> 2;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) {
        float4 pp = reinterpret_cast<float4*>(p)[i], mm = reinterpret_cast<float4*>(m)[i], vv = reinterpret_cast<float4*>(v)[i];
        const float4 gg = __ldg(reinterpret_cast<const float4*>(g) + i);
        upd(pp.x, gg.x, mm.x, vv.x); upd(pp.y, gg.y, mm.y, vv.y); upd(pp.z, gg.z, mm.z, vv.z); upd(pp.w, gg.w, mm.w, vv.w);
        reinterpret_cast<float4*>(p)[i] = pp; reinterpret_cast<float4*>(m)[i] = mm; reinterpret_cast<float4*>(v)[i] = vv;
    }
    if (blockIdx.x == 0 && threadIdx.x < (n & 3)) {
        const size_t i = (n4 << 2) + threadIdx.x;
        upd(p[i], g[i], m[i], v[i]);
    }
}

int grad_sq_norm(const float* g, size_t n, double* out, cudaStream_t s) {
    MP_REQUIRE(g && out && n > 0 && ((uintptr_t)g & 15) == 0, "grad_sq_norm: bad arguments (the buffer must be 16-byte aligned)");
    grad_sq_norm_kernel<<<(unsigned)std::min<size_t>((n / 4 + 255) / 256 + 1, 148 * 8), 256, 0, s>>>(g, n, out);
    MP_CUDA_TRY(cudaGetLastError());
    count_launch();
    return MP_OK;
}

int adamw_step(float* p, const float* g, float* m, float* v, size_t n, float lr, float beta1, float beta2, float eps, float weight_decay,
               int step, const double* sq_norm, float max_norm, float grad_scale, cudaStream_t s) {
    MP_REQUIRE(p && g && m && v && n > 0 && step >= 1, "adamw_step: bad arguments");
    MP_REQUIRE((((uintptr_t)p | (uintptr_t)g | (uintptr_t)m | (uintptr_t)v) & 15) == 0, "adamw_step: the buffers must be 16-byte aligned");
    const double bc1 = 1.0 - pow((double)beta1, step), bc2 = 1.0 - pow((double)beta2, step);
    ProfileScope prof("adamw", 32.0 * (double)n, s);
    adamw_kernel<<<(unsigned)std::min<size_t>((n / 4 + 255) / 256 + 1, 148 * 8), 256, 0, s>>>(p, g, m, v, n, lr, beta1, beta2, eps, weight_decay,
                                                                                             (float)bc1, (float)sqrt(bc2), sq_norm, max_norm,
                                                                                             grad_scale);
    MP_CUDA_TRY(cudaGetLastError());
    count_launch();
    return MP_OK;
}

int joints_loss(const float* pred, const float* target, int B, int T, int D, float t_weight, double* loss, float* dpred, cudaStream_t s) {
    MP_REQUIRE(pred && target && loss && dpred && B > 0 && T > 0 && D > 0, "joints_loss: bad arguments");
    MP_CUDA_TRY(cudaMemsetAsync(loss, 0, sizeof(double), s));
    const size_t n = (size_t)B * T * D;
    joints_loss_kernel<<<(unsigned)std::min<size_t>((n + 255) / 256, 148 * 8), 256, 0, s>>>(pred, target, B, T, D, t_weight, loss, dpred);
    MP_CUDA_TRY(cudaGetLastError());
    count_launch();
    return MP_OK;
}

int poser_loss(const float* pred, const float* pose_t, const float* joints_t, int B, int T, float t_weight, double* loss, float* dpred,
               cudaStream_t s) {
    MP_REQUIRE(pred && pose_t && joints_t && loss && dpred && B > 0 && T > 0, "poser_loss: bad arguments");
    MP_CUDA_TRY(cudaMemsetAsync(loss, 0, sizeof(double), s));
    const long long n = (long long)B * T;
    poser_loss_kernel<<<(unsigned)std::min<long long>((n + 63) / 64, 148 * 16), 64, 0, s>>>(pred, pose_t, joints_t, B, T, t_weight, loss, dpred);
    MP_CUDA_TRY(cudaGetLastError());
    count_launch();
    return MP_OK;
}

int footcontact_loss(const float* pred, const float* target, int B, int T, double* loss, float* dpred, cudaStream_t s) {
    MP_REQUIRE(pred && target && loss && dpred && B > 0 && T > 0, "footcontact_loss: bad arguments");
    MP_CUDA_TRY(cudaMemsetAsync(loss, 0, sizeof(double), s));
    const size_t n = (size_t)B * T * 2;
    bce_logits_loss_kernel<<<(unsigned)std::min<size_t>((n + 255) / 256, 148 * 8), 256, 0, s>>>(pred, target, n, loss, dpred);
    MP_CUDA_TRY(cudaGetLastError());
    count_launch();
    return MP_OK;
}

int velocity_loss(const float* pred, const float* target, int B, int T, int D, double* loss, float* dpred, cudaStream_t s) {
    MP_REQUIRE(pred && target && loss && dpred && B > 0 && T > 0 && D > 0, "velocity_loss: bad arguments");
    MP_CUDA_TRY(cudaMemsetAsync(loss, 0, sizeof(double), s));
    const size_t n = (size_t)B * T * D;
    velocity_loss_kernel<<<(unsigned)std::min<size_t>((n + 255) / 256, 148 * 8), 256, 0, s>>>(pred, target, B, T, D, loss, dpred);
    MP_CUDA_TRY(cudaGetLastError());
    count_launch();
    return MP_OK;
}

}  // namespace mp
