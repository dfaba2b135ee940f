// Kinematic tail of the hot path: r6d -> local SMPL rotations (K5), offline translation fusion
// with the floor clamp (K6) and the online per-tick state machine (K7).
//
// Reference (relative to /root/reference/mobileposer):
//   K5  models/net.py:93-99, articulate/math/angular.py:167-182, articulate/math/general.py:27-39,
//       utils/model_utils.py:18-25, articulate/math/spatial.py:115-123,197-221
//   K6  models/net.py:125-154, articulate/math/general.py:15-24
//   K7  models/net.py:84-88,173-219
// All three are HBM-bound streaming kernels (384 B in / 864 B out per frame for K5); they carry no
// reuse, so no shared-memory staging beyond the per-warp scan buffers of K6.
#include "mp_common.cuh"
#include "mp_constants.cuh"

namespace mp {

namespace {

// joint_set.reduced / ignored (config.py:134-135), the SMPL tree and the zero-pose feet: mp_constants.cuh
__constant__ int c_parent[24] = MP_SMPL_PARENT_INIT;
__constant__ int c_reduced_slot[24] = MP_REDUCED_SLOT_INIT;
constexpr int kIgnored[9] = MP_IGNORED_INIT;
constexpr unsigned ignored_mask() {
    unsigned m = 0;
    for (int i = 0; i < 9; ++i) m |= 1u << kIgnored[i];
    return m;
}
__constant__ unsigned c_ignored_mask = ignored_mask();
__constant__ float c_feet[6] = MP_FEET_INIT;

__device__ __forceinline__ float nan_to_zero(float x) { return (x != x) ? 0.f : x; }

// One lane = one joint.  Computes the 24 local rotations of one frame; result in R[9] (row major).
__device__ __forceinline__ void frame_local_rotations(const float* __restrict__ r6d_frame, int lane, float R[9]) {
    const int j = lane < 24 ? lane : 0;
    float G[9] = {1.f, 0.f, 0.f, 0.f, 1.f, 0.f, 0.f, 0.f, 1.f};
    const int slot = c_reduced_slot[j];
    if (lane < 24 && slot >= 0) {
        const float2* src = reinterpret_cast<const float2*>(r6d_frame + slot * 6);
        const float2 v01 = __ldg(src), v23 = __ldg(src + 1), v45 = __ldg(src + 2);
        const float ax = v01.x, ay = v01.y, az = v23.x, bx = v23.y, by = v45.x, bz = v45.y;
        // normalize_tensor: x / ||x||  (no epsilon: a zero column yields NaN, zeroed below)
        const float na = sqrtf(__fadd_rn(__fadd_rn(__fmul_rn(ax, ax), __fmul_rn(ay, ay)), __fmul_rn(az, az)));
        const float c0x = ax / na, c0y = ay / na, c0z = az / na;
        const float d = __fadd_rn(__fadd_rn(__fmul_rn(c0x, bx), __fmul_rn(c0y, by)), __fmul_rn(c0z, bz));
        const float ux = __fsub_rn(bx, __fmul_rn(d, c0x)), uy = __fsub_rn(by, __fmul_rn(d, c0y)),
                    uz = __fsub_rn(bz, __fmul_rn(d, c0z));
        const float nu = sqrtf(__fadd_rn(__fadd_rn(__fmul_rn(ux, ux), __fmul_rn(uy, uy)), __fmul_rn(uz, uz)));
        const float c1x = ux / nu, c1y = uy / nu, c1z = uz / nu;
        const float c2x = __fsub_rn(__fmul_rn(c0y, c1z), __fmul_rn(c0z, c1y));
        const float c2y = __fsub_rn(__fmul_rn(c0z, c1x), __fmul_rn(c0x, c1z));
        const float c2z = __fsub_rn(__fmul_rn(c0x, c1y), __fmul_rn(c0y, c1x));
        // columns stacked on the last dim: R[r][0]=c0[r], R[r][1]=c1[r], R[r][2]=c2[r]; NaN -> 0
        G[0] = nan_to_zero(c0x); G[1] = nan_to_zero(c1x); G[2] = nan_to_zero(c2x);
        G[3] = nan_to_zero(c0y); G[4] = nan_to_zero(c1y); G[5] = nan_to_zero(c2y);
        G[6] = nan_to_zero(c0z); G[7] = nan_to_zero(c1z); G[8] = nan_to_zero(c2z);
    }
    // parent's global rotation
    const int par = c_parent[j] < 0 ? 0 : c_parent[j];
    float P[9];
#pragma unroll
    for (int i = 0; i < 9; ++i) P[i] = __shfl_sync(0xffffffffu, G[i], par);
    const bool ignored = (c_ignored_mask >> j) & 1u;
    if (j == 0) {
#pragma unroll
        for (int i = 0; i < 9; ++i) R[i] = G[i];                 // root keeps its global rotation
    } else if (ignored) {
#pragma unroll
        for (int i = 0; i < 9; ++i) R[i] = (i % 4 == 0) ? 1.f : 0.f;
    } else {
        // local = parent^T * global  (spatial.py:121: bmm(transpose(parent), child))
#pragma unroll
        for (int r = 0; r < 3; ++r)
#pragma unroll
            for (int c = 0; c < 3; ++c)
                R[r * 3 + c] = fmaf(P[6 + r], G[6 + c], fmaf(P[3 + r], G[3 + c], __fmul_rn(P[r], G[c])));
    }
}

__global__ void __launch_bounds__(256) reduced_global_to_full_kernel(const float* __restrict__ r6d, long long n,
                                                                     float* __restrict__ pose) {
    const int lane = threadIdx.x & 31;
    const long long warps = ((long long)gridDim.x * blockDim.x) >> 5;
    for (long long f = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5; f < n; f += warps) {
        float R[9];
        frame_local_rotations(r6d + f * 96, lane, R);
        if (lane < 24) {
            float* dst = pose + f * 216 + lane * 9;
#pragma unroll
            for (int i = 0; i < 9; ++i) dst[i] = R[i];
        }
    }
}

__device__ __forceinline__ float prob_to_weight(float p) {
    // (p.clamp(0.5, 0.9) - 0.5) / (0.9 - 0.5)   (net.py:90-91; the divisor is the double 0.9-0.5 cast to fp32)
    const float lo = kProbLo, hi = kProbHi;
    const float c = fminf(fmaxf(p, lo), hi);
    return __fsub_rn(c, lo) / (float)(0.9 - 0.5);
}

// K6: one warp per sequence, 32 frames per pass.
__global__ void __launch_bounds__(128) tran_offline_kernel(const float* __restrict__ joints, const float* __restrict__ vel,
                                                           const float* __restrict__ contact,
                                                           const int32_t* __restrict__ lengths, int B, int T,
                                                           float* __restrict__ tran) {
    __shared__ float s_vy[4][32];
    __shared__ float s_my[4][32];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int b = blockIdx.x * 4 + wib;
    if (b >= B) return;
    const int len = lengths ? min(max(lengths[b], 0), T) : T;
    const float* J = joints + (size_t)b * T * 72;
    const float* V = vel + (size_t)b * T * 72;
    const float* Cn = contact + (size_t)b * T * 2;
    float* out = tran + (size_t)b * T * 3;
    double cur = 0.0;                    // current_root_y, a Python float in the reference
    double cx = 0.0, cy = 0.0, cz = 0.0;  // running prefix sums
    for (int t0 = 0; t0 < len; t0 += 32) {
        const int t = t0 + lane;
        const bool valid = t < len;
        float vx = 0.f, vy = 0.f, vz = 0.f, my = 0.f;
        if (valid) {
            const float* jt = J + (size_t)t * 72;
            const float lx = jt[30], ly = jt[31], lz = jt[32], rx = jt[33], ry = jt[34], rz = jt[35];
            float dlx = 0.f, dly = 0.f, dlz = 0.f, drx = 0.f, dry = 0.f, drz = 0.f;
            if (t > 0) {
                const float* jp = jt - 72;
                dlx = __fsub_rn(jp[30], lx); dly = __fsub_rn(jp[31], ly); dlz = __fsub_rn(jp[32], lz);
                drx = __fsub_rn(jp[33], rx); dry = __fsub_rn(jp[34], ry); drz = __fsub_rn(jp[35], rz);
            }
            const float c0 = Cn[t * 2], c1 = Cn[t * 2 + 1];
            const bool right = c1 > c0;                    // argmax, first index on ties
            const float cvx = right ? drx : dlx;
            const float cvy = __fadd_rn(kGravityVel, right ? dry : dly);
            const float cvz = right ? drz : dlz;
            const float* vt = V + (size_t)t * 72;
            const float px = vt[0] / kVelDiv, py = vt[1] / kVelDiv, pz = vt[2] / kVelDiv;
            const float w = prob_to_weight(sigmoidf_acc(fmaxf(c0, c1)));
            const float om = __fsub_rn(1.f, w);
            vx = __fadd_rn(__fmul_rn(px, om), __fmul_rn(cvx, w));
            vy = __fadd_rn(__fmul_rn(py, om), __fmul_rn(cvy, w));
            vz = __fadd_rn(__fmul_rn(pz, om), __fmul_rn(cvz, w));
            my = fminf(ly, ry);
        }
        s_vy[wib][lane] = vy;
        s_my[wib][lane] = my;
        __syncwarp();
        if (lane == 0) {
            const int n = min(32, len - t0);
            for (int i = 0; i < n; ++i) {      // net.py:148-153, float64 scalars on float32 values
                const double foot = cur + (double)s_my[wib][i];
                float v = s_vy[wib][i];
                if (foot + (double)v <= kFloorY) {
                    v = (float)(kFloorY - foot);
                    s_vy[wib][i] = v;
                }
                cur += (double)v;
            }
        }
        __syncwarp();
        vy = s_vy[wib][lane];
        cur = __shfl_sync(0xffffffffu, cur, 0);
        // inclusive prefix sums in float64 (net.py:154 sums v[:i+1] per frame)
        double sx = valid ? (double)vx : 0.0, sy = valid ? (double)vy : 0.0, sz = valid ? (double)vz : 0.0;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const double ux = __shfl_up_sync(0xffffffffu, sx, o);
            const double uy = __shfl_up_sync(0xffffffffu, sy, o);
            const double uz = __shfl_up_sync(0xffffffffu, sz, o);
            if (lane >= o) { sx += ux; sy += uy; sz += uz; }
        }
        sx += cx; sy += cy; sz += cz;
        if (valid) {
            out[(size_t)t * 3 + 0] = (float)sx;
            out[(size_t)t * 3 + 1] = (float)sy;
            out[(size_t)t * 3 + 2] = (float)sz;
        }
        cx = __shfl_sync(0xffffffffu, sx, 31);
        cy = __shfl_sync(0xffffffffu, sy, 31);
        cz = __shfl_sync(0xffffffffu, sz, 31);
        __syncwarp();
    }
    for (int i = len * 3 + lane; i < T * 3; i += 32) out[i] = 0.f;
}

// K7: one warp per stream.
__global__ void __launch_bounds__(128) online_update_kernel(mp_online_state_t* __restrict__ state,
                                                            const float* __restrict__ pose, const float* __restrict__ joints,
                                                            const float* __restrict__ vel, const float* __restrict__ contact,
                                                            int S, int W, int frame, float* __restrict__ pose_out,
                                                            float* __restrict__ root_out, float* __restrict__ contact_out) {
    const int lane = threadIdx.x & 31;
    const int s = blockIdx.x * 4 + (threadIdx.x >> 5);
    if (s >= S) return;
    const size_t f = (size_t)s * W + frame;
    for (int i = lane; i < 216; i += 32) pose_out[(size_t)s * 216 + i] = pose[f * 216 + i];
    if (lane == 0) {
        mp_online_state_t st = state[s];
        const float* jt = joints + f * 72;
        const float lf[3] = {jt[30], jt[31], jt[32]}, rf[3] = {jt[33], jt[34], jt[35]};
        const float c0 = contact[f * 2], c1 = contact[f * 2 + 1];
        float cv[3];
        const bool left = c0 > c1;                       // strict: ties pick the right foot (net.py:189)
#pragma unroll
        for (int i = 0; i < 3; ++i) {
            const float g = (i == 1) ? kGravityVel : 0.f;
            cv[i] = __fadd_rn(__fsub_rn(left ? st.last_lfoot[i] : st.last_rfoot[i], left ? lf[i] : rf[i]), g);
        }
        const float w = prob_to_weight(fmaxf(c0, c1));    // NB: raw logit, no sigmoid (net.py:197)
        const float om = __fsub_rn(1.f, w);
        float v[3];
#pragma unroll
        for (int i = 0; i < 3; ++i) v[i] = __fadd_rn(__fmul_rn(vel[f * 72 + i] / kVelDiv, om), __fmul_rn(cv[i], w));
        const double foot = st.current_root_y + (double)fminf(lf[1], rf[1]);
        if (foot + (double)v[1] <= kFloorY) v[1] = (float)(kFloorY - foot);
        st.current_root_y += (double)v[1];
#pragma unroll
        for (int i = 0; i < 3; ++i) {
            st.last_lfoot[i] = lf[i];
            st.last_rfoot[i] = rf[i];
            st.last_root[i] = __fadd_rn(st.last_root[i], v[i]);
            root_out[(size_t)s * 3 + i] = st.last_root[i];
        }
        contact_out[(size_t)s * 2] = c0;
        contact_out[(size_t)s * 2 + 1] = c1;
        state[s] = st;
    }
}

__global__ void online_reset_kernel(mp_online_state_t* state, int S, int full) {
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= S) return;
    mp_online_state_t st = state[s];
    if (full) {
        for (int i = 0; i < 3; ++i) {
            st.last_lfoot[i] = c_feet[i];
            st.last_rfoot[i] = c_feet[3 + i];
        }
        st.pad_[0] = st.pad_[1] = st.pad_[2] = 0.f;
        st.pad2_ = 0.0;
    }
    st.last_root[0] = st.last_root[1] = st.last_root[2] = 0.f;
    st.current_root_y = 0.0;
    state[s] = st;
}

// sliding IMU window: out[s][w] = in[s][w+1] (w < W-1), out[s][W-1] = frame[s]; cold start
// replicates the frame over the whole window (net.py:175)
__global__ void online_push_kernel(const float* __restrict__ win_in, float* __restrict__ win_out,
                                   const float* __restrict__ frame, int S, int W, int cold) {
    const long long n = (long long)S * W * 60;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const int c = i % 60;
        const int w = (i / 60) % W;
        const int s = i / (60LL * W);
        win_out[i] = (cold || w == W - 1) ? frame[s * 60 + c] : win_in[i + 60];
    }
}

}  // namespace

int launch_reduced_global_to_full(const float* r6d, int64_t n, float* pose, cudaStream_t stream) {
    if (n <= 0) return MP_OK;
    MP_REQUIRE(r6d && pose, "pose: null pointer");
    MP_REQUIRE(((uintptr_t)r6d & 7) == 0, "pose: r6d must be 8-byte aligned");
    ProfileScope prof("k5_pose", 4.0 * (96 + 216) * (double)n, stream);
    const long long blocks = (n + 7) / 8;
    const int grid = (int)(blocks < 148LL * 16 ? blocks : 148LL * 16);
    reduced_global_to_full_kernel<<<grid, 256, 0, stream>>>(r6d, n, pose);
    MP_CUDA_TRY(cudaGetLastError());
    count_launch();
    return MP_OK;
}

// Compact form of the full local pose for transfers: the first two COLUMNS of the 16 non-ignored ("reduced") joints' local
// rotations, [n, 16, 6] = 384 B per frame instead of 864 B.  The ignored joints are the identity by construction (net.py:98) and a
// rotation's third column is the cross product of the first two, so the consumer rebuilds the [24, 3, 3] matrices up to their
// own orthonormality (model_utils.local6d_to_pose).  One thread per (frame, reduced joint).
__global__ void pose_local6d_kernel(const float* __restrict__ pose, long long n, float* __restrict__ out) {
    constexpr int kReduced[16] = MP_REDUCED_INIT;
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n * 16) return;
    const long long f = i >> 4;
    const int slot = (int)(i & 15);
    int joint = 0;
#pragma unroll
    for (int k = 0; k < 16; ++k)
        if (k == slot) joint = kReduced[k];
    const float* R = pose + (f * 24 + joint) * 9;          // row major: R[r][c] = R[3 r + c]
    float* o = out + i * 6;
    o[0] = R[0]; o[1] = R[3]; o[2] = R[6];                  // column 0
    o[3] = R[1]; o[4] = R[4]; o[5] = R[7];                  // column 1
}

int launch_pose_local6d(const float* pose, int64_t n, float* out, cudaStream_t stream) {
    if (n <= 0) return MP_OK;
    MP_REQUIRE(pose && out, "pose_local6d: null pointer");
    ProfileScope prof("pose_local6d", 4.0 * (216 + 96) * (double)n, stream);
    const long long threads = (long long)n * 16;
    pose_local6d_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, stream>>>(pose, n, out);
    MP_CUDA_TRY(cudaGetLastError());
    count_launch();
    return MP_OK;
}

int launch_tran_offline(const float* joints, const float* vel, const float* contact, const int32_t* lengths,
                        int B, int T, float* tran, cudaStream_t stream) {
    if (B <= 0 || T <= 0) return MP_OK;
    MP_REQUIRE(joints && vel && contact && tran, "tran: null pointer");
    ProfileScope prof("k6_tran", 4.0 * (6 + 3 + 2 + 3) * (double)B * T, stream);
    tran_offline_kernel<<<(B + 3) / 4, 128, 0, stream>>>(joints, vel, contact, lengths, B, T, tran);
    MP_CUDA_TRY(cudaGetLastError());
    count_launch();
    return MP_OK;
}

int launch_online_update(mp_online_state_t* st, const float* pose, const float* joints, const float* vel,
                         const float* contact, int S, int W, int frame, float* pose_out, float* root_out,
                         float* contact_out, cudaStream_t stream) {
    MP_REQUIRE(st && pose && joints && vel && contact && pose_out && root_out && contact_out, "online: null pointer");
    MP_REQUIRE(S > 0 && W > 0 && frame >= 0 && frame < W, "online: bad window (S=%d W=%d frame=%d)", S, W, frame);
    online_update_kernel<<<(S + 3) / 4, 128, 0, stream>>>(st, pose, joints, vel, contact, S, W, frame, pose_out, root_out,
                                                          contact_out);
    MP_CUDA_TRY(cudaGetLastError());
    count_launch();
    return MP_OK;
}

int launch_online_reset(mp_online_state_t* st, int S, int full, cudaStream_t stream) {
    MP_REQUIRE(st && S > 0, "online reset: bad arguments");
    online_reset_kernel<<<(S + 127) / 128, 128, 0, stream>>>(st, S, full);
    MP_CUDA_TRY(cudaGetLastError());
    count_launch();
    return MP_OK;
}

int launch_online_push(const float* win_in, float* win_out, const float* frame, int S, int W, int cold,
                       cudaStream_t stream) {
    MP_REQUIRE(win_out && frame && (cold || win_in) && S > 0 && W > 0, "online push: bad arguments");
    const long long n = (long long)S * W * 60;
    const int grid = (int)((n + 255) / 256 < 1184 ? (n + 255) / 256 : 1184);
    online_push_kernel<<<grid, 256, 0, stream>>>(win_in, win_out, frame, S, W, cold);
    MP_CUDA_TRY(cudaGetLastError());
    count_launch();
    return MP_OK;
}

}  // namespace mp
