// K8: kinematic-physics optimizer behind the reference's PHYSICS hook -- one warp per skeleton.
//
// Reference interface (relative to /root/reference/mobileposer): models/net.py:66-69 (`PhysicsOptimizer(debug=False)`,
// `reset_states()`), net.py:157-169 and 211-217 (`optimize_frame(pose, jvel, contact, acc) -> (pose, tran)`).  The module
// behind that hook (`dynamics`) is NOT in the reference tree and its dependency rbdl is neither vendored nor pinned
// (SURVEY.md F2): PARITY UNPINNED.  The algorithm is defined by this repository (float64 statement in
// oracle/physics_port.py); only its SMPL forward kinematics is pinned to the reference (articulate/model.py:208-232).
//
// Per frame (sequential in time, state = root position p + previous world joint positions q_prev):
//   FK of the network pose -> G_j, P_j;   e_j = q_prev_j + jvel_j/fps - (p + P_j);   ec_f = q_prev_f - (p + P_f)
//   min  w_vel sum_j |J_j dth + d - e_j|^2 + sum_f wc_f |J_f dth + d - ec_f|^2 + sum_i (damping H_ii + damping_abs) dth_i^2
//        (45 rotational + 3 translational unknowns; H = J^T W J: Marquardt's diagonal scaling)
//   R_k <- R_k exp([dth_k]x);  floor clamp on d_y;  p += d;  q_prev <- p + FK(new pose)
//
// Mapping: lane = joint for FK / residuals / the rotation update (parents by warp shuffle, level by level), the
// normal equations are assembled from SUBTREE MOMENTS instead of an explicit Jacobian -- column (k, a) of joint j is
// G_k[:,a] x (P_j - P_k), so every 3x3 block of J^T W J between an ancestor k1 and a descendant k2 is a function of
// G_k1, G_k2, P_k2 - P_k1 and the weighted moments (sum c r, sum c r r^T) of k2's subtree -- and solved by an envelope
// Cholesky in shared memory: unknowns are ordered leaves-first (elimination tree = kinematic tree) so the factor has
// no fill outside the ancestor blocks; the right-hand side rides along as row 48.  No cuBLAS / cuSolver.
// HBM traffic per frame and skeleton: 864 B pose in + 288 B velocity + 8 B contact, 864 B pose + 12 B tran out.
#include "mp_common.cuh"

#include <cstdlib>
#include "mp_constants.cuh"

#include <mutex>

namespace mp {

namespace {

constexpr int NJ = 24, NOPT = 15, NX = 48, NROW = 49, NBLK = 16;
// 49 rows (row 48 = right-hand side); every block of 3 unknowns owns 4 columns (the 4th stays zero) so that a block is one
// 16-byte read of a row; row stride 68 floats = 17 x 16 B keeps the 8 lanes of a quarter warp on distinct bank groups
constexpr int LDH = 68;
constexpr int MAX_PAIRS = 64;

struct PhysTables {
    int parent[NJ];
    int depth[NJ];
    unsigned desc[NJ];          // bit j set: j is k or a descendant of k
    float bone[NJ][3];          // J[j] - J[parent[j]] of the zero pose (articulate/model.py:77-92, shape=None)
    int ord[NOPT];              // elimination order (leaves first)
    int slot[NJ];               // joint -> slot in `ord` or -1
    int npair;
    unsigned char pair_a[MAX_PAIRS], pair_d[MAX_PAIRS];   // (ancestor-or-self slot, deeper slot)
    int env[NROW];              // first column of each row's envelope
    int first_blk[NBLK];        // first block of a block row's envelope
    unsigned char rowmap[NBLK][32];   // rows a lane works on in block column K: own 3, ancestors', translation, rhs; 255 = idle
};

__constant__ PhysTables c_tab;

// zero-pose joints and the tree: mp_constants.cuh (tests/test_constants.py holds them to mobileposer_b200/config.py)
const float kJZero[NJ][3] = MP_SMPL_J_ZERO_INIT;
const int kParent[NJ] = MP_SMPL_PARENT_INIT;
// joint_set.reduced without the root (config.py:134), deepest joints first, one kinematic chain after the other
const int kOrd[NOPT] = {15, 12, 18, 16, 13, 19, 17, 14, 9, 6, 3, 4, 1, 5, 2};

PhysTables build_tables() {
    PhysTables t = {};
    for (int j = 0; j < NJ; ++j) {
        t.parent[j] = kParent[j];
        t.depth[j] = kParent[j] < 0 ? 0 : t.depth[kParent[j]] + 1;
        for (int a = 0; a < 3; ++a) t.bone[j][a] = kJZero[j][a] - (kParent[j] < 0 ? 0.f : kJZero[kParent[j]][a]);
        t.slot[j] = -1;
    }
    for (int j = 0; j < NJ; ++j)
        for (int k = j; k >= 0; k = kParent[k]) t.desc[k] |= 1u << j;
    for (int i = 0; i < NOPT; ++i) {
        t.ord[i] = kOrd[i];
        t.slot[kOrd[i]] = i;
    }
    // pairs (ancestor-or-self, deeper); the envelope of a row starts at the first slot that is a descendant of its joint
    int first_desc[NOPT];
    for (int i = 0; i < NOPT; ++i) first_desc[i] = i;
    for (int d = 0; d < NOPT; ++d)
        for (int a = 0; a < NOPT; ++a)
            if ((t.desc[kOrd[a]] >> kOrd[d]) & 1u) {
                t.pair_a[t.npair] = (unsigned char)a;
                t.pair_d[t.npair] = (unsigned char)d;
                ++t.npair;
                if (d < first_desc[a]) first_desc[a] = d;
            }
    int lo = NX;
    for (int i = 0; i < NOPT; ++i) {
        for (int a = 0; a < 3; ++a) t.env[3 * i + a] = 3 * first_desc[i];
        if (3 * first_desc[i] < lo) lo = 3 * first_desc[i];
    }
    for (int r = 45; r < NROW; ++r) t.env[r] = 0;     // translation rows and the right-hand side couple with everything
    for (int K = 0; K < NBLK; ++K) {
        t.first_blk[K] = K == 15 ? 0 : first_desc[K];
        int n = 0;
        for (int l = 0; l < 32; ++l) t.rowmap[K][l] = 255;
        for (int r = 0; r < 3; ++r) t.rowmap[K][n++] = (unsigned char)(3 * K + r);
        if (K < 15) {
            for (int a = K + 1; a < NOPT; ++a)
                if ((t.desc[kOrd[a]] >> kOrd[K]) & 1u)
                    for (int r = 0; r < 3; ++r) t.rowmap[K][n++] = (unsigned char)(3 * a + r);
            for (int r = 45; r < 48; ++r) t.rowmap[K][n++] = (unsigned char)r;
        }
        t.rowmap[K][n++] = 48;
    }
    return t;
}

int upload_tables() {
    static std::mutex mu;
    static bool done[64] = {};
    int dev = 0;
    MP_CUDA_TRY(cudaGetDevice(&dev));
    std::lock_guard<std::mutex> lock(mu);
    if (dev < 64 && done[dev]) return MP_OK;
    const PhysTables t = build_tables();
    MP_REQUIRE(t.npair <= MAX_PAIRS, "physics: pair table overflow (%d)", t.npair);
    MP_CUDA_TRY(cudaMemcpyToSymbol(c_tab, &t, sizeof(t)));
    if (dev < 64) done[dev] = true;
    return MP_OK;
}

struct PhysParams {
    const float* pose;       // [B, T, 24, 9] local rotations (K5 output)
    const float* vel;        // [B, T, 72]    raw velocity-head output
    const float* contact;    // [B, T, 2]     logits
    const int32_t* lengths;  // [B] or null
    float* state;            // [B, 80]: p[3], started, q_prev[72], pad[4]
    float* pose_out;         // [B, T, 24, 9] (may alias pose)
    float* tran_out;         // [B, T, 3] or null
    float* dbg;              // [B, 49*49 + 48] normal equations (+ rhs row) and solution of frame dbg_frame, or null
    int B, T, dbg_frame;
    float jvel_dt;           // vel_scale / fps: raw velocity -> displacement per frame
    float w_vel, w_contact, damping, damping_abs, floor_y;
};

// lane = joint: global rotation and root-relative position by pointer jumping.  After round r a lane holds the
// transform from its 2^r-th ancestor's frame (the world once the pointer has run past the root) down to itself:
// 4 rounds cover the 9 links of the deepest chain, against 8 rounds of a level-by-level walk.  `jump` packs the
// lane's 1st, 2nd, 4th and 8th ancestor (255 = past the root), one byte each.
__device__ __forceinline__ void warp_fk(const float (&R)[9], bool isj, unsigned jump, const float (&bone)[3], float (&G)[9], float (&P)[3]) {
#pragma unroll
    for (int i = 0; i < 9; ++i) G[i] = R[i];
#pragma unroll
    for (int i = 0; i < 3; ++i) P[i] = bone[i];
#pragma unroll
    for (int r = 0; r < 4; ++r) {
        const unsigned a = (jump >> (8 * r)) & 255u;
        const int src = a == 255u ? 0 : (int)a;
        float Ga[9], Pa[3];
#pragma unroll
        for (int i = 0; i < 9; ++i) Ga[i] = __shfl_sync(0xffffffffu, G[i], src);
#pragma unroll
        for (int i = 0; i < 3; ++i) Pa[i] = __shfl_sync(0xffffffffu, P[i], src);
        if (isj && a != 255u) {
            float Gn[9], Pn[3];
#pragma unroll
            for (int rr = 0; rr < 3; ++rr) {
#pragma unroll
                for (int c = 0; c < 3; ++c) Gn[rr * 3 + c] = fmaf(Ga[rr * 3 + 2], G[6 + c], fmaf(Ga[rr * 3 + 1], G[3 + c], Ga[rr * 3] * G[c]));
                Pn[rr] = Pa[rr] + fmaf(Ga[rr * 3 + 2], P[2], fmaf(Ga[rr * 3 + 1], P[1], Ga[rr * 3] * P[0]));
            }
#pragma unroll
            for (int i = 0; i < 9; ++i) G[i] = Gn[i];
#pragma unroll
            for (int i = 0; i < 3; ++i) P[i] = Pn[i];
        }
    }
}

__device__ __forceinline__ float sincf_stable(float x) { return fabsf(x) < 1e-3f ? 1.f - x * x * (1.f / 6.f) : sinf(x) / x; }

__device__ __forceinline__ float phys_prob_to_weight(float logit) {
    const float p = 1.f / (1.f + expf(-logit));
    return (fminf(fmaxf(p, 0.5f), 0.9f) - 0.5f) / 0.4f;       // net.py:90-91
}

// Per-skeleton scratch (one warp each).  A CTA packs `blockDim.x / 32` skeletons: with one warp per CTA the batch's 256 tiny CTAs spread
// over every SM of the chip for the kernel's 3 ms, and their registers / shared memory kept the cluster recurrences and the
// persistent projection CTAs of the other batches in the pipeline off those SMs (1 ms per cfg3 step); packed, the optimizer
// occupies a few SMs and leaves the rest whole.
struct PhysScratch {
    float G[NJ * 9];
    float P[NJ * 3];
    float C[NJ];
    float S[NJ * 3];
    float S1[NJ * 3];
    float S2[NJ * 6];
    float T1[NJ * 3];
    float X[NX];
    uint32_t Item[MAX_PAIRS];     // rotation-rotation blocks: joint a | joint d << 5 | slot a << 10 | slot d << 14
    int Ord[NOPT + 1];
    unsigned char Row[NBLK * 32];
    float H[NROW * LDH] __attribute__((aligned(16)));
};

__global__ void __launch_bounds__(512) physics_optimize_kernel(const PhysParams p) {
    extern __shared__ __align__(16) unsigned char phys_smem[];
    PhysScratch& scr = reinterpret_cast<PhysScratch*>(phys_smem)[threadIdx.x >> 5];
    float* const sG = scr.G;
    float* const sP = scr.P;
    float* const sC = scr.C;
    float* const sS = scr.S;
    float* const sS1 = scr.S1;
    float* const sS2 = scr.S2;
    float* const sT1 = scr.T1;
    float* const sH = scr.H;
    float* const sX = scr.X;
    uint32_t* const sItem = scr.Item;
    int* const sOrd = scr.Ord;
    unsigned char* const sRow = scr.Row;

    const int lane = threadIdx.x & 31;
    const int b = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (b >= p.B) return;
    const bool isj = lane < NJ;
    const int j = isj ? lane : 0;
    unsigned jump = 0;
    {
        for (int r = 0, step = 1; r < 4; ++r, step *= 2) {       // 1st, 2nd, 4th, 8th ancestor of the lane's joint
            int q = j;
            for (int sidx = 0; sidx < step && q >= 0; ++sidx) q = c_tab.parent[q];
            jump |= (unsigned)(q < 0 ? 255 : q) << (8 * r);
        }
    }
    const float bone[3] = {c_tab.bone[j][0], c_tab.bone[j][1], c_tab.bone[j][2]};
    const int slot = isj ? c_tab.slot[j] : -1;
    const unsigned desc = c_tab.desc[j];
    const int len = p.lengths ? min(max(p.lengths[b], 0), p.T) : p.T;

    float* st = p.state + (size_t)b * 80;
    float px = st[0], py = st[1], pz = st[2];
    bool started = st[3] != 0.f;
    float q[3] = {0.f, 0.f, 0.f};
    if (isj) { q[0] = st[4 + j * 3]; q[1] = st[5 + j * 3]; q[2] = st[6 + j * 3]; }

    for (int i = lane; i < NROW * LDH; i += 32) sH[i] = 0.f;
    // pair descriptors of the rotation-rotation blocks (the constant tables are indexed per lane: read them once)
    const int npair = c_tab.npair;
    for (int pr = lane; pr < npair; pr += 32) {
        const int ia = c_tab.pair_a[pr], id = c_tab.pair_d[pr];
        sItem[pr] = (uint32_t)c_tab.ord[ia] | ((uint32_t)c_tab.ord[id] << 5) | ((uint32_t)ia << 10) | ((uint32_t)id << 14);
    }
    if (lane < NOPT) sOrd[lane] = c_tab.ord[lane];
    for (int i = lane; i < NBLK * 32; i += 32) sRow[i] = c_tab.rowmap[i / 32][i % 32];
    __syncwarp();

    const float* pose_b = p.pose + (size_t)b * p.T * 216;
    const float* vel_b = p.vel + (size_t)b * p.T * 72;
    const float* con_b = p.contact + (size_t)b * p.T * 2;
    float* out_b = p.pose_out + (size_t)b * p.T * 216;

    // frame 0 inputs; inside the loop the loads of frame t+1 are issued before the work of frame t
    float Rn[9], vn[3], cn0 = 0.f, cn1 = 0.f;
#pragma unroll
    for (int i = 0; i < 9; ++i) Rn[i] = (isj && len > 0) ? __ldg(pose_b + j * 9 + i) : (i % 4 == 0 ? 1.f : 0.f);
#pragma unroll
    for (int i = 0; i < 3; ++i) vn[i] = (isj && len > 0) ? __ldg(vel_b + j * 3 + i) : 0.f;
    if (len > 0) { cn0 = __ldg(con_b); cn1 = __ldg(con_b + 1); }

    for (int t = 0; t < len; ++t) {
        float R[9], jv[3];
#pragma unroll
        for (int i = 0; i < 9; ++i) R[i] = Rn[i];
#pragma unroll
        for (int i = 0; i < 3; ++i) jv[i] = vn[i] * p.jvel_dt;
        const float c0 = cn0, c1 = cn1;
        if (t + 1 < len) {
            if (isj) {
#pragma unroll
                for (int i = 0; i < 9; ++i) Rn[i] = __ldg(pose_b + (size_t)(t + 1) * 216 + j * 9 + i);
#pragma unroll
                for (int i = 0; i < 3; ++i) vn[i] = __ldg(vel_b + (size_t)(t + 1) * 72 + j * 3 + i);
            }
            cn0 = __ldg(con_b + (size_t)(t + 1) * 2);
            cn1 = __ldg(con_b + (size_t)(t + 1) * 2 + 1);
        }

        float G[9], P[3];
        warp_fk(R, isj, jump, bone, G, P);

        float Rout[9];
#pragma unroll
        for (int i = 0; i < 9; ++i) Rout[i] = R[i];
        float dx = 0.f, dy = 0.f, dz = 0.f;

        if (started) {
            // ---- residuals and weights (lane = joint) ----------------------------------------------------------
            const float wcl = p.w_contact * phys_prob_to_weight(c0), wcr = p.w_contact * phys_prob_to_weight(c1);
            const float wc = (j == 10) ? wcl : ((j == 11) ? wcr : 0.f);
            const float wx = px + P[0], wy = py + P[1], wz = pz + P[2];
            const float cj = isj ? p.w_vel + wc : 0.f;
            float s[3];
            s[0] = isj ? p.w_vel * (q[0] + jv[0] - wx) + wc * (q[0] - wx) : 0.f;
            s[1] = isj ? p.w_vel * (q[1] + jv[1] - wy) + wc * (q[1] - wy) : 0.f;
            s[2] = isj ? p.w_vel * (q[2] + jv[2] - wz) + wc * (q[2] - wz) : 0.f;
            if (isj) {
#pragma unroll
                for (int i = 0; i < 9; ++i) sG[j * 9 + i] = G[i];
#pragma unroll
                for (int i = 0; i < 3; ++i) { sP[j * 3 + i] = P[i]; sS[j * 3 + i] = s[i]; }
                sC[j] = cj;
            }
            float c_tot = cj, s_tot[3] = {s[0], s[1], s[2]};
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                c_tot += __shfl_xor_sync(0xffffffffu, c_tot, o);
#pragma unroll
                for (int i = 0; i < 3; ++i) s_tot[i] += __shfl_xor_sync(0xffffffffu, s_tot[i], o);
            }
            __syncwarp();

            // ---- subtree moments about the joint's own position (lanes of the optimised joints) ----------------
            if (slot >= 0) {
                float S1[3] = {0.f, 0.f, 0.f}, S2[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f}, T1[3] = {0.f, 0.f, 0.f};
#pragma unroll 2
                for (unsigned rest = desc & ~(1u << j); rest; rest &= rest - 1) {
                    const int d = __ffs(rest) - 1;
                    const float rx = sP[d * 3] - P[0], ry = sP[d * 3 + 1] - P[1], rz = sP[d * 3 + 2] - P[2];
                    const float cd = sC[d], sx = sS[d * 3], sy = sS[d * 3 + 1], sz = sS[d * 3 + 2];
                    S1[0] = fmaf(cd, rx, S1[0]); S1[1] = fmaf(cd, ry, S1[1]); S1[2] = fmaf(cd, rz, S1[2]);
                    S2[0] = fmaf(cd * rx, rx, S2[0]); S2[1] = fmaf(cd * rx, ry, S2[1]); S2[2] = fmaf(cd * rx, rz, S2[2]);
                    S2[3] = fmaf(cd * ry, ry, S2[3]); S2[4] = fmaf(cd * ry, rz, S2[4]); S2[5] = fmaf(cd * rz, rz, S2[5]);
                    T1[0] += ry * sz - rz * sy; T1[1] += rz * sx - rx * sz; T1[2] += rx * sy - ry * sx;
                }
#pragma unroll
                for (int i = 0; i < 3; ++i) { sS1[j * 3 + i] = S1[i]; sT1[j * 3 + i] = T1[i]; }
#pragma unroll
                for (int i = 0; i < 6; ++i) sS2[j * 6 + i] = S2[i];
            }
            __syncwarp();

            // ---- normal equations: rotation-rotation blocks of ancestor-related joint pairs --------------------
            // one lane per (ancestor-or-self a, descendant d) pair: entry (i, k) of the 3 x 3 block is
            //   g2_k . [ tr g1_i - S2 g1_i - e (S1 . g1_i) ],   g1_i / g2_k = columns of G_a / G_d, e = P_d - P_a,
            //   S1, S2 = moments of d's subtree about P_d, tr = trace(S2) + e . S1
            for (int pr = lane; pr < npair; pr += 32) {
                const uint32_t ds = sItem[pr];
                const int ka = ds & 31, kd = (ds >> 5) & 31, ia = (ds >> 10) & 15, id = (ds >> 14) & 15;
                float Ga[9], Gd[9];
#pragma unroll
                for (int i = 0; i < 9; ++i) { Ga[i] = sG[ka * 9 + i]; Gd[i] = sG[kd * 9 + i]; }
                const float ex = sP[kd * 3] - sP[ka * 3], ey = sP[kd * 3 + 1] - sP[ka * 3 + 1], ez = sP[kd * 3 + 2] - sP[ka * 3 + 2];
                const float m1x = sS1[kd * 3], m1y = sS1[kd * 3 + 1], m1z = sS1[kd * 3 + 2];
                const float* M = sS2 + kd * 6;     // xx xy xz yy yz zz
                const float mxx = M[0], mxy = M[1], mxz = M[2], myy = M[3], myz = M[4], mzz = M[5];
                const float tr = mxx + myy + mzz + (ex * m1x + ey * m1y + ez * m1z);
                float* dst = sH + (3 * ia) * LDH + 4 * id;
#pragma unroll
                for (int i = 0; i < 3; ++i) {
                    const float gx = Ga[i], gy = Ga[3 + i], gz = Ga[6 + i];
                    const float sg = m1x * gx + m1y * gy + m1z * gz;
                    const float nx = tr * gx - (mxx * gx + mxy * gy + mxz * gz) - ex * sg;
                    const float ny = tr * gy - (mxy * gx + myy * gy + myz * gz) - ey * sg;
                    const float nz = tr * gz - (mxz * gx + myz * gy + mzz * gz) - ez * sg;
#pragma unroll
                    for (int k = 0; k < 3; ++k) {
                        float v = Gd[k] * nx + Gd[3 + k] * ny + Gd[6 + k] * nz;
                        if (ia == id && i == k) v = fmaf(v, p.damping, v) + p.damping_abs;      // Marquardt scaling
                        // lower triangle only: descendants come first in the elimination order, so (row, col) =
                        // (ancestor, descendant); inside a diagonal block keep i >= k
                        if (ia != id || i >= k) dst[i * LDH + k] = v;
                    }
                }
            }
            // rotation-translation blocks, translation block, right-hand side (row 48)
            for (int it = lane; it < NOPT * 3; it += 32) {
                const int i = it / 3, a = it % 3, k = sOrd[i];
                const float gx = sG[k * 9 + a], gy = sG[k * 9 + 3 + a], gz = sG[k * 9 + 6 + a];
                const float mx = sS1[k * 3], my = sS1[k * 3 + 1], mz = sS1[k * 3 + 2];
                const int col = 4 * i + a;
                sH[45 * LDH + col] = gy * mz - gz * my;
                sH[46 * LDH + col] = gz * mx - gx * mz;
                sH[47 * LDH + col] = gx * my - gy * mx;
                sH[48 * LDH + col] = gx * sT1[k * 3] + gy * sT1[k * 3 + 1] + gz * sT1[k * 3 + 2];
            }
            if (lane < 3) {
                sH[(45 + lane) * LDH + 60] = lane == 0 ? c_tot : 0.f;
                sH[(45 + lane) * LDH + 61] = lane == 1 ? c_tot : 0.f;
                sH[(45 + lane) * LDH + 62] = lane == 2 ? c_tot : 0.f;
                sH[48 * LDH + 60 + lane] = s_tot[lane];
            }
            __syncwarp();
            if (p.dbg && t == p.dbg_frame) {
                float* dbg = p.dbg + (size_t)b * (NROW * NROW + NX);
                for (int i = lane; i < NROW * NROW; i += 32) {      // compact 49 x 49 view of the padded storage
                    const int c = i % NROW;
                    dbg[i] = c < NX ? sH[(i / NROW) * LDH + 4 * (c / 3) + c % 3] : 0.f;
                }
            }

            // ---- envelope Cholesky, left looking, one joint (block of 3 unknowns) per round.  Lane l works on row
            //      rowmap[K][l] of block column K: lanes 0..2 hold the diagonal block, then the rows of the joint's ancestors,
            //      the translation rows and the right-hand side (row 48) -- never more than 22 rows.  A block of a row is
            //      one 16-byte read; the dot products run over the blocks of the joint's descendants only (the envelope).
#pragma unroll 1
            for (int K = 0; K < NBLK; ++K) {
                const int c0 = 3 * K;
                const int i1 = sRow[K * 32 + lane];
                const bool h1 = i1 != 255;
                const float* rk = sH + c0 * LDH;
                const float* r1 = sH + (h1 ? i1 : c0) * LDH;
                // idle lanes must not touch the diagonal block lane 0 rewrites below (warp-synchronous, but racecheck is right to ask)
                const float4 ini = h1 ? *reinterpret_cast<const float4*>(r1 + 4 * K) : make_float4(1.f, 0.f, 0.f, 0.f);
                float a0 = ini.x, a1 = ini.y, a2 = ini.z;
#pragma unroll 2
                for (int m = 4 * c_tab.first_blk[K]; m < 4 * K; m += 4) {
                    const float4 u = *reinterpret_cast<const float4*>(r1 + m);
                    const float4 k0 = *reinterpret_cast<const float4*>(rk + m), k1 = *reinterpret_cast<const float4*>(rk + LDH + m),
                                 k2 = *reinterpret_cast<const float4*>(rk + 2 * LDH + m);
                    a0 = fmaf(-u.x, k0.x, a0); a1 = fmaf(-u.x, k1.x, a1); a2 = fmaf(-u.x, k2.x, a2);
                    a0 = fmaf(-u.y, k0.y, a0); a1 = fmaf(-u.y, k1.y, a1); a2 = fmaf(-u.y, k2.y, a2);
                    a0 = fmaf(-u.z, k0.z, a0); a1 = fmaf(-u.z, k1.z, a1); a2 = fmaf(-u.z, k2.z, a2);
                }
                // 3 x 3 diagonal block (rows of lanes 0, 1, 2), factored redundantly by every lane
                const float d00 = __shfl_sync(0xffffffffu, a0, 0);
                const float d10 = __shfl_sync(0xffffffffu, a0, 1), d11 = __shfl_sync(0xffffffffu, a1, 1);
                const float d20 = __shfl_sync(0xffffffffu, a0, 2), d21 = __shfl_sync(0xffffffffu, a1, 2),
                            d22 = __shfl_sync(0xffffffffu, a2, 2);
                const float v0 = rsqrtf(d00);
                const float l10 = d10 * v0, l20 = d20 * v0;
                const float v1 = rsqrtf(fmaf(-l10, l10, d11));
                const float l21 = fmaf(-l20, l10, d21) * v1;
                const float v2 = rsqrtf(fmaf(-l21, l21, fmaf(-l20, l20, d22)));
                float4 out;
                if (lane >= 3) {
                    out.x = a0 * v0;
                    out.y = fmaf(-out.x, l10, a1) * v1;
                    out.z = fmaf(-out.y, l21, fmaf(-out.x, l20, a2)) * v2;
                } else {              // the diagonal keeps 1 / L_kk
                    out.x = lane == 0 ? v0 : (lane == 1 ? l10 : l20);
                    out.y = lane == 0 ? 0.f : (lane == 1 ? v1 : l21);
                    out.z = lane == 2 ? v2 : 0.f;
                }
                out.w = 0.f;
                if (h1) *reinterpret_cast<float4*>(sH + i1 * LDH + 4 * K) = out;
                __syncwarp();
            }
            // ---- back substitution L^T x = y (y = row 48), one joint per round; lane l owns x_l and x_{l+32} --------------
            const int col1 = 4 * (lane / 3) + lane % 3, col2 = 4 * ((lane + 32) / 3) + (lane + 32) % 3;
            float y1 = sH[48 * LDH + col1], y2 = (lane < NX - 32) ? sH[48 * LDH + col2] : 0.f;
#pragma unroll 1
            for (int K = NBLK - 1; K >= 0; --K) {
                const int c0 = 3 * K;
                const float* rk = sH + c0 * LDH;
                const float4 dg0 = *reinterpret_cast<const float4*>(rk + 4 * K), dg1 = *reinterpret_cast<const float4*>(rk + LDH + 4 * K),
                             dg2 = *reinterpret_cast<const float4*>(rk + 2 * LDH + 4 * K);
                const float v0 = dg0.x, l10 = dg1.x, v1 = dg1.y, l20 = dg2.x, l21 = dg2.y, v2 = dg2.z;
                const float u0 = rk[col1], u1 = rk[LDH + col1], u2 = rk[2 * LDH + col1];            // columns l < c0 of the 3 rows
                const int cw = lane < NX - 32 ? col2 : 0;
                const float w0 = rk[cw], w1 = rk[LDH + cw], w2 = rk[2 * LDH + cw];
                const float yk0 = __shfl_sync(0xffffffffu, c0 >= 32 ? y2 : y1, c0 & 31);
                const float yk1 = __shfl_sync(0xffffffffu, c0 + 1 >= 32 ? y2 : y1, (c0 + 1) & 31);
                const float yk2 = __shfl_sync(0xffffffffu, c0 + 2 >= 32 ? y2 : y1, (c0 + 2) & 31);
                const float x2 = yk2 * v2;
                const float x1 = fmaf(-l21, x2, yk1) * v1;
                const float x0 = fmaf(-l20, x2, fmaf(-l10, x1, yk0)) * v0;
#pragma unroll
                for (int c = 0; c < 3; ++c) {
                    const float xc = c == 0 ? x0 : (c == 1 ? x1 : x2);
                    if (lane == ((c0 + c) & 31)) { if (c0 + c >= 32) y2 = xc; else y1 = xc; }
                }
                if (lane < c0) y1 = fmaf(-u2, x2, fmaf(-u1, x1, fmaf(-u0, x0, y1)));
                if (lane + 32 < c0) y2 = fmaf(-w2, x2, fmaf(-w1, x1, fmaf(-w0, x0, y2)));
            }
            sX[lane] = y1;
            if (lane < NX - 32) sX[32 + lane] = y2;
            __syncwarp();
            if (p.dbg && t == p.dbg_frame) {
                float* dbg = p.dbg + (size_t)b * (NROW * NROW + NX) + NROW * NROW;
                for (int i = lane; i < NX; i += 32) dbg[i] = sX[i];
            }

            // ---- rotation update of the optimised joints (lane = joint) ----------------------------------------
            if (slot >= 0) {
                const float w0 = sX[3 * slot], w1 = sX[3 * slot + 1], w2 = sX[3 * slot + 2];
                const float th2 = w0 * w0 + w1 * w1 + w2 * w2, th = sqrtf(th2);
                const float A = sincf_stable(th), hb = sincf_stable(0.5f * th), Bc = 0.5f * hb * hb;
                // E = (1 - B th^2) I + A [w]x + B w w^T
                const float e0 = 1.f - Bc * th2;
                const float E[9] = {e0 + Bc * w0 * w0, Bc * w0 * w1 - A * w2, Bc * w0 * w2 + A * w1,
                                    Bc * w0 * w1 + A * w2, e0 + Bc * w1 * w1, Bc * w1 * w2 - A * w0,
                                    Bc * w0 * w2 - A * w1, Bc * w1 * w2 + A * w0, e0 + Bc * w2 * w2};
#pragma unroll
                for (int r = 0; r < 3; ++r)
#pragma unroll
                    for (int c = 0; c < 3; ++c) Rout[r * 3 + c] = fmaf(R[r * 3 + 2], E[6 + c], fmaf(R[r * 3 + 1], E[3 + c], R[r * 3] * E[c]));
            }
            dx = sX[45]; dy = sX[46]; dz = sX[47];
            __syncwarp();
            warp_fk(Rout, isj, jump, bone, G, P);
        }

        // ---- floor clamp (net.py:148-153's rule on the optimiser's own root), integration, state -----------------
        const float fl = __shfl_sync(0xffffffffu, P[1], 10), fr = __shfl_sync(0xffffffffu, P[1], 11);
        const float foot_y = fminf(fl, fr) + py + dy;
        if (foot_y < p.floor_y) dy += p.floor_y - foot_y;
        px += dx; py += dy; pz += dz;
        q[0] = px + P[0]; q[1] = py + P[1]; q[2] = pz + P[2];
        started = true;

        if (isj) {
            float* dst = out_b + (size_t)t * 216 + j * 9;
#pragma unroll
            for (int i = 0; i < 9; ++i) dst[i] = Rout[i];
        }
        if (p.tran_out && lane < 3) p.tran_out[((size_t)b * p.T + t) * 3 + lane] = lane == 0 ? px : (lane == 1 ? py : pz);
    }

    // frames past the sequence's length pass through (pose) / hold the last root position (tran)
    for (int t = len; t < p.T; ++t) {
        if (isj && p.pose_out != p.pose) {
#pragma unroll
            for (int i = 0; i < 9; ++i) out_b[(size_t)t * 216 + j * 9 + i] = pose_b[(size_t)t * 216 + j * 9 + i];
        }
        if (p.tran_out && lane < 3) p.tran_out[((size_t)b * p.T + t) * 3 + lane] = lane == 0 ? px : (lane == 1 ? py : pz);
    }

    if (lane == 0) { st[0] = px; st[1] = py; st[2] = pz; st[3] = started ? 1.f : 0.f; }
    if (isj) { st[4 + j * 3] = q[0]; st[5 + j * 3] = q[1]; st[6 + j * 3] = q[2]; }
}

__global__ void physics_fk_kernel(const float* __restrict__ pose, long long n, float* __restrict__ glb, float* __restrict__ pos) {
    const int lane = threadIdx.x & 31;
    const long long warps = ((long long)gridDim.x * blockDim.x) >> 5;
    const bool isj = lane < NJ;
    const int j = isj ? lane : 0;
    unsigned jump = 0;
    {
        for (int r = 0, step = 1; r < 4; ++r, step *= 2) {       // 1st, 2nd, 4th, 8th ancestor of the lane's joint
            int q = j;
            for (int sidx = 0; sidx < step && q >= 0; ++sidx) q = c_tab.parent[q];
            jump |= (unsigned)(q < 0 ? 255 : q) << (8 * r);
        }
    }
    const float bone[3] = {c_tab.bone[j][0], c_tab.bone[j][1], c_tab.bone[j][2]};
    for (long long f = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5; f < n; f += warps) {
        float R[9], G[9], P[3];
#pragma unroll
        for (int i = 0; i < 9; ++i) R[i] = isj ? __ldg(pose + f * 216 + j * 9 + i) : 0.f;
        warp_fk(R, isj, jump, bone, G, P);
        if (isj) {
#pragma unroll
            for (int i = 0; i < 9; ++i) glb[f * 216 + j * 9 + i] = G[i];
#pragma unroll
            for (int i = 0; i < 3; ++i) pos[f * 72 + j * 3 + i] = P[i];
        }
    }
}

// ---- N1 (metric side, evaluator.py:292-343): per-frame errors between two motions -------------------------------------
// One warp per frame, lane = joint: SMPL forward kinematics of both poses (same pointer-jumping FK as K8), world joint
// positions, root-aligned joint position error, local and global joint angle errors in degrees.  The frame-to-frame rows
// (jerk, translation drift) and the mean / std reductions are cheap slices of these outputs and stay in torch.
__device__ __forceinline__ float rot_angle_deg(const float (&A)[9], const float (&B)[9]) {
    // angle of A^T B (angular.py:86-99): d[a][b] = sum_k A[k][a] B[k][b]; trace = 1 + 2 cos, |skew| = 2 sin
    float tr = 0.f;
#pragma unroll
    for (int i = 0; i < 9; ++i) tr = fmaf(A[i], B[i], tr);
    const float d21 = A[2] * B[1] + A[5] * B[4] + A[8] * B[7], d12 = A[1] * B[2] + A[4] * B[5] + A[7] * B[8];
    const float d02 = A[0] * B[2] + A[3] * B[5] + A[6] * B[8], d20 = A[2] * B[0] + A[5] * B[3] + A[8] * B[6];
    const float d10 = A[1] * B[0] + A[4] * B[3] + A[7] * B[6], d01 = A[0] * B[1] + A[3] * B[4] + A[6] * B[7];
    const float sx = d21 - d12, sy = d02 - d20, sz = d10 - d01;
    return atan2f(sqrtf(sx * sx + sy * sy + sz * sz), tr - 1.0f) * 57.29577951308232f;
}

__global__ void __launch_bounds__(256) eval_frame_errors_kernel(const float* __restrict__ pose_p, const float* __restrict__ pose_t,
                                                                const float* __restrict__ tran_p, const float* __restrict__ tran_t,
                                                                long long n, float* __restrict__ joint_p, float* __restrict__ joint_t,
                                                                float* __restrict__ je, float* __restrict__ lae, float* __restrict__ gae) {
    const int lane = threadIdx.x & 31;
    const long long warps = ((long long)gridDim.x * blockDim.x) >> 5;
    const bool isj = lane < NJ;
    const int j = isj ? lane : 0;
    unsigned jump = 0;
    for (int r = 0, step = 1; r < 4; ++r, step *= 2) {
        int q = j;
        for (int sidx = 0; sidx < step && q >= 0; ++sidx) q = c_tab.parent[q];
        jump |= (unsigned)(q < 0 ? 255 : q) << (8 * r);
    }
    const float bone[3] = {c_tab.bone[j][0], c_tab.bone[j][1], c_tab.bone[j][2]};
    for (long long f = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5; f < n; f += warps) {
        float Rp[9], Rt[9], Gp[9], Gt[9], Pp[3], Pt[3];
#pragma unroll
        for (int i = 0; i < 9; ++i) {
            Rp[i] = isj ? __ldg(pose_p + f * 216 + j * 9 + i) : 0.f;
            Rt[i] = isj ? __ldg(pose_t + f * 216 + j * 9 + i) : 0.f;
        }
        warp_fk(Rp, isj, jump, bone, Gp, Pp);
        warp_fk(Rt, isj, jump, bone, Gt, Pt);
        float off[3];
#pragma unroll
        for (int i = 0; i < 3; ++i) {
            Pp[i] += tran_p ? __ldg(tran_p + f * 3 + i) : 0.f;
            Pt[i] += tran_t ? __ldg(tran_t + f * 3 + i) : 0.f;
            off[i] = __shfl_sync(0xffffffffu, Pt[i] - Pp[i], 0);          // align the roots (evaluator.py:322)
        }
        if (isj) {
            const float ex = Pp[0] + off[0] - Pt[0], ey = Pp[1] + off[1] - Pt[1], ez = Pp[2] + off[2] - Pt[2];
            je[f * NJ + j] = sqrtf(ex * ex + ey * ey + ez * ez);
            lae[f * NJ + j] = rot_angle_deg(Rp, Rt);
            gae[f * NJ + j] = rot_angle_deg(Gp, Gt);
#pragma unroll
            for (int i = 0; i < 3; ++i) {
                joint_p[f * 72 + j * 3 + i] = Pp[i];
                joint_t[f * 72 + j * 3 + i] = Pt[i];
            }
        }
    }
}


// ---- N1, mesh row (evaluator.py:319-323 with articulate/model.py:233-240): root-aligned vertex position error -----------
// The reference skins both motions to [n, V, 3] vertex sets (V = 6890: 165 KB per frame and motion) and subtracts them.  Linear
// blend skinning is linear in the joint transforms, so the DIFFERENCE of the two meshes is the skinning of the difference
// of the transforms:  v_p - v_t = sum_j w[v][j] * (A_p[j] - A_t[j]) * [v0; 1],  A[j] = [G_j | P_j - G_j J_j]  (model.py:233).
// With the default alignment joint (the root, whose position is the translation in both motions) the offset of
// evaluator.py:322 cancels the translations exactly, so neither a vertex set nor a translation is ever materialised:
// per frame chunk the CTA runs the forward kinematics of both motions (a warp per frame, lane = joint), leaves the 24
// difference transforms in shared memory, and every thread skins its two vertices (weights and rest position in registers)
// against them, accumulating sum and sum of squares of the error per vertex -- all the mean / std rows need.
constexpr int VE_FC = 32;        // frames per chunk
constexpr int VE_VPT = 2;        // vertices per thread
constexpr int VE_THREADS = 256;

__global__ void __launch_bounds__(VE_THREADS)
eval_vertex_errors_kernel(const float* __restrict__ pose_p, const float* __restrict__ pose_t, long long n,
                          const float* __restrict__ v0, const float* __restrict__ weights, int V,
                          double* __restrict__ vsum, double* __restrict__ vsq) {
    __shared__ __align__(16) float D[VE_FC][NJ][12];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const bool isj = lane < NJ;
    const int j = isj ? lane : 0;
    unsigned jump = 0;
    for (int r = 0, step = 1; r < 4; ++r, step *= 2) {
        int q = j;
        for (int sidx = 0; sidx < step && q >= 0; ++sidx) q = c_tab.parent[q];
        jump |= (unsigned)(q < 0 ? 255 : q) << (8 * r);
    }
    const float bone[3] = {c_tab.bone[j][0], c_tab.bone[j][1], c_tab.bone[j][2]};
    float jz[3];                 // zero-pose joint position (root at the origin): FK of the identity pose
    {
        const float I[9] = {1.f, 0.f, 0.f, 0.f, 1.f, 0.f, 0.f, 0.f, 1.f};
        float Gi[9];
        warp_fk(I, isj, jump, bone, Gi, jz);
    }

    float w[VE_VPT][NJ], rest[VE_VPT][3], s1[VE_VPT], s2[VE_VPT];
#pragma unroll
    for (int k = 0; k < VE_VPT; ++k) {
        const int v = blockIdx.x * (VE_THREADS * VE_VPT) + k * VE_THREADS + tid;
#pragma unroll
        for (int jj = 0; jj < NJ; ++jj) w[k][jj] = v < V ? __ldg(weights + (size_t)v * NJ + jj) : 0.f;
#pragma unroll
        for (int i = 0; i < 3; ++i) rest[k][i] = v < V ? __ldg(v0 + (size_t)v * 3 + i) : 0.f;
        s1[k] = s2[k] = 0.f;
    }

    for (long long f0 = (long long)blockIdx.y * VE_FC; f0 < n; f0 += (long long)gridDim.y * VE_FC) {
        __syncthreads();         // the previous chunk has been consumed
        for (int ff = warp; ff < VE_FC; ff += VE_THREADS / 32) {
            const long long f = f0 + ff;
            if (f >= n) break;
            float Rp[9], Rt[9], Gp[9], Gt[9], Pp[3], Pt[3];
#pragma unroll
            for (int i = 0; i < 9; ++i) {
                Rp[i] = isj ? __ldg(pose_p + f * 216 + j * 9 + i) : 0.f;
                Rt[i] = isj ? __ldg(pose_t + f * 216 + j * 9 + i) : 0.f;
            }
            warp_fk(Rp, isj, jump, bone, Gp, Pp);
            warp_fk(Rt, isj, jump, bone, Gt, Pt);
            if (isj) {
#pragma unroll
                for (int r = 0; r < 3; ++r) {
                    float tp = Pp[r], tt = Pt[r];
#pragma unroll
                    for (int c = 0; c < 3; ++c) {
                        tp = fmaf(-Gp[r * 3 + c], jz[c], tp);
                        tt = fmaf(-Gt[r * 3 + c], jz[c], tt);
                        D[ff][j][r * 4 + c] = Gp[r * 3 + c] - Gt[r * 3 + c];
                    }
                    D[ff][j][r * 4 + 3] = tp - tt;
                }
            }
        }
        __syncthreads();
        const int nf = (int)min((long long)VE_FC, n - f0);
        for (int ff = 0; ff < nf; ++ff) {
            float a[VE_VPT][12];
#pragma unroll
            for (int k = 0; k < VE_VPT; ++k)
#pragma unroll
                for (int i = 0; i < 12; ++i) a[k][i] = 0.f;
#pragma unroll
            for (int jj = 0; jj < NJ; ++jj) {
                const float4 d0 = *reinterpret_cast<const float4*>(&D[ff][jj][0]);
                const float4 d1 = *reinterpret_cast<const float4*>(&D[ff][jj][4]);
                const float4 d2 = *reinterpret_cast<const float4*>(&D[ff][jj][8]);
                const float d[12] = {d0.x, d0.y, d0.z, d0.w, d1.x, d1.y, d1.z, d1.w, d2.x, d2.y, d2.z, d2.w};
#pragma unroll
                for (int k = 0; k < VE_VPT; ++k)
#pragma unroll
                    for (int i = 0; i < 12; ++i) a[k][i] = fmaf(w[k][jj], d[i], a[k][i]);
            }
#pragma unroll
            for (int k = 0; k < VE_VPT; ++k) {
                float e[3];
#pragma unroll
                for (int r = 0; r < 3; ++r)
                    e[r] = fmaf(a[k][r * 4 + 2], rest[k][2], fmaf(a[k][r * 4 + 1], rest[k][1], fmaf(a[k][r * 4], rest[k][0], a[k][r * 4 + 3])));
                const float q = fmaf(e[2], e[2], fmaf(e[1], e[1], e[0] * e[0]));
                s1[k] += sqrtf(q);
                s2[k] += q;
            }
        }
    }
#pragma unroll
    for (int k = 0; k < VE_VPT; ++k) {
        const int v = blockIdx.x * (VE_THREADS * VE_VPT) + k * VE_THREADS + tid;
        if (v < V) {
            atomicAdd(vsum + v, (double)s1[k]);
            atomicAdd(vsq + v, (double)s2[k]);
        }
    }
}

}  // namespace

int launch_eval_frame_errors(const float* pose_p, const float* pose_t, const float* tran_p, const float* tran_t, int64_t n,
                             float* joint_p, float* joint_t, float* je, float* lae, float* gae, cudaStream_t stream) {
    MP_REQUIRE(pose_p && pose_t && joint_p && joint_t && je && lae && gae && n > 0, "eval_frame_errors: bad arguments");
    MP_TRY(upload_tables());
    // algorithmic bytes: two poses in, two joint sets + three error planes out
    ProfileScope prof("n1_frame_errors", (double)n * (2 * 864.0 + 24 + 2 * 288 + 3 * 96), stream);
    const int blocks = (int)std::min<int64_t>((n + 7) / 8, 148 * 8);
    eval_frame_errors_kernel<<<blocks, 256, 0, stream>>>(pose_p, pose_t, tran_p, tran_t, n, joint_p, joint_t, je, lae, gae);
    MP_CUDA_TRY(cudaGetLastError());
    count_launch();
    return MP_OK;
}

int launch_eval_vertex_errors(const float* pose_p, const float* pose_t, int64_t n, const float* v0, const float* weights, int V,
                              double* vsum, double* vsq, cudaStream_t stream) {
    MP_REQUIRE(pose_p && pose_t && v0 && weights && vsum && vsq && n > 0 && V > 0, "eval_vertex_errors: bad arguments");
    MP_TRY(upload_tables());
    MP_CUDA_TRY(cudaMemsetAsync(vsum, 0, sizeof(double) * V, stream));
    MP_CUDA_TRY(cudaMemsetAsync(vsq, 0, sizeof(double) * V, stream));
    // algorithmic bytes: the two motions in (what the reference reads to build 2 x [n, V, 3] vertex sets), 2 x V doubles out
    ProfileScope prof("n1_vertex_errors", (double)n * 2 * 864.0 + (double)V * (NJ + 3) * 4 + (double)V * 16, stream);
    const int vtiles = (V + VE_THREADS * VE_VPT - 1) / (VE_THREADS * VE_VPT);
    const int64_t chunks = (n + VE_FC - 1) / VE_FC;
    // about 8 CTAs per SM over the whole grid; every CTA walks its frame chunks with a grid stride
    const int ychunks = (int)std::max<int64_t>(1, std::min<int64_t>(chunks, (148 * 8 + vtiles - 1) / vtiles));
    eval_vertex_errors_kernel<<<dim3(vtiles, ychunks), VE_THREADS, 0, stream>>>(pose_p, pose_t, n, v0, weights, V, vsum, vsq);
    MP_CUDA_TRY(cudaGetLastError());
    count_launch();
    return MP_OK;
}

int launch_physics_optimize(const float* pose, const float* vel, const float* contact, const int32_t* lengths, float* state,
                            int B, int T, const mp_physics_params_t* prm, float* pose_out, float* tran_out, float* dbg,
                            int dbg_frame, cudaStream_t stream) {
    MP_REQUIRE(pose && vel && contact && state && pose_out && prm, "physics: null pointer");
    MP_REQUIRE(B > 0 && T > 0, "physics: B = %d, T = %d", B, T);
    MP_REQUIRE(prm->damping >= 0.f && prm->damping_abs > 0.f && prm->w_vel > 0.f && prm->w_contact >= 0.f && prm->fps > 0.f,
               "physics: w_vel, damping_abs and fps must be positive, damping and w_contact non-negative");
    MP_TRY(upload_tables());
    PhysParams p{pose, vel, contact, lengths, state, pose_out, tran_out, dbg, B, T, dbg_frame, prm->vel_scale / prm->fps,
                 prm->w_vel, prm->w_contact, prm->damping, prm->damping_abs, prm->floor_y};
    // algorithmic bytes: pose in + out, velocity, contact, translation
    ProfileScope prof("k8_physics", (double)B * T * (864.0 * 2 + 288 + 8 + 12), stream);
    // skeletons (warps) per CTA: MP_K8_WARPS, default 4 (measured on the pipelined cfg3 step, 256 skeletons: 1 warp per CTA 6.63 ms per
    // step, 4: 6.12, 8: 6.23, 12: 6.07 -- while the kernel alone slows from 3.48 to 3.50 / 3.90 / 4.35 ms as the warps share schedulers)
    static int wpc = 0;
    if (!wpc) {
        const char* v = getenv("MP_K8_WARPS");
        wpc = v ? atoi(v) : 4;
        wpc = std::min(12, std::max(1, wpc));
        MP_CUDA_TRY(cudaFuncSetAttribute(physics_optimize_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(12 * sizeof(PhysScratch))));
    }
    const int w = std::min(wpc, B);
    physics_optimize_kernel<<<(B + w - 1) / w, 32 * w, (size_t)w * sizeof(PhysScratch), stream>>>(p);
    MP_CUDA_TRY(cudaGetLastError());
    count_launch();
    return MP_OK;
}

int physics_prepare() { return upload_tables(); }

int launch_physics_fk(const float* pose, int64_t n, float* glb, float* pos, cudaStream_t stream) {
    MP_REQUIRE(pose && glb && pos && n > 0, "physics_fk: bad arguments");
    MP_TRY(upload_tables());
    const int blocks = (int)std::min<int64_t>((n + 7) / 8, 148 * 8);
    physics_fk_kernel<<<blocks, 256, 0, stream>>>(pose, n, glb, pos);
    MP_CUDA_TRY(cudaGetLastError());
    count_launch();
    return MP_OK;
}

}  // namespace mp
