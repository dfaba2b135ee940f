/* TEST INFRASTRUCTURE ONLY -- plain-C (float64) statement of the kinematic-physics optimizer behind the PHYSICS hook.
 *
 * PARITY UNPINNED: the reference calls dynamics.PhysicsOptimizer at mobileposer/models/net.py:66-69,157-169,211-217 but
 * the module is not in its tree.  The algorithm is this repository's; the numpy file
 * oracle/physics_port.py is the readable statement, this file is the same thing compiled, so that bench.py's CPU arm
 * (cpu_baseline / --impl reference) times K8 as native multi-threaded code rather than as Python loops.  It follows
 * physics_port.py function by function (explicit Jacobian, dense normal equations, dense Cholesky) and is held to it
 * by tests/test_physics_oracle.py.  Only tests/, __graft_entry__.smoke() and bench.py's CPU arm may load it.
 *
 *   gcc -O3 -fopenmp -shared -fPIC -o oracle/_build/libphysics_port.so oracle/physics_port.c -lm      (oracle/build.py)
 */
#include <math.h>
#include <omp.h>
#include <stdint.h>
#include <string.h>

#define NJ 24
#define NOPT 15
#define NX 48
#define NROWS 78 /* 24 joints x 3 velocity rows + 2 feet x 3 contact rows */

typedef struct mp_oracle_physics_params {
    double w_vel, w_contact, damping, damping_abs, fps, vel_scale, floor_y;
} mp_oracle_physics_params_t;

/* per-skeleton state: p[3], started, q_prev[72] */
#define MP_ORACLE_PHYSICS_STATE_DOUBLES 76

static const int kParent[NJ] = {-1, 0, 0, 0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 9, 9, 12, 13, 14, 16, 17, 18, 19, 20, 21};
static const int kOpt[NOPT] = {1, 2, 3, 4, 5, 6, 9, 12, 13, 14, 15, 16, 17, 18, 19}; /* joint_set.reduced minus the root */

/* articulate/model.py:208-232 with shape=None: G_j = G_parent R_j, P_j = P_parent + G_parent bone_j */
static void fk(const double* bone, const double* R, double* G, double* P) {
    memcpy(G, R, 9 * sizeof(double));
    P[0] = P[1] = P[2] = 0.0;
    for (int j = 1; j < NJ; ++j) {
        const double* Gp = G + 9 * kParent[j];
        const double* Rj = R + 9 * j;
        double* Gj = G + 9 * j;
        for (int r = 0; r < 3; ++r) {
            for (int c = 0; c < 3; ++c) Gj[3 * r + c] = Gp[3 * r] * Rj[c] + Gp[3 * r + 1] * Rj[3 + c] + Gp[3 * r + 2] * Rj[6 + c];
            P[3 * j + r] = P[3 * kParent[j] + r] + Gp[3 * r] * bone[3 * j] + Gp[3 * r + 1] * bone[3 * j + 1] + Gp[3 * r + 2] * bone[3 * j + 2];
        }
    }
}

static double prob_to_weight(double logit) { /* net.py:90-91 on sigmoid(logit) */
    double p = 1.0 / (1.0 + exp(-logit));
    p = p < 0.5 ? 0.5 : (p > 0.9 ? 0.9 : p);
    return (p - 0.5) / (0.9 - 0.5);
}

static void exp_so3(const double* w, double* E) {
    const double th2 = w[0] * w[0] + w[1] * w[1] + w[2] * w[2], th = sqrt(th2);
    double A, B;
    if (th < 1e-8) {
        A = 1.0 - th2 / 6.0;
        B = 0.5 - th2 / 24.0;
    } else {
        A = sin(th) / th;
        B = (1.0 - cos(th)) / th2;
    }
    const double K[9] = {0, -w[2], w[1], w[2], 0, -w[0], -w[1], w[0], 0};
    for (int r = 0; r < 3; ++r)
        for (int c = 0; c < 3; ++c) {
            double k2 = 0.0;
            for (int m = 0; m < 3; ++m) k2 += K[3 * r + m] * K[3 * m + c];
            E[3 * r + c] = (r == c ? 1.0 : 0.0) + A * K[3 * r + c] + B * k2;
        }
}

/* one frame of one skeleton; state = {p[3], started, q_prev[72]} */
static void optimize_frame(const mp_oracle_physics_params_t* prm, const double* bone, const float* pose, const float* vel, const float* contact,
                           double* state, float* pose_out, float* tran_out) {
    double R[NJ * 9], G[NJ * 9], P[NJ * 3], Rn[NJ * 9];
    double* p = state;
    double* q_prev = state + 4;
    for (int i = 0; i < NJ * 9; ++i) Rn[i] = R[i] = pose[i];
    fk(bone, R, G, P);
    double d[3] = {0, 0, 0};
    if (state[3] != 0.0) {
        /* position Jacobian: column (k, a) of joint j = G_k[:, a] x (P_j - P_k) for k an ancestor of j */
        double A[NROWS][NX], bvec[NROWS], wvec[NROWS], H[NX][NX], g[NX];
        memset(A, 0, sizeof(A));
        for (int j = 0; j < NJ; ++j) {
            for (int k = kParent[j]; k >= 0; k = kParent[k]) {
                int ci = -1;
                for (int i = 0; i < NOPT; ++i)
                    if (kOpt[i] == k) ci = i;
                if (ci < 0) continue;
                const double r[3] = {P[3 * j] - P[3 * k], P[3 * j + 1] - P[3 * k + 1], P[3 * j + 2] - P[3 * k + 2]};
                for (int a = 0; a < 3; ++a) {
                    const double gx = G[9 * k + a], gy = G[9 * k + 3 + a], gz = G[9 * k + 6 + a];
                    A[3 * j + 0][3 * ci + a] = gy * r[2] - gz * r[1];
                    A[3 * j + 1][3 * ci + a] = gz * r[0] - gx * r[2];
                    A[3 * j + 2][3 * ci + a] = gx * r[1] - gy * r[0];
                }
            }
            for (int r3 = 0; r3 < 3; ++r3) {
                A[3 * j + r3][45 + r3] = 1.0;
                bvec[3 * j + r3] = q_prev[3 * j + r3] + (double)vel[3 * j + r3] * prm->vel_scale / prm->fps - (p[r3] + P[3 * j + r3]);
                wvec[3 * j + r3] = prm->w_vel;
            }
        }
        for (int f = 0; f < 2; ++f) {
            const int jf = 10 + f;
            const double wc = prm->w_contact * prob_to_weight((double)contact[f]);
            for (int r3 = 0; r3 < 3; ++r3) {
                memcpy(A[72 + 3 * f + r3], A[3 * jf + r3], sizeof(A[0]));
                bvec[72 + 3 * f + r3] = q_prev[3 * jf + r3] - (p[r3] + P[3 * jf + r3]);
                wvec[72 + 3 * f + r3] = wc;
            }
        }
        /* H = A^T W A (lower triangle), g = A^T W b; rows are sparse (a joint only sees its ancestors) */
        memset(H, 0, sizeof(H));
        memset(g, 0, sizeof(g));
        for (int r = 0; r < NROWS; ++r)
            for (int i = 0; i < NX; ++i) {
                const double ai = A[r][i] * wvec[r];
                if (ai == 0.0) continue;
                g[i] += ai * bvec[r];
                for (int j = 0; j <= i; ++j) H[i][j] += ai * A[r][j];
            }
        for (int i = 0; i < 45; ++i) H[i][i] = H[i][i] * (1.0 + prm->damping) + prm->damping_abs;
        /* dense Cholesky H = L L^T, then L y = g, L^T x = y */
        for (int k = 0; k < NX; ++k) {      /* right looking: the trailing update runs over contiguous rows */
            double lk[NX];
            const double dk = sqrt(H[k][k]);
            H[k][k] = dk;
            for (int i = k + 1; i < NX; ++i) lk[i] = H[i][k] = H[i][k] / dk;
            for (int i = k + 1; i < NX; ++i) {
                const double lik = lk[i];
                for (int j = k + 1; j <= i; ++j) H[i][j] -= lik * lk[j];
            }
        }
        for (int i = 0; i < NX; ++i) {
            double acc = g[i];
            for (int m = 0; m < i; ++m) acc -= H[i][m] * g[m];
            g[i] = acc / H[i][i];
        }
        for (int i = NX - 1; i >= 0; --i) {
            double acc = g[i];
            for (int m = i + 1; m < NX; ++m) acc -= H[m][i] * g[m];
            g[i] = acc / H[i][i];
        }
        for (int ci = 0; ci < NOPT; ++ci) {
            const int k = kOpt[ci];
            double E[9];
            exp_so3(g + 3 * ci, E);
            for (int r = 0; r < 3; ++r)
                for (int c = 0; c < 3; ++c)
                    Rn[9 * k + 3 * r + c] = R[9 * k + 3 * r] * E[c] + R[9 * k + 3 * r + 1] * E[3 + c] + R[9 * k + 3 * r + 2] * E[6 + c];
        }
        d[0] = g[45]; d[1] = g[46]; d[2] = g[47];
        fk(bone, Rn, G, P);
    }
    const double foot_y = fmin(P[3 * 10 + 1], P[3 * 11 + 1]) + p[1] + d[1];
    if (foot_y < prm->floor_y) d[1] += prm->floor_y - foot_y;
    for (int r = 0; r < 3; ++r) p[r] += d[r];
    for (int j = 0; j < NJ; ++j)
        for (int r = 0; r < 3; ++r) q_prev[3 * j + r] = p[r] + P[3 * j + r];
    state[3] = 1.0;
    for (int i = 0; i < NJ * 9; ++i) pose_out[i] = (float)Rn[i];
    if (tran_out)
        for (int r = 0; r < 3; ++r) tran_out[r] = (float)p[r];
}

/* torchrun exports OMP_NUM_THREADS=1 to its children; the CPU arm of bench.py asks for the host's cores explicitly. */
void mp_oracle_set_threads(int32_t n) {
    if (n > 0) omp_set_num_threads(n);
}

/* pose [B,T,24,9] f32, vel [B,T,72] raw velocity head, contact [B,T,2] logits, lengths [B] or NULL,
 * state [B, 76] doubles (zero = reset), j_zero [24,3] zero-pose joints; pose_out [B,T,24,9], tran_out [B,T,3] or NULL.
 * Sequences are independent: one OpenMP thread each. */
int mp_oracle_physics_optimize(const float* pose, const float* vel, const float* contact, const int32_t* lengths, double* state, int32_t B,
                               int32_t T, const mp_oracle_physics_params_t* prm, const double* j_zero, float* pose_out, float* tran_out) {
    double bone[NJ * 3];
    for (int j = 0; j < NJ; ++j)
        for (int r = 0; r < 3; ++r) bone[3 * j + r] = j_zero[3 * j + r] - (kParent[j] >= 0 ? j_zero[3 * kParent[j] + r] : 0.0);
#pragma omp parallel for schedule(dynamic, 1)
    for (int b = 0; b < B; ++b) {
        int len = lengths ? lengths[b] : T;
        len = len < 0 ? 0 : (len > T ? T : len);
        double* st = state + (size_t)b * MP_ORACLE_PHYSICS_STATE_DOUBLES;
        for (int t = 0; t < T; ++t) {
            const size_t f = (size_t)b * T + t;
            if (t < len) {
                optimize_frame(prm, bone, pose + f * 216, vel + f * 72, contact + f * 2, st, pose_out + f * 216, tran_out ? tran_out + f * 3 : 0);
            } else {
                if (pose_out != pose) memcpy(pose_out + f * 216, pose + f * 216, 216 * sizeof(float));
                if (tran_out)
                    for (int r = 0; r < 3; ++r) tran_out[f * 3 + r] = (float)st[r];
            }
        }
    }
    return 0;
}
