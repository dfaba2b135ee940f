// Shared helpers of the mobileposer_b200 kernels (sm_100a only).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string>

#include "../../include/mobileposer_b200.h"

namespace mp {

void set_error(const char* fmt, ...);
void count_launch(int n = 1);

// Per-kernel CUDA-event timing on the launching stream (mp_profile_enable / mp_profile_collect).
bool profile_enabled();
struct ProfileScope {
    ProfileScope(const char* name, double algorithmic_bytes, cudaStream_t stream);
    ~ProfileScope();
    cudaStream_t stream_;
    int index_;
};

#define MP_CUDA_TRY(expr)                                                                         \
    do {                                                                                          \
        cudaError_t err__ = (expr);                                                               \
        if (err__ != cudaSuccess) {                                                               \
            mp::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(err__), __FILE__,    \
                          __LINE__);                                                              \
            return MP_ERR_CUDA;                                                                   \
        }                                                                                         \
    } while (0)

#define MP_TRY(expr)                       \
    do {                                   \
        int st__ = (expr);                 \
        if (st__ != MP_OK) return st__;    \
    } while (0)

#define MP_REQUIRE(cond, ...)              \
    do {                                   \
        if (!(cond)) {                     \
            mp::set_error(__VA_ARGS__);    \
            return MP_ERR_INVALID;         \
        }                                  \
    } while (0)

// ---- launchers implemented in the .cu files ------------------------------------------------

// C[M,N] = act(cat(A1[M,K1], A2[M,K2]) * W[N,K1+K2]^T + bias[N]);   relu bit 0: apply max(.,0); bit 1 (FFMA kernel only): write C as
// two [M,N] planes of halves (fp16 hi, fp16 scaled lo) for launch_gemm_f16x3 instead of fp32 -- the same number of bytes
int launch_gemm_bias_act(const float* A1, int K1, const float* A2, int K2, const float* W,
                         const float* bias, float* C, int M, int N, int relu, cudaStream_t stream);

// same contraction on the tensor cores (3xTF32, fp32-grade accuracy), single-source A, no activation
bool gemm_tc_eligible(int M, int N, int K);
int launch_gemm_tf32x3(const float* A, const float* W, const float* bias, float* C, int M, int N, int K,
                       cudaStream_t stream);
// the same with W given as [2N, K] = (hi rows, lo rows), split once by launch_split_weights
int launch_gemm_tf32x3_presplit(const float* A, const float* W_split, const float* bias, float* C, int M, int N, int K,
                                cudaStream_t stream);
int launch_split_weights(const float* W, size_t n, float* W_split, cudaStream_t stream);
// second generation (gemm_f16.cu): fp16 hi / scaled-lo operands, 3 products at the FP16 tensor rate.  Operands arrive split:
// A_split = [2][M, K] halves (hi plane, lo plane), W_split = [2N, K] halves (hi rows, lo rows); launch_split_f16 makes either
// from an fp32 array of n elements (out = n hi halves, then n lo halves)
bool gemm_f16_eligible(int M, int N, int K);
// sched: two zeroed 32-bit words private to the launch (the persistent kernel's dynamic tile counter); left zero on exit
// N_out <= N: C is [M, N_out]; W_split / bias are zero-padded to N rows / entries (narrow outputs: linear2)
int launch_gemm_f16x3(const void* A_split, const void* W_split, const float* bias, float* C, int M, int N, int K, int N_out,
                      unsigned int* sched, cudaStream_t stream);
int launch_split_f16(const float* x, size_t n, void* out_hi_lo, cudaStream_t stream);
// the same kernel with an activation epilogue: flags bit 0 = ReLU, bit 1 = C written as two [M, N_out] planes of halves (hi, scaled lo),
// bit 2 = operands in the interleaved layout (launch_split_f16_il), bit 4 / bit 5 = force the CTA-pair / single-CTA kernel (default: pair
// unless MP_GEMM_PAIR=0)
int launch_gemm_f16x3_act(const void* A_split, const void* W_split, const float* bias, void* C, int M, int N, int K, int N_out,
                          int flags, unsigned int* sched, cudaStream_t stream);
// interleaved operand layout [rows, K/32, {hi[32], lo[32]}] (flags bit 2 of launch_gemm_f16x3_act: 128-byte TMA rows)
int launch_split_f16_il(const float* x, size_t n, void* out, cudaStream_t stream);
extern unsigned long long* g_gemm_dbg;
// cat(A1[M,K1], A2[M,K2]), zero-padded to Kp columns, as (hi, lo) planes [2][M, Kp] of halves (linear1's operand)
int launch_pack_cat_f16(const float* A1, int K1, const float* A2, int K2, size_t M, int Kp, void* out_hi_lo, cudaStream_t stream);
int launch_gemm_ffma(const float* A1, int K1, const float* A2, int K2, const float* W, const float* bias, float* C,
                     int M, int N, int relu, cudaStream_t stream);

struct RecLayerArgs {
    const float* gin;     // [B, T, dirs*4H]  input projection + both biases
    const float4* wpack;  // packed W_hh for this layer (all dirs), see pack_whh
    const float* wT;      // [dirs][H][4H] transposed W_hh (debug kernel)
    const float* w_raw[2];  // per direction: W_hh [4H, H] in torch layout (tensor-core kernel splits it itself)
    float* y;             // [B, T, dirs*H]
    const float* h0;      // [dirs, B, H] or null
    const float* c0;
    float* hn;            // [dirs, B, H] or null
    float* cn;
    const int32_t* lengths;  // [B] or null
    int B, T, H, dirs;
    int tile_hint = 0;    // sequences per cluster tile of the tensor-core recurrence; 0 = one wave covering the batch
    int y_split = 0;      // y is written as two [B,T,dirs*H] planes of halves (fp16 hi, scaled lo) for the next layer's fp16-split projection
                          // (honoured by the fp16-split recurrence only: ask rec_f16_eligible first)
};
int launch_lstm_recurrence(const RecLayerArgs& a, cudaStream_t stream);
// tcgen05 3xTF32 variant for H = 256 and large batches (lstm_rec_tc.cu)
bool rec_tc_eligible(const RecLayerArgs& a);
int launch_lstm_recurrence_tc(const RecLayerArgs& a, cudaStream_t stream);
// second generation: fp16 hi / scaled-lo operands (lstm_rec_f16.cu); supersedes the TF32 kernel (MP_REC_IMPL=tf32 pins the old one)
bool rec_f16_eligible(const RecLayerArgs& a);
int launch_lstm_recurrence_f16(const RecLayerArgs& a, cudaStream_t stream);
// the same kernel with 128 sequences per cluster as four sub-tiles (lstm_rec_f16w.cu): whole tiles of equal length, tile_hint == 128
bool rec_f16w_eligible(const RecLayerArgs& a);
int launch_lstm_recurrence_f16w(const RecLayerArgs& a, cudaStream_t stream);
// number of float4 in the packed recurrent weights of one layer
size_t whh_pack_float4s(int H, int dirs);
// pack W_hh[dirs][4H,H] (device, torch layout) into the register-resident layout + transpose
int launch_pack_whh(const float* const* w_hh_dirs, int H, int dirs, float4* wpack, float* wT,
                    cudaStream_t stream);
int rec_cluster_size(int H);

int launch_reduced_global_to_full(const float* r6d, int64_t n, float* pose, cudaStream_t stream);
// full local pose [n,24,3,3] -> the first two columns of the 16 non-ignored joints [n,16,6] (compact transfer form)
int launch_pose_local6d(const float* pose, int64_t n, float* out, cudaStream_t stream);
int launch_tran_offline(const float* joints, const float* vel, const float* contact,
                        const int32_t* lengths, int B, int T, float* tran, cudaStream_t stream);
int launch_online_update(mp_online_state_t* st, const float* pose, const float* joints,
                         const float* vel, const float* contact, int S, int W, int frame,
                         float* pose_out, float* root_out, float* contact_out, cudaStream_t stream);
int launch_online_reset(mp_online_state_t* st, int S, int full, cudaStream_t stream);
int launch_online_push(const float* win_in, float* win_out, const float* frame, int S, int W, int cold,
                       cudaStream_t stream);

// N2 (inputs.cu)
int launch_imu_assemble(const float* acc, const float* ori, int64_t T, int slots_in, const int32_t* masks_host, int n_combos,
                        float acc_scale, int smooth, float* out, cudaStream_t stream);

int launch_imu_live_normalize(const float* quat, const float* acc_raw, int64_t n, const float* smpl2imu_host,
                              const float* device2bone_host, const float* acc_offsets_host, const int32_t* perm_host,
                              int32_t combo_mask, int phone_as_watch, float acc_scale, float* out, cudaStream_t stream);

// K8 (physics.cu)
int launch_physics_optimize(const float* pose, const float* vel, const float* contact, const int32_t* lengths, float* state,
                            int B, int T, const mp_physics_params_t* prm, float* pose_out, float* tran_out, float* dbg,
                            int dbg_frame, cudaStream_t stream);
int launch_physics_fk(const float* pose, int64_t n, float* glb, float* pos, cudaStream_t stream);
int physics_prepare();
// N1 (physics.cu: shares the SMPL tables and the warp FK)
int launch_eval_frame_errors(const float* pose_p, const float* pose_t, const float* tran_p, const float* tran_t, int64_t n,
                             float* joint_p, float* joint_t, float* je, float* lae, float* gae, cudaStream_t stream);   // uploads the constant tables of the current device (must not happen inside a graph capture)

int launch_eval_vertex_errors(const float* pose_p, const float* pose_t, int64_t n, const float* v0, const float* weights, int V,
                              double* vsum, double* vsq, cudaStream_t stream);
// N1 rows (evaluate.cu): the [10, 2] mean / std table of FullMotionEvaluator from the per-frame errors; row 1 (mesh) left NaN
int launch_eval_motion_rows(const float* jp, const float* jt, const float* je, const float* lae, const float* gae, int64_t n, int fps,
                            unsigned mask_bits, float* rows, cudaStream_t stream);
int launch_eval_motion_rows_batch(const float* jp, const float* jt, const float* je, const float* lae, const float* gae, const long long* offsets,
                                  int n_sequences, int fps, unsigned mask_bits, float* rows, cudaStream_t stream);
// N3 (evaluate.cu)
int launch_eval_tran_windows(const float* tran_p, const float* tran_t, const int32_t* lengths, int S, int T, float* err,
                             int32_t* count, cudaStream_t stream);

// N4 first slice (train.cu): forward with saved activations + backward of one RNN head in torch layouts, and the Joints loss
size_t rnn_train_workspace_bytes(const mp_rnn_weights_t* w, int B, int T);
int rnn_train_forward(const mp_rnn_weights_t* w, const float* x, int B, int T, const int32_t* lengths, const float* mask, float* y,
                      void* workspace, size_t workspace_bytes, cudaStream_t stream);
int rnn_train_backward(const mp_rnn_weights_t* w, const float* x, int B, int T, const int32_t* lengths, const float* mask, const float* dy,
                       const mp_rnn_grads_t* grads, void* workspace, size_t workspace_bytes, cudaStream_t stream);
int poser_loss(const float* pred, const float* pose_t, const float* joints_t, int B, int T, float t_weight, double* loss, float* dpred,
               cudaStream_t stream);
int footcontact_loss(const float* pred, const float* target, int B, int T, double* loss, float* dpred, cudaStream_t stream);
int velocity_loss(const float* pred, const float* target, int B, int T, int D, double* loss, float* dpred, cudaStream_t stream);
int grad_sq_norm(const float* g, size_t n, double* out, cudaStream_t stream);
int adamw_step(float* p, const float* g, float* m, float* v, size_t n, float lr, float beta1, float beta2, float eps, float weight_decay,
               int step, const double* sq_norm, float max_norm, float grad_scale, cudaStream_t stream);
int joints_loss(const float* pred, const float* target, int B, int T, int D, float t_weight, double* loss, float* dpred, cudaStream_t stream);

// ---- PTX helpers (clusters, mbarrier, distributed shared memory) ---------------------------
#ifdef __CUDACC__
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// shared::cta address -> shared::cluster address of the same offset in CTA `rank`
__device__ __forceinline__ uint32_t mapa_u32(uint32_t addr, uint32_t rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
    return r;
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init_cluster() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
// Wait for a phase of a CTA-local mbarrier.  The data it guards arrives through the async proxy
// (st.async / cp.async.bulk with complete_tx), whose writes are made visible by the phase completion
// itself, so the default .acquire.cta semantics suffice -- .acquire.cluster would add a CCTL.IVALL
// (L1 invalidate) per wait, 6% of the first version's stall samples.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "MP_WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra MP_DONE_%=;\n\t"
        "bra MP_WAIT_%=;\n\t"
        "MP_DONE_%=:\n\t"
        "}" ::"r"(bar), "r"(parity)
        : "memory");
}
// 4-byte remote store that completes `4` tx bytes on the destination CTA's mbarrier
__device__ __forceinline__ void st_async_f32(uint32_t remote_addr, float v, uint32_t remote_bar) {
    asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.b32 [%0], %1, [%2];" ::"r"(remote_addr),
                 "r"(__float_as_uint(v)), "r"(remote_bar)
                 : "memory");
}
// bulk copy local smem -> (possibly remote) cluster smem, completing tx bytes on the remote mbarrier
__device__ __forceinline__ void bulk_copy_s2c(uint32_t remote_dst, uint32_t local_src, uint32_t bytes,
                                              uint32_t remote_bar) {
    asm volatile("cp.async.bulk.shared::cluster.shared::cta.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     remote_dst),
                 "r"(local_src), "r"(bytes), "r"(remote_bar)
                 : "memory");
}
__device__ __forceinline__ void fence_proxy_async_smem() {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ float sigmoidf_acc(float x) { return 1.0f / (1.0f + expf(-x)); }
#endif

}  // namespace mp
