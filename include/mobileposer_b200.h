/*
 * mobileposer_b200 -- C ABI of the B200 (sm_100a) hot path.
 *
 * The reference (SPICExLAB/MobilePoser) has no FFI: its boundary for this path is the
 * Python class surface of `MobilePoserNet`.  Each entry point
 * below replaces one torch call chain of that surface; the host-side mirror in
 * mobileposer_b200/{modules,net}.py binds them with ctypes (INTEGRATION.md shows the
 * stub a reference maintainer would add).  Citations are relative to /root/reference.
 *
 * Conventions
 *   - all tensors are fp32, row-major contiguous, DEVICE pointers owned by the caller
 *     (torch), unless the name ends in `_host`;
 *   - `stream` is a cudaStream_t passed as void*; work is enqueued, never synchronised
 *     (except the `_host` entry points, which return after the results are on the host);
 *   - no allocation on the forward path: scratch comes from the caller (`*_workspace_bytes`);
 *   - return value: MP_OK (0) or a negative mp_status; mp_last_error() gives the message
 *     (thread local).  The Python mirror raises RuntimeError on non-zero;
 *   - a handle may be used from one thread / one stream at a time.
 */
#ifndef MOBILEPOSER_B200_H_
#define MOBILEPOSER_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MP_ABI_VERSION 1

typedef enum mp_status {
    MP_OK = 0,
    MP_ERR_INVALID = -1,     /* bad argument (shape, null pointer, alignment)               */
    MP_ERR_CUDA = -2,        /* a CUDA runtime call or launch failed                        */
    MP_ERR_WORKSPACE = -3,   /* workspace smaller than *_workspace_bytes()                   */
    MP_ERR_STATE = -4,       /* carried LSTM state does not match the batch (velocity.py:45) */
    MP_ERR_UNSUPPORTED = -5  /* not an sm_100 device / shape outside the built kernels       */
} mp_status;

typedef struct mp_rnn mp_rnn_t; /* packed weights of one RNN head  */
typedef struct mp_net mp_net_t; /* four heads + streams + graphs   */
typedef void* mp_stream_t;      /* cudaStream_t                    */

int mp_abi_version(void);
const char* mp_last_error(void);

/* The constants the kernels bake in (csrc/mp_constants.cuh), for agreement checks against the host-side mirror of the
 * reference's config (mobileposer/config.py:129-142; models/net.py:47-59; SMPL zero pose of articulate/model.py:77-92).
 * Host call, needs no device. */
typedef struct mp_constants {
    int32_t parent[24];       /* SMPL kinematic tree, -1 = root                         */
    int32_t reduced[16];      /* joint_set.reduced                                      */
    int32_t ignored[9];       /* joint_set.ignored                                      */
    int32_t reduced_slot[24]; /* joint -> index in `reduced` or -1 (what K5 indexes by) */
    float j_zero[24][3];      /* zero-pose joints J - J[0] (K8, evaluator kernels)      */
    float feet[6];            /* J[10], J[11]: initial last_lfoot_pos / last_rfoot_pos  */
    float gravity_velocity;   /* joint_set.gravity_velocity                             */
    float vel_div;            /* datasets.fps / amass.vel_scale                         */
    float prob_lo, prob_hi;   /* prob_threshold                                         */
    double floor_y;           /* min(J[10].y, J[11].y) as float32 widened               */
} mp_constants_t;
int mp_constants(mp_constants_t* out);
/* MP_OK iff the current CUDA device is compute capability 10.x. */
int mp_device_check(void);

/* ------------------------------------------------------------------------------------------
 * One RNN head = mobileposer/models/rnn.py:13-33 (linear1 -> ReLU -> 2-layer LSTM -> linear2).
 * Pointers are the tensors of the head's state_dict in torch's own layouts
 * (weight_ih_l{k}[_reverse] [4H, In], weight_hh [4H, H], biases [4H]; gate rows i,f,g,o).
 * ---------------------------------------------------------------------------------------- */
typedef struct mp_rnn_weights {
    int32_t n_input, n_output, n_hidden, n_layers, bidirectional;
    const float* linear1_w; /* [H, n_input]        */
    const float* linear1_b; /* [H]                 */
    const float* linear2_w; /* [n_output, dirs*H]  */
    const float* linear2_b; /* [n_output]          */
    const float* w_ih[2][2]; /* [layer][dir]       */
    const float* w_hh[2][2];
    const float* b_ih[2][2];
    const float* b_hh[2][2];
} mp_rnn_weights_t;

/* Repack the weights into the library's resident layout (device allocations; cold path). */
int mp_rnn_create(mp_rnn_t** out, const mp_rnn_weights_t* w, mp_stream_t stream);
void mp_rnn_destroy(mp_rnn_t* rnn);
size_t mp_rnn_workspace_bytes(const mp_rnn_t* rnn, int32_t B, int32_t T);

/* RNN.forward(x, seq_lengths, h) -> (y, (h_n, c_n))            [rnn.py:20-33]
 *   input  = concat(xa [B,T,ka], xb [B,T,kb]) on the last dim (xb may be NULL, kb = 0): the
 *            torch.cat of net.py:106,113 is folded into the first GEMM;
 *   lengths [B] int32 device (NULL = all T): packed-sequence semantics -- the reverse
 *            direction of sequence b starts at frame lengths[b]-1, frames >= lengths[b] of the
 *            LSTM output are zero so y there equals linear2.bias (pad_packed_sequence);
 *   h0/c0, hn/cn [layers*dirs, B, H] (NULL = zeros / not wanted);
 *   y [B, T, n_output].                                                                      */
int mp_rnn_forward(const mp_rnn_t* rnn, const float* xa, int32_t ka, const float* xb, int32_t kb,
                   int32_t B, int32_t T, const int32_t* lengths,
                   const float* h0, const float* c0, float* hn, float* cn, float* y,
                   void* workspace, size_t workspace_bytes, mp_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * Training step of one head -- first slice of the training path:
 * forward with saved activations and the full backward pass of RNN.forward [models/rnn.py:20-33], plus the Joints loss
 * [models/joints.py:54-75].  Weights are read in torch's layouts straight from the module's parameters (`mp_rnn_weights_t`, no
 * packing: they change every step); gradients are written (not accumulated) in the same layouts.  `mask` = NULL is eval mode
 * (dropout off), otherwise the [B, T, H] tensor that multiplies relu(linear1(x)) (nn.Dropout's keep / (1 - p) pattern).
 * The workspace carries the saved activations from mp_rnn_train_forward to mp_rnn_train_backward.
 * Plain fp32 kernels (correctness first: pinned to the reference's own shared_step + backward); not the inference kernels.
 * ---------------------------------------------------------------------------------------- */
typedef struct mp_rnn_grads {
    float* linear1_w; /* [H, n_input]        */
    float* linear1_b; /* [H]                 */
    float* linear2_w; /* [n_output, dirs*H]  */
    float* linear2_b; /* [n_output]          */
    float* w_ih[2][2]; /* [layer][dir] [4H, In] */
    float* w_hh[2][2]; /* [4H, H]               */
    float* b_ih[2][2]; /* [4H]                  */
    float* b_hh[2][2]; /* [4H] (equal to b_ih's gradient) */
} mp_rnn_grads_t;
size_t mp_rnn_train_workspace_bytes(const mp_rnn_weights_t* w, int32_t B, int32_t T);
int mp_rnn_train_forward(const mp_rnn_weights_t* w, const float* x, int32_t B, int32_t T, const int32_t* lengths,
                         const float* mask, float* y, void* workspace, size_t workspace_bytes, mp_stream_t stream);
int mp_rnn_train_backward(const mp_rnn_weights_t* w, const float* x, int32_t B, int32_t T, const int32_t* lengths,
                          const float* mask, const float* dy, const mp_rnn_grads_t* grads, void* workspace,
                          size_t workspace_bytes, mp_stream_t stream);
/* Joints.shared_step's loss on a padded prediction [B, T, D]: mean squared error to `target` + t_weight x the batch mean of the
 * summed L1 norm of the second differences along T [joints.py:66-75]; writes the scalar (double, device) and d loss / d pred. */
int mp_joints_loss(const float* pred, const float* target, int32_t B, int32_t T, int32_t D, float t_weight, double* loss,
                   float* dpred, mp_stream_t stream);
/* Poser.shared_step's loss [poser.py:65-98] on a padded r6d prediction [B, T, 96]: MSE to pose_t [B, T, 96] + t_weight x jerk L1
 * [poser.py:100-103] + the joint-position MSE through _reduced_global_to_full and the zero-pose forward kinematics
 * [poser.py:93-96] against joints_t [B, T, 72]; writes the scalar and d loss / d pred (Gram-Schmidt and tree adjoints included). */
int mp_poser_loss(const float* pred, const float* pose_t, const float* joints_t, int32_t B, int32_t T, float t_weight,
                  double* loss, float* dpred, mp_stream_t stream);
/* FootContact.shared_step's loss [footcontact.py:31,63]: nn.BCEWithLogitsLoss over the padded [B, T, 2] logits. */
int mp_footcontact_loss(const float* pred, const float* target, int32_t B, int32_t T, double* loss, float* dpred,
                        mp_stream_t stream);
/* Velocity.shared_step's loss [velocity.py:72-86]: sum over n in {1, 3, 9} of the MSE of the T // n windows of n frames. */
int mp_velocity_loss(const float* pred, const float* target, int32_t B, int32_t T, int32_t D, double* loss, float* dpred,
                     mp_stream_t stream);

/* The optimizer step Lightning runs after shared_step's backward [overfit.py:41-50: Trainer(gradient_clip_val=1);
 * joints.py:113-114 (and the other heads): torch.optim.AdamW(self.parameters(), lr=1e-3)] over FLAT fp32 buffers holding all of a
 * head's parameters / gradients / moments.  mp_grad_sq_norm ADDS the sum of squares of `grads` to *sq_norm (double, device; zero it
 * first) -- torch.nn.utils.clip_grad_norm_'s total norm, squared.  mp_adamw_step is one fused pass: g' = g * grad_scale (e.g. 1 / world
 * size after a gradient all-reduce) * min(1, max_norm / (sqrt(*sq_norm) * grad_scale + 1e-6)) when sq_norm is given, then torch's
 * AdamW update (decoupled weight decay, bias corrections of step `step` >= 1, amsgrad off). */
int mp_grad_sq_norm(const float* grads, size_t n, double* sq_norm, mp_stream_t stream);
int mp_adamw_step(float* params, const float* grads, float* exp_avg, float* exp_avg_sq, size_t n, float lr, float beta1, float beta2,
                  float eps, float weight_decay, int32_t step, const double* sq_norm, float max_norm, float grad_scale,
                  mp_stream_t stream);

/* The dense contraction under every Linear / LSTM input projection of the path (rnn.py:22,27,32):
 *   C[M,N] = act(A[M,K] * W[N,K]^T + bias[N]).  mode 0 = library's choice, 1 = fp32 FFMA kernel,
 *   2 = tcgen05 3xTF32 tensor-core kernel (needs N % 256 == 0, K % 16 == 0, relu == 0),
 *   3 = tcgen05 3xFP16 (fp16 hi / scaled-lo split) tensor-core kernel (N % 256 == 0, K % 32 == 0, relu == 0; the operands are
 *       split into stream-ordered scratch memory first),
 *   4 = the same kernel for a narrow output (N <= 256, N % 4 == 0: linear2), W and bias zero-padded to one 256-row tile.
 *   Exposed for tests.   */
int mp_gemm_bias(const float* A, const float* W, const float* bias, float* C, int32_t M, int32_t N, int32_t K,
                 int32_t relu, int32_t mode, mp_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * Input assembly (the step right before the heads).
 *   PoseDataset._process_file_data / _process_combo_data          [data.py:60-61,69-76]
 *   DataLoader._get_imu + smooth_avg                              [loader.py:39-49, utils/model_utils.py:28-37]
 * acc [T, slots_in, 3] and ori [T, slots_in, 3, 3] are the raw per-slot streams (slots_in >= 5, the first 5 are used);
 * combo_masks_host[c] has bit s set when slot s is worn in combo c (amass.combos, config.py:60-73; HOST array, <= 16
 * entries).  imu_out [n_combos, T, 60] = per combo cat(acc / acc_scale [15], ori [45]) with the other slots zeroed;
 * smooth != 0 additionally applies the viewer's 3-tap moving average to the scaled accelerations.                    */
int mp_imu_assemble(const float* acc, const float* ori, int64_t T, int32_t slots_in, const int32_t* combo_masks_host,
                    int32_t n_combos, float acc_scale, int32_t smooth, float* imu_out, mp_stream_t stream);

/* Live-demo normalisation of n_ticks sensor readings [live_demo.py:210-234] (third part of row N2): per sensor
 * quaternion (wxyz, unnormalised) -> rotation [articulate/math/angular.py:224-236], calibration into the SMPL frame
 * (glb_acc = smpl2imu acc - acc_offset, glb_ori = smpl2imu R device2bone; the calibration of live_demo.py:160-177 is
 * passed as HOST arrays: smpl2imu [3,3], device2bone [5,3,3], acc_offsets [5,3]), slot permutation perm_host[5]
 * (the reference's [1,4,3,0,2]), acc / acc_scale, device-combo mask -- or, phone_as_watch != 0, slot 0 <- permuted slot 3
 * and nothing else [live_demo.py:223-226] -- and the cat into imu_out [n_ticks, 60].
 * quat [n_ticks,5,4], acc_raw [n_ticks,5,3] device pointers.                                                            */
int mp_imu_live_normalize(const float* quat, const float* acc_raw, int64_t n_ticks, const float* smpl2imu_host,
                          const float* device2bone_host, const float* acc_offsets_host, const int32_t* perm_host,
                          int32_t combo_mask, int32_t phone_as_watch, float acc_scale, float* imu_out, mp_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * Kinematic tail.
 * ---------------------------------------------------------------------------------------- */
/* MobilePoserNet._reduced_global_to_full                         [net.py:93-99]
 *   r6d [n_frames, 96] -> local rotations [n_frames, 24, 3, 3]
 *   (angular.py:167-182 Gram-Schmidt with NaN->0, model_utils.py:18-25 scatter,
 *    spatial.py:115-123 inverse tree, ignored joints -> I, root = global root).              */
int mp_pose_reduced_global_to_full(const float* r6d, int64_t n_frames, float* pose, mp_stream_t stream);

/* Translation fusion of forward_offline                          [net.py:125-154]
 *   joints [B,T,72], vel [B,T,72], contact [B,T,2] logits -> tran [B,T,3]; per sequence b only
 *   frames < lengths[b] are written (rest zero).  Floor clamp replayed sequentially with the
 *   reference's float64 scalars; prefix sums accumulated in float64.                          */
int mp_tran_offline(const float* joints, const float* vel, const float* contact, const int32_t* lengths,
                    int32_t B, int32_t T, float* tran, mp_stream_t stream);

/* Per-stream state of forward_online                             [net.py:59-64,84-88,173-219]
 * One block of MP_ONLINE_STATE_FLOATS floats + 1 double per stream, device resident:
 *   [0:3] last_lfoot_pos  [3:6] last_rfoot_pos  [6:9] last_root_pos ; double current_root_y   */
#define MP_ONLINE_STATE_FLOATS 12
typedef struct mp_online_state {
    float last_lfoot[3];
    float last_rfoot[3];
    float last_root[3];
    float pad_[3];
    double current_root_y;
    double pad2_;
} mp_online_state_t;

/* Online tick tail for S streams: takes frame `frame_idx` (= past_frames) of the forward outputs
 * pose [S*W,24,3,3], joints/vel [S,W,72], contact [S,W,2]; updates state[S]; writes
 * pose_out [S,24,9], root_out [S,3], contact_out [S,2].                  [net.py:181-208,219] */
int mp_online_update(mp_online_state_t* state, const float* pose, const float* joints, const float* vel,
                     const float* contact, int32_t S, int32_t W, int32_t frame_idx,
                     float* pose_out, float* root_out, float* contact_out, mp_stream_t stream);
/* Sliding IMU window of forward_online (net.py:175): win_out[s] = cat(win_in[s][1:], frame[s]);
 * cold != 0 replicates frame[s] over the whole window (first tick).  win_in != win_out.       */
int mp_online_push_frame(const float* win_in, float* win_out, const float* frame, int32_t S, int32_t W,
                         int32_t cold, mp_stream_t stream);
/* reset(): current_root_y = 0, last_root_pos = 0 (feet NOT reset, net.py:84-88);
 * full != 0 additionally restores the zero-pose feet (constructor values, net.py:59).         */
int mp_online_reset(mp_online_state_t* state, int32_t S, int32_t full, mp_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * K8: kinematic-physics optimizer behind the PHYSICS hook       [net.py:66-69,157-169,211-217]
 * The reference imports `dynamics.PhysicsOptimizer`, a module that is NOT in its tree (its rbdl
 * dependency is neither vendored nor pinned): PARITY UNPINNED.  The algorithm is defined by this
 * repository (float64 statement: oracle/physics_port.py) behind the hook's
 * signature  optimize_frame(pose, jvel, contact, acc) -> (pose, tran) / reset_states().
 * ---------------------------------------------------------------------------------------- */
typedef struct mp_physics_params {
    float w_vel;     /* weight of the joint-velocity consistency rows (24 joints)            */
    float w_contact; /* weight of a stance foot's stationarity rows at contact probability >= 0.9 */
    float damping;   /* Marquardt damping of the 45 rotational unknowns, relative to diag(J^T W J) */
    float damping_abs; /* absolute part of the damping (> 0; joints without leverage stay SPD)  */
    float fps;       /* datasets.fps (config.py:89)                                           */
    float vel_scale; /* amass.vel_scale (config.py:83): raw velocity head -> m/s (net.py:162)  */
    float floor_y;   /* floor height in the zero-pose root frame (net.py:49)                  */
} mp_physics_params_t;
/* Per-skeleton state carried between calls (reset_states() = zero it): root position [3],
 * started flag, previous world joint positions [24,3], padding.                              */
#define MP_PHYSICS_STATE_FLOATS 80
/* Batched optimize_frame: B skeletons (one warp each) walk their T frames in sequence.
 *   pose [B,T,24,3,3] local rotations (K5), vel [B,T,72] RAW velocity-head output, contact
 *   [B,T,2] logits, lengths [B] or NULL, state [B, MP_PHYSICS_STATE_FLOATS] (read and written),
 *   pose_out [B,T,24,3,3] (may alias pose), tran_out [B,T,3] or NULL.  Frames >= lengths[b]
 *   pass through.  The online tick is the T = 1 case with the state kept by the caller.       */
int mp_physics_optimize(const float* pose, const float* vel, const float* contact, const int32_t* lengths,
                        float* state, int32_t B, int32_t T, const mp_physics_params_t* params,
                        float* pose_out, float* tran_out, mp_stream_t stream);
/* Test hook: same call, additionally dumps the normal equations of frame `dbg_frame`
 * (49 x 49 lower triangle in elimination order, row 48 = right-hand side) and the solution [48]
 * into dbg [B, 49*49 + 48].                                                                   */
int mp_physics_optimize_debug(const float* pose, const float* vel, const float* contact, const int32_t* lengths,
                              float* state, int32_t B, int32_t T, const mp_physics_params_t* params,
                              float* pose_out, float* tran_out, float* dbg, int32_t dbg_frame, mp_stream_t stream);
/* SMPL forward kinematics of the optimizer (ParametricModel.forward_kinematics with shape=None,
 * calc_mesh=False, articulate/model.py:208-232): pose [n,24,3,3] local -> global rotations
 * [n,24,3,3] and root-relative joint positions [n,24,3].                                      */
int mp_physics_fk(const float* pose, int64_t n_frames, float* global_rot, float* joint_pos, mp_stream_t stream);

/* Metric side of evaluate.py (the part that needs no mesh): per-frame errors between a predicted and a
 * true motion, FullMotionEvaluator.__call__ [articulate/evaluator.py:316-327] with shape = None:
 *   pose_p / pose_t [n,24,3,3] local rotations, tran_p / tran_t [n,3] or NULL ->
 *   joint_p / joint_t [n,24,3] world joint positions (forward kinematics + translation),
 *   je [n,24] root-aligned joint position error (m), lae / gae [n,24] local / global joint angle error (degrees).
 * The jerk and translation-drift rows and the mean / std reductions are slices of these outputs (host side, torch).   */
int mp_eval_frame_errors(const float* pose_p, const float* pose_t, const float* tran_p, const float* tran_t,
                         int64_t n_frames, float* joint_p, float* joint_t, float* je, float* lae, float* gae,
                         mp_stream_t stream);

/* Mesh row of the evaluator (FullMotionEvaluator.__call__ row 1 [articulate/evaluator.py:319-323] with the linear blend
 * skinning of ParametricModel.forward_kinematics(calc_mesh=True) [articulate/model.py:233-240], mean shape, no pose
 * blendshapes, root alignment -- the configuration evaluate.py uses):
 *   pose_p / pose_t [n,24,3,3] local rotations, rest_vertices [V,3] = v_template - J[0], weights [V,24] skinning weights ->
 *   err_sum [V], err_sq_sum [V] (float64): per vertex, the sum over the n frames of |v_p - v_t| and of |v_p - v_t|^2
 *   (root-aligned; translations cancel).  mean = sum(err_sum) / (n V); std over frames per vertex from both sums.
 * The vertex sets are never materialised (the difference of two skinned meshes is the skinning of the difference of the
 * joint transforms).  The joint rest positions are the library's SMPL constants.                                       */
int mp_eval_vertex_errors(const float* pose_p, const float* pose_t, int64_t n_frames, const float* rest_vertices,
                          const float* weights, int32_t n_vertices, double* err_sum, double* err_sq_sum, mp_stream_t stream);

/* The [10, 2] (mean, std) rows of FullMotionEvaluator.__call__ [articulate/evaluator.py:326-343] from the per-frame outputs of
 * mp_eval_frame_errors: rows 0, 2-9 (joint position / local angle / global angle errors, jitter of both motions, one-second root
 * translation error in cm, the first three again on the joints of `joint_mask_bits` (bit j = joint j)); row 1 (mesh) is NaN.
 * Every row is x.mean() and x.std(dim=0).mean() of a [frames, joints] array (unbiased std over frames), from per-column sums in double. */
int mp_eval_motion_rows(const float* joint_p, const float* joint_t, const float* je, const float* lae, const float* gae,
                        int64_t n_frames, int32_t fps, uint32_t joint_mask_bits, float* rows, mp_stream_t stream);
/* The same for a GROUP of sequences whose per-frame arrays are concatenated: sequence b owns the frames [offsets[b], offsets[b + 1])
 * (`offsets`: n_sequences + 1 int64 on the device) and gets rows[b] ([n_sequences, 10, 2]); one CTA per sequence, side by side, each
 * with the accumulation order of a launch of its own (bit-identical rows).  What evaluate_pose(batch_size > 1) reduces a group with
 * [evaluate.py:39-107: the loop over sequences]. */
int mp_eval_motion_rows_batch(const float* joint_p, const float* joint_t, const float* je, const float* lae, const float* gae,
                              const int64_t* offsets, int32_t n_sequences, int32_t fps, uint32_t joint_mask_bits, float* rows,
                              mp_stream_t stream);

/* Translation-error windows of evaluate_pose(..., evaluate_tran=True) [evaluate.py:66-92], S sequences
 * per call: tran_p / tran_t [S,T,3] predicted / true root translation (padded to T frames), lengths [S] device ints or NULL ->
 *   err [S,7]   mean of |(tran_t[e]-tran_t[s]) - (tran_p[e]-tran_p[s])| / moved(s,e) * w over the frame pairs (s,e) across which
 *               the ground truth moves at least w = 1..7 m (two-pointer sweep, first start per end); NaN where there is no pair
 *   count [S,7] number of pairs.
 * The travelled distance is accumulated sequentially in fp32 like the reference, so the pair sets are the reference's.
 * T <= 51200 (the running distance lives in shared memory).                                                           */
int mp_eval_tran_windows(const float* tran_p, const float* tran_t, const int32_t* lengths, int32_t S, int32_t T,
                         float* err, int32_t* count, mp_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * Whole net = MobilePoserNet.forward / forward_offline           [net.py:101-171]
 * ---------------------------------------------------------------------------------------- */
/* The net borrows the four heads (they must outlive it). */
int mp_net_create(mp_net_t** out, const mp_rnn_t* joints, const mp_rnn_t* pose,
                  const mp_rnn_t* foot_contact, const mp_rnn_t* velocity);
void mp_net_destroy(mp_net_t* net);
size_t mp_net_workspace_bytes(const mp_net_t* net, int32_t B, int32_t T);
/* PHYSICS hook inside the net (net.py:157-169): with non-NULL params every mp_net_forward that computes the
 * translation (tran != NULL, i.e. forward_offline) finishes with K8 over the pose, in place, each sequence from a
 * fresh optimizer state; NULL switches it off (default).  Parity unpinned, see mp_physics_optimize.            */
int mp_net_set_physics(mp_net_t* net, const mp_physics_params_t* params);
/* Tile policy of the tensor-core recurrence (H = 256, batches large enough for it).  A step of that kernel costs about the
 * same for 16 or 64 sequences per cluster, so fuller clusters mean less SM time per sequence and more SMs for whatever runs
 * beside them (the other heads, other batches in flight).  0 (default) = auto: 64 sequences per cluster from 128 sequences
 * up, one wave of clusters covering the batch below; n > 0: n sequences per cluster (<= 64); -1: always one wave (cfg3:
 * 14 clusters of 37 -- 16.3 ms per forward against 14.2 ms with clusters of 64).  Results do not depend on it.        */
int mp_net_set_rec_tile(mp_net_t* net, int32_t sequences_per_tile);
/* 1 = replay the forward as a cached CUDA graph keyed on (pointers, B, T) (default 1). */
int mp_net_set_graph(mp_net_t* net, int32_t enabled);

/* forward: joints -> {pose -> K5, foot_contact, velocity(stateful)} on forked streams.
 *   imu [B,T,60]; vel_h0/c0 -> vel_hn/cn [2,B,256] carried velocity state (NULL h0 = zeros);
 *   outputs: pose [B*T,24,3,3], joints [B,T,72], vel [B,T,72], contact [B,T,2];
 *   tran [B,T,3] may be NULL (forward) or non-NULL (forward_offline's translation, K6).       */
int mp_net_forward(mp_net_t* net, const float* imu, int32_t B, int32_t T, const int32_t* lengths,
                   const float* vel_h0, const float* vel_c0, float* vel_hn, float* vel_cn,
                   float* pose, float* joints, float* vel, float* contact, float* tran,
                   void* workspace, size_t workspace_bytes, mp_stream_t stream);

/* Same through HOST buffers (pinned recommended): H2D of imu/lengths, forward_offline, D2H of
 * pose/joints/tran/contact, then a stream synchronise.  `dev_io` is device staging of at least
 * mp_net_host_staging_bytes(B,T).  Used for the end-to-end number of bench.py.               */
size_t mp_net_host_staging_bytes(int32_t B, int32_t T);
/* The same enqueue with a COMPACT pose transfer: pose6d_host [B*T,16,6] = the first two columns of the 16 non-ignored joints' local
 * rotations (384 B per frame instead of 864 B; the ignored joints are the identity, the third column is the cross product of the two:
 * `model_utils.local6d_to_pose` rebuilds [B*T,24,3,3] up to the matrices' own orthonormality).  For consumers behind a host link that the full pose saturates
 * (8 GPUs x 90 MB per step through one host).  mp_pose_full_to_local6d is the conversion alone (device pointers). */
int mp_net_enqueue_offline_host_compact(mp_net_t* net, const float* imu_host, int32_t B, int32_t T, const int32_t* lengths_host,
                                        float* pose6d_host, float* joints_host, float* tran_host, float* contact_host,
                                        void* dev_io, void* workspace, size_t workspace_bytes, mp_stream_t stream);
int mp_pose_full_to_local6d(const float* pose, int64_t n_frames, float* pose6d, mp_stream_t stream);

/* The same without the final synchronise: everything (H2D, forward, D2H) is enqueued on `stream` and the call
 * returns; the host buffers are valid once the stream has drained.  Two nets (mp_net_create on the same heads), each
 * with its own staging / workspace / stream, give a depth-2 software pipeline over batches: the copies of one batch
 * overlap the kernels of the next.                                                                */
int mp_net_enqueue_offline_host(mp_net_t* net, const float* imu_host, int32_t B, int32_t T,
                                const int32_t* lengths_host, float* pose_host, float* joints_host,
                                float* tran_host, float* contact_host, void* dev_io,
                                void* workspace, size_t workspace_bytes, mp_stream_t stream);
int mp_net_forward_offline_host(mp_net_t* net, const float* imu_host, int32_t B, int32_t T,
                                const int32_t* lengths_host, float* pose_host, float* joints_host,
                                float* tran_host, float* contact_host, void* dev_io,
                                void* workspace, size_t workspace_bytes, mp_stream_t stream);

/* Per-kernel timing with CUDA events on the launching stream.  While enabled every kernel launch of the
 * library is bracketed by an event pair (graphs are bypassed); mp_profile_collect synchronises the
 * device and returns one aggregated entry per kernel name.  bench.py's `roofline` comes from here.  */
typedef struct mp_profile_entry {
    char name[32];
    int64_t launches;
    double total_ms;          /* sum of the launches' device durations            */
    double algorithmic_bytes; /* sum of the launches' algorithmic bytes */
} mp_profile_entry_t;
int mp_profile_enable(int32_t on);
int mp_profile_collect(mp_profile_entry_t* out, int32_t capacity, int32_t* n_out);

/* How many kernels of this library the last mp_net_forward / mp_rnn_forward enqueued
 * (bench.py's gpu_launches).                                                                 */
int64_t mp_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif /* MOBILEPOSER_B200_H_ */
