// C ABI of mobileposer_b200 (include/mobileposer_b200.h): weight packing, one-head forward,
// whole-net forward on forked streams replayed as a CUDA graph, host-buffer entry point.
#include "mp_common.cuh"
#include "mp_constants.cuh"

#include <stdarg.h>
#include <string.h>

#include <vector>

namespace mp {

static thread_local char g_err[512] = "";
static thread_local int64_t g_launches = 0;

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}
void count_launch(int n) { g_launches += n; }

struct ProfileRec {
    char name[32];
    double bytes;
    cudaEvent_t e0, e1;
};
static bool g_profile = false;
static std::vector<ProfileRec> g_profile_recs;
bool profile_enabled() { return g_profile; }
ProfileScope::ProfileScope(const char* name, double bytes, cudaStream_t stream) : stream_(stream), index_(-1) {
    if (!g_profile) return;
    ProfileRec r;
    snprintf(r.name, sizeof(r.name), "%s", name);
    r.bytes = bytes;
    if (cudaEventCreate(&r.e0) != cudaSuccess || cudaEventCreate(&r.e1) != cudaSuccess) return;
    cudaEventRecord(r.e0, stream);
    index_ = (int)g_profile_recs.size();
    g_profile_recs.push_back(r);
}
ProfileScope::~ProfileScope() {
    if (index_ >= 0) cudaEventRecord(g_profile_recs[index_].e1, stream_);
}

namespace {

inline size_t align_up(size_t x, size_t a = 256) { return (x + a - 1) / a * a; }

// The recurrence kernel wants the 4 gate pre-activations of a hidden unit next to each other, so the
// rows of W_ih (and the summed biases) are re-ordered from torch's (gate, unit) to (unit, gate).
__global__ void bias_sum_permute_kernel(const float* __restrict__ a, const float* __restrict__ b, float* __restrict__ o, int H) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;   // torch row = gate*H + unit
    if (i < 4 * H) o[(i % H) * 4 + i / H] = a[i] + b[i];
}
__global__ void permute_wih_kernel(const float* __restrict__ w, float* __restrict__ o, int H, int in) {
    const size_t n = (size_t)4 * H * in;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const int row = i / in, k = i - (size_t)row * in;
        o[((size_t)(row % H) * 4 + row / H) * in + k] = w[i];
    }
}

}  // namespace
}  // namespace mp

using namespace mp;

struct mp_rnn {
    int n_in = 0, n_out = 0, H = 0, dirs = 1;
    void* blob = nullptr;
    float *w1 = nullptr, *b1 = nullptr, *w2 = nullptr, *b2 = nullptr;
    float* wih[2] = {nullptr, nullptr};    // [dirs*4H, In_l]   both directions stacked on N
    float* wih_split[2] = {nullptr, nullptr};   // [2 * dirs*4H, In_l]: TF32 hi rows, then lo rows (tensor-core projection), or null
    void* wih_f16[2] = {nullptr, nullptr};      // [2 * dirs*4H, In_l] halves: fp16 hi rows, then scaled-lo rows (gemm_f16.cu), or null
    void* w2_f16 = nullptr;                     // linear2 for the same kernel: [2 * 256, dirs*H] halves, rows >= n_out zero; or null
    float* b2_pad = nullptr;                    // [256] linear2 bias, zero padded
    void* w1_f16 = nullptr;                     // linear1 for the same kernel: [2 * 256, k1p] halves (rows >= H and columns >= n_in zero); or null
    float* b1_pad = nullptr;                    // [256] linear1 bias, zero padded
    int k1p = 0;                                // n_in rounded up to the kernel's K step (32)
    float* bsum[2] = {nullptr, nullptr};   // [dirs*4H]         b_ih + b_hh
    float4* whh_pack[2] = {nullptr, nullptr};
    float* whh_t[2] = {nullptr, nullptr};
    float* whh_raw[2] = {nullptr, nullptr};  // [dirs][4H, H] torch layout
};

struct mp_net {
    const mp_rnn *joints = nullptr, *pose = nullptr, *foot = nullptr, *vel = nullptr;
    cudaStream_t s_foot = nullptr, s_vel = nullptr, s_cap = nullptr;
    cudaEvent_t ev_joints = nullptr, ev_foot = nullptr, ev_vel = nullptr;
    int graph_enabled = 1;
    int rec_tile = 0;                // sequences per cluster tile of the tensor-core recurrence (mp_net_set_rec_tile), 0 = auto
    int physics_on = 0;              // K8 tail of forward_offline (mp_net_set_physics)
    uint64_t physics_epoch = 0;      // bumps whenever the parameters change: part of the graph key
    mp_physics_params_t phys = {};
    struct Entry {
        std::vector<uintptr_t> key;
        cudaGraphExec_t exec = nullptr;
        int64_t kernels = 0;
        uint64_t stamp = 0;
    };
    std::vector<Entry> cache;
    uint64_t clock = 0;
};

extern "C" {

int mp_abi_version(void) { return MP_ABI_VERSION; }
const char* mp_last_error(void) { return g_err; }
int64_t mp_launch_count(void) { return g_launches; }

int mp_constants(mp_constants_t* out) {
    MP_REQUIRE(out, "constants: null");
    const int parent[24] = MP_SMPL_PARENT_INIT, reduced[16] = MP_REDUCED_INIT, ignored[9] = MP_IGNORED_INIT, slot[24] = MP_REDUCED_SLOT_INIT;
    const float j0[24][3] = MP_SMPL_J_ZERO_INIT, feet[6] = MP_FEET_INIT;
    memset(out, 0, sizeof(*out));
    memcpy(out->parent, parent, sizeof(parent));
    memcpy(out->reduced, reduced, sizeof(reduced));
    memcpy(out->ignored, ignored, sizeof(ignored));
    memcpy(out->reduced_slot, slot, sizeof(slot));
    memcpy(out->j_zero, j0, sizeof(j0));
    memcpy(out->feet, feet, sizeof(feet));
    out->gravity_velocity = kGravityVel;
    out->vel_div = kVelDiv;
    out->prob_lo = kProbLo;
    out->prob_hi = kProbHi;
    out->floor_y = kFloorY;
    return MP_OK;
}

int mp_profile_enable(int32_t on) {
    for (auto& r : g_profile_recs) {
        cudaEventDestroy(r.e0);
        cudaEventDestroy(r.e1);
    }
    g_profile_recs.clear();
    g_profile = on != 0;
    return MP_OK;
}

int mp_profile_collect(mp_profile_entry_t* out, int32_t capacity, int32_t* n_out) {
    MP_REQUIRE(out && n_out && capacity > 0, "profile_collect: bad arguments");
    MP_CUDA_TRY(cudaDeviceSynchronize());
    int n = 0;
    for (auto& r : g_profile_recs) {
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, r.e0, r.e1) != cudaSuccess) continue;
        int k = 0;
        while (k < n && strncmp(out[k].name, r.name, sizeof(out[k].name)) != 0) ++k;
        if (k == n) {
            if (n == capacity) continue;
            memset(&out[n], 0, sizeof(out[n]));
            snprintf(out[n].name, sizeof(out[n].name), "%s", r.name);
            ++n;
        }
        out[k].launches += 1;
        out[k].total_ms += ms;
        out[k].algorithmic_bytes += r.bytes;
    }
    *n_out = n;
    return MP_OK;
}

int mp_device_check(void) {
    int dev = 0, major = 0;
    MP_CUDA_TRY(cudaGetDevice(&dev));
    MP_CUDA_TRY(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev));
    if (major != 10) {
        set_error("mobileposer_b200 is built for sm_100a only; device %d is compute capability %d.x", dev, major);
        return MP_ERR_UNSUPPORTED;
    }
    return MP_OK;
}

// ------------------------------------------------------------------------------------------------
int mp_rnn_create(mp_rnn_t** out, const mp_rnn_weights_t* w, mp_stream_t stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    MP_REQUIRE(out && w, "rnn_create: null argument");
    MP_REQUIRE(w->n_layers == 2, "rnn_create: only the reference's 2-layer LSTM is built (got %d)", w->n_layers);
    MP_REQUIRE(w->n_hidden == 256 || w->n_hidden == 64, "rnn_create: hidden size %d not built (64, 256)", w->n_hidden);
    MP_REQUIRE(w->n_input > 0 && (w->n_input & 3) == 0, "rnn_create: n_input %d must be a multiple of 4", w->n_input);
    MP_REQUIRE(w->n_output > 0, "rnn_create: n_output");
    MP_TRY(mp_device_check());
    const int H = w->n_hidden, dirs = w->bidirectional ? 2 : 1;
    MP_REQUIRE(w->linear1_w && w->linear1_b && w->linear2_w && w->linear2_b, "rnn_create: null linear weights");
    for (int l = 0; l < 2; ++l)
        for (int d = 0; d < dirs; ++d)
            MP_REQUIRE(w->w_ih[l][d] && w->w_hh[l][d] && w->b_ih[l][d] && w->b_hh[l][d], "rnn_create: null LSTM weights (layer %d dir %d)", l, d);

    mp_rnn* r = new mp_rnn();
    r->n_in = w->n_input; r->n_out = w->n_output; r->H = H; r->dirs = dirs;
    const int in_l[2] = {H, dirs * H};
    size_t off = 0;
    auto take = [&](size_t floats) { size_t o = off; off = align_up(off + floats * sizeof(float)); return o; };
    const size_t o_w1 = take((size_t)H * r->n_in), o_b1 = take(H);
    const size_t o_w2 = take((size_t)r->n_out * dirs * H), o_b2 = take(r->n_out);
    const bool lin2_tc = r->n_out >= 16 && r->n_out <= 256 && (r->n_out & 3) == 0 && (dirs * H) % 32 == 0;
    const size_t o_w2pad = take(lin2_tc ? (size_t)256 * dirs * H : 0), o_w2f16 = take(lin2_tc ? (size_t)256 * dirs * H : 0),
                 o_b2pad = take(lin2_tc ? 256 : 0);
    // linear1 on the same kernel: [H, n_in] padded to one 256-row tile and to a multiple of the K step
    r->k1p = (r->n_in + 31) / 32 * 32;
    const bool lin1_tc = (H & 7) == 0 && H <= 256;
    const size_t o_w1pad = take(lin1_tc ? (size_t)256 * r->k1p : 0), o_w1f16 = take(lin1_tc ? (size_t)256 * r->k1p : 0),
                 o_b1pad = take(lin1_tc ? 256 : 0);
    size_t o_wih[2], o_bs[2], o_pk[2], o_wt[2], o_raw[2], o_split[2], o_f16[2];
    for (int l = 0; l < 2; ++l) {
        o_wih[l] = take((size_t)dirs * 4 * H * in_l[l]);
        o_split[l] = take((size_t)2 * dirs * 4 * H * in_l[l]);
        o_f16[l] = take((size_t)dirs * 4 * H * in_l[l]);          // 2 x n halves = n floats
        o_bs[l] = take((size_t)dirs * 4 * H);
        o_pk[l] = take(whh_pack_float4s(H, dirs) * 4);
        o_wt[l] = take((size_t)dirs * 4 * H * H);
        o_raw[l] = take((size_t)dirs * 4 * H * H);
    }
    cudaError_t e = cudaMalloc(&r->blob, off);
    if (e != cudaSuccess) {
        delete r;
        set_error("rnn_create: cudaMalloc(%zu) failed: %s", off, cudaGetErrorString(e));
        return MP_ERR_CUDA;
    }
    char* base = (char*)r->blob;
    r->w1 = (float*)(base + o_w1); r->b1 = (float*)(base + o_b1);
    r->w2 = (float*)(base + o_w2); r->b2 = (float*)(base + o_b2);
    int st = MP_OK;
    auto copy = [&](float* dst, const float* src, size_t floats) {
        if (st == MP_OK && cudaMemcpyAsync(dst, src, floats * sizeof(float), cudaMemcpyDeviceToDevice, stream) != cudaSuccess) {
            set_error("rnn_create: weight copy failed: %s", cudaGetErrorString(cudaGetLastError()));
            st = MP_ERR_CUDA;
        }
    };
    copy(r->w1, w->linear1_w, (size_t)H * r->n_in);
    copy(r->b1, w->linear1_b, H);
    copy(r->w2, w->linear2_w, (size_t)r->n_out * dirs * H);
    copy(r->b2, w->linear2_b, r->n_out);
    if (lin2_tc && st == MP_OK) {
        // linear2 on the fp16-split tensor-core kernel: the weight padded with zero rows to one 256-row tile, then split
        float* w2pad = (float*)(base + o_w2pad);
        r->b2_pad = (float*)(base + o_b2pad);
        if (cudaMemsetAsync(w2pad, 0, (size_t)256 * dirs * H * sizeof(float), stream) != cudaSuccess ||
            cudaMemsetAsync(r->b2_pad, 0, 256 * sizeof(float), stream) != cudaSuccess) {
            set_error("rnn_create: memset failed: %s", cudaGetErrorString(cudaGetLastError()));
            st = MP_ERR_CUDA;
        }
        copy(w2pad, w->linear2_w, (size_t)r->n_out * dirs * H);
        copy(r->b2_pad, w->linear2_b, r->n_out);
        if (st == MP_OK) {
            r->w2_f16 = base + o_w2f16;
            st = launch_split_f16(w2pad, (size_t)256 * dirs * H, r->w2_f16, stream);
        }
    }
    if (lin1_tc && st == MP_OK) {
        float* w1pad = (float*)(base + o_w1pad);
        r->b1_pad = (float*)(base + o_b1pad);
        if (cudaMemsetAsync(w1pad, 0, (size_t)256 * r->k1p * sizeof(float), stream) != cudaSuccess ||
            cudaMemsetAsync(r->b1_pad, 0, 256 * sizeof(float), stream) != cudaSuccess ||
            cudaMemcpy2DAsync(w1pad, (size_t)r->k1p * sizeof(float), w->linear1_w, (size_t)r->n_in * sizeof(float),
                              (size_t)r->n_in * sizeof(float), H, cudaMemcpyDeviceToDevice, stream) != cudaSuccess) {
            set_error("rnn_create: linear1 padding failed: %s", cudaGetErrorString(cudaGetLastError()));
            st = MP_ERR_CUDA;
        }
        copy(r->b1_pad, w->linear1_b, H);
        if (st == MP_OK) {
            r->w1_f16 = base + o_w1f16;
            st = launch_split_f16(w1pad, (size_t)256 * r->k1p, r->w1_f16, stream);
        }
    }
    for (int l = 0; l < 2 && st == MP_OK; ++l) {
        r->wih[l] = (float*)(base + o_wih[l]);
        r->bsum[l] = (float*)(base + o_bs[l]);
        r->whh_pack[l] = (float4*)(base + o_pk[l]);
        r->whh_t[l] = (float*)(base + o_wt[l]);
        r->whh_raw[l] = (float*)(base + o_raw[l]);
        for (int d = 0; d < dirs; ++d) copy(r->whh_raw[l] + (size_t)d * 4 * H * H, w->w_hh[l][d], (size_t)4 * H * H);
        const float* whh[2] = {w->w_hh[l][0], w->w_hh[l][dirs - 1]};
        for (int d = 0; d < dirs; ++d) {
            permute_wih_kernel<<<296, 256, 0, stream>>>(w->w_ih[l][d], r->wih[l] + (size_t)d * 4 * H * in_l[l], H, in_l[l]);
            bias_sum_permute_kernel<<<(4 * H + 255) / 256, 256, 0, stream>>>(w->b_ih[l][d], w->b_hh[l][d], r->bsum[l] + (size_t)d * 4 * H, H);
        }
        if (st == MP_OK) st = launch_pack_whh(whh, H, dirs, r->whh_pack[l], r->whh_t[l], stream);
        // pre-split copy of the (permuted, stacked) W_ih for the tensor-core projection: the shapes it accepts are N % 256 == 0
        // and K % 16 == 0 (gemm_tc_eligible); the H = 64 head (N = 512, K = 64 / 128) qualifies too
        if (st == MP_OK && (dirs * 4 * H) % 256 == 0 && in_l[l] % 16 == 0) {
            r->wih_split[l] = (float*)(base + o_split[l]);
            st = launch_split_weights(r->wih[l], (size_t)dirs * 4 * H * in_l[l], r->wih_split[l], stream);
        }
        if (st == MP_OK && (dirs * 4 * H) % 256 == 0 && in_l[l] % 32 == 0) {
            r->wih_f16[l] = base + o_f16[l];
            st = launch_split_f16(r->wih[l], (size_t)dirs * 4 * H * in_l[l], r->wih_f16[l], stream);
        }
    }
    if (st == MP_OK && cudaStreamSynchronize(stream) != cudaSuccess) {
        set_error("rnn_create: packing failed: %s", cudaGetErrorString(cudaGetLastError()));
        st = MP_ERR_CUDA;
    }
    if (st != MP_OK) {
        cudaFree(r->blob);
        delete r;
        return st;
    }
    *out = r;
    return MP_OK;
}

void mp_rnn_destroy(mp_rnn_t* r) {
    if (!r) return;
    cudaFree(r->blob);
    delete r;
}

static void rnn_ws_layout(const mp_rnn* r, size_t M, size_t* o_x1, size_t* o_gin, size_t* o_y0, size_t* o_y1, size_t* o_xs, size_t* total) {
    size_t off = 256;      // first 256 bytes: tile-scheduler words of the persistent projection launches (cleared per forward)
    auto take = [&](size_t floats) { size_t o = off; off = align_up(off + floats * sizeof(float)); return o; };
    *o_x1 = take(M * r->H);
    *o_gin = take(M * r->dirs * 4 * r->H);
    *o_y0 = take(M * r->dirs * r->H);
    *o_y1 = take(M * r->dirs * r->H);
    // fp16 (hi, lo) planes of a tensor-core launch's activation operand (2 x 2 B per element): linear1's packed input, or the split of
    // a layer input whose producer could not write the planes itself
    *o_xs = take(M * (size_t)std::max(r->dirs * r->H, r->k1p));
    *total = off;
}

size_t mp_rnn_workspace_bytes(const mp_rnn_t* r, int32_t B, int32_t T) {
    if (!r || B <= 0 || T <= 0) return 0;
    size_t a, b, c, d, e, total;
    rnn_ws_layout(r, (size_t)B * T, &a, &b, &c, &d, &e, &total);
    return total;
}

static int rnn_forward_impl(const mp_rnn_t* r, const float* xa, int32_t ka, const float* xb, int32_t kb, int32_t B, int32_t T,
                            const int32_t* lengths, const float* h0, const float* c0, float* hn, float* cn, float* y,
                            void* workspace, size_t workspace_bytes, cudaStream_t stream, int tile_hint = 0) {
    MP_REQUIRE(r && xa && y && workspace, "rnn_forward: null argument");
    MP_REQUIRE(B > 0 && T > 0, "rnn_forward: empty batch (B=%d T=%d)", B, T);
    MP_REQUIRE(ka + kb == r->n_in, "rnn_forward: input width %d+%d != n_input %d", ka, kb, r->n_in);
    MP_REQUIRE((h0 == nullptr) == (c0 == nullptr), "rnn_forward: h0 and c0 must both be given or both be null");
    MP_REQUIRE(((uintptr_t)workspace & 255) == 0, "rnn_forward: workspace must be 256-byte aligned");
    const size_t M = (size_t)B * T;
    MP_REQUIRE(M * (size_t)r->dirs * 4 * r->H < (size_t)1 << 40, "rnn_forward: batch too large");
    size_t o_x1, o_gin, o_y0, o_y1, o_xs, total;
    rnn_ws_layout(r, M, &o_x1, &o_gin, &o_y0, &o_y1, &o_xs, &total);
    if (workspace_bytes < total) {
        set_error("rnn_forward: workspace %zu < required %zu", workspace_bytes, total);
        return MP_ERR_WORKSPACE;
    }
    char* ws = (char*)workspace;
    float* x1 = (float*)(ws + o_x1);
    float* gin = (float*)(ws + o_gin);
    float* ybuf[2] = {(float*)(ws + o_y0), (float*)(ws + o_y1)};
    const int H = r->H, dirs = r->dirs;
    // which projections run on the fp16-split tensor-core kernel: their activation operand is wanted as (hi, lo) planes of halves,
    // which the producing kernel writes directly where it can (linear1's epilogue, the fp16-split recurrence) -- same bytes as fp32
    const bool proj16[2] = {r->wih_f16[0] && gemm_f16_eligible((int)M, dirs * 4 * H, H),
                            r->wih_f16[1] && gemm_f16_eligible((int)M, dirs * 4 * H, dirs * H)};
    const bool fuse_split = !getenv("MP_NO_FUSED_SPLIT");
    const bool lin2_16 = r->w2_f16 && gemm_f16_eligible((int)M, 256, dirs * H) && !getenv("MP_LINEAR2_FFMA");
    unsigned int* sched = reinterpret_cast<unsigned int*>(ws);
    if (proj16[0] || proj16[1] || lin2_16) MP_CUDA_TRY(cudaMemsetAsync(sched, 0, 256, stream));
    const bool x1_split = proj16[0] && fuse_split;
    // linear1 + ReLU (dropout is the identity in eval)                         rnn.py:22
    const bool lin1_16 = x1_split && r->w1_f16 && gemm_f16_eligible((int)M, 256, r->k1p) && (ka & 3) == 0 && (kb & 3) == 0 &&
                         !getenv("MP_LINEAR1_FFMA");
    if (lin1_16) {
        // tensor-core linear1: the two-source operand is packed into padded (hi, lo) planes, the epilogue applies bias + ReLU and
        // writes the planes the layer-0 projection reads
        MP_TRY(launch_pack_cat_f16(xa, ka, xb, kb, M, r->k1p, ws + o_xs, stream));
        MP_TRY(launch_gemm_f16x3_act(ws + o_xs, r->w1_f16, r->b1_pad, x1, (int)M, 256, r->k1p, H, 3, sched + 6, stream));
    } else {
        MP_TRY(launch_gemm_ffma(xa, ka, xb, kb, r->w1, r->b1, x1, (int)M, H, x1_split ? 3 : 1, stream));
    }
    const float* layer_in = x1;
    bool layer_in_split = x1_split;
    int in_w = H;
    for (int l = 0; l < 2; ++l) {
        // hoisted input projection of both directions                        rnn.py:27 (W_ih x + b_ih + b_hh)
        if (proj16[l]) {
            const void* xs = layer_in;
            if (!layer_in_split) {             // the producer could not write the planes itself: one streaming split pass
                MP_TRY(launch_split_f16(layer_in, M * (size_t)in_w, ws + o_xs, stream));
                xs = ws + o_xs;
            }
            MP_TRY(launch_gemm_f16x3(xs, r->wih_f16[l], r->bsum[l], gin, (int)M, dirs * 4 * H, in_w, dirs * 4 * H, sched + 2 * l, stream));
        } else if (r->wih_split[l] && gemm_tc_eligible((int)M, dirs * 4 * H, in_w) && !getenv("MP_GEMM_NOSPLIT"))
            MP_TRY(launch_gemm_tf32x3_presplit(layer_in, r->wih_split[l], r->bsum[l], gin, (int)M, dirs * 4 * H, in_w, stream));
        else
            MP_TRY(launch_gemm_bias_act(layer_in, in_w, nullptr, 0, r->wih[l], r->bsum[l], gin, (int)M, dirs * 4 * H, 0, stream));
        RecLayerArgs a;
        a.gin = gin; a.wpack = r->whh_pack[l]; a.wT = r->whh_t[l]; a.y = ybuf[l];
        a.w_raw[0] = r->whh_raw[l]; a.w_raw[1] = r->whh_raw[l] + (size_t)(dirs - 1) * 4 * H * H;
        const size_t so = (size_t)l * dirs * B * H;
        a.h0 = h0 ? h0 + so : nullptr; a.c0 = c0 ? c0 + so : nullptr;
        a.hn = hn ? hn + so : nullptr; a.cn = cn ? cn + so : nullptr;
        a.lengths = lengths; a.B = B; a.T = T; a.H = H; a.dirs = dirs;
        a.tile_hint = tile_hint;
        // layer 0 feeds layer 1's projection only: written as (hi, lo) planes when both ends are the fp16-split kernels
        // (layer 1 feeds linear2 only: the same when that runs on the tensor cores)
        a.y_split = (((l == 0 && proj16[1]) || (l == 1 && lin2_16)) && fuse_split && rec_f16_eligible(a)) ? 1 : 0;
        MP_TRY(launch_lstm_recurrence(a, stream));
        layer_in = ybuf[l];
        layer_in_split = a.y_split != 0;
        in_w = dirs * H;
    }
    // linear2                                                                 rnn.py:32
    if (lin2_16) {
        const void* xs = layer_in;
        if (!layer_in_split) {
            MP_TRY(launch_split_f16(layer_in, M * (size_t)in_w, ws + o_xs, stream));
            xs = ws + o_xs;
        }
        MP_TRY(launch_gemm_f16x3(xs, r->w2_f16, r->b2_pad, y, (int)M, 256, in_w, r->n_out, sched + 4, stream));
    } else {
        MP_TRY(launch_gemm_bias_act(layer_in, in_w, nullptr, 0, r->w2, r->b2, y, (int)M, r->n_out, 0, stream));
    }
    return MP_OK;
}

int mp_rnn_forward(const mp_rnn_t* r, const float* xa, int32_t ka, const float* xb, int32_t kb, int32_t B, int32_t T,
                   const int32_t* lengths, const float* h0, const float* c0, float* hn, float* cn, float* y,
                   void* workspace, size_t workspace_bytes, mp_stream_t stream) {
    g_launches = 0;
    return rnn_forward_impl(r, xa, ka, xb, kb, B, T, lengths, h0, c0, hn, cn, y, workspace, workspace_bytes,
                            (cudaStream_t)stream);
}

size_t mp_rnn_train_workspace_bytes(const mp_rnn_weights_t* w, int32_t B, int32_t T) { return rnn_train_workspace_bytes(w, B, T); }
int mp_rnn_train_forward(const mp_rnn_weights_t* w, const float* x, int32_t B, int32_t T, const int32_t* lengths, const float* mask,
                         float* y, void* workspace, size_t workspace_bytes, mp_stream_t stream) {
    g_launches = 0;
    MP_TRY(mp_device_check());
    return rnn_train_forward(w, x, B, T, lengths, mask, y, workspace, workspace_bytes, (cudaStream_t)stream);
}
int mp_rnn_train_backward(const mp_rnn_weights_t* w, const float* x, int32_t B, int32_t T, const int32_t* lengths, const float* mask,
                          const float* dy, const mp_rnn_grads_t* grads, void* workspace, size_t workspace_bytes, mp_stream_t stream) {
    g_launches = 0;
    return rnn_train_backward(w, x, B, T, lengths, mask, dy, grads, workspace, workspace_bytes, (cudaStream_t)stream);
}
int mp_joints_loss(const float* pred, const float* target, int32_t B, int32_t T, int32_t D, float t_weight, double* loss, float* dpred,
                   mp_stream_t stream) {
    return joints_loss(pred, target, B, T, D, t_weight, loss, dpred, (cudaStream_t)stream);
}

int mp_poser_loss(const float* pred, const float* pose_t, const float* joints_t, int32_t B, int32_t T, float t_weight, double* loss,
                  float* dpred, mp_stream_t stream) {
    return poser_loss(pred, pose_t, joints_t, B, T, t_weight, loss, dpred, (cudaStream_t)stream);
}
int mp_footcontact_loss(const float* pred, const float* target, int32_t B, int32_t T, double* loss, float* dpred, mp_stream_t stream) {
    return footcontact_loss(pred, target, B, T, loss, dpred, (cudaStream_t)stream);
}
int mp_velocity_loss(const float* pred, const float* target, int32_t B, int32_t T, int32_t D, double* loss, float* dpred, mp_stream_t stream) {
    return velocity_loss(pred, target, B, T, D, loss, dpred, (cudaStream_t)stream);
}

int mp_grad_sq_norm(const float* grads, size_t n, double* sq_norm, mp_stream_t stream) {
    return grad_sq_norm(grads, n, sq_norm, (cudaStream_t)stream);
}
int mp_adamw_step(float* params, const float* grads, float* exp_avg, float* exp_avg_sq, size_t n, float lr, float beta1, float beta2,
                  float eps, float weight_decay, int32_t step, const double* sq_norm, float max_norm, float grad_scale, mp_stream_t stream) {
    g_launches = 0;
    MP_TRY(mp_device_check());
    return adamw_step(params, grads, exp_avg, exp_avg_sq, n, lr, beta1, beta2, eps, weight_decay, step, sq_norm, max_norm, grad_scale,
                      (cudaStream_t)stream);
}

int mp_gemm_bias(const float* A, const float* W, const float* bias, float* C, int32_t M, int32_t N, int32_t K,
                 int32_t relu, int32_t mode, mp_stream_t stream) {
    g_launches = 0;
    if (mode == 1) return launch_gemm_ffma(A, K, nullptr, 0, W, bias, C, M, N, relu, (cudaStream_t)stream);
    if (mode == 2) {
        MP_REQUIRE(!relu, "gemm_bias: the tensor-core path has no activation");
        return launch_gemm_tf32x3(A, W, bias, C, M, N, K, (cudaStream_t)stream);
    }
    if (mode == 3) {
        // test entry of the fp16-split kernel: both operands are split here into stream-ordered scratch
        MP_REQUIRE(!relu, "gemm_bias: the tensor-core path has no activation");
        cudaStream_t s = (cudaStream_t)stream;
        void *as = nullptr, *wsp = nullptr, *sched = nullptr;
        MP_CUDA_TRY(cudaMallocAsync(&as, (size_t)M * K * 4, s));
        MP_CUDA_TRY(cudaMallocAsync(&wsp, (size_t)N * K * 4, s));
        MP_CUDA_TRY(cudaMallocAsync(&sched, 256, s));
        MP_CUDA_TRY(cudaMemsetAsync(sched, 0, 256, s));
        const bool il = getenv("MP_GEMM_IL") != nullptr;      // experiment: interleaved (hi, lo) operand layout, 128-byte TMA rows
        int st = il ? launch_split_f16_il(A, (size_t)M * K, as, s) : launch_split_f16(A, (size_t)M * K, as, s);
        if (st == MP_OK) st = il ? launch_split_f16_il(W, (size_t)N * K, wsp, s) : launch_split_f16(W, (size_t)N * K, wsp, s);
        if (getenv("MP_GEMM_DBG")) {       // clock stamps of CTA 0's issuing thread land in the last 48 bytes of C's first row... no: own buffer
            static unsigned long long* dbg_buf = nullptr;
            if (!dbg_buf) MP_CUDA_TRY(cudaMallocManaged((void**)&dbg_buf, 64));
            g_gemm_dbg = dbg_buf;
        } else {
            g_gemm_dbg = nullptr;
        }
        const char* pv = getenv("MP_GEMM_PAIR_TEST");       // test entry: 1 = CTA-pair kernel, 0 = single-CTA kernel, unset = default
        const int pf = pv ? (atoi(pv) ? 16 : 32) : 0;
        if (st == MP_OK) st = launch_gemm_f16x3_act(as, wsp, bias, C, M, N, K, N, (il ? 4 : 0) | pf, (unsigned int*)sched, s);
        cudaFreeAsync(as, s);
        cudaFreeAsync(wsp, s);
        cudaFreeAsync(sched, s);
        if (g_gemm_dbg) {
            cudaStreamSynchronize(s);
            fprintf(stderr, "[gemm dbg] M=%d N=%d K=%d flags=%d: CTA0 issuer %llu clk = %llu ns (%.3f GHz), %llu tiles; waiting for operands %llu clk, "
                    "for the epilogue %llu clk, for the tile queue %llu clk\n", M, N, K, (il ? 4 : 0) | pf, g_gemm_dbg[0], g_gemm_dbg[3],
                    (double)g_gemm_dbg[0] / (double)g_gemm_dbg[3], g_gemm_dbg[4], g_gemm_dbg[1], g_gemm_dbg[2], g_gemm_dbg[5]);
            g_gemm_dbg = nullptr;
        }
        return st;
    }
    if (mode == 4) {
        // test entry of the same kernel for narrow outputs (linear2): W and the bias are zero-padded to one 256-row tile first
        MP_REQUIRE(!relu && N <= 256 && (N & 3) == 0 && K % 32 == 0, "gemm_bias mode 4: N <= 256, N %% 4 == 0, K %% 32 == 0, no activation");
        cudaStream_t s = (cudaStream_t)stream;
        void *as = nullptr, *wsp = nullptr, *sched = nullptr;
        float *wpad = nullptr, *bpad = nullptr;
        MP_CUDA_TRY(cudaMallocAsync(&as, (size_t)M * K * 4, s));
        MP_CUDA_TRY(cudaMallocAsync(&wsp, (size_t)256 * K * 4, s));
        MP_CUDA_TRY(cudaMallocAsync((void**)&wpad, (size_t)256 * K * 4, s));
        MP_CUDA_TRY(cudaMallocAsync((void**)&bpad, 256 * 4, s));
        MP_CUDA_TRY(cudaMallocAsync(&sched, 256, s));
        MP_CUDA_TRY(cudaMemsetAsync(sched, 0, 256, s));
        MP_CUDA_TRY(cudaMemsetAsync(wpad, 0, (size_t)256 * K * 4, s));
        MP_CUDA_TRY(cudaMemsetAsync(bpad, 0, 256 * 4, s));
        MP_CUDA_TRY(cudaMemcpyAsync(wpad, W, (size_t)N * K * 4, cudaMemcpyDeviceToDevice, s));
        MP_CUDA_TRY(cudaMemcpyAsync(bpad, bias, (size_t)N * 4, cudaMemcpyDeviceToDevice, s));
        int st = launch_split_f16(A, (size_t)M * K, as, s);
        if (st == MP_OK) st = launch_split_f16(wpad, (size_t)256 * K, wsp, s);
        if (st == MP_OK) st = launch_gemm_f16x3(as, wsp, bpad, C, M, 256, K, N, (unsigned int*)sched, s);
        cudaFreeAsync(as, s); cudaFreeAsync(wsp, s); cudaFreeAsync(wpad, s); cudaFreeAsync(bpad, s); cudaFreeAsync(sched, s);
        return st;
    }
    return launch_gemm_bias_act(A, K, nullptr, 0, W, bias, C, M, N, relu, (cudaStream_t)stream);
}

// ------------------------------------------------------------------------------------------------
int mp_pose_reduced_global_to_full(const float* r6d, int64_t n_frames, float* pose, mp_stream_t stream) {
    return launch_reduced_global_to_full(r6d, n_frames, pose, (cudaStream_t)stream);
}
int mp_tran_offline(const float* joints, const float* vel, const float* contact, const int32_t* lengths, int32_t B,
                    int32_t T, float* tran, mp_stream_t stream) {
    return launch_tran_offline(joints, vel, contact, lengths, B, T, tran, (cudaStream_t)stream);
}
int mp_imu_assemble(const float* acc, const float* ori, int64_t T, int32_t slots_in, const int32_t* combo_masks_host, int32_t n_combos,
                    float acc_scale, int32_t smooth, float* imu_out, mp_stream_t stream) {
    return launch_imu_assemble(acc, ori, T, slots_in, combo_masks_host, n_combos, acc_scale, smooth, imu_out, (cudaStream_t)stream);
}
int mp_imu_live_normalize(const float* quat, const float* acc_raw, int64_t n_ticks, const float* smpl2imu_host, const float* device2bone_host,
                          const float* acc_offsets_host, const int32_t* perm_host, int32_t combo_mask, int32_t phone_as_watch,
                          float acc_scale, float* imu_out, mp_stream_t stream) {
    return launch_imu_live_normalize(quat, acc_raw, n_ticks, smpl2imu_host, device2bone_host, acc_offsets_host, perm_host, combo_mask,
                                     phone_as_watch, acc_scale, imu_out, (cudaStream_t)stream);
}
int mp_physics_optimize(const float* pose, const float* vel, const float* contact, const int32_t* lengths, float* state,
                        int32_t B, int32_t T, const mp_physics_params_t* params, float* pose_out, float* tran_out,
                        mp_stream_t stream) {
    return launch_physics_optimize(pose, vel, contact, lengths, state, B, T, params, pose_out, tran_out, nullptr, -1,
                                   (cudaStream_t)stream);
}
int mp_physics_optimize_debug(const float* pose, const float* vel, const float* contact, const int32_t* lengths,
                              float* state, int32_t B, int32_t T, const mp_physics_params_t* params, float* pose_out,
                              float* tran_out, float* dbg, int32_t dbg_frame, mp_stream_t stream) {
    return launch_physics_optimize(pose, vel, contact, lengths, state, B, T, params, pose_out, tran_out, dbg, dbg_frame,
                                   (cudaStream_t)stream);
}
int mp_eval_frame_errors(const float* pose_p, const float* pose_t, const float* tran_p, const float* tran_t, int64_t n_frames,
                         float* joint_p, float* joint_t, float* je, float* lae, float* gae, mp_stream_t stream) {
    return launch_eval_frame_errors(pose_p, pose_t, tran_p, tran_t, n_frames, joint_p, joint_t, je, lae, gae, (cudaStream_t)stream);
}
int mp_eval_vertex_errors(const float* pose_p, const float* pose_t, int64_t n_frames, const float* rest_vertices, const float* weights,
                           int32_t n_vertices, double* err_sum, double* err_sq_sum, mp_stream_t stream) {
    return launch_eval_vertex_errors(pose_p, pose_t, n_frames, rest_vertices, weights, n_vertices, err_sum, err_sq_sum, (cudaStream_t)stream);
}
int mp_eval_motion_rows(const float* joint_p, const float* joint_t, const float* je, const float* lae, const float* gae, int64_t n_frames,
                        int32_t fps, uint32_t joint_mask_bits, float* rows, mp_stream_t stream) {
    return launch_eval_motion_rows(joint_p, joint_t, je, lae, gae, n_frames, fps, joint_mask_bits, rows, (cudaStream_t)stream);
}
int mp_eval_motion_rows_batch(const float* joint_p, const float* joint_t, const float* je, const float* lae, const float* gae,
                              const int64_t* offsets, int32_t n_sequences, int32_t fps, uint32_t joint_mask_bits, float* rows, mp_stream_t stream) {
    static_assert(sizeof(long long) == sizeof(int64_t), "offsets are 64-bit");
    return launch_eval_motion_rows_batch(joint_p, joint_t, je, lae, gae, reinterpret_cast<const long long*>(offsets), n_sequences, fps,
                                         joint_mask_bits, rows, (cudaStream_t)stream);
}
int mp_eval_tran_windows(const float* tran_p, const float* tran_t, const int32_t* lengths, int32_t S, int32_t T, float* err,
                         int32_t* count, mp_stream_t stream) {
    return launch_eval_tran_windows(tran_p, tran_t, lengths, S, T, err, count, (cudaStream_t)stream);
}
int mp_physics_fk(const float* pose, int64_t n_frames, float* global_rot, float* joint_pos, mp_stream_t stream) {
    return launch_physics_fk(pose, n_frames, global_rot, joint_pos, (cudaStream_t)stream);
}
int mp_online_update(mp_online_state_t* state, const float* pose, const float* joints, const float* vel,
                     const float* contact, int32_t S, int32_t W, int32_t frame_idx, float* pose_out, float* root_out,
                     float* contact_out, mp_stream_t stream) {
    return launch_online_update(state, pose, joints, vel, contact, S, W, frame_idx, pose_out, root_out, contact_out,
                                (cudaStream_t)stream);
}
int mp_online_push_frame(const float* win_in, float* win_out, const float* frame, int32_t S, int32_t W, int32_t cold,
                         mp_stream_t stream) {
    MP_REQUIRE(win_in != win_out, "online push: in-place shift is not supported");
    return launch_online_push(win_in, win_out, frame, S, W, cold, (cudaStream_t)stream);
}
int mp_online_reset(mp_online_state_t* state, int32_t S, int32_t full, mp_stream_t stream) {
    return launch_online_reset(state, S, full, (cudaStream_t)stream);
}

// ------------------------------------------------------------------------------------------------
int mp_net_create(mp_net_t** out, const mp_rnn_t* joints, const mp_rnn_t* pose, const mp_rnn_t* foot,
                  const mp_rnn_t* velocity) {
    MP_REQUIRE(out && joints && pose && foot && velocity, "net_create: null argument");
    MP_REQUIRE(joints->n_in == 60 && joints->n_out == 72, "net_create: joints head must be 60 -> 72");
    MP_REQUIRE(pose->n_in == 132 && pose->n_out == 96, "net_create: pose head must be 132 -> 96");
    MP_REQUIRE(foot->n_in == 132 && foot->n_out == 2, "net_create: foot_contact head must be 132 -> 2");
    MP_REQUIRE(velocity->n_in == 132 && velocity->n_out == 72 && velocity->dirs == 1, "net_create: velocity head must be a unidirectional 132 -> 72");
    MP_TRY(mp_device_check());
    mp_net* n = new mp_net();
    n->joints = joints; n->pose = pose; n->foot = foot; n->vel = velocity;
    cudaError_t e = cudaSuccess;
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&n->s_foot, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&n->s_vel, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&n->s_cap, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&n->ev_joints, cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&n->ev_foot, cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&n->ev_vel, cudaEventDisableTiming);
    if (e != cudaSuccess) {
        set_error("net_create: %s", cudaGetErrorString(e));
        mp_net_destroy(n);
        return MP_ERR_CUDA;
    }
    *out = n;
    return MP_OK;
}

void mp_net_destroy(mp_net_t* n) {
    if (!n) return;
    for (auto& en : n->cache)
        if (en.exec) cudaGraphExecDestroy(en.exec);
    if (n->ev_joints) cudaEventDestroy(n->ev_joints);
    if (n->ev_foot) cudaEventDestroy(n->ev_foot);
    if (n->ev_vel) cudaEventDestroy(n->ev_vel);
    if (n->s_foot) cudaStreamDestroy(n->s_foot);
    if (n->s_vel) cudaStreamDestroy(n->s_vel);
    if (n->s_cap) cudaStreamDestroy(n->s_cap);
    delete n;
}

int mp_net_set_physics(mp_net_t* n, const mp_physics_params_t* params) {
    MP_REQUIRE(n, "net_set_physics: null");
    if (!params) {
        n->physics_on = 0;
        return MP_OK;
    }
    MP_TRY(physics_prepare());
    n->phys = *params;
    n->physics_on = 1;
    ++n->physics_epoch;
    return MP_OK;
}

int mp_net_set_rec_tile(mp_net_t* n, int32_t sequences_per_tile) {
    MP_REQUIRE(n && sequences_per_tile >= -1 && (sequences_per_tile <= 64 || sequences_per_tile == 128),
               "net_set_rec_tile: -1 (one wave), 0 (auto), 1 .. 64, or 128 (the four-sub-tile kernel where the batch holds whole tiles)");
    n->rec_tile = sequences_per_tile;
    return MP_OK;
}

int mp_net_set_graph(mp_net_t* n, int32_t enabled) {
    MP_REQUIRE(n, "net_set_graph: null");
    n->graph_enabled = enabled ? 1 : 0;
    return MP_OK;
}

static void net_ws_layout(const mp_net* n, int B, int T, size_t off[6], size_t* total) {
    size_t o = 0;
    const mp_rnn* heads[4] = {n->joints, n->pose, n->foot, n->vel};
    for (int i = 0; i < 4; ++i) {
        off[i] = o;
        o = align_up(o + mp_rnn_workspace_bytes(heads[i], B, T));
    }
    off[4] = o;   // r6d [B*T, 96]
    o = align_up(o + (size_t)B * T * 96 * sizeof(float));
    off[5] = o;   // K8 state [B, MP_PHYSICS_STATE_FLOATS]
    o = align_up(o + (size_t)B * MP_PHYSICS_STATE_FLOATS * sizeof(float));
    *total = o;
}

size_t mp_net_workspace_bytes(const mp_net_t* n, int32_t B, int32_t T) {
    if (!n || B <= 0 || T <= 0) return 0;
    size_t off[6], total;
    net_ws_layout(n, B, T, off, &total);
    return total;
}

struct NetArgs {
    const float* imu; int B, T; const int32_t* lengths;
    const float *vel_h0, *vel_c0; float *vel_hn, *vel_cn;
    float *pose, *joints, *vel, *contact, *tran;
    void* ws; size_t ws_bytes;
};

// Enqueue the whole forward: joints on `s`, then pose(+K5) on `s`, foot_contact and velocity on the
// net's side streams (forked after joints, joined before K6).          net.py:101-119,125-154
static int net_enqueue(mp_net* n, const NetArgs& a, cudaStream_t s) {
    size_t off[6], total;
    net_ws_layout(n, a.B, a.T, off, &total);
    char* ws = (char*)a.ws;
    auto wsz = [&](int i) { return (i < 3 ? off[i + 1] : off[4]) - off[i]; };
    float* r6d = (float*)(ws + off[4]);
    MP_TRY(rnn_forward_impl(n->joints, a.imu, 60, nullptr, 0, a.B, a.T, a.lengths, nullptr, nullptr, nullptr, nullptr,
                          a.joints, ws + off[0], wsz(0), s, n->rec_tile));
    MP_CUDA_TRY(cudaEventRecord(n->ev_joints, s));
    MP_CUDA_TRY(cudaStreamWaitEvent(n->s_foot, n->ev_joints, 0));
    MP_CUDA_TRY(cudaStreamWaitEvent(n->s_vel, n->ev_joints, 0));
    // velocity first on its own stream: it is the longest side branch (unidirectional, 2 x T steps)
    MP_TRY(rnn_forward_impl(n->vel, a.joints, 72, a.imu, 60, a.B, a.T, a.lengths, a.vel_h0, a.vel_c0, a.vel_hn, a.vel_cn,
                          a.vel, ws + off[3], wsz(3), n->s_vel, n->rec_tile));
    MP_CUDA_TRY(cudaEventRecord(n->ev_vel, n->s_vel));
    MP_TRY(rnn_forward_impl(n->foot, a.joints, 72, a.imu, 60, a.B, a.T, a.lengths, nullptr, nullptr, nullptr, nullptr,
                          a.contact, ws + off[2], wsz(2), n->s_foot, n->rec_tile));
    MP_CUDA_TRY(cudaEventRecord(n->ev_foot, n->s_foot));
    MP_TRY(rnn_forward_impl(n->pose, a.joints, 72, a.imu, 60, a.B, a.T, a.lengths, nullptr, nullptr, nullptr, nullptr, r6d,
                          ws + off[1], wsz(1), s, n->rec_tile));
    MP_TRY(launch_reduced_global_to_full(r6d, (int64_t)a.B * a.T, a.pose, s));
    MP_CUDA_TRY(cudaStreamWaitEvent(s, n->ev_foot, 0));
    MP_CUDA_TRY(cudaStreamWaitEvent(s, n->ev_vel, 0));
    if (a.tran) MP_TRY(launch_tran_offline(a.joints, a.vel, a.contact, a.lengths, a.B, a.T, a.tran, s));
    if (a.tran && n->physics_on) {
        // net.py:157-169: the PHYSICS hook rewrites the pose in place (its translation is discarded); every sequence
        // of the batch starts from a fresh optimizer state
        float* st = (float*)(ws + off[5]);
        MP_CUDA_TRY(cudaMemsetAsync(st, 0, (size_t)a.B * MP_PHYSICS_STATE_FLOATS * sizeof(float), s));
        MP_TRY(launch_physics_optimize(a.pose, a.vel, a.contact, a.lengths, st, a.B, a.T, &n->phys, a.pose, nullptr, nullptr, -1, s));
    }
    return MP_OK;
}

int mp_net_forward(mp_net_t* n, const float* imu, int32_t B, int32_t T, const int32_t* lengths, const float* vel_h0,
                   const float* vel_c0, float* vel_hn, float* vel_cn, float* pose, float* joints, float* vel,
                   float* contact, float* tran, void* workspace, size_t workspace_bytes, mp_stream_t stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    MP_REQUIRE(n && imu && pose && joints && vel && contact && workspace, "net_forward: null argument");
    MP_REQUIRE(B > 0 && T > 0, "net_forward: empty batch (B=%d T=%d)", B, T);
    MP_REQUIRE((vel_h0 == nullptr) == (vel_c0 == nullptr), "net_forward: velocity h0/c0 must come together");
    const size_t need = mp_net_workspace_bytes(n, B, T);
    if (workspace_bytes < need) {
        set_error("net_forward: workspace %zu < required %zu", workspace_bytes, need);
        return MP_ERR_WORKSPACE;
    }
    NetArgs a{imu, B, T, lengths, vel_h0, vel_c0, vel_hn, vel_cn, pose, joints, vel, contact, tran, workspace, workspace_bytes};
    g_launches = 0;
    if (!n->graph_enabled || profile_enabled()) return net_enqueue(n, a, stream);

    std::vector<uintptr_t> key = {(uintptr_t)imu, (uintptr_t)B, (uintptr_t)T, (uintptr_t)lengths, (uintptr_t)vel_h0,
                                  (uintptr_t)vel_c0, (uintptr_t)vel_hn, (uintptr_t)vel_cn, (uintptr_t)pose,
                                  (uintptr_t)joints, (uintptr_t)vel, (uintptr_t)contact, (uintptr_t)tran,
                                  (uintptr_t)workspace, (uintptr_t)(n->physics_on ? n->physics_epoch : 0), (uintptr_t)n->rec_tile};
    mp_net::Entry* hit = nullptr;
    for (auto& en : n->cache)
        if (en.key == key) hit = &en;
    if (!hit) {
        // first sight of this signature: run eagerly (also loads the kernels, which must not happen
        // inside a capture) and remember it; the next call with the same signature is captured.
        if (n->cache.size() >= 32) {
            size_t victim = 0;
            for (size_t i = 1; i < n->cache.size(); ++i)
                if (n->cache[i].stamp < n->cache[victim].stamp) victim = i;
            if (n->cache[victim].exec) cudaGraphExecDestroy(n->cache[victim].exec);
            n->cache.erase(n->cache.begin() + victim);
        }
        mp_net::Entry en;
        en.key = key;
        en.stamp = ++n->clock;
        n->cache.push_back(en);
        return net_enqueue(n, a, stream);
    }
    hit->stamp = ++n->clock;
    if (!hit->exec) {
        cudaGraph_t graph = nullptr;
        MP_CUDA_TRY(cudaStreamBeginCapture(n->s_cap, cudaStreamCaptureModeThreadLocal));
        int st = net_enqueue(n, a, n->s_cap);
        cudaError_t e = cudaStreamEndCapture(n->s_cap, &graph);
        if (st != MP_OK) {
            if (graph) cudaGraphDestroy(graph);
            return st;
        }
        if (e != cudaSuccess) {
            set_error("net_forward: graph capture failed: %s", cudaGetErrorString(e));
            return MP_ERR_CUDA;
        }
        e = cudaGraphInstantiate(&hit->exec, graph, 0);
        cudaGraphDestroy(graph);
        if (e != cudaSuccess) {
            hit->exec = nullptr;
            set_error("net_forward: graph instantiate failed: %s", cudaGetErrorString(e));
            return MP_ERR_CUDA;
        }
        hit->kernels = g_launches;
    }
    MP_CUDA_TRY(cudaGraphLaunch(hit->exec, stream));
    g_launches = hit->kernels;
    return MP_OK;
}

// staging layout: imu | lengths | pose | joints | tran | contact | vel | pose as local 6D (compact transfers)
static void host_staging_layout(size_t B, size_t T, size_t off[8], size_t* total) {
    size_t o = 0;
    const size_t sz[8] = {B * T * 60 * 4, B * 4, B * T * 216 * 4, B * T * 72 * 4, B * T * 3 * 4, B * T * 2 * 4, B * T * 72 * 4, B * T * 96 * 4};
    for (int i = 0; i < 8; ++i) {
        off[i] = o;
        o = align_up(o + sz[i]);
    }
    *total = o;
}

size_t mp_net_host_staging_bytes(int32_t B, int32_t T) {
    if (B <= 0 || T <= 0) return 0;
    size_t off[8], total;
    host_staging_layout(B, T, off, &total);
    return total;
}

static int enqueue_offline_host_impl(mp_net_t* n, const float* imu_host, int32_t B, int32_t T, const int32_t* lengths_host,
                                     float* pose_host, float* joints_host, float* tran_host, float* contact_host, void* dev_io,
                                     void* workspace, size_t workspace_bytes, mp_stream_t stream_, bool compact) {
    cudaStream_t stream = (cudaStream_t)stream_;
    MP_REQUIRE(n && imu_host && dev_io && pose_host && joints_host && tran_host && contact_host, "forward_offline_host: null argument");
    MP_REQUIRE(B > 0 && T > 0, "forward_offline_host: empty batch");
    MP_REQUIRE(((uintptr_t)dev_io & 255) == 0, "forward_offline_host: staging must be 256-byte aligned");
    size_t off[8], total;
    host_staging_layout(B, T, off, &total);
    char* d = (char*)dev_io;
    float* d_imu = (float*)(d + off[0]);
    int32_t* d_len = (int32_t*)(d + off[1]);
    float *d_pose = (float*)(d + off[2]), *d_joints = (float*)(d + off[3]), *d_tran = (float*)(d + off[4]),
          *d_contact = (float*)(d + off[5]), *d_vel = (float*)(d + off[6]), *d_pose6 = (float*)(d + off[7]);
    const size_t F = (size_t)B * T;
    MP_CUDA_TRY(cudaMemcpyAsync(d_imu, imu_host, F * 60 * 4, cudaMemcpyHostToDevice, stream));
    if (lengths_host) MP_CUDA_TRY(cudaMemcpyAsync(d_len, lengths_host, (size_t)B * 4, cudaMemcpyHostToDevice, stream));
    MP_TRY(mp_net_forward(n, d_imu, B, T, lengths_host ? d_len : nullptr, nullptr, nullptr, nullptr, nullptr, d_pose,
                          d_joints, d_vel, d_contact, d_tran, workspace, workspace_bytes, stream));
    if (compact) {
        MP_TRY(launch_pose_local6d(d_pose, (int64_t)F, d_pose6, stream));
        MP_CUDA_TRY(cudaMemcpyAsync(pose_host, d_pose6, F * 96 * 4, cudaMemcpyDeviceToHost, stream));
    } else {
        MP_CUDA_TRY(cudaMemcpyAsync(pose_host, d_pose, F * 216 * 4, cudaMemcpyDeviceToHost, stream));
    }
    MP_CUDA_TRY(cudaMemcpyAsync(joints_host, d_joints, F * 72 * 4, cudaMemcpyDeviceToHost, stream));
    MP_CUDA_TRY(cudaMemcpyAsync(tran_host, d_tran, F * 3 * 4, cudaMemcpyDeviceToHost, stream));
    MP_CUDA_TRY(cudaMemcpyAsync(contact_host, d_contact, F * 2 * 4, cudaMemcpyDeviceToHost, stream));
    return MP_OK;
}

int mp_net_enqueue_offline_host(mp_net_t* n, const float* imu_host, int32_t B, int32_t T, const int32_t* lengths_host,
                                float* pose_host, float* joints_host, float* tran_host, float* contact_host, void* dev_io,
                                void* workspace, size_t workspace_bytes, mp_stream_t stream_) {
    return enqueue_offline_host_impl(n, imu_host, B, T, lengths_host, pose_host, joints_host, tran_host, contact_host, dev_io, workspace,
                                     workspace_bytes, stream_, false);
}

int mp_net_enqueue_offline_host_compact(mp_net_t* n, const float* imu_host, int32_t B, int32_t T, const int32_t* lengths_host,
                                        float* pose6d_host, float* joints_host, float* tran_host, float* contact_host, void* dev_io,
                                        void* workspace, size_t workspace_bytes, mp_stream_t stream_) {
    return enqueue_offline_host_impl(n, imu_host, B, T, lengths_host, pose6d_host, joints_host, tran_host, contact_host, dev_io, workspace,
                                     workspace_bytes, stream_, true);
}

int mp_pose_full_to_local6d(const float* pose, int64_t n_frames, float* pose6d, mp_stream_t stream) {
    return launch_pose_local6d(pose, n_frames, pose6d, (cudaStream_t)stream);
}

int mp_net_forward_offline_host(mp_net_t* n, const float* imu_host, int32_t B, int32_t T, const int32_t* lengths_host,
                                float* pose_host, float* joints_host, float* tran_host, float* contact_host, void* dev_io,
                                void* workspace, size_t workspace_bytes, mp_stream_t stream_) {
    MP_TRY(mp_net_enqueue_offline_host(n, imu_host, B, T, lengths_host, pose_host, joints_host, tran_host, contact_host, dev_io,
                                       workspace, workspace_bytes, stream_));
    MP_CUDA_TRY(cudaStreamSynchronize((cudaStream_t)stream_));
    return MP_OK;
}

}  // extern "C"
