// Input assembly: raw per-slot IMU streams -> the 60-float frames the heads consume, for any
// number of device combos in one launch.
//
// Reference (relative to /root/reference/mobileposer):
//   data.py:60-61,69-76     acc[:, :5] / amass.acc_scale, ori[:, :5]; per combo zero the slots that are not worn and
//                           cat(acc.flatten(1) [15], ori.flatten(1) [45]) -> [T, 60]
//   loader.py:39-49         the same masking followed by smooth_avg over the scaled accelerations
//   utils/model_utils.py:28-37  smooth_avg: 3-tap moving average, nanmean over the taps that exist (borders use 2)
// Pure streaming: 240 B read per frame (once per combo, out of L2 after the first) and 240 B written per (combo, frame);
// one thread per output float, consecutive threads write consecutive addresses.
#include "mp_common.cuh"

namespace mp {

namespace {

constexpr int MAX_COMBOS = 16;
struct ComboMasks {
    int32_t m[MAX_COMBOS];
};

__global__ void __launch_bounds__(256) imu_assemble_kernel(const float* __restrict__ acc, const float* __restrict__ ori, long long T,
                                                           int slots_in, ComboMasks masks, int n_combos, float acc_scale, int smooth,
                                                           float* __restrict__ out) {
    const long long total = (long long)n_combos * T * 60;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int k = (int)(i % 60);
        const long long t = (i / 60) % T;
        const int c = (int)(i / (60 * T));
        const int mask = masks.m[c];
        float v = 0.f;
        if (k < 15) {
            const int slot = k / 3, ax = k - 3 * slot;
            if ((mask >> slot) & 1) {
                const float* a = acc + (size_t)slot * 3 + ax;
                const size_t stride = (size_t)slots_in * 3;
                // the division is the reference's `acc / amass.acc_scale` (not a multiplication by the reciprocal)
                const float cur = __ldg(a + (size_t)t * stride) / acc_scale;
                if (smooth) {
                    // nanmean over the existing taps: taps are added in order (t-1, t, t+1), then ONE division by their count
                    float sum = 0.f;
                    int cnt = 1;
                    if (t > 0) { sum = __ldg(a + (size_t)(t - 1) * stride) / acc_scale; ++cnt; }
                    sum = __fadd_rn(sum, cur);
                    if (t + 1 < T) { sum = __fadd_rn(sum, __ldg(a + (size_t)(t + 1) * stride) / acc_scale); ++cnt; }
                    v = sum / (float)cnt;
                } else {
                    v = cur;
                }
            }
        } else {
            const int kk = k - 15, slot = kk / 9, e = kk - 9 * slot;
            if ((mask >> slot) & 1) v = __ldg(ori + (size_t)t * slots_in * 9 + (size_t)slot * 9 + e);
        }
        out[i] = v;
    }
}

// Live-demo normalisation (live_demo.py:210-234) of one tick (or a buffer of n ticks): per sensor, quaternion -> rotation
// (articulate/math/angular.py:224-236, normalising the quaternion), calibration into the SMPL frame
// (glb_acc = smpl2imu acc - acc_offset, glb_ori = smpl2imu R device2bone), the reference's slot permutation [1,4,3,0,2],
// acc / acc_scale, the device-combo mask (or the phone-as-watch variant), and the cat into the 60-float frame -- a dozen
// small torch ops per tick in the reference, one launch here.  One thread per (tick, output slot).
struct LiveCal {
    float s2i[9];
    float d2b[5][9];
    float off[5][3];
    int perm[5];
};

__global__ void __launch_bounds__(128) imu_live_normalize_kernel(const float* __restrict__ quat, const float* __restrict__ acc_raw,
                                                                 long long n, LiveCal cal, int mask, int phone_as_watch,
                                                                 float acc_scale, float* __restrict__ out) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n * 5) return;
    const long long f = i / 5;
    const int slot = (int)(i - f * 5);
    float* oa = out + f * 60 + slot * 3;
    float* oo = out + f * 60 + 15 + slot * 9;
    // which permuted slot feeds this output slot (phone-as-watch: slot 0 <- permuted slot 3, everything else zero)
    const int src_slot = phone_as_watch ? (slot == 0 ? 3 : -1) : (((mask >> slot) & 1) ? slot : -1);
    if (src_slot < 0) {
#pragma unroll
        for (int e = 0; e < 3; ++e) oa[e] = 0.f;
#pragma unroll
        for (int e = 0; e < 9; ++e) oo[e] = 0.f;
        return;
    }
    const int p = cal.perm[src_slot];                     // physical sensor
    const float* qp = quat + (f * 5 + p) * 4;
    float a = qp[0], b = qp[1], c = qp[2], d = qp[3];
    const float nrm = sqrtf(a * a + b * b + c * c + d * d);
    a /= nrm; b /= nrm; c /= nrm; d /= nrm;
    const float R[9] = {-2 * c * c - 2 * d * d + 1, 2 * b * c - 2 * a * d, 2 * a * c + 2 * b * d,
                        2 * b * c + 2 * a * d, -2 * b * b - 2 * d * d + 1, 2 * c * d - 2 * a * b,
                        2 * b * d - 2 * a * c, 2 * a * b + 2 * c * d, -2 * b * b - 2 * c * c + 1};
    const float* ap = acc_raw + (f * 5 + p) * 3;
    const float ax = ap[0], ay = ap[1], az = ap[2];
#pragma unroll
    for (int r = 0; r < 3; ++r)
        oa[r] = (fmaf(cal.s2i[r * 3 + 2], az, fmaf(cal.s2i[r * 3 + 1], ay, cal.s2i[r * 3] * ax)) - cal.off[p][r]) / acc_scale;
    float SR[9];                                          // smpl2imu * R
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
        for (int cc = 0; cc < 3; ++cc)
            SR[r * 3 + cc] = fmaf(cal.s2i[r * 3 + 2], R[6 + cc], fmaf(cal.s2i[r * 3 + 1], R[3 + cc], cal.s2i[r * 3] * R[cc]));
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
        for (int cc = 0; cc < 3; ++cc)
            oo[r * 3 + cc] = fmaf(SR[r * 3 + 2], cal.d2b[p][6 + cc], fmaf(SR[r * 3 + 1], cal.d2b[p][3 + cc], SR[r * 3] * cal.d2b[p][cc]));
}

}  // namespace

int launch_imu_live_normalize(const float* quat, const float* acc_raw, int64_t n, const float* smpl2imu_host,
                              const float* device2bone_host, const float* acc_offsets_host, const int32_t* perm_host,
                              int32_t combo_mask, int phone_as_watch, float acc_scale, float* out, cudaStream_t stream) {
    MP_REQUIRE(quat && acc_raw && out && smpl2imu_host && device2bone_host && acc_offsets_host && perm_host, "imu_live_normalize: null pointer");
    MP_REQUIRE(n > 0 && acc_scale != 0.f && (combo_mask & ~31) == 0, "imu_live_normalize: n=%lld acc_scale=%g mask=0x%x", (long long)n,
               (double)acc_scale, combo_mask);
    LiveCal cal;
    for (int i = 0; i < 9; ++i) cal.s2i[i] = smpl2imu_host[i];
    for (int s = 0; s < 5; ++s) {
        MP_REQUIRE(perm_host[s] >= 0 && perm_host[s] < 5, "imu_live_normalize: perm[%d] = %d", s, perm_host[s]);
        cal.perm[s] = perm_host[s];
        for (int i = 0; i < 9; ++i) cal.d2b[s][i] = device2bone_host[s * 9 + i];
        for (int i = 0; i < 3; ++i) cal.off[s][i] = acc_offsets_host[s * 3 + i];
    }
    const long long threads = n * 5;
    imu_live_normalize_kernel<<<(unsigned)((threads + 127) / 128), 128, 0, stream>>>(quat, acc_raw, n, cal, combo_mask, phone_as_watch,
                                                                                    acc_scale, out);
    MP_CUDA_TRY(cudaGetLastError());
    count_launch();
    return MP_OK;
}

int launch_imu_assemble(const float* acc, const float* ori, int64_t T, int slots_in, const int32_t* masks_host, int n_combos,
                        float acc_scale, int smooth, float* out, cudaStream_t stream) {
    MP_REQUIRE(acc && ori && out && masks_host, "imu_assemble: null pointer");
    MP_REQUIRE(T > 0 && slots_in >= 5 && n_combos >= 1 && n_combos <= MAX_COMBOS, "imu_assemble: T=%lld slots=%d combos=%d (1..%d)",
               (long long)T, slots_in, n_combos, MAX_COMBOS);
    MP_REQUIRE(acc_scale != 0.f, "imu_assemble: acc_scale must not be zero");
    ComboMasks m = {};
    for (int i = 0; i < n_combos; ++i) {
        MP_REQUIRE((masks_host[i] & ~31) == 0, "imu_assemble: combo %d uses slots outside 0..4 (mask 0x%x)", i, masks_host[i]);
        m.m[i] = masks_host[i];
    }
    const long long total = (long long)n_combos * T * 60;
    ProfileScope prof("n2_imu_assemble", 4.0 * ((double)T * 60 + (double)total), stream);
    const int blocks = (int)std::min<long long>((total + 255) / 256, 148 * 16);
    imu_assemble_kernel<<<blocks, 256, 0, stream>>>(acc, ori, T, slots_in, m, n_combos, acc_scale, smooth, out);
    MP_CUDA_TRY(cudaGetLastError());
    count_launch();
    return MP_OK;
}

}  // namespace mp
