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

}  // namespace

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
