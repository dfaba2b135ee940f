// Tensor-core LSTM recurrence, second generation (H = 256, throughput path): the per-step product
// G[4H/C x NB] = W_hh_slice . h^T on tcgen05 with fp32-grade accuracy from FP16 operands ("3xFP16"), W_hh resident in
// TENSOR MEMORY for the whole sequence.  Replaces the time loop of _VF.lstm (mobileposer/models/rnn.py:27).
//
// Same contract and cluster decomposition as lstm_rec_tc.cu (8-CTA cluster per tile of NB sequences, CTA `rank` owns hidden
// units [32 rank, 32 rank + 32) = 128 gate rows, h_t replicated in every CTA and exchanged through distributed shared memory).
// What changes against the TF32 kernel:
//   * operands are fp16 hi / scaled-lo pairs (x = hi + lo * 2^-11, hi = fp16(x), lo = fp16((x - hi) * 2^11): 22 significant
//     bits like the TF32 pair, see gemm_f16.cu for the range argument -- W_hh is checked at pack time, |h| < 1).  kind::f16 takes
//     K = 16 per instruction: 48 tcgen05.mma per sub-tile and step instead of 96, the same ~27 clk each.
//   * the whole W_hh slice (hi AND lo, 128 rows x 256 k x 2 x 2 B = 128 KB) lives in tensor memory columns [0, 256); nothing of
//     W is in shared memory.  h (hi + lo) is 1 KB per sequence instead of 2 KB, so the exchange moves half the bytes:
//     ONE cp.async.bulk per (destination CTA, sub-tile) carrying the hi and lo slices together.
//   * the epilogue warps report a staged sub-tile with one mbarrier arrival per WARP (16 per sub-tile) instead of one per thread
//     per 16-sequence block (2048 arrivals on one barrier word per step).
//   operands     A = W_hh slice [128 rows (row m = unit_local*4 + gate) x K = 256] from TMEM: W_hi columns [0, 128), W_lo (scaled)
//                    columns [128, 256), two halves per 32-bit column
//                B = h_t [rows x K] hi / lo, K-major with 64-byte rows (32 halves = exactly the slice rank r produces), 64B swizzle:
//                    per sub-tile  [K-block r = source rank][hi | lo][rows x 64 B]
//                D = two fp32 accumulators in TMEM: main (W_hi h_hi) columns [256, 256+N), correction (W_lo h_hi + W_hi h_lo,
//                    scaled by 2^11) columns [256+N, 256+2N)
#include "mp_common.cuh"

#include <cuda_fp16.h>

#include <cstdlib>
#include <cstring>

namespace mp {

namespace {

constexpr int TH = 256;
constexpr int TCC = 8;            // cluster size
constexpr int TUC = TH / TCC;     // 32 units per CTA
constexpr int EPI_WARPS = 16;
constexpr int RF_THREADS = (EPI_WARPS + 3) * 32;   // 16 epilogue warps + MMA warp + one exchange warp per sub-tile
constexpr uint32_t COL_WHI = 0, COL_WLO = 128, COL_D = 256;
constexpr float kLoScale = 2048.0f, kLoInv = 1.0f / 2048.0f;

struct RecF16Params {
    const float* gin;
    const float* w0;      // raw torch W_hh [4H, H] of direction 0 / 1
    const float* w1;
    float* y;
    const float* h0;
    const float* c0;
    float* hn;
    float* cn;
    const int32_t* lengths;
    int B, T, dirs, NB;
    int y_split;          // y = two planes of halves (hi, scaled lo) instead of fp32
    long long* ts;        // bring-up: per-step clock64 stamps of block (0,0) [step][8], or null (MP_RTC_TS)
};

__host__ __device__ inline size_t rec_f16_smem_bytes(int N) {
    // h (hi + lo, 8 K-blocks) | staging [2 parities][hi | lo] | tables | barriers
    return 1024 + (size_t)8 * 2 * N * 64 + (size_t)2 * 2 * N * 64 + (size_t)3 * N * 4 + 128;
}

__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
// D[tmem] (+)= A[tmem] * B[smem desc], fp16 operands, fp32 accumulate
__device__ __forceinline__ void umma_ts_f16(uint32_t d, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d),
        "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(acc)
        : "memory");
}
// K-major operand, 64-byte rows, SWIZZLE_64B: 8-row groups 512 B apart
__device__ __forceinline__ uint64_t umma_desc_sw64(uint32_t smem_addr) {
    return (uint64_t)((smem_addr & 0x3FFFF) >> 4) | (1ull << 16) | ((uint64_t)(512 >> 4) << 32) | (1ull << 46) | (4ull << 61);
}
// byte offset of half `k` (0..31) of row `row` inside one [rows x 64 B] K-block with the 64B swizzle (16-byte chunk ^= (row / 2) % 4)
__device__ __forceinline__ uint32_t sw64_off(int row, int k) {
    return (uint32_t)row * 64u + (uint32_t)((((k >> 3) ^ ((row >> 1) & 3)) << 4) | ((k & 7) << 1));
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&v)[16]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
        "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(taddr),
        "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]),
        "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15])
        : "memory");
}
__device__ __forceinline__ void tmem_ld4(uint32_t taddr, float* v) {
    uint32_t a, b, c, d;
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(a), "=r"(b), "=r"(c), "=r"(d) : "r"(taddr) : "memory");
    v[0] = __uint_as_float(a); v[1] = __uint_as_float(b); v[2] = __uint_as_float(c); v[3] = __uint_as_float(d);
}
// Remote arrive WITHOUT release semantics: it only tells the peers that this CTA's tensor core has finished reading its h rows (a
// fact this thread learned through the acquire of its own `mma` barrier wait); no memory written by this thread has to become
// visible with it.  The default .release.cluster form compiles to MEMBAR.ALL.CTA + MEMBAR.ALL.GPU + ERRBAR + CGAERRBAR in front of
// the arrive -- a GPU-scope fence per sub-tile and step on the exchange's critical path (measured: 2 of 6 k clk per step).
__device__ __forceinline__ void mbar_arrive_remote(uint32_t remote_bar) {
    asm volatile("mbarrier.arrive.relaxed.cluster.shared::cluster.b64 _, [%0];" ::"r"(remote_bar) : "memory");
}
__device__ __forceinline__ void mbar_arrive_local(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile("{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\tselp.u32 %0, 1, 0, P;\n\t}" : "=r"(pred));
    return pred != 0;
}
// x -> (fp16 hi, fp16 scaled lo) as raw 16-bit patterns
__device__ __forceinline__ void split16(float x, unsigned short& hi, unsigned short& lo) {
    const __half h = __float2half_rn(x);
    hi = __half_as_ushort(h);
    lo = __half_as_ushort(__float2half_rn((x - __half2float(h)) * kLoScale));
}

__device__ __forceinline__ uint32_t pk(unsigned short a, unsigned short b) { return (uint32_t)a | ((uint32_t)b << 16); }

// short form: sigma(s) = 1 / (1 + 2^(-s log2 e)) with MUFU ex2 + MUFU rcp; tanh(x) = 2 sigma(2x) - 1.
// ex2.approx is 2 ulp on e, rcp.approx 1 ulp on the quotient: <= 1.5e-7 absolute on the result.  5 instructions against ~20 of the
// expf form below.
__device__ __forceinline__ float act_fast(float x, bool is_tanh) {
    // no clamp, no Newton step: e = +inf gives rcp(inf) = 0, the correct limit; rcp.approx is 1 ulp (5 instructions instead of 8, the
    // epilogue is instruction-bound; same form as lstm_rec_f16w.cu -- the two kernels agree bit for bit)
    float e, r;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(x * (is_tanh ? -2.8853900817779268f : -1.4426950408889634f)));
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(1.0f + e));
    return is_tanh ? fmaf(2.0f, r, -1.0f) : r;
}

__device__ __forceinline__ float act_sigmoid_or_tanh(float x, bool is_tanh) {
    const float s = fminf(fmaxf(is_tanh ? 2.0f * x : x, -30.0f), 30.0f);
    const float e = expf(-s);
    // 1 / (1 + e): MUFU reciprocal + one Newton step (<= 1 ulp; 1 + e is in [1, 1e13], no special cases)
    const float d = 1.0f + e;
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(d));
    r = fmaf(fmaf(-d, r, 1.0f), r, r);
    return is_tanh ? (1.0f - e) * r : r;
}

#define RF_STAMP(slot)                                                                                   \
    do {                                                                                                  \
        if (p.ts && blockIdx.x == 0 && blockIdx.y == 0 && s < 64) p.ts[s * 8 + (slot)] = clock64();       \
    } while (0)

// N = padded sequence count of the tile (multiple of 16, <= 64); SPW = N / 4 sequences per epilogue warp
// FAST: act_fast instead of the expf form; RAGGED = false: every sequence of the launch has T frames (no per-step length checks).
// (Measured and dropped: pushing the new h values straight into the 8 CTAs with st.async -- 4 bytes per lane and destination, no
//  staging, no copy warp -- is correct but 16 k remote stores per CTA and step cost more than they save: 10.5 k clk per step
//  against 6.7 k with the staged bulk copies.)
template <int N, bool FAST, bool RAGGED>
__global__ void __launch_bounds__(RF_THREADS, 1) lstm_rec_f16_kernel(const RecF16Params p) {
    constexpr int SPW = N / 4;
    constexpr int NBLK = N / 16;
    constexpr int NSUB = NBLK >= 2 ? 2 : 1;
    constexpr int BLK_A = (NBLK + 1) / 2;                 // 16-sequence blocks of sub-tile 0
    constexpr int R0 = 16 * BLK_A, R1 = N - R0;           // rows of the two sub-tiles
    constexpr uint32_t SUB1_H = 8u * 2u * R0 * 64u;       // byte offset of sub-tile 1 inside the h region
    constexpr uint32_t HBYTES = 8u * 2u * N * 64u;
    constexpr uint32_t STG_PAR = 2u * N * 64u;            // one parity of the staging: [sub][hi | lo][rows x 64 B]
    constexpr uint32_t SUB1_STG = 2u * R0 * 64u;
    static_assert(N % 16 == 0 && N <= 64 && COL_D + 2 * N <= 512, "tile");

    extern __shared__ unsigned char smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    unsigned char* gen = smem_raw + (base - smem_u32(smem_raw));
    const uint32_t s_h = base, s_stg = s_h + HBYTES;
    unsigned char* g_h = gen;
    unsigned char* g_stg = gen + HBYTES;
    int* lens = reinterpret_cast<int*>(g_stg + 2 * STG_PAR);
    uint32_t* goff = reinterpret_cast<uint32_t*>(lens + N);
    uint32_t* yoff = goff + N;
    const uint32_t s_bars = s_stg + 2 * STG_PAR + 3 * N * 4;
    // per sub-tile: bar_full (h rows arrived), bar_mma (MMAs committed), bar_free (all peers' MMAs done), bar_stage (slice staged)
    const uint32_t bar_full = (s_bars + 7u) & ~7u, bar_mma = bar_full + 16, bar_free = bar_full + 32, bar_stage = bar_full + 48,
                   tmem_slot = bar_full + 64;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int rank = (int)cluster_ctarank();
    const int tile = blockIdx.x / TCC, dir = blockIdx.y;
    const int NB = p.NB;
    const int b_begin = tile * NB;
    const int nb = min(NB, p.B - b_begin);
    const int G4 = p.dirs * 4 * TH, Y2 = p.dirs * TH;

    if (tid == 0) {
        for (int b = 0; b < 2; ++b) {
            mbar_init(bar_full + 8 * b, 1);
            mbar_init(bar_mma + 8 * b, 1);
            mbar_init(bar_free + 8 * b, TCC);
            mbar_init(bar_stage + 8 * b, EPI_WARPS);
        }
        mbar_fence_init_cluster();
    }
    if (warp == EPI_WARPS) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    for (int i = tid; i < N; i += RF_THREADS) {
        lens[i] = (i < nb) ? (p.lengths ? min(max(p.lengths[b_begin + i], 0), p.T) : p.T) : 0;
        goff[i] = (uint32_t)(b_begin + min(i, nb - 1)) * (uint32_t)p.T * (uint32_t)G4 + (uint32_t)(dir * 4 * TH + rank * TUC * 4);
        yoff[i] = (uint32_t)(b_begin + min(i, nb - 1)) * (uint32_t)p.T * (uint32_t)Y2 + (uint32_t)(dir * TH + rank * TUC);
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *reinterpret_cast<volatile uint32_t*>(gen + (tmem_slot - base));

    // epilogue-thread coordinates: TMEM lane quarter, row m = unit_local*4 + gate, column (sequence) part
    const int lq = warp & 3, part = (warp >> 2) & 3;
    const int m = lq * 32 + lane, ul = m >> 2, gate = m & 3;
    const uint32_t lane_base = (uint32_t)(lq * 32) << 16;

    // ---- W_hh slice -> TMEM as fp16 hi + scaled lo (once): lane = row, 2 k per column ----------------------------
    if (warp < EPI_WARPS) {
        const float* wrow = (dir ? p.w1 : p.w0) + (size_t)(gate * TH + rank * TUC + ul) * TH;
        // 128 columns per array = 8 chunks of 16 columns (32 k); the 4 warps of a lane quarter take 2 chunks each
        for (int kc = part; kc < TH / 32; kc += 4) {
            const int k0 = kc * 32;
            uint32_t hi[16], lo[16];
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const float4 v = __ldg(reinterpret_cast<const float4*>(wrow + k0) + i);
                unsigned short h0, l0, h1, l1, h2, l2, h3, l3;
                split16(v.x, h0, l0); split16(v.y, h1, l1); split16(v.z, h2, l2); split16(v.w, h3, l3);
                hi[2 * i] = pk(h0, h1); hi[2 * i + 1] = pk(h2, h3);
                lo[2 * i] = pk(l0, l1); lo[2 * i + 1] = pk(l2, l3);
            }
            tmem_st16(tmem + lane_base + COL_WHI + kc * 16, hi);
            tmem_st16(tmem + lane_base + COL_WLO + kc * 16, lo);
        }
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    }
    // ---- h_0 (hi / lo) for the whole tile, zeros in the padded rows ---------------------------------------------
    for (int i = tid; i < N * (TH / 8); i += RF_THREADS) {
        const int n = i / (TH / 8), ch = i % (TH / 8);        // chunk ch covers k = 8 ch .. 8 ch + 7 (one 16-byte swizzle unit)
        float v[8];
#pragma unroll
        for (int q = 0; q < 8; ++q) v[q] = 0.f;
        if (p.h0 && n < nb) {
            const float4* src = reinterpret_cast<const float4*>(p.h0 + ((size_t)dir * p.B + b_begin + n) * TH) + 2 * ch;
            const float4 a = __ldg(src), b = __ldg(src + 1);
            v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
        }
        unsigned short hh[8], ll[8];
#pragma unroll
        for (int q = 0; q < 8; ++q) split16(v[q], hh[q], ll[q]);
        const int sub = n >= R0 ? 1 : 0;
        const int rs = n - (sub ? R0 : 0), rows = sub ? R1 : R0;
        const int kb = ch >> 2;                               // K-block = source rank
        const uint32_t off = (sub ? SUB1_H : 0u) + (uint32_t)kb * (2u * rows * 64u) + sw64_off(rs, (ch & 3) * 8);
        *reinterpret_cast<uint4*>(g_h + off) = make_uint4(pk(hh[0], hh[1]), pk(hh[2], hh[3]), pk(hh[4], hh[5]), pk(hh[6], hh[7]));
        *reinterpret_cast<uint4*>(g_h + off + rows * 64u) = make_uint4(pk(ll[0], ll[1]), pk(ll[2], ll[3]), pk(ll[4], ll[5]), pk(ll[6], ll[7]));
    }
    // cell state: lane `gate` of a unit's quad owns sequence 16*blk + 4*part + gate of every block of 16 sequences
    float cst[SPW / 4];
#pragma unroll
    for (int blk = 0; blk < SPW / 4; ++blk) {
        const int n = blk * 16 + part * 4 + gate;
        cst[blk] = (warp < EPI_WARPS && p.c0 && n < nb) ? p.c0[((size_t)dir * p.B + b_begin + n) * TH + rank * TUC + ul] : 0.f;
    }
    int maxlen = 0;
    for (int i = 0; i < nb; ++i) maxlen = max(maxlen, lens[i]);

    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();
    tc_fence_after();

    // gate pre-activations of step 0 (gin columns are (unit, gate)-ordered: a warp reads 128 contiguous bytes); go[j] is
    // the running offset of the NEXT frame of sequence j
    float gi[SPW];
    uint32_t go[SPW];
    const uint32_t dG4 = dir ? (uint32_t)(-G4) : (uint32_t)G4;
#pragma unroll
    for (int j = 0; j < SPW; ++j) {
        const int n = (j >> 2) * 16 + part * 4 + (j & 3);      // block j/4 of 16 sequences, this warp's 4 columns in it
        const int l = lens[n];
        go[j] = goff[n] + (uint32_t)(dir ? max(l - 1, 0) : 0) * (uint32_t)G4 + (uint32_t)(ul * 4 + gate);
        gi[j] = (warp < EPI_WARPS && l > 0) ? __ldg(p.gin + go[j]) : 0.f;
        go[j] += dG4;
    }
    // this lane's own sequence of every block: length, staging offset, running output offset
    int len_own[SPW / 4];
    uint32_t so[SPW / 4], yo[SPW / 4];
    const uint32_t dY2 = dir ? (uint32_t)(-Y2) : (uint32_t)Y2;
#pragma unroll
    for (int blk = 0; blk < SPW / 4; ++blk) {
        const int n = blk * 16 + part * 4 + gate;
        len_own[blk] = lens[n];
        const int sub = blk >= BLK_A ? 1 : 0;
        so[blk] = (sub ? SUB1_STG : 0u) + sw64_off(n - (sub ? R0 : 0), ul);      // hi plane; the lo plane is rows * 64 B further
        yo[blk] = yoff[n] + (uint32_t)(dir ? max(len_own[blk] - 1, 0) : 0) * (uint32_t)Y2 + (uint32_t)ul;
    }

    const uint32_t d_main = tmem + COL_D, d_corr = tmem + COL_D + N;

    for (int s = 0; s < maxlen; ++s) {
        const bool send = (s + 1 < maxlen);
        const int par = s & 1;
        if (warp == EPI_WARPS) {
            // ================= MMA issuer =================
            // the whole warp walks the fully unrolled issue sequence, one elected lane executes each tcgen05.mma (warp-uniform
            // control flow keeps the descriptors in uniform registers)
            const bool leader = elect_one();
#pragma unroll
            for (int sub = 0; sub < NSUB; ++sub) {
                const uint32_t rows = sub == 0 ? R0 : R1;
                const uint32_t row0 = sub == 0 ? 0 : R0;
                const uint32_t hb = s_h + (sub ? SUB1_H : 0u);
                const uint32_t kstride = 2u * rows * 64u;               // one K-block (hi + lo planes) of this sub-tile
                const uint32_t idesc = (1u << 4) | ((rows >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
                if (s > 0) mbar_wait(bar_full + 8 * sub, (s - 1) & 1);
                tc_fence_after();
                if (leader && sub == 0) RF_STAMP(0);
                const uint64_t bd_hi = umma_desc_sw64(hb), bd_lo = umma_desc_sw64(hb + rows * 64u);
                const uint32_t dm_ = d_main + row0, dc_ = d_corr + row0;
#pragma unroll
                for (int ks = 0; ks < TH / 16; ++ks)           // correction: W_lo . h_hi
                    if (leader) umma_ts_f16(dc_, tmem + COL_WLO + ks * 8, bd_hi + (uint64_t)(((ks >> 1) * kstride + (ks & 1) * 32) >> 4), idesc, ks != 0);
#pragma unroll
                for (int ks = 0; ks < TH / 16; ++ks)           // correction: W_hi . h_lo
                    if (leader) umma_ts_f16(dc_, tmem + COL_WHI + ks * 8, bd_lo + (uint64_t)(((ks >> 1) * kstride + (ks & 1) * 32) >> 4), idesc, 1u);
#pragma unroll
                for (int ks = 0; ks < TH / 16; ++ks)           // main: W_hi . h_hi
                    if (leader) umma_ts_f16(dm_, tmem + COL_WHI + ks * 8, bd_hi + (uint64_t)(((ks >> 1) * kstride + (ks & 1) * 32) >> 4), idesc, ks != 0);
                if (leader) {
                    if (sub == NSUB - 1) RF_STAMP(1);
                    tc_commit(bar_mma + 8 * sub);
                }
            }
            __syncwarp();
        } else if (warp > EPI_WARPS) {
            // ================= exchange warps (one per sub-tile) =================
            // (a) once the sub-tile's MMAs of this step have completed: arm its `h_full` barrier for the next step's rows and tell all 8
            //     CTAs that this CTA's tensor core no longer reads those rows (`h_free`, relaxed remote arrive);
            // (b) ship the new slice (hi and lo planes, one contiguous block) to all 8 CTAs once the 16 epilogue warps have staged it
            //     and every peer's MMAs on the sub-tile's old rows are done.
            // One warp per sub-tile: with a single warp the second sub-tile's (a) -- which waits for MMAs that start half a step
            // later -- sat in front of the first sub-tile's (b) and delayed its copies by ~1 k clk per step.
            const int sub = warp - EPI_WARPS - 1;
            if (send && sub < NSUB) {
                const uint32_t rows = sub == 0 ? R0 : R1;
                const uint32_t bytes = 2u * rows * 64u;
                mbar_wait(bar_mma + 8 * sub, par);
                // h_{s+1} of this sub-tile arrives as 8 ranks x (hi + lo planes).  Armed only now: MMA(s) has run, so the MMA
                // warp has seen the previous phase of the barrier complete (arming earlier could put two arrivals in one phase)
                if (lane == 0) mbar_arrive_expect_tx(bar_full + 8 * sub, (uint32_t)TCC * bytes);
                if (lane < TCC) mbar_arrive_remote(mapa_u32(bar_free + 8 * sub, lane));
                mbar_wait(bar_stage + 8 * sub, par);
                if (lane == 0 && sub == 0) RF_STAMP(5);
                mbar_wait(bar_free + 8 * sub, par);
                if (p.ts && lane == 0 && sub == 0 && blockIdx.x == 0 && blockIdx.y == 0 && s < 64) p.ts[512 + s * 20 + 16] = clock64();
                if (lane < TCC)
                    bulk_copy_s2c(mapa_u32(s_h + (sub ? SUB1_H : 0u) + (uint32_t)rank * bytes, lane),
                                  s_stg + (uint32_t)par * STG_PAR + (sub ? SUB1_STG : 0u), bytes, mapa_u32(bar_full + 8 * sub, lane));
                if (lane == 0 && sub == 0) RF_STAMP(6);
            }
            __syncwarp();
        } else {
            // ================= epilogue =================
            unsigned char* stg = g_stg + (size_t)par * STG_PAR;
#pragma unroll
            for (int sub = 0; sub < NSUB; ++sub) {
                const int blk0 = sub == 0 ? 0 : BLK_A;
                const int nblk = sub == 0 ? BLK_A : NBLK - BLK_A;
                const uint32_t rows = sub == 0 ? R0 : R1;
                mbar_wait(bar_mma + 8 * sub, par);
                tc_fence_after();
                if (tid == 0 && sub == 0) RF_STAMP(2);
                float dm[4 * (NBLK - NBLK / 2)], dc[4 * (NBLK - NBLK / 2)];
#pragma unroll
                for (int q = 0; q < nblk; ++q) {
                    tmem_ld4(d_main + lane_base + (blk0 + q) * 16 + part * 4, dm + 4 * q);
                    tmem_ld4(d_corr + lane_base + (blk0 + q) * 16 + part * 4, dc + 4 * q);
                }
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                if (tid == 0 && sub == 0) RF_STAMP(3);
                // Blocks of 4 sequences per lane quad: every lane evaluates its gate for the 4 sequences, the quad (4 gate
                // lanes of a unit) exchanges them, and lane g then owns sequence 4*blk + g: ONE cell update per lane.
                constexpr int MAXB = NBLK - NBLK / 2;
                float c_nw[MAXB], h_nw[MAXB];
                unsigned short h_hi16[MAXB], h_lo16[MAXB];
#pragma unroll
                for (int bq = 0; bq < nblk; ++bq) {
                    const int blk = blk0 + bq;
                    float a[4];
#pragma unroll
                    for (int q = 0; q < 4; ++q)
                        a[q] = FAST ? act_fast(fmaf(dc[bq * 4 + q], kLoInv, dm[bq * 4 + q]) + gi[blk * 4 + q], gate == 2)
                                    : act_sigmoid_or_tanh(fmaf(dc[bq * 4 + q], kLoInv, dm[bq * 4 + q]) + gi[blk * 4 + q], gate == 2);
                    // 4 x 4 transpose inside the gate quad: two butterfly rounds, 4 shuffles instead of 16
                    const bool b0 = lane & 1, b1 = lane & 2;
                    const float r0 = __shfl_xor_sync(0xffffffffu, b0 ? a[0] : a[1], 1), r1 = __shfl_xor_sync(0xffffffffu, b0 ? a[2] : a[3], 1);
                    const float u0 = b0 ? r0 : a[0], u1 = b0 ? a[1] : r0, u2 = b0 ? r1 : a[2], u3 = b0 ? a[3] : r1;
                    const float q0 = __shfl_xor_sync(0xffffffffu, b1 ? u0 : u2, 2), q1 = __shfl_xor_sync(0xffffffffu, b1 ? u1 : u3, 2);
                    const float iv = b1 ? q0 : u0, fv = b1 ? q1 : u1, gv = b1 ? u2 : q0, ov = b1 ? u3 : q1;
                    c_nw[bq] = fmaf(fv, cst[blk], iv * gv);
                    h_nw[bq] = ov * (FAST ? act_fast(c_nw[bq], true) : act_sigmoid_or_tanh(c_nw[bq], true));
                }
                // (1) the exchange first: nothing that touches global memory may sit between the activations and the hand-over --
                // fence.proxy.async waits for every earlier memory operation of the thread, and with the y stores and the next step's
                // gin loads in front of it the epilogue stalled a DRAM round trip per sub-tile (ncu: a third of all stall samples)
#pragma unroll
                for (int bq = 0; bq < nblk; ++bq) {
                    const int blk = blk0 + bq;
                    const int n = blk * 16 + part * 4 + gate;
                    const bool active = !RAGGED || s < len_own[blk];
                    split16(active ? h_nw[bq] : 0.f, h_hi16[bq], h_lo16[bq]);
                    if (send) {
                        *reinterpret_cast<unsigned short*>(stg + so[blk]) = h_hi16[bq];
                        *reinterpret_cast<unsigned short*>(stg + so[blk] + rows * 64u) = h_lo16[bq];
                    }
                }
                if (send) {
                    // the sub-tile's slice is staged: make the generic-proxy writes visible to the bulk copy, one arrival per warp
                    fence_proxy_async_smem();
                    __syncwarp();
                    if (lane == 0) mbar_arrive_local(bar_stage + 8 * sub);
                    if (tid == 0 && sub == 0) RF_STAMP(7);
                    if (p.ts && lane == 0 && sub == 0 && blockIdx.x == 0 && blockIdx.y == 0 && s < 64) p.ts[512 + s * 20 + warp] = clock64();
                }
                // (2) then the layer output, the final states and the next step's gate pre-activations
#pragma unroll
                for (int bq = 0; bq < nblk; ++bq) {
                    const int blk = blk0 + bq;
                    const int n = blk * 16 + part * 4 + gate;
                    const int len = len_own[blk];
                    const bool active = !RAGGED || s < len;
                    const float c_new = c_nw[bq], h_new = h_nw[bq];
                    if (active) {
                        cst[blk] = c_new;
                        if (p.y_split) {
                            unsigned short* yh = reinterpret_cast<unsigned short*>(p.y);
                            yh[yo[blk]] = h_hi16[bq];
                            yh[(size_t)p.B * p.T * Y2 + yo[blk]] = h_lo16[bq];
                        } else {
                            p.y[yo[blk]] = h_new;
                        }
                        if (s == len - 1) {
                            if (p.hn) p.hn[((size_t)dir * p.B + b_begin + n) * TH + rank * TUC + ul] = h_new;
                            if (p.cn) p.cn[((size_t)dir * p.B + b_begin + n) * TH + rank * TUC + ul] = c_new;
                        }
                    }
                    yo[blk] += dY2;
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        if (RAGGED ? (s + 1 < lens[blk * 16 + part * 4 + q]) : send) gi[blk * 4 + q] = __ldg(p.gin + go[blk * 4 + q]);
                        go[blk * 4 + q] += dG4;
                    }
                }
            }
            if (tid == 0) RF_STAMP(4);
        }
    }

    // frames >= len of the layer output are zero (pad_packed_sequence, rnn.py:31)
    for (int b = 0; b < nb; ++b) {
        const int len = lens[b];
        const int cnt = (p.T - len) * TUC;
        for (int i = tid; i < cnt; i += RF_THREADS) {
            const int t = len + i / TUC, u = i % TUC;
            const uint32_t idx = yoff[b] + (uint32_t)t * (uint32_t)Y2 + (uint32_t)u;
            if (p.y_split) {
                unsigned short* yh = reinterpret_cast<unsigned short*>(p.y);
                yh[idx] = 0;
                yh[(size_t)p.B * p.T * Y2 + idx] = 0;
            } else {
                p.y[idx] = 0.f;
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();
    if (warp == EPI_WARPS) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
}

template <int N, bool FAST, bool RAGGED>
int launch_f16_n(const RecF16Params& p, cudaStream_t stream) {
    const size_t smem = rec_f16_smem_bytes(N);
    auto kern = lstm_rec_f16_kernel<N, FAST, RAGGED>;
    static bool configured = false;
    if (!configured) {
        MP_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured = true;
    }
    const int n_tiles = (p.B + p.NB - 1) / p.NB;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(TCC * n_tiles, p.dirs, 1);
    cfg.blockDim = dim3(RF_THREADS, 1, 1);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = TCC;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    MP_CUDA_TRY(cudaLaunchKernelEx(&cfg, kern, p));
    count_launch();
    return MP_OK;
}

int f16_cluster_slots() {
    static int slots = 0;
    if (slots > 0) return slots;
    int n = 0;
    auto kern = lstm_rec_f16_kernel<64, false, true>;
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)rec_f16_smem_bytes(64));
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(TCC * 64, 1, 1);
    cfg.blockDim = dim3(RF_THREADS, 1, 1);
    cfg.dynamicSmemBytes = rec_f16_smem_bytes(64);
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = TCC;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    if (cudaOccupancyMaxActiveClusters(&n, kern, &cfg) != cudaSuccess) {
        cudaGetLastError();
        n = 0;
    }
    slots = n > 0 ? n : 8;
    return slots;
}

}  // namespace

// fp16-split tensor-core recurrence for H = 256 when the batch is large enough to fill 16-wide MMA tiles (same policy as the
// TF32 kernel it supersedes; MP_REC_IMPL=tf32 pins the old one, =f16 / =tc force this one)
bool rec_f16_eligible(const RecLayerArgs& a) {
    const char* v = getenv("MP_REC_IMPL");
    if (v && (strcmp(v, "ffma") == 0 || strcmp(v, "simple") == 0 || strcmp(v, "tf32") == 0)) return false;
    if (a.H != TH || !a.w_raw[0]) return false;
    const bool forced = v && (strcmp(v, "tc") == 0 || strcmp(v, "f16") == 0);
    return forced || a.B * a.dirs > 2 * f16_cluster_slots();
}

int launch_lstm_recurrence_f16(const RecLayerArgs& a, cudaStream_t stream) {
    MP_REQUIRE((double)a.B * a.T * a.dirs * 4 * a.H < 4.0e9, "lstm_f16: B*T = %lld frames exceeds 32-bit gate buffer indexing", (long long)a.B * a.T);
    const char* nbv = getenv("MP_REC_NB");
    int NB = (nbv && *nbv) ? atoi(nbv) : a.tile_hint;
    if (NB == 0) {
        if (a.B >= 128) {
            NB = 64;
        } else {
            const int per = std::max(1, f16_cluster_slots() / a.dirs);
            NB = (a.B + per - 1) / per;
        }
    } else if (NB < 0) {      // -1: the one-wave policy whatever the batch
        const int per = std::max(1, f16_cluster_slots() / a.dirs);
        NB = (a.B + per - 1) / per;
    }
    NB = std::min(64, std::max(1, NB));
    const int N = ((NB + 15) / 16) * 16;
    const int n_tiles = (a.B + NB - 1) / NB;     // balance the tiles: same tile count, equal sizes
    NB = (a.B + n_tiles - 1) / n_tiles;
    static long long* ts_dev = nullptr;
    const bool want_ts = getenv("MP_RTC_TS") != nullptr;
    if (want_ts && !ts_dev) {
        cudaMalloc(&ts_dev, (64 * 8 + 64 * 20) * sizeof(long long));
        cudaMemset(ts_dev, 0, (64 * 8 + 64 * 20) * sizeof(long long));
    }
    RecF16Params p{a.gin, a.w_raw[0], a.w_raw[a.dirs - 1], a.y, a.h0, a.c0, a.hn, a.cn, a.lengths, a.B, a.T, a.dirs, NB, a.y_split,
                   want_ts ? ts_dev : nullptr};
    ProfileScope prof("lstm_rec_f16_h256", 4.0 * ((double)a.dirs * 4 * a.H * a.H + (double)a.B * a.T * a.dirs * a.H), stream);
    // the short ex2 / rcp activation by default (MP_RF16_ACT=exact: the expf form); the uniform-length variant when no lengths are given
    // (a tile's padded rows then run as zero-input sequences whose results are never stored: n >= nb is masked by len = 0 only in
    // the ragged variant, so the uniform one is taken only for full tiles)
    const char* actv = getenv("MP_RF16_ACT");
    const bool fast = !(actv && strcmp(actv, "exact") == 0);
    const bool ragged = a.lengths != nullptr || a.B % NB != 0 || NB % 16 != 0 || getenv("MP_RF16_RAGGED") != nullptr;
    int st;
#define RF16_DISPATCH(NN) (fast ? (ragged ? launch_f16_n<NN, true, true>(p, stream) : launch_f16_n<NN, true, false>(p, stream)) \
                                : (ragged ? launch_f16_n<NN, false, true>(p, stream) : launch_f16_n<NN, false, false>(p, stream)))
    switch (N) {
        case 16: st = RF16_DISPATCH(16); break;
        case 32: st = RF16_DISPATCH(32); break;
        case 48: st = RF16_DISPATCH(48); break;
        default: st = RF16_DISPATCH(64); break;
    }
#undef RF16_DISPATCH
#undef RF16_DISPATCH
    if (st == MP_OK && want_ts) {      // bring-up only: synchronous dump of the stamps of block (0,0)
        static long long h[64 * 8 + 64 * 20];
        cudaStreamSynchronize(stream);
        cudaMemcpy(h, ts_dev, sizeof(h), cudaMemcpyDeviceToHost);
        fprintf(stderr, "[rtc ts f16] N=%d NB=%d B=%d\n", N, NB, a.B);
        for (int s = 2; s < 8 && s < a.T - 1; ++s)
            fprintf(stderr, "[rtc ts f16] s=%d  step %lld | sub0: mma start->epi start %lld  tmem ld %lld  act+stage %lld  staged->copy warp %lld  copy issue %lld  "
                            "copies issued->next mma start %lld | both subs: mma issue %lld  epilogue %lld\n", s,
                    h[(s + 1) * 8 + 0] - h[s * 8 + 0], h[s * 8 + 2] - h[s * 8 + 0], h[s * 8 + 3] - h[s * 8 + 2], h[s * 8 + 7] - h[s * 8 + 3],
                    h[s * 8 + 5] - h[s * 8 + 7], h[s * 8 + 6] - h[s * 8 + 5], h[(s + 1) * 8 + 0] - h[s * 8 + 6], h[s * 8 + 1] - h[s * 8 + 0],
                    h[s * 8 + 4] - h[s * 8 + 2]);
        for (int s = 3; s < 6 && s < a.T - 1; ++s) {      // sub-tile 0: when each epilogue warp reported its rows staged, relative to the first
            const long long* w = h + 512 + s * 20;
            long long first = w[0];
            for (int i = 1; i < 16; ++i) first = std::min(first, w[i]);
            fprintf(stderr, "[rtc ts f16] s=%d  sub0 staged, per warp (clk after the first):", s);
            for (int i = 0; i < 16; ++i) fprintf(stderr, " %lld", w[i] - first);
            fprintf(stderr, " | copy warp: stage barrier seen +%lld, free barrier seen +%lld, copies issued +%lld, next mma start +%lld\n",
                    h[s * 8 + 5] - first, w[16] - first, h[s * 8 + 6] - first, h[(s + 1) * 8 + 0] - first);
        }
    }
    return st;
}

}  // namespace mp
