// Tensor-core LSTM recurrence, wide tile: the kernel of lstm_rec_f16.cu with 128 sequences per 8-CTA cluster as FOUR sub-tiles of 32.
// Replaces the time loop of _VF.lstm (mobileposer/models/rnn.py:27) for H = 256 when the batch holds whole 128-sequence tiles of equal
// length (the cfg3 throughput path).
//
// Why: in the 64-sequence kernel a sub-tile's loop  MMA (1.0 k clk) -> epilogue (1.2-1.6 k, 2.2 k until the slowest of the 16 warps
// has staged its rows) -> exchange (1.2 k) -> next MMA  is about as long as the whole step (5 k clk, profiles/r02_rec_f16_stamps.txt):
// with two sub-tiles in flight the tensor pipe is busy 40 % of the step and the epilogue warps 60 %, the rest is waiting on that chain.
// Four sub-tiles keep both fed -- the 16 epilogue warps always find a finished sub-tile -- so a step costs what the epilogue work
// costs and a cluster serves twice the sequences: half the CTAs per layer launch (32 instead of 64 for a bidirectional cfg3 layer) for
// about 1.2 x the time.  Same numerics, operands, exchange protocol and thread coordinates as lstm_rec_f16.cu (read its header first):
//   TMEM     W_hi columns [0, 128), W_lo [128, 256) (fp16 pairs), main accumulators [256, 384), correction [384, 512): all 512 columns
//   smem     h: [sub-tile][K-block = source rank][hi | lo][32 rows x 64 B] (128 KB); staging [parity][sub-tile][hi | lo][32 x 64 B] (32 KB);
//            gate pre-activations of the next step [sub-tile][32 sequences][128 floats] (64 KB), landed by TMA
//   warps    0-15 epilogue (TMEM lane quarter = warp % 4, column part = warp / 4), 16 MMA issuer, 17 / 18 exchange (sub-tiles 0, 2 / 1, 3)
// Differences: (a) the per-sub-tile housekeeping after the MMAs (arm `h_full`, tell the peers the rows are free) is done by epilogue
// warp 0 at the moment it has seen the MMAs complete anyway, the exchange warps only ship; (b) the per-sequence running offsets of the
// 64-wide kernel (2 registers per column) are recomputed from two running bases -- uniform lengths make them affine in the sequence index.
#include "mp_common.cuh"

#include <cuda.h>
#include <cuda_fp16.h>

#include <cstdlib>
#include <cstring>

namespace mp {

namespace {

constexpr int TH = 256;
constexpr int TCC = 8;             // cluster size
constexpr int TUC = TH / TCC;      // 32 units per CTA
constexpr int WN = 128;            // sequences per cluster tile
constexpr int WSUB = 4;            // sub-tiles
constexpr int WR = WN / WSUB;      // 32 rows (sequences) per sub-tile
constexpr int EPI_WARPS = 16;
constexpr int WF_THREADS = (EPI_WARPS + 3) * 32;
constexpr uint32_t COL_WHI = 0, COL_WLO = 128, COL_MAIN = 256, COL_CORR = 384;
constexpr uint32_t PLANE = WR * 64u;                 // one plane (hi or lo) of one K-block of one sub-tile: 2 KB
constexpr uint32_t KBLOCK = 2u * PLANE;              // hi + lo: what one rank ships per sub-tile and step (4 KB)
constexpr uint32_t SUBH = TCC * KBLOCK;              // h of one sub-tile: 32 KB
constexpr uint32_t HBYTES = WSUB * SUBH;             // 128 KB
constexpr uint32_t STG_PAR = WSUB * KBLOCK;          // 16 KB per parity
constexpr uint32_t GIN_SUB = WR * 128u * 4u;           // gate pre-activations of one sub-tile and step: 32 sequences x 128 floats (16 KB)
constexpr uint32_t WF_SMEM = 1024 + HBYTES + 2 * STG_PAR + WSUB * GIN_SUB + 256;
constexpr float kLoScale = 2048.0f, kLoInv = 1.0f / 2048.0f;

struct RecF16WParams {
    const float* gin;
    const float* w0;      // raw torch W_hh [4H, H] of direction 0 / 1
    const float* w1;
    float* y;
    const float* h0;
    const float* c0;
    float* hn;
    float* cn;
    int B, T, dirs;
    int y_split;          // y = two planes of halves (hi, scaled lo) instead of fp32
    long long* ts;        // bring-up (MP_RECW_TS): clock64 stamps of block (0,0), [step][16], or null
    int skip;             // ablation builds only (-DMP_RECW_ABLATION, env MP_RECW_SKIP; results are WRONG with a bit set): bit 0 no y stores,
                          // bit 1 no gin prefetch, bit 2 no activations.  Compiled out of the shipped library.
    int y_tma;            // y_split only: the layer output leaves through TMA stores of the staged slices (map_y_hi / map_y_lo)
};

__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void umma_ts_f16(uint32_t d, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d),
        "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(acc)
        : "memory");
}
__device__ __forceinline__ uint64_t umma_desc_sw64(uint32_t smem_addr) {
    return (uint64_t)((smem_addr & 0x3FFFF) >> 4) | (1ull << 16) | ((uint64_t)(512 >> 4) << 32) | (1ull << 46) | (4ull << 61);
}
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
__device__ __forceinline__ void mbar_arrive_remote_relaxed(uint32_t remote_bar) {
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
__device__ __forceinline__ void split16(float x, unsigned short& hi, unsigned short& lo) {
    const __half h = __float2half_rn(x);
    hi = __half_as_ushort(h);
    lo = __half_as_ushort(__float2half_rn((x - __half2float(h)) * kLoScale));
}
__device__ __forceinline__ uint32_t pk(unsigned short a, unsigned short b) { return (uint32_t)a | ((uint32_t)b << 16); }
// sigma(s) = 1 / (1 + 2^(-s log2 e)) with MUFU ex2 + MUFU rcp; tanh(x) = 2 sigma(2x) - 1.  ex2.approx is 2 ulp on e and rcp.approx 1 ulp
// on the quotient: <= 1.5e-7 absolute on a gate value.  No clamp and no Newton step: e = +inf gives rcp(inf) = 0, the correct limit, and
// nothing downstream multiplies it by d again (5 instructions per activation instead of 8; the epilogue is instruction-bound).
__device__ __forceinline__ float act_fast(float x, bool is_tanh) {
    float e, r;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(x * (is_tanh ? -2.8853900817779268f : -1.4426950408889634f)));
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(1.0f + e));
    return is_tanh ? fmaf(2.0f, r, -1.0f) : r;
}
__device__ __forceinline__ float act_exact(float x, bool is_tanh) {
    const float s = fminf(fmaxf(is_tanh ? 2.0f * x : x, -30.0f), 30.0f);
    const float e = expf(-s);
    const float d = 1.0f + e;
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(d));
    r = fmaf(fmaf(-d, r, 1.0f), r, r);
    return is_tanh ? (1.0f - e) * r : r;
}

#ifdef MP_RECW_ABLATION
#define WF_SKIP(p, bit) ((p).skip & (bit))
#else
#define WF_SKIP(p, bit) 0
#endif
#define WF_STAMP(slot)                                                                                         \
    do {                                                                                                       \
        if (p.ts && blockIdx.x == 0 && blockIdx.y == 0 && s < 64) p.ts[s * 16 + (slot)] = clock64();           \
    } while (0)

// one staged plane (32 sequences x this rank's 32 units, 64-byte rows, 64B swizzle) -> y plane [B, T, dirs * H] at (unit0, t, sequence0)
__device__ __forceinline__ void tma_store_3d(const CUtensorMap* map, uint32_t src, int c0, int c1, int c2) {
    asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];" ::"l"(reinterpret_cast<uint64_t>(map)),
                 "r"(src), "r"(c0), "r"(c1), "r"(c2)
                 : "memory");
}

// this CTA's 128 gate columns of 32 sequences at one frame of gin [B, T, dirs * 4H] -> shared memory, bytes counted on `bar`
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* map, int c0, int c1, int c2, uint32_t bar) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(dst),
        "l"(reinterpret_cast<uint64_t>(map)), "r"(c0), "r"(c1), "r"(c2), "r"(bar)
        : "memory");
}

// GTMA: the next step's gate pre-activations arrive by TMA in shared memory (one 16 KB box per sub-tile and step, issued by the
// exchange warp once the 16 epilogue warps have consumed the previous one) instead of 8 LDG + address arithmetic per thread and sub-tile
template <bool FAST, bool GTMA>
__global__ void __launch_bounds__(WF_THREADS, 1) lstm_rec_f16w_kernel(const RecF16WParams p, const __grid_constant__ CUtensorMap map_y_hi,
                                                                       const __grid_constant__ CUtensorMap map_y_lo,
                                                                       const __grid_constant__ CUtensorMap map_gin) {
    extern __shared__ unsigned char smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    unsigned char* gen = smem_raw + (base - smem_u32(smem_raw));
    const uint32_t s_h = base, s_stg = s_h + HBYTES;
    unsigned char* g_h = gen;
    unsigned char* g_stg = gen + HBYTES;
    // per sub-tile: bar_full (h rows arrived), bar_mma (MMAs committed), bar_free (all peers' MMAs done), bar_stage (slice staged)
    const uint32_t s_gin = s_stg + 2 * STG_PAR;
    const unsigned char* g_gin = gen + HBYTES + 2 * STG_PAR;
    const uint32_t bar_full = s_gin + WSUB * GIN_SUB, bar_mma = bar_full + 8 * WSUB, bar_free = bar_mma + 8 * WSUB,
                   bar_stage = bar_free + 8 * WSUB, bar_gin = bar_stage + 8 * WSUB, tmem_slot = bar_gin + 8 * WSUB;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int rank = (int)cluster_ctarank();
    const int tile = blockIdx.x / TCC, dir = blockIdx.y;
    const int b_begin = tile * WN;
    const int T = p.T;
    const int G4 = p.dirs * 4 * TH, Y2 = p.dirs * TH;

    if (tid == 0) {
        for (int b = 0; b < WSUB; ++b) {
            mbar_init(bar_full + 8 * b, 1);
            mbar_init(bar_mma + 8 * b, 1);
            mbar_init(bar_free + 8 * b, TCC);
            mbar_init(bar_stage + 8 * b, EPI_WARPS);
            mbar_init(bar_gin + 8 * b, 1);
        }
        mbar_fence_init_cluster();
    }
    if (warp == EPI_WARPS) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
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
    // ---- h_0 (hi / lo) for the whole tile ------------------------------------------------------------------------
    for (int i = tid; i < WN * (TH / 8); i += WF_THREADS) {
        const int n = i / (TH / 8), ch = i % (TH / 8);        // chunk ch covers k = 8 ch .. 8 ch + 7 (one 16-byte swizzle unit)
        float v[8];
#pragma unroll
        for (int q = 0; q < 8; ++q) v[q] = 0.f;
        if (p.h0) {
            const float4* src = reinterpret_cast<const float4*>(p.h0 + ((size_t)dir * p.B + b_begin + n) * TH) + 2 * ch;
            const float4 a = __ldg(src), b = __ldg(src + 1);
            v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
        }
        unsigned short hh[8], ll[8];
#pragma unroll
        for (int q = 0; q < 8; ++q) split16(v[q], hh[q], ll[q]);
        const int sub = n / WR, rs = n % WR;
        const int kb = ch >> 2;                               // K-block = source rank
        const uint32_t off = (uint32_t)sub * SUBH + (uint32_t)kb * KBLOCK + sw64_off(rs, (ch & 3) * 8);
        *reinterpret_cast<uint4*>(g_h + off) = make_uint4(pk(hh[0], hh[1]), pk(hh[2], hh[3]), pk(hh[4], hh[5]), pk(hh[6], hh[7]));
        *reinterpret_cast<uint4*>(g_h + off + PLANE) = make_uint4(pk(ll[0], ll[1]), pk(ll[2], ll[3]), pk(ll[4], ll[5]), pk(ll[6], ll[7]));
    }
    // cell state: lane `gate` of a unit's quad owns sequence 16*blk + 4*part + gate of every block of 16 sequences
    float cst[WN / 16];
#pragma unroll
    for (int blk = 0; blk < WN / 16; ++blk) {
        const int n = blk * 16 + part * 4 + gate;
        cst[blk] = (warp < EPI_WARPS && p.c0) ? p.c0[((size_t)dir * p.B + b_begin + n) * TH + rank * TUC + ul] : 0.f;
    }

    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();
    tc_fence_after();

    // running offsets (uniform lengths: frame t = s forward, T - 1 - s reverse, for every sequence): g_run addresses this thread's gate
    // row of sequence b_begin + 4 part of the NEXT frame to prefetch, y_run its own unit of sequence b_begin + 4 part + gate of the
    // CURRENT frame; sequence 16 blk + j further is (16 blk + j) * T * G4 (resp. Y2) away
    const uint32_t TG4 = (uint32_t)T * (uint32_t)G4, TY2 = (uint32_t)T * (uint32_t)Y2;
    const uint32_t dG4 = dir ? (uint32_t)(-G4) : (uint32_t)G4, dY2 = dir ? (uint32_t)(-Y2) : (uint32_t)Y2;
    uint32_t g_run = ((uint32_t)(b_begin + part * 4) * (uint32_t)T + (uint32_t)(dir ? T - 1 : 0)) * (uint32_t)G4 +
                     (uint32_t)(dir * 4 * TH + rank * TUC * 4 + m);
    uint32_t y_run = ((uint32_t)(b_begin + part * 4 + gate) * (uint32_t)T + (uint32_t)(dir ? T - 1 : 0)) * (uint32_t)Y2 +
                     (uint32_t)(dir * TH + rank * TUC + ul);
    // gate pre-activations of step 0 (gin columns are (unit, gate)-ordered: a warp reads 128 contiguous bytes)
    float gi[GTMA ? 1 : WN / 4];
    if (GTMA) {
        if (warp > EPI_WARPS && lane == 0) {
#pragma unroll
            for (int k = 0; k < 2; ++k) {
                const int sub = (warp - EPI_WARPS - 1) + 2 * k;
                mbar_arrive_expect_tx(bar_gin + 8 * sub, GIN_SUB);
                tma_load_3d(s_gin + (uint32_t)sub * GIN_SUB, &map_gin, dir * 4 * TH + rank * TUC * 4, dir ? T - 1 : 0, b_begin + sub * WR,
                            bar_gin + 8 * sub);
            }
        }
    } else if (warp < EPI_WARPS) {
        const float* gp = p.gin + g_run;
#pragma unroll
        for (int j = 0; j < WN / 4; ++j) gi[j] = __ldg(gp + (size_t)((j >> 2) * 16 + (j & 3)) * TG4);
    }
    g_run += dG4;

    const uint32_t d_main = tmem + COL_MAIN, d_corr = tmem + COL_CORR;
    const uint32_t idesc = (1u << 4) | ((uint32_t)(WR >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);

    for (int s = 0; s < T; ++s) {
        const bool send = (s + 1 < T);
        const int par = s & 1;
        if (warp == EPI_WARPS) {
            // ================= MMA issuer: the whole warp walks the unrolled sequence, one elected lane issues =================
            const bool leader = elect_one();
#pragma unroll
            for (int sub = 0; sub < WSUB; ++sub) {
                const uint32_t hb = s_h + (uint32_t)sub * SUBH;
                if (s > 0) mbar_wait(bar_full + 8 * sub, (s - 1) & 1);
                tc_fence_after();
                if (leader) WF_STAMP(sub);
                const uint64_t bd_hi = umma_desc_sw64(hb), bd_lo = umma_desc_sw64(hb + PLANE);
                const uint32_t dm_ = d_main + sub * WR, dc_ = d_corr + sub * WR;
#pragma unroll
                for (int ks = 0; ks < TH / 16; ++ks)           // correction: W_lo . h_hi
                    if (leader) umma_ts_f16(dc_, tmem + COL_WLO + ks * 8, bd_hi + (uint64_t)(((ks >> 1) * KBLOCK + (ks & 1) * 32) >> 4), idesc, ks != 0);
#pragma unroll
                for (int ks = 0; ks < TH / 16; ++ks)           // correction: W_hi . h_lo
                    if (leader) umma_ts_f16(dc_, tmem + COL_WHI + ks * 8, bd_lo + (uint64_t)(((ks >> 1) * KBLOCK + (ks & 1) * 32) >> 4), idesc, 1u);
#pragma unroll
                for (int ks = 0; ks < TH / 16; ++ks)           // main: W_hi . h_hi
                    if (leader) umma_ts_f16(dm_, tmem + COL_WHI + ks * 8, bd_hi + (uint64_t)(((ks >> 1) * KBLOCK + (ks & 1) * 32) >> 4), idesc, ks != 0);
                if (leader) tc_commit(bar_mma + 8 * sub);
            }
            __syncwarp();
        } else if (warp > EPI_WARPS) {
            // ================= exchange warps: ship the staged slices (sub-tiles 0, 2 / 1, 3) to all 8 CTAs =================
            // ... and, when the layer output is wanted as (hi, lo) planes, to global memory: the staged slice IS the output block of
            // this rank for the sub-tile (32 sequences x 32 units, both planes), so two TMA stores replace 4 scattered 2-byte stores per
            // thread, sub-tile and step (those cost 22 % of the kernel: MP_RECW_SKIP ablation, profiles/r02_rec_wide_ab.txt)
#pragma unroll
            for (int k = 0; k < 2; ++k) {
                if (!send && !p.y_tma) break;                  // last step without TMA output: nothing was staged
                const int sub = (warp - EPI_WARPS - 1) + 2 * k;
                const uint32_t src = s_stg + (uint32_t)par * STG_PAR + (uint32_t)sub * KBLOCK;
                mbar_wait(bar_stage + 8 * sub, par);           // the 16 epilogue warps have staged the slice (hi + lo planes, contiguous)
                // the hand-over to the peers first (it is on the step's critical chain), the layer output and the next step's gate
                // pre-activations after it (they have a whole step).  Before the copies of this step let the next-but-one epilogue
                // overwrite the other staging parity, last step's output stores of this sub-tile must have read it out (they were
                // issued a step ago: a guarantee, not a wait)
                if (p.y_tma && lane == 0) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
                __syncwarp();
                if (send) {
                    mbar_wait(bar_free + 8 * sub, par);        // every CTA's MMAs of this step on the sub-tile's old rows are done
                    if (lane < TCC)
                        bulk_copy_s2c(mapa_u32(s_h + (uint32_t)sub * SUBH + (uint32_t)rank * KBLOCK, lane), src, KBLOCK,
                                      mapa_u32(bar_full + 8 * sub, lane));
                    if (lane == 0) WF_STAMP(12 + sub);
                }
                if (p.y_tma && lane == 0) {
                    const int t = dir ? T - 1 - s : s;
                    tma_store_3d(&map_y_hi, src, dir * TH + rank * TUC, t, b_begin + sub * WR);
                    tma_store_3d(&map_y_lo, src + PLANE, dir * TH + rank * TUC, t, b_begin + sub * WR);
                    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                }
                if (GTMA && send && lane == 0) {               // the epilogue warps have consumed this step's gate pre-activations
                    mbar_arrive_expect_tx(bar_gin + 8 * sub, GIN_SUB);
                    tma_load_3d(s_gin + (uint32_t)sub * GIN_SUB, &map_gin, dir * 4 * TH + rank * TUC * 4, dir ? T - 2 - s : s + 1,
                                b_begin + sub * WR, bar_gin + 8 * sub);
                }
            }
            __syncwarp();
        } else {
            // ================= epilogue =================
            unsigned char* stg = g_stg + (size_t)par * STG_PAR;
            // one 64-bit base per step; the per-sequence offsets below are constants times T * G4 (one IMAD.WIDE per access)
            const float* gp = p.gin + g_run;
            float* yp = p.y + y_run;
            unsigned short* yh = reinterpret_cast<unsigned short*>(p.y) + y_run;
            unsigned short* yl = yh + (size_t)p.B * T * Y2;
#pragma unroll
            for (int sub = 0; sub < WSUB; ++sub) {
                mbar_wait(bar_mma + 8 * sub, par);
                tc_fence_after();
                if (tid == 0) WF_STAMP(4 + sub);
                if (warp == 0 && send) {
                    // the sub-tile's MMAs of this step have completed: arm its `h_full` barrier for the next step's rows (8 ranks x
                    // (hi + lo)) and tell all 8 CTAs that this CTA's tensor core no longer reads them (relaxed: no data rides on it)
                    if (lane == 0) mbar_arrive_expect_tx(bar_full + 8 * sub, (uint32_t)TCC * KBLOCK);
                    if (lane < TCC) mbar_arrive_remote_relaxed(mapa_u32(bar_free + 8 * sub, lane));
                }
                float dm[8], dc[8], g8[8];
#pragma unroll
                for (int q = 0; q < 2; ++q) {
                    tmem_ld4(d_main + lane_base + (2 * sub + q) * 16 + part * 4, dm + 4 * q);
                    tmem_ld4(d_corr + lane_base + (2 * sub + q) * 16 + part * 4, dc + 4 * q);
                }
                if (GTMA) {
                    mbar_wait(bar_gin + 8 * sub, par);         // this step's box has landed (row = sequence, 128 floats: lane = column)
#pragma unroll
                    for (int i = 0; i < 8; ++i)
                        g8[i] = *reinterpret_cast<const float*>(g_gin + (size_t)sub * GIN_SUB + (size_t)((i >> 2) * 16 + part * 4 + (i & 3)) * 512 + m * 4);
                } else {
#pragma unroll
                    for (int i = 0; i < 8; ++i) g8[i] = gi[(2 * sub + (i >> 2)) * 4 + (i & 3)];
                }
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                float c_nw[2], h_nw[2];
                unsigned short h_hi16[2], h_lo16[2];
#pragma unroll
                for (int q = 0; q < 2; ++q) {
                    const int blk = 2 * sub + q;
                    float a[4];
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const float pre = fmaf(dc[q * 4 + j], kLoInv, dm[q * 4 + j]) + g8[q * 4 + j];
                        a[j] = WF_SKIP(p, 4) ? pre : FAST ? act_fast(pre, gate == 2) : act_exact(pre, gate == 2);
                    }
                    // 4 x 4 transpose inside the gate quad: lane `gate` ends up with i, f, g, o of sequence 16 blk + 4 part + gate
                    const bool b0 = lane & 1, b1 = lane & 2;
                    const float r0 = __shfl_xor_sync(0xffffffffu, b0 ? a[0] : a[1], 1), r1 = __shfl_xor_sync(0xffffffffu, b0 ? a[2] : a[3], 1);
                    const float u0 = b0 ? r0 : a[0], u1 = b0 ? a[1] : r0, u2 = b0 ? r1 : a[2], u3 = b0 ? a[3] : r1;
                    const float q0 = __shfl_xor_sync(0xffffffffu, b1 ? u0 : u2, 2), q1 = __shfl_xor_sync(0xffffffffu, b1 ? u1 : u3, 2);
                    const float iv = b1 ? q0 : u0, fv = b1 ? q1 : u1, gv = b1 ? u2 : q0, ov = b1 ? u3 : q1;
                    c_nw[q] = fmaf(fv, cst[blk], iv * gv);
                    h_nw[q] = ov * (FAST ? act_fast(c_nw[q], true) : act_exact(c_nw[q], true));
                    cst[blk] = c_nw[q];
                    split16(h_nw[q], h_hi16[q], h_lo16[q]);
                    if (send || p.y_tma) {
                        const uint32_t so = (uint32_t)sub * KBLOCK + sw64_off(q * 16 + part * 4 + gate, ul);
                        *reinterpret_cast<unsigned short*>(stg + so) = h_hi16[q];
                        *reinterpret_cast<unsigned short*>(stg + so + PLANE) = h_lo16[q];
                    }
                }
                if (send || p.y_tma) {
                    // the exchange first, global traffic after (fence.proxy.async waits for every earlier memory operation of the thread)
                    fence_proxy_async_smem();
                    __syncwarp();
                    if (lane == 0) mbar_arrive_local(bar_stage + 8 * sub);
                    if (tid == 0) WF_STAMP(8 + sub);
                }
#pragma unroll
                for (int q = 0; q < 2; ++q) {
                    const int blk = 2 * sub + q;
                    const size_t yo = (size_t)(blk * 16) * TY2;
                    if (WF_SKIP(p, 1) || p.y_tma) {
                    } else if (p.y_split) {
                        yh[yo] = h_hi16[q];
                        yl[yo] = h_lo16[q];
                    } else {
                        yp[yo] = h_nw[q];
                    }
                    if (!send) {
                        const int n = blk * 16 + part * 4 + gate;
                        if (p.hn) p.hn[((size_t)dir * p.B + b_begin + n) * TH + rank * TUC + ul] = h_nw[q];
                        if (p.cn) p.cn[((size_t)dir * p.B + b_begin + n) * TH + rank * TUC + ul] = c_nw[q];
                    } else if (!GTMA && !WF_SKIP(p, 2)) {
#pragma unroll
                        for (int j = 0; j < 4; ++j) gi[blk * 4 + j] = __ldg(gp + (size_t)(blk * 16 + j) * TG4);
                    }
                }
            }
            g_run += dG4;
            y_run += dY2;
        }
    }

    if (warp > EPI_WARPS && lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");      // this thread's y stores are complete
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();
    if (warp == EPI_WARPS) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

// one plane of the split layer output, [B, T, Y2] halves, as a 3-D tensor map with a (32 units, 1 frame, 32 sequences) box whose
// shared-memory image is the staged slice (64-byte rows, 64B swizzle)
EncodeTiledFn encode_tiled() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* q = nullptr;
        cudaDriverEntryPointQueryResult r;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &q, cudaEnableDefault, &r) == cudaSuccess && r == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(q);
    }
    return fn;
}

int make_map_y(CUtensorMap* map, void* ptr, int B, int T, int Y2) {
    EncodeTiledFn fn = encode_tiled();
    if (!fn) {
        set_error("lstm_f16w: cuTensorMapEncodeTiled is not available from this driver");
        return MP_ERR_CUDA;
    }
    const cuuint64_t dims[3] = {(cuuint64_t)Y2, (cuuint64_t)T, (cuuint64_t)B};
    const cuuint64_t strides[2] = {(cuuint64_t)Y2 * 2, (cuuint64_t)T * Y2 * 2};
    const cuuint32_t box[3] = {(cuuint32_t)TUC, 1, (cuuint32_t)WR};
    const cuuint32_t estr[3] = {1, 1, 1};
    const CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3, ptr, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                          CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("lstm_f16w: cuTensorMapEncodeTiled (layer output) failed with CUresult %d (B=%d T=%d Y2=%d)", (int)r, B, T, Y2);
        return MP_ERR_CUDA;
    }
    return MP_OK;
}

// gin [B, T, G4] floats as a 3-D tensor map with a (128 gate columns, 1 frame, 32 sequences) box, no swizzle (512-byte rows)
int make_map_gin(CUtensorMap* map, const float* ptr, int B, int T, int G4) {
    EncodeTiledFn fn = encode_tiled();
    if (!fn) {
        set_error("lstm_f16w: cuTensorMapEncodeTiled is not available from this driver");
        return MP_ERR_CUDA;
    }
    const cuuint64_t dims[3] = {(cuuint64_t)G4, (cuuint64_t)T, (cuuint64_t)B};
    const cuuint64_t strides[2] = {(cuuint64_t)G4 * 4, (cuuint64_t)T * G4 * 4};
    const cuuint32_t box[3] = {128, 1, (cuuint32_t)WR};
    const cuuint32_t estr[3] = {1, 1, 1};
    const CUresult res = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(ptr), dims, strides, box, estr,
                            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (res != CUDA_SUCCESS) {
        set_error("lstm_f16w: cuTensorMapEncodeTiled (gate pre-activations) failed with CUresult %d (B=%d T=%d G4=%d)", (int)res, B, T, G4);
        return MP_ERR_CUDA;
    }
    return MP_OK;
}

}  // namespace

// whole 128-sequence tiles of equal length, H = 256, both ends of the exchange inside one 8-CTA cluster; MP_REC_WIDE=0 keeps the
// 64-sequence kernel, =1 takes this one whenever the shape allows (default: when the tile policy asks for 128, i.e. the pipelined path)
bool rec_f16w_eligible(const RecLayerArgs& a) {
    const char* v = getenv("MP_REC_WIDE");
    if (v && atoi(v) == 0) return false;
    if (a.H != TH || !a.w_raw[0] || a.lengths != nullptr || a.B % WN != 0 || a.T < 2) return false;
    if (!a.y_split) return false;      // the layer output leaves through the TMA stores of the (hi, lo) planes; fp32 output keeps the 64-sequence kernel
    if (getenv("MP_RF16_RAGGED") || getenv("MP_RTC_TS") || getenv("MP_REC_NB")) return false;
    return (v && atoi(v) != 0) || a.tile_hint == WN;
}

int launch_lstm_recurrence_f16w(const RecLayerArgs& a, cudaStream_t stream) {
    MP_REQUIRE((double)a.B * a.T * a.dirs * 4 * a.H < 4.0e9, "lstm_f16w: B*T = %lld frames exceeds 32-bit gate buffer indexing", (long long)a.B * a.T);
    static long long* ts_dev = nullptr;
    const bool want_ts = getenv("MP_RECW_TS") != nullptr;
    if (want_ts && !ts_dev) {
        cudaMalloc(&ts_dev, 64 * 16 * sizeof(long long));
        cudaMemset(ts_dev, 0, 64 * 16 * sizeof(long long));
    }
    RecF16WParams p{a.gin, a.w_raw[0], a.w_raw[a.dirs - 1], a.y, a.h0, a.c0, a.hn, a.cn, a.B, a.T, a.dirs, a.y_split, want_ts ? ts_dev : nullptr,
                    getenv("MP_RECW_SKIP") ? atoi(getenv("MP_RECW_SKIP")) : 0, 0};
    alignas(64) CUtensorMap map_hi, map_lo, map_gin;
    memset(&map_hi, 0, sizeof(map_hi));
    memset(&map_lo, 0, sizeof(map_lo));
    memset(&map_gin, 0, sizeof(map_gin));
    const bool gtma = !getenv("MP_RECW_NO_TMA_GIN");
    if (gtma) MP_TRY(make_map_gin(&map_gin, a.gin, a.B, a.T, a.dirs * 4 * a.H));
    if (a.y_split && !getenv("MP_RECW_NO_TMA_Y")) {
        const int Y2 = a.dirs * a.H;
        MP_TRY(make_map_y(&map_hi, a.y, a.B, a.T, Y2));
        MP_TRY(make_map_y(&map_lo, reinterpret_cast<unsigned short*>(a.y) + (size_t)a.B * a.T * Y2, a.B, a.T, Y2));
        p.y_tma = 1;
    }
    ProfileScope prof("lstm_rec_f16_h256", 4.0 * ((double)a.dirs * 4 * a.H * a.H + (double)a.B * a.T * a.dirs * a.H), stream);
    const char* actv = getenv("MP_RF16_ACT");
    const bool fast = !(actv && strcmp(actv, "exact") == 0);
    static bool configured = false;
    if (!configured) {
        MP_CUDA_TRY(cudaFuncSetAttribute(lstm_rec_f16w_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)WF_SMEM));
        MP_CUDA_TRY(cudaFuncSetAttribute(lstm_rec_f16w_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)WF_SMEM));
        MP_CUDA_TRY(cudaFuncSetAttribute(lstm_rec_f16w_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)WF_SMEM));
        MP_CUDA_TRY(cudaFuncSetAttribute(lstm_rec_f16w_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)WF_SMEM));
        configured = true;
    }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(TCC * (a.B / WN), a.dirs, 1);
    cfg.blockDim = dim3(WF_THREADS, 1, 1);
    cfg.dynamicSmemBytes = WF_SMEM;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = TCC;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    if (fast && gtma) MP_CUDA_TRY(cudaLaunchKernelEx(&cfg, lstm_rec_f16w_kernel<true, true>, p, map_hi, map_lo, map_gin));
    else if (fast) MP_CUDA_TRY(cudaLaunchKernelEx(&cfg, lstm_rec_f16w_kernel<true, false>, p, map_hi, map_lo, map_gin));
    else if (gtma) MP_CUDA_TRY(cudaLaunchKernelEx(&cfg, lstm_rec_f16w_kernel<false, true>, p, map_hi, map_lo, map_gin));
    else MP_CUDA_TRY(cudaLaunchKernelEx(&cfg, lstm_rec_f16w_kernel<false, false>, p, map_hi, map_lo, map_gin));
    count_launch();
    if (want_ts) {      // bring-up only: synchronous dump of the stamps of block (0,0)
        static long long h[64 * 16];
        cudaStreamSynchronize(stream);
        cudaMemcpy(h, ts_dev, sizeof(h), cudaMemcpyDeviceToHost);
        for (int s = 3; s < 8 && s < a.T - 1; ++s) {
            const long long t0 = h[s * 16];
            fprintf(stderr, "[recw ts] s=%d step %lld |", s, h[(s + 1) * 16] - t0);
            for (int sub = 0; sub < 4; ++sub)
                fprintf(stderr, " sub%d: mma issue +%lld, epi start +%lld, staged +%lld, copies out +%lld |", sub, h[s * 16 + sub] - t0,
                        h[s * 16 + 4 + sub] - t0, h[s * 16 + 8 + sub] - t0, h[s * 16 + 12 + sub] - t0);
            fprintf(stderr, "\n");
        }
    }
    return MP_OK;
}

}  // namespace mp
