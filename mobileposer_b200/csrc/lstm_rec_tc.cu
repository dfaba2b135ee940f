// Tensor-core LSTM recurrence (H = 256, throughput path): the per-step product  G[4H/C x NB] = W_hh_slice . h^T
// runs on tcgen05 with fp32-grade accuracy (3xTF32), W_hh resident in TENSOR MEMORY for the whole sequence.
//
// Same contract as lstm_rec_kernel<256, 8, 4> (lstm_rec.cu): one launch walks all T steps of one LSTM layer for a
// tile of NB sequences per 8-CTA cluster; CTA `rank` owns hidden units [32 rank, 32 rank + 32) = 128 gate rows.
// What changes is where the FLOPs go.  The FFMA kernel is bound by issue slots (reduce-scatter + activations
// around every 256 FMAs); here the 128 x N x 256 product of a step is 96 UMMA instructions issued by one thread:
//
//   operands     A = W_hh slice [128 rows (row m = unit_local*4 + gate) x K = 256], split once at kernel start into
//                    hi = w & 0xFFFFE000 and lo = rn_tf32(w - hi):
//                      W_hi  -> TMEM columns [0, 256)            (lane = row, column = k; tcgen05.st)
//                      W_lo  -> TMEM columns [384, 512) for k < 128, shared memory (K-major, 128B swizzle) for k >= 128
//                B = h_t [N sequences x K = 256] as hi / lo, K-major 128B-swizzled in shared memory; K-block `r`
//                    (32 floats) of every row is exactly the slice produced by cluster rank r, so a rank pushes its
//                    new slice to a peer with ONE cp.async.bulk per array
//                D = two fp32 accumulators in TMEM: main (W_hi h_hi) columns [256, 256+N), correction
//                    (W_lo h_hi + W_hi h_lo) columns [256+N, 256+2N)
//   per step     MMA warp: wait h_t complete -> 96 x tcgen05.mma.kind::tf32 (A from TMEM or smem) -> tcgen05.commit
//                16 epilogue warps: wait commit -> tcgen05.ld both accumulators -> + gate pre-activation (gin) ->
//                    sigmoid/tanh -> i/f/g/o gather by shuffle (the 4 gates of a unit are adjacent TMEM lanes) ->
//                    cell update (c in registers) -> y store, hi/lo of h_{t+1} into the staging slice ->
//                    bulk copies to all 8 CTAs (transaction bytes on their `h_full` mbarrier)
//   h is single buffered: a rank may overwrite a peer's h only after that peer's MMAs of the step have finished, which
//   every CTA announces with a remote mbarrier arrive on all peers' `h_free` barrier right after its commit lands.
#include "mp_common.cuh"

#include <cstdlib>
#include <cstring>

namespace mp {

namespace {

constexpr int TH = 256;
constexpr int TCC = 8;            // cluster size
constexpr int TUC = TH / TCC;     // 32 units per CTA
constexpr int EPI_WARPS = 16;
constexpr int RTC_THREADS = (EPI_WARPS + 2) * 32;   // 16 epilogue warps + MMA warp + copy warp
constexpr uint32_t COL_WHI = 0, COL_D = 256, COL_WLO = 384;
constexpr int WLO_TMEM_K = 128;                       // k < 128 of W_lo lives in TMEM
constexpr uint32_t WLO_S_BYTES = (TH - WLO_TMEM_K) / 32 * 128 * 128;   // 4 K-blocks x 128 rows x 128 B = 64 KiB

struct RecTcParams {
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
    long long* ts;   // bring-up: per-step clock64 stamps of block (0,0) [step][8], or null (MP_RTC_TS)
};

__host__ __device__ inline size_t rec_tc_smem_bytes(int N) {
    // W_lo half | H_hi | H_lo | staging [2 parity][2 arrays] | tables | barriers
    return 1024 + WLO_S_BYTES + (size_t)2 * 8 * N * 128 + (size_t)4 * N * 128 + (size_t)3 * N * 4 + 128;
}

__device__ __forceinline__ uint32_t tf32_hi(float x) { return __float_as_uint(x) & 0xFFFFE000u; }
__device__ __forceinline__ uint32_t tf32_lo(float x, uint32_t hi) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x - __uint_as_float(hi)));
    return r;
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
// D[tmem] (+)= A[tmem] * B[smem desc]
__device__ __forceinline__ void umma_ts(uint32_t d, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d),
        "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(acc)
        : "memory");
}
// D[tmem] (+)= A[smem desc] * B[smem desc]
__device__ __forceinline__ void umma_ss(uint32_t d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(d),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc)
        : "memory");
}
// K-major operand, 128-byte rows, SWIZZLE_128B: 8-row groups 1024 B apart
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t smem_addr) {
    return (uint64_t)((smem_addr & 0x3FFFF) >> 4) | (1ull << 16) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
}
// byte offset of float `k32` (0..31) of row `row` inside one [rows x 128 B] K-block with the 128B swizzle
__device__ __forceinline__ uint32_t sw128_off(int row, int k32) {
    return (uint32_t)row * 128u + (uint32_t)((((k32 >> 2) ^ (row & 7)) << 4) | ((k32 & 3) << 2));
}
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t (&v)[32]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
        "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
        "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};" ::"r"(taddr),
        "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]),
        "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]), "r"(v[16]), "r"(v[17]), "r"(v[18]),
        "r"(v[19]), "r"(v[20]), "r"(v[21]), "r"(v[22]), "r"(v[23]), "r"(v[24]), "r"(v[25]), "r"(v[26]), "r"(v[27]),
        "r"(v[28]), "r"(v[29]), "r"(v[30]), "r"(v[31])
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

__device__ __forceinline__ float act_sigmoid_or_tanh(float x, bool is_tanh) {
    const float s = fminf(fmaxf(is_tanh ? 2.0f * x : x, -30.0f), 30.0f);
    // expf, not ex2.approx on s log2(e): the MUFU form is as accurate for a sigmoid (measured: max |tc - ffma| 8.2e-8 either way)
    // and 9 instructions shorter, but an A/B on the same box showed no difference (1.50 ms per launch both): after the
    // reciprocal / transpose / offset diet the epilogue is bound by its dependent chains, not by issue slots.
    const float e = expf(-s);
    // 1 / (1 + e): MUFU reciprocal + one Newton step (<= 1 ulp; 1 + e is in [1, 1e13], no special cases) -- the correctly
    // rounded __frcp_rn cost as many instructions as expf itself (ncu: 15 % of the kernel's instructions)
    const float d = 1.0f + e;
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(d));
    r = fmaf(fmaf(-d, r, 1.0f), r, r);
    return is_tanh ? (1.0f - e) * r : r;
}

#define RTC_STAMP(slot)                                                                                  \
    do {                                                                                                  \
        if (p.ts && blockIdx.x == 0 && blockIdx.y == 0 && s < 64) p.ts[s * 8 + (slot)] = clock64();       \
    } while (0)

// N = padded sequence count of the tile (multiple of 16, <= 64); SPW = N / 4 sequences per epilogue warp
template <int N>
__global__ void __launch_bounds__(RTC_THREADS, 1) lstm_rec_tc_kernel(const RecTcParams p) {
    constexpr int SPW = N / 4;
    constexpr uint32_t HBYTES = 8u * N * 128u;            // one h array: 8 K-blocks x N rows x 128 B
    constexpr uint32_t SLICE = (uint32_t)N * 128u;        // one K-block = one rank's slice
    static_assert(N % 16 == 0 && N <= 64 && COL_D + 2 * N <= COL_WLO, "tile");

    extern __shared__ unsigned char smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    unsigned char* gen = smem_raw + (base - smem_u32(smem_raw));
    const uint32_t s_wlo = base, s_hhi = s_wlo + WLO_S_BYTES, s_hlo = s_hhi + HBYTES, s_stg = s_hlo + HBYTES;
    unsigned char* g_wlo = gen;
    unsigned char* g_hhi = g_wlo + WLO_S_BYTES;
    unsigned char* g_hlo = g_hhi + HBYTES;
    unsigned char* g_stg = g_hlo + HBYTES;                // [par][arr][N rows][128 B]
    int* lens = reinterpret_cast<int*>(g_stg + 4 * SLICE);
    uint32_t* goff = reinterpret_cast<uint32_t*>(lens + N);
    uint32_t* yoff = goff + N;
    const uint32_t s_bars = s_stg + 4 * SLICE + 3 * N * 4;
    // per sub-tile: bar_full[2] (h rows arrived), bar_mma[2] (MMAs committed), bar_free[2] (all peers' MMAs done);
    // bar_chunk[0..3]: "rows [16 b, 16 b + 16) of the new slice are staged"
    const uint32_t bar_full = (s_bars + 7u) & ~7u, bar_mma = bar_full + 16, bar_free = bar_full + 32, bar_chunk = bar_full + 48,
                   tmem_slot = bar_full + 48 + 8 * 4;

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
        }
        for (int b = 0; b < 4; ++b) mbar_init(bar_chunk + 8 * b, EPI_WARPS * 32);
        mbar_fence_init_cluster();
    }
    if (warp == EPI_WARPS) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    for (int i = tid; i < N; i += RTC_THREADS) {
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

    // ---- W_hh slice -> TMEM / shared memory as TF32 hi + lo (once) ---------------------------------------------
    if (warp < EPI_WARPS) {
        const float* wrow = (dir ? p.w1 : p.w0) + (size_t)(gate * TH + rank * TUC + ul) * TH;
        for (int kc = part; kc < TH / 32; kc += 4) {
            const int k0 = kc * 32;
            uint32_t hi[32], lo[32];
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const float4 v = __ldg(reinterpret_cast<const float4*>(wrow + k0) + i);
                hi[4 * i + 0] = tf32_hi(v.x); lo[4 * i + 0] = tf32_lo(v.x, hi[4 * i + 0]);
                hi[4 * i + 1] = tf32_hi(v.y); lo[4 * i + 1] = tf32_lo(v.y, hi[4 * i + 1]);
                hi[4 * i + 2] = tf32_hi(v.z); lo[4 * i + 2] = tf32_lo(v.z, hi[4 * i + 2]);
                hi[4 * i + 3] = tf32_hi(v.w); lo[4 * i + 3] = tf32_lo(v.w, hi[4 * i + 3]);
            }
            tmem_st32(tmem + lane_base + COL_WHI + k0, hi);
            if (k0 < WLO_TMEM_K) {
                tmem_st32(tmem + lane_base + COL_WLO + k0, lo);
            } else {
                unsigned char* blk = g_wlo + (size_t)((k0 - WLO_TMEM_K) / 32) * (128 * 128);
#pragma unroll
                for (int c = 0; c < 8; ++c)
                    *reinterpret_cast<uint4*>(blk + sw128_off(m, c * 4)) = make_uint4(lo[4 * c], lo[4 * c + 1], lo[4 * c + 2], lo[4 * c + 3]);
            }
        }
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    }
    // ---- h_0 (hi / lo) for the whole tile, zeros in the padded rows ---------------------------------------------
    for (int i = tid; i < N * (TH / 4); i += RTC_THREADS) {
        const int n = i / (TH / 4), ch = i % (TH / 4);       // float4 chunk ch covers k = 4 ch .. 4 ch + 3
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (p.h0 && n < nb) v = __ldg(reinterpret_cast<const float4*>(p.h0 + ((size_t)dir * p.B + b_begin + n) * TH) + ch);
        const uint32_t off = (uint32_t)(ch >> 3) * SLICE + sw128_off(n, (ch & 7) * 4);
        uint4 h, l;
        h.x = tf32_hi(v.x); h.y = tf32_hi(v.y); h.z = tf32_hi(v.z); h.w = tf32_hi(v.w);
        l.x = tf32_lo(v.x, h.x); l.y = tf32_lo(v.y, h.y); l.z = tf32_lo(v.z, h.z); l.w = tf32_lo(v.w, h.w);
        *reinterpret_cast<uint4*>(g_hhi + off) = h;
        *reinterpret_cast<uint4*>(g_hlo + off) = l;
    }
    // cell state of this thread's (unit, sequences): every one of the 4 gate lanes of a unit keeps a copy
    // cell state: lane `gate` of a unit's quad owns sequence 4*blk + gate of every block of 4 sequences
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
    // the running offset of the NEXT frame of sequence j (one add per step instead of rebuilding it from the tables)
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
        so[blk] = sw128_off(n, ul);
        yo[blk] = yoff[n] + (uint32_t)(dir ? max(len_own[blk] - 1, 0) : 0) * (uint32_t)Y2 + (uint32_t)ul;
    }

    const uint32_t d_main = tmem + COL_D, d_corr = tmem + COL_D + N;

    // The tile is split into two sub-tiles of whole 16-sequence blocks that ping-pong through the step: while the
    // epilogue warps run the activations of one sub-tile the tensor core already works on the other, and the exchange of
    // a finished sub-tile hides behind both.  Each sub-tile has its own full / mma / free barriers.
    constexpr int NBLK = N / 16;
    constexpr int NSUB = NBLK >= 2 ? 2 : 1;
    constexpr int BLK_A = (NBLK + 1) / 2;

    for (int s = 0; s < maxlen; ++s) {
        const bool send = (s + 1 < maxlen);
        const int par = s & 1;
        if (warp == EPI_WARPS) {
            // ================= MMA issuer =================
            // The whole warp walks the (fully unrolled) issue sequence and one elected lane executes each tcgen05.mma:
            // in warp-uniform control flow the descriptors live in uniform registers; under `if (lane == 0)` every
            // UMMA was wrapped in an ELECT / R2UR / BRA.U.ANY loop (48 clk per issue, 4.6 k clk per step).
            const bool leader = elect_one();
#pragma unroll
            for (int sub = 0; sub < NSUB; ++sub) {
                const int blk0 = sub == 0 ? 0 : BLK_A;
                const int nblk = sub == 0 ? BLK_A : NBLK - BLK_A;
                const uint32_t row0 = 16u * blk0;                       // first sequence row / accumulator column of the sub-tile
                const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)((16 * nblk) >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
                if (s > 0) mbar_wait(bar_full + 8 * sub, (s - 1) & 1);
                tc_fence_after();
                if (leader && sub == 0) RTC_STAMP(0);
                // descriptors advance by compile-time constants (fully unrolled): one 64-bit add per operand
                const uint64_t bd_hi = umma_desc_sw128(s_hhi + row0 * 128u), bd_lo = umma_desc_sw128(s_hlo + row0 * 128u),
                               ad_wlo = umma_desc_sw128(s_wlo);
                const uint32_t dm_ = d_main + row0, dc_ = d_corr + row0;
#pragma unroll
                for (int ks = 0; ks < TH / 8; ++ks) {          // correction: W_lo . h_hi
                    const uint64_t b = bd_hi + (uint64_t)(((ks >> 2) * SLICE + (ks & 3) * 32) >> 4);
                    if (ks < WLO_TMEM_K / 8) { if (leader) umma_ts(dc_, tmem + COL_WLO + ks * 8, b, idesc, ks != 0); }
                    else if (leader)
                        umma_ss(dc_, ad_wlo + (uint64_t)((((ks - WLO_TMEM_K / 8) >> 2) * (128 * 128) + (ks & 3) * 32) >> 4), b, idesc, 1u);
                }
#pragma unroll
                for (int ks = 0; ks < TH / 8; ++ks)            // correction: W_hi . h_lo
                    if (leader) umma_ts(dc_, tmem + COL_WHI + ks * 8, bd_lo + (uint64_t)(((ks >> 2) * SLICE + (ks & 3) * 32) >> 4), idesc, 1u);
#pragma unroll
                for (int ks = 0; ks < TH / 8; ++ks)            // main: W_hi . h_hi
                    if (leader) umma_ts(dm_, tmem + COL_WHI + ks * 8, bd_hi + (uint64_t)(((ks >> 2) * SLICE + (ks & 3) * 32) >> 4), idesc, ks != 0);
                if (leader) {
                    if (sub == NSUB - 1) RTC_STAMP(1);
                    tc_commit(bar_mma + 8 * sub);
                }
            }
            __syncwarp();
        } else if (warp == EPI_WARPS + 1) {
            // ================= copy warp =================
            // ships rows [16 b, 16 b + 16) of the new slice to all 8 CTAs as soon as the 16 epilogue warps have staged
            // them, so the exchange overlaps the activation work of the following blocks and costs them no barrier
            if (send && lane < 2 * TCC) {
                const int r = lane >> 1, arr = lane & 1;
                const uint32_t chunk = 16u * 128u;
#pragma unroll
                for (int blk = 0; blk < NBLK; ++blk) {
                    const int sub = blk < BLK_A ? 0 : 1;
                    mbar_wait(bar_chunk + 8 * blk, par);
                    if (blk == 0 || blk == BLK_A) mbar_wait(bar_free + 8 * sub, par);   // every peer's MMAs on this sub-tile are done
                    bulk_copy_s2c(mapa_u32((arr ? s_hlo : s_hhi) + (uint32_t)rank * SLICE + blk * chunk, r),
                                  s_stg + (uint32_t)(par * 2 + arr) * SLICE + blk * chunk, chunk, mapa_u32(bar_full + 8 * sub, r));
                }
            }
            __syncwarp();
        } else {
            // ================= epilogue =================
            unsigned char* stg_hi = g_stg + (size_t)(par * 2 + 0) * SLICE;
            unsigned char* stg_lo = g_stg + (size_t)(par * 2 + 1) * SLICE;
#pragma unroll
            for (int sub = 0; sub < NSUB; ++sub) {
                const int blk0 = sub == 0 ? 0 : BLK_A;
                const int nblk = sub == 0 ? BLK_A : NBLK - BLK_A;
                mbar_wait(bar_mma + 8 * sub, par);
                tc_fence_after();
                // h_{s+1} of this sub-tile will arrive as 8 ranks x (hi, lo) x nblk chunks.  Armed only now: MMA(s) has run,
                // so the MMA warp has seen the previous phase of the barrier complete -- arming earlier could put two
                // arrivals into one phase.
                if (tid == 0 && send) mbar_arrive_expect_tx(bar_full + 8 * sub, 2u * TCC * (uint32_t)nblk * 16u * 128u);
                if (tid == 0 && sub == 0) RTC_STAMP(2);
                if (send && tid < TCC) mbar_arrive_remote(mapa_u32(bar_free + 8 * sub, tid));     // my MMAs no longer read these rows
                float dm[4 * (NBLK - NBLK / 2)], dc[4 * (NBLK - NBLK / 2)];
#pragma unroll
                for (int q = 0; q < nblk; ++q) {
                    tmem_ld4(d_main + lane_base + (blk0 + q) * 16 + part * 4, dm + 4 * q);
                    tmem_ld4(d_corr + lane_base + (blk0 + q) * 16 + part * 4, dc + 4 * q);
                }
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                if (tid == 0 && sub == 0) RTC_STAMP(3);
                // Blocks of 4 sequences per lane quad: every lane evaluates its gate for the 4 sequences, the quad (4 gate
                // lanes of a unit) exchanges them, and lane g then owns sequence 4*blk + g: ONE cell update per lane.
                // Phase 1 is register-only for all blocks of the sub-tile, so their dependent chains (activation -> transpose ->
                // cell -> tanh) interleave; the stores, fences and barrier arrivals (compiler barriers) come after, per block.
                constexpr int MAXB = NBLK - NBLK / 2;
                float c_nw[MAXB], h_nw[MAXB];
#pragma unroll
                for (int bq = 0; bq < nblk; ++bq) {
                    const int blk = blk0 + bq;
                    float a[4];
#pragma unroll
                    for (int q = 0; q < 4; ++q) a[q] = act_sigmoid_or_tanh((dm[bq * 4 + q] + dc[bq * 4 + q]) + gi[blk * 4 + q], gate == 2);
                    // 4 x 4 transpose inside the gate quad (lanes = gates i, f, g, o of one unit; registers = the 4 sequences):
                    // two butterfly rounds, 4 shuffles instead of 16
                    const bool b0 = lane & 1, b1 = lane & 2;
                    const float r0 = __shfl_xor_sync(0xffffffffu, b0 ? a[0] : a[1], 1), r1 = __shfl_xor_sync(0xffffffffu, b0 ? a[2] : a[3], 1);
                    const float u0 = b0 ? r0 : a[0], u1 = b0 ? a[1] : r0, u2 = b0 ? r1 : a[2], u3 = b0 ? a[3] : r1;
                    const float q0 = __shfl_xor_sync(0xffffffffu, b1 ? u0 : u2, 2), q1 = __shfl_xor_sync(0xffffffffu, b1 ? u1 : u3, 2);
                    const float iv = b1 ? q0 : u0, fv = b1 ? q1 : u1, gv = b1 ? u2 : q0, ov = b1 ? u3 : q1;
                    c_nw[bq] = fmaf(fv, cst[blk], iv * gv);
                    h_nw[bq] = ov * act_sigmoid_or_tanh(c_nw[bq], true);
                }
#pragma unroll
                for (int bq = 0; bq < nblk; ++bq) {
                    const int blk = blk0 + bq;
                    // this lane's sequence of the block
                    const int n = blk * 16 + part * 4 + gate;
                    const int len = len_own[blk];
                    const bool active = s < len;
                    const float c_new = c_nw[bq], h_new = h_nw[bq];
                    if (active) {
                        cst[blk] = c_new;
                        p.y[yo[blk]] = h_new;
                        if (s == len - 1) {
                            if (p.hn) p.hn[((size_t)dir * p.B + b_begin + n) * TH + rank * TUC + ul] = h_new;
                            if (p.cn) p.cn[((size_t)dir * p.B + b_begin + n) * TH + rank * TUC + ul] = c_new;
                        }
                    }
                    yo[blk] += dY2;
                    if (send) {
                        const float hv = active ? h_new : 0.f;
                        const uint32_t hh = tf32_hi(hv);
                        *reinterpret_cast<uint32_t*>(stg_hi + so[blk]) = hh;
                        *reinterpret_cast<uint32_t*>(stg_lo + so[blk]) = tf32_lo(hv, hh);
                        // rows [16 blk, 16 blk + 16) of the new slice are staged: hand them to the copy warp (non-blocking)
                        fence_proxy_async_smem();
                        mbar_arrive_local(bar_chunk + 8 * blk);
                    }
                    // next step's gate pre-activations of the block (this lane's gate, all 4 sequences)
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        if (s + 1 < lens[blk * 16 + part * 4 + q]) gi[blk * 4 + q] = __ldg(p.gin + go[blk * 4 + q]);
                        go[blk * 4 + q] += dG4;
                    }
                }
            }
            if (tid == 0) RTC_STAMP(4);
        }
    }

    // frames >= len of the layer output are zero (pad_packed_sequence, rnn.py:31)
    for (int b = 0; b < nb; ++b) {
        const int len = lens[b];
        const int cnt = (p.T - len) * TUC;
        for (int i = tid; i < cnt; i += RTC_THREADS) {
            const int t = len + i / TUC, u = i % TUC;
            p.y[yoff[b] + (uint32_t)t * (uint32_t)Y2 + (uint32_t)u] = 0.f;
        }
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();
    if (warp == EPI_WARPS) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
}

template <int N>
int launch_tc_n(const RecTcParams& p, cudaStream_t stream) {
    const size_t smem = rec_tc_smem_bytes(N);
    auto kern = lstm_rec_tc_kernel<N>;
    static bool configured = false;
    if (!configured) {
        MP_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured = true;
    }
    const int n_tiles = (p.B + p.NB - 1) / p.NB;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(TCC * n_tiles, p.dirs, 1);
    cfg.blockDim = dim3(RTC_THREADS, 1, 1);
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

int tc_cluster_slots() {
    static int slots = 0;
    if (slots > 0) return slots;
    int n = 0;
    auto kern = lstm_rec_tc_kernel<48>;
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)rec_tc_smem_bytes(48));
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(TCC * 64, 1, 1);
    cfg.blockDim = dim3(RTC_THREADS, 1, 1);
    cfg.dynamicSmemBytes = rec_tc_smem_bytes(48);
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

// Tensor-core recurrence for H = 256 when the batch is large enough to fill 16-wide MMA tiles.
bool rec_tc_eligible(const RecLayerArgs& a) {
    const char* v = getenv("MP_REC_IMPL");
    if (v && (strcmp(v, "ffma") == 0 || strcmp(v, "simple") == 0)) return false;
    if (a.H != TH || !a.w_raw[0]) return false;
    const bool forced = v && (strcmp(v, "tc") == 0 || strcmp(v, "tf32") == 0);
    return forced || a.B * a.dirs > 2 * tc_cluster_slots();
}

int launch_lstm_recurrence_tc(const RecLayerArgs& a, cudaStream_t stream) {
    MP_REQUIRE((double)a.B * a.T * a.dirs * 4 * a.H < 4.0e9, "lstm_tc: B*T = %lld frames exceeds 32-bit gate buffer indexing", (long long)a.B * a.T);
    const char* nbv = getenv("MP_REC_NB");
    int NB = (nbv && *nbv) ? atoi(nbv) : a.tile_hint;
    if (NB == 0) {
        // auto: a step of this kernel costs about the same for 16 or 64 sequences per cluster (ncu, cfg3: 1.49 ms with 14
        // clusters of 37, 1.45 ms with 8 clusters of 64), so large batches take the fullest tile and leave the other SMs to
        // whatever runs beside them (the other heads, other batches); small batches spread over one wave of clusters
        if (a.B >= 128) {
            NB = 64;
        } else {
            const int per = std::max(1, tc_cluster_slots() / a.dirs);
            NB = (a.B + per - 1) / per;
        }
    } else if (NB < 0) {      // -1: the one-wave policy whatever the batch (lowest occupancy per cluster)
        const int per = std::max(1, tc_cluster_slots() / a.dirs);
        NB = (a.B + per - 1) / per;
    }
    NB = std::min(64, std::max(1, NB));
    const int N = ((NB + 15) / 16) * 16;
    // balance the tiles: same tile count, equal sizes
    const int n_tiles = (a.B + NB - 1) / NB;
    NB = (a.B + n_tiles - 1) / n_tiles;
    static long long* ts_dev = nullptr;
    const bool want_ts = getenv("MP_RTC_TS") != nullptr;
    if (want_ts && !ts_dev) {
        cudaMalloc(&ts_dev, 64 * 8 * sizeof(long long));
        cudaMemset(ts_dev, 0, 64 * 8 * sizeof(long long));
    }
    RecTcParams p{a.gin, a.w_raw[0], a.w_raw[a.dirs - 1], a.y, a.h0, a.c0, a.hn, a.cn, a.lengths, a.B, a.T, a.dirs, NB,
                  want_ts ? ts_dev : nullptr};
    ProfileScope prof("lstm_rec_tc_h256", 4.0 * ((double)a.dirs * 4 * a.H * a.H + (double)a.B * a.T * a.dirs * a.H), stream);
    int st;
    switch (N) {
        case 16: st = launch_tc_n<16>(p, stream); break;
        case 32: st = launch_tc_n<32>(p, stream); break;
        case 48: st = launch_tc_n<48>(p, stream); break;
        default: st = launch_tc_n<64>(p, stream); break;
    }
    if (st == MP_OK && want_ts) {      // bring-up only: synchronous dump of the stamps of block (0,0)
        long long h[64 * 8];
        cudaStreamSynchronize(stream);
        cudaMemcpy(h, ts_dev, sizeof(h), cudaMemcpyDeviceToHost);
        fprintf(stderr, "[rtc ts] N=%d NB=%d B=%d\n", N, NB, a.B);
        for (int s = 2; s < 8 && s < a.T - 1; ++s)
            fprintf(stderr, "[rtc ts] s=%d  mma issue %lld | commit->epi %lld  tmem ld %lld  activations+staging %lld  tail->next mma %lld  step %lld\n", s,
                    h[s * 8 + 1] - h[s * 8 + 0], h[s * 8 + 2] - h[s * 8 + 1], h[s * 8 + 3] - h[s * 8 + 2], h[s * 8 + 4] - h[s * 8 + 3],
                    h[(s + 1) * 8 + 0] - h[s * 8 + 4], h[(s + 1) * 8 + 0] - h[s * 8 + 0]);
    }
    return st;
}

}  // namespace mp
