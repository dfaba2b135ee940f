// Persistent LSTM recurrence: one launch walks every timestep of one layer (all directions,
// all sequences), W_hh resident in REGISTERS across timesteps, h exchanged between the CTAs of a
// thread-block cluster through distributed shared memory.
//
// Replaces the time loop inside torch's `_VF.lstm` as configured at mobileposer/models/rnn.py:15,27
// (2 layers, gate rows i,f,g,o, b_ih + b_hh, zero or carried initial state, packed-sequence
// semantics for ragged lengths: rnn.py:24-31).  The input projection W_ih x_t + b_ih + b_hh is
// hoisted into gemm.cu (`gin`); this kernel adds W_hh h_{t-1}, applies the gate nonlinearities
// and the cell update, and emits h_t.
//
// Decomposition (H = hidden, C = cluster size, UC = H / C hidden units per CTA):
//   * CTA `rank` of a cluster owns hidden units [rank*UC, rank*UC+UC) => the 4*UC gate rows of
//     W_hh that produce them, so the cell update never leaves the CTA.
//   * a warp owns 8 gate rows (2 units x 4 gates) and splits the K = H reduction across its lanes:
//     lane kl holds, for R rows, the float4 chunks {kl, kl + KQ, ..} of each row (KQ = 32 lanes for
//     H = 256, 16 for H = 64), 64 (16) weight registers per thread for the whole sequence.  One
//     LDS.128 per lane then fetches 512 DISTINCT bytes of h per warp and feeds R*4 FFMAs -- ncu on the
//     first version (4 lanes per row) showed a broadcast LDS.128 still costs 4 shared-memory
//     wavefronts, which made the recurrence 4x LSU-bound (profiles/r01_notes.md).
//   * the per-lane partial sums (rows x sequences) are combined with a shuffle reduce-scatter whose
//     result layout is lane = unit_in_warp*16 + gate*4 + sequence_in_group, so the i/f/g/o gather and
//     the cell update stay inside the warp -- no shared-memory round trip, no block barrier.
//   * h_{t} lives in shared memory of EVERY CTA of the cluster, double buffered by step parity.
//     After the cell update the owning CTA pushes its UC new values to all C CTAs, either with
//     per-value `st.async ... mbarrier::complete_tx` (latency path, one sequence per cluster) or
//     staged + ONE `cp.async.bulk shared::cluster` per destination (throughput path).  Each CTA waits
//     on its own mbarrier (transaction bytes) -- there is no cluster-wide barrier inside the time loop.
//   * a cluster serves a tile of NB sequences (weights reused NB times per step); the grid is
//     (C * n_tiles, dirs).  Reverse direction of sequence b walks t = len_b-1 .. 0.
#include "mp_common.cuh"

#include <algorithm>
#include <cstdlib>
#include <cstring>

namespace mp {

namespace {

struct RecParams {
    const float* gin;
    const float4* wpack;
    float* y;
    const float* h0;
    const float* c0;
    float* hn;
    float* cn;
    const int32_t* lengths;
    int B, T, dirs, NB, bulk;
};

template <int H, int C>
struct RecCfg {
    static constexpr int UC = H / C;                 // hidden units owned by one CTA
    static constexpr int THREADS = UC * 16;          // one warp per 2 units (8 gate rows)
    static constexpr int KQ = (H >= 128) ? 32 : 16;  // lanes that split K
    static constexpr int RG = 32 / KQ;               // row groups per warp (1 or 2)
    static constexpr int R = 8 / RG;                 // gate rows held by one lane
    static constexpr int CPL = H / 4 / KQ;           // float4 k-chunks per lane and row
    static constexpr int NW4 = R * CPL;              // float4 weight registers per thread
    static_assert(UC % 16 == 0, "unit slice must hold whole 16-float groups");
    static_assert(THREADS <= 1024 && THREADS >= 32, "block size");
    static_assert(CPL >= 1 && NW4 * 4 * THREADS == 4 * UC * H, "weights must tile the CTA's rows exactly");
};

// Shuffle reduce-scatter without selects.  Every lane holds NV partial sums in SLOTS; slot j of lane l carries
// the value whose index is j ^ key(l), where key(l) are the lane bits {topbit, topbit/2, ..} read as a number
// (the weight rows and the sequence order are pre-permuted per lane to make that so).  In the round for lane
// bit `bit` a lane keeps the lower half of its slots and receives the partner's upper half, which carries the
// same value indices; after log2(NV) rounds slot 0 of lane l holds the total of value key(l).
template <int NV>
__device__ __forceinline__ float reduce_scatter(float (&v)[NV], int topbit) {
    int bit = topbit;
#pragma unroll
    for (int half = NV / 2; half >= 1; half >>= 1) {
#pragma unroll
        for (int i = 0; i < half; ++i) v[i] += __shfl_xor_sync(0xffffffffu, v[i + half], bit);
        bit >>= 1;
    }
    return v[0];
}

// Blackwell packed fp32: one FFMA2 issue slot does two FMAs (SASS `FFMA2 Rd, Ra.F32x2, Rb.F32, Rc.F32x2` -- the
// scalar operand is broadcast).  The recurrence is issue-slot bound (ncu: 50 % of the instruction stream was
// non-FMA work), so halving the FMA issue count is a direct win.  A pair = the same k of TWO gate rows.
typedef unsigned long long f32x2;
__device__ __forceinline__ f32x2 pack2(float lo, float hi) {
    f32x2 r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ void unpack2(f32x2 v, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ f32x2 ffma2(f32x2 a, float b, f32x2 c) {
    f32x2 d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(pack2(b, b)), "l"(c));
    return d;
}

// Dynamic shared memory carve-up (floats unless noted):
//   hbuf   [2][C][NB][UC]  h_t of every sequence of the tile, double buffered by step parity; the slice
//                          written by source CTA `src` is contiguous so it travels as ONE bulk copy
//   hstage [2][NB][UC]     this CTA's new slice, source of the bulk copies
//   cbuf   [2][UC][NB]     cell state of the owned units, double buffered by step parity (no read/write hazard
//                          inside a step, so the sequence-group loop has no warp barrier and can be unrolled)
//   lens   [NB] int, goff [NB] uint32 (element offset of sequence b in gin), yoff [NB] uint32 (same for y)
//   bars   [2] uint64      mbarriers, one per hbuf parity
template <int H, int C>
__host__ __device__ inline size_t rec_smem_bytes(int NB) {
    return sizeof(float) * ((size_t)2 * NB * H + (size_t)2 * NB * RecCfg<H, C>::UC + (size_t)2 * NB * RecCfg<H, C>::UC) +
           sizeof(int) * 4 * NB + 2 * sizeof(unsigned long long) + 16;
}

// sigmoid(x) or tanh(x) from ONE exponential, branch free, on the step's serial chain twice (gates, then tanh(c)):
//   sigma(s) = 1 / (1 + 2^(-s log2 e)),  tanh(x) = 2 sigma(2x) - 1
// MUFU ex2 + MUFU reciprocal + one Newton step (7 dependent instructions).  ex2.approx is 2 ulp on the exponential, i.e. <= 6e-8
// absolute on the result -- the same form the tensor-core recurrence uses, inside the parity budget (tests hold the kernels to the
// oracle at 1e-4 on the head outputs, measured ~1e-7).  The first version used expf and the correctly rounded __frcp_rn: two ~10-deep
// dependent sequences per activation, ~200 of the latency path's 1500 clk per step.  The clamp keeps 1 + e finite (rcp(inf) = 0
// would feed inf * 0 into the Newton step).
__device__ __forceinline__ float sigmoid_or_tanh(float x, bool is_tanh) {
#ifdef MP_EXACT_ACT
    const float s = fminf(fmaxf(is_tanh ? 2.0f * x : x, -30.0f), 30.0f);
    const float e = expf(-s);
    const float r = __frcp_rn(1.0f + e);
    return is_tanh ? (1.0f - e) * r : r;
#else
    const float t = fminf(x * (is_tanh ? -2.8853900817779268f : -1.4426950408889634f), 126.0f);
    float e;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(t));
    const float d = 1.0f + e;
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(d));
    r = fmaf(fmaf(-d, r, 1.0f), r, r);
    return is_tanh ? fmaf(2.0f, r, -1.0f) : r;
#endif
}

// gate nonlinearity + i/f/g/o gather + cell update for one (unit, sequence); every lane of the 4x4
// (gate, sequence) group of a unit returns the h_new / c_new of ITS sequence.
__device__ __forceinline__ void lstm_cell(float pre, int gate, int lane, float c_old, float& c_new, float& h_new) {
    const float act = sigmoid_or_tanh(pre, gate == 2);
    const int base = lane & ~0xC;
    const float iv = __shfl_sync(0xffffffffu, act, base);
    const float fv = __shfl_sync(0xffffffffu, act, base | 4);
    const float gv = __shfl_sync(0xffffffffu, act, base | 8);
    const float ov = __shfl_sync(0xffffffffu, act, base | 12);
    c_new = fmaf(fv, c_old, iv * gv);
    h_new = ov * sigmoid_or_tanh(c_new, true);
}

template <int H, int C, int BG>
__global__ void __launch_bounds__(RecCfg<H, C>::THREADS, 1) lstm_rec_kernel(const RecParams p) {
    using Cfg = RecCfg<H, C>;
    constexpr int UC = Cfg::UC, THREADS = Cfg::THREADS, KQ = Cfg::KQ, R = Cfg::R, CPL = Cfg::CPL, NW4 = Cfg::NW4;
    static_assert(BG == 1 || BG == 4, "batch group");

    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int NB = (BG == 1) ? 1 : p.NB;      // latency path: one sequence per cluster
    const int SRC = NB * UC;                  // floats in one source CTA's slice of hbuf
    float* hbuf = reinterpret_cast<float*>(smem_raw);
    float* hstage = hbuf + (size_t)2 * NB * H;
    float* cbuf = hstage + (size_t)2 * NB * UC;          // [2][UC][NB]
    int* lens = reinterpret_cast<int*>(cbuf + (size_t)2 * NB * UC);
    uint32_t* goff = reinterpret_cast<uint32_t*>(lens + NB);
    uint32_t* yoff = goff + NB;
    int* glen = reinterpret_cast<int*>(yoff + NB);       // [NB/4] longest sequence of each group of 4
    unsigned long long* bars =
        reinterpret_cast<unsigned long long*>((reinterpret_cast<uintptr_t>(glen + NB) + 15) & ~uintptr_t(15));

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int rank = (C > 1) ? (int)cluster_ctarank() : 0;
    const int tile = blockIdx.x / C, dir = blockIdx.y;
    const int b_begin = tile * NB;
    const int nb = min(NB, p.B - b_begin);
    // after the reduce-scatter lane = unit_in_warp*16 + gate*4 + q owns (unit, gate, sequence g0+q)
    const int u_local = warp * 2 + (lane >> 4);
    const int gate = (lane >> 2) & 3, qn = lane & 3;
    const int unit = rank * UC + u_local;           // hidden unit in [0, H)
    const int G4 = p.dirs * 4 * H, Y2 = p.dirs * H;
    const int gcol0 = dir * 4 * H;                   // this direction's block of gin columns
    const int gcol = gcol0 + unit * 4 + gate;        // gin columns are (unit, gate)-ordered (api.cu permutes W_ih)
    const int gcol_l = unit * 4 + gate;
    const int kl = lane % KQ;                        // this lane's slot in the K split
    // (weight slot rs of this lane holds gate row rs ^ ((lane >> 2) & (R - 1)); the permutation lives in pack_whh_kernel)
    const int qkey = lane & 3;                       // sequence slot qs of this lane holds sequence g0 + (qs ^ qkey)

    // ---- resident weights: NW4 float4 per thread (slot j: row slot j / CPL, k-chunk kl + (j % CPL)*KQ); the row
    // slots are permuted per lane (row = slot ^ rkey, see pack_whh_kernel) so the reduce-scatter needs no selects
    // Register layout: w2[(rp*CPL + c)*4 + e] = pair (row slot 2rp, row slot 2rp+1) at k = chunk c, element e.
    f32x2 w2[NW4 * 2];
    {
        const float4* wp = p.wpack + ((size_t)(dir * C + rank) * NW4) * THREADS + tid;
#pragma unroll
        for (int j = 0; j < NW4; ++j) {
            const float4 v = __ldg(wp + (size_t)j * THREADS);
            w2[2 * j] = pack2(v.x, v.y);
            w2[2 * j + 1] = pack2(v.z, v.w);
        }
    }
    // float offsets of this lane's k-chunks inside one parity buffer of hbuf ([src][b][u], b = 0)
    int hoff[CPL];
#pragma unroll
    for (int c = 0; c < CPL; ++c) {
        const int k = (c * KQ + kl) * 4;
        hoff[c] = (k / UC) * SRC + (k % UC);
    }

    // ---- tile state ------------------------------------------------------------------------
    for (int i = tid; i < NB; i += THREADS) {
        lens[i] = (i < nb) ? (p.lengths ? min(max(p.lengths[b_begin + i], 0), p.T) : p.T) : 0;
        goff[i] = (uint32_t)(b_begin + min(i, nb - 1)) * (uint32_t)p.T * (uint32_t)G4 + (uint32_t)gcol0;
        yoff[i] = (uint32_t)(b_begin + min(i, nb - 1)) * (uint32_t)p.T * (uint32_t)Y2 + (uint32_t)(dir * H);
    }
    for (int i = tid; i < NB * H; i += THREADS) {      // parity-0 buffer <- h0 (or zeros)
        const int src = i / SRC, rem = i - src * SRC;
        const int b = rem / UC, u = rem - b * UC;
        hbuf[i] = (p.h0 && b < nb) ? p.h0[((size_t)dir * p.B + b_begin + b) * H + src * UC + u] : 0.f;
    }
    for (int i = tid; i < NB * UC; i += THREADS) {
        const int u = i / NB, b = i - u * NB;
        cbuf[i] = (p.c0 && b < nb) ? p.c0[((size_t)dir * p.B + b_begin + b) * H + rank * UC + u] : 0.f;
    }
    for (int i = tid; i < NB / 4; i += THREADS) {
        int m = 0;
        for (int q = 0; q < 4; ++q) {
            const int b = i * 4 + q;
            const int l = (b < nb) ? (p.lengths ? min(max(p.lengths[b_begin + b], 0), p.T) : p.T) : 0;
            m = max(m, l);
        }
        glen[i] = m;
    }
    const uint32_t bar0 = smem_u32(&bars[0]), bar1 = bar0 + 8;
    if (C > 1 && tid == 0) {
        mbar_init(bar0, 1);
        mbar_init(bar1, 1);
        mbar_fence_init_cluster();
    }
    __syncthreads();
    if (C > 1) cluster_sync_all();

    int maxlen = 0;
    for (int i = 0; i < nb; ++i) maxlen = max(maxlen, lens[i]);

    // latency path: this lane feeds CTA (lane & 15) of the cluster; remote views of hbuf / mbarriers
    uint32_t rem_hbuf = 0, rem_bar0 = 0;
    if constexpr (C > 1 && BG == 1) {
        const uint32_t r = (lane & 15) < C ? (lane & 15) : 0;
        rem_hbuf = mapa_u32(smem_u32(hbuf), r);
        rem_bar0 = mapa_u32(bar0, r);
    }
    // gate pre-activations are fetched ahead of use and kept in a register (never on the stack):
    // one step ahead on the latency path, one sequence group ahead on the throughput path
    float gi_next = 0.f;
    const int len0 = lens[0];
    float c_reg = cbuf[u_local];              // latency path: the unit's cell state stays in a register
    int g_hi = ((nb + 3) / 4) * 4;            // throughput path: sequences [0, g_hi) may still be active
    if constexpr (BG == 1) {
        if (len0 > 0) gi_next = __ldg(p.gin + ((size_t)b_begin * p.T + (dir ? len0 - 1 : 0)) * G4 + gcol);
    } else {
        const int l = lens[qn];
        if (l > 0) gi_next = __ldg(p.gin + (goff[qn] + (uint32_t)(dir ? l - 1 : 0) * (uint32_t)G4 + (uint32_t)gcol_l));
    }

    for (int s = 0; s < maxlen; ++s) {
        const int par = s & 1;
        const bool send = (s + 1 < maxlen);
        if constexpr (C > 1) {
            if (s > 0) mbar_wait(par ? bar1 : bar0, ((s - 1) >> 1) & 1);
            if (tid == 0 && send) {
                const int nsend = (BG == 4) ? NB : nb;     // throughput path: whole staged slices travel
                mbar_arrive_expect_tx(par ? bar0 : bar1, (uint32_t)nsend * H * sizeof(float));
            }
        }
        const float* hcur = hbuf + (size_t)par * NB * H;
        float* hnext = hbuf + (size_t)(par ^ 1) * NB * H;
        float* hst = hstage + (size_t)par * NB * UC;

        if constexpr (BG == 1) {
            // ---------------- latency path: the cluster's single sequence ------------------------
            const int t = dir ? len0 - 1 - s : s;
            const float gi = gi_next;
            if (send) gi_next = __ldg(p.gin + ((size_t)b_begin * p.T + (dir ? t - 1 : t + 1)) * G4 + gcol);
            f32x2 acc2[R / 2];
#pragma unroll
            for (int rp = 0; rp < R / 2; ++rp) acc2[rp] = 0ull;
#pragma unroll
            for (int c = 0; c < CPL; ++c) {
                const float4 hv = *reinterpret_cast<const float4*>(hcur + hoff[c]);
#pragma unroll
                for (int rp = 0; rp < R / 2; ++rp) {
                    const int wi = (rp * CPL + c) * 4;
                    acc2[rp] = ffma2(w2[wi + 0], hv.x, acc2[rp]);
                    acc2[rp] = ffma2(w2[wi + 1], hv.y, acc2[rp]);
                    acc2[rp] = ffma2(w2[wi + 2], hv.z, acc2[rp]);
                    acc2[rp] = ffma2(w2[wi + 3], hv.w, acc2[rp]);
                }
            }
            float acc[R];
#pragma unroll
            for (int rp = 0; rp < R / 2; ++rp) unpack2(acc2[rp], acc[2 * rp], acc[2 * rp + 1]);
            // rows -> lane bits 4..2 (unit, gate); then the 4 lanes of a (unit, gate) all get the total
            float tot = reduce_scatter<R>(acc, KQ / 2);
            tot += __shfl_xor_sync(0xffffffffu, tot, 2);
            tot += __shfl_xor_sync(0xffffffffu, tot, 1);
            float c_new, h_new;
            lstm_cell(tot + gi, gate, lane, c_reg, c_new, h_new);
            c_reg = c_new;
            if (send) {
                if constexpr (C == 1) {
                    if ((lane & 15) == 0) hnext[unit] = h_new;
                } else {
                    // the 16 lanes of a unit all hold h_new; lane r of them feeds CTA r
                    if ((lane & 15) < C)
                        st_async_f32(rem_hbuf + (uint32_t)((par ^ 1) * H + unit) * 4u, h_new, rem_bar0 + (par ? 0u : 8u));
                }
            }
            if ((lane & 15) == 0) {
                p.y[((size_t)b_begin * p.T + t) * Y2 + dir * H + unit] = h_new;
                if (s == len0 - 1) {
                    if (p.hn) p.hn[((size_t)dir * p.B + b_begin) * H + unit] = h_new;
                    if (p.cn) p.cn[((size_t)dir * p.B + b_begin) * H + unit] = c_new;
                }
            }
        } else {
            // ---------------- throughput path: groups of 4 sequences; after the reduce lane q owns sequence g0+q.
            // The body is straight-line (stores are predicated, the cell state is double buffered) so that two
            // groups can be unrolled together and the scheduler overlaps the reduce / activation latency chain of
            // one with the FFMA2 stream of the other -- with 4 warps per scheduler the kernel is latency bound
            // otherwise (ncu: issue slots 53 % busy, FMA pipe 50 %).
            while (g_hi > 0 && glen[g_hi / 4 - 1] <= s) g_hi -= 4;     // trailing groups that have finished
            const float* ccur = cbuf + (size_t)par * NB * UC;
            float* cnxt = cbuf + (size_t)(par ^ 1) * NB * UC;
#pragma unroll 2
            for (int g0 = 0; g0 < g_hi; g0 += 4) {
                const int b = g0 + qn;
                const int len = lens[b];
                const bool active = s < len;
                const int t = dir ? len - 1 - s : s;
                const float gi = gi_next;
                {   // prefetch the gate pre-activation of the NEXT group (or of group 0 of the next step)
                    const bool wrap = (g0 + 4 >= g_hi);
                    const int bn = wrap ? qn : b + 4;
                    const int sn = wrap ? s + 1 : s;
                    const int ln = lens[bn];
                    if (sn < ln) gi_next = __ldg(p.gin + (goff[bn] + (uint32_t)(dir ? ln - 1 - sn : sn) * (uint32_t)G4 + (uint32_t)gcol_l));
                }
                f32x2 acc2[R / 2 * 4];
#pragma unroll
                for (int a = 0; a < R / 2 * 4; ++a) acc2[a] = 0ull;
                const float* hg = hcur + g0 * UC;
#pragma unroll
                for (int q = 0; q < 4; ++q) {
#pragma unroll
                    for (int c = 0; c < CPL; ++c) {
                        const float4 hv = *reinterpret_cast<const float4*>(hg + (q ^ qkey) * UC + hoff[c]);
#pragma unroll
                        for (int rp = 0; rp < R / 2; ++rp) {
                            const int wi = (rp * CPL + c) * 4;
                            f32x2 a = acc2[rp * 4 + q];
                            a = ffma2(w2[wi + 0], hv.x, a);
                            a = ffma2(w2[wi + 1], hv.y, a);
                            a = ffma2(w2[wi + 2], hv.z, a);
                            a = ffma2(w2[wi + 3], hv.w, a);
                            acc2[rp * 4 + q] = a;
                        }
                    }
                }
                float acc[R * 4];     // value slot = row_slot*4 + q, row_slot = 2*rp + {0,1}
#pragma unroll
                for (int rp = 0; rp < R / 2; ++rp)
#pragma unroll
                    for (int q = 0; q < 4; ++q) unpack2(acc2[rp * 4 + q], acc[(2 * rp) * 4 + q], acc[(2 * rp + 1) * 4 + q]);
                // (row, sequence) -> lane bits: lane = unit_in_warp*16 + gate*4 + q
                const float tot = reduce_scatter<R * 4>(acc, KQ / 2);
                float c_new, h_new;
                lstm_cell(tot + gi, gate, lane, ccur[u_local * NB + b], c_new, h_new);
                const bool owner = active && gate == 0;
                if (owner) {
                    cnxt[u_local * NB + b] = c_new;
                    p.y[yoff[b] + (uint32_t)t * (uint32_t)Y2 + (uint32_t)unit] = h_new;
                    if (send) {
                        if constexpr (C == 1) hnext[b * UC + u_local] = h_new;
                        else hst[b * UC + u_local] = h_new;
                    }
                }
                if (owner && s == len - 1) {
                    if (p.hn) p.hn[((size_t)dir * p.B + b_begin + b) * H + unit] = h_new;
                    if (p.cn) p.cn[((size_t)dir * p.B + b_begin + b) * H + unit] = c_new;
                }
            }
        }

        if constexpr (C == 1) {
            __syncthreads();
        } else if (BG == 4 && send) {
            // staged slice -> every CTA of the cluster: ONE contiguous NB*UC*4-byte bulk copy per destination
            fence_proxy_async_smem();
            __syncthreads();
            if (tid < C)
                bulk_copy_s2c(mapa_u32(smem_u32(hnext + rank * SRC), tid), smem_u32(hst), SRC * sizeof(float),
                              mapa_u32(par ? bar0 : bar1, tid));
        }
    }

    // frames >= len of the layer output are zero (pad_packed_sequence, rnn.py:31)
    for (int b = 0; b < nb; ++b) {
        const int len = lens[b];
        const int n = (p.T - len) * UC;
        for (int i = tid; i < n; i += THREADS) {
            const int t = len + i / UC, u = i % UC;
            p.y[((size_t)(b_begin + b) * p.T + t) * Y2 + dir * H + rank * UC + u] = 0.f;
        }
    }
    if (C > 1) cluster_sync_all();   // nobody exits while a peer may still address its shared memory
}

// ---------------------------------------------------------------------------------------------
// H = 64, many sequences (the foot-contact head at cfg3 / cfg4): one thread per gate row.
//
// The cluster kernel above splits K across the lanes of a warp, which is right for H = 256 (64 k per lane) and wrong for
// H = 64: 2 k per lane and a 31-shuffle reduction per step -- 1024 threads per CTA busy reducing (measured 2 950 clk per step,
// 0.45 ms per layer on 128 SMs).  Here a CTA of 256 threads serves 4 sequences; thread t owns gate row (unit t/4, gate t%4)
// -- the column order of `gin`, so its pre-activation read is coalesced -- with the row's 64 weights in registers, walks
// h (shared memory, [k][sequence]: one broadcast LDS.128 feeds two packed FFMA2 = 4 sequences) and needs NO reduction.  The 4
// gates of a unit sit in one lane quad: each lane applies its gate's nonlinearity to its 4 sequences, a two-round butterfly
// (4 shuffles) transposes gates x sequences so that lane q owns sequence q of the unit, and the cell update, the h store and
// the y store happen there.  One __syncthreads per step, 2 KB of shared memory, 128 CTAs of 256 threads at B = 256.
// ---------------------------------------------------------------------------------------------
struct RowsParams {
    const float* gin;
    const float* w0;      // W_hh [4H, H] of direction 0 / 1 (torch layout: row = gate*H + unit)
    const float* w1;
    float* y;
    const float* h0;
    const float* c0;
    float* hn;
    float* cn;
    const int32_t* lengths;
    int B, T, dirs;
};

__global__ void __launch_bounds__(256) lstm_rec_h64_rows_kernel(const RowsParams p) {
    constexpr int H = 64, NBT = 4;
    __shared__ __align__(16) float hs[2][H][NBT];      // h_t of the tile's 4 sequences, [parity][k][sequence]
    const int tid = threadIdx.x, lane = tid & 31;
    const int unit = tid >> 2, gate = tid & 3, q = gate;   // after the transpose lane `q` of the quad owns sequence q
    const int dir = blockIdx.y, b_begin = blockIdx.x * NBT;
    const int nb = min(NBT, p.B - b_begin);
    const int G4 = p.dirs * 4 * H, Y2 = p.dirs * H;
    const bool b0 = (lane & 1) != 0, b1 = (lane & 2) != 0;

    // the row's weights, as pairs (w, w) are not needed: FFMA2 broadcasts the scalar operand
    float w[H];
    {
        const float4* src = reinterpret_cast<const float4*>((dir ? p.w1 : p.w0) + (size_t)(gate * H + unit) * H);
#pragma unroll
        for (int i = 0; i < H / 4; ++i) {
            const float4 v = __ldg(src + i);
            w[4 * i] = v.x; w[4 * i + 1] = v.y; w[4 * i + 2] = v.z; w[4 * i + 3] = v.w;
        }
    }
    int len[NBT], maxlen = 0;
#pragma unroll
    for (int i = 0; i < NBT; ++i) {
        len[i] = (i < nb) ? (p.lengths ? min(max(p.lengths[b_begin + i], 0), p.T) : p.T) : 0;
        maxlen = max(maxlen, len[i]);
    }
    int len_q = len[0];                                  // this lane's own sequence (the one it owns after the transpose)
#pragma unroll
    for (int i = 1; i < NBT; ++i) len_q = (q == i) ? len[i] : len_q;
    const int bq = b_begin + min(q, nb - 1);
    // initial state: h of (sequence q, unit) goes to the parity-0 buffer, c stays in a register of lane q
    hs[0][unit][q] = (p.h0 && q < nb) ? p.h0[((size_t)dir * p.B + bq) * H + unit] : 0.f;
    hs[1][unit][q] = 0.f;
    float c_reg = (p.c0 && q < nb) ? p.c0[((size_t)dir * p.B + bq) * H + unit] : 0.f;
    // gate pre-activations one step ahead: column (unit, gate) = tid of every sequence of the tile
    float gi_next[NBT];
#pragma unroll
    for (int i = 0; i < NBT; ++i) {
        const int b = b_begin + min(i, nb - 1);
        gi_next[i] = len[i] > 0 ? __ldg(p.gin + ((size_t)b * p.T + (dir ? len[i] - 1 : 0)) * G4 + dir * 4 * H + tid) : 0.f;
    }
    __syncthreads();

    for (int s = 0; s < maxlen; ++s) {
        const int par = s & 1;
        float gi[NBT];
#pragma unroll
        for (int i = 0; i < NBT; ++i) {
            gi[i] = gi_next[i];
            if (s + 1 < len[i]) {
                const int b = b_begin + i;
                gi_next[i] = __ldg(p.gin + ((size_t)b * p.T + (dir ? len[i] - 2 - s : s + 1)) * G4 + dir * 4 * H + tid);
            }
        }
        // W_hh row . h for the 4 sequences: two independent packed chains per half of K
        f32x2 a01 = 0ull, a23 = 0ull, b01 = 0ull, b23 = 0ull;
#pragma unroll
        for (int k = 0; k < H; k += 2) {
            const float4 h0v = *reinterpret_cast<const float4*>(&hs[par][k][0]);
            const float4 h1v = *reinterpret_cast<const float4*>(&hs[par][k + 1][0]);
            a01 = ffma2(pack2(h0v.x, h0v.y), w[k], a01);
            a23 = ffma2(pack2(h0v.z, h0v.w), w[k], a23);
            b01 = ffma2(pack2(h1v.x, h1v.y), w[k + 1], b01);
            b23 = ffma2(pack2(h1v.z, h1v.w), w[k + 1], b23);
        }
        float v[NBT], t0, t1;
        unpack2(a01, v[0], v[1]);
        unpack2(a23, v[2], v[3]);
        unpack2(b01, t0, t1);
        v[0] += t0; v[1] += t1;
        unpack2(b23, t0, t1);
        v[2] += t0; v[3] += t1;
        // this lane's gate for the 4 sequences
#pragma unroll
        for (int i = 0; i < NBT; ++i) v[i] = sigmoid_or_tanh(v[i] + gi[i], gate == 2);
        // gates x sequences transpose inside the lane quad (lane = gate -> lane = sequence)
        const float r0 = __shfl_xor_sync(0xffffffffu, b0 ? v[0] : v[1], 1);
        const float r1 = __shfl_xor_sync(0xffffffffu, b0 ? v[2] : v[3], 1);
        const float k0 = b0 ? v[1] : v[0], k1 = b0 ? v[3] : v[2];
        const float e0 = b0 ? r0 : k0, o0 = b0 ? k0 : r0;      // even / odd gate of the pair, sequence (lane & 1)
        const float e1 = b0 ? r1 : k1, o1 = b0 ? k1 : r1;      // ... sequence (lane & 1) + 2
        const float rE = __shfl_xor_sync(0xffffffffu, b1 ? e0 : e1, 2);
        const float rO = __shfl_xor_sync(0xffffffffu, b1 ? o0 : o1, 2);
        const float kE = b1 ? e1 : e0, kO = b1 ? o1 : o0;
        const float iv = b1 ? rE : kE, fv = b1 ? rO : kO, gv = b1 ? kE : rE, ov = b1 ? kO : rO;
        // cell update of (unit, sequence q)
        const bool active = s < len_q;
        const float c_new = fmaf(fv, c_reg, iv * gv);
        const float h_new = ov * sigmoid_or_tanh(c_new, true);
        if (active) {
            c_reg = c_new;
            const int t = dir ? len_q - 1 - s : s;
            hs[par ^ 1][unit][q] = h_new;
            p.y[((size_t)(b_begin + q) * p.T + t) * Y2 + dir * H + unit] = h_new;
            if (s == len_q - 1) {
                if (p.hn) p.hn[((size_t)dir * p.B + b_begin + q) * H + unit] = h_new;
                if (p.cn) p.cn[((size_t)dir * p.B + b_begin + q) * H + unit] = c_new;
            }
        }
        __syncthreads();
    }
    // frames >= len of the layer output are zero (pad_packed_sequence, rnn.py:31)
    for (int i = 0; i < nb; ++i) {
        const int n = (p.T - len[i]) * H;
        for (int j = tid; j < n; j += 256) {
            const int t = len[i] + j / H, u = j % H;
            p.y[((size_t)(b_begin + i) * p.T + t) * Y2 + dir * H + u] = 0.f;
        }
    }
}

// ---------------------------------------------------------------------------------------------
// Debug kernel (MP_REC_IMPL=simple): one CTA per (sequence, direction), W_hh^T streamed from L2
// every step.  Same arithmetic, no clusters / mbarriers -- used to bisect failures of the kernel
// above on hardware.  Never selected by default.
// ---------------------------------------------------------------------------------------------
template <int H>
__global__ void __launch_bounds__(4 * H > 1024 ? 1024 : 4 * H) lstm_rec_simple_kernel(
    const float* __restrict__ gin, const float* __restrict__ wT, float* __restrict__ y,
    const float* __restrict__ h0, const float* __restrict__ c0, float* __restrict__ hn,
    float* __restrict__ cn, const int32_t* __restrict__ lengths, int B, int T, int dirs) {
    __shared__ float h[H];
    __shared__ float g[4 * H];
    const int b = blockIdx.x, dir = blockIdx.y, tid = threadIdx.x, nthr = blockDim.x;
    const int len = lengths ? min(max(lengths[b], 0), T) : T;
    const float* wt = wT + (size_t)dir * H * 4 * H;
    float c = 0.f;
    if (tid < H) {
        h[tid] = h0 ? h0[((size_t)dir * B + b) * H + tid] : 0.f;
        c = c0 ? c0[((size_t)dir * B + b) * H + tid] : 0.f;
    }
    __syncthreads();
    for (int s = 0; s < len; ++s) {
        const int t = dir ? len - 1 - s : s;
        for (int r = tid; r < 4 * H; r += nthr) {
            float acc = 0.f;
            for (int k = 0; k < H; ++k) acc = fmaf(wt[(size_t)k * 4 * H + r], h[k], acc);
            g[r] = acc + gin[((size_t)b * T + t) * (dirs * 4 * H) + dir * 4 * H + (r % H) * 4 + r / H];   // (unit, gate) columns
        }
        __syncthreads();
        if (tid < H) {
            const float iv = sigmoidf_acc(g[tid]), fv = sigmoidf_acc(g[H + tid]);
            const float gv = tanhf(g[2 * H + tid]), ov = sigmoidf_acc(g[3 * H + tid]);
            c = fmaf(fv, c, iv * gv);
            const float hv = ov * tanhf(c);
            h[tid] = hv;
            y[((size_t)b * T + t) * (dirs * H) + dir * H + tid] = hv;
            if (s == len - 1) {
                if (hn) hn[((size_t)dir * B + b) * H + tid] = hv;
                if (cn) cn[((size_t)dir * B + b) * H + tid] = c;
            }
        }
        __syncthreads();
    }
    if (tid < H)
        for (int t = len; t < T; ++t) y[((size_t)b * T + t) * (dirs * H) + dir * H + tid] = 0.f;
}

// W_hh [4H, H] (torch) -> wpack[dir][rank][j][tid] float4 (j = row_of_lane * CPL + chunk) + wT[dir][k][row]
template <int H, int C>
__global__ void pack_whh_kernel(const float* __restrict__ w0, const float* __restrict__ w1,
                                float4* __restrict__ wpack, float* __restrict__ wT, int dirs) {
    using Cfg = RecCfg<H, C>;
    const size_t total = (size_t)dirs * C * Cfg::NW4 * Cfg::THREADS;
    for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
         idx += (size_t)gridDim.x * blockDim.x) {
        const int tid = idx % Cfg::THREADS;
        const int j = (idx / Cfg::THREADS) % Cfg::NW4;
        const int rank = (idx / ((size_t)Cfg::THREADS * Cfg::NW4)) % C;
        const int dir = idx / ((size_t)Cfg::THREADS * Cfg::NW4 * C);
        const int lane = tid & 31, warp = tid >> 5;
        // float4 j = ((rp*CPL + c)*2 + hf) holds (W[r0][k], W[r1][k], W[r0][k+1], W[r1][k+1]), k = chunk c + 2*hf, for the
        // row-slot pair (2rp, 2rp+1).  Row slot rs of lane l is gate row rs ^ key(l), key = lane bits 4..2 (KQ = 32) or
        // 3..2 (KQ = 16): the shuffle reduce-scatter of lstm_rec_kernel then needs no per-lane selects.
        const int hf = j & 1, c = (j >> 1) % Cfg::CPL, rp = (j >> 1) / Cfg::CPL;
        const int key = (lane >> 2) & (Cfg::R - 1);
        const int k = (c * Cfg::KQ + lane % Cfg::KQ) * 4 + 2 * hf;
        const float* W = dir ? w1 : w0;
        float v[2][2];
        for (int e = 0; e < 2; ++e) {
            const int r = (2 * rp + e) ^ key;
            // row r -> (unit_in_warp, gate): with one row group (KQ = 32) r = u*4 + g; with two (KQ = 16) the row group is
            // the unit and r is the gate
            const int u_in_warp = (Cfg::RG == 1) ? (r >> 2) : (lane / Cfg::KQ);
            const int gate = r & 3;
            const int unit = rank * Cfg::UC + warp * 2 + u_in_warp;
            const float* src = W + (size_t)(gate * H + unit) * H + k;
            v[e][0] = src[0];
            v[e][1] = src[1];
        }
        wpack[idx] = make_float4(v[0][0], v[1][0], v[0][1], v[1][1]);
    }
    const size_t tt = (size_t)dirs * 4 * H * H;
    for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < tt;
         idx += (size_t)gridDim.x * blockDim.x) {
        const int r = idx % (4 * H);
        const int k = (idx / (4 * H)) % H;
        const int dir = idx / ((size_t)4 * H * H);
        const float* W = dir ? w1 : w0;
        wT[idx] = W[(size_t)r * H + k];
    }
}

int env_int(const char* name, int dflt) {
    const char* v = getenv(name);
    return (v && *v) ? atoi(v) : dflt;
}
bool env_is(const char* name, const char* val) {
    const char* v = getenv(name);
    return v && strcmp(v, val) == 0;
}

constexpr size_t kMaxSmem = 227 * 1024;
constexpr int kMaxTile = 64;   // sequences per cluster (H=256: 2*64*1 KiB of h + staging fits 227 KB)

template <int H, int C, int BG>
int configure_kernel() {
    static bool configured = false;   // one device per process (one rank per GPU)
    if (!configured) {
        auto kern = lstm_rec_kernel<H, C, BG>;
        MP_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kMaxSmem));
        if (C > 8) MP_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
        configured = true;
    }
    return MP_OK;
}

template <int H, int C, int BG>
int launch_cluster(const RecLayerArgs& a, int NB, int bulk, cudaStream_t stream) {
    using Cfg = RecCfg<H, C>;
    RecParams p{a.gin, a.wpack, a.y, a.h0, a.c0, a.hn, a.cn, a.lengths, a.B, a.T, a.dirs, NB, bulk};
    const size_t smem = rec_smem_bytes<H, C>(NB);
    auto kern = lstm_rec_kernel<H, C, BG>;
    MP_REQUIRE(smem <= kMaxSmem, "lstm: tile of %d sequences needs %zu B of shared memory", NB, smem);
    MP_TRY((configure_kernel<H, C, BG>()));
    const int n_tiles = (a.B + NB - 1) / NB;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(C * n_tiles, a.dirs, 1);
    cfg.blockDim = dim3(Cfg::THREADS, 1, 1);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = C;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = (C > 1) ? 1 : 0;
    MP_CUDA_TRY(cudaLaunchKernelEx(&cfg, kern, p));
    count_launch();
    return MP_OK;
}

// How many clusters of this kernel the device can hold at once (GPC / TPC packing decides; on B200 a
// cluster of 8 one-CTA-per-SM blocks does NOT get 148 / 8 slots).  Queried once per kernel.
template <int H, int C, int BG>
int cluster_slots() {
    static int slots = 0;
    if (slots > 0) return slots;
    using Cfg = RecCfg<H, C>;
    configure_kernel<H, C, BG>();
    int n = 0;
    if (C == 1) {
        int per_sm = 0, sms = 0, dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, lstm_rec_kernel<H, C, BG>, Cfg::THREADS, rec_smem_bytes<H, C>(4));
        n = per_sm * sms;
    } else {
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(C * 64, 1, 1);
        cfg.blockDim = dim3(Cfg::THREADS, 1, 1);
        cfg.dynamicSmemBytes = rec_smem_bytes<H, C>(BG == 1 ? 1 : 32);
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = C;
        attr[0].val.clusterDim.y = 1;
        attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr;
        cfg.numAttrs = 1;
        if (cudaOccupancyMaxActiveClusters(&n, lstm_rec_kernel<H, C, BG>, &cfg) != cudaSuccess) {
            cudaGetLastError();
            n = 0;
        }
    }
    const int forced = env_int("MP_REC_SLOTS", 0);
    slots = forced > 0 ? forced : (n > 0 ? n : (C == 1 ? 148 : 8));
    return slots;
}

template <int H, int C>
int launch_for(const RecLayerArgs& a, cudaStream_t stream) {
    // tile policy: if every (sequence, direction) can have its own cluster, run the latency path; otherwise size
    // the tile so that ONE wave of resident clusters covers the batch (a second, nearly empty wave would double the
    // time), capped by shared memory.
    const int forced = env_int("MP_REC_NB", 0);
    const int bulk = env_is("MP_REC_SEND", "stasync") ? 0 : 1;
    int NB;
    if (forced > 0) {
        NB = forced;
    } else if (a.B * a.dirs <= cluster_slots<H, C, 1>()) {
        NB = 1;
    } else {
        const int per = std::max(1, cluster_slots<H, C, 4>() / a.dirs);
        NB = (a.B + per - 1) / per;
        NB = std::min(kMaxTile, ((NB + 3) / 4) * 4);
    }
    if (NB == 1) return launch_cluster<H, C, 1>(a, 1, 0, stream);
    NB = ((NB + 3) / 4) * 4;
    return launch_cluster<H, C, 4>(a, NB, bulk, stream);
}

}  // namespace

int rec_cluster_size(int H) {
    if (H == 64) return 1;
    return env_int("MP_REC_CLUSTER", 8) == 16 ? 16 : 8;
}

size_t whh_pack_float4s(int H, int dirs) { return (size_t)dirs * 4 * H * H / 4; }

int launch_pack_whh(const float* const* w_hh_dirs, int H, int dirs, float4* wpack, float* wT, cudaStream_t stream) {
    const float* w0 = w_hh_dirs[0];
    const float* w1 = dirs > 1 ? w_hh_dirs[1] : w_hh_dirs[0];
    const int C = rec_cluster_size(H);
    if (H == 256 && C == 8) pack_whh_kernel<256, 8><<<296, 256, 0, stream>>>(w0, w1, wpack, wT, dirs);
    else if (H == 256 && C == 16) pack_whh_kernel<256, 16><<<296, 256, 0, stream>>>(w0, w1, wpack, wT, dirs);
    else if (H == 64) pack_whh_kernel<64, 1><<<296, 256, 0, stream>>>(w0, w1, wpack, wT, dirs);
    else {
        set_error("lstm: hidden size %d not built (64, 256)", H);
        return MP_ERR_UNSUPPORTED;
    }
    MP_CUDA_TRY(cudaGetLastError());
    return MP_OK;
}

int launch_lstm_recurrence(const RecLayerArgs& a, cudaStream_t stream) {
    MP_REQUIRE(a.gin && a.wpack && a.y && a.B > 0 && a.T > 0 && (a.dirs == 1 || a.dirs == 2), "lstm: bad arguments");
    MP_REQUIRE((double)a.B * a.T * a.dirs * 4 * a.H < 4.0e9, "lstm: B*T = %lld frames exceeds the 32-bit gate buffer indexing of one launch",
               (long long)a.B * a.T);
    if (!env_is("MP_REC_IMPL", "simple") && rec_f16_eligible(a) && rec_f16w_eligible(a)) return launch_lstm_recurrence_f16w(a, stream);
    if (!env_is("MP_REC_IMPL", "simple") && rec_f16_eligible(a)) return launch_lstm_recurrence_f16(a, stream);
    if (!env_is("MP_REC_IMPL", "simple") && rec_tc_eligible(a)) return launch_lstm_recurrence_tc(a, stream);
    // algorithmic bytes of the recurrence: W_hh once + the layer's h output; the gate
    // pre-activations it reads are an intermediate of this design, charged as what a fully fused
    // layer would read instead (the layer input, counted with the input-projection GEMM).
    ProfileScope prof(a.H == 256 ? "lstm_rec_h256" : "lstm_rec_h64",
                      4.0 * ((double)a.dirs * 4 * a.H * a.H + (double)a.B * a.T * a.dirs * a.H), stream);
    if (env_is("MP_REC_IMPL", "simple")) {
        dim3 grid(a.B, a.dirs);
        if (a.H == 256)
            lstm_rec_simple_kernel<256><<<grid, 1024, 0, stream>>>(a.gin, a.wT, a.y, a.h0, a.c0, a.hn, a.cn, a.lengths, a.B, a.T, a.dirs);
        else if (a.H == 64)
            lstm_rec_simple_kernel<64><<<grid, 256, 0, stream>>>(a.gin, a.wT, a.y, a.h0, a.c0, a.hn, a.cn, a.lengths, a.B, a.T, a.dirs);
        else {
            set_error("lstm: hidden size %d not built", a.H);
            return MP_ERR_UNSUPPORTED;
        }
        MP_CUDA_TRY(cudaGetLastError());
        count_launch();
        return MP_OK;
    }
    if (a.H == 256) return rec_cluster_size(256) == 16 ? launch_for<256, 16>(a, stream) : launch_for<256, 8>(a, stream);
    if (a.H == 64) {
        // more sequences than the one-CTA-per-sequence latency path has room for: one thread per gate row, 4 sequences per CTA
        if (a.B * a.dirs > cluster_slots<64, 1, 1>() && a.w_raw[0] && (a.dirs == 1 || a.w_raw[1]) && !env_is("MP_REC_H64", "cluster")) {
            RowsParams p{a.gin, a.w_raw[0], a.dirs > 1 ? a.w_raw[1] : a.w_raw[0], a.y, a.h0, a.c0, a.hn, a.cn, a.lengths, a.B, a.T, a.dirs};
            lstm_rec_h64_rows_kernel<<<dim3((a.B + 3) / 4, a.dirs), 256, 0, stream>>>(p);
            MP_CUDA_TRY(cudaGetLastError());
            count_launch();
            return MP_OK;
        }
        return launch_for<64, 1>(a, stream);
    }
    set_error("lstm: hidden size %d not built (64, 256)", a.H);
    return MP_ERR_UNSUPPORTED;
}

}  // namespace mp
