// crank-b200: generic channels-last dilated conv / dgrad on tcgen05 (K-major TF32 / 3xTF32 UMMA).
//
// Tensor-core version of k_conv (crk_conv.cuh), same ConvParams contract (prologue / epilogue
// options), used for: dgrad of the gated conv (X = dg, K = 128, N = 64, k taps), the plain
// Conv1d+LeakyReLU stacks (speaker classifier C, SPKRADV; crank/bin/train.py:78-89,
// crank/net/module/spkradv.py:49-60) forward and dgrad, and the 1x1 convs of the WaveNet stacks.
//   CTA = 128 frames of one utterance;  acc[128 x Npad] (TMEM) += X[j*dil + rows][seg] . W_{j,seg}^T
// over (tap j, 64-channel K segment) steps; the weight blobs stream through a 2-slot smem ring while
// the previous step's MMAs run; the A tile (with halo) is staged once, chunk-major (crk_tc.cuh).
#pragma once
#include "crk_common.cuh"
#include "crk_conv.cuh"
#include "crk_resblock_tc.cuh"
#include "crk_tc.cuh"

namespace crk {

struct ConvTcParams {
    ConvParams p;
    const float* Wtc;     // per tap: hi [Kpad/4][chunk_rows(Npad)][4] | lo [same]
    int Kpad, Npad;       // K rounded up to 8, UMMA N (multiple of 16, <= 128)
    int dbg;
    int vec_epi;          // epilogue may use 128-bit accesses (Cout, row strides and pointers all 16 B aligned)
    int opt_stage;        // (unused)
    int kshift, nshift;   // FAST mode: log2(Kpad / 4), log2(Cout / 4)  (both powers of two there)
    // gate-backward mode (gate != 0), the tensor-core version of k_resblock_bwd_gate:
    //   A[row][4q+r] = r<2 ? sqrt(.5)*dH[row][2q+r] : dS[row][2q+r-2]  (also written to GOS),  acc = A . Wos^T = dz
    //   epilogue: dg = gate'(dz; ta, sb) -> DG (interleaved gate order),  z = ta*sb -> Z
    int gate;
    const float* g_dH; const float* g_dS; const float* g_TaSb;
    float* g_DG; float* g_GOS; float* g_Z;
};

// A-tile staging of the gate-backward mode: 128 rows x 128 interleaved [sqrt(.5)*dH | dS] columns
// KEEP: the staged values stay in `keep` (one round: U * 256 == 128 * 32) and the GOS global store is issued
// later by the caller, overlapped with the MMAs, instead of competing with the dH / dS loads here
// pshift / q0: the pairs [q0, q0 + (1 << pshift)) are staged (a K phase of the 2-CTA/SM variant: 16 pairs; all: 32)
template <bool SPLIT, int U, bool KEEP>
__device__ __forceinline__ void tc_stage_gos(float* hi, float* lo, int cs_floats, const ConvTcParams& q, int b, int t0,
                                             float4 (&keep)[KEEP ? U : 1], int pshift = 5, int q0 = 0) {
    const ConvParams& p = q.p;
    const int total = CRK_TC_TM << pshift;                 // (row, pair q): one float4 of A each
    for (int base = threadIdx.x; base < total; base += blockDim.x * U) {
        float2 h[U], sg[U];
        int rr[U], qq[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int idx = base + u * blockDim.x;
            rr[u] = idx >> pshift; qq[u] = q0 + (idx & ((1 << pshift) - 1));
            h[u] = make_float2(0.f, 0.f); sg[u] = make_float2(0.f, 0.f);
            const int t = t0 + rr[u];
            if (idx < total && t < p.T) {
                const size_t row = (size_t)b * p.T + t;
                if (q.g_dH) h[u] = __ldg(reinterpret_cast<const float2*>(q.g_dH + row * 64) + qq[u]);
                sg[u] = __ldg(reinterpret_cast<const float2*>(q.g_dS + row * 64) + qq[u]);
            }
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int idx = base + u * blockDim.x;
            if (idx >= total) continue;
            const float4 v = make_float4(h[u].x * CRK_SQRT_HALF, h[u].y * CRK_SQRT_HALF, sg[u].x, sg[u].y);
            const int t = t0 + rr[u];
            if (KEEP) keep[KEEP ? u : 0] = v;
            else if (t < p.T) reinterpret_cast<float4*>(q.g_GOS + ((size_t)b * p.T + t) * 128)[qq[u]] = v;
            const int off = (qq[u] - q0) * cs_floats + rr[u] * 4;  // chunk = pair index (4 interleaved columns)
            if (SPLIT) {
                float4 hh, ll;
                tc::split_tf32(v.x, hh.x, ll.x); tc::split_tf32(v.y, hh.y, ll.y);
                tc::split_tf32(v.z, hh.z, ll.z); tc::split_tf32(v.w, hh.w, ll.w);
                *reinterpret_cast<float4*>(hi + off) = hh;
                *reinterpret_cast<float4*>(lo + off) = ll;
            } else {
                *reinterpret_cast<float4*>(hi + off) = v;
            }
        }
    }
}

// VECONLY: the host has checked that 128-bit loads are legal (the scalar fallback is compiled out: these
// kernels execute their straight-line code once per CTA, so every unrolled path that is not taken still
// costs instruction-cache footprint)
// c4shift >= 0: Kpad / 4 is a power of two (index split by shift: the runtime division is ~30 instructions per
// element in code that runs once per CTA)
// 128-bit version for the overlapped gate mode: one item = (row, 4 consecutive z channels) = TWO gate pairs, so dH
// and dS are read with 16 B loads (8 + 8 per thread instead of 16 + 16 of 8 B); the staged values stay in
// `keep` (the caller stores them to GOS under the MMAs).  Exactly one round: 8 * 256 == 128 rows * 16.
template <bool SPLIT>
__device__ __forceinline__ void tc_stage_gos4(float* hi, float* lo, int cs_floats, const ConvTcParams& q, int b, int t0,
                                              float4 (&keep)[16]) {
    const ConvParams& p = q.p;
    float4 h[8], sg[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) {
        const int idx = threadIdx.x + u * 256, rr = idx >> 4, pp = idx & 15;
        h[u] = make_float4(0.f, 0.f, 0.f, 0.f); sg[u] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (t0 + rr < p.T) {
            const size_t row = (size_t)b * p.T + t0 + rr;
            if (q.g_dH) h[u] = __ldg(reinterpret_cast<const float4*>(q.g_dH + row * 64) + pp);
            sg[u] = __ldg(reinterpret_cast<const float4*>(q.g_dS + row * 64) + pp);
        }
    }
#pragma unroll
    for (int u = 0; u < 8; ++u) {
        const int idx = threadIdx.x + u * 256, rr = idx >> 4, pp = idx & 15;
        const float4 v0 = make_float4(h[u].x * CRK_SQRT_HALF, h[u].y * CRK_SQRT_HALF, sg[u].x, sg[u].y);
        const float4 v1 = make_float4(h[u].z * CRK_SQRT_HALF, h[u].w * CRK_SQRT_HALF, sg[u].z, sg[u].w);
        keep[2 * u] = v0; keep[2 * u + 1] = v1;
        const int off = (2 * pp) * cs_floats + rr * 4;          // chunk = pair index
        if (SPLIT) {
            float4 hh, ll;
            tc::split_tf32(v0.x, hh.x, ll.x); tc::split_tf32(v0.y, hh.y, ll.y);
            tc::split_tf32(v0.z, hh.z, ll.z); tc::split_tf32(v0.w, hh.w, ll.w);
            *reinterpret_cast<float4*>(hi + off) = hh;
            *reinterpret_cast<float4*>(lo + off) = ll;
            tc::split_tf32(v1.x, hh.x, ll.x); tc::split_tf32(v1.y, hh.y, ll.y);
            tc::split_tf32(v1.z, hh.z, ll.z); tc::split_tf32(v1.w, hh.w, ll.w);
            *reinterpret_cast<float4*>(hi + off + cs_floats) = hh;
            *reinterpret_cast<float4*>(lo + off + cs_floats) = ll;
        } else {
            *reinterpret_cast<float4*>(hi + off) = v0;
            *reinterpret_cast<float4*>(hi + off + cs_floats) = v1;
        }
    }
}

// cbeg / c4n_sel: only the channel chunks [cbeg, cbeg + c4n_sel) are staged, at chunk offset 0 of the tile (a K phase)
template <bool SPLIT, int U, bool HASMUL, bool VECONLY>
__device__ __forceinline__ void tc_stage_act_pro(float* hi, float* lo, int cs_floats, const ConvParams& p, int Kpad,
                                                 int b, int tstart, int rows, int c4shift = -1, int cbeg = 0, int c4n_sel = -1) {
    const int c4n = c4n_sel >= 0 ? c4n_sel : (Kpad >> 2);
    const int total = rows * c4n;
    const bool vec = VECONLY ||
                     (((p.ldx & 3) == 0) && ((p.Cin & 3) == 0) && ((reinterpret_cast<uintptr_t>(p.X) & 15) == 0) &&
                      (p.xmul == nullptr || (((p.ldxmul & 3) == 0) && ((reinterpret_cast<uintptr_t>(p.xmul) & 15) == 0))));
    // U independent float4 loads per thread per round: with one CTA (8 warps) per SM every global round
    // trip costs ~2-3K cycles (measured), so the whole tile should be in flight at once when it fits
    for (int base = threadIdx.x; base < total; base += blockDim.x * U) {
        float4 v[U], m[HASMUL ? U : 1];
        int off[U];
        if (!HASMUL) m[0] = make_float4(1.f, 1.f, 1.f, 1.f);
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int idx = base + u * blockDim.x;
            v[u] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (HASMUL) m[u] = make_float4(1.f, 1.f, 1.f, 1.f);
            off[u] = -1;
            if (idx < total) {
                const int r = (VECONLY && c4shift >= 0) ? (idx >> c4shift) : idx / c4n, cl = idx - r * c4n;
                const int c4 = cbeg + cl;
                off[u] = cl * cs_floats + r * 4;
                const int tt = tstart + r;
                const int c = c4 * 4;
                if (tt >= 0 && tt < p.T && c < p.Cin) {
                    const size_t row = (size_t)b * p.T + tt;
                    if (VECONLY || vec) {
                        v[u] = __ldg(reinterpret_cast<const float4*>(p.X + row * p.ldx) + c4);
                        if (HASMUL && p.xmul) m[u] = __ldg(reinterpret_cast<const float4*>(p.xmul + row * p.ldxmul) + c4);
                    } else if (!VECONLY) {
                        const float* sp = p.X + row * p.ldx + c;
                        v[u].x = __ldg(sp);
                        v[u].y = c + 1 < p.Cin ? __ldg(sp + 1) : 0.f;
                        v[u].z = c + 2 < p.Cin ? __ldg(sp + 2) : 0.f;
                        v[u].w = c + 3 < p.Cin ? __ldg(sp + 3) : 0.f;
                        if (HASMUL && p.xmul) {
                            const float* ms = p.xmul + row * p.ldxmul + c;
                            m[u].x = __ldg(ms);
                            m[u].y = c + 1 < p.Cin ? __ldg(ms + 1) : 0.f;
                            m[u].z = c + 2 < p.Cin ? __ldg(ms + 2) : 0.f;
                            m[u].w = c + 3 < p.Cin ? __ldg(ms + 3) : 0.f;
                        }
                    }
                }
            }
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            if (off[u] < 0) continue;
            const float4 mm = m[HASMUL ? u : 0];
            float4 x;
            x.x = apply_act(v[u].x * p.pro_scale, p.pro_act, p.pro_slope) * mm.x;
            x.y = apply_act(v[u].y * p.pro_scale, p.pro_act, p.pro_slope) * mm.y;
            x.z = apply_act(v[u].z * p.pro_scale, p.pro_act, p.pro_slope) * mm.z;
            x.w = apply_act(v[u].w * p.pro_scale, p.pro_act, p.pro_slope) * mm.w;
            if (SPLIT) {
                float4 h, l;
                tc::split_tf32(x.x, h.x, l.x); tc::split_tf32(x.y, h.y, l.y);
                tc::split_tf32(x.z, h.z, l.z); tc::split_tf32(x.w, h.w, l.w);
                *reinterpret_cast<float4*>(hi + off[u]) = h;
                *reinterpret_cast<float4*>(lo + off[u]) = l;
            } else {
                *reinterpret_cast<float4*>(hi + off[u]) = x;
            }
        }
    }
}

// MODE 0: generic conv / dgrad (every prologue / epilogue option, scalar fallbacks);  MODE 1: gate backward;
// MODE 2: conv / dgrad whose operands allow 128-bit accesses throughout and have no input multiplier (the
// hot instances: every dgrad and 1x1 conv of the WaveNet stacks) -- one staging path, one epilogue path.
#ifndef CRK_CONV_NSLOT2
#define CRK_CONV_NSLOT2 2          // ring depth of the two-CTA/SM variant: 2 slots x 8 chunks.  Measured: 4 x 4 chunks in the same 33 KB is SLOWER (conv family 5.6 -> 6.6 ms per step: the per-step handshake outweighs the deeper prefetch)
#endif
#define CRK_CONV_GENERIC 0
#define CRK_CONV_GATE 1
#define CRK_CONV_FAST 2
// K PHASES (round 2).  Round 1 staged the whole K (up to 128 channels, hi | lo = 140 KB for the dgrad tile) at once, which
// left room for ONE CTA per SM: 256 tiles on 148 SMs ran as two strictly serial waves whose global-memory phases
// (staging, epilogue) hit HBM in lock step while the tensor pipe idled, and vice versa.  Now a CTA stages at most
// PHC = 16 channel chunks (64 channels) at a time and the weight ring holds SEG = 8 chunks per slot: 70 + 33 KB, two
// CTAs per SM (and <= 128 registers), so one CTA's MMAs run under the other's staging / epilogue and every tile of a
// 64 x 500-frame launch is resident in a single wave.  Phase ph: wait until the MMAs of phase ph-1 have read the A tile,
// restage, issue (tap, segment) steps; the weight producer runs NSLOT steps ahead across phase boundaries.
// WIDE: round 1's footprint (whole K staged at once, 16-chunk ring slots, one CTA per SM in the 3xTF32 mode) -- the shorter
// per-CTA latency wins when the launch has no more tiles than SMs (8 utterances per GPU: 32 tiles), where nothing can co-reside.
template <bool SPLIT, int MODE, bool WIDE>
__global__ void __launch_bounds__(256, (SPLIT && WIDE) ? 1 : 2) k_conv_tc(const ConvTcParams q) {
    const ConvParams& p = q.p;
    extern __shared__ float4 crk_smem4[];
    float* smem = reinterpret_cast<float*>(crk_smem4);
    // ring: NSLOT slots of SEG channel chunks (see CRK_CONV_NSLOT2)
    constexpr int NSLOT = (SPLIT && !WIDE) ? CRK_CONV_NSLOT2 : 2;
    constexpr int SEG = (SPLIT && !WIDE) ? (32 / CRK_CONV_NSLOT2) / 2 : 16;     // channel chunks per ring slot
    constexpr int PHC = (SPLIT && !WIDE) ? 16 : 32;    // channel chunks staged per K phase
    __shared__ uint64_t bar_full[NSLOT];
    __shared__ uint64_t bar_free[NSLOT];
    __shared__ uint64_t bar_acc;                   // all MMAs of a phase have completed (one completion per phase)
    __shared__ uint32_t tmem_base_s;
    __shared__ int timeout_s;

    const int tiles_per_utt = (p.T + CRK_TC_TM - 1) / CRK_TC_TM;
    const int b = blockIdx.x / tiles_per_utt;
    const int t0 = (blockIdx.x - b * tiles_per_utt) * CRK_TC_TM;
    const int rowsX = CRK_TC_TM + (p.k - 1) * p.dil;
    const int csx = tc::chunk_rows(rowsX) * 4;                 // floats per A chunk
    const int kch = q.Kpad >> 2;                               // A / B chunks over the whole K
    const int csw = tc::chunk_rows(q.Npad) * 4;                // floats per B chunk
    const int nph = (kch + PHC - 1) / PHC;                     // K phases
    const int phc_max = kch < PHC ? kch : PHC;
    const int whalf_tap = kch * csw;                           // floats of the hi half of one tap blob
    const int seg_max = (kch < SEG ? kch : SEG) * csw;         // floats of one segment half in a slot
    float* Xh = smem;
    float* Xl = Xh + phc_max * csx;
    float* ring = Xl + (SPLIT ? phc_max * csx : 0);
    auto slot_hi = [&](int sl) -> float* { return ring + sl * (SPLIT ? 2 : 1) * seg_max; };
    auto slot_lo = [&](int sl) -> float* { return ring + sl * (SPLIT ? 2 : 1) * seg_max + seg_max; };
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    auto ph_chunks = [&](int ph) -> int { const int r = kch - ph * PHC; return r < PHC ? r : PHC; };
    auto ph_nseg = [&](int ph) -> int { return (ph_chunks(ph) + SEG - 1) / SEG; };
    // steps are enumerated phase-major: (ph, tap j, segment sg);  steps_before(ph) = first step index of phase ph
    auto steps_before = [&](int ph) -> int { int n = 0; for (int i = 0; i < ph; ++i) n += p.k * ph_nseg(i); return n; };
    const int nsteps = steps_before(nph);

    // TMA bulk copy of step st's weight slice into its ring slot
    auto produce = [&](int st) {
        int ph = 0, rem = st;
        while (rem >= p.k * ph_nseg(ph)) { rem -= p.k * ph_nseg(ph); ++ph; }
        const int ns = ph_nseg(ph);
        const int j = rem / ns, sg = rem - j * ns;
        const int ch0 = ph * PHC + sg * SEG;
        const int chend = ph * PHC + ph_chunks(ph);
        const int nch = (chend - ch0) < SEG ? (chend - ch0) : SEG;
        const float* blob = q.Wtc + (size_t)j * 2 * whalf_tap + (size_t)ch0 * csw;
        const int sl = st & (NSLOT - 1);
        tc_bulk_blob<SPLIT>(slot_hi(sl), slot_lo(sl), blob, nch * csw, whalf_tap, &bar_full[sl]);
    };

    if (threadIdx.x == 0) {
        for (int i = 0; i < NSLOT; ++i) { tc::mbar_init(&bar_full[i], 1); tc::mbar_init(&bar_free[i], 1); }
        tc::mbar_init(&bar_acc, 1);
        tc::fence_mbar_init();
        timeout_s = 0;
    }
    if (warp == 1) tc::tmem_alloc<128>(&tmem_base_s);
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    const uint32_t tmem = tmem_base_s;
    bool ok = true;
    pdl_trigger();                                      // only now: this CTA already owns its TMEM columns (see crk_common.cuh)
    pdl_wait();
    dbg_stamp(q.dbg, 0);
    int produced = 0;                                   // (meaningful in thread 0 only)
    if (threadIdx.x == 0)
        for (; produced < NSLOT && produced < nsteps; ++produced) produce(produced);
    float4 gkeep[1];
    const uint32_t idesc = tc::make_idesc_tf32(128, q.Npad, 0, 0);
    uint32_t acc = 0;
    long long wfull = 0, wfree = 0;
    for (int ph = 0; ph < nph; ++ph) {
        const int pc0 = ph * PHC, pcn = ph_chunks(ph);
        if (ph > 0) {                                   // the MMAs of the previous phase have read the A tile
            ok &= tc::mbar_wait(&bar_acc, (ph - 1) & 1);
            tc::tc_fence_after();
        }
        if constexpr (MODE == CRK_CONV_GATE) tc_stage_gos<SPLIT, 4, false>(Xh, Xl, csx, q, b, t0, gkeep, PHC == 16 ? 4 : 5, pc0);
        else if constexpr (MODE == CRK_CONV_FAST)
            tc_stage_act_pro<SPLIT, 4, false, true>(Xh, Xl, csx, p, q.Kpad, b, t0 - p.padl, rowsX,
                                                    (pcn == 16) ? 4 : ((pcn == 32) ? 5 : ((pcn == 8) ? 3 : -1)), pc0, pcn);
        else tc_stage_act_pro<SPLIT, 4, true, false>(Xh, Xl, csx, p, q.Kpad, b, t0 - p.padl, rowsX, -1, pc0, pcn);
        tc::fence_proxy_async_smem();
        __syncthreads();
        if (ph == 0) dbg_stamp(q.dbg, 1);
        const int st0 = steps_before(ph), st1 = st0 + p.k * ph_nseg(ph);
        // one elected lane per role; the other 31 lanes of that warp park at __syncwarp
        if (warp == 0) {
            if (lane == 0) {
                const int upto = (st1 + NSLOT) < nsteps ? (st1 + NSLOT) : nsteps;       // runs NSLOT steps into the next phase
                for (; produced < upto; ++produced) {
                    const long long c0 = q.dbg ? clock64() : 0;
                    ok &= tc::mbar_wait(&bar_free[produced & (NSLOT - 1)], ((produced - NSLOT) / NSLOT) & 1);
                    if (q.dbg) wfree += clock64() - c0;
                    produce(produced);
                }
            }
            __syncwarp();
        } else if (warp == 1) {
            // MMA issuer: the whole warp runs the (uniform) loop, one elected lane issues (see tc_issue_kmajor_w)
            const uint32_t xh_s = tc::smem_u32(Xh), xl_s = tc::smem_u32(Xl);
            const int ns = ph_nseg(ph);
            for (int st = st0; st < st1; ++st) {
                const int rem = st - st0;
                const int j = rem / ns, sg = rem - j * ns;
                const int cl0 = sg * SEG;                               // first chunk of the segment inside the staged tile
                const int nch = (pcn - cl0) < SEG ? (pcn - cl0) : SEG;
                const int sl = st & (NSLOT - 1);
                const long long c0 = q.dbg ? clock64() : 0;
                ok &= tc::mbar_wait(&bar_full[sl], (st / NSLOT) & 1);
                if (q.dbg) wfull += clock64() - c0;
                tc::tc_fence_after();
                tc_issue_kmajor_w<SPLIT>(tmem, xh_s + cl0 * csx * 4, xl_s + cl0 * csx * 4, csx * 4, j * p.dil,
                                         tc::smem_u32(slot_hi(sl)), tc::smem_u32(slot_lo(sl)), csw * 4,
                                         nch * 4, idesc, acc);
                if (tc::elect_one()) tc::umma_commit(&bar_free[sl]);
            }
            if (tc::elect_one()) tc::umma_commit(&bar_acc);
            __syncwarp();
        }
    }
    if (threadIdx.x == 0) dbg_put(q.dbg, 9, wfree);
    if (threadIdx.x == 32) dbg_put(q.dbg, 8, wfull);
    // FAST mode: side inputs of the epilogue (residual gradient, dropout multiplier, activation-derivative
    // source, old output) -- EU float4 each, fetched after the accumulator wait.
    constexpr int EU = (MODE == CRK_CONV_FAST) ? 4 : 1;
    float4 mulv[EU], rv[EU], dv[EU], oldv[EU];
    auto epi_side_load = [&](int e0) {
        const int c4n = p.Cout >> 2;
        const int total = min(CRK_TC_TM, p.T - t0) * c4n;
        const size_t row0 = (size_t)b * p.T + t0;
        const float4 one4 = make_float4(1.f, 1.f, 1.f, 1.f), zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int u = 0; u < EU; ++u) {
            const int e = e0 + u * blockDim.x;
            mulv[u] = one4; rv[u] = zero4; dv[u] = one4; oldv[u] = zero4;
            if (e < total) {
                const int rr = e >> q.nshift, c4 = e - rr * c4n;
                const size_t row = row0 + rr;
                if (p.mul_src) mulv[u] = __ldg(reinterpret_cast<const float4*>(p.mul_src + row * p.ldmul) + c4);
                if (p.R) rv[u] = __ldg(reinterpret_cast<const float4*>(p.R + row * p.ldr) + c4);
                if (p.dact_src) dv[u] = __ldg(reinterpret_cast<const float4*>(p.dact_src + row * p.lddact) + c4);
                if (p.accumulate) oldv[u] = *(reinterpret_cast<const float4*>(p.Y + row * p.ldy) + c4);
            }
        }
    };
    ok &= tc::mbar_wait(&bar_acc, (nph - 1) & 1);
    tc::tc_fence_after();
    if (!ok) timeout_s = 1;
    dbg_stamp(q.dbg, 2);

    // ---- epilogue (same option order as k_conv) ----
    // TMEM -> registers is thread-per-row; the accumulators are transposed through a padded smem tile
    // (all pipeline buffers are free: every MMA has completed) so that the global pass is row-coalesced.
    const int sst = q.Npad | 1;                                 // odd row stride
    float* S = smem;
    {
        const int r = (warp & 3) * 32 + lane;
        const uint32_t tlane = tmem + ((uint32_t)((warp & 3) * 32) << 16);
        const int nblk = (q.Npad + 31) >> 5;
        for (int blk = warp >> 2; blk < nblk; blk += 2) {
            float v[32];
            tc::tmem_ld32(tlane + blk * 32, v);
#pragma unroll
            for (int i = 0; i < 32; ++i)
                if (blk * 32 + i < q.Npad) S[r * sst + blk * 32 + i] = v[i];
        }
    }
    __syncthreads();
    dbg_stamp(q.dbg, 3);
    if constexpr (MODE == CRK_CONV_GATE) {
        // lanes over gate pairs (2 z channels each): TaSb read and DG write are one float4 per lane
        const int nlive = min(CRK_TC_TM, p.T - t0);
        const size_t row0 = (size_t)b * p.T + t0;
        const int total = nlive * 32;
        constexpr int U = 4;
        for (int e0 = threadIdx.x; e0 < total; e0 += blockDim.x * U) {
            float4 ts[U];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int e = e0 + u * blockDim.x;
                if (e < total) ts[u] = __ldg(reinterpret_cast<const float4*>(q.g_TaSb + (row0 + (e >> 5)) * 128) + (e & 31));
            }
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int e = e0 + u * blockDim.x;
                if (e >= total) continue;
                const int rr = e >> 5, qi = e & 31;
                const float dz0 = S[rr * sst + 2 * qi], dz1 = S[rr * sst + 2 * qi + 1];
                float4 dg;
                dg.x = (dz0 * ts[u].z) * (1.f - ts[u].x * ts[u].x);
                dg.y = (dz1 * ts[u].w) * (1.f - ts[u].y * ts[u].y);
                dg.z = (dz0 * ts[u].x) * ((1.f - ts[u].z) * ts[u].z);
                dg.w = (dz1 * ts[u].y) * ((1.f - ts[u].w) * ts[u].w);
                reinterpret_cast<float4*>(q.g_DG + (row0 + rr) * 128)[qi] = dg;
                reinterpret_cast<float2*>(q.g_Z + (row0 + rr) * 64)[qi] = make_float2(ts[u].x * ts[u].z, ts[u].y * ts[u].w);
            }
        }
    } else if constexpr (MODE == CRK_CONV_FAST) {
        // 128-bit epilogue: one float4 (4 output channels of a frame) per thread per slot, U slots in
        // flight, so the side inputs of a whole 128 x 64 tile (dgrad: residual grad + dropout
        // multiplier) need 1-2 global round trips instead of 8
        const int nlive = min(CRK_TC_TM, p.T - t0);
        const size_t row0 = (size_t)b * p.T + t0;
        const int c4n = p.Cout >> 2;
        const int total = nlive * c4n;
        constexpr int U = 4;                                // (two CTAs per SM: 128 registers per thread)
        for (int e0 = threadIdx.x; e0 < total; e0 += blockDim.x * U) {
            epi_side_load(e0);
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int e = e0 + u * blockDim.x;
                if (e >= total) continue;
                const int rr = e >> q.nshift, c4 = e - rr * c4n;
                const float* sp = S + rr * sst + 4 * c4;
                float y[4] = {sp[0], sp[1], sp[2], sp[3]};
                const float m4[4] = {mulv[u].x, mulv[u].y, mulv[u].z, mulv[u].w};
                const float r4[4] = {rv[u].x, rv[u].y, rv[u].z, rv[u].w};
                const float d4[4] = {dv[u].x, dv[u].y, dv[u].z, dv[u].w};
                const float o4[4] = {oldv[u].x, oldv[u].y, oldv[u].z, oldv[u].w};
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    float v = y[i];                     // same operation order as the scalar path below
                    if (p.bias) v += __ldg(p.bias + 4 * c4 + i);
                    v = apply_act(v, p.epi_act, p.epi_slope);
                    v *= m4[i];
                    v += p.rscale * r4[i];
                    if (p.dact_src) v *= act_grad(d4[i], p.dact_mode, p.dact_slope);
                    v *= p.out_scale;
                    v += o4[i];
                    y[i] = v;
                }
                reinterpret_cast<float4*>(p.Y + (row0 + rr) * p.ldy)[c4] = make_float4(y[0], y[1], y[2], y[3]);
            }
        }
    } else {
        const int nlive = min(CRK_TC_TM, p.T - t0);
        const size_t row0 = (size_t)b * p.T + t0;
        const int total = nlive * p.Cout;
        constexpr int U = 4;      // independent elements per thread per round: their global loads overlap
        for (int e0 = threadIdx.x; e0 < total; e0 += blockDim.x * U) {
            float y[U], mulv[U], rv[U], dv[U], oldv[U];
            size_t rowv[U];
            int cov[U];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int e = e0 + u * blockDim.x;
                cov[u] = -1;
                if (e < total) {
                    const int rr = e / p.Cout, co = e - rr * p.Cout;
                    const size_t row = row0 + rr;
                    cov[u] = co; rowv[u] = row;
                    y[u] = S[rr * sst + co];
                    mulv[u] = p.mul_src ? __ldg(p.mul_src + row * p.ldmul + co) : 1.f;
                    rv[u] = p.R ? __ldg(p.R + row * p.ldr + co) : 0.f;
                    dv[u] = p.dact_src ? __ldg(p.dact_src + row * p.lddact + co) : 1.f;
                    oldv[u] = p.accumulate ? p.Y[row * p.ldy + co] : 0.f;
                }
            }
#pragma unroll
            for (int u = 0; u < U; ++u) {
                if (cov[u] < 0) continue;
                float v = y[u];
                if (p.bias) v += __ldg(p.bias + cov[u]);
                v = apply_act(v, p.epi_act, p.epi_slope);
                v *= mulv[u];
                v += p.rscale * rv[u];
                if (p.dact_src) v *= act_grad(dv[u], p.dact_mode, p.dact_slope);
                v *= p.out_scale;
                v += oldv[u];
                p.Y[rowv[u] * p.ldy + cov[u]] = v;
            }
        }
    }
    tc::tc_fence_before();
    __syncthreads();
    dbg_stamp(q.dbg, 4);
    if (timeout_s && threadIdx.x == 0)
        (MODE == CRK_CONV_GATE ? q.g_DG : p.Y)[((size_t)b * p.T + t0) * (MODE == CRK_CONV_GATE ? 128 : p.ldy)] = __int_as_float(0x7fc00000);
    if (warp == 1) tc::tmem_dealloc<128>(tmem);
}

inline size_t conv_tc_smem(const ConvTcParams& q, bool split, bool wide = false) {
    const int rowsX = CRK_TC_TM + (q.p.k - 1) * q.p.dil;
    const int kch = q.Kpad >> 2;
    const int phc = (split && !wide) ? 16 : 32, segc = (split && !wide) ? (32 / CRK_CONV_NSLOT2) / 2 : 16;     // PHC / SEG of the kernel
    const int nslot = (split && !wide) ? CRK_CONV_NSLOT2 : 2;
    const size_t a = (size_t)(kch < phc ? kch : phc) * tc::chunk_rows(rowsX) * 4;
    const size_t seg = (size_t)(kch < segc ? kch : segc) * tc::chunk_rows(q.Npad) * 4;   // one half of a ring slot
    const size_t pipe = (split ? 2 : 1) * a + (size_t)nslot * (split ? 2 : 1) * seg;
    const size_t stage = (size_t)CRK_TC_TM * (q.Npad | 1);      // epilogue transposition tile
    return (pipe > stage ? pipe : stage) * sizeof(float);
}
inline bool conv_tc_ok(const ConvTcParams& q, bool split) {
    return q.Wtc != nullptr && q.Npad >= 16 && q.Npad <= 128 && (q.Npad % 16) == 0 && q.Kpad >= 8 && q.Kpad <= 128 &&
           (q.p.k - 1) * q.p.dil <= 32 && conv_tc_smem(q, split) <= 220 * 1024;
}
inline int device_sm_count_conv() {
    static int n = 0;
    if (n == 0) {
        int dev = 0;
        if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    }
    return n;
}

template <bool SPLIT, int MODE, bool WIDE>
inline cudaError_t launch_conv_tc_w(const ConvTcParams& q, cudaStream_t s) {
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(k_conv_tc<SPLIT, MODE, WIDE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 225 * 1024);
        if (e != cudaSuccess) return e;
        attr_set = true;
    }
    const int tiles = q.p.B * cdiv(q.p.T, CRK_TC_TM);
    TimedLaunch tl(MODE == CRK_CONV_GATE ? CRK_K_BWD_GATE : CRK_K_CONV, s,
                   MODE == CRK_CONV_GATE ? 2.0 * q.p.B * q.p.T * 128.0 * 64 : 2.0 * q.p.B * q.p.T * q.p.Cin * q.p.Cout * q.p.k);
    ConvTcParams qq = q;
    qq.dbg = dbg_take(MODE == CRK_CONV_GATE ? CRK_K_BWD_GATE : CRK_K_CONV);
    cudaError_t le = launch_pdl(k_conv_tc<SPLIT, MODE, WIDE>, dim3(tiles), dim3(256), conv_tc_smem(q, SPLIT, WIDE), s, qq);
    if (le != cudaSuccess) return le;
    return launch_check();
}
template <bool SPLIT, int MODE>
inline cudaError_t launch_conv_tc_t(const ConvTcParams& q, cudaStream_t s) {
    const int tiles = q.p.B * cdiv(q.p.T, CRK_TC_TM);
    // 3xTF32: the two-CTA/SM footprint pays when tiles outnumber SMs; otherwise the whole-K variant has the shorter latency
    if (SPLIT && (tiles <= device_sm_count_conv() || (opt_disable_mask() & 128)) && conv_tc_smem(q, true, true) <= 220 * 1024)
        return launch_conv_tc_w<SPLIT, MODE, true>(q, s);
    return launch_conv_tc_w<SPLIT, MODE, false>(q, s);
}

// persistent pipelined variant (crk_conv_pt.cuh, included after this header)
template <bool SPLIT> inline cudaError_t launch_conv_pt(const ConvTcParams& q, cudaStream_t s);
inline bool conv_pt_ok(const ConvTcParams& q, bool split);

// precision-aware dispatch: tensor cores when enabled and the shape fits, else the fp32 kernel
inline cudaError_t conv_dispatch(const ConvParams& p, int cpt, const float* wtc, int kpad, int npad, cudaStream_t s) {
    if (p.Cout > 128) {
        // more than 128 output columns: the dgrad of a stack's first conv when its input is wider than 128 channels
        // (n_vq_stacks = 3: 192).  fp32 kernel over 128-column slices of the packed matrix (row stride wide_ld(Cout)).
        const int ldw = wide_ld(p.Cout);
        for (int n0 = 0; n0 < p.Cout; n0 += 128) {
            ConvParams q = p;
            q.W = p.W + n0; q.ldw = ldw; q.Cout = (p.Cout - n0) < 128 ? (p.Cout - n0) : 128;
            if (p.bias) q.bias = p.bias + n0;
            q.Y = p.Y + n0;
            if (p.mul_src) q.mul_src = p.mul_src + n0;
            if (p.R) q.R = p.R + n0;
            if (p.dact_src) q.dact_src = p.dact_src + n0;
            cudaError_t e = launch_conv(q, cpt_for(q.Cout), s);
            if (e != cudaSuccess) return e;
        }
        return cudaSuccess;
    }
    const int mode = precision_mode();
    if (mode != CRK_PREC_FP32 && !(tc_disable_mask() & 2)) {
        ConvTcParams q;
        q.p = p; q.Wtc = wtc; q.Kpad = kpad; q.Npad = npad; q.dbg = 0; q.gate = 0;
        auto al = [](const float* ptr, int ld) { return ptr == nullptr || (((ld & 3) == 0) && ((reinterpret_cast<uintptr_t>(ptr) & 15) == 0)); };
        q.opt_stage = 0;
        q.vec_epi = (p.Cout & 3) == 0 && p.Y != nullptr && al(p.Y, p.ldy) && al(p.mul_src, p.ldmul) &&
                    al(p.R, p.ldr) && al(p.dact_src, p.lddact);
        auto lg2 = [](int v) { int sft = 0; while ((1 << sft) < v) ++sft; return (1 << sft) == v ? sft : -1; };
        q.kshift = lg2(kpad >> 2); q.nshift = lg2(p.Cout >> 2);
        const bool fast = !(opt_disable_mask() & 4) && q.vec_epi && p.xmul == nullptr && (p.Cin & 3) == 0 && al(p.X, p.ldx) && p.X != nullptr &&
                          q.kshift >= 0 && q.nshift >= 0;
        q.g_dH = q.g_dS = q.g_TaSb = nullptr; q.g_DG = q.g_GOS = q.g_Z = nullptr;
        const bool split = mode == CRK_PREC_TF32X3;
        if (fast && (opt_enable_mask() & 2) && conv_pt_ok(q, split))
            return split ? launch_conv_pt<true>(q, s) : launch_conv_pt<false>(q, s);
        if (conv_tc_ok(q, split)) {
            if (fast) return split ? launch_conv_tc_t<true, CRK_CONV_FAST>(q, s) : launch_conv_tc_t<false, CRK_CONV_FAST>(q, s);
            return split ? launch_conv_tc_t<true, CRK_CONV_GENERIC>(q, s) : launch_conv_tc_t<false, CRK_CONV_GENERIC>(q, s);
        }
    }
    return launch_conv(p, cpt, s);
}

// tensor-core gate backward (replaces k_resblock_bwd_gate when a tensor-core mode is on)
inline bool gate_bwd_tc(const ResBwdGateParams& g, const float* wos_tct, cudaStream_t s, cudaError_t* err) {
    const int mode = precision_mode();
    if (mode == CRK_PREC_FP32 || (tc_disable_mask() & 8)) return false;
    ConvTcParams q;
    q.p = conv_params_default();
    q.p.B = g.B; q.p.T = g.T; q.p.Cin = 128; q.p.Cout = 64; q.p.k = 1; q.p.dil = 1; q.p.padl = 0;
    q.Wtc = wos_tct; q.Kpad = 128; q.Npad = 64; q.dbg = 0; q.gate = 1; q.vec_epi = 0; q.opt_stage = 0; q.kshift = q.nshift = -1;
    q.g_dH = g.dH; q.g_dS = g.dS; q.g_TaSb = g.TaSb; q.g_DG = g.DG; q.g_GOS = g.GOS; q.g_Z = g.Z;
    const bool split = mode == CRK_PREC_TF32X3;
    if (conv_tc_smem(q, split) > 220 * 1024) return false;
    *err = split ? launch_conv_tc_t<true, CRK_CONV_GATE>(q, s) : launch_conv_tc_t<false, CRK_CONV_GATE>(q, s);
    return true;
}


}  // namespace crk
