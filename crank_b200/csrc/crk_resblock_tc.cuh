// crank-b200: fused WaveNet residual block FORWARD on the 5th-gen tensor cores (tcgen05 / TMEM).
//
// Same math as k_resblock_fwd (crk_resblock.cuh) -- parallel_wavegan ResidualBlock.forward as used at
// crank/net/module/vqvae2.py:236-273 and crank/bin/train.py:107-115 -- as an implicit GEMM:
//   CTA = 128 consecutive frames of one utterance (UMMA M = 128).
//   GEMM1: for each tap j: acc1[128 x 128] += X[j*dil + (0..127)][0..63] . W_j^T    (+ aux 1x1)
//          A operand = ONE staged (128+halo) x 64 tile in chunk-major layout (crk_tc.cuh); tap j is
//          just a different descriptor start row.  B operand = pre-packed weight blobs streamed
//          through a 2-slot shared-memory ring while the previous tap's MMAs run.
//   epilogue 1 (8 warps, TMEM -> registers): +bias, tanh * sigmoid, save (tanh, sigmoid) for backward,
//          z -> shared memory (chunk-major, aliases the X tile).
//   GEMM2: acc2[128 x 128] = z . [Wout | Wskip]^T ;  epilogue 2: +bias, residual, *sqrt(.5), skip +=.
// Precision: SPLIT=true runs the 3xTF32 error-compensated product (x = hi + lo, products hi*hi +
// hi*lo + lo*hi accumulate in fp32 TMEM): ~fp32 accuracy, needed for the <=1e-4 parity contract;
// SPLIT=false is plain TF32.
#pragma once
#include "crk_common.cuh"
#include "crk_resblock.cuh"
#include "crk_tc.cuh"

namespace crk {

#define CRK_TC_TM 128

struct ResFwdTcParams {
    ResFwdParams p;          // same tensors as the fp32 kernel
    const float* WcTc;       // per tap blob: hi [16][129][4] | lo [16][129][4]
    const float* WaTc;       // aux blob: hi [KaPad/4][129][4] | lo   (KaPad = round_up(Ca, 8))
    const float* WosTc;      // blob hi [16][129][4] | lo
    int KaPad;
    int dbg;
};

__host__ __device__ constexpr int tc_blob_half(int kdim, int nrows) { return (kdim / 4) * tc::chunk_rows(nrows) * 4; }

// stage rows [tstart, tstart+rows) x ncols channels (zero outside [0,T) / beyond ncols) into chunk-major
// hi/lo tiles.  Loads are issued in batches of 4 per thread before any use, so one L2 round trip covers
// the batch (a plain per-iteration load->store loop is latency-serialised: measured 10x slower).
template <bool SPLIT, int U = 4>
__device__ __forceinline__ void tc_stage_act(float* hi, float* lo, int cs_floats, const float* __restrict__ src,
                                             int ld, int ncols, int ncols_pad, int b, int T, int tstart, int rows,
                                             const float* __restrict__ mul, int ldmul) {
    const int c4n = ncols_pad >> 2;
    const int total = rows * c4n;
    const bool vec = ((ld & 3) == 0) && ((ncols & 3) == 0) && ((reinterpret_cast<uintptr_t>(src) & 15) == 0);
    for (int base = threadIdx.x; base < total; base += blockDim.x * U) {
        float4 v[U], m[U];
        int off[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int idx = base + u * blockDim.x;
            v[u] = make_float4(0.f, 0.f, 0.f, 0.f);
            m[u] = make_float4(1.f, 1.f, 1.f, 1.f);
            off[u] = -1;
            if (idx < total) {
                const int r = idx / c4n, c4 = idx - r * c4n;
                off[u] = c4 * cs_floats + r * 4;
                const int tt = tstart + r;
                if (tt >= 0 && tt < T) {
                    const size_t row = (size_t)b * T + tt;
                    if (vec) {
                        v[u] = __ldg(reinterpret_cast<const float4*>(src + row * ld) + c4);
                    } else {
                        const float* sp = src + row * ld + c4 * 4;
                        const int c = c4 * 4;
                        v[u].x = c + 0 < ncols ? __ldg(sp + 0) : 0.f;
                        v[u].y = c + 1 < ncols ? __ldg(sp + 1) : 0.f;
                        v[u].z = c + 2 < ncols ? __ldg(sp + 2) : 0.f;
                        v[u].w = c + 3 < ncols ? __ldg(sp + 3) : 0.f;
                    }
                    if (mul) m[u] = __ldg(reinterpret_cast<const float4*>(mul + row * ldmul) + c4);
                }
            }
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            if (off[u] < 0) continue;
            float4 x = v[u];
            x.x *= m[u].x; x.y *= m[u].y; x.z *= m[u].z; x.w *= m[u].w;
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

// one thread: stream a pre-packed weight blob (hi | lo halves, each half_floats long) into a ring slot
// with the TMA bulk-copy engine; `full` completes when all bytes have landed
template <bool SPLIT>
__device__ __forceinline__ void tc_bulk_blob(float* slot_hi, float* slot_lo, const float* __restrict__ blob,
                                             int half_floats, int lo_offset_floats, uint64_t* full) {
    const uint32_t bytes = (uint32_t)half_floats * 4u;
    tc::mbar_arrive_expect_tx(full, SPLIT ? 2u * bytes : bytes);
    tc::bulk_g2s(slot_hi, blob, bytes, full);
    if (SPLIT) tc::bulk_g2s(slot_lo, blob + lo_offset_floats, bytes, full);
}

// issue the MMAs of one (A tile rows a_row0.., B blob) product with K = kdim (multiple of 8).
// ONE thread issues them, so its instruction stream is the rate limit for narrow tiles (measured: 112
// cycles per 128x64x8 MMA with a descriptor rebuilt from scratch for every MMA, against ~48 of
// shared-memory operand fetch): the descriptors are built once per pass and advanced by a 32-bit add on
// their address field (start address >> 4 in bits 0..13; two 4-channel chunks per K = 8 step; the tile
// lives below 256 KB so the field cannot carry into the LBO field).
template <bool SPLIT>
__device__ __forceinline__ void tc_issue_kmajor(uint32_t tmem_d, uint32_t a_hi, uint32_t a_lo, uint32_t a_cs_bytes,
                                                int a_row0, uint32_t b_hi, uint32_t b_lo, uint32_t b_cs_bytes,
                                                int kdim, uint32_t idesc, uint32_t& acc) {
    const int npass = SPLIT ? 3 : 1;
    const uint32_t inc_a = (2u * a_cs_bytes) >> 4, inc_b = (2u * b_cs_bytes) >> 4;
    const int nk = kdim >> 3;
    for (int pass = 0; pass < npass; ++pass) {
        const uint32_t as = (SPLIT && pass == 0) ? a_lo : a_hi;      // lo*hi, hi*lo, hi*hi
        const uint32_t bs = (SPLIT && pass == 1) ? b_lo : b_hi;
        const uint64_t da0 = tc::make_smem_desc(as + a_row0 * 16, a_cs_bytes, 128);
        const uint64_t db0 = tc::make_smem_desc(bs, b_cs_bytes, 128);
        uint32_t da_lo = (uint32_t)da0, db_lo = (uint32_t)db0;
        const uint32_t da_hi = (uint32_t)(da0 >> 32), db_hi = (uint32_t)(db0 >> 32);
#pragma unroll 4
        for (int i = 0; i < nk; ++i) {
            tc::umma_tf32(tmem_d, ((uint64_t)da_hi << 32) | da_lo, ((uint64_t)db_hi << 32) | db_lo, idesc, acc);
            acc = 1;
            da_lo += inc_a;
            db_lo += inc_b;
        }
    }
}

// Warp-collective form of tc_issue_kmajor: the WHOLE warp runs the (warp-uniform) descriptor arithmetic and
// loop, one elected lane executes the tcgen05.mma.  With a single divergent lane issuing, every operand of
// every MMA is moved from a vector to a uniform register first (R2UR) and the loop overhead sits on the
// critical path: measured ~50 cycles of issue cost per MMA + ~110 per pass (profiles/mma_rate.py); code
// the compiler can prove uniform stays in the uniform datapath.
template <bool SPLIT>
__device__ __forceinline__ void tc_issue_kmajor_w(uint32_t tmem_d, uint32_t a_hi, uint32_t a_lo, uint32_t a_cs_bytes,
                                                  int a_row0, uint32_t b_hi, uint32_t b_lo, uint32_t b_cs_bytes,
                                                  int kdim, uint32_t idesc, uint32_t& acc) {
    const int npass = SPLIT ? 3 : 1;
    const uint32_t inc_a = (2u * a_cs_bytes) >> 4, inc_b = (2u * b_cs_bytes) >> 4;
    const int nk = kdim >> 3;
    const bool leader = tc::elect_one();
    for (int pass = 0; pass < npass; ++pass) {
        const uint32_t as = (SPLIT && pass == 0) ? a_lo : a_hi;
        const uint32_t bs = (SPLIT && pass == 1) ? b_lo : b_hi;
        const uint64_t da0 = tc::make_smem_desc(as + a_row0 * 16, a_cs_bytes, 128);
        const uint64_t db0 = tc::make_smem_desc(bs, b_cs_bytes, 128);
        uint32_t da_lo = (uint32_t)da0, db_lo = (uint32_t)db0;
        const uint32_t da_hi = (uint32_t)(da0 >> 32), db_hi = (uint32_t)(db0 >> 32);
#pragma unroll 4
        for (int i = 0; i < nk; ++i) {
            if (leader) tc::umma_tf32(tmem_d, ((uint64_t)da_hi << 32) | da_lo, ((uint64_t)db_hi << 32) | db_lo, idesc, acc);
            acc = 1;
            da_lo += inc_a;
            db_lo += inc_b;
        }
    }
}

// TS form: A (128 rows x kdim) lives in tensor memory at columns a_hi.. / a_lo.. (one column per k element)
template <bool SPLIT>
__device__ __forceinline__ void tc_issue_ts(uint32_t tmem_d, uint32_t a_hi, uint32_t a_lo, uint32_t b_hi, uint32_t b_lo,
                                            uint32_t b_cs_bytes, int kdim, uint32_t idesc, uint32_t& acc) {
    const int npass = SPLIT ? 3 : 1;
    const uint32_t inc_b = (2u * b_cs_bytes) >> 4;
    const int nk = kdim >> 3;
    for (int pass = 0; pass < npass; ++pass) {
        uint32_t ta = (SPLIT && pass == 0) ? a_lo : a_hi;            // lo*hi, hi*lo, hi*hi
        const uint32_t bs = (SPLIT && pass == 1) ? b_lo : b_hi;
        const uint64_t db0 = tc::make_smem_desc(bs, b_cs_bytes, 128);
        uint32_t db_lo = (uint32_t)db0;
        const uint32_t db_hi = (uint32_t)(db0 >> 32);
#pragma unroll 4
        for (int i = 0; i < nk; ++i) {
            tc::umma_tf32_ts(tmem_d, ta, ((uint64_t)db_hi << 32) | db_lo, idesc, acc);
            acc = 1;
            ta += 8;
            db_lo += inc_b;
        }
    }
}

__device__ __forceinline__ float gate_sigmoid(float x) { return __fdividef(1.f, 1.f + __expf(-x)); }
__device__ __forceinline__ float gate_tanh(float x) { return 1.f - __fdividef(2.f, __expf(2.f * x) + 1.f); }

template <bool SPLIT>
__global__ void __launch_bounds__(256, SPLIT ? 1 : 2) k_resblock_fwd_tc(const ResFwdTcParams q) {
    const ResFwdParams& p = q.p;
    extern __shared__ float4 crk_smem4[];
    float* smem = reinterpret_cast<float*>(crk_smem4);
    __shared__ uint64_t bar_full[2];     // TMA bytes of a weight blob have landed in ring slot s
    __shared__ uint64_t bar_free[2];     // the MMAs reading ring slot s have completed
    __shared__ uint64_t bar_acc[3];      // 0: tap MMAs done, 1: GEMM1 (incl. aux) done, 2: GEMM2 done
    __shared__ uint32_t tmem_base_s;
    __shared__ int timeout_s;

    const int tiles_per_utt = (p.T + CRK_TC_TM - 1) / CRK_TC_TM;
    const int b = blockIdx.x / tiles_per_utt;
    const int t0 = (blockIdx.x - b * tiles_per_utt) * CRK_TC_TM;
    const int halo = (p.k - 1) * p.dil;
    const int rowsX = CRK_TC_TM + halo;
    const int csx = tc::chunk_rows(rowsX) * 4;          // floats per X chunk (odd row count)
    constexpr int CRW = 129, CSW = CRW * 4;             // weight / z / aux tiles: 128 rows -> 129
    constexpr int WHALF = 16 * CSW;                     // floats of one 64-K blob half
    const int xhalf = 16 * csx;
    float* Xh = smem;                                   // region A: X tile (hi|lo); later the aux tile, then z
    float* Xl = Xh + xhalf;
    float* ring = Xl + xhalf;                           // 2 slots x (hi | lo)
    float* slot_hi[2] = {ring, ring + 2 * WHALF};
    float* slot_lo[2] = {ring + WHALF, ring + 3 * WHALF};
    float* Zh = smem;
    float* Zl = Zh + WHALF;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const bool has_aux = p.Ca > 0;
    const int nblobs = p.k + (has_aux ? 1 : 0) + 1;     // taps, [aux], out|skip
    const int kcha = q.KaPad >> 2;

    // blob b: source, half length (floats)
    auto blob_src = [&](int bi) -> const float* {
        if (bi < p.k) return q.WcTc + (size_t)bi * 2 * WHALF;
        if (has_aux && bi == p.k) return q.WaTc;
        return q.WosTc;
    };
    auto blob_half = [&](int bi) -> int { return (has_aux && bi == p.k) ? kcha * CSW : WHALF; };

    if (threadIdx.x == 0) {
        tc::mbar_init(&bar_full[0], 1); tc::mbar_init(&bar_full[1], 1);
        tc::mbar_init(&bar_free[0], 1); tc::mbar_init(&bar_free[1], 1);
        tc::mbar_init(&bar_acc[0], 1); tc::mbar_init(&bar_acc[1], 1); tc::mbar_init(&bar_acc[2], 1);
        tc::fence_mbar_init();
        timeout_s = 0;
    }
    if (warp == 1) tc::tmem_alloc<256>(&tmem_base_s);
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    const uint32_t tmem = tmem_base_s;
    bool ok = true;
    pdl_trigger();                                      // only now: this CTA already owns its TMEM columns (see crk_common.cuh)
    pdl_wait();                                         // predecessor complete: global memory may be touched
    dbg_stamp(q.dbg, 0);

    // ---- producer: first two blobs in flight while everybody stages the activation tile ----
    if (threadIdx.x == 0) {
        for (int bi = 0; bi < 2 && bi < nblobs; ++bi)
            tc_bulk_blob<SPLIT>(slot_hi[bi & 1], slot_lo[bi & 1], blob_src(bi), blob_half(bi), blob_half(bi), &bar_full[bi & 1]);
    }
    tc_stage_act<SPLIT, SPLIT ? 9 : 5>(Xh, Xl, csx, p.Hin, 64, 64, 64, b, p.T, t0 - p.padl, rowsX, p.dropmul, 64);
    tc::fence_proxy_async_smem();
    __syncthreads();
    dbg_stamp(q.dbg, 1);

    const uint32_t idesc = tc::make_idesc_tf32(128, 128, 0, 0);
    // one elected lane per role; the other 31 lanes of that warp park at __syncwarp (no spinning)
    if (warp == 0) {
        // ===== TMA producer: refill a ring slot as soon as the MMAs that read it have completed =====
        if (lane == 0)
            for (int bi = 2; bi < nblobs; ++bi) {
                ok &= tc::mbar_wait(&bar_free[bi & 1], ((bi - 2) >> 1) & 1);
                tc_bulk_blob<SPLIT>(slot_hi[bi & 1], slot_lo[bi & 1], blob_src(bi), blob_half(bi), blob_half(bi), &bar_full[bi & 1]);
            }
        __syncwarp();
    } else if (warp == 1) {
        // ===== MMA issuer: conv taps (warp-collective loop, one elected lane issues: tc_issue_kmajor_w) =====
        const uint32_t xh_s = tc::smem_u32(Xh), xl_s = tc::smem_u32(Xl);
        uint32_t acc = 0;
        long long wfull = 0;
        for (int j = 0; j < p.k; ++j) {
            const long long c0 = q.dbg ? clock64() : 0;
            ok &= tc::mbar_wait(&bar_full[j & 1], (j >> 1) & 1);
            if (q.dbg) wfull += clock64() - c0;
            tc::tc_fence_after();
            tc_issue_kmajor_w<SPLIT>(tmem, xh_s, xl_s, csx * 4, j * p.dil, tc::smem_u32(slot_hi[j & 1]),
                                     tc::smem_u32(slot_lo[j & 1]), CSW * 4, 64, idesc, acc);
            if (tc::elect_one()) tc::umma_commit(&bar_free[j & 1]);
        }
        if (tc::elect_one()) {
            tc::umma_commit(&bar_acc[0]);
            if (!has_aux) tc::umma_commit(&bar_acc[1]);
        }
        if (lane == 0) dbg_put(q.dbg, 8, wfull);
        __syncwarp();
    }
    // ---- aux 1x1 (decoder 0): its tile reuses the X region -> all tap MMAs must have completed ----
    if (has_aux) {
        ok &= tc::mbar_wait(&bar_acc[0], 0);
        tc::tc_fence_after();
        float* Ch = smem;
        float* Cl = Ch + kcha * CSW;
        tc_stage_act<SPLIT>(Ch, Cl, CSW, p.Caux, p.ldc, p.Ca, q.KaPad, b, p.T, t0, CRK_TC_TM, nullptr, 0);
        tc::fence_proxy_async_smem();
        tc::tc_fence_before();
        __syncthreads();
        tc::tc_fence_after();
        if (warp == 1) {
            const int bi = p.k;
            uint32_t acc = 1;
            ok &= tc::mbar_wait(&bar_full[bi & 1], (bi >> 1) & 1);
            tc::tc_fence_after();
            tc_issue_kmajor_w<SPLIT>(tmem, tc::smem_u32(Ch), tc::smem_u32(Cl), CSW * 4, 0, tc::smem_u32(slot_hi[bi & 1]),
                                     tc::smem_u32(slot_lo[bi & 1]), CSW * 4, q.KaPad, idesc, acc);
            if (tc::elect_one()) {
                tc::umma_commit(&bar_free[bi & 1]);
                tc::umma_commit(&bar_acc[1]);
            }
            __syncwarp();
        }
    }
    ok &= tc::mbar_wait(&bar_acc[1], 0);
    tc::tc_fence_after();
    dbg_stamp(q.dbg, 2);

    // ---- epilogue 1: gate ----
    // TMEM -> registers is thread-per-row; a thread-per-row GLOBAL access pattern costs 32 line
    // transactions per warp instruction (measured: 11K + 17K cycles per tile in the two epilogues), so
    // results are transposed through a padded shared-memory tile and written out row-coalesced.
    const int r = (warp & 3) * 32 + lane;           // frame row in the tile == TMEM lane
    const int hh = warp >> 2;                       // column half
    const uint32_t tlane = tmem + ((uint32_t)((warp & 3) * 32) << 16);
    constexpr int SST = 129;                        // staging row stride (odd: conflict-free column writes)
    float* S1 = slot_hi[(nblobs - 2) & 1];          // ring slot NOT holding [out|skip] (free: GEMM1 is done)
#pragma unroll
    for (int cc = 0; cc < 2; ++cc) {
        const int col0 = hh * 64 + cc * 32;
        float v[32];
        tc::tmem_ld32(tlane + col0, v);
#pragma unroll
        for (int g = 0; g < 8; ++g) {
            const int qi = (col0 >> 2) + g;          // gate pair index: channels 2qi, 2qi+1
            const float4 bv = __ldg(reinterpret_cast<const float4*>(p.bc) + qi);
            // tanh / sigmoid through ex2.approx + approximate division: absolute error ~1e-7 (two orders
            // below the 3xTF32 noise), ~8x fewer instructions than tanhf/expf (the gate was the longest
            // phase of the kernel: 9K of 33K cycles per tile)
            const float ta0 = gate_tanh(v[4 * g + 0] + bv.x);
            const float ta1 = gate_tanh(v[4 * g + 1] + bv.y);
            const float sb0 = gate_sigmoid(v[4 * g + 2] + bv.z);
            const float sb1 = gate_sigmoid(v[4 * g + 3] + bv.w);
            if (p.TaSb) {
                float* sp = S1 + r * SST + 4 * qi;
                sp[0] = ta0; sp[1] = ta1; sp[2] = sb0; sp[3] = sb1;
            }
            const float z0 = ta0 * sb0, z1 = ta1 * sb1;
            const int zo = (qi >> 1) * CSW + r * 4 + 2 * (qi & 1);
            if (SPLIT) {
                float h0, l0, h1, l1;
                tc::split_tf32(z0, h0, l0); tc::split_tf32(z1, h1, l1);
                *reinterpret_cast<float2*>(Zh + zo) = make_float2(h0, h1);
                *reinterpret_cast<float2*>(Zl + zo) = make_float2(l0, l1);
            } else {
                *reinterpret_cast<float2*>(Zh + zo) = make_float2(z0, z1);
            }
        }
    }
    tc::fence_proxy_async_smem();
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    dbg_stamp(q.dbg, 3);

    // ---- GEMM2: [out | skip] (async) while every warp streams the saved gates out, row-coalesced ----
    if (warp == 1) {
        const int bi = nblobs - 1;
        uint32_t acc2 = 0;
        ok &= tc::mbar_wait(&bar_full[bi & 1], (bi >> 1) & 1);
        tc::tc_fence_after();
        tc_issue_kmajor_w<SPLIT>(tmem + 128, tc::smem_u32(Zh), tc::smem_u32(Zl), CSW * 4, 0, tc::smem_u32(slot_hi[bi & 1]),
                                 tc::smem_u32(slot_lo[bi & 1]), CSW * 4, 64, idesc, acc2);
        if (tc::elect_one()) tc::umma_commit(&bar_acc[2]);
    }
    __syncwarp();
    const int nlive = min(CRK_TC_TM, p.T - t0);     // valid rows of this tile
    if (p.TaSb) {
        float* dstbase = p.TaSb + ((size_t)b * p.T + t0) * 128;
        for (int rr = warp; rr < nlive; rr += 8) {
            const float* sp = S1 + rr * SST;
            float* dp = dstbase + (size_t)rr * 128;
            const float a0 = sp[lane], a1 = sp[lane + 32], a2 = sp[lane + 64], a3 = sp[lane + 96];
            dp[lane] = a0; dp[lane + 32] = a1; dp[lane + 64] = a2; dp[lane + 96] = a3;
        }
    }
    ok &= tc::mbar_wait(&bar_acc[2], 0);
    tc::tc_fence_after();
    dbg_stamp(q.dbg, 4);

    // ---- epilogue 2: (acc2 + bias) -> padded smem tile -> coalesced residual / skip pass ----
    if (!ok) timeout_s = 1;
    float* S2 = smem;                               // X/z region: free, GEMM2 has completed
#pragma unroll
    for (int cc = 0; cc < 2; ++cc) {
        const int col0 = hh * 64 + cc * 32;
        float v[32];
        tc::tmem_ld32(tlane + 128 + col0, v);
#pragma unroll
        for (int g = 0; g < 8; ++g) {
            const int qi = (col0 >> 2) + g;
            const float4 bv = __ldg(reinterpret_cast<const float4*>(p.bos) + qi);
            float* sp = S2 + r * SST + 4 * qi;
            sp[0] = v[4 * g + 0] + bv.x; sp[1] = v[4 * g + 1] + bv.y;
            sp[2] = v[4 * g + 2] + bv.z; sp[3] = v[4 * g + 3] + bv.w;
        }
    }
    __syncthreads();
    {
        const size_t base = ((size_t)b * p.T + t0) * 64;
        // lane owns channels (2*lane, 2*lane+1): packed columns 4*lane+{0,1} = out, 4*lane+{2,3} = skip
        for (int rr0 = warp; rr0 < nlive; rr0 += 32) {
            float2 res[4], sko[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int rr = rr0 + 8 * u;
                if (rr < nlive) {
                    res[u] = __ldg(reinterpret_cast<const float2*>(p.Hin + base + (size_t)rr * 64) + lane);
                    if (!p.skip_init) sko[u] = *(reinterpret_cast<const float2*>(p.Skip + base + (size_t)rr * 64) + lane);
                }
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int rr = rr0 + 8 * u;
                if (rr >= nlive) continue;
                const float* sp = S2 + rr * SST + 4 * lane;
                float2 ho, sk;
                ho.x = (sp[0] + res[u].x) * CRK_SQRT_HALF;
                ho.y = (sp[1] + res[u].y) * CRK_SQRT_HALF;
                sk.x = sp[2]; sk.y = sp[3];
                if (!p.skip_init) { sk.x += sko[u].x; sk.y += sko[u].y; }
                reinterpret_cast<float2*>(p.Hout + base + (size_t)rr * 64)[lane] = ho;
                reinterpret_cast<float2*>(p.Skip + base + (size_t)rr * 64)[lane] = sk;
            }
        }
    }
    tc::tc_fence_before();
    __syncthreads();
    dbg_stamp(q.dbg, 5);
    if (timeout_s && threadIdx.x == 0) p.Hout[((size_t)b * p.T + t0) * 64] = __int_as_float(0x7fc00000);  // poison: test must fail
    if (warp == 1) tc::tmem_dealloc<256>(tmem);
}

inline size_t resblock_fwd_tc_smem(int k, int dil) {
    const int rowsX = CRK_TC_TM + (k - 1) * dil;
    return (size_t)(2 * 16 * tc::chunk_rows(rowsX) * 4 + 4 * 16 * 129 * 4) * sizeof(float);
}

template <bool SPLIT>
inline cudaError_t launch_resblock_fwd_tc(const ResFwdTcParams& q, cudaStream_t s) {
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(k_resblock_fwd_tc<SPLIT>, cudaFuncAttributeMaxDynamicSharedMemorySize, 225 * 1024);
        if (e != cudaSuccess) return e;
        attr_set = true;
    }
    const int tiles = q.p.B * cdiv(q.p.T, CRK_TC_TM);
    TimedLaunch tl(CRK_K_RESBLOCK_FWD, s, 2.0 * q.p.B * q.p.T * (64.0 * 128 * q.p.k + q.p.Ca * 128.0 + 64.0 * 128));
    ResFwdTcParams qq = q;
    qq.dbg = dbg_take(CRK_K_RESBLOCK_FWD);
    cudaError_t le = launch_pdl(k_resblock_fwd_tc<SPLIT>, dim3(tiles), dim3(256), resblock_fwd_tc_smem(q.p.k, q.p.dil), s, qq);
    if (le != cudaSuccess) return le;
    return launch_check();
}

}  // namespace crk
