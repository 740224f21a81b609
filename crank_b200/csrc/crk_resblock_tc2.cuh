// crank-b200: fused WaveNet residual block FORWARD, round 2: the round-1 kernel (crk_resblock_tc.cuh) on a shared-memory
// diet so that TWO CTAs share an SM.
//
// Same math, operands and packed weight blobs as k_resblock_fwd_tc (parallel_wavegan ResidualBlock.forward as built at
// crank/net/module/vqvae2.py:236-273 / crank/bin/train.py:107-115).  Round 1 needed 202 KB per CTA (X hi|lo 70 KB + a
// weight ring of two 66 KB slots), so the 256 tiles of a 64 x 500-frame launch ran as two strictly serial waves on 148 SMs,
// all CTAs in lock step: HBM idle during the MMAs, the tensor pipe idle during staging / epilogues.  Here
//   * the weight ring holds SEG = 4 channel chunks (K = 16) per slot, hi | lo: 2 x 16.5 KB instead of 2 x 66 KB;
//   * the saved gates go from registers straight to global memory (16x256b tensor-memory fragments: every quad of lanes
//     writes one full 32 B sector), so the 66 KB transposition tile of epilogue 1 is gone;
//   * <= 128 registers (launch bound 2 CTAs/SM);
// = 103..107 KB per CTA: all 256 tiles are resident at once and one CTA's MMAs run under the other's global-memory
// phases.  The weight producer advances in three stages (taps / aux / out|skip) because every warp -- including the
// producer's -- takes part in the staging and epilogue phases between them.
#pragma once
#include "crk_common.cuh"
#include "crk_resblock_pt.cuh"
#include "crk_resblock_tc.cuh"
#include "crk_tc.cuh"

namespace crk {

template <bool SPLIT>
__global__ void __launch_bounds__(256, 2) k_resblock_fwd_tc2(const ResFwdTcParams q) {
    const ResFwdParams& p = q.p;
    extern __shared__ float4 crk_smem4[];
    float* smem = reinterpret_cast<float*>(crk_smem4);
    constexpr int NSLOT = 2;
    constexpr int SEG = SPLIT ? 4 : 8;                 // channel chunks per ring slot
    __shared__ uint64_t bar_full[NSLOT];
    __shared__ uint64_t bar_free[NSLOT];
    __shared__ uint64_t bar_acc[3];                    // 0: tap MMAs done, 1: GEMM1 (incl. aux) done, 2: GEMM2 done
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
    float* Xh = smem;                                   // region A: X tile (hi|lo); later the aux tile, then z, then S2
    float* Xl = Xh + xhalf;
    float* ring = Xh + (SPLIT ? 2 : 1) * xhalf;
    constexpr int SLOTF = SEG * CSW;                    // floats of one slot half
    auto slot_hi = [&](int sl) -> float* { return ring + sl * (SPLIT ? 2 : 1) * SLOTF; };
    auto slot_lo = [&](int sl) -> float* { return ring + sl * (SPLIT ? 2 : 1) * SLOTF + SLOTF; };
    float* Zh = smem;
    float* Zl = Zh + WHALF;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const bool has_aux = p.Ca > 0;
    const int kcha = q.KaPad >> 2;
    constexpr int NSEG64 = 16 / SEG;                    // steps per 64-channel blob
    const int nseg_aux = has_aux ? (kcha + SEG - 1) / SEG : 0;
    const int nt = p.k * NSEG64;                        // tap steps
    const int n1 = nt + nseg_aux;                       // GEMM1 steps
    const int nsteps = n1 + NSEG64;                     // + [out|skip]

    // step -> weight slice (TMA bulk copy into the step's ring slot)
    auto produce = [&](int st) {
        const float* blob;
        int kch_blob, sg;
        if (st < nt) { const int j = st / NSEG64; sg = st - j * NSEG64; blob = q.WcTc + (size_t)j * 2 * WHALF; kch_blob = 16; }
        else if (st < n1) { sg = st - nt; blob = q.WaTc; kch_blob = kcha; }
        else { sg = st - n1; blob = q.WosTc; kch_blob = 16; }
        const int ch0 = sg * SEG;
        const int nch = (kch_blob - ch0) < SEG ? (kch_blob - ch0) : SEG;
        const int sl = st & (NSLOT - 1);
        tc_bulk_blob<SPLIT>(slot_hi(sl), slot_lo(sl), blob + (size_t)ch0 * CSW, nch * CSW, kch_blob * CSW, &bar_full[sl]);
    };

    if (threadIdx.x == 0) {
        for (int i = 0; i < NSLOT; ++i) { tc::mbar_init(&bar_full[i], 1); tc::mbar_init(&bar_free[i], 1); }
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

    int produced = 0;                                   // (thread 0 only)
    if (threadIdx.x == 0)
        for (; produced < NSLOT && produced < nsteps; ++produced) produce(produced);
    // thread 0: refill ring slots up to step `upto` (exclusive); a slot is free once the MMAs of the step that used it completed
    auto producer_run = [&](int upto) {
        if (upto > nsteps) upto = nsteps;
        for (; produced < upto; ++produced) {
            ok &= tc::mbar_wait(&bar_free[produced & (NSLOT - 1)], ((produced - NSLOT) / NSLOT) & 1);
            produce(produced);
        }
    };
    // warp 1: issue the MMAs of steps [s0, s1) whose A operand is the K-major tile (a_hi | a_lo) of chunk stride a_cs floats
    const uint32_t idesc = tc::make_idesc_tf32(128, 128, 0, 0);
    auto issue_steps = [&](int s0, int s1, int step_base, int nseg_blob, int kch_blob, uint32_t tmem_d, const float* a_hi,
                           const float* a_lo, int a_cs, bool taps, uint32_t& acc) {
        for (int st = s0; st < s1; ++st) {
            const int rem = st - step_base;
            const int j = rem / nseg_blob, sg = rem - j * nseg_blob;
            const int ch0 = sg * SEG;
            const int nch = (kch_blob - ch0) < SEG ? (kch_blob - ch0) : SEG;
            const int sl = st & (NSLOT - 1);
            ok &= tc::mbar_wait(&bar_full[sl], (st / NSLOT) & 1);
            tc::tc_fence_after();
            tc_issue_kmajor_w<SPLIT>(tmem_d, tc::smem_u32(a_hi + ch0 * a_cs), tc::smem_u32(a_lo + ch0 * a_cs), a_cs * 4,
                                     taps ? j * p.dil : 0, tc::smem_u32(slot_hi(sl)), tc::smem_u32(slot_lo(sl)), CSW * 4,
                                     nch * 4, idesc, acc);
            if (tc::elect_one()) tc::umma_commit(&bar_free[sl]);
        }
    };

    tc_stage_act<SPLIT, 4>(Xh, Xl, csx, p.Hin, 64, 64, 64, b, p.T, t0 - p.padl, rowsX, p.dropmul, 64);
    tc::fence_proxy_async_smem();
    __syncthreads();
    dbg_stamp(q.dbg, 1);

    // ---- GEMM1, conv taps ----
    uint32_t acc1 = 0;
    if (warp == 0) {
        if (lane == 0) producer_run(nt + NSLOT);
        __syncwarp();
    } else if (warp == 1) {
        issue_steps(0, nt, 0, NSEG64, 16, tmem, Xh, Xl, csx, true, acc1);
        if (tc::elect_one()) {
            tc::umma_commit(&bar_acc[0]);
            if (!has_aux) tc::umma_commit(&bar_acc[1]);
        }
        __syncwarp();
    }
    // ---- aux 1x1 (decoder 0): its tile reuses the X region -> all tap MMAs must have completed ----
    if (has_aux) {
        ok &= tc::mbar_wait(&bar_acc[0], 0);
        tc::tc_fence_after();
        float* Ch = smem;
        float* Cl = Ch + kcha * CSW;
        tc_stage_act<SPLIT, 4>(Ch, Cl, CSW, p.Caux, p.ldc, p.Ca, q.KaPad, b, p.T, t0, CRK_TC_TM, nullptr, 0);
        tc::fence_proxy_async_smem();
        tc::tc_fence_before();
        __syncthreads();
        tc::tc_fence_after();
        if (warp == 0) {
            if (lane == 0) producer_run(n1 + NSLOT);
            __syncwarp();
        } else if (warp == 1) {
            issue_steps(nt, n1, nt, nseg_aux, kcha, tmem, Ch, Cl, CSW, false, acc1);
            if (tc::elect_one()) tc::umma_commit(&bar_acc[1]);
            __syncwarp();
        }
    }
    ok &= tc::mbar_wait(&bar_acc[1], 0);
    tc::tc_fence_after();
    dbg_stamp(q.dbg, 2);

    // ---- epilogue 1: gate.  16x256b fragments: reg[4b + 2h + e] = (lane0 + t/4 + 8h, col0 + 8b + 2(t%4) + e); even lanes
    //      hold the tanh pair of a gate quad, odd lanes its sigmoid pair (interleaved gate order), partners differ in bit 0 ----
    const int wq = warp & 3;                        // TMEM lane quarter this warp may access
    const int hh = warp >> 2;                       // column half
    const int nlive = min(CRK_TC_TM, p.T - t0);     // valid rows of this tile
    const size_t row0 = (size_t)b * p.T + t0;
    {
        const bool odd = lane & 1;
        const float s_arg = odd ? -1.f : 2.f;       // sigmoid: exp(-x);  tanh: exp(2x)
        const float s_num = odd ? 1.f : 2.f;
#pragma unroll
        for (int half = 0; half < 2; ++half) {
            const int lane0 = wq * 32 + half * 16;
            float v[32];
            tmem_ld_16x256b_x8(tmem + ((uint32_t)lane0 << 16) + hh * 64, v);
#pragma unroll
            for (int bb = 0; bb < 8; ++bb) {
                const int col = hh * 64 + 8 * bb + 2 * (lane & 3);       // packed column of e = 0
                const float2 bv = __ldg(reinterpret_cast<const float2*>(p.bc + col));
                const int qi = col >> 2;                                 // gate quad: z channels 2qi, 2qi+1
#pragma unroll
                for (int hr = 0; hr < 2; ++hr) {
                    const int r = lane0 + (lane >> 2) + 8 * hr;
                    // same expressions as gate_tanh / gate_sigmoid: tanh = 1 - 2/(exp(2x)+1), sigmoid = 1/(1+exp(-x))
                    const float e0 = __expf(s_arg * (v[4 * bb + 2 * hr] + bv.x));
                    const float e1 = __expf(s_arg * (v[4 * bb + 2 * hr + 1] + bv.y));
                    const float d0 = __fdividef(s_num, e0 + 1.f), d1 = __fdividef(s_num, e1 + 1.f);
                    const float a0 = odd ? d0 : 1.f - d0, a1 = odd ? d1 : 1.f - d1;
                    if (p.TaSb && r < nlive)
                        *reinterpret_cast<float2*>(p.TaSb + (row0 + r) * 128 + col) = make_float2(a0, a1);
                    const float o0 = __shfl_xor_sync(0xffffffffu, a0, 1), o1 = __shfl_xor_sync(0xffffffffu, a1, 1);
                    const float z0 = a0 * o0, z1 = a1 * o1;              // tanh * sigmoid (both lanes of the pair)
                    const int zo = (qi >> 1) * CSW + r * 4 + 2 * (qi & 1);
                    if (SPLIT) {
                        float h0, l0, h1, l1;
                        tc::split_tf32(z0, h0, l0); tc::split_tf32(z1, h1, l1);
                        if (!odd) *reinterpret_cast<float2*>(Zh + zo) = make_float2(h0, h1);
                        else *reinterpret_cast<float2*>(Zl + zo) = make_float2(l0, l1);
                    } else if (!odd) {
                        *reinterpret_cast<float2*>(Zh + zo) = make_float2(z0, z1);
                    }
                }
            }
        }
    }
    tc::fence_proxy_async_smem();
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    dbg_stamp(q.dbg, 3);

    // ---- GEMM2: [out | skip] ----
    // the residual / old-skip rows of epilogue 2 do not depend on it: fetched while the MMAs run (a second CTA shares the SM,
    // so this is the only latency hiding left inside the CTA).  warp w owns rows w + 8j, lane owns channels (2 lane, 2 lane + 1)
    if (warp == 0) {
        if (lane == 0) producer_run(nsteps);
        __syncwarp();
    } else if (warp == 1) {
        uint32_t acc2 = 0;
        issue_steps(n1, nsteps, n1, NSEG64, 16, tmem + 128, Zh, Zl, CSW, false, acc2);
        if (tc::elect_one()) tc::umma_commit(&bar_acc[2]);
        __syncwarp();
    }
    ok &= tc::mbar_wait(&bar_acc[2], 0);
    tc::tc_fence_after();
    dbg_stamp(q.dbg, 4);

    // ---- epilogue 2: (acc2 + bias) -> padded smem tile -> coalesced residual / skip pass ----
    if (!ok) timeout_s = 1;
    constexpr int SST = 129;                        // staging row stride (odd: conflict-free column writes)
    float* S2 = smem;                               // X/z region: free, GEMM2 has completed
    {
        const int r = wq * 32 + lane;               // frame row in the tile == TMEM lane
        const uint32_t tlane = tmem + ((uint32_t)(wq * 32) << 16);
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
    }
    __syncthreads();
    {
        const size_t base = row0 * 64;
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
    if (timeout_s && threadIdx.x == 0) p.Hout[row0 * 64] = __int_as_float(0x7fc00000);  // poison: test must fail
    if (warp == 1) tc::tmem_dealloc<256>(tmem);
}

inline size_t resblock_fwd_tc2_smem(int k, int dil, bool split) {
    const int rowsX = CRK_TC_TM + (k - 1) * dil;
    const size_t x = (size_t)16 * tc::chunk_rows(rowsX) * 4 * (split ? 2 : 1);
    const size_t a = x > (size_t)CRK_TC_TM * 129 ? x : (size_t)CRK_TC_TM * 129;      // region A also hosts z and the S2 tile
    const size_t ring = (size_t)2 * (split ? 2 : 1) * (split ? 4 : 8) * 129 * 4;
    return (a + ring) * sizeof(float);
}
inline bool resblock_fwd_tc2_ok(const ResFwdTcParams& q, bool split) {
    if (q.p.Ca > 0 && (q.KaPad & 7)) return false;
    return resblock_fwd_tc2_smem(q.p.k, q.p.dil, split) <= 112 * 1024;
}

template <bool SPLIT>
inline cudaError_t launch_resblock_fwd_tc2(const ResFwdTcParams& q, cudaStream_t s) {
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(k_resblock_fwd_tc2<SPLIT>, cudaFuncAttributeMaxDynamicSharedMemorySize, 112 * 1024);
        if (e != cudaSuccess) return e;
        attr_set = true;
    }
    const int tiles = q.p.B * cdiv(q.p.T, CRK_TC_TM);
    TimedLaunch tl(CRK_K_RESBLOCK_FWD, s, 2.0 * q.p.B * q.p.T * (64.0 * 128 * q.p.k + q.p.Ca * 128.0 + 64.0 * 128));
    ResFwdTcParams qq = q;
    qq.dbg = dbg_take(CRK_K_RESBLOCK_FWD);
    cudaError_t le = launch_pdl(k_resblock_fwd_tc2<SPLIT>, dim3(tiles), dim3(256), resblock_fwd_tc2_smem(q.p.k, q.p.dil, SPLIT), s, qq);
    if (le != cudaSuccess) return le;
    return launch_check();
}

}  // namespace crk
