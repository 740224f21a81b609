// crank-b200: PERSISTENT, warp-specialised fused residual-block forward (round 2).
//
// Same math, operands and packed layouts as k_resblock_fwd_tc (crk_resblock_tc.cuh) -- parallel_wavegan
// ResidualBlock.forward as built at crank/net/module/vqvae2.py:236-273 / crank/bin/train.py:107-115 -- but the
// phases of consecutive tiles overlap instead of running back to back (round-1 profile: stage 3.4K / GEMM1 9.8K /
// gate 3.9K / GEMM2 2.9K / epilogue 7.0K cycles, strictly serial, tensor pipe 21 % active):
//
//   grid = min(#tiles, #SMs) CTAs, CTA c owns tiles c, c+grid, ...           320 threads = 10 warps
//     warp 0      TMA producer : streams weight HALF-blobs (hi or lo, K = 64, 33 KB) through a 2-slot ring
//     warp 1      MMA issuer   : GEMM1(i) -> TMEM acc1[i&1];  GEMM2(i-1) is issued in the MIDDLE of GEMM1(i)
//     warps 2..9  workers      : E1(i) gate epilogue (acc1 -> tanh*sigmoid -> saved gates, z tile), E2(i) output
//                                epilogue (acc2 -> residual / skip), and the staging of X(i+2), in that order
//   shared memory: two X tiles (hi|lo, with halo) + the ring.  z(i) aliases X(i)'s buffer (free once GEMM1(i) has
//   completed); E2(i) transposes through the same buffer (free once GEMM2(i) has completed) before X(i+2) lands there.
//   tensor memory: acc1 x 2 (columns 0, 128), acc2 (256), aux operand hi|lo (384, 448).
//   The aux 1x1 (decoder 0) reads its A operand from TENSOR MEMORY (tcgen05.mma TS form): there is no shared memory
//   left for a third activation tile, and K = 34 makes it a 15-instruction GEMM.
// Hardware facts this relies on are pinned by profiles/hwprobe (TS form numerics, aligned TMEM operand columns,
// the 16x256b tcgen05.ld fragment map) and tests/test_gpu_tc.py.
#pragma once
#include "crk_common.cuh"
#include "crk_resblock_tc.cuh"
#include "crk_tc.cuh"

namespace crk {

#define CRK_PT_THREADS 320
#define CRK_PT_WORKERS 256

__device__ __forceinline__ void pt_worker_sync() { asm volatile("bar.sync 1, 256;" ::: "memory"); }

// bounded mbarrier wait that gives up at once when another role of this CTA has already timed out (a mis-programmed
// pipeline must fail its test within seconds, not serialise a time-out per barrier)
__device__ __forceinline__ bool pt_wait(uint64_t* bar, uint32_t parity, volatile int* abort_flag) {
    for (uint32_t it = 0; it < (1u << 24); ++it) {
        if (tc::mbar_try_wait(bar, parity)) return true;
        if ((it & 1023u) == 1023u && *abort_flag) return false;
    }
    *abort_flag = 1;
    return false;
}

// one pass of warp-collectively issued MMAs, both operands in shared memory (K-major chunk tiles)
__device__ __forceinline__ void pt_issue_ss(uint32_t tmem_d, uint32_t a_s, uint32_t a_cs_bytes, int a_row0, uint32_t b_s,
                                            uint32_t b_cs_bytes, int kdim, uint32_t idesc, uint32_t& acc, bool leader) {
    const uint64_t da0 = tc::make_smem_desc(a_s + a_row0 * 16, a_cs_bytes, 128);
    const uint64_t db0 = tc::make_smem_desc(b_s, b_cs_bytes, 128);
    uint32_t da_lo = (uint32_t)da0, db_lo = (uint32_t)db0;
    const uint32_t da_hi = (uint32_t)(da0 >> 32), db_hi = (uint32_t)(db0 >> 32);
    const uint32_t inc_a = (2u * a_cs_bytes) >> 4, inc_b = (2u * b_cs_bytes) >> 4;
    const int nk = kdim >> 3;
#pragma unroll 4
    for (int i = 0; i < nk; ++i) {
        if (leader) tc::umma_tf32(tmem_d, ((uint64_t)da_hi << 32) | da_lo, ((uint64_t)db_hi << 32) | db_lo, idesc, acc);
        acc = 1;
        da_lo += inc_a;
        db_lo += inc_b;
    }
}
// A operand in tensor memory (lane = row, one column per k element)
__device__ __forceinline__ void pt_issue_ts(uint32_t tmem_d, uint32_t tmem_a, uint32_t b_s, uint32_t b_cs_bytes, int kdim,
                                            uint32_t idesc, uint32_t& acc, bool leader) {
    const uint64_t db0 = tc::make_smem_desc(b_s, b_cs_bytes, 128);
    uint32_t db_lo = (uint32_t)db0;
    const uint32_t db_hi = (uint32_t)(db0 >> 32);
    const uint32_t inc_b = (2u * b_cs_bytes) >> 4;
    const int nk = kdim >> 3;
#pragma unroll 4
    for (int i = 0; i < nk; ++i) {
        if (leader) tc::umma_tf32_ts(tmem_d, tmem_a, ((uint64_t)db_hi << 32) | db_lo, idesc, acc);
        acc = 1;
        tmem_a += 8;
        db_lo += inc_b;
    }
}

// 16 lanes x 64 columns: reg[4b + 2h + e] = (lane0 + t/4 + 8h, col0 + 8b + 2(t%4) + e)   (profiles/hwprobe T5)
__device__ __forceinline__ void tmem_ld_16x256b_x8(uint32_t taddr, float (&v)[32]) {
    uint32_t r[32];
    asm volatile(
        "tcgen05.ld.sync.aligned.16x256b.x8.b32 "
        "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
          "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
          "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr)
        : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}

// 16 lanes x 32 columns: reg[4b + 2h + e], b = 0..3
__device__ __forceinline__ void tmem_ld_16x256b_x4(uint32_t taddr, float (&v)[16]) {
    uint32_t r[16];
    asm volatile(
        "tcgen05.ld.sync.aligned.16x256b.x4.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr)
        : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

// the order in which weight half-blobs are consumed; producer and issuer enumerate the SAME program
//   kind 0: conv tap j of tile `tile`;  1: aux 1x1 of tile `tile`;  2: [out|skip] of tile `tile` (GEMM2)
//   h: half index -- SPLIT: 0 = lo blob (pass A_hi x B_lo), 1 = hi blob (passes A_lo x B_hi, A_hi x B_hi);  !SPLIT: hi
template <bool SPLIT, class F>
__device__ __forceinline__ void pt_program(int n_my, int k, bool has_aux, int ksplit, F&& f) {
    constexpr int NH = SPLIT ? 2 : 1;
    for (int i = 0; i < n_my; ++i) {
        for (int j = 0; j < k; ++j) {
            if (j == ksplit && i >= 1)
                for (int h = 0; h < NH; ++h) f(2, i - 1, 0, h);
            for (int h = 0; h < NH; ++h) f(0, i, j, h);
        }
        if (has_aux)
            for (int h = 0; h < NH; ++h) f(1, i, 0, h);
    }
    if (n_my > 0)
        for (int h = 0; h < NH; ++h) f(2, n_my - 1, 0, h);
}

struct PtXRegs {
    float4 v[9];
    float4 m[9];
};

template <bool SPLIT>
__global__ void __launch_bounds__(CRK_PT_THREADS, 1) k_resblock_fwd_pt(const ResFwdTcParams q) {
    const ResFwdParams& p = q.p;
    extern __shared__ float4 crk_smem4[];
    float* smem = reinterpret_cast<float*>(crk_smem4);
    __shared__ uint64_t bar_rfull[2], bar_rfree[2];     // weight ring
    __shared__ uint64_t bar_xfull[2];                   // X tile staged (256 worker arrivals)
    __shared__ uint64_t bar_acc1[2];                    // GEMM1 (incl. aux) of the tile in acc1[b] has completed
    __shared__ uint64_t bar_zfull[2];                   // z tile written (256 worker arrivals)
    __shared__ uint64_t bar_acc2;                       // GEMM2 has completed
    __shared__ uint64_t bar_aux;                        // aux operand in tensor memory (128 worker arrivals)
    __shared__ uint32_t tmem_base_s;
    __shared__ int timeout_s;

    constexpr int NH = SPLIT ? 2 : 1;
    const int tiles_per_utt = (p.T + CRK_TC_TM - 1) / CRK_TC_TM;
    const int ntiles = p.B * tiles_per_utt;
    const int n_my = (int)blockIdx.x < ntiles ? (ntiles - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;
    const int halo = (p.k - 1) * p.dil;
    const int rowsX = CRK_TC_TM + halo;
    const int csx = tc::chunk_rows(rowsX) * 4;          // floats per X chunk
    constexpr int CRW = 129, CSW = CRW * 4;
    constexpr int WHALF = 16 * CSW;                     // floats of one K = 64 half-blob / of one z half
    const int xhalf = 16 * csx;
    const int xbuf = max(NH * xhalf, CRK_TC_TM * 129);   // also hosts the 128 x 129 epilogue-2 transposition tile
    float* Xb[2] = {smem, smem + xbuf};
    float* ring = smem + 2 * xbuf;
    float* slot[2] = {ring, ring + WHALF};
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const bool has_aux = p.Ca > 0;
    const int kcha = q.KaPad >> 2;
    const int ksplit = min(p.k - 1, (p.k + 1) / 2);

    if (threadIdx.x == 0) {
        for (int i = 0; i < 2; ++i) {
            tc::mbar_init(&bar_rfull[i], 1); tc::mbar_init(&bar_rfree[i], 1);
            tc::mbar_init(&bar_xfull[i], CRK_PT_WORKERS); tc::mbar_init(&bar_acc1[i], 1);
            tc::mbar_init(&bar_zfull[i], CRK_PT_WORKERS);
        }
        tc::mbar_init(&bar_acc2, 1);
        tc::mbar_init(&bar_aux, 128);
        tc::fence_mbar_init();
        timeout_s = 0;
    }
    if (warp == 1) tc::tmem_alloc<512>(&tmem_base_s);
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    const uint32_t tmem = tmem_base_s;
    bool ok = true;
    pdl_trigger();
    pdl_wait();

    auto tile_of = [&](int i, int& b, int& t0) {
        const int tile = (int)blockIdx.x + i * (int)gridDim.x;
        b = tile / tiles_per_utt;
        t0 = (tile - b * tiles_per_utt) * CRK_TC_TM;
    };

    if (warp == 0) {
        // ===================== TMA producer =====================
        if (lane == 0) {
            int step = 0;
            pt_program<SPLIT>(n_my, p.k, has_aux, ksplit, [&](int kind, int, int j, int h) {
                const int sl = step & 1;
                if (step >= 2) ok &= pt_wait(&bar_rfree[sl], ((step - 2) >> 1) & 1, &timeout_s);
                const float* src;
                int halfn;
                if (kind == 0) { src = q.WcTc + (size_t)j * 2 * WHALF; halfn = WHALF; }
                else if (kind == 1) { src = q.WaTc; halfn = kcha * CSW; }
                else { src = q.WosTc; halfn = WHALF; }
                if (SPLIT && h == 0) src += halfn;                       // lo half
                tc::mbar_arrive_expect_tx(&bar_rfull[sl], (uint32_t)halfn * 4u);
                tc::bulk_g2s(slot[sl], src, (uint32_t)halfn * 4u, &bar_rfull[sl]);
                ++step;
            });
            if (!ok) timeout_s = 1;
        }
        __syncwarp();
    } else if (warp == 1) {
        // ===================== MMA issuer (whole warp runs the uniform program, one elected lane issues) ==========
        const uint32_t idesc = tc::make_idesc_tf32(128, 128, 0, 0);
        const bool leader = tc::elect_one();
        int step = 0;
        uint32_t acc1f = 0, acc2f = 0;
        pt_program<SPLIT>(n_my, p.k, has_aux, ksplit, [&](int kind, int tile, int j, int h) {
            const int sl = step & 1, bi = tile & 1;
            const bool lo_blob = SPLIT && h == 0;
            if (h == 0) {
                if (kind == 0 && j == 0) { ok &= pt_wait(&bar_xfull[bi], (tile >> 1) & 1, &timeout_s); acc1f = 0; }
                if (kind == 1) ok &= pt_wait(&bar_aux, tile & 1, &timeout_s);
                if (kind == 2) { ok &= pt_wait(&bar_zfull[bi], (tile >> 1) & 1, &timeout_s); acc2f = 0; }
            }
            ok &= pt_wait(&bar_rfull[sl], (step >> 1) & 1, &timeout_s);
            tc::tc_fence_after();
            const uint32_t b_s = tc::smem_u32(slot[sl]);
            if (kind == 0) {
                const uint32_t xh = tc::smem_u32(Xb[bi]), xl = tc::smem_u32(Xb[bi] + xhalf);
                const uint32_t d = tmem + bi * 128;
                if (lo_blob) pt_issue_ss(d, xh, csx * 4, j * p.dil, b_s, CSW * 4, 64, idesc, acc1f, leader);
                else {
                    if (SPLIT) pt_issue_ss(d, xl, csx * 4, j * p.dil, b_s, CSW * 4, 64, idesc, acc1f, leader);
                    pt_issue_ss(d, xh, csx * 4, j * p.dil, b_s, CSW * 4, 64, idesc, acc1f, leader);
                }
            } else if (kind == 1) {
                const uint32_t d = tmem + bi * 128;
                if (lo_blob) pt_issue_ts(d, tmem + 384, b_s, CSW * 4, q.KaPad, idesc, acc1f, leader);
                else {
                    if (SPLIT) pt_issue_ts(d, tmem + 448, b_s, CSW * 4, q.KaPad, idesc, acc1f, leader);
                    pt_issue_ts(d, tmem + 384, b_s, CSW * 4, q.KaPad, idesc, acc1f, leader);
                }
            } else {
                const uint32_t zh = tc::smem_u32(Xb[bi]), zl = tc::smem_u32(Xb[bi] + WHALF);
                const uint32_t d = tmem + 256;
                if (lo_blob) pt_issue_ss(d, zh, CSW * 4, 0, b_s, CSW * 4, 64, idesc, acc2f, leader);
                else {
                    if (SPLIT) pt_issue_ss(d, zl, CSW * 4, 0, b_s, CSW * 4, 64, idesc, acc2f, leader);
                    pt_issue_ss(d, zh, CSW * 4, 0, b_s, CSW * 4, 64, idesc, acc2f, leader);
                }
            }
            if (leader) {
                tc::umma_commit(&bar_rfree[sl]);
                if (h == NH - 1) {
                    if ((kind == 0 && j == p.k - 1 && !has_aux) || kind == 1) tc::umma_commit(&bar_acc1[bi]);
                    if (kind == 2) tc::umma_commit(&bar_acc2);
                }
            }
            __syncwarp();
            ++step;
        });
        if (!ok) timeout_s = 1;
        __syncwarp();
    } else {
        // ===================== workers =====================
        const int wt = threadIdx.x - 64;                 // 0..255
        const int ww = wt >> 5;                          // worker warp 0..7
        const int wq = warp & 3;                         // TMEM lane quarter this warp may access
        const int hh = ww >> 2;                          // column half handled by this warp
        const uint32_t tq = tmem + ((uint32_t)(wq * 32) << 16);
        PtXRegs XR;

        auto x_load = [&](int i) {
            int b, t0;
            tile_of(i, b, t0);
            const int tstart = t0 - p.padl;
            const int total = rowsX * 16;
#pragma unroll
            for (int u = 0; u < 9; ++u) {
                const int idx = wt + u * CRK_PT_WORKERS;
                XR.v[u] = make_float4(0.f, 0.f, 0.f, 0.f);
                XR.m[u] = make_float4(1.f, 1.f, 1.f, 1.f);
                if (idx < total) {
                    const int r = idx >> 4, c4 = idx & 15;
                    const int tt = tstart + r;
                    if (tt >= 0 && tt < p.T) {
                        const size_t row = (size_t)b * p.T + tt;
                        XR.v[u] = __ldg(reinterpret_cast<const float4*>(p.Hin + row * 64) + c4);
                        if (p.dropmul) XR.m[u] = __ldg(reinterpret_cast<const float4*>(p.dropmul + row * 64) + c4);
                    }
                }
            }
        };
        auto x_store = [&](int bi) {
            float* hi = Xb[bi];
            float* lo = Xb[bi] + xhalf;
            const int total = rowsX * 16;
#pragma unroll
            for (int u = 0; u < 9; ++u) {
                const int idx = wt + u * CRK_PT_WORKERS;
                if (idx >= total) continue;
                const int r = idx >> 4, c4 = idx & 15;
                const int off = c4 * csx + r * 4;
                float4 x = XR.v[u];
                x.x *= XR.m[u].x; x.y *= XR.m[u].y; x.z *= XR.m[u].z; x.w *= XR.m[u].w;
                if (SPLIT) {
                    float4 h, l;
                    tc::split_tf32(x.x, h.x, l.x); tc::split_tf32(x.y, h.y, l.y);
                    tc::split_tf32(x.z, h.z, l.z); tc::split_tf32(x.w, h.w, l.w);
                    *reinterpret_cast<float4*>(hi + off) = h;
                    *reinterpret_cast<float4*>(lo + off) = l;
                } else {
                    *reinterpret_cast<float4*>(hi + off) = x;
                }
            }
            tc::fence_proxy_async_smem();
            tc::tc_fence_before();
            tc::mbar_arrive(&bar_xfull[bi]);
        };
        // aux operand of tile i -> tensor memory (hi: columns 384.., lo: 448..); threads of column-half 0 only
        auto aux_stage = [&](int i) {
            if (!has_aux || hh != 0) return;
            int b, t0;
            tile_of(i, b, t0);
            const int r = wq * 32 + lane;
            const bool live = t0 + r < p.T;
            const float* src = p.Caux + ((size_t)b * p.T + t0 + r) * p.ldc;
            for (int c0 = 0; c0 < q.KaPad; c0 += 32) {
                float h[32], l[32];
#pragma unroll
                for (int e = 0; e < 32; ++e) {
                    const float x = (live && c0 + e < p.Ca) ? __ldg(src + c0 + e) : 0.f;
                    if (SPLIT) tc::split_tf32(x, h[e], l[e]);
                    else { h[e] = x; l[e] = 0.f; }
                }
                tc::tmem_st32(tq + 384 + c0, h);
                if (SPLIT) tc::tmem_st32(tq + 448 + c0, l);
            }
            tc::tmem_st_wait();
            tc::tc_fence_before();
            tc::mbar_arrive(&bar_aux);
        };

        dbg_stamp(q.dbg, 0);
        if (n_my > 0) { x_load(0); x_store(0); }
        if (n_my > 1) { x_load(1); x_store(1); }
        if (n_my > 0) aux_stage(0);
        dbg_stamp(q.dbg, 1);

        for (int i = 0; i < n_my; ++i) {
            const int bi = i & 1;
            int b, t0;
            tile_of(i, b, t0);
            const int nlive = min(CRK_TC_TM, p.T - t0);
            const size_t row0 = (size_t)b * p.T + t0;

            // ---------------- E1: gate ----------------
            ok &= pt_wait(&bar_acc1[bi], (i >> 1) & 1, &timeout_s);
            tc::tc_fence_after();
            if (i == 0) dbg_stamp(q.dbg, 2);
            {
                float* Zh = Xb[bi];
                float* Zl = Xb[bi] + WHALF;
                const bool odd = lane & 1;
                const float s_arg = odd ? -1.f : 2.f;       // sigmoid: exp(-x);  tanh: exp(2x)
                const float s_num = odd ? 1.f : 2.f;
#pragma unroll
                for (int half = 0; half < 2; ++half) {
                    const int lane0 = wq * 32 + half * 16;
                    float v[32];
                    tmem_ld_16x256b_x8(tmem + ((uint32_t)lane0 << 16) + bi * 128 + hh * 64, v);
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
                tc::fence_proxy_async_smem();
                tc::tc_fence_before();
                tc::mbar_arrive(&bar_zfull[bi]);
            }
            if (i == 0) dbg_stamp(q.dbg, 3);
            if (i + 1 < n_my) aux_stage(i + 1);          // acc1(i) complete => the aux MMAs of tile i have read the operand

            // ---------------- E2: (acc2 + bias) -> padded smem tile -> coalesced residual / skip pass ----------------
            // The residual / old-skip rows this thread will combine are fetched BEFORE waiting for GEMM2 (they do not
            // depend on it): the old loop issued 4 dependent load->store rounds per warp, ~2.4K cycles each under load.
            // warp ww owns rows ww + 8j (j = 0..15), lane owns channels (2 lane, 2 lane + 1).
            float2 res[16], sko[16];
            {
                const size_t base = row0 * 64;
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                    const int rr = ww + 8 * j;
                    res[j] = make_float2(0.f, 0.f); sko[j] = make_float2(0.f, 0.f);
                    if (rr < nlive) {
                        res[j] = __ldg(reinterpret_cast<const float2*>(p.Hin + base + (size_t)rr * 64) + lane);
                        if (!p.skip_init) sko[j] = *(reinterpret_cast<const float2*>(p.Skip + base + (size_t)rr * 64) + lane);
                    }
                }
            }
            ok &= pt_wait(&bar_acc2, i & 1, &timeout_s);
            tc::tc_fence_after();
            if (i == 0) dbg_stamp(q.dbg, 4);
            const bool more = i + 2 < n_my;
            if (more) x_load(i + 2);                     // in flight under the epilogue
            constexpr int SST = 129;
            float* S2 = Xb[bi];                          // z has been consumed: GEMM2(i) is complete
            {
                const int r = wq * 32 + lane;
#pragma unroll
                for (int cc = 0; cc < 2; ++cc) {
                    const int col0 = hh * 64 + cc * 32;
                    float v[32];
                    tc::tmem_ld32(tq + 256 + col0, v);
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
            tc::tc_fence_before();
            pt_worker_sync();
            {
                const size_t base = row0 * 64;
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                    const int rr = ww + 8 * j;
                    if (rr >= nlive) continue;
                    const float* sp = S2 + rr * SST + 4 * lane;
                    float2 ho, sk;
                    ho.x = (sp[0] + res[j].x) * CRK_SQRT_HALF;
                    ho.y = (sp[1] + res[j].y) * CRK_SQRT_HALF;
                    sk.x = sp[2]; sk.y = sp[3];
                    if (!p.skip_init) { sk.x += sko[j].x; sk.y += sko[j].y; }
                    reinterpret_cast<float2*>(p.Hout + base + (size_t)rr * 64)[lane] = ho;
                    reinterpret_cast<float2*>(p.Skip + base + (size_t)rr * 64)[lane] = sk;
                }
            }
            pt_worker_sync();                            // S2 fully read before X(i+2) / z(i+2) overwrite the buffer
            if (more) x_store(bi);
            if (i == 0) dbg_stamp(q.dbg, 5);
        }
        if (!ok) timeout_s = 1;
        dbg_stamp(q.dbg, 6);
    }
    tc::tc_fence_before();
    __syncthreads();
    if (timeout_s && threadIdx.x == 0) {
        int b, t0;
        tile_of(0, b, t0);
        p.Hout[((size_t)b * p.T + t0) * 64] = __int_as_float(0x7fc00000);      // poison: the test must fail
    }
    if (warp == 1) tc::tmem_dealloc<512>(tmem);
}

inline size_t resblock_fwd_pt_smem(int k, int dil, bool split) {
    const int rowsX = CRK_TC_TM + (k - 1) * dil;
    const int xh = 16 * tc::chunk_rows(rowsX) * 4;
    const int xbuf = (split ? 2 : 1) * xh > CRK_TC_TM * 129 ? (split ? 2 : 1) * xh : CRK_TC_TM * 129;
    return (size_t)(2 * xbuf + 2 * 16 * 129 * 4) * sizeof(float);
}
inline int device_sm_count() {
    static int n = 0;
    if (n == 0) {
        int dev = 0;
        if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    }
    return n;
}
inline bool resblock_fwd_pt_ok(const ResFwdTcParams& q, bool split) {
    const int halo = (q.p.k - 1) * q.p.dil;
    if (halo > 16 || q.p.k < 1) return false;                    // X tile <= 144 rows = 9 float4 per worker
    if (q.p.Ca > 0 && (q.KaPad > 64 || (q.KaPad & 7))) return false;   // aux operand: 64 + 64 tensor-memory columns
    return resblock_fwd_pt_smem(q.p.k, q.p.dil, split) <= 225 * 1024;
}

template <bool SPLIT>
inline cudaError_t launch_resblock_fwd_pt(const ResFwdTcParams& q, cudaStream_t s) {
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(k_resblock_fwd_pt<SPLIT>, cudaFuncAttributeMaxDynamicSharedMemorySize, 225 * 1024);
        if (e != cudaSuccess) return e;
        attr_set = true;
    }
    const int tiles = q.p.B * cdiv(q.p.T, CRK_TC_TM);
    const int grid = tiles < device_sm_count() ? tiles : device_sm_count();
    TimedLaunch tl(CRK_K_RESBLOCK_FWD, s, 2.0 * q.p.B * q.p.T * (64.0 * 128 * q.p.k + q.p.Ca * 128.0 + 64.0 * 128));
    ResFwdTcParams qq = q;
    qq.dbg = dbg_take(CRK_K_RESBLOCK_FWD);
    cudaError_t le = launch_pdl(k_resblock_fwd_pt<SPLIT>, dim3(grid), dim3(CRK_PT_THREADS), resblock_fwd_pt_smem(q.p.k, q.p.dil, SPLIT), s, qq);
    if (le != cudaSuccess) return le;
    return launch_check();
}

}  // namespace crk
