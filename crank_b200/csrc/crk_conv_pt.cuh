// crank-b200: PERSISTENT, warp-specialised channels-last conv / dgrad on tcgen05 (round 2).
//
// Same contract as k_conv_tc<.., CRK_CONV_FAST> (crk_conv_tc.cuh): the dgrad of the gated dilated conv
// (X = dg, K = 128, N = 64, k taps) and the 1x1 convs of the WaveNet stacks (crank/net/module/vqvae2.py:236-273),
// for operands that allow 128-bit accesses and N <= 64.  Round-1 profile of that kernel: stage 8.6K / MMA 17.1K /
// epilogue 7.3K cycles per 128-frame tile, strictly serial, tensor pipe 5 % active.  Here
//   * the A tile is staged per 64-CHANNEL SEGMENT into one of two shared-memory buffers: the MMAs of segment g run
//     while the workers stage segment g+2 (the next tile's), so staging, MMA and epilogue of consecutive tiles overlap;
//   * accumulators are double-buffered in tensor memory (2 x 128 columns);
//   * the epilogue reads tensor memory with the 16x256b fragment shape (a thread owns 2 adjacent channels of 4 rows:
//     every global access of a quad is one full 32 B sector) -- no shared-memory transposition tile, no block barrier.
//   grid = min(#tiles, #SMs);  320 threads: warp 0 TMA producer (weight blobs, 2-slot ring of hi|lo K = 64 slices),
//   warp 1 MMA issuer, warps 2..9 workers.
#pragma once
#include "crk_common.cuh"
#include "crk_conv_tc.cuh"
#include "crk_resblock_pt.cuh"
#include "crk_tc.cuh"

namespace crk {

struct PtARegs {
    float4 v[9];
};

template <bool SPLIT>
__global__ void __launch_bounds__(CRK_PT_THREADS, 1) k_conv_pt(const ConvTcParams q) {
    const ConvParams& p = q.p;
    extern __shared__ float4 crk_smem4[];
    float* smem = reinterpret_cast<float*>(crk_smem4);
    __shared__ uint64_t bar_rfull[2], bar_rfree[2];     // weight ring
    __shared__ uint64_t bar_afull[2];                   // A segment staged (256 worker arrivals)
    __shared__ uint64_t bar_adone[2];                   // the MMAs reading A buffer b have completed
    __shared__ uint64_t bar_acc[2];                     // all MMAs of the tile in acc[b] have completed
    __shared__ uint32_t tmem_base_s;
    __shared__ int timeout_s;

    constexpr int SEG = 16;                             // K chunks (of 4 channels) per segment
    const int tiles_per_utt = (p.T + CRK_TC_TM - 1) / CRK_TC_TM;
    const int ntiles = p.B * tiles_per_utt;
    const int n_my = (int)blockIdx.x < ntiles ? (ntiles - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;
    const int rowsX = CRK_TC_TM + (p.k - 1) * p.dil;
    const int csx = tc::chunk_rows(rowsX) * 4;          // floats per A chunk
    const int kch = q.Kpad >> 2;                        // chunks over the whole K
    const int nseg = (kch + SEG - 1) / SEG;             // 1 or 2
    const int csw = tc::chunk_rows(q.Npad) * 4;         // floats per B chunk
    const int whalf_tap = kch * csw;                    // floats of the hi half of one tap blob
    const int ahalf = SEG * csx;                        // floats of one A segment half
    const int abuf = (SPLIT ? 2 : 1) * ahalf;
    const int shalf = SEG * csw;                        // floats of one ring-slot half
    float* Ab[2] = {smem, smem + abuf};
    float* ring = smem + 2 * abuf;
    float* slot[2] = {ring, ring + (SPLIT ? 2 : 1) * shalf};
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int nsegs_total = n_my * nseg;

    if (threadIdx.x == 0) {
        for (int i = 0; i < 2; ++i) {
            tc::mbar_init(&bar_rfull[i], 1); tc::mbar_init(&bar_rfree[i], 1);
            tc::mbar_init(&bar_afull[i], CRK_PT_WORKERS); tc::mbar_init(&bar_adone[i], 1);
            tc::mbar_init(&bar_acc[i], 1);
        }
        tc::fence_mbar_init();
        timeout_s = 0;
    }
    if (warp == 1) tc::tmem_alloc<256>(&tmem_base_s);
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
    auto seg_chunks = [&](int sg) { return (kch - sg * SEG) < SEG ? (kch - sg * SEG) : SEG; };

    if (warp == 0) {
        // ===================== TMA producer: step = (global segment g, tap j) =====================
        if (lane == 0) {
            int step = 0;
            for (int g = 0; g < nsegs_total; ++g) {
                const int sg = g % nseg;
                const int nch = seg_chunks(sg);
                for (int j = 0; j < p.k; ++j, ++step) {
                    const int sl = step & 1;
                    if (step >= 2) ok &= pt_wait(&bar_rfree[sl], ((step - 2) >> 1) & 1, &timeout_s);
                    const float* blob = q.Wtc + (size_t)j * 2 * whalf_tap + (size_t)sg * SEG * csw;
                    const uint32_t bytes = (uint32_t)(nch * csw) * 4u;
                    tc::mbar_arrive_expect_tx(&bar_rfull[sl], SPLIT ? 2u * bytes : bytes);
                    tc::bulk_g2s(slot[sl], blob, bytes, &bar_rfull[sl]);
                    if (SPLIT) tc::bulk_g2s(slot[sl] + shalf, blob + whalf_tap, bytes, &bar_rfull[sl]);
                }
            }
            if (!ok) timeout_s = 1;
        }
        __syncwarp();
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        const uint32_t idesc = tc::make_idesc_tf32(128, q.Npad, 0, 0);
        const bool leader = tc::elect_one();
        int step = 0;
        uint32_t accf = 0;
        for (int g = 0; g < nsegs_total; ++g) {
            const int i = g / nseg, sg = g - i * nseg;
            const int ab = g & 1;
            const int nch = seg_chunks(sg);
            if (sg == 0) accf = 0;
            ok &= pt_wait(&bar_afull[ab], (g >> 1) & 1, &timeout_s);
            const uint32_t ah = tc::smem_u32(Ab[ab]), al = tc::smem_u32(Ab[ab] + ahalf);
            const uint32_t d = tmem + (i & 1) * 128;
            for (int j = 0; j < p.k; ++j, ++step) {
                const int sl = step & 1;
                ok &= pt_wait(&bar_rfull[sl], (step >> 1) & 1, &timeout_s);
                tc::tc_fence_after();
                const uint32_t bh = tc::smem_u32(slot[sl]), bl = tc::smem_u32(slot[sl] + shalf);
                if (SPLIT) {
                    pt_issue_ss(d, al, csx * 4, j * p.dil, bh, csw * 4, nch * 4, idesc, accf, leader);    // lo * hi
                    pt_issue_ss(d, ah, csx * 4, j * p.dil, bl, csw * 4, nch * 4, idesc, accf, leader);    // hi * lo
                }
                pt_issue_ss(d, ah, csx * 4, j * p.dil, bh, csw * 4, nch * 4, idesc, accf, leader);        // hi * hi
                if (leader) tc::umma_commit(&bar_rfree[sl]);
                __syncwarp();
            }
            if (leader) {
                tc::umma_commit(&bar_adone[ab]);
                if (sg == nseg - 1) tc::umma_commit(&bar_acc[i & 1]);
            }
            __syncwarp();
        }
        if (!ok) timeout_s = 1;
        __syncwarp();
    } else {
        // ===================== workers =====================
        const int wt = threadIdx.x - 64;                 // 0..255
        const int ww = wt >> 5;
        const int wq = warp & 3;                         // TMEM lane quarter
        const int hh = ww >> 2;                          // which half of the 8-column blocks
        PtARegs AR;

        // segment g = (tile i, K segment sg): rows x nch chunks, 9 float4 per worker at most (rows <= 144)
        auto a_load = [&](int g) {
            const int i = g / nseg, sg = g - i * nseg;
            int b, t0;
            tile_of(i, b, t0);
            const int tstart = t0 - p.padl;
            const int nch = seg_chunks(sg);
            const int total = rowsX * SEG;
#pragma unroll
            for (int u = 0; u < 9; ++u) {
                const int idx = wt + u * CRK_PT_WORKERS;
                AR.v[u] = make_float4(0.f, 0.f, 0.f, 0.f);
                if (idx < total) {
                    const int r = idx >> 4, c4 = idx & 15;
                    const int tt = tstart + r;
                    const int c = (sg * SEG + c4) * 4;
                    if (c4 < nch && tt >= 0 && tt < p.T && c < p.Cin)
                        AR.v[u] = __ldg(reinterpret_cast<const float4*>(p.X + ((size_t)b * p.T + tt) * p.ldx) + sg * SEG + c4);
                }
            }
        };
        auto a_store = [&](int g) {
            const int ab = g & 1;
            const int sg = g % nseg;
            const int nch = seg_chunks(sg);
            float* hi = Ab[ab];
            float* lo = Ab[ab] + ahalf;
            const int total = rowsX * SEG;
#pragma unroll
            for (int u = 0; u < 9; ++u) {
                const int idx = wt + u * CRK_PT_WORKERS;
                if (idx >= total) continue;
                const int r = idx >> 4, c4 = idx & 15;
                if (c4 >= nch) continue;
                const int off = c4 * csx + r * 4;
                float4 x;
                x.x = apply_act(AR.v[u].x * p.pro_scale, p.pro_act, p.pro_slope);
                x.y = apply_act(AR.v[u].y * p.pro_scale, p.pro_act, p.pro_slope);
                x.z = apply_act(AR.v[u].z * p.pro_scale, p.pro_act, p.pro_slope);
                x.w = apply_act(AR.v[u].w * p.pro_scale, p.pro_act, p.pro_slope);
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
            tc::mbar_arrive(&bar_afull[ab]);
        };

        dbg_stamp(q.dbg, 0);
        if (nsegs_total > 0) { a_load(0); a_store(0); }
        if (nsegs_total > 1) { a_load(1); a_store(1); }
        dbg_stamp(q.dbg, 1);

        for (int i = 0; i < n_my; ++i) {
            int b, t0;
            tile_of(i, b, t0);
            const int nlive = min(CRK_TC_TM, p.T - t0);
            const size_t row0 = (size_t)b * p.T + t0;
            // all but the last segment of this tile: as soon as its MMAs have completed, its buffer takes segment g + 2
            for (int sg = 0; sg + 1 < nseg; ++sg) {
                const int g = i * nseg + sg;
                ok &= pt_wait(&bar_adone[g & 1], (g >> 1) & 1, &timeout_s);
                if (g + 2 < nsegs_total) { a_load(g + 2); a_store(g + 2); }
            }
            const int gl = i * nseg + nseg - 1;          // last segment: its completion == the tile's accumulator
            // side inputs of the epilogue (residual gradient, dropout multiplier) fetched BEFORE the accumulator wait:
            // slot s = (cg, half, bb, hr) of this thread; one round trip instead of one per 8-column block
            const int ncol64 = (q.Npad + 63) >> 6;       // 64-column groups (1 for N <= 64)
            float2 pmul[16], pres[16];
            {
#pragma unroll
                for (int sidx = 0; sidx < 16; ++sidx) {
                    const int half = sidx >> 3, bb = (sidx >> 1) & 3, hr = sidx & 1;
                    const int r = wq * 32 + half * 16 + (lane >> 2) + 8 * hr;
                    const int col = hh * 32 + 8 * bb + 2 * (lane & 3);
                    pmul[sidx] = make_float2(1.f, 1.f); pres[sidx] = make_float2(0.f, 0.f);
                    if (r < nlive && col < p.Cout) {
                        const size_t row = row0 + r;
                        if (p.mul_src) pmul[sidx] = __ldg(reinterpret_cast<const float2*>(p.mul_src + row * p.ldmul + col));
                        if (p.R) pres[sidx] = __ldg(reinterpret_cast<const float2*>(p.R + row * p.ldr + col));
                    }
                }
            }
            ok &= pt_wait(&bar_acc[i & 1], (i >> 1) & 1, &timeout_s);
            tc::tc_fence_after();
            if (i == 0) dbg_stamp(q.dbg, 2);
            const bool more = gl + 2 < nsegs_total;
            if (more) a_load(gl + 2);                    // in flight under the epilogue
            if (i == 0) dbg_stamp(q.dbg, 5);

            // ---------------- epilogue: 16x256b fragments, this warp's 32 of each group's 64 columns ----------------
            for (int cg = 0; cg < ncol64; ++cg) {
#pragma unroll
                for (int half = 0; half < 2; ++half) {
                    const int lane0 = wq * 32 + half * 16;
                    float v[16];
                    tmem_ld_16x256b_x4(tmem + ((uint32_t)lane0 << 16) + (i & 1) * 128 + cg * 64 + hh * 32, v);
#pragma unroll
                    for (int bb = 0; bb < 4; ++bb) {
                        const int col = cg * 64 + hh * 32 + 8 * bb + 2 * (lane & 3);
                        if (col >= p.Cout) continue;
                        float2 bia = make_float2(0.f, 0.f);
                        if (p.bias) bia = __ldg(reinterpret_cast<const float2*>(p.bias + col));
#pragma unroll
                        for (int hr = 0; hr < 2; ++hr) {
                            const int r = lane0 + (lane >> 2) + 8 * hr;
                            if (r >= nlive) continue;
                            const size_t row = row0 + r;
                            const int sidx = half * 8 + bb * 2 + hr;
                            float2 mulv = pmul[sidx], rv = pres[sidx];
                            if (cg > 0) {                                   // (N > 64: not prefetched)
                                mulv = p.mul_src ? __ldg(reinterpret_cast<const float2*>(p.mul_src + row * p.ldmul + col)) : make_float2(1.f, 1.f);
                                rv = p.R ? __ldg(reinterpret_cast<const float2*>(p.R + row * p.ldr + col)) : make_float2(0.f, 0.f);
                            }
                            float2 dv = make_float2(1.f, 1.f), oldv = make_float2(0.f, 0.f);
                            if (p.dact_src) dv = __ldg(reinterpret_cast<const float2*>(p.dact_src + row * p.lddact + col));
                            if (p.accumulate) oldv = *reinterpret_cast<const float2*>(p.Y + row * p.ldy + col);
                            float y[2] = {v[4 * bb + 2 * hr], v[4 * bb + 2 * hr + 1]};
                            const float b2[2] = {bia.x, bia.y}, m2[2] = {mulv.x, mulv.y}, r2[2] = {rv.x, rv.y};
                            const float d2[2] = {dv.x, dv.y}, o2[2] = {oldv.x, oldv.y};
#pragma unroll
                            for (int e = 0; e < 2; ++e) {
                                float t = y[e];                               // same operation order as k_conv / k_conv_tc
                                if (p.bias) t += b2[e];
                                t = apply_act(t, p.epi_act, p.epi_slope);
                                t *= m2[e];
                                t += p.rscale * r2[e];
                                if (p.dact_src) t *= act_grad(d2[e], p.dact_mode, p.dact_slope);
                                t *= p.out_scale;
                                t += o2[e];
                                y[e] = t;
                            }
                            *reinterpret_cast<float2*>(p.Y + row * p.ldy + col) = make_float2(y[0], y[1]);
                        }
                    }
                }
            }
            tc::tc_fence_before();
            if (i == 0) dbg_stamp(q.dbg, 6);
            if (more) a_store(gl + 2);
            if (i == 0) dbg_stamp(q.dbg, 3);
        }
        if (!ok) timeout_s = 1;
        dbg_stamp(q.dbg, 4);
    }
    tc::tc_fence_before();
    __syncthreads();
    if (timeout_s && threadIdx.x == 0) {
        int b, t0;
        tile_of(0, b, t0);
        p.Y[((size_t)b * p.T + t0) * p.ldy] = __int_as_float(0x7fc00000);      // poison: the test must fail
    }
    if (warp == 1) tc::tmem_dealloc<256>(tmem);
}

inline size_t conv_pt_smem(const ConvTcParams& q, bool split) {
    const int rowsX = CRK_TC_TM + (q.p.k - 1) * q.p.dil;
    const size_t a = (size_t)16 * tc::chunk_rows(rowsX) * 4;
    const size_t sl = (size_t)16 * tc::chunk_rows(q.Npad) * 4;
    return ((split ? 4 : 2) * a + (split ? 4 : 2) * sl) * sizeof(float);
}
// FAST contract (128-bit operands, no input multiplier) + N <= 64 (ring slot of hi|lo K = 64 slices) + halo <= 16
inline bool conv_pt_ok(const ConvTcParams& q, bool split) {
    return q.Wtc != nullptr && q.Npad >= 16 && q.Npad <= 64 && (q.Npad % 16) == 0 && q.Kpad >= 8 && q.Kpad <= 128 &&
           (q.Kpad % 8) == 0 && (q.p.k - 1) * q.p.dil <= 16 && (q.p.Cout & 1) == 0 && conv_pt_smem(q, split) <= 225 * 1024;
}

template <bool SPLIT>
inline cudaError_t launch_conv_pt(const ConvTcParams& q, cudaStream_t s) {
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(k_conv_pt<SPLIT>, cudaFuncAttributeMaxDynamicSharedMemorySize, 225 * 1024);
        if (e != cudaSuccess) return e;
        attr_set = true;
    }
    const int tiles = q.p.B * cdiv(q.p.T, CRK_TC_TM);
    const int grid = tiles < device_sm_count() ? tiles : device_sm_count();
    TimedLaunch tl(CRK_K_CONV, s, 2.0 * q.p.B * q.p.T * q.p.Cin * q.p.Cout * q.p.k);
    ConvTcParams qq = q;
    qq.dbg = dbg_take(CRK_K_CONV);
    cudaError_t le = launch_pdl(k_conv_pt<SPLIT>, dim3(grid), dim3(CRK_PT_THREADS), conv_pt_smem(q, SPLIT), s, qq);
    if (le != cudaSuccess) return le;
    return launch_check();
}

}  // namespace crk
