// crank-b200: convolution weight gradients on tcgen05 (K-major TF32 / 3xTF32, reduction over frames).
//
//   dW_j[ci][co] = sum_frames X[f + j*dil - padl][ci] * G[f][co]
// as  D_j^T[co (M=128)][ci (N)] = G^T . X_j  with the frame index as the UMMA K dimension.  Both
// operands are staged TRANSPOSED into chunk-major tiles ([4-frame chunk][channel row][4 frames]) so
// that they are K-major (MN-major tf32 operands are not usable with the no-swizzle layout -- measured,
// tests/test_gpu_tc.py).  A tap shift is a shift along K, which a descriptor can only express in
// multiples of 4 frames, so X^T is re-staged per tap (through a 2-slot ring, overlapped with the
// previous tap's MMAs) while G^T is staged once per 64-frame tile.  All k taps accumulate in TMEM
// (k x N columns) across every tile of the CTA's frame chunk; one epilogue per CTA writes the
// per-chunk partial in the packed [tap][ci][co] layout k_reduce expects.
#pragma once
#include "crk_common.cuh"
#include "crk_conv.cuh"
#include "crk_resblock_tc.cuh"
#include "crk_tc.cuh"

namespace crk {

#define CRK_WG_TF 64   // frames per tile (UMMA K = 64 per tile and tap)

struct WgradTcParams {
    int gbuf;          // k_wgrad_tc: number of G^T buffers (1 or 2)
    long long stage_floats;   // dynamic shared memory of the launch in floats (the epilogue stages the partial block there)
    WgradParams p;     // same contract as the fp32 kernel (p.part = [nchunk][k][Rows][TN])
    int TN;            // packed columns of G / dW (32*cpt)
    int Npad;          // UMMA N = round_up(Cin, 16)
    int dbg;
    int nslot;         // X^T ring depth (2..4): hides the MMA completion latency of the per-tap handshake
    int bias;          // != 0: also write the per-chunk column sums of G (bias-gradient partial) behind the
                       // chunk's [k][Rows][TN] block, i.e. what k_colsum would have produced in a second pass over G
    int tpart;         // != 0: the chunk's weight block is written TRANSPOSED, [k][TN][Rows]: a thread owns one co and 32
                       // consecutive ci (the tensor-memory layout), so it stores 8 x 16 B instead of 32 x 4 B -- the epilogue
                       // was suspected to be bound by the number of store instructions (128 B per warp instruction).
                       // k_reduce_t sums the partials in this layout and writes dW in the [k][Rows][TN] layout.
                       // Measured (opt-in, crk_debug_opt_enable 8): no gain -- k = 5 epilogue 15 K -> 20 K cycles, k = 1 6.1 K -> 5.5 K:
                       // 16 B per lane at a 256 B stride writes half-filled sectors; the epilogue is bound by L2 sector writes.
};

// Transposed staging of src[frame t0+shift .. +64)[0..ncols):  elem(frame f, channel c) -> (f>>2)*cs + c*4 + (f&3).
// Split in two phases so that the global loads can be issued BEFORE waiting on the mbarrier that
// guards the destination buffer (their latency overlaps the wait / the running MMAs):
//   wg_load  : NB float4 per thread into registers (consecutive lanes = consecutive frames)
//   wg_store : prologue, tf32 hi/lo split, conflict-free scalar scatter into the chunk-major tile
template <int NB>
struct WgRegs {
    float4 v[NB];
    float4 m[NB];
};

// VECONLY: host-checked (16 B aligned rows, ncols % 4 == 0) -> the scalar tail path is compiled out
template <int NB, bool VECONLY = false, int TW = 256>
__device__ __forceinline__ void wg_load(WgRegs<NB>& R, int nrows_pad, const float* __restrict__ src, int ld, int ncols,
                                        int b, int T, int t0, int shift, const float* __restrict__ mul, int ldmul) {
    const int total = CRK_WG_TF * (nrows_pad >> 2);
    const bool vec = VECONLY || (((ld & 3) == 0) && ((reinterpret_cast<uintptr_t>(src) & 15) == 0) &&
                     (mul == nullptr || (((ldmul & 3) == 0) && ((reinterpret_cast<uintptr_t>(mul) & 15) == 0))));
#pragma unroll
    for (int u = 0; u < NB; ++u) {
        const int idx = threadIdx.x + u * TW;
        R.v[u] = make_float4(0.f, 0.f, 0.f, 0.f);
        R.m[u] = make_float4(1.f, 1.f, 1.f, 1.f);
        if (idx < total) {
            const int c4 = idx >> 6, f = idx & 63;
            const int c = c4 * 4;
            const int tg = t0 + f;
            const int tt = tg + shift;
            if (tg < T && tt >= 0 && tt < T && c < ncols) {
                const size_t row = (size_t)b * T + tt;
                if (VECONLY || (vec && c + 3 < ncols)) {
                    R.v[u] = __ldg(reinterpret_cast<const float4*>(src + row * ld + c));
                    if (mul) R.m[u] = __ldg(reinterpret_cast<const float4*>(mul + row * ldmul + c));
                } else if (!VECONLY) {
                    float t4[4] = {0.f, 0.f, 0.f, 0.f}, m4[4] = {1.f, 1.f, 1.f, 1.f};
#pragma unroll
                    for (int e = 0; e < 4; ++e)
                        if (c + e < ncols) {
                            t4[e] = __ldg(src + row * ld + c + e);
                            if (mul) m4[e] = __ldg(mul + row * ldmul + c + e);
                        }
                    R.v[u] = make_float4(t4[0], t4[1], t4[2], t4[3]);
                    R.m[u] = make_float4(m4[0], m4[1], m4[2], m4[3]);
                }
            }
        }
    }
}

template <bool SPLIT, int NB, int TW = 256>
__device__ __forceinline__ void wg_store(const WgRegs<NB>& R, float* hi, float* lo, int cs_floats, int nrows_pad,
                                         int pro_act, float pro_slope, float pro_scale) {
    const int total = CRK_WG_TF * (nrows_pad >> 2);
#pragma unroll
    for (int u = 0; u < NB; ++u) {
        const int idx = threadIdx.x + u * TW;
        if (idx >= total) continue;
        const int c4 = idx >> 6, f = idx & 63;
        const int off = (f >> 2) * cs_floats + c4 * 16 + (f & 3);
        const float x[4] = {apply_act(R.v[u].x * pro_scale, pro_act, pro_slope) * R.m[u].x,
                            apply_act(R.v[u].y * pro_scale, pro_act, pro_slope) * R.m[u].y,
                            apply_act(R.v[u].z * pro_scale, pro_act, pro_slope) * R.m[u].z,
                            apply_act(R.v[u].w * pro_scale, pro_act, pro_slope) * R.m[u].w};
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            if (SPLIT) {
                float h, l;
                tc::split_tf32(x[e], h, l);
                hi[off + e * 4] = h;
                lo[off + e * 4] = l;
            } else {
                hi[off + e * 4] = x[e];
            }
        }
    }
}

// TW threads: 512 in the 3xTF32 mode since round 2 (see k_wgrad_tc_raw: the per-tile stores are latency-bound with 8 warps)
template <bool SPLIT, int NX, bool VEC, int TW>
__global__ void __launch_bounds__(TW, (SPLIT || NX > 4 || TW > 256) ? 1 : 2) k_wgrad_tc(const WgradTcParams q) {
    constexpr int SC = TW / 256;
    constexpr int NG = 8 / SC, NXT = NX / SC;              // G^T / X^T float4 per thread
    const WgradParams& p = q.p;
    extern __shared__ float4 crk_smem4[];
    float* smem = reinterpret_cast<float*>(crk_smem4);
    __shared__ uint64_t bar_slot[4];
    __shared__ uint64_t bar_tile[2];                       // all MMAs of the tile in G^T buffer b have completed
    __shared__ uint32_t tmem_base_s;
    __shared__ int timeout_s;

    constexpr int KCH = CRK_WG_TF / 4;                     // 16 frame chunks per tile
    constexpr int CSG = 129 * 4;                           // G^T: 128 rows -> 129
    const int csx = tc::chunk_rows(q.Npad) * 4;
    // G^T buffers: q.gbuf = 2 (round 2, when shared memory allows): the tile n+1 stores run under the MMAs of tile n -- with
    // one buffer every tile waited for the previous tile's MMAs before its first store (k = 1: the whole tile was serial)
    const int NGB = q.gbuf;
    const int gstride = (SPLIT ? 2 : 1) * KCH * CSG;       // floats per G^T buffer (hi | lo)
    float* Gh = smem;
    float* Gl = Gh + KCH * CSG;
    float* ring = Gh + NGB * gstride;
    const int xhalf = KCH * csx;
    const int slot_floats = (SPLIT ? 2 : 1) * xhalf;
    const int NS = q.nslot;
    auto slot_hi = [&](int sl) -> float* { return ring + sl * slot_floats; };
    auto slot_lo = [&](int sl) -> float* { return ring + sl * slot_floats + xhalf; };
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    const int tiles_per_utt = (p.T + CRK_WG_TF - 1) / CRK_WG_TF;
    const int ntiles = p.B * tiles_per_utt;
    const int tile_beg = blockIdx.x * p.tiles_per_chunk;
    const int tile_end = min(ntiles, tile_beg + p.tiles_per_chunk);

    if (threadIdx.x == 0) {
        for (int i = 0; i < 4; ++i) tc::mbar_init(&bar_slot[i], 1);
        tc::mbar_init(&bar_tile[0], 1); tc::mbar_init(&bar_tile[1], 1);
        tc::fence_mbar_init();
        timeout_s = 0;
    }
    if (warp == 1) tc::tmem_alloc<512>(&tmem_base_s);
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    const uint32_t tmem = tmem_base_s;
    const uint32_t idesc = tc::make_idesc_tf32(128, q.Npad, 0, 0);
    const uint32_t gh_s = tc::smem_u32(Gh), gl_s = tc::smem_u32(Gl);
    bool ok = true;
    int step = 0;            // global (tile, tap) step counter -> ring slot + mbarrier phase
    int ntile_done = 0;

    // G^T tile = 64 frames x 128 channels = 8 float4 per thread; X^T tile = NX float4 per thread (4 when
    // Npad <= 64).  The X loads of tap j+1 are issued BEFORE the MMAs of tap j are even launched and land
    // while the previous tap is stored / synchronised; G of the next tile is prefetched during the last tap.
    WgRegs<NG> RG;
    WgRegs<NXT> RX;
    // bias gradient: every G element passes through this thread's registers exactly once (wg_load of the
    // tile), always the same (frame f = tid & 63, channel quad tid/64 + 4u) slot -> per-thread running
    // sums, combined once per CTA in fixed order (deterministic); no second pass over G
    float4 bsum[NG];
#pragma unroll
    for (int u = 0; u < NG; ++u) bsum[u] = make_float4(0.f, 0.f, 0.f, 0.f);
    pdl_trigger();                                      // only now: this CTA already owns its TMEM columns (see crk_common.cuh)
    pdl_wait();
    auto load_x = [&](int tile, int j) {
        const int bb = tile / tiles_per_utt;
        const int tt0 = (tile - bb * tiles_per_utt) * CRK_WG_TF;
        wg_load<NXT, VEC, TW>(RX, q.Npad, p.X, p.ldx, p.Cin, bb, p.T, tt0, j * p.dil - p.padl, p.xmul, p.ldxmul);
    };
    auto store_x = [&](int slot) {
        wg_store<SPLIT, NXT, TW>(RX, slot_hi(slot), slot_lo(slot), csx, q.Npad, p.pro_act, p.pro_slope, p.pro_scale);
    };
    auto load_g = [&](int tile) {
        const int bb = tile / tiles_per_utt;
        const int tt0 = (tile - bb * tiles_per_utt) * CRK_WG_TF;
        wg_load<NG, VEC, TW>(RG, 128, p.G, p.ldg, p.N, bb, p.T, tt0, 0, nullptr, 0);
    };

    if (tile_beg < tile_end) { load_g(tile_beg); load_x(tile_beg, 0); }
    for (int tile = tile_beg; tile < tile_end; ++tile) {
        if (tile == tile_beg) dbg_stamp(q.dbg, 0);
        // G and first-tap X of this tile are already in registers (prefetched); wait for the MMAs of the tile that used this
        // G^T buffer last, and for the MMAs that read the ring slot we are about to overwrite
        const int gb = ntile_done % NGB, guse = ntile_done / NGB;
        if (guse > 0) { ok &= tc::mbar_wait(&bar_tile[gb], (guse - 1) & 1); tc::tc_fence_after(); }
        if (step >= NS) { ok &= tc::mbar_wait(&bar_slot[step % NS], ((step - NS) / NS) & 1); tc::tc_fence_after(); }
        if (tile == tile_beg) dbg_stamp(q.dbg, 1);
        wg_store<SPLIT, NG, TW>(RG, Gh + gb * gstride, Gl + gb * gstride, CSG, 128, CRK_ACT_NONE, 0.f, 1.f);
        if (q.bias) {
#pragma unroll
            for (int u = 0; u < NG; ++u) {
                bsum[u].x += RG.v[u].x; bsum[u].y += RG.v[u].y; bsum[u].z += RG.v[u].z; bsum[u].w += RG.v[u].w;
            }
        }
        store_x(step % NS);
        if (p.k > 1) load_x(tile, 1);                          // next tap's loads in flight during the sync + MMAs
        else if (tile + 1 < tile_end) { load_g(tile + 1); load_x(tile + 1, 0); }
        tc::fence_proxy_async_smem();
        tc::tc_fence_before();
        __syncthreads();
        tc::tc_fence_after();
        if (tile == tile_beg) dbg_stamp(q.dbg, 2);
        for (int j = 0; j < p.k; ++j, ++step) {
            if (warp == 1) {                               // warp-collective issue, one elected lane (tc_issue_kmajor_w)
                uint32_t acc = ntile_done > 0 ? 1u : 0u;
                // A = G^T (M = 128 rows = co), B = X_j^T (N rows = ci), K = 64 frames
                tc_issue_kmajor_w<SPLIT>(tmem + j * q.Npad, gh_s + gb * gstride * 4, gl_s + gb * gstride * 4, CSG * 4, 0, tc::smem_u32(slot_hi(step % NS)),
                                         tc::smem_u32(slot_lo(step % NS)), csx * 4, CRK_WG_TF, idesc, acc);
                if (tc::elect_one()) {
                    tc::umma_commit(&bar_slot[step % NS]);
                    if (j == p.k - 1) tc::umma_commit(&bar_tile[gb]);
                }
                __syncwarp();
            }
            if (j + 1 < p.k) {
                const int nstep = step + 1;
                // slot (nstep % NS) was last used by step nstep-NS (this tile or the previous one): with
                // NS slots the MMAs of the last NS-1 taps may still be in flight while we store
                if (nstep >= NS) ok &= tc::mbar_wait(&bar_slot[nstep % NS], ((nstep - NS) / NS) & 1);
                store_x(nstep % NS);                       // tap j+1 (loaded one tap ago)
                if (j + 2 < p.k) load_x(tile, j + 2);      // prefetch the tap after
                else if (tile + 1 < tile_end) { load_g(tile + 1); load_x(tile + 1, 0); }
                tc::fence_proxy_async_smem();
                __syncthreads();
            }
        }
        if (tile == tile_beg) dbg_stamp(q.dbg, 3);
        ++ntile_done;
    }
    dbg_stamp(q.dbg, 4);
    if (ntile_done > 0) ok &= tc::mbar_wait(&bar_tile[(ntile_done - 1) % NGB], ((ntile_done - 1) / NGB) & 1);   // (MMAs complete in order)
    tc::tc_fence_after();
    if (!ok) timeout_s = 1;
    __syncthreads();
    dbg_stamp(q.dbg, 5);

    // ---- epilogue: D_j^T[co][ci] -> part[chunk][j][ci][co] ----
    const int co = (warp & 3) * 32 + lane;
    const uint32_t tlane = tmem + ((uint32_t)((warp & 3) * 32) << 16);
    float* gout = p.part + (size_t)blockIdx.x * p.part_stride;
    // the chunk's whole partial block is assembled in shared memory (every operand buffer is free now) and leaves through
    // ONE TMA bulk store; direct stores only when the block does not fit / is not 16 B granular
    const bool staged = q.stage_floats >= p.part_stride + 512 && (p.part_stride & 3) == 0 && (reinterpret_cast<uintptr_t>(gout) & 15) == 0;
    float* out = staged ? smem : gout;
    const int nblk = (q.Npad + 31) >> 5;
    const float poison = __int_as_float(0x7fc00000);
    // (every element of the [k][Rows][TN] (+ [TN]) block is written below: no clearing needed)
    for (int pair = warp >> 2; pair < p.k * nblk; pair += TW / 128) {          // (tap j, 32-column block) pairs over the warps
        const int j = pair / nblk, blk = pair - j * nblk;
        float v[32];
        if (ntile_done > 0) tc::tmem_ld32(tlane + j * q.Npad + blk * 32, v);
        if (co >= q.TN) continue;
        if (q.tpart) {
            float* dst = out + ((size_t)j * q.TN + co) * p.Rows + blk * 32;
#pragma unroll
            for (int i = 0; i < 32; i += 4)
                if (blk * 32 + i < p.Rows) {
                    float4 w4 = ntile_done > 0 ? make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]) : make_float4(0.f, 0.f, 0.f, 0.f);
                    if (timeout_s) w4 = make_float4(poison, poison, poison, poison);
                    *reinterpret_cast<float4*>(dst + i) = w4;
                }
            continue;
        }
#pragma unroll
        for (int i = 0; i < 32; ++i) {
            const int ci = blk * 32 + i;
            if (ci < p.Rows)
                out[((size_t)j * p.Rows + ci) * q.TN + co] = timeout_s ? poison : (ntile_done > 0 ? v[i] : 0.f);
        }
    }
    if (q.bias) {
        // frames live on the 64 threads that share tid/64: warp-shuffle sum over 32 frames, then the two
        // warps of a pair through shared memory (all MMAs have completed: the G^T buffer is free)
        float* red = staged ? smem + p.part_stride : smem;         // (the staged block occupies [0, part_stride))                                 // [8 warps][32]
#pragma unroll
        for (int u = 0; u < NG; ++u) {
            float c4[4] = {bsum[u].x, bsum[u].y, bsum[u].z, bsum[u].w};
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                float sv = c4[e];
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) sv += __shfl_xor_sync(0xffffffffu, sv, o);
                if (lane == 0) red[warp * 32 + u * 4 + e] = sv;
            }
        }
        __syncthreads();
        if (threadIdx.x < q.TN) {
            const int c = threadIdx.x;                     // G column: quad c/4 = g + (TW/64) u (g = tid/64 of its owners)
            constexpr int GQ = TW / 64;
            const int quad = c >> 2, g = quad & (GQ - 1), u = quad / GQ;
            const float sv = red[(2 * g) * 32 + u * 4 + (c & 3)] + red[(2 * g + 1) * 32 + u * 4 + (c & 3)];
            out[(size_t)p.k * p.Rows * q.TN + c] = timeout_s ? poison : sv;
        }
    }
    if (staged) {
        tc::fence_proxy_async_smem();
        __syncthreads();
        if (threadIdx.x == 0) {
            tc::bulk_s2g(gout, smem, (uint32_t)(p.part_stride * sizeof(float)));
            tc::bulk_commit_wait_all();
        }
    }
    tc::tc_fence_before();
    __syncthreads();
    dbg_stamp(q.dbg, 6);
    if (warp == 1) tc::tmem_dealloc<512>(tmem);
}

__host__ __device__ inline size_t wgrad_raw_half(int Npad, int k, int dil) { return (size_t)(CRK_WG_TF + (k - 1) * dil) * (Npad + 4); }

// ---- raw-tile variant (k > 1) --------------------------------------------------------------------
// The k tap tiles of one 64-frame tile are the SAME (64 + halo) input rows shifted by j*dil frames.
// k_wgrad_tc fetches each of them from global memory (k latency-bound round trips per tile, one per
// tap iteration); here the rows are fetched ONCE per tile (register-prefetched a whole tile ahead),
// parked row-major in shared memory with the prologue applied, and every tap's transposed hi/lo
// operand tile is produced from that copy (shared -> shared).
//   xraw[r][c]  (row stride Npad + 4 floats: the 128-bit reads of 8 consecutive rows hit 8 distinct
//   bank groups), r <-> input time t0 - padl + r, zero outside [0, T).
// C4L: log2(c4n) when the channel-quad count is a power of two (the index split becomes a shift; the runtime
// division cost ~1.8K cycles per tile in the load-issue and store phases each), -1: generic
template <int NR, bool VECONLY, int C4L, int TW = 256>
__device__ __forceinline__ void wg_load_raw(WgRegs<NR>& R, int c4n, int rows, const float* __restrict__ src, int ld,
                                            int ncols, int b, int T, int tstart, const float* __restrict__ mul, int ldmul) {
    const int total = rows * c4n;
    const bool vec = VECONLY || (((ld & 3) == 0) && ((reinterpret_cast<uintptr_t>(src) & 15) == 0) &&
                     (mul == nullptr || (((ldmul & 3) == 0) && ((reinterpret_cast<uintptr_t>(mul) & 15) == 0))));
#pragma unroll
    for (int u = 0; u < NR; ++u) {
        const int idx = threadIdx.x + u * TW;
        R.v[u] = make_float4(0.f, 0.f, 0.f, 0.f);
        R.m[u] = make_float4(1.f, 1.f, 1.f, 1.f);
        if (idx < total) {
            const int r = C4L >= 0 ? (idx >> (C4L >= 0 ? C4L : 0)) : idx / c4n, c4 = idx - r * c4n;
            const int c = c4 * 4;
            const int tt = tstart + r;
            if (tt >= 0 && tt < T && c < ncols) {
                const size_t row = (size_t)b * T + tt;
                if (VECONLY || (vec && c + 3 < ncols)) {
                    R.v[u] = __ldg(reinterpret_cast<const float4*>(src + row * ld + c));
                    if (mul) R.m[u] = __ldg(reinterpret_cast<const float4*>(mul + row * ldmul + c));
                } else if (!VECONLY) {
                    float t4[4] = {0.f, 0.f, 0.f, 0.f}, m4[4] = {1.f, 1.f, 1.f, 1.f};
#pragma unroll
                    for (int e = 0; e < 4; ++e)
                        if (c + e < ncols) {
                            t4[e] = __ldg(src + row * ld + c + e);
                            if (mul) m4[e] = __ldg(mul + row * ldmul + c + e);
                        }
                    R.v[u] = make_float4(t4[0], t4[1], t4[2], t4[3]);
                    R.m[u] = make_float4(m4[0], m4[1], m4[2], m4[3]);
                }
            }
        }
    }
}

// SPLIT: the hi / lo parts are produced HERE, once per input element, into two raw copies (xraw | xraw_lo): round 1 re-split
// the same element for every one of the k taps that read it (ncu, round 2: the split arithmetic + its stores were ~40 % of
// the kernel's issue slots for k = 5)
template <bool SPLIT, int NR, int C4L, int TW = 256>
__device__ __forceinline__ void wg_store_raw(const WgRegs<NR>& R, float* xraw, float* xraw_lo, int stride, int c4n, int rows,
                                             int pro_act, float pro_slope, float pro_scale) {
    const int total = rows * c4n;
#pragma unroll
    for (int u = 0; u < NR; ++u) {
        const int idx = threadIdx.x + u * TW;
        if (idx >= total) continue;
        const int r = C4L >= 0 ? (idx >> (C4L >= 0 ? C4L : 0)) : idx / c4n, c4 = idx - r * c4n;
        float4 x;
        x.x = apply_act(R.v[u].x * pro_scale, pro_act, pro_slope) * R.m[u].x;
        x.y = apply_act(R.v[u].y * pro_scale, pro_act, pro_slope) * R.m[u].y;
        x.z = apply_act(R.v[u].z * pro_scale, pro_act, pro_slope) * R.m[u].z;
        x.w = apply_act(R.v[u].w * pro_scale, pro_act, pro_slope) * R.m[u].w;
        if (SPLIT) {
            float4 h, l;
            tc::split_tf32(x.x, h.x, l.x); tc::split_tf32(x.y, h.y, l.y);
            tc::split_tf32(x.z, h.z, l.z); tc::split_tf32(x.w, h.w, l.w);
            *reinterpret_cast<float4*>(xraw + r * stride + c4 * 4) = h;
            *reinterpret_cast<float4*>(xraw_lo + r * stride + c4 * 4) = l;
        } else {
            *reinterpret_cast<float4*>(xraw + r * stride + c4 * 4) = x;
        }
    }
}

// transposed hi/lo operand tile of one tap from the raw rows: elem(frame f, channel c) = xraw[f + shift][c]
template <bool SPLIT, int NX, int TW = 256>
__device__ __forceinline__ void wg_transpose_raw(const float* xraw, const float* xraw_lo, int stride, float* hi, float* lo,
                                                 int cs_floats, int nrows_pad, int shift) {
    const int total = CRK_WG_TF * (nrows_pad >> 2);
#pragma unroll
    for (int u = 0; u < NX; ++u) {
        const int idx = threadIdx.x + u * TW;
        if (idx >= total) continue;
        const int c4 = idx >> 6, f = idx & 63;
        const float4 v = *reinterpret_cast<const float4*>(xraw + (f + shift) * stride + c4 * 4);
        const int off = (f >> 2) * cs_floats + c4 * 16 + (f & 3);
        hi[off] = v.x; hi[off + 4] = v.y; hi[off + 8] = v.z; hi[off + 12] = v.w;
        if (SPLIT) {
            const float4 w = *reinterpret_cast<const float4*>(xraw_lo + (f + shift) * stride + c4 * 4);
            lo[off] = w.x; lo[off + 4] = w.y; lo[off + 8] = w.z; lo[off + 12] = w.w;
        }
    }
}

// TW worker threads (+ one issuer warp).  Round 2: TW = 512 -- with 8 worker warps (2 per scheduler) the per-tile stores, address
// arithmetic and transpositions ran at 6-20 cycles per instruction (measured per phase: G store 1.4 K, raw store 3.0 K, issuing
// the next tile's loads 4.1 K, 5 tap transpositions 4.1 K cycles for ~500 instructions per thread): latency-bound issue, not
// throughput.  16 worker warps halve the per-thread work at the same total.
template <bool SPLIT, int NX, bool VEC, int C4L, int TW>
__global__ void __launch_bounds__(TW + 32, 1) k_wgrad_tc_raw(const WgradTcParams q) {
    // 9 warps: warps 0..7 ("workers") load, stage and transpose; warp 8 only issues MMAs.  With the issue
    // inside a worker warp every tap waited ~1K cycles at the block barrier for that warp's 24 MMAs
    // (measured: 19K of 85K cycles per CTA); now workers hand a finished slot to the issuer through an
    // mbarrier and go straight to the next tap.
    constexpr int SC = TW / 256;                           // worker scale
    constexpr int NR = (NX == 4 ? 6 : 10) / SC;            // raw-tile float4 per thread (host checks the fit)
    constexpr int NG = 8 / SC;                             // G-tile float4 per thread
    constexpr int NT = NX / SC;                            // tap-tile float4 per thread
    constexpr int NWORK = TW;
    const WgradParams& p = q.p;
    extern __shared__ float4 crk_smem4[];
    float* smem = reinterpret_cast<float*>(crk_smem4);
    __shared__ uint64_t bar_full[4];                       // slot written by all workers (count NWORK)
    __shared__ uint64_t bar_slot[4];                       // the MMAs that read the slot have completed
    __shared__ uint64_t bar_tile;                          // all MMAs of a tile have completed
    __shared__ uint32_t tmem_base_s;
    __shared__ int timeout_s;

    constexpr int KCH = CRK_WG_TF / 4;
    constexpr int CSG = 129 * 4;
    const int csx = tc::chunk_rows(q.Npad) * 4;
    float* Gh = smem;
    float* Gl = Gh + KCH * CSG;
    float* ring = Gl + (SPLIT ? KCH * CSG : 0);
    const int xhalf = KCH * csx;
    const int slot_floats = (SPLIT ? 2 : 1) * xhalf;
    const int NS = q.nslot;
    float* xraw = ring + NS * slot_floats;
    float* xraw_lo = xraw + wgrad_raw_half(q.Npad, p.k, p.dil);        // (SPLIT only)
    const int rstride = q.Npad + 4;
    const int c4n = q.Npad >> 2;
    const int rows_raw = CRK_WG_TF + (p.k - 1) * p.dil;
    auto slot_hi = [&](int sl) -> float* { return ring + sl * slot_floats; };
    auto slot_lo = [&](int sl) -> float* { return ring + sl * slot_floats + xhalf; };
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const bool worker = threadIdx.x < NWORK;

    const int tiles_per_utt = (p.T + CRK_WG_TF - 1) / CRK_WG_TF;
    const int ntiles = p.B * tiles_per_utt;
    const int tile_beg = blockIdx.x * p.tiles_per_chunk;
    const int tile_end = min(ntiles, tile_beg + p.tiles_per_chunk);

    if (threadIdx.x == 0) {
        for (int i = 0; i < 4; ++i) { tc::mbar_init(&bar_slot[i], 1); tc::mbar_init(&bar_full[i], NWORK); }
        tc::mbar_init(&bar_tile, 1);
        tc::fence_mbar_init();
        timeout_s = 0;
    }
    if (warp == 1) tc::tmem_alloc<512>(&tmem_base_s);
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    const uint32_t tmem = tmem_base_s;
    bool ok = true;
    int ntile_done = 0;
    float4 bsum[NG];
#pragma unroll
    for (int u = 0; u < NG; ++u) bsum[u] = make_float4(0.f, 0.f, 0.f, 0.f);
    pdl_trigger();                                      // only now: this CTA already owns its TMEM columns (see crk_common.cuh)
    pdl_wait();

    if (worker) {
        WgRegs<NG> RG;
        WgRegs<NR> RX;
        auto load_g = [&](int tile) {
            const int bb = tile / tiles_per_utt;
            const int tt0 = (tile - bb * tiles_per_utt) * CRK_WG_TF;
            wg_load<NG, VEC, TW>(RG, 128, p.G, p.ldg, p.N, bb, p.T, tt0, 0, nullptr, 0);
        };
        auto load_x = [&](int tile) {
            const int bb = tile / tiles_per_utt;
            const int tt0 = (tile - bb * tiles_per_utt) * CRK_WG_TF;
            wg_load_raw<NR, VEC, C4L, TW>(RX, c4n, rows_raw, p.X, p.ldx, p.Cin, bb, p.T, tt0 - p.padl, p.xmul, p.ldxmul);
        };
        int step = 0;
        if (tile_beg < tile_end) { load_x(tile_beg); load_g(tile_beg); }
        for (int tile = tile_beg; tile < tile_end; ++tile) {
            if (tile == tile_beg) dbg_stamp(q.dbg, 0);
            const bool dbg1 = q.dbg && tile == tile_beg + 1 && threadIdx.x == 64;      // second tile: steady-state phase costs
            const long long c0 = dbg1 ? clock64() : 0;
            // Order (round 2): everything that does not touch what the previous tile's MMAs still read -- the raw rows (free once
            // every worker has finished its last transposition: workers-only barrier) and the next tile's X loads -- goes BEFORE
            // the wait for those MMAs; only the G^T store needs them complete.  (The wait was 4.0 K of 14.6 K cycles per tile.)
            if (tile != tile_beg) asm volatile("bar.sync 1, %0;" ::"n"(TW) : "memory");
            wg_store_raw<SPLIT, NR, C4L, TW>(RX, xraw, xraw_lo, rstride, c4n, rows_raw, p.pro_act, p.pro_slope, p.pro_scale);
            const long long c1 = dbg1 ? clock64() : 0;
            if (tile + 1 < tile_end) load_x(tile + 1);      // a whole tile (k tap iterations) of latency cover
            const long long c2 = dbg1 ? clock64() : 0;
            // the previous tile's MMAs read the G^T buffer and the ring slots
            if (ntile_done > 0) { ok &= tc::mbar_wait(&bar_tile, (ntile_done - 1) & 1); tc::tc_fence_after(); }
            if (tile == tile_beg) dbg_stamp(q.dbg, 1);
            const long long c3 = dbg1 ? clock64() : 0;
            wg_store<SPLIT, NG, TW>(RG, Gh, Gl, CSG, 128, CRK_ACT_NONE, 0.f, 1.f);
            if (q.bias) {
#pragma unroll
                for (int u = 0; u < NG; ++u) {
                    bsum[u].x += RG.v[u].x; bsum[u].y += RG.v[u].y; bsum[u].z += RG.v[u].z; bsum[u].w += RG.v[u].w;
                }
            }
            if (tile + 1 < tile_end) load_g(tile + 1);
            const long long c4 = dbg1 ? clock64() : 0;
            asm volatile("bar.sync 1, %0;" ::"n"(TW) : "memory");  // workers only: xraw and G^T complete
            if (dbg1) {
                const long long c5 = clock64();
                // slots: 11 wait prev MMAs, 12 G store (+ next G loads), 13 raw store (+ barrier), 14 issue next X loads, 15 barrier
                dbg_put(1, 11, c3 - c2); dbg_put(1, 12, c4 - c3); dbg_put(1, 13, c1 - c0); dbg_put(1, 14, c2 - c1); dbg_put(1, 15, c5 - c4);
            }
            if (tile == tile_beg) dbg_stamp(q.dbg, 2);
            long long tw = 0, tt = 0;
            for (int j = 0; j < p.k; ++j, ++step) {
                const int sl = step % NS;
                const long long d0 = dbg1 ? clock64() : 0;
                // the MMAs of step - NS (this tile; earlier tiles are covered by bar_tile) read this slot
                if (j >= NS) ok &= tc::mbar_wait(&bar_slot[sl], ((step - NS) / NS) & 1);
                const long long d1 = dbg1 ? clock64() : 0;
                wg_transpose_raw<SPLIT, NT, TW>(xraw, xraw_lo, rstride, slot_hi(sl), slot_lo(sl), csx, q.Npad, j * p.dil);
                tc::fence_proxy_async_smem();               // generic-proxy writes (G^T, slot) -> async proxy
                tc::mbar_arrive(&bar_full[sl]);
                if (dbg1) { tw += d1 - d0; tt += clock64() - d1; }
            }
            if (dbg1) { dbg_put(1, 8, tw); dbg_put(1, 9, tt); }
            if (tile == tile_beg) dbg_stamp(q.dbg, 3);
            ++ntile_done;
        }
        dbg_stamp(q.dbg, 4);
        if (ntile_done > 0) ok &= tc::mbar_wait(&bar_tile, (ntile_done - 1) & 1);
    } else {
        // ===== MMA issuer warp =====
        const uint32_t idesc = tc::make_idesc_tf32(128, q.Npad, 0, 0);
        const uint32_t gh_s = tc::smem_u32(Gh), gl_s = tc::smem_u32(Gl);
        int step = 0;
        for (int tile = tile_beg; tile < tile_end; ++tile) {
            for (int j = 0; j < p.k; ++j, ++step) {
                const int sl = step % NS;
                ok &= tc::mbar_wait(&bar_full[sl], (step / NS) & 1);
                tc::tc_fence_after();
                uint32_t acc = tile > tile_beg ? 1u : 0u;
                // A = G^T (M = 128 rows = co), B = X_j^T (N rows = ci), K = 64 frames
                tc_issue_kmajor_w<SPLIT>(tmem + j * q.Npad, gh_s, gl_s, CSG * 4, 0, tc::smem_u32(slot_hi(sl)),
                                         tc::smem_u32(slot_lo(sl)), csx * 4, CRK_WG_TF, idesc, acc);
                if (tc::elect_one()) {
                    tc::umma_commit(&bar_slot[sl]);
                    if (j == p.k - 1) tc::umma_commit(&bar_tile);
                }
                __syncwarp();
            }
            ++ntile_done;
        }
    }
    tc::tc_fence_after();
    if (!ok) timeout_s = 1;
    __syncthreads();
    dbg_stamp(q.dbg, 5);

    // ---- epilogue (workers; same layout as k_wgrad_tc) ----
    const int co = (warp & 3) * 32 + lane;
    const uint32_t tlane = tmem + ((uint32_t)((warp & 3) * 32) << 16);
    float* gout = p.part + (size_t)blockIdx.x * p.part_stride;
    const bool staged = q.stage_floats >= p.part_stride + 512 && (p.part_stride & 3) == 0 && (reinterpret_cast<uintptr_t>(gout) & 15) == 0;
    float* out = staged ? smem : gout;                     // (see k_wgrad_tc: one TMA bulk store per chunk)
    const int nblk = (q.Npad + 31) >> 5;
    const float poison = __int_as_float(0x7fc00000);
    if (worker)
        for (int pair = warp >> 2; pair < p.k * nblk; pair += TW / 128) {      // (tap j, 32-column block) pairs over the worker warps
            const int j = pair / nblk, blk = pair - j * nblk;
            float v[32];
            if (ntile_done > 0) tc::tmem_ld32(tlane + j * q.Npad + blk * 32, v);
            if (co >= q.TN) continue;
            if (q.tpart) {
                float* dst = out + ((size_t)j * q.TN + co) * p.Rows + blk * 32;
#pragma unroll
                for (int i = 0; i < 32; i += 4)
                    if (blk * 32 + i < p.Rows) {
                        float4 w4 = ntile_done > 0 ? make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]) : make_float4(0.f, 0.f, 0.f, 0.f);
                        if (timeout_s) w4 = make_float4(poison, poison, poison, poison);
                        *reinterpret_cast<float4*>(dst + i) = w4;
                    }
                continue;
            }
#pragma unroll
            for (int i = 0; i < 32; ++i) {
                const int ci = blk * 32 + i;
                if (ci < p.Rows)
                    out[((size_t)j * p.Rows + ci) * q.TN + co] = timeout_s ? poison : (ntile_done > 0 ? v[i] : 0.f);
            }
        }
    if (q.bias) {
        float* red = staged ? smem + p.part_stride : smem;         // (the staged block occupies [0, part_stride))
        if (worker) {
#pragma unroll
            for (int u = 0; u < NG; ++u) {
                float c4[4] = {bsum[u].x, bsum[u].y, bsum[u].z, bsum[u].w};
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    float sv = c4[e];
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1) sv += __shfl_xor_sync(0xffffffffu, sv, o);
                    if (lane == 0) red[warp * 32 + u * 4 + e] = sv;
                }
            }
        }
        __syncthreads();
        if (threadIdx.x < q.TN) {
            const int c = threadIdx.x;                     // G column: quad c/4 = g + (TW/64) u  (g = tid/64 of its owners)
            constexpr int GQ = TW / 64;
            const int quad = c >> 2, g = quad & (GQ - 1), u = quad / GQ;
            const float sv = red[(2 * g) * 32 + u * 4 + (c & 3)] + red[(2 * g + 1) * 32 + u * 4 + (c & 3)];
            out[(size_t)p.k * p.Rows * q.TN + c] = timeout_s ? poison : sv;
        }
    }
    if (staged) {
        tc::fence_proxy_async_smem();
        __syncthreads();
        if (threadIdx.x == 0) {
            tc::bulk_s2g(gout, smem, (uint32_t)(p.part_stride * sizeof(float)));
            tc::bulk_commit_wait_all();
        }
    }
    tc::tc_fence_before();
    __syncthreads();
    dbg_stamp(q.dbg, 6);
    if (warp == 1) tc::tmem_dealloc<512>(tmem);
}

// raw-tile variant: shared-memory plan and applicability
inline size_t wgrad_raw_floats(int Npad, int k, int dil, bool split) { return (split ? 2 : 1) * wgrad_raw_half(Npad, k, dil); }
inline int wgrad_raw_nslot(int Npad, int k, int dil, bool split) {
    const size_t g = (size_t)16 * 129 * 4 * (split ? 2 : 1), x = (size_t)16 * tc::chunk_rows(Npad) * 4 * (split ? 2 : 1);
    const size_t budget = 215 * 1024 / sizeof(float);
    const size_t fixed = g + wgrad_raw_floats(Npad, k, dil, split);
    if (fixed + 2 * x > budget) return 0;
    const size_t n = (budget - fixed) / x;
    return n >= 4 ? 4 : (int)n;
}
inline bool wgrad_raw_ok(const WgradParams& p, int Npad, bool split) {
    if (p.k < 2 || (opt_disable_mask() & 8)) return false;
    const int nr = Npad <= 64 ? 6 : 10;                    // per thread at 256 workers (3 / 5 at 512: the same total)
    if ((CRK_WG_TF + (p.k - 1) * p.dil) * (Npad >> 2) > nr * 256) return false;
    return wgrad_raw_nslot(Npad, p.k, p.dil, split) >= 2;
}
// 128-bit loads legal for both operands (and the multiplier): lets the kernels drop their scalar paths
inline bool wgrad_vec_ok(const WgradParams& p) {
    auto al = [](const float* ptr, int ld) { return ptr == nullptr || (((ld & 3) == 0) && ((reinterpret_cast<uintptr_t>(ptr) & 15) == 0)); };
    return (p.Cin & 3) == 0 && (p.N & 3) == 0 && al(p.X, p.ldx) && al(p.G, p.ldg) && al(p.xmul, p.ldxmul);
}
template <bool SPLIT, int NX, bool VEC, int C4L, int TW>
inline cudaError_t launch_wgrad_tc_raw_tw(const WgradTcParams& q, int nchunk, cudaStream_t s) {
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(k_wgrad_tc_raw<SPLIT, NX, VEC, C4L, TW>, cudaFuncAttributeMaxDynamicSharedMemorySize, 225 * 1024);
        if (e != cudaSuccess) return e;
        attr_set = true;
    }
    const size_t g = (size_t)16 * 129 * 4 * (SPLIT ? 2 : 1), x = (size_t)16 * tc::chunk_rows(q.Npad) * 4 * (SPLIT ? 2 : 1);
    const size_t smem = (g + q.nslot * x + wgrad_raw_floats(q.Npad, q.p.k, q.p.dil, SPLIT)) * sizeof(float);
    TimedLaunch tl(CRK_K_WGRAD, s, 2.0 * q.p.B * q.p.T * q.p.Cin * q.p.N * q.p.k);
    WgradTcParams qq = q;
    qq.stage_floats = !(opt_enable_mask() & 4) ? 0 : (long long)(smem / sizeof(float));
    cudaError_t le = launch_pdl(k_wgrad_tc_raw<SPLIT, NX, VEC, C4L, TW>, dim3(nchunk), dim3(TW + 32), smem, s, qq);
    if (le != cudaSuccess) return le;
    return launch_check();
}
template <bool SPLIT, int NX, bool VEC, int C4L>
inline cudaError_t launch_wgrad_tc_raw_nx(const WgradTcParams& q, int nchunk, cudaStream_t s) {
    // 512 workers need the per-thread counts (8 G, 6 / 10 raw, NX tap float4 at 256 workers) to halve evenly: NX = 4 and 8 both do
    if (opt_disable_mask() & 1024) return launch_wgrad_tc_raw_tw<SPLIT, NX, VEC, C4L, 256>(q, nchunk, s);
    return launch_wgrad_tc_raw_tw<SPLIT, NX, VEC, C4L, 512>(q, nchunk, s);
}

inline int wgrad_tc_gbuf(int Npad, bool split) {       // two G^T buffers when they fit next to two X^T slots
    const size_t g = (size_t)16 * 129 * 4 * (split ? 2 : 1), x = (size_t)16 * tc::chunk_rows(Npad) * 4 * (split ? 2 : 1);
    return (2 * g + 2 * x) * sizeof(float) <= 215 * 1024 ? 2 : 1;
}
inline int wgrad_tc_nslot(int Npad, bool split) {
    const size_t g = (size_t)16 * 129 * 4 * (split ? 2 : 1) * wgrad_tc_gbuf(Npad, split), x = (size_t)16 * tc::chunk_rows(Npad) * 4 * (split ? 2 : 1);
    size_t n = (215 * 1024 / sizeof(float) - g) / x;
    return n >= 4 ? 4 : (n >= 3 ? 3 : 2);
}
inline size_t wgrad_tc_smem(int Npad, bool split) {
    const size_t g = (size_t)16 * 129 * 4, x = (size_t)16 * tc::chunk_rows(Npad) * 4;
    return ((split ? 2 : 1) * wgrad_tc_gbuf(Npad, split) * g + (split ? 2 : 1) * wgrad_tc_nslot(Npad, split) * x) * sizeof(float);
}
inline bool wgrad_tc_ok(const WgradParams& p, int TN, int Npad, bool split) {
    return TN <= 128 && Npad >= 16 && Npad <= 128 && p.k * Npad <= 512 && p.Rows <= Npad &&  // X^T tile <= 8 float4/thread
          
           wgrad_tc_smem(Npad, split) <= 220 * 1024;
}

struct WgradTcWork { int nchunk, tiles_per_chunk; };
inline WgradTcWork wgrad_tc_work(int B, int T) {
    const int ntiles = B * cdiv(T, CRK_WG_TF);
    int nchunk = ntiles < 148 ? ntiles : 148;
    WgradTcWork w;
    w.tiles_per_chunk = cdiv(ntiles, nchunk);
    w.nchunk = cdiv(ntiles, w.tiles_per_chunk);
    return w;
}

template <bool SPLIT, int NX, bool VEC, int TW>
inline cudaError_t launch_wgrad_tc_tw(const WgradTcParams& q, int nchunk, cudaStream_t s) {
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(k_wgrad_tc<SPLIT, NX, VEC, TW>, cudaFuncAttributeMaxDynamicSharedMemorySize, 225 * 1024);
        if (e != cudaSuccess) return e;
        attr_set = true;
    }
    TimedLaunch tl(CRK_K_WGRAD, s, 2.0 * q.p.B * q.p.T * q.p.Cin * q.p.N * q.p.k);
    WgradTcParams qq = q;
    qq.stage_floats = !(opt_enable_mask() & 4) ? 0 : (long long)(wgrad_tc_smem(q.Npad, SPLIT) / sizeof(float));
    cudaError_t le = launch_pdl(k_wgrad_tc<SPLIT, NX, VEC, TW>, dim3(nchunk), dim3(TW), wgrad_tc_smem(q.Npad, SPLIT), s, qq);
    if (le != cudaSuccess) return le;
    return launch_check();
}
template <bool SPLIT, int NX, bool VEC>
inline cudaError_t launch_wgrad_tc_nx(const WgradTcParams& q, int nchunk, cudaStream_t s) {
    if (SPLIT && !(opt_disable_mask() & 1024)) return launch_wgrad_tc_tw<SPLIT, NX, VEC, 512>(q, nchunk, s);
    return launch_wgrad_tc_tw<SPLIT, NX, VEC, 256>(q, nchunk, s);
}
template <bool SPLIT>
inline cudaError_t launch_wgrad_tc_t(const WgradTcParams& q, int nchunk, cudaStream_t s) {
    if (wgrad_vec_ok(q.p))
        return q.Npad <= 64 ? launch_wgrad_tc_nx<SPLIT, 4, true>(q, nchunk, s) : launch_wgrad_tc_nx<SPLIT, 8, true>(q, nchunk, s);
    return q.Npad <= 64 ? launch_wgrad_tc_nx<SPLIT, 4, false>(q, nchunk, s) : launch_wgrad_tc_nx<SPLIT, 8, false>(q, nchunk, s);
}
template <bool SPLIT>
inline cudaError_t launch_wgrad_tc_raw_t(const WgradTcParams& q, int nchunk, cudaStream_t s) {
    if (wgrad_vec_ok(q.p)) {
        if (q.Npad == 64) return launch_wgrad_tc_raw_nx<SPLIT, 4, true, 4>(q, nchunk, s);     // the WaveNet blocks
        return q.Npad <= 64 ? launch_wgrad_tc_raw_nx<SPLIT, 4, true, -1>(q, nchunk, s) : launch_wgrad_tc_raw_nx<SPLIT, 8, true, -1>(q, nchunk, s);
    }
    return q.Npad <= 64 ? launch_wgrad_tc_raw_nx<SPLIT, 4, false, -1>(q, nchunk, s) : launch_wgrad_tc_raw_nx<SPLIT, 8, false, -1>(q, nchunk, s);
}

// fused_bias: the caller wants the bias-gradient partial behind each chunk's weight partial; *bias_done
// tells it whether this kernel produced it (else k_colsum must run)
inline bool wgrad_tc_try(WgradParams& p, int TN, float* part, cudaStream_t s, int* nchunk, cudaError_t* err,
                         bool fused_bias, bool* bias_done, bool* tpart) {
    const int mode = precision_mode();
    if (mode == CRK_PREC_FP32 || (tc_disable_mask() & 4)) return false;
    const bool split = mode == CRK_PREC_TF32X3;
    const int Npad = round_up(p.Cin, 16);
    if (!wgrad_tc_ok(p, TN, Npad, split)) return false;
    const WgradTcWork w = wgrad_tc_work(p.B, p.T);
    WgradTcParams q;
    q.p = p; q.p.part = part; q.p.tiles_per_chunk = w.tiles_per_chunk; q.TN = TN; q.Npad = Npad; q.dbg = dbg_take(CRK_K_WGRAD); q.nslot = wgrad_tc_nslot(Npad, split);
    q.gbuf = wgrad_tc_gbuf(Npad, split);
    q.tpart = ((p.Rows & 3) == 0 && (p.part_stride & 3) == 0 && (reinterpret_cast<uintptr_t>(part) & 15) == 0 && (opt_enable_mask() & 8)) ? 1 : 0;
    *tpart = q.tpart != 0;
    q.bias = (fused_bias && !(opt_disable_mask() & 2)) ? 1 : 0;
    *bias_done = q.bias != 0;
    *nchunk = w.nchunk;
    if (wgrad_raw_ok(p, Npad, split)) {
        q.nslot = wgrad_raw_nslot(Npad, p.k, p.dil, split);
        *err = split ? launch_wgrad_tc_raw_t<true>(q, w.nchunk, s) : launch_wgrad_tc_raw_t<false>(q, w.nchunk, s);
        return true;
    }
    *err = split ? launch_wgrad_tc_t<true>(q, w.nchunk, s) : launch_wgrad_tc_t<false>(q, w.nchunk, s);
    return true;
}

}  // namespace crk
