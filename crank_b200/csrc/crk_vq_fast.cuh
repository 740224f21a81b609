// crank-b200: VQ L2-argmin, round 2: ONE TF32 tensor-core pass + exact fp32 re-score of the candidates, persistent CTAs
// with the whole codebook resident in shared memory.
//
// Reference: Quantizer.vq (crank/net/module/vqvae2.py:338-347: (F, K) distance matrix, argmin) + the gather and the
// straight-through output of Quantizer.forward (:306-336).  Round 1 (crk_vq_tc.cuh) ran the distance GEMM in 3xTF32
// (96 MMAs of N = 128 per tile, 198 KB of hi/lo operand tiles, the codebook streamed through a ring by every CTA, X read
// twice) and took 62 us per 32 000 frames.  The argmin does not need fp32-accurate distances, only the right index:
//   1. dots[128 frames][K] = X . W^T as PLAIN TF32 (the tensor core truncates the raw fp32 operands): 16 MMAs of
//      128 x 256 x 8 per tile; raw fp32 tiles in shared memory, so the SAME bytes serve the exact re-score and the gather;
//   2. rigorous radius: |dist~ - dist_fp32| <= c (|x|^2 + |w|^2), c = 2.5e-3 >= 2^-9 (two truncated operands, Cauchy-
//      Schwarz, 2 dot) -- pass 1 over tensor memory finds the best upper bound, pass 2 lists every code whose lower
//      bound does not exceed it (1-2 codes per frame in practice), and those are re-scored with EXACTLY the fp32
//      arithmetic of k_vq_argmin (ascending-d FMA dot, fl(fl(wn - 2 dot) + xn), lowest index wins ties);
//   3. gather e = W[idx] (from shared memory), qx = x + (e - x), idx as int64.
// The index is therefore identical to k_vq_argmin's whenever the radius bound holds (tests/test_gpu_tc.py compares them
// on fresh, EMA-warmed and dead-code codebooks).  grid = min(tiles, SMs); a CTA loads the operand blob (131 KB for
// K = 512, written by k_vq_pack_op / the EMA kernel) ONCE with TMA bulk copies and loops over its tiles.
// D = 64, K a multiple of 128, K <= 512.
#pragma once
#include "crk_common.cuh"
#include "crk_resblock_tc.cuh"
#include "crk_tc.cuh"
#include "crk_vq.cuh"

namespace crk {

#define CRK_VQ_TF32_RADIUS 2.5e-3f
#define CRK_VQ_MAXCAND 6

__host__ __device__ constexpr int vq_op_rows(int K) { return tc::chunk_rows(K); }
__host__ __device__ constexpr long long vq_op_floats(int K) { return 16LL * vq_op_rows(K) * 4 + K; }

// operand blob: raw fp32 codebook, chunk-major [16][rows(K)][4], then wn[K] (|w|^2 in k_vq_prepare's summation order)
__global__ void k_vq_pack_op(const float* __restrict__ W, float* __restrict__ blob, int K) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= K) return;
    const int cr4 = vq_op_rows(K) * 4;
    float s = 0.f;
    for (int d = 0; d < 64; ++d) {
        const float w = W[(size_t)k * 64 + d];
        s = fmaf(w, w, s);
        blob[(size_t)(d >> 2) * cr4 + k * 4 + (d & 3)] = w;
    }
    blob[(size_t)16 * cr4 + k] = s;
}

struct VqFastParams {
    VqArgminParams p;        // W / WT / wn unused
    const float* opblob;
};

// 512 threads: the bound scans are instruction-bound, and with 8 warps (2 per scheduler) latency-bound on top (round 2: the same
// finding as in the wgrad kernels); 16 warps split each frame's K codes in four column parts instead of two.
#define CRK_VQ_THREADS 512
#define CRK_VQ_PARTS (CRK_VQ_THREADS / 128)
__global__ void __launch_bounds__(CRK_VQ_THREADS, 1) k_vq_argmin_tf32(const VqFastParams q) {
    const VqArgminParams& p = q.p;
    extern __shared__ float4 crk_smem4[];
    float* smem = reinterpret_cast<float*>(crk_smem4);
    __shared__ uint64_t bar_w, bar_acc;
    __shared__ uint32_t tmem_base_s;
    __shared__ int timeout_s;
    __shared__ float xn_s[128];
    __shared__ float mbu_s[CRK_VQ_PARTS][128];
    __shared__ float med_s[CRK_VQ_PARTS][128];
    __shared__ int mek_s[CRK_VQ_PARTS][128];
    __shared__ int best_s[128];

    const int K = p.K;
    const int CR = vq_op_rows(K), CR4 = CR * 4;
    constexpr int CSX = 129 * 4;
    float* Wb = smem;                          // [16][CR][4] raw fp32 codebook (tensor-core B operand AND exact copy)
    float* wn_s = Wb + 16 * CR4;               // [K]
    float* Xt = wn_s + K;                      // [16][129][4] raw fp32 X tile (A operand)
    float* xrows = Xt + 16 * CSX;              // [128][65] row-major copy (|x|^2, re-score, outputs)
    float* wa_s = xrows + 128 * 65;            // [K] wn (1 + c)
    float* wb_s = wa_s + K;                    // [K] wn (1 - c)
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long long ntiles = (p.F + 127) / 128;

    if (threadIdx.x == 0) {
        tc::mbar_init(&bar_w, 1);
        tc::mbar_init(&bar_acc, 1);
        tc::fence_mbar_init();
        timeout_s = 0;
    }
    if (warp == 1) tc::tmem_alloc<512>(&tmem_base_s);
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    const uint32_t tmem = tmem_base_s;
    bool ok = true;
    if (threadIdx.x == 0) {
        const uint32_t chunk_bytes = (uint32_t)CR4 * 4u;
        tc::mbar_arrive_expect_tx(&bar_w, 16u * chunk_bytes + (uint32_t)K * 4u);
        for (int c = 0; c < 16; ++c) tc::bulk_g2s(Wb + c * CR4, q.opblob + (size_t)c * CR4, chunk_bytes, &bar_w);
        tc::bulk_g2s(wn_s, q.opblob + (size_t)16 * CR4, (uint32_t)K * 4u, &bar_w);
    }
    const int NB = (K % 256 == 0) ? 256 : 128;
    const uint32_t idesc = tc::make_idesc_tf32(128, NB, 0, 0);
    const bool vec = ((p.ldx & 3) == 0) && ((reinterpret_cast<uintptr_t>(p.x) & 15) == 0);
    const bool vec_e = ((p.lde & 3) == 0) && ((reinterpret_cast<uintptr_t>(p.e) & 15) == 0);
    const bool vec_q = ((p.ldqx & 3) == 0) && ((reinterpret_cast<uintptr_t>(p.qx) & 15) == 0);
    const float INF = __int_as_float(0x7f800000);
    uint32_t it = 0;

    for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
        const long long f0 = tile * 128;
        const int nlive = (int)min((long long)128, p.F - f0);
        // ---- stage the X tile once: chunk-major operand + row-major copy ----
        {
            constexpr int NU = 2048 / CRK_VQ_THREADS;
            float4 v[NU];
#pragma unroll
            for (int u = 0; u < NU; ++u) {
                const int i = threadIdx.x + u * CRK_VQ_THREADS, r = i >> 4, c4 = i & 15;
                v[u] = make_float4(0.f, 0.f, 0.f, 0.f);
                if (r < nlive) {
                    const float* src = p.x + (size_t)(f0 + r) * p.ldx + c4 * 4;
                    if (vec) v[u] = __ldg(reinterpret_cast<const float4*>(src));
                    else v[u] = make_float4(__ldg(src), __ldg(src + 1), __ldg(src + 2), __ldg(src + 3));
                }
            }
#pragma unroll
            for (int u = 0; u < NU; ++u) {
                const int i = threadIdx.x + u * CRK_VQ_THREADS, r = i >> 4, c4 = i & 15;
                *reinterpret_cast<float4*>(Xt + c4 * CSX + r * 4) = v[u];
                float* d = xrows + r * 65 + c4 * 4;
                d[0] = v[u].x; d[1] = v[u].y; d[2] = v[u].z; d[3] = v[u].w;
            }
        }
        tc::fence_proxy_async_smem();
        __syncthreads();
        if (warp == 1) {
            if (it == 0) ok &= tc::mbar_wait(&bar_w, 0);
            tc::tc_fence_after();
            for (int nb = 0; nb < K / NB; ++nb) {
                uint32_t acc = 0;
                tc_issue_kmajor_w<false>(tmem + nb * NB, tc::smem_u32(Xt), 0u, CSX * 4, 0,
                                         tc::smem_u32(Wb) + (uint32_t)nb * NB * 16u, 0u, CR4 * 4, 64, idesc, acc);
            }
            if (tc::elect_one()) tc::umma_commit(&bar_acc);
            __syncwarp();
        } else if (threadIdx.x >= 128 && threadIdx.x < 256) {
            const int r = threadIdx.x - 128;
            const float* xr = xrows + r * 65;
            float s = 0.f;
            for (int d = 0; d < 64; ++d) s = fmaf(xr[d], xr[d], s);      // same order as k_vq_argmin
            xn_s[r] = s;
        }
        if (it == 0) {
            ok &= tc::mbar_wait(&bar_w, 0);                               // wn_s / Wb visible to every thread
            for (int k = threadIdx.x; k < K; k += CRK_VQ_THREADS) {
                wa_s[k] = wn_s[k] * (1.f + CRK_VQ_TF32_RADIUS);
                wb_s[k] = wn_s[k] * (1.f - CRK_VQ_TF32_RADIUS);
            }
        }
        ok &= tc::mbar_wait(&bar_acc, it & 1);
        tc::tc_fence_after();
        __syncthreads();

        // ---- epilogue: thread = (frame row r, code half) ----
        // bounds with one FMA per code (the scan is instruction-bound: 512 codes x 128 frames per tile):
        //   upper_k = dist~_k + m_k = fma(-2, dot_k, wn_k (1 + c)) + xn (1 + c),   lower_k = fma(-2, dot_k, wn_k (1 - c)) + xn (1 - c)
        // with the per-code tables wa = wn (1 + c), wb = wn (1 - c) in shared memory (c carries 28 % of slack over the
        // rigorous 2^-9: fp32 rounding of these expressions, ~1e-7 relative, cannot break the bounds)
        const int r = (warp & 3) * 32 + lane;
        const int half = warp >> 2;                      // column part of this warp (0 .. CRK_VQ_PARTS-1)
        const int kbeg = half * (K / CRK_VQ_PARTS), kend = kbeg + (K / CRK_VQ_PARTS);
        const float xn = xn_s[r];
        const uint32_t tlane = tmem + ((uint32_t)((warp & 3) * 32) << 16);
        float mub = INF;
        for (int k0 = kbeg; k0 < kend; k0 += 32) {
            float v[32];
            tc::tmem_ld32(tlane + k0, v);
#pragma unroll
            for (int i = 0; i < 32; i += 4) {
                const float4 a = *reinterpret_cast<const float4*>(wa_s + k0 + i);
                mub = fminf(mub, fminf(fminf(fmaf(-2.f, v[i], a.x), fmaf(-2.f, v[i + 1], a.y)),
                                       fminf(fmaf(-2.f, v[i + 2], a.z), fmaf(-2.f, v[i + 3], a.w))));
            }
        }
        mbu_s[half][r] = mub;
        __syncthreads();
        float mall = mbu_s[0][r];
#pragma unroll
        for (int pp = 1; pp < CRK_VQ_PARTS; ++pp) mall = fminf(mall, mbu_s[pp][r]);
        const float bu = fmaf(xn, 1.f + CRK_VQ_TF32_RADIUS, mall);     // best upper bound
        const float thr = bu - xn * (1.f - CRK_VQ_TF32_RADIUS);
        int cand[CRK_VQ_MAXCAND];
        int ncand = 0;
#pragma unroll
        for (int c = 0; c < CRK_VQ_MAXCAND; ++c) cand[c] = 0;
        for (int k0 = kbeg; k0 < kend; k0 += 32) {
            float v[32];
            tc::tmem_ld32(tlane + k0, v);
            uint32_t hit = 0;
#pragma unroll
            for (int i = 0; i < 32; i += 4) {
                const float4 bq = *reinterpret_cast<const float4*>(wb_s + k0 + i);
                hit |= (fmaf(-2.f, v[i], bq.x) <= thr ? 1u : 0u) << i;
                hit |= (fmaf(-2.f, v[i + 1], bq.y) <= thr ? 1u : 0u) << (i + 1);
                hit |= (fmaf(-2.f, v[i + 2], bq.z) <= thr ? 1u : 0u) << (i + 2);
                hit |= (fmaf(-2.f, v[i + 3], bq.w) <= thr ? 1u : 0u) << (i + 3);
            }
            while (hit) {                                   // rare: 1-2 candidates per frame over all 512 codes
                const int i = __ffs(hit) - 1;
                hit &= hit - 1;
#pragma unroll
                for (int c = 0; c < CRK_VQ_MAXCAND; ++c)
                    if (c == ncand) cand[c] = k0 + i;
                ++ncand;
            }
        }
        // exact re-score, warp-converged: round c handles every lane's c-th candidate
        float ed = INF;
        int ek = 0x7fffffff;
        const float* xr = xrows + r * 65;
        const int maxc = __reduce_max_sync(0xffffffffu, min(ncand, CRK_VQ_MAXCAND));
        for (int c = 0; c < maxc; ++c) {
            int kk = 0;
#pragma unroll
            for (int cc = 0; cc < CRK_VQ_MAXCAND; ++cc)
                if (cc == c) kk = cand[cc];
            if (c < ncand) {
                const float4* w4 = reinterpret_cast<const float4*>(Wb) + kk;
                float acc = 0.f;
#pragma unroll
                for (int c4 = 0; c4 < 16; ++c4) {
                    const float4 w = w4[(size_t)c4 * CR];
                    acc = fmaf(xr[4 * c4 + 0], w.x, acc);
                    acc = fmaf(xr[4 * c4 + 1], w.y, acc);
                    acc = fmaf(xr[4 * c4 + 2], w.z, acc);
                    acc = fmaf(xr[4 * c4 + 3], w.w, acc);
                }
                const float ex = __fadd_rn(__fsub_rn(wn_s[kk], 2.f * acc), xn);
                if (ex < ed) { ed = ex; ek = kk; }              // candidates ascend in k: first minimum wins ties
            }
        }
        if (ncand > CRK_VQ_MAXCAND) {
            // rare overflow (a frame with more than MAXCAND near-ties): exact scan of this thread's whole half
            ed = INF; ek = 0x7fffffff;
            for (int kk = kbeg; kk < kend; ++kk) {
                const float4* w4 = reinterpret_cast<const float4*>(Wb) + kk;
                float acc = 0.f;
                for (int c4 = 0; c4 < 16; ++c4) {
                    const float4 w = w4[(size_t)c4 * CR];
                    acc = fmaf(xr[4 * c4 + 0], w.x, acc);
                    acc = fmaf(xr[4 * c4 + 1], w.y, acc);
                    acc = fmaf(xr[4 * c4 + 2], w.z, acc);
                    acc = fmaf(xr[4 * c4 + 3], w.w, acc);
                }
                const float ex = __fadd_rn(__fsub_rn(wn_s[kk], 2.f * acc), xn);
                if (ex < ed) { ed = ex; ek = kk; }
            }
        }
        med_s[half][r] = ed; mek_s[half][r] = ek;
        __syncthreads();
        if (half == 0) {
#pragma unroll
            for (int pp = 1; pp < CRK_VQ_PARTS; ++pp)           // ascending parts: ties keep the lower index
                if (med_s[pp][r] < ed) { ed = med_s[pp][r]; ek = mek_s[pp][r]; }
            best_s[r] = ek == 0x7fffffff ? 0 : ek;
        }
        __syncthreads();
        if (threadIdx.x < nlive) p.idx[f0 + threadIdx.x] = (long long)best_s[threadIdx.x];
#pragma unroll 2
        for (int i = threadIdx.x; i < 128 * 16; i += CRK_VQ_THREADS) {
            const int rr = i >> 4, c4 = i & 15;
            if (rr >= nlive) continue;
            const float4 ev = *(reinterpret_cast<const float4*>(Wb) + (size_t)c4 * CR + best_s[rr]);
            const float* xq = xrows + rr * 65 + c4 * 4;
            float4 qv;
            qv.x = __fadd_rn(xq[0], __fsub_rn(ev.x, xq[0]));
            qv.y = __fadd_rn(xq[1], __fsub_rn(ev.y, xq[1]));
            qv.z = __fadd_rn(xq[2], __fsub_rn(ev.z, xq[2]));
            qv.w = __fadd_rn(xq[3], __fsub_rn(ev.w, xq[3]));
            float* edst = p.e + (size_t)(f0 + rr) * p.lde + c4 * 4;
            float* qdst = p.qx + (size_t)(f0 + rr) * p.ldqx + c4 * 4;
            if (vec_e) *reinterpret_cast<float4*>(edst) = ev;
            else { edst[0] = ev.x; edst[1] = ev.y; edst[2] = ev.z; edst[3] = ev.w; }
            if (vec_q) *reinterpret_cast<float4*>(qdst) = qv;
            else { qdst[0] = qv.x; qdst[1] = qv.y; qdst[2] = qv.z; qdst[3] = qv.w; }
        }
        tc::tc_fence_before();
        __syncthreads();                 // TMEM, Xt, xrows and the merge arrays are free for the next tile
    }
    if (!ok) timeout_s = 1;
    __syncthreads();
    if (timeout_s && threadIdx.x == 0 && (long long)blockIdx.x < ntiles) p.idx[(long long)blockIdx.x * 128] = -1;   // poison
    if (warp == 1) tc::tmem_dealloc<512>(tmem);
}

// ---- EMA codebook update in ONE launch (vqvae2.py:315-330), writing the next call's operand blob ------------------------
// Round 1: k_vq_ema_size (1 CTA) -> k_vq_ema_w -> (next forward) k_vq_prepare -> k_vq_pack_tc.  Here every CTA (512
// threads = 8 codes x 64 dims) recomputes the K cluster sizes and their Laplace-smoothed normalisation in shared memory
// with the SAME reduction order as k_vq_ema_size (so all CTAs hold identical values), updates its 8 codes' running sums,
// codebook rows, operand-blob entries and |w|^2.  ema_size is updated in place by whichever CTA finishes reading it
// LAST (ticket counter, zeroed by the statistics kernel and reset here): no CTA can still need the old values.
__global__ void __launch_bounds__(512) k_vq_ema_fused(const float* __restrict__ counts, const float* __restrict__ esum,
                                                      float* __restrict__ ema_size, float* __restrict__ ema_w,
                                                      float* __restrict__ W, float* __restrict__ opblob, int* ticket,
                                                      float decay, float one_m_decay, float eps, float keps, int K) {
    extern __shared__ float4 crk_smem4[];
    float* size_s = reinterpret_cast<float*>(crk_smem4);       // [K] new cluster sizes
    __shared__ float red[512];
    __shared__ float wrow[8][64];
    __shared__ int last_s;
    float loc = 0.f;
    for (int k = threadIdx.x; k < K; k += 512) {
        const float sv = __fadd_rn(__fmul_rn(decay, ema_size[k]), __fmul_rn(one_m_decay, counts[k]));
        size_s[k] = sv;
        loc += sv;
    }
    red[threadIdx.x] = loc;
    __syncthreads();
    for (int o = 256; o > 0; o >>= 1) {
        if (threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
        __syncthreads();
    }
    const float n = red[0];
    const float den = __fadd_rn(n, keps);
    for (int k = threadIdx.x; k < K; k += 512) size_s[k] = __fmul_rn(__fdiv_rn(__fadd_rn(size_s[k], eps), den), n);
    __syncthreads();                                            // every read of the old ema_size by this CTA is done
    if (threadIdx.x == 0) {
        __threadfence();
        const int t = atomicAdd(ticket, 1);
        last_s = (t == (int)gridDim.x - 1);
    }
    const int kk = threadIdx.x >> 6, d = threadIdx.x & 63;
    const int k = blockIdx.x * 8 + kk;
    if (k < K) {
        const int i = d * K + k;
        const float wv = __fadd_rn(__fmul_rn(decay, ema_w[i]), __fmul_rn(one_m_decay, esum[i]));
        ema_w[i] = wv;
        const float w = __fdiv_rn(wv, size_s[k]);
        W[(size_t)k * 64 + d] = w;
        wrow[kk][d] = w;
        if (opblob) opblob[(size_t)(d >> 2) * (vq_op_rows(K) * 4) + k * 4 + (d & 3)] = w;
    }
    __syncthreads();
    if (k < K && d == 0 && opblob) {
        float sv = 0.f;
        for (int dd = 0; dd < 64; ++dd) sv = fmaf(wrow[kk][dd], wrow[kk][dd], sv);     // k_vq_prepare's order
        opblob[(size_t)16 * (vq_op_rows(K) * 4) + k] = sv;
    }
    if (last_s) {
        __threadfence();
        for (int kq = threadIdx.x; kq < K; kq += 512) ema_size[kq] = size_s[kq];
        if (threadIdx.x == 0) *ticket = 0;
    }
}

inline size_t vq_fast_smem(int K) {
    return (size_t)(16 * vq_op_rows(K) * 4 + 3 * K + 16 * 129 * 4 + 128 * 65) * sizeof(float);
}

}  // namespace crk
