// crank-b200: vector quantiser kernels (fp32 exact path).
//
// Restates Quantizer.vq / Quantizer.forward of the reference (crank/net/module/vqvae2.py:306-347):
//   dist = sum(W^2,1) - 2*x@W^T + sum(x^2,1)  (that association order, fp32) ; idx = argmin(dist,1)
//   e = W[idx]  (the reference does it as one_hot(idx).float() @ W)
//   qx = x + (e - x)   (value of the straight-through estimator, vqvae2.py:333)
//   EMA: counts / per-code sums of x / decay / Laplace smoothing / new codebook (vqvae2.py:315-330)
// The reference materialises dist (2 KB/frame), an int64 one-hot (4 KB/frame) and its float copy;
// here a frame costs one 256 B read, one int64 + two 256 B writes.
#pragma once
#include "crk_common.cuh"

namespace crk {

// WT[d][k] = W[k][d];  wn[k] = sum_d W[k][d]^2 (d ascending)
__global__ void k_vq_prepare(const float* __restrict__ W, float* __restrict__ WT, float* __restrict__ wn,
                             int K, int D) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= K) return;
    float s = 0.f;
    for (int d = 0; d < D; ++d) {
        const float w = W[(size_t)k * D + d];
        s = fmaf(w, w, s);
        WT[(size_t)d * K + k] = w;
    }
    wn[k] = s;
}

struct VqArgminParams {
    const float* x; int ldx;
    const float* W; const float* WT; const float* wn;
    long long* idx; float* e; int lde; float* qx; int ldqx;
    long long F; int K;
};

// D == 64.  CTA = 64 frames; distance GEMM in passes of 128 codes.
__global__ void __launch_bounds__(CRK_THREADS) k_vq_argmin(const VqArgminParams p) {
    extern __shared__ float4 crk_smem4[];
    float* smem = reinterpret_cast<float*>(crk_smem4);
    float* xs = smem;                 // [64][64]
    float* ws = smem + 64 * 64;       // [64][128]
    float* xn = ws + 64 * 128;        // [64]
    int* best = reinterpret_cast<int*>(xn + 64);  // [64]
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const long long f0 = (long long)blockIdx.x * 64;

    for (int i = threadIdx.x; i < 64 * 16; i += CRK_THREADS) {
        const int r = i >> 4, c4 = i & 15;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (f0 + r < p.F) {
            const float* src = p.x + (size_t)(f0 + r) * p.ldx;
            if (((p.ldx & 3) == 0) && ((reinterpret_cast<uintptr_t>(p.x) & 15) == 0))
                v = __ldg(reinterpret_cast<const float4*>(src) + c4);
            else
                v = make_float4(src[c4 * 4], src[c4 * 4 + 1], src[c4 * 4 + 2], src[c4 * 4 + 3]);
        }
        reinterpret_cast<float4*>(xs)[i] = v;
    }
    __syncthreads();
    if (threadIdx.x < 64) {
        float s = 0.f;
        for (int d = 0; d < 64; ++d) { const float v = xs[threadIdx.x * 64 + d]; s = fmaf(v, v, s); }
        xn[threadIdx.x] = s;
    }
    float bd[8]; int bk[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) { bd[i] = __int_as_float(0x7f800000); bk[i] = 0x7fffffff; }

    for (int pass = 0; pass < p.K / 128; ++pass) {
        __syncthreads();   // previous pass done with ws (and xn visible on first pass)
        for (int i = threadIdx.x; i < 64 * 32; i += CRK_THREADS) {
            const int d = i >> 5, c4 = i & 31;
            reinterpret_cast<float4*>(ws)[i] =
                __ldg(reinterpret_cast<const float4*>(p.WT + (size_t)d * p.K + pass * 128) + c4);
        }
        __syncthreads();
        float acc[8][4];
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
            for (int c = 0; c < 4; ++c) acc[i][c] = 0.f;
        tile_mac_rowA<4>(acc, xs + ty * 8 * 64, 64, ws + tx * 4, 128, 64);
        const float4 wn4 = __ldg(reinterpret_cast<const float4*>(p.wn + pass * 128) + tx);
        const float wnv[4] = {wn4.x, wn4.y, wn4.z, wn4.w};
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const float xnv = xn[ty * 8 + i];
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                const float t2 = 2.f * acc[i][c];
                const float dist = __fadd_rn(__fsub_rn(wnv[c], t2), xnv);
                const int k = pass * 128 + tx * 4 + c;
                if (dist < bd[i]) { bd[i] = dist; bk[i] = k; }   // k ascending per thread => first min wins
            }
        }
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const float od = __shfl_xor_sync(0xffffffffu, bd[i], o);
            const int ok = __shfl_xor_sync(0xffffffffu, bk[i], o);
            if (od < bd[i] || (od == bd[i] && ok < bk[i])) { bd[i] = od; bk[i] = ok; }
        }
        if (tx == 0) best[ty * 8 + i] = bk[i] == 0x7fffffff ? 0 : bk[i];
    }
    __syncthreads();
    if (threadIdx.x < 64 && f0 + threadIdx.x < p.F) p.idx[f0 + threadIdx.x] = (long long)best[threadIdx.x];
    for (int i = threadIdx.x; i < 64 * 16; i += CRK_THREADS) {
        const int r = i >> 4, c4 = i & 15;
        if (f0 + r >= p.F) continue;
        const float4 ev = __ldg(reinterpret_cast<const float4*>(p.W + (size_t)best[r] * 64) + c4);
        const float4 xv = reinterpret_cast<const float4*>(xs)[i];
        float4 q;
        q.x = __fadd_rn(xv.x, __fsub_rn(ev.x, xv.x));
        q.y = __fadd_rn(xv.y, __fsub_rn(ev.y, xv.y));
        q.z = __fadd_rn(xv.z, __fsub_rn(ev.z, xv.z));
        q.w = __fadd_rn(xv.w, __fsub_rn(ev.w, xv.w));
        float* ed = p.e + (size_t)(f0 + r) * p.lde + c4 * 4;
        float* qd = p.qx + (size_t)(f0 + r) * p.ldqx + c4 * 4;
        if (((p.lde & 3) == 0) && ((reinterpret_cast<uintptr_t>(p.e) & 15) == 0)) *reinterpret_cast<float4*>(ed) = ev;
        else { ed[0] = ev.x; ed[1] = ev.y; ed[2] = ev.z; ed[3] = ev.w; }
        if (((p.ldqx & 3) == 0) && ((reinterpret_cast<uintptr_t>(p.qx) & 15) == 0)) *reinterpret_cast<float4*>(qd) = q;
        else { qd[0] = q.x; qd[1] = q.y; qd[2] = q.z; qd[3] = q.w; }
    }
}

// ---- EMA statistics (deterministic): warp w of a CTA owns the codes with (k & 7) == w, scans the
// chunk's indices and accumulates the owned frames in frame order into a shared-memory table.
struct VqStatsParams {
    const float* x; int ldx; const long long* idx;
    float* part;          // [nchunk][K*64 + K]
    long long F; int K; long long frames_per_chunk;
};

__global__ void __launch_bounds__(CRK_THREADS) k_vq_stats(const VqStatsParams p) {
    extern __shared__ float4 crk_smem4[];
    float* tab = reinterpret_cast<float*>(crk_smem4);   // [K][64]
    float* cnt = tab + (size_t)p.K * 64;                // [K]
    // the 32 frames of a batch are staged in shared memory by the whole CTA (coalesced, all loads in flight)
    // before the owning warps accumulate them: the first version fetched each owned row from global memory
    // inside the serial per-frame loop (one exposed L2 round trip per frame: 73 us per call)
    float* xs = cnt + (((size_t)p.K + 3) & ~(size_t)3); // [2][32][64], 16 B aligned
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < p.K * 65; i += CRK_THREADS) tab[i] = 0.f;
    const long long beg = (long long)blockIdx.x * p.frames_per_chunk;
    const long long end = min(p.F, beg + p.frames_per_chunk);
    const bool vec = ((p.ldx & 3) == 0) && ((reinterpret_cast<uintptr_t>(p.x) & 15) == 0);
    int buf = 0;
    for (long long fb = beg; fb < end; fb += 32, buf ^= 1) {
        float* xb = xs + buf * (32 * 64);
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            const int i = threadIdx.x + u * CRK_THREADS;          // 512 float4 = 32 rows x 16
            const int r = i >> 4, c4 = i & 15;
            const long long f = fb + r;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (f < end) {
                const float* src = p.x + (size_t)f * p.ldx + c4 * 4;
                if (vec) v = __ldg(reinterpret_cast<const float4*>(src));
                else v = make_float4(__ldg(src), __ldg(src + 1), __ldg(src + 2), __ldg(src + 3));
            }
            reinterpret_cast<float4*>(xb)[i] = v;
        }
        const long long f = fb + lane;
        int k = -1;
        if (f < end) k = (int)p.idx[f];
        __syncthreads();            // batch staged (and, first time, table zeroed); the other buffer is free again
        const bool owned = (k >= 0) && ((k & 7) == w);
        unsigned m = __ballot_sync(0xffffffffu, owned);
        while (m) {
            const int l = __ffs(m) - 1;
            m &= m - 1;
            const int kk = __shfl_sync(0xffffffffu, k, l);
            tab[kk * 64 + lane] += xb[l * 64 + lane];
            tab[kk * 64 + lane + 32] += xb[l * 64 + lane + 32];
            if (lane == 0) cnt[kk] += 1.f;
            __syncwarp();
        }
    }
    __syncthreads();
    float* out = p.part + (size_t)blockIdx.x * ((size_t)p.K * 65);
    for (int i = threadIdx.x; i < p.K * 65; i += CRK_THREADS) out[i] = tab[i];
}

// counts[k] = sum_c part[c].cnt[k];  esum[d][k] = sum_c part[c].tab[k][d]
__global__ void k_vq_stats_reduce(const float* __restrict__ part, int nchunk, int K, float* __restrict__ counts,
                                  float* __restrict__ esum) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int n = K * 64;
    const size_t stride = (size_t)K * 65;
    if (i < n) {
        const int d = i / K, k = i - d * K;    // output index (d,k): coalesced writes
        float s = 0.f;
        for (int c = 0; c < nchunk; ++c) s += part[c * stride + (size_t)k * 64 + d];
        esum[i] = s;
    } else if (i < n + K) {
        const int k = i - n;
        float s = 0.f;
        for (int c = 0; c < nchunk; ++c) s += part[c * stride + n + k];
        counts[k] = s;
    }
}

// EMA update (vqvae2.py:315-330) in two launches: k_vq_ema_size (one CTA: the K cluster sizes and their
// Laplace-smoothed normalisation, same reduction order as before) and k_vq_ema_w (many CTAs: the D x K
// running sums and the new codebook; the single-CTA version took 46 us per call for 32 K elements).
// ema_w / esum are (D,K) like the reference buffer.
__global__ void __launch_bounds__(512) k_vq_ema_size(const float* __restrict__ counts, float* __restrict__ ema_size,
                                                     float decay, float one_m_decay, float eps, float keps, int K) {
    __shared__ float red[512];
    float loc = 0.f;
    for (int k = threadIdx.x; k < K; k += 512) {
        const float s = __fadd_rn(__fmul_rn(decay, ema_size[k]), __fmul_rn(one_m_decay, counts[k]));
        ema_size[k] = s;
        loc += s;
    }
    red[threadIdx.x] = loc;
    __syncthreads();
    for (int o = 256; o > 0; o >>= 1) {
        if (threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
        __syncthreads();
    }
    const float n = red[0];
    const float den = __fadd_rn(n, keps);
    for (int k = threadIdx.x; k < K; k += 512) {
        const float s = __fmul_rn(__fdiv_rn(__fadd_rn(ema_size[k], eps), den), n);
        ema_size[k] = s;
    }
}
__global__ void __launch_bounds__(256) k_vq_ema_w(const float* __restrict__ esum, const float* __restrict__ ema_size,
                                                  float* __restrict__ ema_w, float* __restrict__ W, float decay,
                                                  float one_m_decay, int K, int D) {
    const int i = blockIdx.x * 256 + threadIdx.x;
    if (i >= K * D) return;
    const int d = i / K, k = i - d * K;
    const float wv = __fadd_rn(__fmul_rn(decay, ema_w[i]), __fmul_rn(one_m_decay, esum[i]));
    ema_w[i] = wv;
    W[(size_t)k * D + d] = __fdiv_rn(wv, ema_size[k]);
}

__global__ void k_vq_scatter_grad(const float* __restrict__ g, int ldg, const long long* __restrict__ idx,
                                  float* __restrict__ dW, long long F, int D) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= F * D) return;
    const long long f = i / D;
    const int d = (int)(i - f * D);
    atomicAdd(dW + (size_t)idx[f] * D + d, g[(size_t)f * ldg + d]);
}

inline int vq_stats_chunks(long long F) {
    long long n = cdivl(F, 512);
    if (n > 148) n = 148;
    if (n < 1) n = 1;
    return (int)n;
}

}  // namespace crk
