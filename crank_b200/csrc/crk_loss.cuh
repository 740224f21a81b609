// crank-b200: loss kernels (sum/count reductions, no masked_select, no host sync) + Adam.
//
// Restates: CustomFeatureLoss / STFTLoss (crank/net/module/loss.py:18-114), the masked MSE terms of
// the trainers (crank/net/trainer/trainer_vqvae.py:210-239, trainer_lsgan.py:146-173),
// CrossEntropyLoss(ignore_index=-100) (crank/net/trainer/utils.py:26) and torch.optim.Adam.step.
// All reductions are two-pass with a fixed summation order (deterministic).
#pragma once
#include "crk_common.cuh"

namespace crk {

__device__ __forceinline__ float block_sum_256(float v, float* red /*[8]*/) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += red[i];
    return s;
}

// ---- masked L1 / MSE -----------------------------------------------------------------------
struct MaskedLossParams {
    const float* x; int ldx; const float* y; int ldy; float yconst;
    const unsigned char* mask; int B, T, D, shift;
};

__device__ __forceinline__ bool ml_fetch(const MaskedLossParams& p, long long e, float& d, size_t& xoff) {
    const int Tp = p.T - (p.shift < 0 ? -p.shift : p.shift);
    const int c = (int)(e % p.D);
    const long long r = e / p.D;
    const int tp = (int)(r % Tp);
    const int b = (int)(r / Tp);
    const int tx = tp + (p.shift > 0 ? p.shift : 0);
    const int ty = tp + (p.shift < 0 ? -p.shift : 0);
    if (p.mask && !p.mask[(size_t)b * p.T + tx]) return false;
    xoff = ((size_t)b * p.T + tx) * p.ldx + c;
    const float yv = p.y ? p.y[((size_t)b * p.T + ty) * p.ldy + c] : p.yconst;
    d = p.x[xoff] - yv;
    return true;
}

__global__ void __launch_bounds__(CRK_THREADS) k_masked_loss_part(const MaskedLossParams p, float* __restrict__ part) {
    __shared__ float red[8];
    const int Tp = p.T - (p.shift < 0 ? -p.shift : p.shift);
    const long long N = (long long)p.B * Tp * p.D;
    float sa = 0.f, sq = 0.f, cn = 0.f;
    for (long long e = (long long)blockIdx.x * CRK_THREADS + threadIdx.x; e < N; e += (long long)gridDim.x * CRK_THREADS) {
        float d; size_t xo;
        if (ml_fetch(p, e, d, xo)) { sa += fabsf(d); sq = fmaf(d, d, sq); cn += 1.f; }
    }
    sa = block_sum_256(sa, red);
    sq = block_sum_256(sq, red);
    cn = block_sum_256(cn, red);
    if (threadIdx.x == 0) { part[blockIdx.x * 3] = sa; part[blockIdx.x * 3 + 1] = sq; part[blockIdx.x * 3 + 2] = cn; }
}

// out[j] = sum_b part[b*m + j] ; then out[0..nmean-1] /= out[cnt_idx]  (one CTA, fixed order)
__global__ void __launch_bounds__(CRK_THREADS) k_finalize(const float* __restrict__ part, int nblk, int m,
                                                          float* __restrict__ out, int nmean, int cnt_idx) {
    __shared__ float red[8];
    __shared__ float tot[8];
    for (int j = 0; j < m; ++j) {
        float s = 0.f;
        for (int b = threadIdx.x; b < nblk; b += CRK_THREADS) s += part[(size_t)b * m + j];
        s = block_sum_256(s, red);
        if (threadIdx.x == 0) tot[j] = s;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int j = 0; j < m; ++j) out[j] = (j < nmean && cnt_idx >= 0) ? tot[j] / tot[cnt_idx] : tot[j];
    }
}

__global__ void __launch_bounds__(CRK_THREADS) k_masked_loss_bwd(const MaskedLossParams p, const float* __restrict__ out,
                                                                 const float* __restrict__ g_l1,
                                                                 const float* __restrict__ g_mse,
                                                                 float* __restrict__ dx, int lddx) {
    // dx covers the full (B,T,D) panel: zero where not selected
    const long long N = (long long)p.B * p.T * p.D;
    const long long e = (long long)blockIdx.x * CRK_THREADS + threadIdx.x;
    if (e >= N) return;
    const int c = (int)(e % p.D);
    const long long r = e / p.D;
    const int t = (int)(r % p.T);
    const int b = (int)(r / p.T);
    const int lo = p.shift > 0 ? p.shift : 0;
    const int hi = p.shift < 0 ? p.T + p.shift : p.T;
    float gval = 0.f;
    if (t >= lo && t < hi && (!p.mask || p.mask[(size_t)b * p.T + t])) {
        const int tp = t - lo;
        const int ty = tp + (p.shift < 0 ? -p.shift : 0);
        const float yv = p.y ? p.y[((size_t)b * p.T + ty) * p.ldy + c] : p.yconst;
        const float d = p.x[((size_t)b * p.T + t) * p.ldx + c] - yv;
        const float n = out[2];
        if (g_l1) gval += g_l1[0] * (d > 0.f ? 1.f : (d < 0.f ? -1.f : 0.f)) / n;
        if (g_mse) gval += g_mse[0] * 2.f * d / n;
    }
    dx[((size_t)b * p.T + t) * lddx + c] = gval;
}

inline int loss_blocks(long long N) {
    long long n = cdivl(N, CRK_THREADS * 8);
    if (n > 1024) n = 1024;
    if (n < 1) n = 1;
    return (int)n;
}

// ---- cross entropy with ignore_index -----------------------------------------------------------
__global__ void __launch_bounds__(CRK_THREADS) k_ce_part(const float* __restrict__ logits, int ldl,
                                                         const long long* __restrict__ labels, long long F, int S,
                                                         long long ignore, float* __restrict__ part) {
    __shared__ float red[8];
    float sl = 0.f, cn = 0.f;
    for (long long r = (long long)blockIdx.x * CRK_THREADS + threadIdx.x; r < F; r += (long long)gridDim.x * CRK_THREADS) {
        const long long lab = labels[r];
        if (lab == ignore) continue;
        const float* row = logits + (size_t)r * ldl;
        float mx = row[0];
        for (int c = 1; c < S; ++c) mx = fmaxf(mx, row[c]);
        float se = 0.f;
        for (int c = 0; c < S; ++c) se += expf(row[c] - mx);
        sl += -((row[lab] - mx) - logf(se));
        cn += 1.f;
    }
    sl = block_sum_256(sl, red);
    cn = block_sum_256(cn, red);
    if (threadIdx.x == 0) { part[blockIdx.x * 2] = sl; part[blockIdx.x * 2 + 1] = cn; }
}

__global__ void __launch_bounds__(CRK_THREADS) k_ce_bwd(const float* __restrict__ logits, int ldl,
                                                        const long long* __restrict__ labels, long long F, int S,
                                                        long long ignore, const float* __restrict__ out,
                                                        const float* __restrict__ g, float* __restrict__ dl, int lddl) {
    const long long r = (long long)blockIdx.x * CRK_THREADS + threadIdx.x;
    if (r >= F) return;
    const long long lab = labels[r];
    float* drow = dl + (size_t)r * lddl;
    if (lab == ignore) {
        for (int c = 0; c < S; ++c) drow[c] = 0.f;
        return;
    }
    const float* row = logits + (size_t)r * ldl;
    float mx = row[0];
    for (int c = 1; c < S; ++c) mx = fmaxf(mx, row[c]);
    float se = 0.f;
    for (int c = 0; c < S; ++c) se += expf(row[c] - mx);
    const float scale = g[0] / out[1];
    for (int c = 0; c < S; ++c) {
        const float sm = expf(row[c] - mx) / se;
        drow[c] = (sm - (c == lab ? 1.f : 0.f)) * scale;
    }
}

// ---- STFT-magnitude trajectory L1 --------------------------------------------------------------
struct StftParams {
    const float* x; int ldx; const float* y; int ldy;
    int B, T, D, n_fft, hop, win, M, bins;
};

__device__ __forceinline__ int reflect_idx(int p, int T) {
    if (p < 0) p = -p;
    if (p >= T) p = 2 * (T - 1) - p;
    return p;
}

// cos/sin tables + periodic hann window in shared memory
__device__ __forceinline__ void stft_tables(float* ct, float* st, float* wv, int n_fft, int win) {
    for (int j = threadIdx.x; j < n_fft; j += blockDim.x) {
        double s, c;
        sincospi(2.0 * (double)j / (double)n_fft, &s, &c);
        ct[j] = (float)c; st[j] = (float)s;
    }
    for (int n = threadIdx.x; n < win; n += blockDim.x)
        wv[n] = (float)(0.5 - 0.5 * cospi(2.0 * (double)n / (double)win));
}

__device__ __forceinline__ void stft_bin(const StftParams& p, const float* src, int ld, int b, int d, int m, int bin,
                                         const float* ct, const float* st, const float* wv, float& re, float& im) {
    const int woff = (p.n_fft - p.win) / 2;
    re = 0.f; im = 0.f;
    for (int n = 0; n < p.win; ++n) {
        const int pos = reflect_idx(m * p.hop + woff + n - p.n_fft / 2, p.T);
        const float a = wv[n] * src[((size_t)b * p.T + pos) * ld + d];
        const int ph = (bin * (woff + n)) % p.n_fft;
        re = fmaf(a, ct[ph], re);
        im = fmaf(-a, st[ph], im);
    }
}

// element e -> (b, m, bin, d) with d fastest (coalesced across the feature dimension)
__global__ void __launch_bounds__(CRK_THREADS) k_stft_loss_part(const StftParams p, float* __restrict__ part) {
    extern __shared__ float4 crk_smem4[];
    float* ct = reinterpret_cast<float*>(crk_smem4);
    float* st = ct + p.n_fft;
    float* wv = st + p.n_fft;
    __shared__ float red[8];
    stft_tables(ct, st, wv, p.n_fft, p.win);
    __syncthreads();
    const long long N = (long long)p.B * p.M * p.bins * p.D;
    float sm = 0.f, slg = 0.f;
    for (long long e = (long long)blockIdx.x * CRK_THREADS + threadIdx.x; e < N; e += (long long)gridDim.x * CRK_THREADS) {
        const int d = (int)(e % p.D);
        long long r = e / p.D;
        const int bin = (int)(r % p.bins); r /= p.bins;
        const int m = (int)(r % p.M);
        const int b = (int)(r / p.M);
        float rx, ix, ry, iy;
        stft_bin(p, p.x, p.ldx, b, d, m, bin, ct, st, wv, rx, ix);
        stft_bin(p, p.y, p.ldy, b, d, m, bin, ct, st, wv, ry, iy);
        const float mx = sqrtf(fmaxf(fmaf(rx, rx, ix * ix), 1e-7f));
        const float my = sqrtf(fmaxf(fmaf(ry, ry, iy * iy), 1e-7f));
        sm += fabsf(mx - my);
        slg += fabsf(logf(mx) - logf(my));
    }
    sm = block_sum_256(sm, red);
    slg = block_sum_256(slg, red);
    if (threadIdx.x == 0) { part[blockIdx.x * 2] = sm; part[blockIdx.x * 2 + 1] = slg; }
}

// ---- frame-per-CTA versions (used when a frame's working set fits shared memory) ---------------------
// One CTA per (utterance b, STFT frame m): the win x D windowed samples of x and y are staged ONCE in shared
// memory and every (bin, feature) pair reads them from there.  The element-per-thread kernels above re-read
// each sample from global memory once per bin (33-65x) and measured 200 us per call on the recipe shapes,
// 3% of the whole LSGAN step for ~0.1 GFLOP.  Same per-bin arithmetic and summation order over n.
struct StftFrameSmem {
    float *ct, *st, *wv, *xw, *yw, *cre, *cim;
};
__device__ __forceinline__ StftFrameSmem stft_frame_smem(float* base, const StftParams& p, bool bwd) {
    StftFrameSmem s;
    s.ct = base; s.st = s.ct + p.n_fft; s.wv = s.st + p.n_fft;
    s.xw = s.wv + p.n_fft; s.yw = s.xw + p.win * p.D;
    s.cre = s.yw + p.win * p.D; s.cim = bwd ? s.cre + p.bins * p.D : s.cre;
    return s;
}
inline size_t stft_frame_smem_bytes(int n_fft, int win, int D, bool bwd) {
    return (size_t)(3 * n_fft + 2 * win * D + (bwd ? 2 * (n_fft / 2 + 1) * D : 0)) * sizeof(float);
}
__device__ __forceinline__ void stft_frame_stage(const StftParams& p, const StftFrameSmem& s, int b, int m) {
    stft_tables(s.ct, s.st, s.wv, p.n_fft, p.win);
    __syncthreads();
    const int woff = (p.n_fft - p.win) / 2;
    for (int i = threadIdx.x; i < p.win * p.D; i += blockDim.x) {
        const int n = i / p.D, d = i - n * p.D;
        const int pos = reflect_idx(m * p.hop + woff + n - p.n_fft / 2, p.T);
        const size_t row = (size_t)b * p.T + pos;
        s.xw[i] = s.wv[n] * p.x[row * p.ldx + d];
        s.yw[i] = s.wv[n] * p.y[row * p.ldy + d];
    }
    __syncthreads();
}
__device__ __forceinline__ void stft_frame_bin(const StftParams& p, const StftFrameSmem& s, const float* sw, int d, int bin,
                                               float& re, float& im) {
    const int woff = (p.n_fft - p.win) / 2;
    re = 0.f; im = 0.f;
    int ph = (bin * woff) % p.n_fft;
    for (int n = 0; n < p.win; ++n) {
        const float a = sw[n * p.D + d];
        re = fmaf(a, s.ct[ph], re);
        im = fmaf(-a, s.st[ph], im);
        ph += bin;
        if (ph >= p.n_fft) ph -= p.n_fft;
    }
}

__global__ void __launch_bounds__(CRK_THREADS) k_stft_loss_frame(const StftParams p, float* __restrict__ part) {
    extern __shared__ float4 crk_smem4[];
    const StftFrameSmem s = stft_frame_smem(reinterpret_cast<float*>(crk_smem4), p, false);
    __shared__ float red[8];
    const int b = blockIdx.x / p.M, m = blockIdx.x - b * p.M;
    stft_frame_stage(p, s, b, m);
    float sm = 0.f, slg = 0.f;
    for (int e = threadIdx.x; e < p.bins * p.D; e += CRK_THREADS) {
        const int bin = e / p.D, d = e - bin * p.D;
        float rx, ix, ry, iy;
        stft_frame_bin(p, s, s.xw, d, bin, rx, ix);
        stft_frame_bin(p, s, s.yw, d, bin, ry, iy);
        const float mx = sqrtf(fmaxf(fmaf(rx, rx, ix * ix), 1e-7f));
        const float my = sqrtf(fmaxf(fmaf(ry, ry, iy * iy), 1e-7f));
        sm += fabsf(mx - my);
        slg += fabsf(logf(mx) - logf(my));
    }
    sm = block_sum_256(sm, red);
    slg = block_sum_256(slg, red);
    if (threadIdx.x == 0) { part[blockIdx.x * 2] = sm; part[blockIdx.x * 2 + 1] = slg; }
}

__global__ void __launch_bounds__(CRK_THREADS) k_stft_loss_bwd_frame(const StftParams p, const float* __restrict__ g,
                                                                     const float* __restrict__ glog, float scale,
                                                                     float* __restrict__ dx, int lddx) {
    extern __shared__ float4 crk_smem4[];
    const StftFrameSmem s = stft_frame_smem(reinterpret_cast<float*>(crk_smem4), p, true);
    const int b = blockIdx.x / p.M, m = blockIdx.x - b * p.M;
    stft_frame_stage(p, s, b, m);
    const float inv_n = scale / (float)((long long)p.B * p.D * p.M * p.bins);
    const float gg = g ? g[0] * inv_n : 0.f;
    const float gl = glog ? glog[0] * inv_n : 0.f;
    for (int e = threadIdx.x; e < p.bins * p.D; e += CRK_THREADS) {
        const int bin = e / p.D, d = e - bin * p.D;
        float rx, ix, ry, iy;
        stft_frame_bin(p, s, s.xw, d, bin, rx, ix);
        stft_frame_bin(p, s, s.yw, d, bin, ry, iy);
        const float px = fmaf(rx, rx, ix * ix);
        const float mx = sqrtf(fmaxf(px, 1e-7f));
        const float my = sqrtf(fmaxf(fmaf(ry, ry, iy * iy), 1e-7f));
        float coef = 0.f;
        if (px >= 1e-7f) {
            const float df = mx - my;
            const float sgn = df > 0.f ? 1.f : (df < 0.f ? -1.f : 0.f);
            coef = sgn * (gg / mx);
            if (glog) coef += sgn * (gl / (mx * mx));
        }
        s.cre[e] = coef * rx;
        s.cim[e] = coef * ix;
    }
    __syncthreads();
    const int woff = (p.n_fft - p.win) / 2;
    for (int i = threadIdx.x; i < p.win * p.D; i += CRK_THREADS) {
        const int n = i / p.D, d = i - n * p.D;
        float acc = 0.f;
        int ph = 0;
        const int stepph = (woff + n) % p.n_fft;
        for (int bin = 0; bin < p.bins; ++bin) {
            acc += s.cre[bin * p.D + d] * s.ct[ph] - s.cim[bin * p.D + d] * s.st[ph];
            ph += stepph;
            if (ph >= p.n_fft) ph -= p.n_fft;
        }
        const int pos = reflect_idx(m * p.hop + woff + n - p.n_fft / 2, p.T);
        atomicAdd(dx + ((size_t)b * p.T + pos) * lddx + d, acc * s.wv[n]);
    }
}

__global__ void k_scale2(float* out, float inv) { out[0] *= inv; out[1] *= inv; }

__global__ void k_zero_panel(float* dx, int ld, int D, long long rows) {
    const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= rows * D) return;
    dx[(e / D) * ld + (e % D)] = 0.f;
}

// one warp per (b, d, m) frame: lanes over bins build the per-bin coefficient, then lanes over
// window samples scatter-add into dx (atomics: frames may overlap / reflect onto the same sample).
// g / glog: upstream gradients of the magnitude and the log-magnitude means (either may be null):
//   d|mx - my|/dre = sgn * re/mx,   d|log mx - log my|/dre = sgn * re/mx^2   (same sign: log is monotonic)
__global__ void __launch_bounds__(CRK_THREADS) k_stft_loss_bwd(const StftParams p, const float* __restrict__ g,
                                                               const float* __restrict__ glog,
                                                               float scale, float* __restrict__ dx, int lddx) {
    extern __shared__ float4 crk_smem4[];
    float* ct = reinterpret_cast<float*>(crk_smem4);
    float* st = ct + p.n_fft;
    float* wv = st + p.n_fft;
    float* cre = wv + p.n_fft;              // [8 warps][bins]
    float* cim = cre + 8 * p.bins;
    stft_tables(ct, st, wv, p.n_fft, p.win);
    __syncthreads();
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const long long nfr = (long long)p.B * p.D * p.M;
    const float inv_n = scale / (float)((long long)p.B * p.D * p.M * p.bins);
    const float gg = g ? g[0] * inv_n : 0.f;
    const float gl = glog ? glog[0] * inv_n : 0.f;
    const int woff = (p.n_fft - p.win) / 2;
    for (long long fr = (long long)blockIdx.x * 8 + w; fr < nfr; fr += (long long)gridDim.x * 8) {
        const int d = (int)(fr % p.D);
        long long r = fr / p.D;
        const int m = (int)(r % p.M);
        const int b = (int)(r / p.M);
        for (int bin = lane; bin < p.bins; bin += 32) {
            float rx, ix, ry, iy;
            stft_bin(p, p.x, p.ldx, b, d, m, bin, ct, st, wv, rx, ix);
            stft_bin(p, p.y, p.ldy, b, d, m, bin, ct, st, wv, ry, iy);
            const float px = fmaf(rx, rx, ix * ix);
            const float mx = sqrtf(fmaxf(px, 1e-7f));
            const float my = sqrtf(fmaxf(fmaf(ry, ry, iy * iy), 1e-7f));
            float coef = 0.f;
            if (px >= 1e-7f) {
                const float df = mx - my;
                const float sgn = df > 0.f ? 1.f : (df < 0.f ? -1.f : 0.f);
                coef = sgn * (gg / mx);
                if (glog) coef += sgn * (gl / (mx * mx));
            }
            cre[w * p.bins + bin] = coef * rx;
            cim[w * p.bins + bin] = coef * ix;
        }
        __syncwarp();
        for (int n = lane; n < p.win; n += 32) {
            float acc = 0.f;
            for (int bin = 0; bin < p.bins; ++bin) {
                const int ph = (bin * (woff + n)) % p.n_fft;
                // re = sum a cos, im = -sum a sin  =>  d/da = cre*cos - cim*sin
                acc += cre[w * p.bins + bin] * ct[ph] - cim[w * p.bins + bin] * st[ph];
            }
            const int pos = reflect_idx(m * p.hop + woff + n - p.n_fft / 2, p.T);
            atomicAdd(dx + ((size_t)b * p.T + pos) * lddx + d, acc * wv[n]);
        }
        __syncwarp();
    }
}

// ---- Adam (torch.optim.Adam single-tensor formula) ------------------------------------------
__global__ void k_adam(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                       float* __restrict__ v, long long n, float lr, float beta1, float beta2, float eps,
                       float step_size, float bc2_sqrt) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float gi = g[i];
    const float mi = m[i] + (gi - m[i]) * (1.f - beta1);          // exp_avg.lerp_(grad, 1-beta1)
    const float vi = fmaf(gi * gi, 1.f - beta2, v[i] * beta2);    // exp_avg_sq.mul_(b2).addcmul_(g,g,1-b2)
    m[i] = mi; v[i] = vi;
    const float denom = sqrtf(vi) / bc2_sqrt + eps;
    p[i] = p[i] - step_size * (mi / denom);
}


// Graph-capturable variant: the step count lives in device memory (a captured launch cannot carry a host value
// that changes every replay).  k_step_inc runs first; every block of k_adam_dev then derives the same bias
// corrections from the incremented counter, in double like the host path of crk_adam_step.
__global__ void k_step_inc(long long* step) { step[0] += 1; }
__global__ void k_adam_dev(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                           float* __restrict__ v, long long n, float lr, float beta1, float beta2, float eps,
                           const long long* __restrict__ step) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double st = (double)step[0];
    const double bc1 = 1.0 - pow((double)beta1, st);
    const double bc2 = 1.0 - pow((double)beta2, st);
    const float step_size = (float)((double)lr / bc1);
    const float bc2_sqrt = (float)sqrt(bc2);
    const float gi = g[i];
    const float mi = m[i] + (gi - m[i]) * (1.f - beta1);
    const float vi = fmaf(gi * gi, 1.f - beta2, v[i] * beta2);
    m[i] = mi; v[i] = vi;
    const float denom = sqrtf(vi) / bc2_sqrt + eps;
    p[i] = p[i] - step_size * (mi / denom);
}

// ---- RAdam (torch_optimizer.RAdam as constructed at crank/net/trainer/utils.py:44-45: lr, betas (0.9, 0.999), eps 1e-8,
// no weight decay).  The rectification term and bias corrections are host scalars (python floats in the original):
//   rect != 0:  p -= step_size * m / (sqrt(v) + eps)        rect == 0 (variance not yet tractable, N_sma < 5):  p -= step_size * m
__global__ void k_radam(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
                        long long n, float beta1, float beta2, float eps, float step_size, int rect) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float gi = g[i];
    const float vi = __fadd_rn(__fmul_rn(v[i], beta2), __fmul_rn(__fmul_rn(1.f - beta2, gi), gi));   // mul_(b2).addcmul_(g, g, 1-b2)
    const float mi = __fadd_rn(__fmul_rn(m[i], beta1), __fmul_rn(1.f - beta1, gi));                    // mul_(b1).add_(g, 1-b1)
    m[i] = mi; v[i] = vi;
    p[i] = rect ? p[i] - step_size * (mi / (sqrtf(vi) + eps)) : p[i] - step_size * mi;
}

// ---- LAMB (pytorch_lamb.Lamb as constructed at crank/net/trainer/utils.py:46-47: lr, betas (0.9, 0.999), eps 1e-6, no
// weight decay, no bias correction).  The trust ratio is per PARAMETER TENSOR of the reference (weight_g / weight_v / bias
// of every conv): one segment of the flat parameter pack each.  Pass 1 (one CTA per segment): moments, the Adam direction
// u = m / (sqrt(v) + eps) into `upd`, ||p|| (clamped to [0, 10]) and ||u|| -> trust[seg];  pass 2: p -= lr * trust * u.
__global__ void __launch_bounds__(256) k_lamb_moments(const float* __restrict__ p, const float* __restrict__ g,
                                                      float* __restrict__ m, float* __restrict__ v, float* __restrict__ upd,
                                                      const long long* __restrict__ seg_off, const long long* __restrict__ seg_len,
                                                      float* __restrict__ trust, float beta1, float beta2, float eps) {
    __shared__ double red[2][256];
    const long long o = seg_off[blockIdx.x], n = seg_len[blockIdx.x];
    double sp = 0.0, su = 0.0;
    for (long long i = threadIdx.x; i < n; i += 256) {
        const float gi = g[o + i];
        const float mi = __fadd_rn(__fmul_rn(m[o + i], beta1), __fmul_rn(1.f - beta1, gi));
        const float vi = __fadd_rn(__fmul_rn(v[o + i], beta2), __fmul_rn(__fmul_rn(1.f - beta2, gi), gi));
        m[o + i] = mi; v[o + i] = vi;
        const float u = mi / (sqrtf(vi) + eps);
        upd[o + i] = u;
        const float pi = p[o + i];
        sp += (double)pi * pi;
        su += (double)u * u;
    }
    red[0][threadIdx.x] = sp; red[1][threadIdx.x] = su;
    __syncthreads();
    for (int s = 128; s > 0; s >>= 1) {
        if (threadIdx.x < s) { red[0][threadIdx.x] += red[0][threadIdx.x + s]; red[1][threadIdx.x] += red[1][threadIdx.x + s]; }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        const float wn = fminf(fmaxf((float)sqrt(red[0][0]), 0.f), 10.f);
        const float an = (float)sqrt(red[1][0]);
        trust[blockIdx.x] = (wn == 0.f || an == 0.f) ? 1.f : wn / an;
    }
}
__global__ void __launch_bounds__(256) k_lamb_apply(float* __restrict__ p, const float* __restrict__ upd,
                                                    const long long* __restrict__ seg_off, const long long* __restrict__ seg_len,
                                                    const float* __restrict__ trust, float lr) {
    const long long o = seg_off[blockIdx.x], n = seg_len[blockIdx.x];
    const float s = lr * trust[blockIdx.x];
    for (long long i = threadIdx.x; i < n; i += 256) p[o + i] = p[o + i] - s * upd[o + i];
}

}  // namespace crk
