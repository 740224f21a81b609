// crank-b200: fused log-mel front end (round 2): framing + window + FFT-1024 + |.| + mel + log10 (+ scaler) in ONE
// kernel, one pass over the waveform.
//
// Reference: LogMelFilterBankLayer.forward (crank/net/module/mlfb.py:134-171: torch.stft -> sqrt(re^2+im^2) ->
// matmul(513 x 80 mel basis) -> clamp(eps) -> log10 -> optional (x - mean) / std) and the offline
// Feature._analyze_mlfb (crank/feature/feature.py:126-145).  Round 1 ran k_frame_window -> cuFFT R2C -> k_mel and
// moved ~16 KB per frame through HBM (materialised 1024-sample frames and the complex spectrum) against 832 B
// algorithmic (512 B of new samples in, 320 B out).  Here a CTA owns 16 consecutive frames of one utterance:
//   * their 1024 + 15*hop samples are read ONCE into shared memory;
//   * two real frames ride one complex FFT (z = a + i b;  A[k] = (Z[k] + conj Z[N-k]) / 2,  B[k] = (Z[k] - conj Z[N-k]) / 2i):
//     radix-4 Stockham, 5 stages, one butterfly per thread per stage, twiddle table in shared memory;
//   * the mel projection uses the STRUCTURE of the basis: every triangular Slaney filter covers a short run of bins
//     (~1 100 non-zeros of 41 040 entries), so it is a banded sum per mel channel -- 2.6 % of the dense GEMM's MACs.
//     (A tcgen05 GEMM here would spend 97 % of its work on zeros: this path is bandwidth-, not tensor-bound.)
// n_fft = 1024 only (every recipe: egs/vaevc/template/conf/default.yml:10); other sizes keep the cuFFT path.
#pragma once
#include "crk_common.cuh"

namespace crk {

#define CRK_MEL_FPC 16
#define CRK_MEL_MAXNNZ 2048

struct LogmelParams {
    const float* wav; long long n_samples;
    const float* window;
    const int* band_start; const int* band_len; const int* band_off; const float* band_w; int nnz;
    int hop, n_mels, M;
    float eps; const float* mean; const float* stdv;
    float* out;
};

__device__ __forceinline__ float2 cmulf(float2 a, float2 b) { return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }

__device__ __forceinline__ void logmel_fft_stages(float2*& in, float2*& out, const float2* tw, int tid) {
#pragma unroll
    for (int s = 0, Ns = 1; s < 5; ++s, Ns *= 4) {
        const int j = tid, k = j & (Ns - 1);
        float2 v0 = in[j], v1 = in[j + 256], v2 = in[j + 512], v3 = in[j + 768];
        const float2* twp = tw + (Ns - 1) + k;
        v1 = cmulf(v1, twp[0]); v2 = cmulf(v2, twp[Ns]); v3 = cmulf(v3, twp[2 * Ns]);
        const float2 a0 = make_float2(v0.x + v2.x, v0.y + v2.y), a1 = make_float2(v0.x - v2.x, v0.y - v2.y);
        const float2 a2 = make_float2(v1.x + v3.x, v1.y + v3.y), a3 = make_float2(v1.x - v3.x, v1.y - v3.y);
        const float2 m3 = make_float2(a3.y, -a3.x);
        const int base = ((j - k) << 2) + k;
        out[base] = make_float2(a0.x + a2.x, a0.y + a2.y);
        out[base + Ns] = make_float2(a1.x + m3.x, a1.y + m3.y);
        out[base + 2 * Ns] = make_float2(a0.x - a2.x, a0.y - a2.y);
        out[base + 3 * Ns] = make_float2(a1.x - m3.x, a1.y - m3.y);
        __syncthreads();
        float2* t = in; in = out; out = t;
    }
}

__global__ void __launch_bounds__(256) k_logmel_fft1024(const LogmelParams p) {
    constexpr int N = 1024, MAGS = 520;
    extern __shared__ float4 crk_smem4[];
    // twiddles, one conflict-free table per (stage, t): entry (Ns, t, k) = exp(-2 pi i t k / (4 Ns)) at (Ns - 1) + (t - 1) Ns + k
    // (1 023 entries).  A single exp(-2 pi i n / N) table is read at stride 256 / Ns: 16-way bank conflicts at Ns = 16
    // (ncu, round 2: 54 % of the kernel's shared-memory wavefronts were conflict replays).
    float2* tw = reinterpret_cast<float2*>(crk_smem4);
    float2* bufA = tw + N;
    float2* bufB = bufA + N;
    float* mag = reinterpret_cast<float*>(bufB + N);        // [2][MAGS]
    float* wts = mag + 2 * MAGS;                            // [CRK_MEL_MAXNNZ]
    float* win = wts + CRK_MEL_MAXNNZ;                      // [N]
    float* xs = win + N;                                    // [N + (FPC-1)*hop]
    const int tid = threadIdx.x;
    const int groups = (p.M + CRK_MEL_FPC - 1) / CRK_MEL_FPC;
    const int b = blockIdx.x / groups;
    const int m0 = (blockIdx.x - b * groups) * CRK_MEL_FPC;
    const int nfr = min(CRK_MEL_FPC, p.M - m0);
    for (int i = tid; i < N; i += 256) {
        win[i] = __ldg(p.window + i);
        if (i < N - 1) {
            const int Ns = i < 3 ? 1 : (i < 15 ? 4 : (i < 63 ? 16 : (i < 255 ? 64 : 256)));
            const int e = i - (Ns - 1), t = e / Ns + 1, k = e - (t - 1) * Ns;
            float s, c;
            sincospif(-2.0f * (float)(t * k) / (float)(4 * Ns), &s, &c);
            tw[i] = make_float2(c, s);
        }
    }
    for (int i = tid; i < p.nnz; i += 256) wts[i] = __ldg(p.band_w + i);
    const float* src = p.wav + (size_t)b * p.n_samples + (size_t)m0 * p.hop;
    const int nsmp = N + (nfr - 1) * p.hop;
    for (int i = tid; i < nsmp; i += 256) xs[i] = __ldg(src + i);
    __syncthreads();

    for (int f = 0; f < nfr; f += 2) {
        const bool two = f + 1 < nfr;
        const float* xa = xs + f * p.hop;
        const float* xb = xa + p.hop;
        for (int n = tid; n < N; n += 256) bufA[n] = make_float2(win[n] * xa[n], two ? win[n] * xb[n] : 0.f);
        __syncthreads();
        float2* in = bufA;
        float2* out = bufB;
        logmel_fft_stages(in, out, tw, tid);
        // spectrum of the pair in `in`: split into the two real frames, magnitudes of bins 0..512
        for (int k = tid; k <= N / 2; k += 256) {
            const float2 Z = in[k], Zc = in[(N - k) & (N - 1)];
            const float ra = 0.5f * (Z.x + Zc.x), ia = 0.5f * (Z.y - Zc.y);
            const float rb = 0.5f * (Z.y + Zc.y), ib = -0.5f * (Z.x - Zc.x);
            mag[k] = sqrtf(ra * ra + ia * ia);
            mag[MAGS + k] = sqrtf(rb * rb + ib * ib);
        }
        __syncthreads();
        {
            const int which = tid >> 7, m = tid & 127;
            if (m < p.n_mels && (which == 0 || two)) {
                const int st = __ldg(p.band_start + m), ln = __ldg(p.band_len + m), of = __ldg(p.band_off + m);
                const float* mg = mag + which * MAGS + st;
                float acc = 0.f;
                for (int i = 0; i < ln; ++i) acc = fmaf(mg[i], wts[of + i], acc);
                float v = log10f(fmaxf(acc, p.eps));
                if (p.mean) v = (v - __ldg(p.mean + m)) / __ldg(p.stdv + m);
                p.out[((size_t)b * p.M + m0 + f + which) * p.n_mels + m] = v;
            }
        }
        __syncthreads();
    }
}

// ---- backward of the fused front end (learnable STFT windows: crank/net/module/mlfb.py:72-90, "param" / "conv") --------
// d out / d window and d out / d wav.  Per frame pair: recompute the forward spectrum, mel sums and their log10 / scaler
// derivative, pull the gradient back through the banded mel projection (transposed band table: per bin the run of mel
// channels covering it) and through |X| (Y[k] = g_mag[k] X[k] / |X[k]|), then ONE more complex FFT returns the gradient
// of both real frames: g_frame = Re FFT(conj(Ypad)) = FFT(Hermitian part of conj(Ypad)), two Hermitian spectra per transform.
//   d window[n] = sum_frames g_frame[n] * x[n]     (per-thread register accumulators -> per-CTA partial row -> fixed-order reduce)
//   d wav[m hop + n] += g_frame[n] * window[n]      (overlap-add in shared memory, one atomicAdd per sample per CTA: a sample
//                                                   is touched by <= 2 CTAs when 17 hop >= 1024, so the sum is order-independent)
struct LogmelBwdParams {
    LogmelParams f;                    // forward tensors (f.out unused)
    const float* dout;                 // (B, M, n_mels) upstream gradient
    const int* bin_start; const int* bin_len; const int* bin_off; const float* bin_w;   // transposed bands, per bin 0..512
    float* dwin_part;                  // [grid][1024] or nullptr
    float* dwav;                       // (B, n_samples), zero-initialised by the caller, or nullptr
};

__global__ void __launch_bounds__(256) k_logmel_bwd_fft1024(const LogmelBwdParams q) {
    const LogmelParams& p = q.f;
    constexpr int N = 1024, MAGS = 520;
    extern __shared__ float4 crk_smem4[];
    float2* tw = reinterpret_cast<float2*>(crk_smem4);
    float2* bufA = tw + N;
    float2* bufB = bufA + N;
    float* mag = reinterpret_cast<float*>(bufB + N);        // [2][MAGS]
    float* wts = mag + 2 * MAGS;                            // [CRK_MEL_MAXNNZ]
    float* win = wts + CRK_MEL_MAXNNZ;                      // [N]
    float* gacc = win + N;                                  // [2][128] d loss / d mel sum
    float* xs = gacc + 256;                                 // [N + (FPC-1)*hop]
    const int tid = threadIdx.x;
    const int groups = (p.M + CRK_MEL_FPC - 1) / CRK_MEL_FPC;
    const int b = blockIdx.x / groups;
    const int m0 = (blockIdx.x - b * groups) * CRK_MEL_FPC;
    const int nfr = min(CRK_MEL_FPC, p.M - m0);
    const int nsmp = N + (nfr - 1) * p.hop;
    float* gx = xs + N + (CRK_MEL_FPC - 1) * p.hop;         // [nsmp] overlap-added d loss / d wav of this CTA's span
    for (int i = tid; i < N; i += 256) {
        win[i] = __ldg(p.window + i);
        if (i < N - 1) {
            const int Ns = i < 3 ? 1 : (i < 15 ? 4 : (i < 63 ? 16 : (i < 255 ? 64 : 256)));
            const int e = i - (Ns - 1), t = e / Ns + 1, k = e - (t - 1) * Ns;
            float s, c;
            sincospif(-2.0f * (float)(t * k) / (float)(4 * Ns), &s, &c);
            tw[i] = make_float2(c, s);
        }
    }
    for (int i = tid; i < p.nnz; i += 256) wts[i] = __ldg(p.band_w + i);
    const float* src = p.wav + (size_t)b * p.n_samples + (size_t)m0 * p.hop;
    for (int i = tid; i < nsmp; i += 256) { xs[i] = __ldg(src + i); gx[i] = 0.f; }
    __syncthreads();
    float dwin_acc[4] = {0.f, 0.f, 0.f, 0.f};

    for (int f = 0; f < nfr; f += 2) {
        const bool two = f + 1 < nfr;
        const float* xa = xs + f * p.hop;
        const float* xb = xa + p.hop;
        for (int n = tid; n < N; n += 256) bufA[n] = make_float2(win[n] * xa[n], two ? win[n] * xb[n] : 0.f);
        __syncthreads();
        float2* in = bufA;
        float2* out = bufB;
        logmel_fft_stages(in, out, tw, tid);
        for (int k = tid; k <= N / 2; k += 256) {
            const float2 Z = in[k], Zc = in[(N - k) & (N - 1)];
            const float ra = 0.5f * (Z.x + Zc.x), ia = 0.5f * (Z.y - Zc.y);
            const float rb = 0.5f * (Z.y + Zc.y), ib = -0.5f * (Z.x - Zc.x);
            mag[k] = sqrtf(ra * ra + ia * ia);
            mag[MAGS + k] = sqrtf(rb * rb + ib * ib);
        }
        __syncthreads();
        {
            const int which = tid >> 7, m = tid & 127;
            float g = 0.f;
            if (m < p.n_mels && (which == 0 || two)) {
                const int st = __ldg(p.band_start + m), ln = __ldg(p.band_len + m), of = __ldg(p.band_off + m);
                const float* mg = mag + which * MAGS + st;
                float acc = 0.f;
                for (int i = 0; i < ln; ++i) acc = fmaf(mg[i], wts[of + i], acc);
                g = __ldg(q.dout + ((size_t)b * p.M + m0 + f + which) * p.n_mels + m);
                if (p.mean) g /= __ldg(p.stdv + m);
                // d log10(max(acc, eps)) / d acc
                g = acc > p.eps ? g / (acc * 2.302585092994046f) : 0.f;
            }
            gacc[which * 128 + m] = g;
        }
        __syncthreads();
        // Y = g_mag X / |X| per frame, packed as W = Ph + i Qh (Hermitian parts of conj(Ya_pad), conj(Yb_pad)) into `out`
        for (int k = tid; k <= N / 2; k += 256) {
            const float2 Z = in[k], Zc = in[(N - k) & (N - 1)];
            const float ra = 0.5f * (Z.x + Zc.x), ia = 0.5f * (Z.y - Zc.y);
            const float rb = 0.5f * (Z.y + Zc.y), ib = -0.5f * (Z.x - Zc.x);
            const int st = __ldg(q.bin_start + k), ln = __ldg(q.bin_len + k), of = __ldg(q.bin_off + k);
            float ga = 0.f, gb = 0.f;
            for (int i = 0; i < ln; ++i) {
                const float w = __ldg(q.bin_w + of + i);
                ga = fmaf(w, gacc[st + i], ga);
                gb = fmaf(w, gacc[128 + st + i], gb);
            }
            const float ma = mag[k], mb = mag[MAGS + k];
            const float sa = ma > 0.f ? ga / ma : 0.f, sb = mb > 0.f ? gb / mb : 0.f;
            const float2 Ya = make_float2(sa * ra, sa * ia), Yb = make_float2(sb * rb, sb * ib);
            if (k == 0 || k == N / 2) {
                out[k] = make_float2(Ya.x, Yb.x);                                  // Ph = Re Ya, Qh = Re Yb (real)
            } else {
                // Ph_k = conj(Ya)/2, Qh_k = conj(Yb)/2:  W_k = Ph_k + i Qh_k
                out[k] = make_float2(0.5f * (Ya.x + Yb.y), 0.5f * (-Ya.y + Yb.x));
                // Ph_{N-k} = Ya/2, Qh_{N-k} = Yb/2
                out[N - k] = make_float2(0.5f * (Ya.x - Yb.y), 0.5f * (Ya.y + Yb.x));
            }
        }
        __syncthreads();
        { float2* t = in; in = out; out = t; }
        logmel_fft_stages(in, out, tw, tid);
        // in[n] = (g_frame_a[n], g_frame_b[n])
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int n = tid + 256 * j;
            const float2 g = in[n];
            dwin_acc[j] += g.x * xa[n] + (two ? g.y * xb[n] : 0.f);
            gx[f * p.hop + n] += g.x * win[n];
        }
        __syncthreads();
        if (two) {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int n = tid + 256 * j;
                gx[(f + 1) * p.hop + n] += in[n].y * win[n];
            }
        }
        __syncthreads();
    }
    if (q.dwin_part) {
#pragma unroll
        for (int j = 0; j < 4; ++j) q.dwin_part[(size_t)blockIdx.x * N + tid + 256 * j] = dwin_acc[j];
    }
    if (q.dwav) {
        float* dst = q.dwav + (size_t)b * p.n_samples + (size_t)m0 * p.hop;
        for (int i = tid; i < nsmp; i += 256) atomicAdd(dst + i, gx[i]);
    }
}

// d window[n] = sum over CTAs of their partial rows, fixed order
__global__ void k_logmel_dwin_reduce(const float* __restrict__ part, int rows, float* __restrict__ dwin) {
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= 1024) return;
    float s = 0.f;
    for (int r = 0; r < rows; ++r) s += part[(size_t)r * 1024 + n];
    dwin[n] = s;
}

inline size_t logmel_bwd_smem(int hop) {
    return (size_t)(3 * 1024 * 2 + 2 * 520 + CRK_MEL_MAXNNZ + 1024 + 256 + 2 * (1024 + (CRK_MEL_FPC - 1) * hop)) * sizeof(float);
}

inline size_t logmel_fused_smem(int hop) {
    return (size_t)(3 * 1024 * 2 + 2 * 520 + CRK_MEL_MAXNNZ + 1024 + 1024 + (CRK_MEL_FPC - 1) * hop) * sizeof(float);
}

}  // namespace crk
