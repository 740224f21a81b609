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

__global__ void __launch_bounds__(256) k_logmel_fft1024(const LogmelParams p) {
    constexpr int N = 1024, MAGS = 520;
    extern __shared__ float4 crk_smem4[];
    float2* tw = reinterpret_cast<float2*>(crk_smem4);      // [N] exp(-2 pi i t / N)
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
        float s, c;
        sincospif(-2.0f * (float)i / (float)N, &s, &c);
        tw[i] = make_float2(c, s);
        win[i] = __ldg(p.window + i);
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
#pragma unroll
        for (int s = 0, Ns = 1; s < 5; ++s, Ns *= 4) {
            const int j = tid, k = j & (Ns - 1);
            float2 v0 = in[j], v1 = in[j + 256], v2 = in[j + 512], v3 = in[j + 768];
            const int ts = k * (256 / Ns);
            v1 = cmulf(v1, tw[ts]); v2 = cmulf(v2, tw[2 * ts]); v3 = cmulf(v3, tw[3 * ts]);
            const float2 a0 = make_float2(v0.x + v2.x, v0.y + v2.y), a1 = make_float2(v0.x - v2.x, v0.y - v2.y);
            const float2 a2 = make_float2(v1.x + v3.x, v1.y + v3.y), a3 = make_float2(v1.x - v3.x, v1.y - v3.y);
            const float2 m3 = make_float2(a3.y, -a3.x);                   // -i * a3
            const int base = ((j - k) << 2) + k;
            out[base] = make_float2(a0.x + a2.x, a0.y + a2.y);
            out[base + Ns] = make_float2(a1.x + m3.x, a1.y + m3.y);
            out[base + 2 * Ns] = make_float2(a0.x - a2.x, a0.y - a2.y);
            out[base + 3 * Ns] = make_float2(a1.x - m3.x, a1.y - m3.y);
            __syncthreads();
            float2* t = in; in = out; out = t;
        }
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

inline size_t logmel_fused_smem(int hop) {
    return (size_t)(3 * 1024 * 2 + 2 * 520 + CRK_MEL_MAXNNZ + 1024 + 1024 + (CRK_MEL_FPC - 1) * hop) * sizeof(float);
}

}  // namespace crk
