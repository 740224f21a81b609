// crank-b200: generic channels-last 1-D (dilated) convolution kernels, fp32 CUDA-core path.
//
//   k_conv   : Y = epilogue( bias + sum_j prologue(X)[t + j*dil - padl] . W_j )      (fwd and dgrad)
//   k_wgrad  : dW_j[ci][co] = sum_frames prologue(X)[t + j*dil - padl][ci] * G[t][co] (per frame-chunk partials)
//   k_colsum : column sums of a (F, N) panel (bias gradients), per-chunk partials
//   k_reduce : deterministic fixed-order sum of the per-chunk partials
//
// These restate, for channels-last panels, what the reference gets from `F.conv1d` + autograd
// (cuDNN/oneDNN) at every Conv1d of parallel_wavegan (call sites crank/net/module/vqvae2.py:236-273,
// crank/bin/train.py:78-115, crank/net/module/spkradv.py:49-60).
#pragma once
#include "crk_common.cuh"

namespace crk {

struct ConvParams {
    const float* X; int ldx; int Cin; int CinPad;   // CinPad = round_up(Cin,4): smem row stride
    const float* W;      // [k][CinPad][TN]   zero padded, TN = 32*CPT
    int ldw;             // 0: rows of W are TN apart; else the row stride of a wider matrix of which W is a TN-column slice
    const float* bias;   // [TN] or null
    float* Y; int ldy; int Cout;
    int B, T;
    int k, dil, padl;    // tap j reads X[t + j*dil - padl]
    // prologue on X:  v = act(pro_scale * x) * xmul
    int pro_act; float pro_slope; float pro_scale;
    const float* xmul; int ldxmul;
    // epilogue: y=acc+bias; y=act(y); y*=mul; y+=rscale*R; y*=act'(dact_src); y*=out_scale; (y+=Yold)
    int epi_act; float epi_slope;
    const float* mul_src; int ldmul;
    const float* R; int ldr; float rscale;
    const float* dact_src; int lddact; int dact_mode; float dact_slope;
    float out_scale;
    int accumulate;
};

inline ConvParams conv_params_default() {
    ConvParams p;
    p.X = nullptr; p.ldx = 0; p.Cin = 0; p.CinPad = 0; p.W = nullptr; p.ldw = 0; p.bias = nullptr;
    p.Y = nullptr; p.ldy = 0; p.Cout = 0; p.B = 0; p.T = 0; p.k = 1; p.dil = 1; p.padl = 0;
    p.pro_act = CRK_ACT_NONE; p.pro_slope = 0.f; p.pro_scale = 1.f; p.xmul = nullptr; p.ldxmul = 0;
    p.epi_act = CRK_ACT_NONE; p.epi_slope = 0.f; p.mul_src = nullptr; p.ldmul = 0;
    p.R = nullptr; p.ldr = 0; p.rscale = 1.f;
    p.dact_src = nullptr; p.lddact = 0; p.dact_mode = CRK_ACT_NONE; p.dact_slope = 0.f;
    p.out_scale = 1.f; p.accumulate = 0;
    return p;
}

// stage `rows` frames starting at time `tstart` of utterance b into smem xs[rows][CinPad],
// zero for out-of-range times / pad columns, prologue applied.
__device__ __forceinline__ void stage_x(float* xs, const float* __restrict__ X, int ldx, int Cin,
                                        int CinPad, int b, int T, int tstart, int rows, int pro_act,
                                        float pro_slope, float pro_scale,
                                        const float* __restrict__ xmul, int ldxmul) {
    const bool vec = ((ldx & 3) == 0) && ((Cin & 3) == 0) && ((reinterpret_cast<uintptr_t>(X) & 15) == 0) &&
                     (xmul == nullptr || (((ldxmul & 3) == 0) && ((reinterpret_cast<uintptr_t>(xmul) & 15) == 0)));
    if (vec) {
        const int c4n = CinPad >> 2;  // == Cin/4
        for (int idx = threadIdx.x; idx < rows * c4n; idx += CRK_THREADS) {
            const int r = idx / c4n, c4 = idx - r * c4n;
            const int tt = tstart + r;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (tt >= 0 && tt < T) {
                const size_t row = (size_t)b * T + tt;
                v = __ldg(reinterpret_cast<const float4*>(X + row * ldx) + c4);
                v.x = apply_act(v.x * pro_scale, pro_act, pro_slope);
                v.y = apply_act(v.y * pro_scale, pro_act, pro_slope);
                v.z = apply_act(v.z * pro_scale, pro_act, pro_slope);
                v.w = apply_act(v.w * pro_scale, pro_act, pro_slope);
                if (xmul) {
                    const float4 m = __ldg(reinterpret_cast<const float4*>(xmul + row * ldxmul) + c4);
                    v.x *= m.x; v.y *= m.y; v.z *= m.z; v.w *= m.w;
                }
            }
            reinterpret_cast<float4*>(xs)[idx] = v;
        }
    } else {
        for (int idx = threadIdx.x; idx < rows * CinPad; idx += CRK_THREADS) {
            const int r = idx / CinPad, c = idx - r * CinPad;
            const int tt = tstart + r;
            float v = 0.f;
            if (tt >= 0 && tt < T && c < Cin) {
                const size_t row = (size_t)b * T + tt;
                v = apply_act(__ldg(X + row * ldx + c) * pro_scale, pro_act, pro_slope);
                if (xmul) v *= __ldg(xmul + row * ldxmul + c);
            }
            xs[idx] = v;
        }
    }
}

template <int CPT>
__global__ void __launch_bounds__(CRK_THREADS) k_conv(const ConvParams p) {
    extern __shared__ float4 crk_smem4[];
    float* smem = reinterpret_cast<float*>(crk_smem4);
    constexpr int TN = 32 * CPT;
    const int tiles_per_utt = (p.T + CRK_TM - 1) / CRK_TM;
    const int b = blockIdx.x / tiles_per_utt;
    const int t0 = (blockIdx.x - b * tiles_per_utt) * CRK_TM;
    const int halo = (p.k - 1) * p.dil;
    const int rows = CRK_TM + halo;
    float* xs = smem;
    float* ws = smem + rows * p.CinPad;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;

    stage_x(xs, p.X, p.ldx, p.Cin, p.CinPad, b, p.T, t0 - p.padl, rows, p.pro_act, p.pro_slope,
            p.pro_scale, p.xmul, p.ldxmul);

    float acc[8][CPT];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int c = 0; c < CPT; ++c) acc[i][c] = 0.f;

    for (int j = 0; j < p.k; ++j) {
        if (j > 0) __syncthreads();  // everyone done with the previous tap's weights
        if (p.ldw == 0 || p.ldw == TN) {
            copy_to_smem(ws, p.W + (size_t)j * p.CinPad * TN, p.CinPad * TN);
        } else {                                             // TN-column slice of a wider packed matrix
            for (int idx = threadIdx.x; idx < p.CinPad * TN; idx += CRK_THREADS) {
                const int r = idx / TN, cc = idx - r * TN;
                ws[idx] = __ldg(p.W + ((size_t)j * p.CinPad + r) * p.ldw + cc);
            }
        }
        __syncthreads();
        tile_mac_rowA<CPT>(acc, xs + (ty * 8 + j * p.dil) * p.CinPad, p.CinPad, ws + tx * CPT, TN,
                           p.CinPad);
    }

#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int t = t0 + ty * 8 + i;
        if (t >= p.T) continue;
        const size_t row = (size_t)b * p.T + t;
#pragma unroll
        for (int c = 0; c < CPT; ++c) {
            const int co = tx * CPT + c;
            if (co >= p.Cout) continue;
            float y = acc[i][c];
            if (p.bias) y += __ldg(p.bias + co);
            y = apply_act(y, p.epi_act, p.epi_slope);
            if (p.mul_src) y *= __ldg(p.mul_src + row * p.ldmul + co);
            if (p.R) y += p.rscale * __ldg(p.R + row * p.ldr + co);
            if (p.dact_src) y *= act_grad(__ldg(p.dact_src + row * p.lddact + co), p.dact_mode, p.dact_slope);
            y *= p.out_scale;
            float* dst = p.Y + row * p.ldy + co;
            if (p.accumulate) y += *dst;
            *dst = y;
        }
    }
}

inline size_t conv_smem_bytes(const ConvParams& p, int cpt) {
    const int rows = CRK_TM + (p.k - 1) * p.dil;
    return (size_t)(rows * p.CinPad + p.CinPad * 32 * cpt) * sizeof(float);
}

template <int CPT>
inline cudaError_t launch_conv_t(const ConvParams& p, cudaStream_t s) {
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(k_conv<CPT>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
        if (e != cudaSuccess) return e;
        attr_set = true;
    }
    const int tiles = p.B * cdiv(p.T, CRK_TM);
    TimedLaunch tl(CRK_K_CONV, s, 2.0 * p.B * p.T * p.Cin * p.Cout * p.k);
    k_conv<CPT><<<tiles, CRK_THREADS, conv_smem_bytes(p, CPT), s>>>(p);
    return launch_check();
}

// `cpt` must match the packing of p.W / p.bias (TN = 32*cpt)
inline cudaError_t launch_conv(const ConvParams& p, int cpt, cudaStream_t s) {
    switch (cpt) {
        case 1: return launch_conv_t<1>(p, s);
        case 2: return launch_conv_t<2>(p, s);
        case 3: return launch_conv_t<3>(p, s);
        default: return launch_conv_t<4>(p, s);
    }
}

// ---------------------------------------------------------------------------------------------
// wgrad
struct WgradParams {
    const float* X; int ldx; int Cin; int Rows;   // Rows = round_up(Cin,4): rows of dW per tap
    int pro_act; float pro_slope; float pro_scale;
    const float* xmul; int ldxmul;
    const float* G; int ldg; int N;                // N real columns of G (<= TN)
    float* part;                                   // [nchunk][part_stride]: per chunk [k][Rows][TN] (+ bias partial)
    long long part_stride;
    int B, T, k, dil, padl;
    int tiles_per_chunk;
};

// grid = (nchunk, k, row blocks of 64)
template <int CPT>
__global__ void __launch_bounds__(CRK_THREADS) k_wgrad(const WgradParams p) {
    extern __shared__ float4 crk_smem4[];
    float* smem = reinterpret_cast<float*>(crk_smem4);
    constexpr int TN = 32 * CPT;
    float* xs = smem;                  // [TM][64]
    float* gs = smem + CRK_TM * 64;    // [TM][TN]
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const int j = blockIdx.y;
    const int rb = blockIdx.z * 64;
    const int tiles_per_utt = (p.T + CRK_TM - 1) / CRK_TM;
    const int ntiles = p.B * tiles_per_utt;
    const int tile_beg = blockIdx.x * p.tiles_per_chunk;
    const int tile_end = min(ntiles, tile_beg + p.tiles_per_chunk);

    float acc[8][CPT];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int c = 0; c < CPT; ++c) acc[i][c] = 0.f;

    for (int tile = tile_beg; tile < tile_end; ++tile) {
        const int b = tile / tiles_per_utt;
        const int t0 = (tile - b * tiles_per_utt) * CRK_TM;
        if (tile > tile_beg) __syncthreads();
        // xs[f][ci] = X[b, t0+f + j*dil - padl, rb+ci]
        for (int idx = threadIdx.x; idx < CRK_TM * 64; idx += CRK_THREADS) {
            const int f = idx >> 6, ci = idx & 63;
            const int tt = t0 + f + j * p.dil - p.padl;
            const int c = rb + ci;
            float v = 0.f;
            if (t0 + f < p.T && tt >= 0 && tt < p.T && c < p.Cin) {
                const size_t row = (size_t)b * p.T + tt;
                v = apply_act(__ldg(p.X + row * p.ldx + c) * p.pro_scale, p.pro_act, p.pro_slope);
                if (p.xmul) v *= __ldg(p.xmul + row * p.ldxmul + c);
            }
            xs[idx] = v;
        }
        for (int idx = threadIdx.x; idx < CRK_TM * TN; idx += CRK_THREADS) {
            const int f = idx / TN, n = idx - f * TN;
            float v = 0.f;
            if (t0 + f < p.T && n < p.N) v = __ldg(p.G + ((size_t)b * p.T + t0 + f) * p.ldg + n);
            gs[idx] = v;
        }
        __syncthreads();
        tile_mac_colA<CPT>(acc, xs + ty * 8, 64, gs + tx * CPT, TN, CRK_TM);
    }
    float* out = p.part + (size_t)blockIdx.x * p.part_stride + (size_t)j * p.Rows * TN;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int r = rb + ty * 8 + i;
        if (r >= p.Rows) continue;
#pragma unroll
        for (int c = 0; c < CPT; ++c) out[(size_t)r * TN + tx * CPT + c] = acc[i][c];
    }
}

// out[e] (+)= sum_{chunk} part[chunk][e]   (fixed order => deterministic).  block = (32 x 8): each
// thread owns 4 consecutive outputs (float4) and every 8th chunk, so 8x more loads are in flight than
// a one-thread-per-output loop; the 8 partial sums are combined through shared memory in fixed order.
__global__ void __launch_bounds__(256) k_reduce(const float* __restrict__ part, int nchunk, int n,
                                                float* __restrict__ out, int accumulate) {
    __shared__ float4 red[8][32];
    pdl_trigger();
    pdl_wait();
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const int e = (blockIdx.x * 32 + tx) * 4;
    float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
    if (e < n) {
        if (((n & 3) == 0) && ((reinterpret_cast<uintptr_t>(part) & 15) == 0)) {
            for (int c = ty; c < nchunk; c += 8) {
                const float4 v = *reinterpret_cast<const float4*>(part + (size_t)c * n + e);
                s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
            }
        } else {
            for (int c = ty; c < nchunk; c += 8) {
                const float* q = part + (size_t)c * n + e;
                s.x += q[0];
                if (e + 1 < n) s.y += q[1];
                if (e + 2 < n) s.z += q[2];
                if (e + 3 < n) s.w += q[3];
            }
        }
    }
    red[ty][tx] = s;
    __syncthreads();
    if (ty == 0 && e < n) {
        float4 t = red[0][tx];
#pragma unroll
        for (int i = 1; i < 8; ++i) { t.x += red[i][tx].x; t.y += red[i][tx].y; t.z += red[i][tx].z; t.w += red[i][tx].w; }
        float r4[4] = {t.x, t.y, t.z, t.w};
        for (int i = 0; i < 4 && e + i < n; ++i) out[e + i] = accumulate ? out[e + i] + r4[i] : r4[i];
    }
}

// k_reduce for partial blocks whose weight part is TRANSPOSED, [k][TN][Rows] (+ [TN] bias partial; see WgradTcParams::tpart):
// same fixed-order sums, read in the partials' own order (coalesced), written to out in the [k][Rows][TN] (+ [TN]) layout.
// Rows % 4 == 0 and n % 4 == 0 (host-checked), so the 4 consecutive sources of a thread are 4 consecutive ci of one (j, co).
__global__ void __launch_bounds__(256) k_reduce_t(const float* __restrict__ part, int nchunk, int n, float* __restrict__ out,
                                                  int k, int Rows, int TN) {
    __shared__ float4 red[8][32];
    pdl_trigger();
    pdl_wait();
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const int e = (blockIdx.x * 32 + tx) * 4;
    float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
    if (e < n)
        for (int c = ty; c < nchunk; c += 8) {
            const float4 v = *reinterpret_cast<const float4*>(part + (size_t)c * n + e);
            s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
        }
    red[ty][tx] = s;
    __syncthreads();
    if (ty == 0 && e < n) {
        float4 t = red[0][tx];
#pragma unroll
        for (int i = 1; i < 8; ++i) { t.x += red[i][tx].x; t.y += red[i][tx].y; t.z += red[i][tx].z; t.w += red[i][tx].w; }
        const float r4[4] = {t.x, t.y, t.z, t.w};
        const int nW = k * Rows * TN;
        if (e < nW) {
            const int jc = e / Rows, ci = e - jc * Rows;          // (j * TN + co), first of 4 consecutive ci
            const int j = jc / TN, co = jc - j * TN;
#pragma unroll
            for (int i = 0; i < 4; ++i) out[((size_t)j * Rows + ci + i) * TN + co] = r4[i];
        } else {
#pragma unroll
            for (int i = 0; i < 4; ++i)
                if (e + i < n) out[e + i] = r4[i];
        }
    }
}

// per-chunk column sums (bias gradients): part[chunk * stride + n].  block = 8 row groups x 128 columns
// (1024 threads), 4 independent accumulators per thread: 32 rows in flight per column; the 8 group
// sums are combined through shared memory in fixed order (deterministic).
#define CRK_COLSUM_THREADS 1024
__global__ void __launch_bounds__(CRK_COLSUM_THREADS) k_colsum(const float* __restrict__ G, int ldg, int N,
                                                                long long F, int rows_per_chunk,
                                                                float* __restrict__ part, int TN, long long stride) {
    __shared__ float red[8][128];
    pdl_trigger();
    pdl_wait();
    const int n = threadIdx.x & 127, g = threadIdx.x >> 7;
    const long long beg = (long long)blockIdx.x * rows_per_chunk;
    const long long end = min(F, beg + rows_per_chunk);
    float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
    if (n < N) {
        long long r = beg + g;
        for (; r + 24 < end; r += 32) {
            s0 += __ldg(G + r * ldg + n);
            s1 += __ldg(G + (r + 8) * ldg + n);
            s2 += __ldg(G + (r + 16) * ldg + n);
            s3 += __ldg(G + (r + 24) * ldg + n);
        }
        for (; r < end; r += 8) s0 += __ldg(G + r * ldg + n);
    }
    red[g][n] = (s0 + s1) + (s2 + s3);
    __syncthreads();
    if (g == 0 && n < TN) {
        float s = red[0][n];
#pragma unroll
        for (int i = 1; i < 8; ++i) s += red[i][n];
        part[(size_t)blockIdx.x * stride + n] = s;
    }
}

// Work-partition policy of wgrad: number of per-chunk partials so that the (chunk, tap, rowblock)
// grid is ~2 CTAs per SM.
struct WgradWork {
    int nchunk, tiles_per_chunk;
};
inline WgradWork wgrad_work(int B, int T, int k, int rows) {
    const int ntiles = B * cdiv(T, CRK_TM);
    int target = (148 * 2) / (k * cdiv(rows, 64));
    if (target < 1) target = 1;
    int nchunk = ntiles < target ? ntiles : target;
    if (nchunk < 1) nchunk = 1;
    WgradWork w;
    w.tiles_per_chunk = cdiv(ntiles, nchunk);
    w.nchunk = cdiv(ntiles, w.tiles_per_chunk);
    return w;
}
// Scratch floats conv_wgrad() needs for a conv of k taps, `rows` packed input rows, tn packed columns.
inline size_t wgrad_part_floats(int B, int T, int k, int rows, int tn) {
    WgradWork w = wgrad_work(B, T, k, rows);
    // per chunk: [k][rows][tn] weight-gradient partial followed by a [tn] bias-gradient partial
    size_t a = (size_t)w.nchunk * ((size_t)k * rows * tn + tn);
    // tensor-core wgrad policy (crk_wgrad_tc.cuh): up to 148 chunks of 64-frame tiles
    const int ntiles64 = B * cdiv(T, 64);
    size_t c = (size_t)(ntiles64 < 148 ? ntiles64 : 148) * ((size_t)k * rows * tn + tn);
    return a > c ? a : c;
}

template <int CPT>
inline cudaError_t launch_wgrad_t(const WgradParams& p, dim3 grid, cudaStream_t s) {
    const size_t smem = (size_t)(CRK_TM * 64 + CRK_TM * 32 * CPT) * sizeof(float);
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(k_wgrad<CPT>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
        if (e != cudaSuccess) return e;
        attr_set = true;
    }
    TimedLaunch tl(CRK_K_WGRAD, s, 2.0 * p.B * p.T * p.Cin * p.N * p.k / (grid.z > 0 ? 1 : 1));
    k_wgrad<CPT><<<grid, CRK_THREADS, smem, s>>>(p);
    return launch_check();
}

// tensor-core path (defined in crk_wgrad_tc.cuh): returns true when it handled the partials
bool wgrad_tc_try(WgradParams& p, int TN, float* part, cudaStream_t s, int* nchunk, cudaError_t* err,
                  bool fused_bias, bool* bias_done, bool* tpart);

// dW (and db when non-null) of one convolution.  cpt gives the packing of the G columns (TN=32*cpt).
// dW and db must be adjacent (db == dW + k*Rows*TN, the packed weff layout) so that ONE deterministic
// reduction launch sums both partial sets.
inline cudaError_t conv_wgrad(WgradParams p, int cpt, float* dW, float* db, float* part,
                              cudaStream_t s) {
    const int TN = 32 * cpt;
    const int nW = p.k * p.Rows * TN;
    const bool fused_bias = db != nullptr && db == dW + nW;
    const long long stride = nW + (fused_bias ? TN : 0);     // floats per chunk in `part`
    p.part = part;
    p.part_stride = stride;
    cudaError_t e = cudaSuccess;
    int nchunk = 0;
    bool bias_done = false, tpart = false;
    if (!wgrad_tc_try(p, TN, part, s, &nchunk, &e, fused_bias, &bias_done, &tpart)) {
        const WgradWork w = wgrad_work(p.B, p.T, p.k, p.Rows);
        nchunk = w.nchunk;
        p.tiles_per_chunk = w.tiles_per_chunk;
        dim3 grid(w.nchunk, p.k, cdiv(p.Rows, 64));
        switch (cpt) {
            case 1: e = launch_wgrad_t<1>(p, grid, s); break;
            case 2: e = launch_wgrad_t<2>(p, grid, s); break;
            case 3: e = launch_wgrad_t<3>(p, grid, s); break;
            default: e = launch_wgrad_t<4>(p, grid, s); break;
        }
    }
    if (e != cudaSuccess) return e;
    const long long F = (long long)p.B * p.T;
    if (fused_bias && !bias_done) {
        const int rows_per_chunk = (int)cdivl(F, nchunk);     // chunk c may be empty: it then writes zeros
        e = launch_pdl(k_colsum, dim3(nchunk), dim3(CRK_COLSUM_THREADS), 0, s, p.G, p.ldg, p.N, F, rows_per_chunk,
                       part + nW, TN, stride);
        if (e != cudaSuccess) return e;
        e = launch_check();
        if (e != cudaSuccess) return e;
    }
    const int n = (int)stride;
    if (tpart) e = launch_pdl(k_reduce_t, dim3(cdiv(n, 128)), dim3(256), 0, s, (const float*)part, nchunk, n, dW, p.k, p.Rows, TN);
    else e = launch_pdl(k_reduce, dim3(cdiv(n, 128)), dim3(256), 0, s, (const float*)part, nchunk, n, dW, 0);
    if (e != cudaSuccess) return e;
    e = launch_check();
    if (e != cudaSuccess) return e;
    if (db && !fused_bias) {
        int nch = (int)cdivl(F, 256);
        if (nch > nchunk) nch = nchunk;
        if (nch < 1) nch = 1;
        const int rows_per_chunk = (int)cdivl(F, nch);
        e = launch_pdl(k_colsum, dim3(nch), dim3(CRK_COLSUM_THREADS), 0, s, p.G, p.ldg, p.N, F, rows_per_chunk, part, TN,
                       (long long)TN);
        if (e != cudaSuccess) return e;
        e = launch_check();
        if (e != cudaSuccess) return e;
        e = launch_pdl(k_reduce, dim3(cdiv(TN, 128)), dim3(256), 0, s, (const float*)part, nch, TN, db, 0);
        if (e != cudaSuccess) return e;
        e = launch_check();
    }
    return e;
}

}  // namespace crk
