// crank-b200: fused WaveNet gated residual block, fp32 CUDA-core path.
//
// Restates parallel_wavegan's ResidualBlock.forward as used by the reference's encoder/decoder
// stacks (crank/net/module/vqvae2.py:236-273) and residual discriminator (crank/bin/train.py:107-115):
//     x = dropout(x); x = conv_dilated(x); xa,xb = split(x); (+ aux 1x1);
//     z = tanh(xa)*sigmoid(xb); s = conv1x1_skip(z); x = (conv1x1_out(z) + residual) * sqrt(.5)
// One CTA = 64 frames of one utterance: k tap-GEMMs (64x64 . 64x128) + aux GEMM accumulate the
// 128 gate pre-activations in registers, the gate is applied in registers, z goes to shared
// memory and a second GEMM (64x64 . 64x128) produces [out | skip] -- activations make exactly one
// HBM round trip per block (read h, write h', skip+=, save tanh/sigmoid for backward).
//
// Gate-channel interleave: packed column p = 4q+r holds  r=0:a[2q] r=1:a[2q+1] r=2:b[2q] r=3:b[2q+1]
// (a = tanh half, b = sigmoid half), and [out|skip] uses the same pattern, so each thread owns both
// halves of the channels it gates and every global access is a float4/float2.
#pragma once
#include "crk_common.cuh"
#include "crk_conv.cuh"

namespace crk {

struct ResFwdParams {
    const float* Hin;       // (F,64) block input (also the residual)
    float* Hout;            // (F,64)
    float* Skip;            // (F,64) running skip sum
    int skip_init;          // 1: Skip = s (first layer), 0: Skip += s
    const float* Wc;        // [k][64][128] gate-interleaved
    const float* bc;        // [128] gate-interleaved
    const float* Caux; int ldc; int Ca; int CaPad; const float* Wa;  // aux (F,Ca), Wa [CaPad][128]
    const float* Wos;       // [64][128] out|skip interleaved
    const float* bos;       // [128]
    const float* dropmul;   // (F,64) dropout multiplier (mask/(1-p)) or null
    float* TaSb;            // (F,128) saved tanh/sigmoid (interleaved) or null (no-grad)
    int B, T, k, dil, padl;
};

__global__ void __launch_bounds__(CRK_THREADS) k_resblock_fwd(const ResFwdParams p) {
    extern __shared__ float4 crk_smem4[];
    float* smem = reinterpret_cast<float*>(crk_smem4);
    const int tiles_per_utt = (p.T + CRK_TM - 1) / CRK_TM;
    const int b = blockIdx.x / tiles_per_utt;
    const int t0 = (blockIdx.x - b * tiles_per_utt) * CRK_TM;
    const int halo = (p.k - 1) * p.dil;
    const int rows = CRK_TM + halo;
    float* xs = smem;                       // [rows][64]
    float* ws = xs + rows * 64;             // [64][128] (also holds Wa: [CaPad][128])
    float* zs = ws + 64 * 128;              // [64][64]
    float* cs = zs + 64 * 64;               // [64][CaPad]
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;

    stage_x(xs, p.Hin, 64, 64, 64, b, p.T, t0 - p.padl, rows, CRK_ACT_NONE, 0.f, 1.f, p.dropmul, 64);
    if (p.Ca > 0)
        stage_x(cs, p.Caux, p.ldc, p.Ca, p.CaPad, b, p.T, t0, CRK_TM, CRK_ACT_NONE, 0.f, 1.f, nullptr, 0);

    float acc[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int c = 0; c < 4; ++c) acc[i][c] = 0.f;

    for (int j = 0; j < p.k; ++j) {
        if (j > 0) __syncthreads();
        copy_to_smem(ws, p.Wc + (size_t)j * 64 * 128, 64 * 128);
        __syncthreads();
        tile_mac_rowA<4>(acc, xs + (ty * 8 + j * p.dil) * 64, 64, ws + tx * 4, 128, 64);
    }
    if (p.Ca > 0) {
        __syncthreads();
        copy_to_smem(ws, p.Wa, p.CaPad * 128);
        __syncthreads();
        tile_mac_rowA<4>(acc, cs + ty * 8 * p.CaPad, p.CaPad, ws + tx * 4, 128, p.CaPad);
    }

    // gate (registers) -> z tile (smem), save tanh/sigmoid for backward
    const float4 bcv = __ldg(reinterpret_cast<const float4*>(p.bc) + tx);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int r = ty * 8 + i;
        const float ta0 = tanhf(acc[i][0] + bcv.x);
        const float ta1 = tanhf(acc[i][1] + bcv.y);
        const float sb0 = 1.f / (1.f + expf(-(acc[i][2] + bcv.z)));
        const float sb1 = 1.f / (1.f + expf(-(acc[i][3] + bcv.w)));
        *reinterpret_cast<float2*>(zs + r * 64 + 2 * tx) = make_float2(ta0 * sb0, ta1 * sb1);
        const int t = t0 + r;
        if (p.TaSb && t < p.T)
            reinterpret_cast<float4*>(p.TaSb + ((size_t)b * p.T + t) * 128)[tx] = make_float4(ta0, ta1, sb0, sb1);
    }
    __syncthreads();   // all tap/aux GEMMs done with ws; zs complete
    copy_to_smem(ws, p.Wos, 64 * 128);
    __syncthreads();
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int c = 0; c < 4; ++c) acc[i][c] = 0.f;
    tile_mac_rowA<4>(acc, zs + ty * 8 * 64, 64, ws + tx * 4, 128, 64);

    const float4 bov = __ldg(reinterpret_cast<const float4*>(p.bos) + tx);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int t = t0 + ty * 8 + i;
        if (t >= p.T) continue;
        const size_t row = (size_t)b * p.T + t;
        const float2 res = __ldg(reinterpret_cast<const float2*>(p.Hin + row * 64) + tx);
        float2 ho;
        ho.x = ((acc[i][0] + bov.x) + res.x) * CRK_SQRT_HALF;
        ho.y = ((acc[i][1] + bov.y) + res.y) * CRK_SQRT_HALF;
        reinterpret_cast<float2*>(p.Hout + row * 64)[tx] = ho;
        float2 sk = make_float2(acc[i][2] + bov.z, acc[i][3] + bov.w);
        float2* sp = reinterpret_cast<float2*>(p.Skip + row * 64) + tx;
        if (!p.skip_init) {
            const float2 o = *sp;
            sk.x += o.x; sk.y += o.y;
        }
        *sp = sk;
    }
}

inline size_t resblock_fwd_smem(int k, int dil, int CaPad) {
    return (size_t)((CRK_TM + (k - 1) * dil) * 64 + 64 * 128 + 64 * 64 + 64 * CaPad) * sizeof(float);
}

inline cudaError_t launch_resblock_fwd(const ResFwdParams& p, cudaStream_t s) {
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(k_resblock_fwd, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
        if (e != cudaSuccess) return e;
        attr_set = true;
    }
    const int tiles = p.B * cdiv(p.T, CRK_TM);
    TimedLaunch tl(CRK_K_RESBLOCK_FWD, s, 2.0 * p.B * p.T * (64.0 * 128 * p.k + p.Ca * 128.0 + 64.0 * 128));
    k_resblock_fwd<<<tiles, CRK_THREADS, resblock_fwd_smem(p.k, p.dil, p.CaPad), s>>>(p);
    return launch_check();
}

// ---------------------------------------------------------------------------------------------
// backward through [out|skip] 1x1 convs and the gate:
//   gos = [sqrt(.5)*dH | dS] (interleaved)         -> global GOS (wgrad operand)
//   dz  = gos . Wos^T                               (64x128 . 128x64)
//   dxa = (dz*sb)*(1-ta^2),  dxb = (dz*ta)*((1-sb)*sb)  -> global DG (interleaved gate order)
//   z   = ta*sb                                     -> global Z (wgrad operand)
struct ResBwdGateParams {
    const float* dH;     // (F,64) grad wrt block output, or null (last block: output unused)
    const float* dS;     // (F,64) grad wrt skip output
    const float* TaSb;   // (F,128)
    const float* WosT;   // [128][64]
    float* DG;           // (F,128)
    float* GOS;          // (F,128)
    float* Z;            // (F,64)
    int B, T;
};

__global__ void __launch_bounds__(CRK_THREADS) k_resblock_bwd_gate(const ResBwdGateParams p) {
    extern __shared__ float4 crk_smem4[];
    float* smem = reinterpret_cast<float*>(crk_smem4);
    float* gs = smem;              // [64][128]
    float* ws = smem + 64 * 128;   // [128][64]
    const int tiles_per_utt = (p.T + CRK_TM - 1) / CRK_TM;
    const int b = blockIdx.x / tiles_per_utt;
    const int t0 = (blockIdx.x - b * tiles_per_utt) * CRK_TM;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;

    copy_to_smem(ws, p.WosT, 128 * 64);
    for (int idx = threadIdx.x; idx < CRK_TM * 32; idx += CRK_THREADS) {
        const int f = idx >> 5, q = idx & 31;
        const int t = t0 + f;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (t < p.T) {
            const size_t row = (size_t)b * p.T + t;
            if (p.dH) {
                const float2 h = __ldg(reinterpret_cast<const float2*>(p.dH + row * 64) + q);
                v.x = h.x * CRK_SQRT_HALF; v.y = h.y * CRK_SQRT_HALF;
            }
            const float2 sgr = __ldg(reinterpret_cast<const float2*>(p.dS + row * 64) + q);
            v.z = sgr.x; v.w = sgr.y;
            reinterpret_cast<float4*>(p.GOS + row * 128)[q] = v;
        }
        reinterpret_cast<float4*>(gs + f * 128)[q] = v;
    }
    __syncthreads();
    float acc[8][2];
#pragma unroll
    for (int i = 0; i < 8; ++i) { acc[i][0] = 0.f; acc[i][1] = 0.f; }
    tile_mac_rowA<2>(acc, gs + ty * 8 * 128, 128, ws + tx * 2, 64, 128);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int t = t0 + ty * 8 + i;
        if (t >= p.T) continue;
        const size_t row = (size_t)b * p.T + t;
        const float4 ts = __ldg(reinterpret_cast<const float4*>(p.TaSb + row * 128) + tx);
        const float dz0 = acc[i][0], dz1 = acc[i][1];
        float4 dg;
        dg.x = (dz0 * ts.z) * (1.f - ts.x * ts.x);
        dg.y = (dz1 * ts.w) * (1.f - ts.y * ts.y);
        dg.z = (dz0 * ts.x) * ((1.f - ts.z) * ts.z);
        dg.w = (dz1 * ts.y) * ((1.f - ts.w) * ts.w);
        reinterpret_cast<float4*>(p.DG + row * 128)[tx] = dg;
        reinterpret_cast<float2*>(p.Z + row * 64)[tx] = make_float2(ts.x * ts.z, ts.y * ts.w);
    }
}

inline cudaError_t launch_resblock_bwd_gate(const ResBwdGateParams& p, cudaStream_t s) {
    const size_t smem = (size_t)(64 * 128 + 128 * 64) * sizeof(float);
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(k_resblock_bwd_gate, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
        if (e != cudaSuccess) return e;
        attr_set = true;
    }
    const int tiles = p.B * cdiv(p.T, CRK_TM);
    TimedLaunch tl(CRK_K_BWD_GATE, s, 2.0 * p.B * p.T * 128.0 * 64);
    k_resblock_bwd_gate<<<tiles, CRK_THREADS, smem, s>>>(p);
    return launch_check();
}

}  // namespace crk
