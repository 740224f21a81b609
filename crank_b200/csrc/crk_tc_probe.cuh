// crank-b200: tcgen05 probe kernel -- a single-CTA TF32 GEMM through the exact operand layout,
// descriptors, TMEM accumulator and completion mechanism the tensor-core conv kernels use.
// It exists so that the hardware-facing assumptions are pinned by a test of their own
// (tests/test_gpu_tc.py) before / independently of the fused kernels.
//
//   mode 0 (K-major, forward/dgrad style):  D[m][n] = sum_k A[row_shift+m][k] * B[n][k]
//   mode 1 (MN-major, wgrad style):         D[m][n] = sum_f A[f][m]           * B[f][n]
// M = 128; N in {64,128}; fp32 inputs; `split`=1 runs the 3xTF32 error-compensated product.
#pragma once
#include "crk_common.cuh"
#include "crk_tc.cuh"

namespace crk {

struct TcProbeParams {
    const float* A; int lda; int rowsA;
    const float* B; int ldb; int rowsB;
    float* D;
    int N, K;          // mode 0: K = reduction (multiple of 8) ; mode 1: K = #frames (multiple of 8)
    int row_shift, mode, split, variant;   // variant bit0: swap LBO/SBO of MN-major descriptors (diagnostic)
};

// stage src[rows][cols] (row-major, ld) into chunk-major smem tiles (hi and optionally lo)
__device__ __forceinline__ void tc_stage(float* hi, float* lo, const float* __restrict__ src, int ld, int rows,
                                         int cols, int cs_floats) {
    const int c4n = cols >> 2;
    for (int idx = threadIdx.x; idx < rows * c4n; idx += blockDim.x) {
        const int r = idx / c4n, c4 = idx - r * c4n;
        const float* s = src + (size_t)r * ld + c4 * 4;
        float4 v = make_float4(s[0], s[1], s[2], s[3]);
        float4 h, l;
        tc::split_tf32(v.x, h.x, l.x); tc::split_tf32(v.y, h.y, l.y);
        tc::split_tf32(v.z, h.z, l.z); tc::split_tf32(v.w, h.w, l.w);
        *reinterpret_cast<float4*>(hi + (size_t)c4 * cs_floats + r * 4) = lo ? h : v;
        if (lo) *reinterpret_cast<float4*>(lo + (size_t)c4 * cs_floats + r * 4) = l;
    }
}

__global__ void __launch_bounds__(128) k_tc_probe(const TcProbeParams p) {
    extern __shared__ float4 crk_smem4[];
    float* smem = reinterpret_cast<float*>(crk_smem4);
    __shared__ uint64_t mbar;
    __shared__ uint32_t tmem_base;
    const int colsA = p.mode == 0 ? p.K : 128;
    const int colsB = p.mode == 0 ? p.K : p.N;
    const int csA = tc::chunk_stride_bytes(p.rowsA) / 4, csB = tc::chunk_stride_bytes(p.rowsB) / 4;
    float* a_hi = smem;
    float* a_lo = a_hi + (p.split ? (colsA / 4) * csA : 0);
    float* b_hi = a_lo + (colsA / 4) * csA;
    float* b_lo = b_hi + (p.split ? (colsB / 4) * csB : 0);
    const int warp = threadIdx.x >> 5;

    tc_stage(a_hi, p.split ? a_lo : nullptr, p.A, p.lda, p.rowsA, colsA, csA);
    tc_stage(b_hi, p.split ? b_lo : nullptr, p.B, p.ldb, p.rowsB, colsB, csB);
    if (threadIdx.x == 0) { tc::mbar_init(&mbar, 1); tc::fence_mbar_init(); }
    if (warp == 0) tc::tmem_alloc<128>(&tmem_base);
    tc::fence_proxy_async_smem();
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    const uint32_t tmem = tmem_base;

    if (threadIdx.x == 0) {
        const uint32_t idesc = tc::make_idesc_tf32(128, p.N, p.mode, p.mode);
        const uint32_t a_hi_s = tc::smem_u32(a_hi), a_lo_s = tc::smem_u32(a_lo);
        const uint32_t b_hi_s = tc::smem_u32(b_hi), b_lo_s = tc::smem_u32(b_lo);
        const uint32_t csAb = csA * 4, csBb = csB * 4;
        uint32_t acc = 0;
        const int npass = p.split ? 3 : 1;
        for (int pass = 0; pass < npass; ++pass) {
            // small terms first: lo*hi, hi*lo, then hi*hi
            const uint32_t as = (p.split && pass == 0) ? a_lo_s : a_hi_s;
            const uint32_t bs = (p.split && pass == 1) ? b_lo_s : b_hi_s;
            for (int k0 = 0; k0 < p.K; k0 += 8) {
                uint64_t da, db;
                if (p.mode == 0) {
                    // K-major: rows = M/N index; the 2 K-chunks of this MMA are CS apart (LBO), 8-row groups 128 B (SBO)
                    da = tc::make_smem_desc(as + (k0 / 4) * csAb + p.row_shift * 16, csAb, 128);
                    db = tc::make_smem_desc(bs + (k0 / 4) * csBb, csBb, 128);
                } else {
                    // MN-major: rows = K index (frames, 8 per MMA, 16 B apart); M/N chunks of 4 are CS apart (SBO)
                    if (p.variant & 1) {
                        da = tc::make_smem_desc(as + k0 * 16, csAb, 128);
                        db = tc::make_smem_desc(bs + k0 * 16, csBb, 128);
                    } else {
                        da = tc::make_smem_desc(as + k0 * 16, 128, csAb);
                        db = tc::make_smem_desc(bs + k0 * 16, 128, csBb);
                    }
                }
                tc::umma_tf32(tmem, da, db, idesc, acc);
                acc = 1;
            }
        }
        tc::umma_commit(&mbar);
    }
    const bool arrived = tc::mbar_wait(&mbar, 0);
    tc::tc_fence_after();
    // warp w reads TMEM lanes [32w, 32w+32): thread = one output row
    const int row = warp * 32 + (threadIdx.x & 31);
    for (int c0 = 0; c0 < p.N; c0 += 32) {
        float v[32];
        tc::tmem_ld32(tmem + ((uint32_t)(warp * 32) << 16) + c0, v);
#pragma unroll
        for (int i = 0; i < 32; ++i) p.D[(size_t)row * p.N + c0 + i] = arrived ? v[i] : __int_as_float(0x7fc00000);
    }
    tc::tc_fence_before();
    __syncthreads();
    if (warp == 0) tc::tmem_dealloc<128>(tmem);
}


// ---- MMA-rate microbenchmark ---------------------------------------------------------------------
// Every CTA issues `reps` groups of (npass x K/8) 128 x N x 8 TF32 MMAs on zero-filled operand tiles laid
// out exactly like the conv kernels' (A: 136+1 rows x K channels hi|lo, B: N rows x K hi|lo) through
// tc_issue_kmajor, commits, and waits; cycles[blockIdx] = clock64 span of thread 0 from the first issue to the
// completion barrier.  Answers "what does one MMA of this shape cost in the pipeline" (execution, not issue:
// the issuing thread runs ahead of the tensor pipe), per N, with grid = 1 or one CTA per SM.
template <bool SPLIT> __device__ __forceinline__ void tc_issue_kmajor(uint32_t, uint32_t, uint32_t, uint32_t, int, uint32_t, uint32_t, uint32_t, int, uint32_t, uint32_t&);
template <bool SPLIT> __device__ __forceinline__ void tc_issue_kmajor_w(uint32_t, uint32_t, uint32_t, uint32_t, int, uint32_t, uint32_t, uint32_t, int, uint32_t, uint32_t&);
template <bool SPLIT> __device__ __forceinline__ void tc_issue_ts(uint32_t, uint32_t, uint32_t, uint32_t, uint32_t, uint32_t, int, uint32_t, uint32_t&);

__global__ void __launch_bounds__(128) k_tc_mma_rate(int N, int K, int reps, int split, long long* __restrict__ cycles) {
    extern __shared__ float4 crk_smem4[];
    float* smem = reinterpret_cast<float*>(crk_smem4);
    __shared__ uint64_t mbar;
    __shared__ uint32_t tmem_base;
    const int rowsA = 137, csA = rowsA * 4, csB = tc::chunk_rows(N) * 4;
    const int kch = K >> 2;
    float* a_hi = smem;
    float* a_lo = a_hi + kch * csA;
    float* b_hi = a_lo + kch * csA;
    float* b_lo = b_hi + kch * csB;
    const int total = 2 * kch * (csA + csB);
    // reps < 0: pseudo-random operand data instead of zeros (is the MMA cost data dependent?)
    const bool rnd = reps < 0;
    if (rnd) reps = -reps;
    for (int i = threadIdx.x; i < total; i += blockDim.x)
        smem[i] = rnd ? (float)((i * 2654435761u) >> 8) * (1.0f / 8388608.0f) - 1.0f : 0.f;
    if (threadIdx.x == 0) { tc::mbar_init(&mbar, 1); tc::fence_mbar_init(); }
    if ((threadIdx.x >> 5) == 0) tc::tmem_alloc<512>(&tmem_base);      // D: columns 0..255 (N <= 256), A (TS form): 256..383
    tc::fence_proxy_async_smem();
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    const uint32_t tmem = tmem_base;
    const int a_tmem = (split >> 1) & 1;           // bit 1 of `split`: A operand from tensor memory (columns 256..)
    const int commit_each = (split >> 2) & 1;      // bit 2: tcgen05.commit to a scratch mbarrier after every group
    const int warp_issue = (split >> 3) & 1;       // bit 3: warp-collective issue (uniform datapath), elected lane
    split &= 1;
    __shared__ uint64_t scratch_bar;
    if (threadIdx.x == 0) tc::mbar_init(&scratch_bar, 1);
    if (a_tmem) {                                  // zero A region: 128 lanes x 128 columns
        float z[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) z[i] = 0.f;
        for (int c = 0; c < 128; c += 32) tc::tmem_st32(tmem + ((uint32_t)((threadIdx.x >> 5) * 32) << 16) + 256 + c, z);
        tc::tmem_st_wait();
        tc::tc_fence_before();
        __syncthreads();
        tc::tc_fence_after();
    }
    long long c0 = 0;
    if (warp_issue) {
        if ((threadIdx.x >> 5) == 0) {
            const uint32_t idesc = tc::make_idesc_tf32(128, N, 0, 0);
            uint32_t acc = 0;
            c0 = clock64();
            for (int r = 0; r < reps; ++r) {
                if (split) tc_issue_kmajor_w<true>(tmem, tc::smem_u32(a_hi), tc::smem_u32(a_lo), csA * 4, r & 7, tc::smem_u32(b_hi), tc::smem_u32(b_lo), csB * 4, K, idesc, acc);
                else tc_issue_kmajor_w<false>(tmem, tc::smem_u32(a_hi), tc::smem_u32(a_lo), csA * 4, r & 7, tc::smem_u32(b_hi), tc::smem_u32(b_lo), csB * 4, K, idesc, acc);
            }
            if (tc::elect_one()) tc::umma_commit(&mbar);
            if (threadIdx.x == 0) cycles[2 * blockIdx.x + 1] = clock64() - c0;
            __syncwarp();
        }
    } else if (threadIdx.x == 0) {
        const uint32_t idesc = tc::make_idesc_tf32(128, N, 0, 0);
        uint32_t acc = 0;
        c0 = clock64();
        for (int r = 0; r < reps; ++r) {
            if (a_tmem) {
                if (split) tc_issue_ts<true>(tmem, tmem + 256, tmem + 256 + 64, tc::smem_u32(b_hi), tc::smem_u32(b_lo), csB * 4, K > 64 ? 64 : K, idesc, acc);
                else tc_issue_ts<false>(tmem, tmem + 256, tmem + 256 + 64, tc::smem_u32(b_hi), tc::smem_u32(b_lo), csB * 4, K > 64 ? 64 : K, idesc, acc);
            } else if (split) tc_issue_kmajor<true>(tmem, tc::smem_u32(a_hi), tc::smem_u32(a_lo), csA * 4, r & 7, tc::smem_u32(b_hi), tc::smem_u32(b_lo), csB * 4, K, idesc, acc);
            else tc_issue_kmajor<false>(tmem, tc::smem_u32(a_hi), tc::smem_u32(a_lo), csA * 4, r & 7, tc::smem_u32(b_hi), tc::smem_u32(b_lo), csB * 4, K, idesc, acc);
            if (commit_each) tc::umma_commit(&scratch_bar);
        }
        tc::umma_commit(&mbar);
        cycles[2 * blockIdx.x + 1] = clock64() - c0;          // issue time of the whole batch
    }
    const bool ok = tc::mbar_wait(&mbar, 0);
    tc::tc_fence_after();
    if (threadIdx.x == 0) cycles[2 * blockIdx.x] = ok ? clock64() - c0 : -1;
    tc::tc_fence_before();
    __syncthreads();
    if ((threadIdx.x >> 5) == 0) tc::tmem_dealloc<512>(tmem);
}

}  // namespace crk
