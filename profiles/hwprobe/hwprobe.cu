// crank-b200 round 2: hardware probe for the tcgen05 features the redesigned kernels want to rely on.
// Standalone program (nvcc -gencode arch=compute_100a,code=sm_100a); prints one line per question.
//   T1  MN-major tf32 operands in the no-swizzle chunk-major layout: which (LBO, SBO) roles / strides work
//   T2  ... with a start-row shift (conv taps along the K = frame dimension)
//   T3  TS form (A from tensor memory): numerics, unaligned start column
//   T4  does the tensor core truncate fp32 operands to tf32 (raw operand == masked-hi operand, bitwise)?
//   T5  tcgen05.ld.16x256b fragment layout
//   T6  MMA cost per shape: SS N=64/128/256, TS N=64/128, bf16 N=128 (warp-collective issue, 148 CTAs)
#include <cuda_runtime.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <vector>
#include "../../crank_b200/csrc/crk_tc.cuh"

using namespace crk;

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)

struct P {
    const float* A; const float* B; float* D;
    int rowsA_alloc, rowsB_alloc;   // chunk stride (rows) of the staged tiles
    int a_mn, b_mn;                 // 1: MN-major operand (tile rows = K index)
    int variant;                    // MN-major descriptor variant
    int shiftA, shiftB;             // start-row shift of MN-major operands
    int N, K;
    int ts, coff;                   // TS form: A in TMEM at column offset coff
    int raw;                        // no hi masking (T4)
    int split;
};

// generic stage: src[rows][cols] row-major -> chunk-major (chunk = 4 cols), chunk stride rows_alloc*16 B
__device__ void stage(float* dst, const float* src, int rows, int cols, int rows_alloc, int mode /*0 raw,1 hi,2 lo*/) {
    for (int i = threadIdx.x; i < rows * cols; i += blockDim.x) {
        const int r = i / cols, c = i % cols;
        float x = src[i], hi, lo;
        tc::split_tf32(x, hi, lo);
        dst[(c >> 2) * rows_alloc * 4 + r * 4 + (c & 3)] = mode == 0 ? x : (mode == 1 ? hi : lo);
    }
}

__global__ void __launch_bounds__(128) k_probe(P p) {
    extern __shared__ float4 sm4[];
    float* smem = reinterpret_cast<float*>(sm4);
    __shared__ uint64_t mbar;
    __shared__ uint32_t tmem_base;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // A tile: K-major: rows = 128 (M), cols = K.   MN-major: rows = K + shiftA, cols = 128
    const int rA = p.a_mn ? p.K + p.shiftA : 128, cA = p.a_mn ? 128 : p.K;
    const int rB = p.b_mn ? p.K + p.shiftB : p.N, cB = p.b_mn ? p.N : p.K;
    const int szA = (cA / 4) * p.rowsA_alloc * 4, szB = (cB / 4) * p.rowsB_alloc * 4;
    float* a_hi = smem; float* a_lo = a_hi + szA; float* b_hi = a_lo + szA; float* b_lo = b_hi + szB;
    stage(a_hi, p.A, rA, cA, p.rowsA_alloc, p.raw ? 0 : 1);
    stage(a_lo, p.A, rA, cA, p.rowsA_alloc, 2);
    stage(b_hi, p.B, rB, cB, p.rowsB_alloc, p.raw ? 0 : 1);
    stage(b_lo, p.B, rB, cB, p.rowsB_alloc, 2);
    if (threadIdx.x == 0) { tc::mbar_init(&mbar, 1); tc::fence_mbar_init(); }
    if (warp == 0) tc::tmem_alloc<512>(&tmem_base);
    tc::fence_proxy_async_smem();
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    const uint32_t tmem = tmem_base;
    if (p.ts) {
        // A (K-major source A[m][k]) -> TMEM: lane m, columns 256+coff+k (hi) and 384+coff+k (lo)
        const int m = warp * 32 + lane;
        for (int k0 = 0; k0 < p.K; k0 += 32) {
            float h[32], l[32];
            for (int i = 0; i < 32; ++i) {
                float x = (k0 + i < p.K) ? p.A[m * p.K + k0 + i] : 0.f;
                tc::split_tf32(x, h[i], l[i]);
                if (p.raw) h[i] = x;
            }
            // tcgen05.st needs an aligned column? write through a 32-column window that starts at coff
            tc::tmem_st32(tmem + ((uint32_t)(warp * 32) << 16) + 256 + p.coff + k0, h);
            tc::tmem_st32(tmem + ((uint32_t)(warp * 32) << 16) + 384 + p.coff + k0, l);
        }
        tc::tmem_st_wait();
        tc::tc_fence_before();
        __syncthreads();
        tc::tc_fence_after();
    }
    if (threadIdx.x == 0) {
        const uint32_t idesc = tc::make_idesc_tf32(128, p.N, p.a_mn, p.b_mn);
        const uint32_t csA = p.rowsA_alloc * 16, csB = p.rowsB_alloc * 16;
        uint32_t acc = 0;
        const int npass = p.split ? 3 : 1;
        for (int pass = 0; pass < npass; ++pass) {
            const uint32_t as = tc::smem_u32((p.split && pass == 0) ? a_lo : a_hi);
            const uint32_t bs = tc::smem_u32((p.split && pass == 1) ? b_lo : b_hi);
            const uint32_t ta = tmem + ((p.split && pass == 0) ? 384 : 256) + p.coff;
            for (int k0 = 0; k0 < p.K; k0 += 8) {
                uint64_t da, db;
                if (!p.a_mn) da = tc::make_smem_desc(as + (k0 / 4) * csA, csA, 128);
                else if (p.variant == 0) da = tc::make_smem_desc(as + (k0 + p.shiftA) * 16, 128, csA);
                else if (p.variant == 1) da = tc::make_smem_desc(as + (k0 + p.shiftA) * 16, csA, 128);
                else da = tc::make_smem_desc(as + (k0 + p.shiftA) * 16, 128, csA) | (1ull << 52);   // lbo_mode bit
                if (!p.b_mn) db = tc::make_smem_desc(bs + (k0 / 4) * csB, csB, 128);
                else if (p.variant == 0) db = tc::make_smem_desc(bs + (k0 + p.shiftB) * 16, 128, csB);
                else if (p.variant == 1) db = tc::make_smem_desc(bs + (k0 + p.shiftB) * 16, csB, 128);
                else db = tc::make_smem_desc(bs + (k0 + p.shiftB) * 16, 128, csB) | (1ull << 52);
                if (p.ts) tc::umma_tf32_ts(tmem, ta + k0, db, idesc, acc);
                else tc::umma_tf32(tmem, da, db, idesc, acc);
                acc = 1;
            }
        }
        tc::umma_commit(&mbar);
    }
    const bool ok = tc::mbar_wait(&mbar, 0);
    tc::tc_fence_after();
    const int row = warp * 32 + lane;
    for (int c0 = 0; c0 < p.N; c0 += 32) {
        float v[32];
        tc::tmem_ld32(tmem + ((uint32_t)(warp * 32) << 16) + c0, v);
        for (int i = 0; i < 32; ++i) p.D[row * p.N + c0 + i] = ok ? v[i] : NAN;
    }
    tc::tc_fence_before();
    __syncthreads();
    if (warp == 0) tc::tmem_dealloc<512>(tmem);
}


// ---- T7: MN-major tf32 operands in the 128B-SWIZZLED layout (rows = frames, 32 channels = 128 B per row) ----------
//   elem(frame f, channel c) at panel(c/32) + (f/8)*1024 + (f%8)*128 + (((c%32)/4) ^ (f%8))*16 + (c%4)*4
//   descriptor: layout SWIZZLE_128B (2), LBO = panel stride (next 32 channels), SBO = 1024 (next 8 frames)
struct PS { const float* A; const float* B; float* D; int K, N, rows_alloc, shiftA, shiftB, boff_mode; };
__device__ void stage_sw(float* dst, const float* src, int rows, int cols, int rows_alloc) {
    for (int i = threadIdx.x; i < rows * cols; i += blockDim.x) {
        const int f = i / cols, c = i % cols;
        const int off = (c / 32) * rows_alloc * 32 + (f / 8) * 256 + (f % 8) * 32 + ((((c % 32) / 4) ^ (f % 8)) * 4) + (c % 4);
        dst[off] = src[i];
    }
}
__global__ void __launch_bounds__(128) k_probe_sw(PS p) {
    extern __shared__ __align__(1024) float4 sm4[];
    float* smem = reinterpret_cast<float*>((reinterpret_cast<uintptr_t>(sm4) + 1023) & ~(uintptr_t)1023);
    __shared__ uint64_t mbar;
    __shared__ uint32_t tmem_base;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int rA = p.K + p.shiftA, rB = p.K + p.shiftB;
    float* a = smem;
    float* b = a + 4 * p.rows_alloc * 32;               // 4 panels of A (128 channels)
    stage_sw(a, p.A, rA, 128, p.rows_alloc);
    stage_sw(b, p.B, rB, p.N, p.rows_alloc);
    if (threadIdx.x == 0) { tc::mbar_init(&mbar, 1); tc::fence_mbar_init(); }
    if (warp == 0) tc::tmem_alloc<128>(&tmem_base);
    tc::fence_proxy_async_smem();
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    const uint32_t tmem = tmem_base;
    if (threadIdx.x == 0) {
        const uint32_t idesc = tc::make_idesc_tf32(128, p.N, 1, 1);
        const uint32_t pstride = p.rows_alloc * 128;
        uint32_t acc = 0;
        for (int k0 = 0; k0 < p.K; k0 += 8) {
            const uint32_t sa = tc::smem_u32(a) + (k0 + p.shiftA) * 128, sb = tc::smem_u32(b) + (k0 + p.shiftB) * 128;
            uint64_t da = tc::make_smem_desc(sa, pstride, 1024) | (2ull << 61);
            uint64_t db = tc::make_smem_desc(sb, pstride, 1024) | (2ull << 61);
            if (p.boff_mode == 1) { da |= (uint64_t)((sa >> 7) & 7) << 49; db |= (uint64_t)((sb >> 7) & 7) << 49; }
            tc::umma_tf32(tmem, da, db, idesc, acc);
            acc = 1;
        }
        tc::umma_commit(&mbar);
    }
    const bool ok = tc::mbar_wait(&mbar, 0);
    tc::tc_fence_after();
    const int row = warp * 32 + lane;
    for (int c0 = 0; c0 < p.N; c0 += 32) {
        float v[32];
        tc::tmem_ld32(tmem + ((uint32_t)(warp * 32) << 16) + c0, v);
        for (int i = 0; i < 32; ++i) p.D[row * p.N + c0 + i] = ok ? v[i] : NAN;
    }
    tc::tc_fence_before();
    __syncthreads();
    if (warp == 0) tc::tmem_dealloc<128>(tmem);
}

// ---- T5: 16x256b fragment layout -----------------------------------------------------------------
__global__ void __launch_bounds__(128) k_frag(float* out /*[128 thr][16 regs][1]*/) {
    __shared__ uint32_t tmem_base;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp == 0) tc::tmem_alloc<64>(&tmem_base);
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    const uint32_t tmem = tmem_base;
    // fill: lane m, column c <- m*256 + c  (32x32b store: thread = lane)
    {
        float v[32];
        for (int c0 = 0; c0 < 64; c0 += 32) {
            for (int i = 0; i < 32; ++i) v[i] = (float)((warp * 32 + lane) * 256 + c0 + i);
            tc::tmem_st32(tmem + ((uint32_t)(warp * 32) << 16) + c0, v);
        }
        tc::tmem_st_wait();
    }
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    uint32_t r[16];
    // .16x256b.x4: 16 lanes x 32 columns -> 16 registers per thread; lanes base .. base+15
    asm volatile(
        "tcgen05.ld.sync.aligned.16x256b.x4.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(tmem + ((uint32_t)(warp * 32 + 16) << 16) + 8)      // second half of the warp's lanes, columns 8..39
        : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    for (int i = 0; i < 16; ++i) out[threadIdx.x * 16 + i] = __uint_as_float(r[i]);
    tc::tc_fence_before();
    __syncthreads();
    if (warp == 0) tc::tmem_dealloc<64>(tmem);
}

// ---- T6: MMA rate --------------------------------------------------------------------------------
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// kind: 0 SS tf32, 1 TS tf32, 2 SS bf16 (K = 16 per MMA; tiles of 8-element 16 B chunks)
__global__ void __launch_bounds__(128) k_rate(int kind, int N, int nmma, long long* cycles) {
    extern __shared__ float4 sm4[];
    float* smem = reinterpret_cast<float*>(sm4);
    __shared__ uint64_t mbar;
    __shared__ uint32_t tmem_base;
    const int warp = threadIdx.x >> 5;
    const int rowsA = 137, rowsB = tc::chunk_rows(N);
    // K = 64 fp32 channels = 16 chunks per operand (bf16: the same bytes hold K = 128)
    const int szA = 16 * rowsA * 4, szB = 16 * rowsB * 4;
    for (int i = threadIdx.x; i < szA + szB; i += blockDim.x) smem[i] = 0.f;
    if (threadIdx.x == 0) { tc::mbar_init(&mbar, 1); tc::fence_mbar_init(); }
    if (warp == 0) tc::tmem_alloc<512>(&tmem_base);
    tc::fence_proxy_async_smem();
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    const uint32_t tmem = tmem_base;
    if (kind == 1) {
        float z[32];
        for (int i = 0; i < 32; ++i) z[i] = 0.f;
        for (int c = 0; c < 64; c += 32) tc::tmem_st32(tmem + ((uint32_t)(warp * 32) << 16) + 256 + c, z);
        tc::tmem_st_wait();
        tc::tc_fence_before();
        __syncthreads();
        tc::tc_fence_after();
    }
    long long c0 = 0;
    if (warp == 0) {
        const uint32_t idesc = kind == 2 ? ((1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | (8u << 24))
                                         : tc::make_idesc_tf32(128, N, 0, 0);
        const uint32_t a_s = tc::smem_u32(smem), b_s = tc::smem_u32(smem + szA);
        const uint32_t csA = rowsA * 16, csB = rowsB * 16;
        const bool leader = tc::elect_one();
        uint32_t acc = 0;
        c0 = clock64();
        for (int i = 0; i < nmma; ++i) {
            const int ks = i & 7;                     // 8 K-steps over the tile, cyclic
            const uint64_t da = tc::make_smem_desc(a_s + ks * 2 * csA, csA, 128);
            const uint64_t db = tc::make_smem_desc(b_s + ks * 2 * csB, csB, 128);
            if (leader) {
                if (kind == 0) tc::umma_tf32(tmem, da, db, idesc, acc);
                else if (kind == 1) tc::umma_tf32_ts(tmem, tmem + 256 + ks * 8, db, idesc, acc);
                else umma_f16(tmem, da, db, idesc, acc);
            }
            acc = 1;
        }
        if (leader) tc::umma_commit(&mbar);
        __syncwarp();
    }
    const bool ok = tc::mbar_wait(&mbar, 0);
    tc::tc_fence_after();
    if (threadIdx.x == 0) cycles[blockIdx.x] = ok ? clock64() - c0 : -1;
    tc::tc_fence_before();
    __syncthreads();
    if (warp == 0) tc::tmem_dealloc<512>(tmem);
}

// ---------------------------------------------------------------------------------------------------
static float frand() { return (float)rand() / RAND_MAX * 2.f - 1.f; }

int main(int argc, char** argv) {
    auto want = [&](const char* g) { return argc < 2 || strcmp(argv[1], g) == 0; };
    CK(cudaFuncSetAttribute(k_probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
    CK(cudaFuncSetAttribute(k_rate, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
    srand(7);
    const int K = 64, M = 128, KS = K + 8;
    // frame-major sources (rows = frames, with 8 spare rows for the shift tests) and their transposes
    std::vector<float> Afm(KS * M), Bfn(KS * 128), Amk(M * K), Bnk(128 * K);
    for (auto& x : Afm) x = frand();
    for (auto& x : Bfn) x = frand();
    float *dA, *dB, *dD;
    CK(cudaMalloc(&dA, 1 << 20)); CK(cudaMalloc(&dB, 1 << 20)); CK(cudaMalloc(&dD, 1 << 20));
    std::vector<float> D(M * 256);

    auto run = [&](P p, const std::vector<float>& A, const std::vector<float>& B, int rA, int cA, int rB, int cB) -> bool {
        CK(cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice));
        CK(cudaMemcpy(dB, B.data(), B.size() * 4, cudaMemcpyHostToDevice));
        CK(cudaMemset(dD, 0xff, M * p.N * 4));
        p.A = dA; p.B = dB; p.D = dD;
        const size_t smem = 2 * ((size_t)(cA / 4) * p.rowsA_alloc * 16 + (size_t)(cB / 4) * p.rowsB_alloc * 16);
        if (smem > 220 * 1024) { printf("smem too large\n"); return false; }
        k_probe<<<1, 128, smem>>>(p);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("  kernel error: %s\n", cudaGetErrorString(e)); exit(2); }
        CK(cudaMemcpy(D.data(), dD, M * p.N * 4, cudaMemcpyDeviceToHost));
        return true;
    };
    // reference for frame-major inputs with shifts: D[m][n] = sum_f A[f+sa][m] * B[f+sb][n]
    auto ref_fm = [&](int N, int sa, int sb, int ldb, std::vector<double>& R) {
        R.assign(M * N, 0.0);
        for (int m = 0; m < M; ++m) for (int n = 0; n < N; ++n) {
            double s = 0; for (int f = 0; f < K; ++f) s += (double)Afm[(f + sa) * M + m] * Bfn[(f + sb) * ldb + n];
            R[m * N + n] = s;
        }
    };
    auto relerr = [&](const std::vector<double>& R, int N) {
        double mx = 0, mr = 0; bool nan = false;
        for (int i = 0; i < M * N; ++i) { if (D[i] != D[i]) nan = true; mx = fmax(mx, fabs(D[i] - R[i])); mr = fmax(mr, fabs(R[i])); }
        return nan ? -1.0 : mx / mr;
    };

    printf("== T1/T2: MN-major tf32, no-swizzle chunk-major (rows = frames).  err<3e-3 (tf32) means the layout works\n");
    for (int N : {64, 128}) {
        // B source with N columns
        std::vector<float> Bsrc(KS * N);
        for (int f = 0; f < KS; ++f) for (int n = 0; n < N; ++n) Bsrc[f * N + n] = Bfn[f * 128 + n];
        for (int f = 0; f < K; ++f) for (int m = 0; m < M; ++m) Amk[m * K + f] = Afm[f * M + m];
        std::vector<float> Bnk2(N * K);
        for (int f = 0; f < K; ++f) for (int n = 0; n < N; ++n) Bnk2[n * K + f] = Bfn[f * 128 + n];
        for (int ralloc : {73, 72, 80}) for (int variant = 0; variant < 3; ++variant) for (int maj = 0; maj < 3; ++maj) for (int shift : {0, 3}) {
            char gname[16]; snprintf(gname, sizeof gname, "T1v%d", variant);
            if (!want(gname)) continue;
            P p = {}; p.N = N; p.K = K; p.variant = variant; p.split = 0;
            p.a_mn = (maj == 0 || maj == 1); p.b_mn = (maj == 0 || maj == 2);
            p.shiftA = 0; p.shiftB = p.b_mn ? shift : 0;
            if (!p.b_mn && shift) continue;
            p.rowsA_alloc = p.a_mn ? ralloc : 129; p.rowsB_alloc = p.b_mn ? ralloc : tc::chunk_rows(N);
            const std::vector<float>& A = p.a_mn ? Afm : Amk;
            const std::vector<float>& B = p.b_mn ? Bsrc : Bnk2;
            std::vector<float> Acut(A.begin(), A.begin() + (p.a_mn ? (K + p.shiftA) * M : M * K));
            std::vector<float> Bcut(B.begin(), B.begin() + (p.b_mn ? (K + p.shiftB) * N : N * K));
            run(p, Acut, Bcut, 0, p.a_mn ? 128 : K, 0, p.b_mn ? N : K);
            std::vector<double> R; ref_fm(N, 0, p.shiftB, 128, R);
            printf("T1 N=%3d rows_alloc=%2d variant=%d a_mn=%d b_mn=%d shiftB=%d  relerr=%.3e\n", N, ralloc, variant, p.a_mn, p.b_mn, p.shiftB, relerr(R, N));
        }
    }


    if (want("T7")) {
        printf("== T7: MN-major tf32, SWIZZLE_128B (rows = frames).  err<3e-3 means it works\n");
        CK(cudaFuncSetAttribute(k_probe_sw, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        for (int N : {64, 128}) for (int boff = 0; boff < 2; ++boff) for (int sa : {0, 8}) for (int sb : {0, 1, 3, 8}) {
            std::vector<float> Bsrc((K + 8) * N);
            for (int f = 0; f < K + 8; ++f) for (int n = 0; n < N; ++n) Bsrc[f * N + n] = Bfn[f * 128 + n];
            CK(cudaMemcpy(dA, Afm.data(), Afm.size() * 4, cudaMemcpyHostToDevice));
            CK(cudaMemcpy(dB, Bsrc.data(), Bsrc.size() * 4, cudaMemcpyHostToDevice));
            CK(cudaMemset(dD, 0xff, M * N * 4));
            PS ps = {dA, dB, dD, K, N, 72, sa, sb, boff};
            k_probe_sw<<<1, 128, (4 + N / 32) * 72 * 128 + 2048>>>(ps);
            cudaError_t e = cudaDeviceSynchronize();
            if (e != cudaSuccess) { printf("  kernel error: %s\n", cudaGetErrorString(e)); exit(2); }
            CK(cudaMemcpy(D.data(), dD, M * N * 4, cudaMemcpyDeviceToHost));
            std::vector<double> R(M * N, 0.0);
            for (int m = 0; m < M; ++m) for (int n = 0; n < N; ++n) {
                double sum = 0; for (int f = 0; f < K; ++f) sum += (double)Afm[(f + sa) * M + m] * Bfn[(f + sb) * 128 + n];
                R[m * N + n] = sum;
            }
            printf("T7 N=%3d base_offset_mode=%d shiftA=%d shiftB=%d relerr=%.3e\n", N, boff, sa, sb, relerr(R, N));
        }
    }

    printf("== T3: TS form (A from TMEM), K-major B.  split=1 expects ~1e-6, split=0 ~1e-3\n");
    for (int N : {64, 128}) for (int split : {0, 1}) for (int coff : {0, 1, 2, 4, 8}) {
        if (!want((coff == 1 || coff == 2) ? "T3u" : "T3a")) continue;
        for (int f = 0; f < K; ++f) for (int m = 0; m < M; ++m) Amk[m * K + f] = Afm[f * M + m];
        std::vector<float> Bnk2(N * K);
        for (int f = 0; f < K; ++f) for (int n = 0; n < N; ++n) Bnk2[n * K + f] = Bfn[f * 128 + n];
        P p = {}; p.N = N; p.K = K; p.split = split; p.ts = 1; p.coff = coff;
        p.rowsA_alloc = 129; p.rowsB_alloc = tc::chunk_rows(N);
        run(p, Amk, Bnk2, 0, K, 0, K);
        std::vector<double> R; ref_fm(N, 0, 0, 128, R);
        printf("T3 N=%3d split=%d coff=%d relerr=%.3e\n", N, split, coff, relerr(R, N));
    }

    printf("== T4: raw fp32 operands vs masked-hi operands (bitwise equal => the tensor core truncates)\n");
    if (want("T4")) {
        const int N = 128;
        for (int f = 0; f < K; ++f) for (int m = 0; m < M; ++m) Amk[m * K + f] = Afm[f * M + m];
        std::vector<float> Bnk2(N * K);
        for (int f = 0; f < K; ++f) for (int n = 0; n < N; ++n) Bnk2[n * K + f] = Bfn[f * 128 + n];
        P p = {}; p.N = N; p.K = K; p.rowsA_alloc = 129; p.rowsB_alloc = 129;
        run(p, Amk, Bnk2, 0, K, 0, K);
        std::vector<float> D0(D.begin(), D.begin() + M * N);
        p.raw = 1;
        run(p, Amk, Bnk2, 0, K, 0, K);
        int diff = 0; double mx = 0;
        for (int i = 0; i < M * N; ++i) { if (memcmp(&D0[i], &D[i], 4)) ++diff; mx = fmax(mx, fabs(D0[i] - D[i])); }
        printf("T4 SS: %d of %d outputs differ bitwise, max abs diff %.3e\n", diff, M * N, mx);
        p.raw = 0; p.ts = 1;
        run(p, Amk, Bnk2, 0, K, 0, K);
        D0.assign(D.begin(), D.begin() + M * N);
        p.raw = 1;
        run(p, Amk, Bnk2, 0, K, 0, K);
        diff = 0; mx = 0;
        for (int i = 0; i < M * N; ++i) { if (memcmp(&D0[i], &D[i], 4)) ++diff; mx = fmax(mx, fabs(D0[i] - D[i])); }
        printf("T4 TS: %d of %d outputs differ bitwise, max abs diff %.3e\n", diff, M * N, mx);
    }

    printf("== T5: tcgen05.ld.16x256b.x4 at lane base+16, column 8: value = lane*256 + column\n");
    if (want("T5")) {
        float* dO; CK(cudaMalloc(&dO, 128 * 16 * 4));
        k_frag<<<1, 128>>>(dO);
        CK(cudaDeviceSynchronize());
        std::vector<float> O(128 * 16);
        CK(cudaMemcpy(O.data(), dO, O.size() * 4, cudaMemcpyDeviceToHost));
        for (int t : {0, 1, 2, 3, 4, 5, 31, 32, 33, 64, 127}) {
            printf("T5 thread %3d:", t);
            for (int i = 0; i < 16; ++i) { int v = (int)O[t * 16 + i]; printf(" (%d,%d)", v / 256, v % 256); }
            printf("\n");
        }
        // check the assumed mapping: reg[4*b + 2*h + e] = (lane 16 + 32*warp + t/4 + 8*h, col 8 + 8*b + 2*(t%4) + e)
        int bad = 0;
        for (int t = 0; t < 128; ++t) for (int b = 0; b < 4; ++b) for (int h = 0; h < 2; ++h) for (int e = 0; e < 2; ++e) {
            const int lane = (t / 32) * 32 + 16 + (t % 32) / 4 + 8 * h, col = 8 + 8 * b + 2 * (t % 4) + e;
            if ((int)O[t * 16 + 4 * b + 2 * h + e] != lane * 256 + col) ++bad;
        }
        printf("T5 assumed mapping reg[4b+2h+e] = (lane0 + t/4 + 8h, col0 + 8b + 2(t%%4) + e): %d mismatches\n", bad);
    }

    printf("== T6: cycles per MMA, warp-collective issue, 148 CTAs (M = 128)\n");
    if (want("T6")) {
        long long* dC; CK(cudaMalloc(&dC, 148 * 8));
        std::vector<long long> C(148);
        const int nmma = 960;
        struct Cfg { int kind, N; const char* name; } cfgs[] = {
            {0, 64, "SS tf32 N=64"}, {0, 128, "SS tf32 N=128"}, {0, 256, "SS tf32 N=256"},
            {1, 64, "TS tf32 N=64"}, {1, 128, "TS tf32 N=128"}, {1, 256, "TS tf32 N=256"},
            {2, 64, "SS bf16 N=64 (K=16)"}, {2, 128, "SS bf16 N=128 (K=16)"}, {2, 256, "SS bf16 N=256 (K=16)"}};
        for (auto& c : cfgs) {
            const size_t smem = (size_t)16 * (137 + tc::chunk_rows(c.N)) * 16;
            for (int rep = 0; rep < 2; ++rep) { k_rate<<<148, 128, smem>>>(c.kind, c.N, nmma, dC); CK(cudaDeviceSynchronize()); }
            CK(cudaMemcpy(C.data(), dC, 148 * 8, cudaMemcpyDeviceToHost));
            double s = 0; for (auto v : C) s += (double)v;
            printf("T6 %-24s %7.1f cycles/MMA\n", c.name, s / 148 / nmma);
        }
    }
    printf("done\n");
    return 0;
}
