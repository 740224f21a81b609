// crank-b200: shared device helpers for the fp32 (CUDA-core) kernel family.
//
// Data layout everywhere: channels-last.  An activation is a row-major (B*T, C) panel,
// "(batch*time) x channel", row stride `ld` floats (so sub-ranges of a concatenated buffer can be
// addressed without a copy).  A CTA owns a tile of TM=64 consecutive frames of ONE utterance,
// so "same"/causal zero padding never bleeds across utterances.
//
// Thread mapping of every tile GEMM (256 threads): ty = tid/32 owns 8 consecutive tile rows,
// tx = tid%32 owns CPT consecutive output columns (TN = 32*CPT <= 128).  All lanes of a warp
// share ty, so A-operand shared-memory reads are broadcasts; B-operand reads are CPT*4-byte
// vectors, conflict free.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#define CRK_THREADS 256
#define CRK_TM 64

#define CRK_ACT_NONE 0
#define CRK_ACT_RELU 1
#define CRK_ACT_LRELU 2

#define CRK_SQRT_HALF 0.70710678118654752440f

// error codes of the C ABI (see include/crank_b200.h)
#define CRK_OK 0
#define CRK_ERR_ARG (-1)
#define CRK_ERR_CUDA (-2)
#define CRK_ERR_UNSUPPORTED (-3)

namespace crk {

// ---- instrumentation: launch counter + optional per-kernel CUDA-event timing -------------------
// Every kernel launch of the library bumps the counter (bench.py reports it as gpu_launches).
// crk_timing_enable(id) makes the launch helpers of kernel family `id` record a cudaEvent pair
// on the launch stream around each launch; crk_timing_read() synchronises and sums them.
enum { CRK_K_RESBLOCK_FWD = 1, CRK_K_WGRAD = 2, CRK_K_CONV = 3, CRK_K_BWD_GATE = 4, CRK_K_VQ_ARGMIN = 5, CRK_K_LOGMEL = 6, CRK_K_MAX = 8 };
struct Instr {
    unsigned long long launches = 0;
    int enabled_id = 0;
    static const int kMaxPairs = 8192;
    cudaEvent_t* ev = nullptr;   // 2*kMaxPairs
    int npairs = 0;
    double flops = 0.0;          // algorithmic FLOPs of the timed launches
};
inline Instr& instr() { static Instr i; return i; }
inline cudaError_t launch_check() { ++instr().launches; return cudaGetLastError(); }
struct TimedLaunch {   // RAII: records events around a launch when timing of `id` is on
    bool on; cudaStream_t s; int slot;
    TimedLaunch(int id, cudaStream_t st, double flops = 0.0) : on(false), s(st), slot(0) {
        Instr& I = instr();
        if (I.enabled_id == id && I.ev && I.npairs < Instr::kMaxPairs) {
            on = true; slot = I.npairs++;
            I.flops += flops;
            cudaEventRecord(I.ev[2 * slot], s);
        }
    }
    ~TimedLaunch() { if (on) cudaEventRecord(instr().ev[2 * slot + 1], s); }
};

// arithmetic of the dense conv contractions: fp32 CUDA cores, 3xTF32 (error-compensated, ~fp32
// accuracy) or plain TF32 on the tcgen05 tensor cores.  Process-wide; set by crk_set_precision().
enum { CRK_PREC_FP32 = 0, CRK_PREC_TF32X3 = 1, CRK_PREC_TF32 = 2 };
inline int& precision_mode() { static int m = CRK_PREC_FP32; return m; }
// debugging: bit mask of tensor-core kernel families forced back to the fp32 kernels
// (1 fused forward, 2 conv/dgrad, 4 wgrad, 8 gate backward)
inline int& tc_disable_mask() { static int m = 0; return m; }

// debugging / A-B switches for optional optimisations (crk_debug_opt_mask):
//   1 programmatic dependent launch off, 2 bias column sums fused into k_wgrad_tc off,
//   4 k_conv_tc FAST instance (128-bit staging + epilogue) off -> GENERIC instance, 8 raw-tile k_wgrad_tc_raw off,
//   16 persistent pipelined fused forward (k_resblock_fwd_pt) off -> k_resblock_fwd_tc,
//   32 persistent pipelined conv / dgrad (k_conv_pt) off -> k_conv_tc
//   64 two-CTA/SM fused forward (k_resblock_fwd_tc2) off -> round 1's k_resblock_fwd_tc
//   128 two-CTA/SM K-phased k_conv_tc off -> whole-K variant (one CTA per SM in the 3xTF32 mode)
//   1024 16 worker warps in k_wgrad_tc_raw off -> 8
//   512 weight gradients of small launches on a side stream (crk_stacks.cuh wavenet_bwd) off -> everything on the caller's stream
inline int& opt_disable_mask() { static int m = 0; return m; }
// opt-IN switches (crk_debug_opt_enable / CRANK_B200_OPT_ENABLE): experimental paths that are correct (parity-tested)
// but not yet faster than the default ones:
//   1 persistent pipelined fused forward k_resblock_fwd_pt, 2 persistent pipelined conv / dgrad k_conv_pt,
//   8 transposed wgrad partial blocks (16 B stores in the epilogue, k_reduce_t): measured no gain
//   4 wgrad partial block staged in shared memory + one TMA bulk store per CTA (measured: wgrad family 8.0 vs 7.75 ms per step
//     with direct stores -- the serialised stage -> fence -> bulk store costs more than the 128 B-per-instruction stores it replaces)
inline int& opt_enable_mask() { static int m = 0; return m; }

// ---- programmatic dependent launch (PDL) -----------------------------------------------------
// The steps of a stack are hundreds of short dependent kernels on one stream.  A kernel launched with
// the programmatic-stream-serialisation attribute may become resident while its predecessor's last
// wave is still running (every predecessor CTA has executed pdl_trigger() or exited): block
// scheduling, barrier init and the TMEM allocation then overlap the predecessor's tail instead of
// following its drain.  Contract kept by every kernel launched through launch_pdl(): NO global memory
// access before pdl_wait() (which returns once the predecessor grid has completed and its writes are
// visible), so the result is identical to plain stream order.  Both instructions are no-ops in a
// kernel launched without the attribute.  Kernels that allocate tensor memory trigger AFTER the
// allocation: a dependent CTA that became co-resident on the same SM (possible in the plain-TF32 variants,
// two CTAs per SM) and grabbed the TMEM columns first would then wait for a predecessor that can never
// get its own -- a deadlock by construction, however unlikely the interleaving.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t s,
                              Args... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = s;
    cudaLaunchAttribute at[1];
    if (!(opt_disable_mask() & 1)) {
        at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        at[0].val.programmaticStreamSerializationAllowed = 1;
        cfg.attrs = at; cfg.numAttrs = 1;
    }
    return cudaLaunchKernelEx(&cfg, kern, KArgs(args)...);
}

__host__ __device__ inline int round_up(int a, int b) { return (a + b - 1) / b * b; }
__host__ __device__ inline int cdiv(int a, int b) { return (a + b - 1) / b; }
__host__ __device__ inline long long cdivl(long long a, long long b) { return (a + b - 1) / b; }

// columns-per-thread needed for n output columns (TN = 32*cpt)
__host__ __device__ inline int cpt_for(int n) { return n <= 32 ? 1 : n <= 64 ? 2 : n <= 96 ? 3 : 4; }
// leading dimension of a packed matrix with n columns: 32*cpt up to 128 columns; wider (the first conv of a stack fed by
// three concatenated 64-channel code streams, n_vq_stacks = 3: 192 input channels) rounded up to 32 -- such a matrix is
// consumed in 128-column slices (ConvParams::ldw)
__host__ __device__ inline int wide_ld(int n) { return n <= 128 ? 32 * cpt_for(n) : ((n + 31) / 32) * 32; }
#define CRK_MAX_IN_CH 256

__device__ __forceinline__ float apply_act(float v, int act, float slope) {
    if (act == CRK_ACT_RELU) return v > 0.f ? v : 0.f;
    if (act == CRK_ACT_LRELU) return v > 0.f ? v : v * slope;
    return v;
}
// derivative of the activation, evaluated from the activation's OUTPUT (sign-preserving acts)
__device__ __forceinline__ float act_grad(float out, int act, float slope) {
    if (act == CRK_ACT_RELU) return out > 0.f ? 1.f : 0.f;
    if (act == CRK_ACT_LRELU) return out > 0.f ? 1.f : slope;
    return 1.f;
}

template <int CPT>
__device__ __forceinline__ void load_b(float (&b)[CPT], const float* p) {
    if constexpr (CPT == 4) {
        float4 v = *reinterpret_cast<const float4*>(p);
        b[0] = v.x; b[1] = v.y; b[2] = v.z; b[3] = v.w;
    } else if constexpr (CPT == 2) {
        float2 v = *reinterpret_cast<const float2*>(p);
        b[0] = v.x; b[1] = v.y;
    } else {
#pragma unroll
        for (int c = 0; c < CPT; ++c) b[c] = p[c];
    }
}

// acc[i][c] += sum_kk A[i][kk] * B[kk][c];  A row-major in smem (row stride lda, lda%4==0,
// 16B-aligned), K%4==0.  `A` already points at this thread's first row, `Bm` at its first column.
template <int CPT>
__device__ __forceinline__ void tile_mac_rowA(float (&acc)[8][CPT], const float* __restrict__ A,
                                              int lda, const float* __restrict__ Bm, int ldb, int K) {
    for (int kk = 0; kk < K; kk += 4) {
        float4 a[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) a[i] = *reinterpret_cast<const float4*>(A + i * lda + kk);
        float b[CPT];
        load_b<CPT>(b, Bm + (kk + 0) * ldb);
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
            for (int c = 0; c < CPT; ++c) acc[i][c] = fmaf(a[i].x, b[c], acc[i][c]);
        load_b<CPT>(b, Bm + (kk + 1) * ldb);
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
            for (int c = 0; c < CPT; ++c) acc[i][c] = fmaf(a[i].y, b[c], acc[i][c]);
        load_b<CPT>(b, Bm + (kk + 2) * ldb);
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
            for (int c = 0; c < CPT; ++c) acc[i][c] = fmaf(a[i].z, b[c], acc[i][c]);
        load_b<CPT>(b, Bm + (kk + 3) * ldb);
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
            for (int c = 0; c < CPT; ++c) acc[i][c] = fmaf(a[i].w, b[c], acc[i][c]);
    }
}

// acc[i][c] += sum_kk A[kk][i] * B[kk][c];  A "column-major": smem [K][lda], the thread's 8 rows
// are 8 consecutive floats (16B aligned).  Used by wgrad (reduction index = frame).
template <int CPT>
__device__ __forceinline__ void tile_mac_colA(float (&acc)[8][CPT], const float* __restrict__ A,
                                              int lda, const float* __restrict__ Bm, int ldb, int K) {
#pragma unroll 4
    for (int kk = 0; kk < K; ++kk) {
        const float4 a0 = *reinterpret_cast<const float4*>(A + kk * lda);
        const float4 a1 = *reinterpret_cast<const float4*>(A + kk * lda + 4);
        float b[CPT];
        load_b<CPT>(b, Bm + kk * ldb);
        const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
            for (int c = 0; c < CPT; ++c) acc[i][c] = fmaf(a[i], b[c], acc[i][c]);
    }
}

// cooperative copy of n floats (n%4==0, both 16B aligned) global->shared
__device__ __forceinline__ void copy_to_smem(float* dst, const float* __restrict__ src, int n) {
    const float4* s4 = reinterpret_cast<const float4*>(src);
    float4* d4 = reinterpret_cast<float4*>(dst);
    for (int i = threadIdx.x; i < n / 4; i += CRK_THREADS) d4[i] = __ldg(s4 + i);
}

}  // namespace crk
