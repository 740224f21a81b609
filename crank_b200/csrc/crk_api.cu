// crank-b200: the C ABI (include/crank_b200.h) over the fp32 kernel family.  Single translation
// unit: all kernels are header-defined and instantiated here.
#include <cufft.h>
#include <stdio.h>

#include <map>
#include <utility>

#include "../../include/crank_b200.h"
#include "crk_common.cuh"
#include "crk_conv.cuh"
#include "crk_loss.cuh"
#include "crk_logmel.cuh"
#include "crk_resblock.cuh"
#include "crk_stacks.cuh"
#include "crk_tc_probe.cuh"
#include "crk_vq.cuh"
#include "crk_vq_tc.cuh"
#include "crk_vq_fast.cuh"

namespace crk {
static thread_local char g_cuda_err[256] = "";
int set_cuda_error(cudaError_t e) {
    snprintf(g_cuda_err, sizeof(g_cuda_err), "%s: %s", cudaGetErrorName(e), cudaGetErrorString(e));
    return CRK_ERR_CUDA;
}
}  // namespace crk

using namespace crk;

#define API_TRY(expr)                                     \
    do {                                                  \
        cudaError_t _e = (expr);                          \
        if (_e != cudaSuccess) return set_cuda_error(_e); \
    } while (0)

extern "C" {

const char* crk_strerror(int code) {
    switch (code) {
        case CRK_OK: return "ok";
        case CRK_ERR_ARG: return "invalid argument";
        case CRK_ERR_CUDA: return "CUDA error (see crk_last_cuda_error)";
        case CRK_ERR_UNSUPPORTED: return "unsupported configuration";
        default: return "unknown error";
    }
}
const char* crk_last_cuda_error(void) { return g_cuda_err; }
int crk_version(void) { return 100; }

int crk_set_precision(int mode) {
    if (mode < CRK_PREC_FP32 || mode > CRK_PREC_TF32) return CRK_ERR_ARG;
    precision_mode() = mode;
    return CRK_OK;
}
int crk_get_precision(void) { return precision_mode(); }
int crk_debug_tc_disable(int mask) { tc_disable_mask() = mask; return CRK_OK; }
int crk_debug_opt_disable(int mask) { opt_disable_mask() = mask; return CRK_OK; }
int crk_debug_opt_enable(int mask) { opt_enable_mask() = mask; return CRK_OK; }
int crk_debug_timestamps(long long* device_buffer, int kernel_id, int launch_index) {
    API_TRY(cudaMemcpyToSymbol(g_crk_dbg, &device_buffer, sizeof(device_buffer)));
    DbgSel& d = dbg_sel();
    d.kind = device_buffer ? kernel_id : 0; d.target = launch_index; d.count = 0;
    return CRK_OK;
}
unsigned long long crk_launch_count(void) { return instr().launches; }
int crk_timing_enable(int kernel_id) {
    Instr& I = instr();
    if (kernel_id < 0 || kernel_id >= CRK_K_MAX) return CRK_ERR_ARG;
    if (kernel_id > 0 && !I.ev) {
        I.ev = new cudaEvent_t[2 * Instr::kMaxPairs];
        for (int i = 0; i < 2 * Instr::kMaxPairs; ++i) API_TRY(cudaEventCreate(&I.ev[i]));
    }
    I.enabled_id = kernel_id;
    I.npairs = 0;
    I.flops = 0.0;
    return CRK_OK;
}
double crk_timing_flops(void) { return instr().flops; }
int crk_timing_read(int* count, float* total_ms) {
    Instr& I = instr();
    if (!count || !total_ms) return CRK_ERR_ARG;
    API_TRY(cudaDeviceSynchronize());
    float tot = 0.f;
    for (int i = 0; i < I.npairs; ++i) {
        float ms = 0.f;
        API_TRY(cudaEventElapsedTime(&ms, I.ev[2 * i], I.ev[2 * i + 1]));
        tot += ms;
    }
    *count = I.npairs;
    *total_ms = tot;
    I.npairs = 0;
    return CRK_OK;
}

// ---- tcgen05 probe ---------------------------------------------------------------------------
int crk_tc_probe(const float* A, int lda, int rowsA, const float* B, int ldb, int rowsB, float* D, int N,
                 int K, int row_shift, int mode, int split, void* stream) {
    if (!A || !B || !D || (N != 64 && N != 128) || K < 8 || (K % 8) != 0 || row_shift < 0) return CRK_ERR_ARG;
    if ((mode & 1) == 0 && (rowsA < 128 + row_shift || rowsB < N)) return CRK_ERR_ARG;
    if ((mode & 1) == 1 && (rowsA < K || rowsB < K)) return CRK_ERR_ARG;
    TcProbeParams p;
    p.A = A; p.lda = lda; p.rowsA = rowsA; p.B = B; p.ldb = ldb; p.rowsB = rowsB; p.D = D;
    p.N = N; p.K = K; p.row_shift = row_shift; p.mode = mode & 1; p.split = split; p.variant = mode >> 1;
    const int colsA = p.mode == 0 ? K : 128, colsB = p.mode == 0 ? K : N;
    const size_t smem = (size_t)(split ? 2 : 1) * ((colsA / 4) * tc::chunk_stride_bytes(rowsA) + (colsB / 4) * tc::chunk_stride_bytes(rowsB));
    if (smem > 200 * 1024) return CRK_ERR_UNSUPPORTED;
    API_TRY(cudaFuncSetAttribute(k_tc_probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    k_tc_probe<<<1, 128, smem, (cudaStream_t)stream>>>(p);
    API_TRY(launch_check());
    return CRK_OK;
}

int crk_tc_mma_rate(int N, int K, int reps, int split, int grid, long long* cycles, void* stream) {
    if (!cycles || N < 16 || N > 256 || (N % 16) != 0 || K < 8 || K > 128 || (K % 8) != 0 || reps == 0 || grid < 1) return CRK_ERR_ARG;
    const size_t smem = (size_t)2 * (K / 4) * (137 * 4 + tc::chunk_rows(N) * 4) * sizeof(float);
    if (smem > 220 * 1024) return CRK_ERR_UNSUPPORTED;
    API_TRY(cudaFuncSetAttribute(k_tc_mma_rate, cudaFuncAttributeMaxDynamicSharedMemorySize, 225 * 1024));
    k_tc_mma_rate<<<grid, 128, smem, (cudaStream_t)stream>>>(N, K, reps, split, cycles);
    API_TRY(launch_check());
    return CRK_OK;
}

// ---- WaveNet stack ---------------------------------------------------------------------------
int crk_wavenet_describe(const crk_wavenet_cfg* cfg, crk_conv_desc* descs, int* n_convs,
                         long long* theta_floats, long long* weff_floats) {
    WavenetLayout L;
    int rc = wavenet_layout(cfg, &L);
    if (rc) return rc;
    if (n_convs) *n_convs = L.tab.n;
    if (theta_floats) *theta_floats = L.theta;
    if (weff_floats) *weff_floats = L.weff;
    if (descs) memcpy(descs, L.tab.d, sizeof(crk_conv_desc) * L.tab.n);
    return CRK_OK;
}
long long crk_wavenet_act_floats(const crk_wavenet_cfg* cfg, int B, int T) {
    WavenetLayout L;
    if (wavenet_layout(cfg, &L) || B < 1 || T < 1) return -1;
    return wavenet_act(cfg, (long long)B * T).total;
}
long long crk_wavenet_ws_floats(const crk_wavenet_cfg* cfg, int B, int T) {
    WavenetLayout L;
    if (wavenet_layout(cfg, &L) || B < 1 || T < 1) return -1;
    return wavenet_ws(cfg, L, B, T).total;
}
int crk_wavenet_weights(const crk_wavenet_cfg* cfg, const float* theta, float* weff, void* stream) {
    WavenetLayout L;
    int rc = wavenet_layout(cfg, &L);
    if (rc) return rc;
    if (!theta || !weff) return CRK_ERR_ARG;
    API_TRY(launch_weightnorm(L.tab, theta, weff, (cudaStream_t)stream));
    return CRK_OK;
}
int crk_wavenet_fwd(const crk_wavenet_cfg* cfg, const float* weff, const float* x, int ldx,
                    const float* c, int ldc, const float* dropmul, float* y, int ldy, float* act,
                    int B, int T, void* stream) {
    return wavenet_fwd(cfg, weff, x, ldx, c, ldc, dropmul, y, ldy, act, B, T, (cudaStream_t)stream);
}
int crk_wavenet_infer(const crk_wavenet_cfg* cfg, const float* weff, const float* x, int ldx,
                      const float* c, int ldc, const float* dropmul, float* y, int ldy, float* act,
                      int B, int T, void* stream) {
    return wavenet_fwd(cfg, weff, x, ldx, c, ldc, dropmul, y, ldy, act, B, T, (cudaStream_t)stream, false);
}
int crk_wavenet_bwd(const crk_wavenet_cfg* cfg, const float* theta, const float* weff,
                    const float* x, int ldx, const float* c, int ldc, const float* dropmul,
                    const float* act, const float* dy, int lddy, float* dx, int lddx, float* dc,
                    int lddc, float* gtheta, float* ws, int B, int T, void* stream) {
    return wavenet_bwd(cfg, theta, weff, x, ldx, c, ldc, dropmul, act, dy, lddy, dx, lddx, dc, lddc,
                       gtheta, ws, B, T, (cudaStream_t)stream);
}

// ---- plain conv stack ------------------------------------------------------------------------
int crk_convstack_describe(const crk_convstack_cfg* cfg, crk_conv_desc* descs, int* n_convs,
                           long long* theta_floats, long long* weff_floats) {
    ConvstackLayout L;
    int rc = convstack_layout(cfg, &L);
    if (rc) return rc;
    if (n_convs) *n_convs = L.tab.n;
    if (theta_floats) *theta_floats = L.theta;
    if (weff_floats) *weff_floats = L.weff;
    if (descs) memcpy(descs, L.tab.d, sizeof(crk_conv_desc) * L.tab.n);
    return CRK_OK;
}
long long crk_convstack_act_floats(const crk_convstack_cfg* cfg, int B, int T) {
    if (B < 1 || T < 1) return -1;
    return convstack_act_floats(cfg, (long long)B * T);
}
long long crk_convstack_ws_floats(const crk_convstack_cfg* cfg, int B, int T) {
    ConvstackLayout L;
    if (convstack_layout(cfg, &L) || B < 1 || T < 1) return -1;
    return convstack_ws(cfg, L, B, T).total;
}
int crk_convstack_weights(const crk_convstack_cfg* cfg, const float* theta, float* weff, void* stream) {
    ConvstackLayout L;
    int rc = convstack_layout(cfg, &L);
    if (rc) return rc;
    if (!theta || !weff) return CRK_ERR_ARG;
    API_TRY(launch_weightnorm(L.tab, theta, weff, (cudaStream_t)stream));
    return CRK_OK;
}
int crk_convstack_fwd(const crk_convstack_cfg* cfg, const float* weff, const float* x, int ldx,
                      float* y, int ldy, float* act, int B, int T, void* stream) {
    return convstack_fwd(cfg, weff, x, ldx, y, ldy, act, B, T, (cudaStream_t)stream);
}
int crk_convstack_bwd(const crk_convstack_cfg* cfg, const float* theta, const float* weff,
                      const float* x, int ldx, const float* act, const float* dy, int lddy,
                      float* dx, int lddx, float dx_scale, float* gtheta, float* ws, int B, int T,
                      void* stream) {
    return convstack_bwd(cfg, theta, weff, x, ldx, act, dy, lddy, dx, lddx, dx_scale, gtheta, ws, B, T,
                         (cudaStream_t)stream);
}

// ---- vector quantiser ------------------------------------------------------------------------
int crk_vq_prepare(const float* W, float* WT, float* wn, int K, int D, void* stream) {
    if (!W || !WT || !wn || K < 1 || D < 1) return CRK_ERR_ARG;
    k_vq_prepare<<<cdiv(K, 128), 128, 0, (cudaStream_t)stream>>>(W, WT, wn, K, D);
    API_TRY(launch_check());
    return CRK_OK;
}
int crk_vq_argmin(const float* x, int ldx, const float* W, const float* WT, const float* wn,
                  long long* idx, float* e, int lde, float* qx, int ldqx, long long F, int K, int D,
                  void* stream) {
    if (!x || !W || !WT || !wn || !idx || !e || !qx || F < 1) return CRK_ERR_ARG;
    if (D != 64 || K < 128 || (K % 128) != 0) return CRK_ERR_UNSUPPORTED;
    VqArgminParams p;
    p.x = x; p.ldx = ldx; p.W = W; p.WT = WT; p.wn = wn; p.idx = idx; p.e = e; p.lde = lde;
    p.qx = qx; p.ldqx = ldqx; p.F = F; p.K = K;
    const size_t smem = (size_t)(64 * 64 + 64 * 128 + 64 + 64) * sizeof(float);
    static bool attr_set = false;
    if (!attr_set) {
        API_TRY(cudaFuncSetAttribute(k_vq_argmin, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
        attr_set = true;
    }
    TimedLaunch tl(CRK_K_VQ_ARGMIN, (cudaStream_t)stream, 2.0 * F * 64.0 * K);
    k_vq_argmin<<<(unsigned)cdivl(F, 64), CRK_THREADS, smem, (cudaStream_t)stream>>>(p);
    API_TRY(launch_check());
    return CRK_OK;
}
long long crk_vq_tc_blob_floats(int K, int D) {
    if (D != 64 || K < 128 || (K % 128) != 0) return -1;
    return (long long)(K / 128) * 2 * 16 * 129 * 4;
}
int crk_vq_pack_tc(const float* W, float* blob, int K, int D, void* stream) {
    if (!W || !blob || D != 64 || K < 128 || (K % 128) != 0) return CRK_ERR_ARG;
    k_vq_pack_tc<<<cdiv(K * 64, 256), 256, 0, (cudaStream_t)stream>>>(W, blob, K);
    API_TRY(launch_check());
    return CRK_OK;
}
int crk_vq_argmin_tc(const float* x, int ldx, const float* W, const float* blob, const float* wn,
                     long long* idx, float* e, int lde, float* qx, int ldqx, long long F, int K, int D,
                     void* stream) {
    if (!x || !W || !blob || !wn || !idx || !e || !qx || F < 1) return CRK_ERR_ARG;
    if (D != 64 || K < 128 || (K % 128) != 0 || K > 512) return CRK_ERR_UNSUPPORTED;
    VqTcParams q;
    q.p.x = x; q.p.ldx = ldx; q.p.W = W; q.p.WT = nullptr; q.p.wn = wn; q.p.idx = idx; q.p.e = e; q.p.lde = lde;
    q.p.qx = qx; q.p.ldqx = ldqx; q.p.F = F; q.p.K = K;
    q.blob = blob;
    static bool attr_set = false;
    if (!attr_set) {
        API_TRY(cudaFuncSetAttribute(k_vq_argmin_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, 225 * 1024));
        attr_set = true;
    }
    TimedLaunch tl(CRK_K_VQ_ARGMIN, (cudaStream_t)stream, 2.0 * F * 64.0 * K);
    k_vq_argmin_tc<<<(unsigned)cdivl(F, 128), 256, vq_tc_smem(), (cudaStream_t)stream>>>(q);
    API_TRY(launch_check());
    return CRK_OK;
}
long long crk_vq_stats_ws_floats(long long F, int K, int D) {
    if (F < 1 || K < 1 || D != 64) return -1;
    return (long long)vq_stats_chunks(F) * K * 65;
}
int crk_vq_stats(const float* x, int ldx, const long long* idx, float* counts, float* esum,
                 float* ws, long long F, int K, int D, void* stream) {
    if (!x || !idx || !counts || !esum || !ws || F < 1) return CRK_ERR_ARG;
    if (D != 64 || (size_t)K * 65 * sizeof(float) > 200 * 1024) return CRK_ERR_UNSUPPORTED;
    const int nch = vq_stats_chunks(F);
    VqStatsParams p;
    p.x = x; p.ldx = ldx; p.idx = idx; p.part = ws; p.F = F; p.K = K;
    p.frames_per_chunk = cdivl(cdivl(F, nch), 32) * 32;
    const int nchunk = (int)cdivl(F, p.frames_per_chunk);
    static bool attr_set = false;
    if (!attr_set) {
        API_TRY(cudaFuncSetAttribute(k_vq_stats, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        attr_set = true;
    }
    const size_t stats_smem = ((size_t)K * 64 + (((size_t)K + 3) & ~(size_t)3) + 2 * 32 * 64) * sizeof(float);
    if (stats_smem > 200 * 1024) return CRK_ERR_UNSUPPORTED;
    k_vq_stats<<<nchunk, CRK_THREADS, stats_smem, (cudaStream_t)stream>>>(p);
    API_TRY(launch_check());
    k_vq_stats_reduce<<<cdiv(K * 65, 256), 256, 0, (cudaStream_t)stream>>>(ws, nchunk, K, counts, esum);
    API_TRY(launch_check());
    return CRK_OK;
}
int crk_vq_ema(const float* counts, const float* esum, float* ema_size, float* ema_w, float* W,
               float decay, float eps, int K, int D, void* stream) {
    if (!counts || !esum || !ema_size || !ema_w || !W || K < 1 || D < 1) return CRK_ERR_ARG;
    // python-double scalars of the reference, rounded once to fp32 like torch does for scalar operands
    const float one_m_decay = (float)(1.0 - (double)decay);
    const float keps = (float)((double)K * (double)eps);
    k_vq_ema_size<<<1, 512, 0, (cudaStream_t)stream>>>(counts, ema_size, decay, one_m_decay, eps, keps, K);
    API_TRY(launch_check());
    k_vq_ema_w<<<cdiv(K * D, 256), 256, 0, (cudaStream_t)stream>>>(esum, ema_size, ema_w, W, decay, one_m_decay, K, D);
    API_TRY(launch_check());
    return CRK_OK;
}
// ---- round 2: single-pass TF32 argmin with a resident codebook, fused EMA (crk_vq_fast.cuh) ----
long long crk_vq_op_floats(int K, int D) {
    if (D != 64 || K < 128 || (K % 128) != 0 || K > 512) return -1;
    return vq_op_floats(K);
}
int crk_vq_pack_op(const float* W, float* opblob, int K, int D, void* stream) {
    if (!W || !opblob) return CRK_ERR_ARG;
    if (D != 64 || K < 128 || (K % 128) != 0 || K > 512) return CRK_ERR_UNSUPPORTED;
    k_vq_pack_op<<<cdiv(K, 128), 128, 0, (cudaStream_t)stream>>>(W, opblob, K);
    API_TRY(launch_check());
    return CRK_OK;
}
int crk_vq_argmin_fast(const float* x, int ldx, const float* opblob, long long* idx, float* e, int lde, float* qx,
                       int ldqx, long long F, int K, int D, void* stream) {
    if (!x || !opblob || !idx || !e || !qx || F < 1) return CRK_ERR_ARG;
    if (D != 64 || K < 128 || (K % 128) != 0 || K > 512) return CRK_ERR_UNSUPPORTED;
    VqFastParams q;
    q.p.x = x; q.p.ldx = ldx; q.p.W = nullptr; q.p.WT = nullptr; q.p.wn = nullptr; q.p.idx = idx; q.p.e = e; q.p.lde = lde;
    q.p.qx = qx; q.p.ldqx = ldqx; q.p.F = F; q.p.K = K;
    q.opblob = opblob;
    static bool attr_set = false;
    if (!attr_set) {
        // (dynamic + ~3 KB of static shared memory must stay below the 227 KB per-CTA limit)
        API_TRY(cudaFuncSetAttribute(k_vq_argmin_tf32, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)vq_fast_smem(512)));
        attr_set = true;
    }
    const long long tiles = cdivl(F, 128);
    const int sms = device_sm_count();
    TimedLaunch tl(CRK_K_VQ_ARGMIN, (cudaStream_t)stream, 2.0 * F * 64.0 * K);
    k_vq_argmin_tf32<<<(unsigned)(tiles < sms ? tiles : sms), CRK_VQ_THREADS, vq_fast_smem(K), (cudaStream_t)stream>>>(q);
    API_TRY(launch_check());
    return CRK_OK;
}
// statistics of one quantiser call into `stats` = [counts K | esum D*K | ticket (1 int, zeroed here)]
long long crk_vq_stats_floats(int K, int D) { return (K < 1 || D < 1) ? -1 : (long long)K + (long long)K * D + 4; }
int crk_vq_stats_fused(const float* x, int ldx, const long long* idx, float* stats, float* ws, long long F, int K, int D,
                       void* stream) {
    if (!stats) return CRK_ERR_ARG;
    int rc = crk_vq_stats(x, ldx, idx, stats, stats + K, ws, F, K, D, stream);
    if (rc != CRK_OK) return rc;
    API_TRY(cudaMemsetAsync(stats + (size_t)K + (size_t)K * D, 0, 4 * sizeof(float), (cudaStream_t)stream));
    return CRK_OK;
}
int crk_vq_ema_fused(float* stats, float* ema_size, float* ema_w, float* W, float* opblob, float decay, float eps, int K,
                     int D, void* stream) {
    if (!stats || !ema_size || !ema_w || !W || K < 1) return CRK_ERR_ARG;
    if (D != 64 || (K % 8) != 0 || K > 8192) return CRK_ERR_UNSUPPORTED;
    if (opblob && (K < 128 || (K % 128) != 0 || K > 512)) return CRK_ERR_UNSUPPORTED;
    const float one_m_decay = (float)(1.0 - (double)decay);
    const float keps = (float)((double)K * (double)eps);
    int* ticket = reinterpret_cast<int*>(stats + (size_t)K + (size_t)K * D);
    k_vq_ema_fused<<<K / 8, 512, (size_t)K * sizeof(float), (cudaStream_t)stream>>>(stats, stats + K, ema_size, ema_w, W, opblob,
                                                                                   ticket, decay, one_m_decay, eps, keps, K);
    API_TRY(launch_check());
    return CRK_OK;
}
int crk_vq_scatter_grad(const float* g, int ldg, const long long* idx, float* dW, long long F,
                        int K, int D, void* stream) {
    if (!g || !idx || !dW || F < 1 || K < 1 || D < 1) return CRK_ERR_ARG;
    k_vq_scatter_grad<<<(unsigned)cdivl(F * D, 256), 256, 0, (cudaStream_t)stream>>>(g, ldg, idx, dW, F, D);
    API_TRY(launch_check());
    return CRK_OK;
}

// ---- losses ----------------------------------------------------------------------------------
long long crk_masked_loss_ws_floats(int B, int T, int D) {
    if (B < 1 || T < 1 || D < 1) return -1;
    return 3LL * loss_blocks((long long)B * T * D);
}
static int masked_params(MaskedLossParams* p, const float* x, int ldx, const float* y, int ldy, float yconst,
                         const unsigned char* mask, int B, int T, int D, int shift) {
    if (!x || B < 1 || T < 1 || D < 1) return CRK_ERR_ARG;
    if ((shift < 0 ? -shift : shift) >= T) return CRK_ERR_ARG;
    p->x = x; p->ldx = ldx; p->y = y; p->ldy = ldy; p->yconst = yconst; p->mask = mask;
    p->B = B; p->T = T; p->D = D; p->shift = shift;
    return CRK_OK;
}
int crk_masked_loss_fwd(const float* x, int ldx, const float* y, int ldy, float yconst,
                        const unsigned char* mask, int B, int T, int D, int shift, float* out,
                        float* ws, void* stream) {
    MaskedLossParams p;
    int rc = masked_params(&p, x, ldx, y, ldy, yconst, mask, B, T, D, shift);
    if (rc) return rc;
    if (!out || !ws) return CRK_ERR_ARG;
    const int Tp = T - (shift < 0 ? -shift : shift);
    const int nblk = loss_blocks((long long)B * Tp * D);
    k_masked_loss_part<<<nblk, CRK_THREADS, 0, (cudaStream_t)stream>>>(p, ws);
    API_TRY(launch_check());
    k_finalize<<<1, CRK_THREADS, 0, (cudaStream_t)stream>>>(ws, nblk, 3, out, 2, 2);
    API_TRY(launch_check());
    return CRK_OK;
}
int crk_masked_loss_bwd(const float* x, int ldx, const float* y, int ldy, float yconst,
                        const unsigned char* mask, int B, int T, int D, int shift,
                        const float* out, const float* g_l1, const float* g_mse, float* dx, int lddx,
                        void* stream) {
    MaskedLossParams p;
    int rc = masked_params(&p, x, ldx, y, ldy, yconst, mask, B, T, D, shift);
    if (rc) return rc;
    if (!out || !dx) return CRK_ERR_ARG;
    const long long N = (long long)B * T * D;
    k_masked_loss_bwd<<<(unsigned)cdivl(N, CRK_THREADS), CRK_THREADS, 0, (cudaStream_t)stream>>>(p, out, g_l1, g_mse, dx, lddx);
    API_TRY(launch_check());
    return CRK_OK;
}

static int stft_params(StftParams* p, const float* x, int ldx, const float* y, int ldy, int B, int T, int D,
                       int n_fft, int hop, int win) {
    if (!x || !y || B < 1 || T < 1 || D < 1 || n_fft < 2 || hop < 1 || win < 1 || win > n_fft) return CRK_ERR_ARG;
    if (n_fft / 2 >= T) return CRK_ERR_ARG;          // reflect padding needs pad < T (torch.stft errors too)
    if (n_fft > 2048) return CRK_ERR_UNSUPPORTED;
    p->x = x; p->ldx = ldx; p->y = y; p->ldy = ldy; p->B = B; p->T = T; p->D = D;
    p->n_fft = n_fft; p->hop = hop; p->win = win;
    p->M = 1 + T / hop;
    p->bins = n_fft / 2 + 1;
    return CRK_OK;
}
long long crk_stft_loss_ws_floats(int B, int T, int D, int n_fft, int hop) {
    if (B < 1 || T < 1 || D < 1 || n_fft < 2 || hop < 1) return -1;
    const long long a = 2LL * loss_blocks((long long)B * (1 + T / hop) * (n_fft / 2 + 1) * D);
    const long long f = 2LL * B * (1 + T / hop);             // frame-per-CTA kernel: one partial pair per frame
    return a > f ? a : f;
}
int crk_stft_loss_fwd(const float* x, int ldx, const float* y, int ldy, int B, int T, int D,
                      int n_fft, int hop, int win, float* out, float* ws, void* stream) {
    StftParams p;
    int rc = stft_params(&p, x, ldx, y, ldy, B, T, D, n_fft, hop, win);
    if (rc) return rc;
    if (!out || !ws) return CRK_ERR_ARG;
    const long long N = (long long)B * p.M * p.bins * D;
    int nblk = loss_blocks(N);
    const size_t fsmem = stft_frame_smem_bytes(n_fft, win, D, false);
    if (fsmem <= 160 * 1024 && (long long)B * p.M <= (1 << 20)) {
        static bool attr = false;
        if (!attr) { API_TRY(cudaFuncSetAttribute(k_stft_loss_frame, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024)); attr = true; }
        nblk = B * p.M;
        k_stft_loss_frame<<<nblk, CRK_THREADS, fsmem, (cudaStream_t)stream>>>(p, ws);
    } else {
        const size_t smem = (size_t)3 * n_fft * sizeof(float);
        k_stft_loss_part<<<nblk, CRK_THREADS, smem, (cudaStream_t)stream>>>(p, ws);
    }
    API_TRY(launch_check());
    k_finalize<<<1, CRK_THREADS, 0, (cudaStream_t)stream>>>(ws, nblk, 2, out, 0, -1);
    API_TRY(launch_check());
    k_scale2<<<1, 1, 0, (cudaStream_t)stream>>>(out, 1.0f / (float)N);
    API_TRY(launch_check());
    return CRK_OK;
}
int crk_stft_loss_bwd(const float* x, int ldx, const float* y, int ldy, int B, int T, int D,
                      int n_fft, int hop, int win, const float* g, const float* glog, float scale, float* dx,
                      int lddx, int accumulate, void* stream) {
    StftParams p;
    int rc = stft_params(&p, x, ldx, y, ldy, B, T, D, n_fft, hop, win);
    if (rc) return rc;
    if ((!g && !glog) || !dx) return CRK_ERR_ARG;
    cudaStream_t s = (cudaStream_t)stream;
    if (!accumulate) {
        const long long rows = (long long)B * T;
        k_zero_panel<<<(unsigned)cdivl(rows * D, 256), 256, 0, s>>>(dx, lddx, D, rows);
        API_TRY(launch_check());
    }
    const size_t fsmem = stft_frame_smem_bytes(n_fft, win, D, true);
    if (fsmem <= 160 * 1024 && (long long)B * p.M <= (1 << 20)) {
        static bool attr = false;
        if (!attr) { API_TRY(cudaFuncSetAttribute(k_stft_loss_bwd_frame, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024)); attr = true; }
        k_stft_loss_bwd_frame<<<B * p.M, CRK_THREADS, fsmem, s>>>(p, g, glog, scale, dx, lddx);
    } else {
        const long long nfr = (long long)B * D * p.M;
        long long nblk = cdivl(nfr, 8);
        if (nblk > 148 * 8) nblk = 148 * 8;
        const size_t smem = (size_t)(3 * n_fft + 16 * p.bins) * sizeof(float);
        k_stft_loss_bwd<<<(unsigned)nblk, CRK_THREADS, smem, s>>>(p, g, glog, scale, dx, lddx);
    }
    API_TRY(launch_check());
    return CRK_OK;
}

long long crk_ce_ws_floats(long long F) {
    if (F < 1) return -1;
    return 2LL * loss_blocks(F * 8);
}
int crk_ce_fwd(const float* logits, int ldl, const long long* labels, long long F, int S,
               long long ignore_index, float* out, float* ws, void* stream) {
    if (!logits || !labels || !out || !ws || F < 1 || S < 1) return CRK_ERR_ARG;
    const int nblk = loss_blocks(F * 8);
    k_ce_part<<<nblk, CRK_THREADS, 0, (cudaStream_t)stream>>>(logits, ldl, labels, F, S, ignore_index, ws);
    API_TRY(launch_check());
    k_finalize<<<1, CRK_THREADS, 0, (cudaStream_t)stream>>>(ws, nblk, 2, out, 1, 1);
    API_TRY(launch_check());
    return CRK_OK;
}
int crk_ce_bwd(const float* logits, int ldl, const long long* labels, long long F, int S,
               long long ignore_index, const float* out, const float* g, float* dlogits, int lddl,
               void* stream) {
    if (!logits || !labels || !out || !g || !dlogits || F < 1 || S < 1) return CRK_ERR_ARG;
    k_ce_bwd<<<(unsigned)cdivl(F, CRK_THREADS), CRK_THREADS, 0, (cudaStream_t)stream>>>(logits, ldl, labels, F, S, ignore_index, out, g, dlogits, lddl);
    API_TRY(launch_check());
    return CRK_OK;
}

// ---- Adam ------------------------------------------------------------------------------------
int crk_adam_step(float* p, const float* g, float* m, float* v, long long n, float lr, float beta1,
                  float beta2, float eps, int step_count, void* stream) {
    if (!p || !g || !m || !v || n < 1 || step_count < 1) return CRK_ERR_ARG;
    // python-double bias corrections of torch.optim.Adam (_single_tensor_adam)
    const double bc1 = 1.0 - pow((double)beta1, (double)step_count);
    const double bc2 = 1.0 - pow((double)beta2, (double)step_count);
    const float step_size = (float)((double)lr / bc1);
    const float bc2_sqrt = (float)sqrt(bc2);
    k_adam<<<(unsigned)cdivl(n, 256), 256, 0, (cudaStream_t)stream>>>(p, g, m, v, n, lr, beta1, beta2, eps, step_size, bc2_sqrt);
    API_TRY(launch_check());
    return CRK_OK;
}

int crk_radam_step(float* p, const float* g, float* m, float* v, long long n, float lr, float beta1, float beta2,
                   float eps, int step_count, void* stream) {
    if (!p || !g || !m || !v || n < 1 || step_count < 1) return CRK_ERR_ARG;
    // python-double scalars of torch_optimizer.RAdam.step
    const double b1 = beta1, b2 = beta2, st = step_count;
    const double beta2_t = pow(b2, st);
    const double n_sma_max = 2.0 / (1.0 - b2) - 1.0;
    const double n_sma = n_sma_max - 2.0 * st * beta2_t / (1.0 - beta2_t);
    const int rect = n_sma >= 5.0;
    double step_size = (double)lr / (1.0 - pow(b1, st));
    if (rect)
        step_size = (double)lr * sqrt((1.0 - beta2_t) * (n_sma - 4.0) / (n_sma_max - 4.0) * (n_sma - 2.0) / n_sma * n_sma_max / (n_sma_max - 2.0)) /
                    (1.0 - pow(b1, st));
    k_radam<<<(unsigned)cdivl(n, 256), 256, 0, (cudaStream_t)stream>>>(p, g, m, v, n, beta1, beta2, eps, (float)step_size, rect);
    API_TRY(launch_check());
    return CRK_OK;
}

int crk_lamb_step(float* p, const float* g, float* m, float* v, float* upd, const long long* seg_off, const long long* seg_len,
                  int nseg, float* trust, float lr, float beta1, float beta2, float eps, void* stream) {
    if (!p || !g || !m || !v || !upd || !seg_off || !seg_len || !trust || nseg < 1) return CRK_ERR_ARG;
    k_lamb_moments<<<nseg, 256, 0, (cudaStream_t)stream>>>(p, g, m, v, upd, seg_off, seg_len, trust, beta1, beta2, eps);
    API_TRY(launch_check());
    k_lamb_apply<<<nseg, 256, 0, (cudaStream_t)stream>>>(p, upd, seg_off, seg_len, trust, lr);
    API_TRY(launch_check());
    return CRK_OK;
}

int crk_adam_step_dev(float* p, const float* g, float* m, float* v, long long n, float lr, float beta1,
                      float beta2, float eps, long long* step_dev, void* stream) {
    if (!p || !g || !m || !v || !step_dev || n < 1) return CRK_ERR_ARG;
    k_step_inc<<<1, 1, 0, (cudaStream_t)stream>>>(step_dev);
    API_TRY(launch_check());
    k_adam_dev<<<(unsigned)cdivl(n, 256), 256, 0, (cudaStream_t)stream>>>(p, g, m, v, n, lr, beta1, beta2, eps, step_dev);
    API_TRY(launch_check());
    return CRK_OK;
}

}  // extern "C"

// ---- log-mel front end -------------------------------------------------------------------------
namespace crk {

__global__ void k_frame_window(const float* __restrict__ wav, long long n_samples, const float* __restrict__ window,
                               int n_fft, int hop, int M, long long total, float* __restrict__ frames) {
    const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= total) return;
    const int n = (int)(e % n_fft);
    const long long fm = e / n_fft;
    const int m = (int)(fm % M);
    const long long b = fm / M;
    frames[e] = wav[b * n_samples + (long long)m * hop + n] * window[n];
}

// out[f][mel] = log10(max(eps, sum_bin |X[f][bin]| * basis[bin][mel])) (optionally standardised)
// CTA = 64 frames x 96 (>=n_mels) columns, K loop over bins in chunks of 64.
__global__ void __launch_bounds__(CRK_THREADS) k_mel(const float2* __restrict__ spec, int bins, const float* __restrict__ basis,
                                                     int n_mels, float eps, const float* __restrict__ mean,
                                                     const float* __restrict__ stdv, long long F, float* __restrict__ out) {
    __shared__ __align__(16) float as[64 * 64];
    __shared__ __align__(16) float bs[64 * 96];
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const long long f0 = (long long)blockIdx.x * 64;
    float acc[8][3];
#pragma unroll
    for (int i = 0; i < 8; ++i) { acc[i][0] = 0.f; acc[i][1] = 0.f; acc[i][2] = 0.f; }
    for (int k0 = 0; k0 < bins; k0 += 64) {
        __syncthreads();
        for (int i = threadIdx.x; i < 64 * 64; i += CRK_THREADS) {
            const int r = i >> 6, c = i & 63;
            float v = 0.f;
            if (f0 + r < F && k0 + c < bins) {
                const float2 z = spec[(size_t)(f0 + r) * bins + k0 + c];
                v = sqrtf(z.x * z.x + z.y * z.y);
            }
            as[i] = v;
        }
        for (int i = threadIdx.x; i < 64 * 96; i += CRK_THREADS) {
            const int r = i / 96, c = i - r * 96;
            bs[i] = (k0 + r < bins && c < n_mels) ? basis[(size_t)(k0 + r) * n_mels + c] : 0.f;
        }
        __syncthreads();
        tile_mac_rowA<3>(acc, as + ty * 8 * 64, 64, bs + tx * 3, 96, 64);
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const long long f = f0 + ty * 8 + i;
        if (f >= F) continue;
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const int mel = tx * 3 + c;
            if (mel >= n_mels) continue;
            float v = log10f(fmaxf(acc[i][c], eps));
            if (mean) v = (v - mean[mel]) / stdv[mel];
            out[(size_t)f * n_mels + mel] = v;
        }
    }
}

static std::map<std::pair<int, long long>, cufftHandle>& fft_plans() {
    static std::map<std::pair<int, long long>, cufftHandle> plans;
    return plans;
}

}  // namespace crk

extern "C" {

long long crk_logmel_ws_floats(int B, int n_frames, int n_fft) {
    if (B < 1 || n_frames < 1 || n_fft < 2) return -1;
    const long long F = (long long)B * n_frames;
    return F * n_fft + F * (n_fft / 2 + 1) * 2 + 8;
}

int crk_logmel_fwd(const float* wav, int B, long long n_samples, const float* window,
                   const float* mel_basis, int n_fft, int hop, int n_mels, float eps,
                   const float* mean, const float* stdv, float* out, float* ws, void* stream) {
    if (!wav || !window || !mel_basis || !out || !ws || B < 1 || n_fft < 2 || hop < 1) return CRK_ERR_ARG;
    if (n_samples < n_fft || n_mels < 1 || n_mels > 96) return CRK_ERR_ARG;
    cudaStream_t s = (cudaStream_t)stream;
    const int M = (int)(1 + (n_samples - n_fft) / hop);
    const long long F = (long long)B * M;
    const int bins = n_fft / 2 + 1;
    float* frames = ws;
    long long spec_off = F * n_fft;
    spec_off = (spec_off + 3) / 4 * 4;
    float2* spec = reinterpret_cast<float2*>(ws + spec_off);
    const long long total = F * n_fft;
    k_frame_window<<<(unsigned)cdivl(total, 256), 256, 0, s>>>(wav, n_samples, window, n_fft, hop, M, total, frames);
    API_TRY(launch_check());
    auto key = std::make_pair(n_fft, F);
    auto& plans = fft_plans();
    auto it = plans.find(key);
    if (it == plans.end()) {
        cufftHandle h;
        int n[1] = {n_fft};
        if (cufftPlanMany(&h, 1, n, nullptr, 1, n_fft, nullptr, 1, bins, CUFFT_R2C, (int)F) != CUFFT_SUCCESS)
            return CRK_ERR_CUDA;
        it = plans.emplace(key, h).first;
    }
    if (cufftSetStream(it->second, s) != CUFFT_SUCCESS) return CRK_ERR_CUDA;
    if (cufftExecR2C(it->second, frames, reinterpret_cast<cufftComplex*>(spec)) != CUFFT_SUCCESS) return CRK_ERR_CUDA;
    k_mel<<<(unsigned)cdivl(F, 64), CRK_THREADS, 0, s>>>(spec, bins, mel_basis, n_mels, eps, mean, stdv, F, out);
    API_TRY(launch_check());
    return CRK_OK;
}


// fused front end (crk_logmel.cuh): n_fft = 1024, banded mel projection.  `band_*`: for mel channel m the bins
// [band_start[m], band_start[m] + band_len[m]) carry its non-zero weights band_w[band_off[m] ...].
int crk_logmel_fused_fwd(const float* wav, int B, long long n_samples, const float* window, const int* band_start,
                         const int* band_len, const int* band_off, const float* band_w, int nnz, int n_fft, int hop,
                         int n_mels, float eps, const float* mean, const float* stdv, float* out, void* stream) {
    if (!wav || !window || !band_start || !band_len || !band_off || !band_w || !out || B < 1 || hop < 1) return CRK_ERR_ARG;
    if (n_fft != 1024 || n_mels < 1 || n_mels > 128 || nnz < 1 || nnz > CRK_MEL_MAXNNZ || hop > 512) return CRK_ERR_UNSUPPORTED;
    if (n_samples < n_fft) return CRK_ERR_ARG;
    LogmelParams p;
    p.wav = wav; p.n_samples = n_samples; p.window = window; p.band_start = band_start; p.band_len = band_len;
    p.band_off = band_off; p.band_w = band_w; p.nnz = nnz; p.hop = hop; p.n_mels = n_mels;
    p.M = (int)(1 + (n_samples - n_fft) / hop); p.eps = eps; p.mean = mean; p.stdv = stdv; p.out = out;
    const size_t smem = logmel_fused_smem(hop);
    static bool attr_set = false;
    if (!attr_set) {
        API_TRY(cudaFuncSetAttribute(k_logmel_fft1024, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
        attr_set = true;
    }
    const long long grid = (long long)B * ((p.M + CRK_MEL_FPC - 1) / CRK_MEL_FPC);
    TimedLaunch tl(CRK_K_LOGMEL, (cudaStream_t)stream, 0.0);
    k_logmel_fft1024<<<(unsigned)grid, 256, smem, (cudaStream_t)stream>>>(p);
    API_TRY(launch_check());
    return CRK_OK;
}

// backward of crk_logmel_fused_fwd w.r.t. the window and / or the waveform (learnable STFT windows, mlfb.py:72-90).
// `bin_*`: the transposed band table (per FFT bin 0..512 the run of mel channels whose filter covers it).
// dwin (1024 floats) and / or dwav (B x n_samples, accumulated into: the caller zero-initialises it) may be NULL;
// ws: crk_logmel_bwd_ws_floats(B, n_samples, hop) floats (per-CTA partial rows of d window).
long long crk_logmel_bwd_ws_floats(int B, long long n_samples, int hop) {
    if (B < 1 || hop < 1 || n_samples < 1024) return 0;
    const long long M = 1 + (n_samples - 1024) / hop;
    return (long long)B * ((M + CRK_MEL_FPC - 1) / CRK_MEL_FPC) * 1024;
}
int crk_logmel_fused_bwd(const float* wav, int B, long long n_samples, const float* window, const int* band_start,
                         const int* band_len, const int* band_off, const float* band_w, int nnz, const int* bin_start,
                         const int* bin_len, const int* bin_off, const float* bin_w, int n_fft, int hop, int n_mels,
                         float eps, const float* mean, const float* stdv, const float* dout, float* dwin, float* dwav,
                         float* ws, void* stream) {
    if (!wav || !window || !band_start || !band_len || !band_off || !band_w || !bin_start || !bin_len || !bin_off || !bin_w ||
        !dout || B < 1 || hop < 1 || (dwin && !ws))
        return CRK_ERR_ARG;
    if (n_fft != 1024 || n_mels < 1 || n_mels > 128 || nnz < 1 || nnz > CRK_MEL_MAXNNZ || hop > 512) return CRK_ERR_UNSUPPORTED;
    if (n_samples < n_fft) return CRK_ERR_ARG;
    LogmelBwdParams q;
    LogmelParams& p = q.f;
    p.wav = wav; p.n_samples = n_samples; p.window = window; p.band_start = band_start; p.band_len = band_len;
    p.band_off = band_off; p.band_w = band_w; p.nnz = nnz; p.hop = hop; p.n_mels = n_mels;
    p.M = (int)(1 + (n_samples - n_fft) / hop); p.eps = eps; p.mean = mean; p.stdv = stdv; p.out = nullptr;
    q.dout = dout; q.bin_start = bin_start; q.bin_len = bin_len; q.bin_off = bin_off; q.bin_w = bin_w;
    q.dwin_part = dwin ? ws : nullptr; q.dwav = dwav;
    static bool attr_set = false;
    if (!attr_set) {
        API_TRY(cudaFuncSetAttribute(k_logmel_bwd_fft1024, cudaFuncAttributeMaxDynamicSharedMemorySize, 128 * 1024));
        attr_set = true;
    }
    const long long grid = (long long)B * ((p.M + CRK_MEL_FPC - 1) / CRK_MEL_FPC);
    k_logmel_bwd_fft1024<<<(unsigned)grid, 256, logmel_bwd_smem(hop), (cudaStream_t)stream>>>(q);
    API_TRY(launch_check());
    if (dwin) {
        k_logmel_dwin_reduce<<<4, 256, 0, (cudaStream_t)stream>>>(ws, (int)grid, dwin);
        API_TRY(launch_check());
    }
    return CRK_OK;
}

}  // extern "C"
