/* crank_b200 -- C ABI of libcrank_b200.so (sm_100a CUDA kernels for crank's VQ-VAE train step).
 *
 * The reference (k2kobayashi/crank) is pure Python on PyTorch: it has no FFI layer.  Its seam is
 * the L2/L1 boundary of SURVEY.md section 8b -- Python classes whose arithmetic is stock torch ops.  This
 * header is what a maintainer binds (ctypes) to replace that arithmetic; each
 * entry point cites the reference code it replaces.
 *
 * Conventions
 *   - every function returns 0 (CRK_OK) or a negative error code; crk_strerror() names it.
 *     No exceptions cross the ABI.
 *   - all tensor arguments are DEVICE pointers owned by the caller (PyTorch allocations); the
 *     library never allocates persistent memory -- scratch sizes come from the *_floats() queries.
 *   - activations are channels-last fp32 panels (B*T, C) with an explicit row stride ("ld", floats).
 *   - `stream` is a cudaStream_t passed as void*; calls are asynchronous on it, capturable in CUDA
 *     graphs (no host syncs, no mallocs), re-entrant per stream.
 *   - single training thread per process, one process per GPU.
 */
#ifndef CRANK_B200_H
#define CRANK_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CRK_OK 0
#define CRK_ERR_ARG (-1)
#define CRK_ERR_CUDA (-2)
#define CRK_ERR_UNSUPPORTED (-3)

const char* crk_strerror(int code);
/* last CUDA error string seen by the library on this thread (for CRK_ERR_CUDA) */
const char* crk_last_cuda_error(void);
int crk_version(void);

/* arithmetic of the dense conv contractions (process-wide): 0 = fp32 CUDA cores, 1 = 3xTF32 on the
 * tcgen05 tensor cores (error-compensated, ~fp32 accuracy: the parity mode), 2 = plain TF32 (fast mode). */
int crk_set_precision(int mode);
int crk_get_precision(void);
/* debugging: force tensor-core kernel families back to the fp32 kernels
 * (bit 1 fused forward, 2 conv/dgrad, 4 wgrad, 8 gate backward) */
int crk_debug_tc_disable(int mask);
/* debugging / A-B measurement: switch optional optimisations off (results are unchanged up to the
 * summation order of bias gradients): bit 1 programmatic dependent launch, 2 bias column sums fused into
 * the tensor-core wgrad kernel, 4 128-bit epilogue of the tensor-core conv kernel, 8 shared-memory raw tile
 * feeding the k taps of the tensor-core wgrad kernel (off: one global fetch per tap).  Bit 4 also selects the
 * generic (all-options) instance of the conv kernel instead of the 128-bit-only one. */
int crk_debug_opt_disable(int mask);
/* opt-in experimental paths (parity-tested, not yet faster than the defaults): bit 1 persistent
 * warp-specialised fused residual-block forward (k_resblock_fwd_pt), 2 persistent pipelined conv / dgrad (k_conv_pt) */
int crk_debug_opt_enable(int mask);

/* instrumentation: number of kernels the library has launched in this process; optional CUDA-event
 * timing of one kernel family (ids: 1 resblock_fwd, 2 wgrad, 3 conv, 4 resblock_bwd_gate, 5 vq_argmin;
 * 0 disables).  crk_timing_read synchronises the device and returns (#launches, total ms) since enable. */
unsigned long long crk_launch_count(void);
/* performance debugging: the `launch_index`-th launch (counted from this call) of tensor-core kernel
 * family `kernel_id` (1 resblock_fwd, 2 wgrad, 3 conv) writes clock64() phase stamps into
 * device_buffer[gridDim][16] (NULL disables) */
int crk_debug_timestamps(long long* device_buffer, int kernel_id, int launch_index);
int crk_timing_enable(int kernel_id);
int crk_timing_read(int* count, float* total_ms);
/* algorithmic FLOPs (2*MACs of the real, unpadded contraction) of the launches timed since crk_timing_enable */
double crk_timing_flops(void);

/* tcgen05 probe: single-CTA TF32 GEMM through the tensor-core kernels' operand layout / descriptors /
 * TMEM path.  mode 0: D[m][n] = sum_k A[row_shift+m][k]*B[n][k]  (K-major operands, any row shift);
 * mode 1: D[m][n] = sum_f A[f][m]*B[f][n]  (MN-major operands, K = #frames).  M = 128, N in {64,128},
 * K % 8 == 0; split != 0 -> 3xTF32 error-compensated product.  D is (128, N) row-major. */
int crk_tc_probe(const float* A, int lda, int rowsA, const float* B, int ldb, int rowsB, float* D, int N,
                 int K, int row_shift, int mode, int split, void* stream);
/* tcgen05 MMA-rate microbenchmark (performance characterisation, profiles/mma_rate.py): `grid` CTAs each issue
 * reps x (split & 1 ? 3 : 1) x K/8 MMAs of shape 128 x N x 8 (TF32, N <= 256; split bit 1: A operand from
 * tensor memory, bit 2: commit after every group, bit 3: warp-collective issue; reps < 0: random data) on tiles in the conv kernels' operand
 * layout; cycles[2*cta] = clock64 span issue..completion, cycles[2*cta+1] = span of the issue loop alone. */
int crk_tc_mma_rate(int N, int K, int reps, int split, int grid, long long* cycles, void* stream);

/* ------------------------------------------------------------------------------------------
 * WaveNet stack = parallel_wavegan.models.ParallelWaveGANGenerator (upsample off) and
 * ResidualParallelWaveGANDiscriminator.   Replaces: crank/net/module/vqvae2.py:236-273 (encoders /
 * decoders), crank/bin/train.py:107-115 (residual discriminator D).
 * 64 residual / 128 gate / 64 skip channels are fixed (the reference never changes them).
 */
typedef struct {
    int in_ch, out_ch, aux_ch;   /* aux_ch <= 0: no conditioning */
    int layers, stacks;          /* dilation of layer l = 2^(l % (layers/stacks)) */
    int kernel_size;
    int causal;                  /* 0: "same" padding, 1: left padding (k-1)*d */
    int first_act;               /* 0: none (generator), 2: LeakyReLU after first 1x1 (discriminator) */
    int head_act;                /* 1: ReLU (generator), 2: LeakyReLU (discriminator) */
    float slope;                 /* LeakyReLU negative slope */
} crk_wavenet_cfg;

/* One weight-normalised Conv1d inside a parameter pack ("theta": [g | v | bias] per conv,
 * PyTorch weight_norm layout: weight_g (Cout,1,1), weight_v (Cout,Cin,k), bias (Cout)). */
typedef struct {
    int g_off, v_off, b_off;     /* offsets (floats) into theta; b_off < 0: no bias */
    int cout, cin, k;
    int w_off, bias_off;         /* offsets into the packed effective-weight buffer ("weff") */
    int cin_pad, ldw, perm;      /* fwd packing  W[j][cin_pad][ldw], column permutation id */
    int wt_off, wt_rows, ldwt;   /* transposed, tap-flipped copy for dgrad; wt_off < 0: none */
    /* tensor-core B-operand blobs (chunk-major, tf32 hi|lo split), one per tap:
     *   forward: rows n = packed output column (tc_n rows, UMMA N), K = input channel (tc_kpad)
     *   dgrad  : rows n = input channel (tct_n), K = packed output column (tct_kpad), taps flipped */
    int tc_off, tc_kpad, tc_n;
    int tct_off, tct_kpad, tct_n;
} crk_conv_desc;

#define CRK_MAX_CONVS 64

/* number of convs, canonical order: first_conv, then per layer {conv, [conv1x1_aux], conv1x1_out,
 * conv1x1_skip}, then last_conv_layers.1, last_conv_layers.3.  descs may be NULL. */
int crk_wavenet_describe(const crk_wavenet_cfg* cfg, crk_conv_desc* descs, int* n_convs,
                         long long* theta_floats, long long* weff_floats);
long long crk_wavenet_act_floats(const crk_wavenet_cfg* cfg, int B, int T);
long long crk_wavenet_ws_floats(const crk_wavenet_cfg* cfg, int B, int T);

/* weight-norm forward: theta -> weff  (w = g * v/||v||, packed + transposed copies).
 * Replaces torch._weight_norm evaluated at every conv call (511x per LSGAN step in the reference). */
int crk_wavenet_weights(const crk_wavenet_cfg* cfg, const float* theta, float* weff, void* stream);

/* forward.  x (B*T,in_ch) ld ldx; c (B*T,aux_ch) ld ldc or NULL; dropmul [layers][B*T][64]
 * dropout multipliers (mask/(1-p)) or NULL; y (B*T,out_ch) ld ldy; act: saved activations
 * (crk_wavenet_act_floats), needed by backward. */
int crk_wavenet_fwd(const crk_wavenet_cfg* cfg, const float* weff, const float* x, int ldx,
                    const float* c, int ldc, const float* dropmul, float* y, int ldy, float* act,
                    int B, int T, void* stream);

/* forward without the state backward needs (inference: reconstruction / eval / conversion, and the no-grad
 * generator passes of the GAN train steps): same arguments and result as crk_wavenet_fwd, `act` is scratch of the
 * same size, the per-layer (tanh, sigmoid) pairs are not written. */
int crk_wavenet_infer(const crk_wavenet_cfg* cfg, const float* weff, const float* x, int ldx,
                      const float* c, int ldc, const float* dropmul, float* y, int ldy, float* act,
                      int B, int T, void* stream);

/* backward.  dy (B*T,out_ch) ld lddy -> gtheta (same layout as theta, overwritten; NULL = parameters
 * frozen: only input gradients are computed and every weight-gradient kernel is skipped),
 * dx (B*T,in_ch) ld lddx or NULL, dc (B*T,aux_ch) ld lddc or NULL (overwritten).
 * ws: scratch of crk_wavenet_ws_floats() floats. */
int crk_wavenet_bwd(const crk_wavenet_cfg* cfg, const float* theta, const float* weff,
                    const float* x, int ldx, const float* c, int ldc, const float* dropmul,
                    const float* act, const float* dy, int lddy, float* dx, int lddx, float* dc,
                    int lddc, float* gtheta, float* ws, int B, int T, void* stream);

/* ------------------------------------------------------------------------------------------
 * Plain conv stack = parallel_wavegan.models.ParallelWaveGANDiscriminator (Conv1d + LeakyReLU).
 * Replaces: crank/bin/train.py:78-89 (speaker classifier C), crank/net/module/spkradv.py:49-60.
 */
typedef struct {
    int in_ch, out_ch, layers, kernel_size, conv_ch, dilation_factor;
    float slope;
} crk_convstack_cfg;

int crk_convstack_describe(const crk_convstack_cfg* cfg, crk_conv_desc* descs, int* n_convs,
                           long long* theta_floats, long long* weff_floats);
long long crk_convstack_act_floats(const crk_convstack_cfg* cfg, int B, int T);
long long crk_convstack_ws_floats(const crk_convstack_cfg* cfg, int B, int T);
int crk_convstack_weights(const crk_convstack_cfg* cfg, const float* theta, float* weff, void* stream);
int crk_convstack_fwd(const crk_convstack_cfg* cfg, const float* weff, const float* x, int ldx,
                      float* y, int ldy, float* act, int B, int T, void* stream);
/* dx_scale multiplies dx (gradient-reversal layer of crank/net/module/spkradv.py:63-72: -lambda);
 * gtheta may be NULL (frozen parameters: input gradient only) */
int crk_convstack_bwd(const crk_convstack_cfg* cfg, const float* theta, const float* weff,
                      const float* x, int ldx, const float* act, const float* dy, int lddy,
                      float* dx, int lddx, float dx_scale, float* gtheta, float* ws, int B, int T,
                      void* stream);

/* ------------------------------------------------------------------------------------------
 * Vector quantiser.  Replaces Quantizer.vq / Quantizer.forward, crank/net/module/vqvae2.py:306-347.
 *   crk_vq_prepare : codebook W (K,D) -> WT (D,K) and wn[k] = sum_d W[k][d]^2
 *   crk_vq_argmin  : idx[f] = argmin_k fl(fl(wn[k] - 2*dot(x_f, w_k)) + |x_f|^2)   (lowest index wins ties)
 *                    e[f] = W[idx[f]],  qx[f] = x[f] + (e[f] - x[f])   (straight-through value)
 *   crk_vq_stats   : counts[k] = #{f: idx[f]=k},  esum[d][k] = sum_{f: idx[f]=k} x[f][d]   (deterministic)
 *   crk_vq_ema     : EMA update + Laplace smoothing + new codebook (vqvae2.py:315-330)
 * D must be 64, K a multiple of 128 (the reference uses D=64, K=512).
 */
int crk_vq_prepare(const float* W, float* WT, float* wn, int K, int D, void* stream);
int crk_vq_argmin(const float* x, int ldx, const float* W, const float* WT, const float* wn,
                  long long* idx, float* e, int lde, float* qx, int ldqx, long long F, int K, int D,
                  void* stream);
/* tensor-core argmin (3xTF32 distance GEMM on tcgen05 + exact fp32 re-score of near-ties; same result as
 * crk_vq_argmin).  blob: crk_vq_tc_blob_floats() floats filled by crk_vq_pack_tc from the codebook. K <= 512 */
long long crk_vq_tc_blob_floats(int K, int D);
int crk_vq_pack_tc(const float* W, float* blob, int K, int D, void* stream);
int crk_vq_argmin_tc(const float* x, int ldx, const float* W, const float* blob, const float* wn,
                     long long* idx, float* e, int lde, float* qx, int ldqx, long long F, int K, int D,
                     void* stream);
long long crk_vq_stats_ws_floats(long long F, int K, int D);
int crk_vq_stats(const float* x, int ldx, const long long* idx, float* counts, float* esum,
                 float* ws, long long F, int K, int D, void* stream);
int crk_vq_ema(const float* counts, const float* esum, float* ema_size, float* ema_w, float* W,
               float decay, float eps, int K, int D, void* stream);
/* Round-2 quantiser path (crk_vq_fast.cuh), same reference code as crk_vq_argmin / crk_vq_ema
 * (crank/net/module/vqvae2.py:306-347): ONE plain-TF32 tensor-core pass over a codebook that stays resident in shared
 * memory + exact fp32 re-score of every code inside a rigorous error radius (identical indices to crk_vq_argmin), and
 * the EMA update in ONE launch that also writes the operand blob of the next call.
 *   opblob: crk_vq_op_floats(K, D) floats = raw fp32 codebook in the tensor-core operand layout | |w|^2.
 *   stats : crk_vq_stats_floats(K, D) floats = [counts K | per-code sums D*K (D-major, like ema_w) | ticket].
 * Data parallel: all-reduce(sum) the first K + D*K floats of `stats` between the two calls. */
long long crk_vq_op_floats(int K, int D);
int crk_vq_pack_op(const float* W, float* opblob, int K, int D, void* stream);
int crk_vq_argmin_fast(const float* x, int ldx, const float* opblob, long long* idx, float* e, int lde, float* qx,
                       int ldqx, long long F, int K, int D, void* stream);
long long crk_vq_stats_floats(int K, int D);
int crk_vq_stats_fused(const float* x, int ldx, const long long* idx, float* stats, float* ws, long long F, int K, int D,
                       void* stream);
int crk_vq_ema_fused(float* stats, float* ema_size, float* ema_w, float* W, float* opblob, float decay, float eps, int K,
                     int D, void* stream);
/* dW[k][d] += sum_{f: idx[f]=k} g[f][d]  (gradient of the codebook gather; only without EMA) */
int crk_vq_scatter_grad(const float* g, int ldg, const long long* idx, float* dW, long long F,
                        int K, int D, void* stream);

/* ------------------------------------------------------------------------------------------
 * Losses.  Each *_fwd writes a small fp32 record to `out` (device) and *_bwd reads the upstream
 * scalar gradient(s) from device memory -- no host synchronisation (the reference syncs via
 * masked_select and .item(), crank/net/trainer/trainer_vqvae.py:229-237, basetrainer.py:208-215).
 */
/* masked L1 / MSE with causal shift (crank/net/module/loss.py:30-47, trainer_vqvae.py:214-232,
 * trainer_lsgan.py:154-171).  x,y (B,T,D); y NULL => compare against the constant yconst;
 * mask (B,T) uint8 or NULL.  shift s>=0: x[:, s:] vs y[:, :T-s] with mask[:, s:];
 * s<0: x[:, :T+s] vs y[:, -s:] with mask[:, :T+s].  out[0]=mean|d|, out[1]=mean d^2, out[2]=#elements. */
int crk_masked_loss_fwd(const float* x, int ldx, const float* y, int ldy, float yconst,
                        const unsigned char* mask, int B, int T, int D, int shift, float* out,
                        float* ws, void* stream);
long long crk_masked_loss_ws_floats(int B, int T, int D);
/* dx = g_l1[0]*sign(d)/n + g_mse[0]*2d/n on selected elements, 0 elsewhere (g_* may be NULL) */
int crk_masked_loss_bwd(const float* x, int ldx, const float* y, int ldy, float yconst,
                        const unsigned char* mask, int B, int T, int D, int shift,
                        const float* out, const float* g_l1, const float* g_mse, float* dx, int lddx,
                        void* stream);

/* STFT-magnitude trajectory L1 (crank/net/module/loss.py:50-85): each feature dimension's time
 * trajectory -> STFT(n_fft, hop, hann(win) centred in n_fft, center=True reflect) ->
 * sqrt(clamp(re^2+im^2, 1e-7)) -> out[0] = mean |mag_x - mag_y|, out[1] = mean |log mag_x - log mag_y|
 * (STFTLoss.forward combines them as (1 - logratio) * out[0] + logratio * out[1], loss.py:80-84). */
long long crk_stft_loss_ws_floats(int B, int T, int D, int n_fft, int hop);
int crk_stft_loss_fwd(const float* x, int ldx, const float* y, int ldy, int B, int T, int D,
                      int n_fft, int hop, int win, float* out, float* ws, void* stream);
/* dx (+)= scale * (g[0] * d out[0]/dx + glog[0] * d out[1]/dx); g or glog may be NULL (term absent);
 * accumulate != 0 adds into dx */
int crk_stft_loss_bwd(const float* x, int ldx, const float* y, int ldy, int B, int T, int D,
                      int n_fft, int hop, int win, const float* g, const float* glog, float scale, float* dx,
                      int lddx, int accumulate, void* stream);

/* cross entropy with ignore_index (torch.nn.CrossEntropyLoss(ignore_index=-100),
 * crank/net/trainer/utils.py:26).  logits (F,S) ld ldl, S<=64; out[0]=loss, out[1]=#valid rows */
long long crk_ce_ws_floats(long long F);
int crk_ce_fwd(const float* logits, int ldl, const long long* labels, long long F, int S,
               long long ignore_index, float* out, float* ws, void* stream);
int crk_ce_bwd(const float* logits, int ldl, const long long* labels, long long F, int S,
               long long ignore_index, const float* out, const float* g, float* dlogits, int lddl,
               void* stream);

/* ------------------------------------------------------------------------------------------
 * Adam (torch.optim.Adam defaults, crank/net/trainer/utils.py:43; step_model trainer_vqvae.py:200-208).
 * state: m, v same size as p; step_count is the 1-based step number AFTER this update. */
int crk_adam_step(float* p, const float* g, float* m, float* v, long long n, float lr, float beta1,
                  float beta2, float eps, int step_count, void* stream);
/* The other two optimizers crank/net/trainer/utils.py:40-58 can build.  crk_radam_step: torch_optimizer.RAdam (lr, betas
 * (0.9, 0.999), eps 1e-8; rectified update once the variance is tractable, plain momentum step before).  crk_lamb_step:
 * pytorch_lamb.Lamb (eps 1e-6, no bias correction, weight norm clamped to [0, 10]); the trust ratio is per parameter
 * tensor of the reference, i.e. per segment [seg_off[i], seg_off[i] + seg_len[i]) of the flat pack (device arrays of
 * nseg int64); upd (n floats) and trust (nseg floats) are scratch. */
int crk_radam_step(float* p, const float* g, float* m, float* v, long long n, float lr, float beta1, float beta2,
                   float eps, int step_count, void* stream);
int crk_lamb_step(float* p, const float* g, float* m, float* v, float* upd, const long long* seg_off, const long long* seg_len,
                  int nseg, float* trust, float lr, float beta1, float beta2, float eps, void* stream);
/* Same update with the step count in DEVICE memory (*step_dev is incremented first, then used for the bias
 * corrections): no per-step host value in the launch arguments, so the call can be captured in a CUDA graph and
 * replayed (crank_b200/net/graph.py). */
int crk_adam_step_dev(float* p, const float* g, float* m, float* v, long long n, float lr, float beta1,
                      float beta2, float eps, long long* step_dev, void* stream);

/* ------------------------------------------------------------------------------------------
 * Log-mel front end (crank/net/module/mlfb.py:134-171 online; crank/feature/feature.py:126-145 offline):
 * frames of n_fft samples every `hop`, times window, |rFFT|, x mel basis (n_bins x n_mels), clamp eps,
 * log10, optional (x-mean)/std.   wav (B, n_samples) already padded by the caller (center handling is
 * host-side index arithmetic); n_frames = 1 + (n_samples - n_fft)/hop. */
long long crk_logmel_ws_floats(int B, int n_frames, int n_fft);
int crk_logmel_fwd(const float* wav, int B, long long n_samples, const float* window,
                   const float* mel_basis, int n_fft, int hop, int n_mels, float eps,
                   const float* mean, const float* stdv, float* out, float* ws, void* stream);
/* fused front end, one kernel and one pass over the waveform (n_fft = 1024; else CRK_ERR_UNSUPPORTED -> use
 * crk_logmel_fwd): framing + window + radix-4 FFT (two real frames per complex transform) + |.| + BANDED mel projection +
 * clamp / log10 / optional scaler.  Replaces the same reference code as crk_logmel_fwd (crank/net/module/mlfb.py:134-171,
 * crank/feature/feature.py:126-145).  The mel basis is passed by its non-zero runs: mel channel m sums bins
 * [band_start[m], band_start[m] + band_len[m]) with weights band_w[band_off[m] ...] (nnz <= 2048 in total). */
int crk_logmel_fused_fwd(const float* wav, int B, long long n_samples, const float* window, const int* band_start,
                         const int* band_len, const int* band_off, const float* band_w, int nnz, int n_fft, int hop,
                         int n_mels, float eps, const float* mean, const float* stdv, float* out, void* stream);
/* backward of crk_logmel_fused_fwd w.r.t. the STFT window (dwin: 1024 floats, overwritten) and / or the waveform
 * (dwav: B x n_samples, ACCUMULATED into -- zero it first); either may be NULL.  Serves the learnable windows of
 * crank/net/module/mlfb.py:72-90 ("param": the window is an nn.Parameter; "conv": a learnable pre-filter feeds the
 * STFT, so the gradient must reach the waveform), whose backward the reference gets from autograd through torch.stft.
 * bin_*: transposed band table (per FFT bin 0..512 the run [bin_start, bin_start + bin_len) of mel channels whose
 * filter covers the bin, weights bin_w[bin_off ...]).  ws: crk_logmel_bwd_ws_floats() floats. */
long long crk_logmel_bwd_ws_floats(int B, long long n_samples, int hop);
int crk_logmel_fused_bwd(const float* wav, int B, long long n_samples, const float* window, const int* band_start,
                         const int* band_len, const int* band_off, const float* band_w, int nnz, const int* bin_start,
                         const int* bin_len, const int* bin_off, const float* bin_w, int n_fft, int hop, int n_mels,
                         float eps, const float* mean, const float* stdv, const float* dout, float* dwin, float* dwav,
                         float* ws, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* CRANK_B200_H */
