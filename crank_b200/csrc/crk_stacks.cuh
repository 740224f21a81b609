// crank-b200: parameter packing (weight-norm) and the two conv-stack programs
// (WaveNet stack, plain LeakyReLU conv stack) built from the kernels in crk_conv.cuh /
// crk_resblock.cuh.  Host-side orchestration only launches kernels on the caller's stream.
#pragma once
#include <string.h>
#include "../../include/crank_b200.h"
#include "crk_common.cuh"
#include "crk_conv.cuh"
#include "crk_resblock.cuh"
#include "crk_resblock_tc.cuh"
#include "crk_resblock_pt.cuh"
#include "crk_resblock_tc2.cuh"
#include "crk_conv_tc.cuh"
#include "crk_conv_pt.cuh"
#include "crk_wgrad_tc.cuh"

namespace crk {

#define CRK_TRY(expr)                                   \
    do {                                                \
        cudaError_t _e = (expr);                        \
        if (_e != cudaSuccess) return set_cuda_error(_e); \
    } while (0)

int set_cuda_error(cudaError_t e);  // defined in crk_api.cu

// packed column of output channel co under permutation `perm`
__host__ __device__ inline int pcol(int co, int perm) {
    switch (perm) {
        case 1: return co < 64 ? 4 * (co >> 1) + (co & 1) : 4 * ((co - 64) >> 1) + 2 + (co & 1);
        case 2: return 4 * (co >> 1) + (co & 1);
        case 3: return 4 * (co >> 1) + 2 + (co & 1);
        default: return co;
    }
}

struct DescTable {
    int n;
    crk_conv_desc d[CRK_MAX_CONVS];
};

// ---- weight norm -----------------------------------------------------------------------------
// forward: w = v * (g / ||v||)  per output channel (torch._weight_norm, dim=0), scattered into the
// packed fwd layout W[j][ci][pcol] and the transposed tap-flipped layout WT[k-1-j][pcol][ci].
__global__ void __launch_bounds__(128) k_weightnorm_fwd(const DescTable tab, const float* __restrict__ theta,
                                                         float* __restrict__ weff) {
    pdl_trigger();
    pdl_wait();
    const crk_conv_desc d = tab.d[blockIdx.y];
    const int co = blockIdx.x;
    if (co >= d.cout) return;
    const int n = d.cin * d.k;
    const float* v = theta + d.v_off + (size_t)co * n;
    __shared__ float red[4];
    float s = 0.f;
    for (int e = threadIdx.x; e < n; e += 128) s += v[e] * v[e];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
    __syncthreads();
    const float norm = sqrtf(red[0] + red[1] + red[2] + red[3]);
    const float scale = theta[d.g_off + co] / norm;
    const int pc = pcol(co, d.perm);
    for (int e = threadIdx.x; e < n; e += 128) {
        const int ci = e / d.k, j = e - ci * d.k;
        const float w = v[e] * scale;
        weff[d.w_off + ((size_t)j * d.cin_pad + ci) * d.ldw + pc] = w;
        if (d.wt_off >= 0)
            weff[d.wt_off + ((size_t)(d.k - 1 - j) * d.wt_rows + pc) * d.ldwt + ci] = w;
        if (d.tc_off >= 0) {
            float hi, lo;
            tc::split_tf32(w, hi, lo);
            // forward B operand of tap j: [K chunk (ci/4)][n = packed column][ci%4], hi then lo
            const int half = tc_blob_half(d.tc_kpad, d.tc_n);
            const size_t o = (size_t)d.tc_off + (size_t)j * 2 * half + (size_t)(ci >> 2) * tc::chunk_rows(d.tc_n) * 4 + pc * 4 + (ci & 3);
            weff[o] = hi;
            weff[o + half] = lo;
            // dgrad B operand of flipped tap k-1-j: [K chunk (pc/4)][n = ci][pc%4]
            const int halft = tc_blob_half(d.tct_kpad, d.tct_n);
            const size_t ot = (size_t)d.tct_off + (size_t)(d.k - 1 - j) * 2 * halft + (size_t)(pc >> 2) * tc::chunk_rows(d.tct_n) * 4 + ci * 4 + (pc & 3);
            weff[ot] = hi;
            weff[ot + halft] = lo;
        }
    }
    if (threadIdx.x == 0 && d.b_off >= 0) weff[d.bias_off + pc] = theta[d.b_off + co];
}

// backward (aten::_weight_norm_interface_backward):  a = g/||v||, dot = <gw, v>,
//   grad_v = a*gw - (a*dot/||v||^2) * v,   grad_g = dot/||v||,   grad_bias = db
__global__ void __launch_bounds__(128) k_weightnorm_bwd(const DescTable tab, const float* __restrict__ theta,
                                                         const float* __restrict__ gweff,
                                                         float* __restrict__ gtheta) {
    pdl_trigger();
    pdl_wait();
    const crk_conv_desc d = tab.d[blockIdx.y];
    const int co = blockIdx.x;
    if (co >= d.cout) return;
    const int n = d.cin * d.k;
    const float* v = theta + d.v_off + (size_t)co * n;
    const int pc = pcol(co, d.perm);
    __shared__ float red[2][4];
    float s = 0.f, dt = 0.f;
    for (int e = threadIdx.x; e < n; e += 128) {
        const int ci = e / d.k, j = e - ci * d.k;
        const float gw = gweff[d.w_off + ((size_t)j * d.cin_pad + ci) * d.ldw + pc];
        s += v[e] * v[e];
        dt += gw * v[e];
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        s += __shfl_xor_sync(0xffffffffu, s, o);
        dt += __shfl_xor_sync(0xffffffffu, dt, o);
    }
    if ((threadIdx.x & 31) == 0) { red[0][threadIdx.x >> 5] = s; red[1][threadIdx.x >> 5] = dt; }
    __syncthreads();
    const float nsq = red[0][0] + red[0][1] + red[0][2] + red[0][3];
    const float dot = red[1][0] + red[1][1] + red[1][2] + red[1][3];
    const float norm = sqrtf(nsq);
    const float a = theta[d.g_off + co] / norm;
    const float bcoef = a * dot / nsq;
    for (int e = threadIdx.x; e < n; e += 128) {
        const int ci = e / d.k, j = e - ci * d.k;
        const float gw = gweff[d.w_off + ((size_t)j * d.cin_pad + ci) * d.ldw + pc];
        gtheta[d.v_off + (size_t)co * n + e] = a * gw - bcoef * v[e];
    }
    if (threadIdx.x == 0) {
        gtheta[d.g_off + co] = dot / norm;
        if (d.b_off >= 0) gtheta[d.b_off + co] = gweff[d.bias_off + pc];
    }
}

inline cudaError_t launch_weightnorm(const DescTable& tab, const float* theta, float* weff, cudaStream_t s) {
    int maxc = 1;
    for (int i = 0; i < tab.n; ++i) maxc = tab.d[i].cout > maxc ? tab.d[i].cout : maxc;
    cudaError_t e = launch_pdl(k_weightnorm_fwd, dim3(maxc, tab.n), dim3(128), 0, s, tab, theta, weff);
    if (e != cudaSuccess) return e;
    return launch_check();
}
inline cudaError_t launch_weightnorm_bwd(const DescTable& tab, const float* theta, const float* gweff,
                                         float* gtheta, cudaStream_t s) {
    int maxc = 1;
    for (int i = 0; i < tab.n; ++i) maxc = tab.d[i].cout > maxc ? tab.d[i].cout : maxc;
    cudaError_t e = launch_pdl(k_weightnorm_bwd, dim3(maxc, tab.n), dim3(128), 0, s, tab, theta, gweff, gtheta);
    if (e != cudaSuccess) return e;
    return launch_check();
}

// ---- layout builder ---------------------------------------------------------------------------
struct LayoutBuilder {
    DescTable tab;
    long long theta = 0, weff = 0;
    LayoutBuilder() { tab.n = 0; }
    // returns index; share >= 0: reuse the packed W/bias/WT slots of conv `share` (out|skip pair)
    int add(int cout, int cin, int k, bool bias, int perm, int ldw, int wt_rows, int ldwt, int share = -1) {
        crk_conv_desc& d = tab.d[tab.n];
        d.cout = cout; d.cin = cin; d.k = k;
        d.g_off = (int)theta; theta += cout;
        d.v_off = (int)theta; theta += (long long)cout * cin * k;
        if (bias) { d.b_off = (int)theta; theta += cout; } else d.b_off = -1;
        d.cin_pad = round_up(cin, 4); d.ldw = ldw; d.perm = perm;
        d.wt_rows = wt_rows; d.ldwt = ldwt;
        d.tc_kpad = round_up(cin, 8); d.tc_n = perm != 0 ? 128 : round_up(cout, 16);
        d.tct_kpad = perm != 0 ? 128 : round_up(cout, 8); d.tct_n = round_up(cin, 16);
        if (share >= 0) {
            d.w_off = tab.d[share].w_off; d.bias_off = tab.d[share].bias_off; d.wt_off = tab.d[share].wt_off;
            d.tc_off = tab.d[share].tc_off; d.tct_off = tab.d[share].tct_off;
        } else {
            d.w_off = (int)weff; weff += (long long)k * d.cin_pad * ldw;
            d.bias_off = (int)weff; weff += ldw;
            d.wt_off = (int)weff; weff += (long long)k * wt_rows * ldwt;
            d.tc_off = (int)weff; weff += (long long)k * 2 * tc_blob_half(d.tc_kpad, d.tc_n);
            d.tct_off = (int)weff; weff += (long long)k * 2 * tc_blob_half(d.tct_kpad, d.tct_n);
        }
        return tab.n++;
    }
};

// ---- WaveNet stack ----------------------------------------------------------------------------
struct WavenetLayout {
    DescTable tab;
    long long theta, weff;
    int first, last1, last2;
    int conv[32], aux[32], out[32], skip[32];
};

inline int wavenet_layout(const crk_wavenet_cfg* c, WavenetLayout* L) {
    if (!c || c->layers < 1 || c->layers > 32 || c->stacks < 1 || c->layers % c->stacks != 0) return CRK_ERR_ARG;
    if (c->in_ch < 1 || c->in_ch > CRK_MAX_IN_CH || c->out_ch < 1 || c->out_ch > 128 || c->aux_ch > 128) return CRK_ERR_ARG;
    if (c->kernel_size < 1 || c->kernel_size > 9) return CRK_ERR_ARG;
    if (!c->causal && (c->kernel_size % 2) == 0) return CRK_ERR_ARG;
    const int n_convs = 1 + c->layers * (c->aux_ch > 0 ? 4 : 3) + 2;
    if (n_convs > CRK_MAX_CONVS) return CRK_ERR_UNSUPPORTED;
    LayoutBuilder b;
    L->first = b.add(64, c->in_ch, 1, true, 0, 64, 64, wide_ld(c->in_ch));
    for (int l = 0; l < c->layers; ++l) {
        L->conv[l] = b.add(128, 64, c->kernel_size, true, 1, 128, 128, 64);
        L->aux[l] = c->aux_ch > 0 ? b.add(128, c->aux_ch, 1, false, 1, 128, 128, 32 * cpt_for(c->aux_ch)) : -1;
        L->out[l] = b.add(64, 64, 1, true, 2, 128, 128, 64);
        L->skip[l] = b.add(64, 64, 1, true, 3, 128, 128, 64, L->out[l]);
    }
    L->last1 = b.add(64, 64, 1, true, 0, 64, 64, 64);
    L->last2 = b.add(c->out_ch, 64, 1, true, 0, 32 * cpt_for(c->out_ch), round_up(c->out_ch, 4), 64);
    L->tab = b.tab; L->theta = b.theta; L->weff = b.weff;
    return CRK_OK;
}

inline int wn_dilation(const crk_wavenet_cfg* c, int l) { return 1 << (l % (c->layers / c->stacks)); }
inline int wn_padl(const crk_wavenet_cfg* c, int dil) {
    return c->causal ? (c->kernel_size - 1) * dil : (c->kernel_size - 1) / 2 * dil;
}

struct WavenetAct {   // offsets (floats) into the saved-activation buffer
    long long h, tasb, skips, head1, total;
};
inline WavenetAct wavenet_act(const crk_wavenet_cfg* c, long long F) {
    WavenetAct a;
    a.h = 0;
    a.tasb = a.h + (long long)(c->layers + 1) * F * 64;
    a.skips = a.tasb + (long long)c->layers * F * 128;
    a.head1 = a.skips + F * 64;
    a.total = a.head1 + F * 64;
    return a;
}
// ---- weight gradients on a side stream (round 2) ------------------------------------------------------------------
// With few tiles per launch (8 utterances per GPU = 32 tiles on 148 SMs: BASELINE config 3 at 8 GPUs) a step is a chain of
// ~900 dependent kernels of ~13 us each, most SMs idle.  The weight-gradient kernels (wgrad + partial-sum reduce: 4-6 of
// the ~8 launches of a block's backward) do not feed the dgrad chain, so they run on a side stream: fork after the kernel
// that produces their operands (an event), join before the weight-norm backward.  Needs the per-layer operands (dg, gos,
// z) to stay alive until their wgrad ran: one buffer per layer instead of one per stack (41 MB per layer at 32 000 frames).
// Measured: 8 utterances per GPU 12.0 -> 10.0 ms graphed; 64 utterances per GPU 19.5 -> 19.1 ms (the wgrad CTAs, one per SM,
// fill the SMs a 256-tile dgrad / gate launch at two CTAs per SM leaves idle).
struct SideStream {
    cudaStream_t st = nullptr;
    cudaEvent_t ev[64];
    bool ok = false;
};
inline SideStream* side_stream() {
    static thread_local SideStream per_dev[16];           // one per device of this thread (streams belong to a device)
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 16) return nullptr;
    SideStream& S = per_dev[dev];
    if (!S.ok) {
        if (cudaStreamCreateWithFlags(&S.st, cudaStreamNonBlocking) != cudaSuccess) return nullptr;
        for (int i = 0; i < 64; ++i)
            if (cudaEventCreateWithFlags(&S.ev[i], cudaEventDisableTiming) != cudaSuccess) return nullptr;
        S.ok = true;
    }
    return &S;
}
inline bool wavenet_side_mode(int B, int T) {
    (void)B; (void)T;
    return !(opt_disable_mask() & 512);
}
// `to` waits for everything `from` has been given so far
inline cudaError_t stream_fork(SideStream* S, int& evn, cudaStream_t from, cudaStream_t to) {
    cudaEvent_t e = S->ev[evn++ & 63];
    cudaError_t rc = cudaEventRecord(e, from);
    if (rc != cudaSuccess) return rc;
    return cudaStreamWaitEvent(to, e, 0);
}

struct WavenetWs {
    long long gweff, dhA, dhB, ds, dg, gos, z, dhead1, part, total;
    int nb;             // per-layer copies of dg / gos / z (1, or `layers` in side-stream mode)
};
inline WavenetWs wavenet_ws(const crk_wavenet_cfg* c, const WavenetLayout& L, int B, int T) {
    const long long F = (long long)B * T;
    WavenetWs w;
    w.nb = wavenet_side_mode(B, T) ? c->layers : 1;
    w.gweff = 0;
    w.dhA = round_up((int)L.weff, 4);
    w.dhB = w.dhA + F * 64;
    w.ds = w.dhB + F * 64;
    w.dg = w.ds + F * 64;
    w.gos = w.dg + (long long)w.nb * F * 128;
    w.z = w.gos + (long long)w.nb * F * 128;
    w.dhead1 = w.z + (long long)w.nb * F * 64;
    w.part = w.dhead1 + F * 64;
    size_t m = 0;
    for (int i = 0; i < L.tab.n; ++i) {
        const crk_conv_desc& d = L.tab.d[i];
        const size_t q = wgrad_part_floats(B, T, d.k, d.cin_pad, d.ldw);
        m = q > m ? q : m;
    }
    w.total = w.part + (long long)m;
    return w;
}

// save_gates = false: inference / no-grad forward -- the (tanh, sigmoid) pairs backward would need are not
// written (512 B per frame and layer, 40% of the fused block's output traffic)
inline int wavenet_fwd(const crk_wavenet_cfg* c, const float* weff, const float* x, int ldx,
                       const float* cond, int ldc, const float* dropmul, float* y, int ldy, float* act,
                       int B, int T, cudaStream_t s, bool save_gates = true) {
    WavenetLayout L;
    int rc = wavenet_layout(c, &L);
    if (rc) return rc;
    if (!weff || !x || !y || !act || B < 1 || T < 1) return CRK_ERR_ARG;
    if (c->aux_ch > 0 && !cond) return CRK_ERR_ARG;
    const long long F = (long long)B * T;
    const WavenetAct A = wavenet_act(c, F);
    float* h = act + A.h;
    // first 1x1
    {
        const crk_conv_desc& d = L.tab.d[L.first];
        ConvParams p = conv_params_default();
        p.X = x; p.ldx = ldx; p.Cin = c->in_ch; p.CinPad = d.cin_pad;
        p.W = weff + d.w_off; p.bias = weff + d.bias_off;
        p.Y = h; p.ldy = 64; p.Cout = 64; p.B = B; p.T = T;
        p.epi_act = c->first_act; p.epi_slope = c->slope;
        CRK_TRY(conv_dispatch(p, 2, weff + d.tc_off, d.tc_kpad, d.tc_n, s));
    }
    for (int l = 0; l < c->layers; ++l) {
        const int dil = wn_dilation(c, l);
        ResFwdParams p;
        p.Hin = h + (long long)l * F * 64;
        p.Hout = h + (long long)(l + 1) * F * 64;
        p.Skip = act + A.skips; p.skip_init = (l == 0);
        p.Wc = weff + L.tab.d[L.conv[l]].w_off; p.bc = weff + L.tab.d[L.conv[l]].bias_off;
        if (c->aux_ch > 0) {
            p.Caux = cond; p.ldc = ldc; p.Ca = c->aux_ch; p.CaPad = round_up(c->aux_ch, 4);
            p.Wa = weff + L.tab.d[L.aux[l]].w_off;
        } else { p.Caux = nullptr; p.ldc = 0; p.Ca = 0; p.CaPad = 0; p.Wa = nullptr; }
        p.Wos = weff + L.tab.d[L.out[l]].w_off; p.bos = weff + L.tab.d[L.out[l]].bias_off;
        p.dropmul = dropmul ? dropmul + (long long)l * F * 64 : nullptr;
        p.TaSb = save_gates ? act + A.tasb + (long long)l * F * 128 : nullptr;
        p.B = B; p.T = T; p.k = c->kernel_size; p.dil = dil; p.padl = wn_padl(c, dil);
        const int mode = precision_mode();
        if (mode == CRK_PREC_FP32 || (tc_disable_mask() & 1) || (c->kernel_size - 1) * dil > 16) {
            CRK_TRY(launch_resblock_fwd(p, s));
        } else {
            ResFwdTcParams q;
            q.p = p;
            q.WcTc = weff + L.tab.d[L.conv[l]].tc_off;
            q.WosTc = weff + L.tab.d[L.out[l]].tc_off;
            q.WaTc = c->aux_ch > 0 ? weff + L.tab.d[L.aux[l]].tc_off : nullptr;
            q.KaPad = c->aux_ch > 0 ? L.tab.d[L.aux[l]].tc_kpad : 0;
            const bool split = mode == CRK_PREC_TF32X3;
            if ((opt_enable_mask() & 1) && resblock_fwd_pt_ok(q, split)) {       // persistent pipelined kernel (round 2)
                if (split) CRK_TRY(launch_resblock_fwd_pt<true>(q, s));
                else CRK_TRY(launch_resblock_fwd_pt<false>(q, s));
            } else if (!(opt_disable_mask() & 64) && resblock_fwd_tc2_ok(q, split) &&
                       B * cdiv(T, CRK_TC_TM) > device_sm_count()) {   // 2-CTA/SM kernel (round 2): pays when tiles outnumber SMs
                if (split) CRK_TRY(launch_resblock_fwd_tc2<true>(q, s));
                else CRK_TRY(launch_resblock_fwd_tc2<false>(q, s));
            } else if (split) CRK_TRY(launch_resblock_fwd_tc<true>(q, s));
            else CRK_TRY(launch_resblock_fwd_tc<false>(q, s));
        }
    }
    const float hscale = sqrtf(1.0f / (float)c->layers);
    {
        const crk_conv_desc& d = L.tab.d[L.last1];
        ConvParams p = conv_params_default();
        p.X = act + A.skips; p.ldx = 64; p.Cin = 64; p.CinPad = 64;
        p.pro_act = c->head_act; p.pro_slope = c->slope; p.pro_scale = hscale;
        p.W = weff + d.w_off; p.bias = weff + d.bias_off;
        p.Y = act + A.head1; p.ldy = 64; p.Cout = 64; p.B = B; p.T = T;
        CRK_TRY(conv_dispatch(p, 2, weff + d.tc_off, d.tc_kpad, d.tc_n, s));
    }
    {
        const crk_conv_desc& d = L.tab.d[L.last2];
        ConvParams p = conv_params_default();
        p.X = act + A.head1; p.ldx = 64; p.Cin = 64; p.CinPad = 64;
        p.pro_act = c->head_act; p.pro_slope = c->slope;
        p.W = weff + d.w_off; p.bias = weff + d.bias_off;
        p.Y = y; p.ldy = ldy; p.Cout = c->out_ch; p.B = B; p.T = T;
        CRK_TRY(conv_dispatch(p, cpt_for(c->out_ch), weff + d.tc_off, d.tc_kpad, d.tc_n, s));
    }
    return CRK_OK;
}

inline int wavenet_bwd(const crk_wavenet_cfg* c, const float* theta, const float* weff, const float* x,
                       int ldx, const float* cond, int ldc, const float* dropmul, const float* act,
                       const float* dy, int lddy, float* dx, int lddx, float* dc, int lddc,
                       float* gtheta, float* ws, int B, int T, cudaStream_t s) {
    WavenetLayout L;
    int rc = wavenet_layout(c, &L);
    if (rc) return rc;
    if (!theta || !weff || !x || !act || !dy || !ws || B < 1 || T < 1) return CRK_ERR_ARG;
    const bool need_w = gtheta != nullptr;    // NULL: input gradients only (frozen parameters): every wgrad is skipped
    const long long F = (long long)B * T;
    const WavenetAct A = wavenet_act(c, F);
    const WavenetWs W = wavenet_ws(c, L, B, T);
    float* gweff = ws + W.gweff;
    float* part = ws + W.part;
    const float* h = act + A.h;
    const float* skips = act + A.skips;
    const float* head1 = act + A.head1;
    const float hscale = sqrtf(1.0f / (float)c->layers);
    const int cpt_out = cpt_for(c->out_ch);
    // side-stream mode: every wgrad (+ reduce) goes to `sw`, forked from `s` after its operands were produced
    SideStream* SS = (need_w && W.nb > 1) ? side_stream() : nullptr;
    const bool side = SS != nullptr;
    cudaStream_t sw = side ? SS->st : s;
    int evn = 0;
    if (side) CRK_TRY(stream_fork(SS, evn, s, sw));       // (dy and the saved activations are ready on `s`)

    // ---- head: y = W2.act(head1)+b2 ; head1 = W1.act(hscale*skips)+b1
    {
        const crk_conv_desc& d = L.tab.d[L.last2];
        WgradParams g;
        g.X = head1; g.ldx = 64; g.Cin = 64; g.Rows = 64;
        g.pro_act = c->head_act; g.pro_slope = c->slope; g.pro_scale = 1.f; g.xmul = nullptr; g.ldxmul = 0;
        g.G = dy; g.ldg = lddy; g.N = c->out_ch; g.B = B; g.T = T; g.k = 1; g.dil = 1; g.padl = 0;
        if (need_w) CRK_TRY(conv_wgrad(g, cpt_out, gweff + d.w_off, gweff + d.bias_off, part, sw));
        ConvParams p = conv_params_default();   // dhead1 = (dy . W2^T) * act'(head1)
        p.X = dy; p.ldx = lddy; p.Cin = c->out_ch; p.CinPad = d.wt_rows;
        p.W = weff + d.wt_off; p.Y = ws + W.dhead1; p.ldy = 64; p.Cout = 64; p.B = B; p.T = T;
        p.dact_src = head1; p.lddact = 64; p.dact_mode = c->head_act; p.dact_slope = c->slope;
        CRK_TRY(conv_dispatch(p, 2, weff + d.tct_off, d.tct_kpad, d.tct_n, s));
        if (side) CRK_TRY(stream_fork(SS, evn, s, sw));   // dhead1 -> wgrad(last1)
    }
    {
        const crk_conv_desc& d = L.tab.d[L.last1];
        WgradParams g;
        g.X = skips; g.ldx = 64; g.Cin = 64; g.Rows = 64;
        g.pro_act = c->head_act; g.pro_slope = c->slope; g.pro_scale = hscale; g.xmul = nullptr; g.ldxmul = 0;
        g.G = ws + W.dhead1; g.ldg = 64; g.N = 64; g.B = B; g.T = T; g.k = 1; g.dil = 1; g.padl = 0;
        if (need_w) CRK_TRY(conv_wgrad(g, 2, gweff + d.w_off, gweff + d.bias_off, part, sw));
        ConvParams p = conv_params_default();   // ds = (dhead1 . W1^T) * act'(skips) * hscale
        p.X = ws + W.dhead1; p.ldx = 64; p.Cin = 64; p.CinPad = 64;
        p.W = weff + d.wt_off; p.Y = ws + W.ds; p.ldy = 64; p.Cout = 64; p.B = B; p.T = T;
        p.dact_src = skips; p.lddact = 64; p.dact_mode = c->head_act; p.dact_slope = c->slope;
        p.out_scale = hscale;
        CRK_TRY(conv_dispatch(p, 2, weff + d.tct_off, d.tct_kpad, d.tct_n, s));
    }
    // ---- residual blocks, last to first
    float* dh_cur = nullptr;              // grad wrt output of layer l (null for the last layer)
    float* dh_bufs[2] = {ws + W.dhA, ws + W.dhB};
    int flip = 0;
    for (int l = c->layers - 1; l >= 0; --l) {
        const int dil = wn_dilation(c, l);
        const int padl = wn_padl(c, dil);
        const float* hin = h + (long long)l * F * 64;
        const float* dm = dropmul ? dropmul + (long long)l * F * 64 : nullptr;
        const long long lb = (long long)(l % W.nb);           // this layer's dg / gos / z buffers
        float* DGl = ws + W.dg + lb * F * 128;
        float* GOSl = ws + W.gos + lb * F * 128;
        float* Zl = ws + W.z + lb * F * 64;
        {
            ResBwdGateParams p;
            p.dH = dh_cur; p.dS = ws + W.ds; p.TaSb = act + A.tasb + (long long)l * F * 128;
            p.WosT = weff + L.tab.d[L.out[l]].wt_off;
            p.DG = DGl; p.GOS = GOSl; p.Z = Zl; p.B = B; p.T = T;
            cudaError_t ge = cudaSuccess;
            if (gate_bwd_tc(p, weff + L.tab.d[L.out[l]].tct_off, s, &ge)) CRK_TRY(ge);
            else CRK_TRY(launch_resblock_bwd_gate(p, s));
            if (side) CRK_TRY(stream_fork(SS, evn, s, sw));   // dg / gos / z of this layer -> its three wgrads
        }
        {   // [out|skip] weights:  dWos = z^T . gos
            const crk_conv_desc& d = L.tab.d[L.out[l]];
            WgradParams g;
            g.X = Zl; g.ldx = 64; g.Cin = 64; g.Rows = 64;
            g.pro_act = CRK_ACT_NONE; g.pro_slope = 0.f; g.pro_scale = 1.f; g.xmul = nullptr; g.ldxmul = 0;
            g.G = GOSl; g.ldg = 128; g.N = 128; g.B = B; g.T = T; g.k = 1; g.dil = 1; g.padl = 0;
            if (need_w) CRK_TRY(conv_wgrad(g, 4, gweff + d.w_off, gweff + d.bias_off, part, sw));
        }
        {   // dilated conv weights
            const crk_conv_desc& d = L.tab.d[L.conv[l]];
            WgradParams g;
            g.X = hin; g.ldx = 64; g.Cin = 64; g.Rows = 64;
            g.pro_act = CRK_ACT_NONE; g.pro_slope = 0.f; g.pro_scale = 1.f; g.xmul = dm; g.ldxmul = 64;
            g.G = DGl; g.ldg = 128; g.N = 128; g.B = B; g.T = T;
            g.k = c->kernel_size; g.dil = dil; g.padl = padl;
            if (need_w) CRK_TRY(conv_wgrad(g, 4, gweff + d.w_off, gweff + d.bias_off, part, sw));
        }
        if (c->aux_ch > 0) {
            const crk_conv_desc& d = L.tab.d[L.aux[l]];
            WgradParams g;
            g.X = cond; g.ldx = ldc; g.Cin = c->aux_ch; g.Rows = d.cin_pad;
            g.pro_act = CRK_ACT_NONE; g.pro_slope = 0.f; g.pro_scale = 1.f; g.xmul = nullptr; g.ldxmul = 0;
            g.G = DGl; g.ldg = 128; g.N = 128; g.B = B; g.T = T; g.k = 1; g.dil = 1; g.padl = 0;
            if (need_w) CRK_TRY(conv_wgrad(g, 4, gweff + d.w_off, nullptr, part, sw));
            if (dc) {
                ConvParams p = conv_params_default();
                p.X = DGl; p.ldx = 128; p.Cin = 128; p.CinPad = 128;
                p.W = weff + d.wt_off; p.Y = dc; p.ldy = lddc; p.Cout = c->aux_ch; p.B = B; p.T = T;
                p.accumulate = (l != c->layers - 1);
                CRK_TRY(conv_dispatch(p, cpt_for(c->aux_ch), weff + d.tct_off, d.tct_kpad, d.tct_n, s));
            }
        }
        {   // dgrad: dh_{l-1} = convT(dg) [* dropmul] + sqrt(.5)*dh_l ; (l==0: * first_act'(h0))
            const crk_conv_desc& d = L.tab.d[L.conv[l]];
            float* dst = dh_bufs[flip];
            ConvParams p = conv_params_default();
            p.X = DGl; p.ldx = 128; p.Cin = 128; p.CinPad = 128;
            p.W = weff + d.wt_off; p.Y = dst; p.ldy = 64; p.Cout = 64; p.B = B; p.T = T;
            p.k = c->kernel_size; p.dil = dil; p.padl = (c->kernel_size - 1) * dil - padl;
            p.mul_src = dm; p.ldmul = 64;
            if (dh_cur) { p.R = dh_cur; p.ldr = 64; p.rscale = CRK_SQRT_HALF; }
            if (l == 0 && c->first_act != CRK_ACT_NONE) {
                p.dact_src = h; p.lddact = 64; p.dact_mode = c->first_act; p.dact_slope = c->slope;
            }
            CRK_TRY(conv_dispatch(p, 2, weff + d.tct_off, d.tct_kpad, d.tct_n, s));
            dh_cur = dst;
            flip ^= 1;
        }
    }
    // ---- first 1x1
    {
        const crk_conv_desc& d = L.tab.d[L.first];
        WgradParams g;
        g.X = x; g.ldx = ldx; g.Cin = c->in_ch; g.Rows = d.cin_pad;
        g.pro_act = CRK_ACT_NONE; g.pro_slope = 0.f; g.pro_scale = 1.f; g.xmul = nullptr; g.ldxmul = 0;
        g.G = dh_cur; g.ldg = 64; g.N = 64; g.B = B; g.T = T; g.k = 1; g.dil = 1; g.padl = 0;
        if (side) CRK_TRY(stream_fork(SS, evn, s, sw));       // the input gradient of layer 0 -> wgrad(first)
        if (need_w) CRK_TRY(conv_wgrad(g, 2, gweff + d.w_off, gweff + d.bias_off, part, sw));
        if (dx) {
            ConvParams p = conv_params_default();
            p.X = dh_cur; p.ldx = 64; p.Cin = 64; p.CinPad = 64;
            p.W = weff + d.wt_off; p.Y = dx; p.ldy = lddx; p.Cout = c->in_ch; p.B = B; p.T = T;
            CRK_TRY(conv_dispatch(p, cpt_for(c->in_ch), weff + d.tct_off, d.tct_kpad, d.tct_n, s));
        }
    }
    if (side) CRK_TRY(stream_fork(SS, evn, sw, s));           // join: every weight gradient is complete
    if (need_w) CRK_TRY(launch_weightnorm_bwd(L.tab, theta, gweff, gtheta, s));
    return CRK_OK;
}

// ---- plain conv stack (ParallelWaveGANDiscriminator) -------------------------------------------
struct ConvstackLayout {
    DescTable tab;
    long long theta, weff;
    int cin[32], cout[32], dil[32];
};
inline int convstack_layout(const crk_convstack_cfg* c, ConvstackLayout* L) {
    if (!c || c->layers < 1 || c->layers > 32 || c->kernel_size < 1 || c->kernel_size > 9 ||
        (c->kernel_size % 2) == 0 || c->dilation_factor < 1)
        return CRK_ERR_ARG;
    if (c->in_ch < 1 || c->in_ch > CRK_MAX_IN_CH || c->out_ch < 1 || c->out_ch > 128 || c->conv_ch < 1 || c->conv_ch > 128)
        return CRK_ERR_ARG;
    LayoutBuilder b;
    for (int i = 0; i < c->layers; ++i) {
        const bool last = (i == c->layers - 1);
        int dil = 1, cin;
        if (last) {
            // upstream quirk kept: the last conv's input width is the loop's final conv_in_channels,
            // which is still in_ch unless the loop ran with i >= 1
            cin = (c->layers - 1 >= 2) ? c->conv_ch : c->in_ch;
            if (c->layers == 2 && c->in_ch != c->conv_ch) return CRK_ERR_UNSUPPORTED;
        } else if (i == 0) {
            cin = c->in_ch;
        } else {
            cin = c->conv_ch;
            if (c->dilation_factor == 1) dil = i;
            else for (int e = 0; e < i; ++e) dil *= c->dilation_factor;
        }
        const int cout = last ? c->out_ch : c->conv_ch;
        if ((c->kernel_size - 1) * dil > 64) return CRK_ERR_UNSUPPORTED;
        L->cin[i] = cin; L->cout[i] = cout; L->dil[i] = dil;
        b.add(cout, cin, c->kernel_size, true, 0, 32 * cpt_for(cout), round_up(cout, 4), wide_ld(cin));
    }
    L->tab = b.tab; L->theta = b.theta; L->weff = b.weff;
    return CRK_OK;
}

inline long long convstack_act_floats(const crk_convstack_cfg* c, long long F) {
    long long n = 0;
    ConvstackLayout L;
    if (convstack_layout(c, &L)) return -1;
    for (int i = 0; i + 1 < c->layers; ++i) n += F * L.cout[i];
    return n > 0 ? n : 4;
}
struct ConvstackWs { long long gweff, dA, dB, part, total; };
inline ConvstackWs convstack_ws(const crk_convstack_cfg* c, const ConvstackLayout& L, int B, int T) {
    const long long F = (long long)B * T;
    ConvstackWs w;
    w.gweff = 0;
    w.dA = round_up((int)L.weff, 4);
    w.dB = w.dA + F * 128;
    w.part = w.dB + F * 128;
    size_t m = 0;
    for (int i = 0; i < L.tab.n; ++i) {
        const crk_conv_desc& d = L.tab.d[i];
        const size_t q = wgrad_part_floats(B, T, d.k, d.cin_pad, d.ldw);
        m = q > m ? q : m;
    }
    w.total = w.part + (long long)m;
    return w;
}

inline int convstack_fwd(const crk_convstack_cfg* c, const float* weff, const float* x, int ldx, float* y,
                         int ldy, float* act, int B, int T, cudaStream_t s) {
    ConvstackLayout L;
    int rc = convstack_layout(c, &L);
    if (rc) return rc;
    if (!weff || !x || !y || !act || B < 1 || T < 1) return CRK_ERR_ARG;
    const long long F = (long long)B * T;
    const float* in = x; int ldin = ldx;
    long long off = 0;
    for (int i = 0; i < c->layers; ++i) {
        const crk_conv_desc& d = L.tab.d[i];
        const bool last = (i == c->layers - 1);
        ConvParams p = conv_params_default();
        p.X = in; p.ldx = ldin; p.Cin = L.cin[i]; p.CinPad = d.cin_pad;
        p.W = weff + d.w_off; p.bias = weff + d.bias_off;
        p.B = B; p.T = T; p.k = c->kernel_size; p.dil = L.dil[i]; p.padl = (c->kernel_size - 1) / 2 * L.dil[i];
        if (last) { p.Y = y; p.ldy = ldy; }
        else { p.Y = act + off; p.ldy = L.cout[i]; p.epi_act = CRK_ACT_LRELU; p.epi_slope = c->slope; }
        p.Cout = L.cout[i];
        CRK_TRY(conv_dispatch(p, cpt_for(L.cout[i]), weff + d.tc_off, d.tc_kpad, d.tc_n, s));
        in = p.Y; ldin = p.ldy;
        if (!last) off += F * L.cout[i];
    }
    return CRK_OK;
}

inline int convstack_bwd(const crk_convstack_cfg* c, const float* theta, const float* weff, const float* x,
                         int ldx, const float* act, const float* dy, int lddy, float* dx, int lddx,
                         float dx_scale, float* gtheta, float* ws, int B, int T, cudaStream_t s) {
    ConvstackLayout L;
    int rc = convstack_layout(c, &L);
    if (rc) return rc;
    if (!theta || !weff || !x || !act || !dy || !ws || B < 1 || T < 1) return CRK_ERR_ARG;
    const bool need_w = gtheta != nullptr;    // NULL: input gradients only (frozen parameters)
    const long long F = (long long)B * T;
    const ConvstackWs W = convstack_ws(c, L, B, T);
    float* gweff = ws + W.gweff;
    float* part = ws + W.part;
    long long offs[32];
    long long off = 0;
    for (int i = 0; i + 1 < c->layers; ++i) { offs[i] = off; off += F * L.cout[i]; }
    const float* dcur = dy; int lddcur = lddy;
    float* bufs[2] = {ws + W.dA, ws + W.dB};
    int flip = 0;
    for (int i = c->layers - 1; i >= 0; --i) {
        const crk_conv_desc& d = L.tab.d[i];
        const float* xin = (i == 0) ? x : act + offs[i - 1];
        const int ldxin = (i == 0) ? ldx : L.cout[i - 1];
        const int dil = L.dil[i], padl = (c->kernel_size - 1) / 2 * dil;
        WgradParams g;
        g.X = xin; g.ldx = ldxin; g.Cin = L.cin[i]; g.Rows = d.cin_pad;
        g.pro_act = CRK_ACT_NONE; g.pro_slope = 0.f; g.pro_scale = 1.f; g.xmul = nullptr; g.ldxmul = 0;
        g.G = dcur; g.ldg = lddcur; g.N = L.cout[i]; g.B = B; g.T = T; g.k = c->kernel_size; g.dil = dil; g.padl = padl;
        if (need_w) CRK_TRY(conv_wgrad(g, cpt_for(L.cout[i]), gweff + d.w_off, gweff + d.bias_off, part, s));
        if (i > 0 || dx) {
            ConvParams p = conv_params_default();
            p.X = dcur; p.ldx = lddcur; p.Cin = L.cout[i]; p.CinPad = d.wt_rows;
            p.W = weff + d.wt_off; p.B = B; p.T = T; p.k = c->kernel_size; p.dil = dil;
            p.padl = (c->kernel_size - 1) * dil - padl;
            p.Cout = L.cin[i];
            if (i > 0) {
                p.Y = bufs[flip]; p.ldy = L.cin[i];
                p.dact_src = xin; p.lddact = ldxin; p.dact_mode = CRK_ACT_LRELU; p.dact_slope = c->slope;
            } else {
                p.Y = dx; p.ldy = lddx; p.out_scale = dx_scale;
            }
            CRK_TRY(conv_dispatch(p, cpt_for(L.cin[i]), weff + d.tct_off, d.tct_kpad, d.tct_n, s));
            dcur = p.Y; lddcur = p.ldy;
            flip ^= 1;
        }
    }
    if (need_w) CRK_TRY(launch_weightnorm_bwd(L.tab, theta, gweff, gtheta, s));
    return CRK_OK;
}

}  // namespace crk
