// crank-b200: Blackwell (sm_100a) tensor-core primitives -- tcgen05.mma (kind::tf32) with TMEM
// accumulators, mbarrier completion, and the shared-memory operand layout used by every
// tensor-core kernel of this library.
//
// Operand layout in shared memory ("chunk-major", UMMA canonical INTERLEAVE / no-swizzle):
//     elem(row r, channel c)  at  base + (c/4)*CS + r*16 + (c%4)*4     [bytes]
// i.e. [channel-chunk of 4 floats][row][4 floats], chunk stride CS = rows_alloc*16 with
// rows_alloc chosen ODD so that the per-warp staging stores (16 consecutive chunks of one row)
// hit 8 distinct bank groups (no conflicts).  One physical tile serves as
//   * a K-major operand   (rows = M or N index, channels = K):  SBO = 128 B, LBO = CS
//   * an MN-major operand (rows = K index (frames), channels = M or N): SBO = CS, LBO = 128 B
// and -- the reason for choosing it over a 128B-swizzled TMA tile -- a descriptor may start at
// ANY row (start address += rows*16 B), so the k dilated taps of a conv read ONE staged tile with a
// halo instead of k shifted copies.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace crk {
// optional per-CTA phase timestamps (clock64) for performance debugging: crk_debug_timestamps(ptr)
__device__ long long* g_crk_dbg = nullptr;
__device__ __forceinline__ void dbg_stamp(int enabled, int slot) {
    if (enabled && g_crk_dbg && threadIdx.x == 64) g_crk_dbg[(size_t)blockIdx.x * 16 + slot] = clock64();
}
// single-thread diagnostic accumulators (slots 8..15 of the CTA's stamp row): e.g. cycles an MMA issuer
// spent waiting for weight bytes, or a TMA producer for a free ring slot
__device__ __forceinline__ void dbg_put(int enabled, int slot, long long v) {
    if (enabled && g_crk_dbg) g_crk_dbg[(size_t)blockIdx.x * 16 + slot] = v;
}
// host side: which launch of which kernel family records stamps
struct DbgSel { int kind = 0; int target = 0; int count = 0; };
inline DbgSel& dbg_sel() { static DbgSel d; return d; }
inline int dbg_take(int kind) {
    DbgSel& d = dbg_sel();
    if (d.kind != kind) return 0;
    return (d.count++ == d.target) ? 1 : 0;
}
namespace tc {

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

// ---- mbarrier ----------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void fence_mbar_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ uint32_t mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t"
        "}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok;
}
// bounded spin (~seconds): a mis-programmed pipeline must fail a test, not hang the GPU box
__device__ __forceinline__ bool mbar_wait(uint64_t* bar, uint32_t parity) {
    for (uint32_t it = 0; it < (1u << 24); ++it)
        if (mbar_try_wait(bar, parity)) return true;
    return false;
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// one arrival + expected transaction bytes (the bulk copies below complete the transaction count)
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
// TMA 1-D bulk copy global -> shared, completion signalled on an mbarrier (one thread issues; the copy
// engine streams the bytes: no register staging, no per-thread load latency).  16 B aligned, size % 16 == 0.
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst_smem)),
                 "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
// TMA 1-D bulk copy shared -> global (one thread issues; 16 B aligned, size % 16 == 0), then wait until the whole group has
// completed.  Round 2 experiment (opt-in, crk_debug_opt_enable 4): the wgrad kernels' partial-sum epilogue stores 128 B per
// warp instruction from the thread-per-row tensor-memory layout at ~5.4 B/clk per SM (30 K of 97 K cycles for k = 5); staging
// the block in shared memory and handing it to the copy engine was measured SLOWER (8.0 vs 7.75 ms per step).
__device__ __forceinline__ void bulk_s2g(void* dst_gmem, const void* src_smem, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst_gmem), "r"(smem_u32(src_smem)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void bulk_commit_wait_all() {
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}
// generic-proxy smem writes (st.shared) -> visible to the async proxy (tensor core operand reads)
__device__ __forceinline__ void fence_proxy_async_smem() {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

// ---- TMEM --------------------------------------------------------------------------------------
// one full warp; writes the allocated base address (lane<<16 | column) to *dst (shared memory)
template <int NCOLS>
__device__ __forceinline__ void tmem_alloc(uint32_t* dst) {
    static_assert(NCOLS == 32 || NCOLS == 64 || NCOLS == 128 || NCOLS == 256 || NCOLS == 512, "pow2 >= 32");
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst)), "n"(NCOLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
template <int NCOLS>
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "n"(NCOLS) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// 32 lanes x 32 consecutive fp32 columns: thread i of the warp gets lane (warp%4)*32+i
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float (&v)[32]) {
    uint32_t r[32];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
          "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
          "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr)
        : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}

// ---- descriptors -------------------------------------------------------------------------------
// shared-memory matrix descriptor, no swizzle (layout_type 0), descriptor version 1 (Blackwell)
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;
    return d;
}
// instruction descriptor: kind::tf32, fp32 accumulate; a_mn / b_mn = 1 for MN-major operands
__host__ __device__ constexpr uint32_t make_idesc_tf32(int M, int N, int a_mn, int b_mn) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)a_mn << 15) | ((uint32_t)b_mn << 16) |
           ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
// D[tmem] (+)= A[smem] . B[smem]^T   (one thread issues)
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                          uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
        "}" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// D[tmem] (+)= A[tmem] . B[smem]^T : A operand read from tensor memory (lane = row m, one 32-bit column per
// k element, 8 columns per MMA) instead of shared memory.  Measured (profiles/mma_rate.py): an SS-form
// 128 x N x 8 TF32 MMA costs ~66 cycles for every N <= 128 -- the 4 KB A fetch from shared memory is the
// floor -- so for N = 64 tiles half the tensor pipe idles; the TS form removes that fetch.
__device__ __forceinline__ void umma_tf32_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc,
                                             uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t"
        "}" ::"r"(tmem_d), "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// registers -> TMEM, 32 lanes x 32 consecutive columns (thread i of the warp writes lane (warp%4)*32+i)
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const float (&v)[32]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
        "{%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,%32};"
        ::"r"(taddr), "r"(__float_as_uint(v[0])), "r"(__float_as_uint(v[1])), "r"(__float_as_uint(v[2])), "r"(__float_as_uint(v[3])),
          "r"(__float_as_uint(v[4])), "r"(__float_as_uint(v[5])), "r"(__float_as_uint(v[6])), "r"(__float_as_uint(v[7])),
          "r"(__float_as_uint(v[8])), "r"(__float_as_uint(v[9])), "r"(__float_as_uint(v[10])), "r"(__float_as_uint(v[11])),
          "r"(__float_as_uint(v[12])), "r"(__float_as_uint(v[13])), "r"(__float_as_uint(v[14])), "r"(__float_as_uint(v[15])),
          "r"(__float_as_uint(v[16])), "r"(__float_as_uint(v[17])), "r"(__float_as_uint(v[18])), "r"(__float_as_uint(v[19])),
          "r"(__float_as_uint(v[20])), "r"(__float_as_uint(v[21])), "r"(__float_as_uint(v[22])), "r"(__float_as_uint(v[23])),
          "r"(__float_as_uint(v[24])), "r"(__float_as_uint(v[25])), "r"(__float_as_uint(v[26])), "r"(__float_as_uint(v[27])),
          "r"(__float_as_uint(v[28])), "r"(__float_as_uint(v[29])), "r"(__float_as_uint(v[30])), "r"(__float_as_uint(v[31]))
        : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// one deterministic leader lane of a fully active warp (the same lane for every call)
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile(
        "{\n\t"
        ".reg .pred P;\n\t"
        "elect.sync _|P, 0xffffffff;\n\t"
        "selp.u32 %0, 1, 0, P;\n\t"
        "}"
        : "=r"(pred));
    return pred != 0;
}

// arrive on an mbarrier when all previously issued tcgen05.mma of this thread have completed
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// ---- operand staging ---------------------------------------------------------------------------
// chunk stride (bytes) for a tile of `rows` rows: rows rounded up to an odd count
__host__ __device__ constexpr int chunk_rows(int rows) { return (rows & 1) ? rows : rows + 1; }
__host__ __device__ constexpr int chunk_stride_bytes(int rows) { return chunk_rows(rows) * 16; }

// split x into two tf32-representable parts (3xTF32 error compensation), BOTH rounded to nearest:
//   hi = rn_tf32(x)            |x - hi| <= 2^-12 |x|, exact difference in fp32
//   lo = rn_tf32(x - hi)       |x - hi - lo| <= 2^-12 |x - hi| <= 2^-24 |x|
// so hi + lo carries x to fp32's own precision and the dropped lo*lo term of a product is <= 2^-24 of it.
// Round 1 truncated instead (hi = x & 0xFFFFE000, lo = x - hi left for the tensor core to truncate to 11 bits):
// a one-sided error of up to 2^-21 |x| per operand, 8x larger -- measured as 5e-6 forward error per stack, enough to
// flip ReLU derivatives of near-zero pre-activations in backward ~10x more often than the fp32 kernels do
// (profiles/diag_r2c.py).  The tensor core truncates its fp32 operands to tf32 (profiles/hwprobe T4), which is exact
// for values that are already tf32-representable.
// (cvt.rna.tf32.f32 compiles to FSETP |x| < inf + predicated IADD + LOP3; the same rounding -- nearest, ties away from
// zero, on the magnitude bits -- is an add of half a tf32 ulp and a mask: inf stays inf, NaN stays NaN, and a value that
// rounds past the largest finite tf32 becomes inf exactly as cvt does.  One instruction less on every operand element.)
__device__ __forceinline__ float rn_tf32(float x) {
    return __uint_as_float((__float_as_uint(x) + 0x1000u) & 0xffffe000u);
}
__device__ __forceinline__ void split_tf32(float x, float& hi, float& lo) {
    hi = rn_tf32(x);
    lo = rn_tf32(x - hi);
}

}  // namespace tc
}  // namespace crk
