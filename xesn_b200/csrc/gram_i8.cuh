// Error-free Gram accumulation on the 5th-gen tensor cores (tcgen05, kind::i8) -- device side.
//
// FP64 has no tcgen05 kind; DMMA tops out at 37 TFLOP/s.  But reservoir states are bounded,
// |r| <= 1 (tanh + convex leak), so they have an exact fixed-point image
//       q = rint(r * 2^55)  =  sum_{a=0..7} d_a * 128^(7-a),   d_a in [-64, 64]  (balanced base 128)
// and the Gram matrix  sum_t r_i r_j  splits into integer matrix products of the digit planes
//       sum_t r_i(t) r_j(t)  =  sum_{a,b} 2^(-12-7(a+b)) * sum_t d_a,i(t) d_b,j(t).
// Every plane product is an int8 x int8 -> int32 GEMM that the tensor cores compute EXACTLY
// (|d| <= 64, K <= 2^17 steps per chunk keeps the int32 sums below 2^31 even with 8 pairs per
// accumulator).  Pairs with a+b = s share one TMEM accumulator (same weight 2^(-12-7s)); pairs with
// a+b >= 8 are dropped (< 2^-68 per product).  The only roundings are the fixed-point image itself
// (2^-56 absolute, below FP64's own resolution for |r| > 2^-3) and the eight FP64 adds that
// combine the levels -- tighter than a DGEMM whose every FMA rounds.  36 int8 MMAs replace the
// FP64 FMAs: 36 * 2 N^2 K int-ops at 4.5 Pop/s vs 2 N^2 K flops at 37 TFLOP/s.
//
// Data layout: digit planes [8][rows][ldp] int8, row = time step, ldp >= N rounded up to 128:
// MN-major operands for BOTH sides of R^T R, read straight by tcgen05.mma (a_major = b_major = MN).
// The recurrence writes the digits of a state next to the FP64 value, so no transpose or
// re-quantisation pass exists.
//
// Tile: 128 (M) x 64 (N) outputs, K-block = 32 time steps.  TMEM: 8 levels x 64 columns = 512.
// Shared memory per stage: A 8 planes x 32 rows x 128 B (SWIZZLE_128B) + B 8 x 32 x 64 B
// (SWIZZLE_64B) = 48 KB, 4 stages.  All 256 threads copy with cp.async (16 B chunks straight into
// the swizzled position), one thread issues the 36 MMAs of a K-block and commits to an mbarrier.
#pragma once
#include "common.cuh"

namespace xesn {
namespace gi8 {

constexpr int TM = 128, TN = 64, KB = 32, NDIG = 8;
constexpr int STAGES = 4;
constexpr int A_PLANE = KB * 128;              // bytes per plane per stage (A)
constexpr int B_PLANE = KB * 64;               // (B)
constexpr int A_BYTES = NDIG * A_PLANE;        // 32 KB
constexpr int STAGE_BYTES = NDIG * (A_PLANE + B_PLANE);  // 48 KB
constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /*alignment slack*/ + 128 /*barriers + tmem ptr*/;
constexpr int THREADS = 256;
constexpr int ISSUERS = 4;  // threads issuing tcgen05.mma (one per warp 0..3), levels split modulo ISSUERS

struct GramI8Args {
    const int8_t* planes;   // [NDIG][rows][ldp]
    long long ldp;          // bytes between time steps
    long long plane_stride; // bytes between digit planes
    long long group_stride; // bytes between groups (batch)
    int K;                  // time steps, multiple of KB (tail rows zero)
    int N;                  // reservoir size (planes zero-padded up to ldp)
    double* G;              // [batch][N][ldg], lower triangle accumulated: G += R^T R
    long long ldg;
    long long strideG;
    int batch;
};

__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(unsigned bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned bar, unsigned parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t}\n" ::"r"(bar),
        "r"(parity)
        : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory"); }

__device__ __forceinline__ void cp16(unsigned dst, const void* src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_wait() {
    asm volatile("cp.async.wait_group %0;\n" ::"n"(N) : "memory");
}

// shared-memory matrix descriptor (cute::UMMA::SmemDescriptor): MN-major, swizzled, sm_100 version
__device__ __forceinline__ unsigned long long smem_desc(unsigned addr, unsigned sbo_bytes, unsigned layout_type) {
    unsigned long long d = 0;
    d |= (unsigned long long)((addr >> 4) & 0x3FFF);          // start address
    d |= (unsigned long long)(1) << 16;                         // leading byte offset (unused: one MN block)
    d |= (unsigned long long)((sbo_bytes >> 4) & 0x3FFF) << 32;  // stride byte offset: next group of 8 K rows
    d |= (unsigned long long)(1) << 46;                         // descriptor version (Blackwell)
    d |= (unsigned long long)(layout_type & 7) << 61;           // 2 = SWIZZLE_128B, 4 = SWIZZLE_64B
    return d;
}

// kind::i8 instruction descriptor (cute::UMMA::InstrDescriptor): S32 accumulate, signed 8-bit A and B,
// both MN-major, M = 128, N = 64
constexpr unsigned IDESC = (2u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) | ((TN >> 3) << 17) |
                           ((TM >> 4) << 24);

__device__ __forceinline__ void mma_i8(unsigned tmem_d, unsigned long long da, unsigned long long db, unsigned accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, {%5, %5, %5, %5}, p;\n\t}\n" ::"r"(tmem_d),
        "l"(da), "l"(db), "r"(IDESC), "r"(accumulate), "r"(0u)
        : "memory");
}
__device__ __forceinline__ void mma_commit(unsigned bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" ::"r"(bar) : "memory");
}

// 32 consecutive TMEM columns of this thread's lane
__device__ __forceinline__ void tmem_ld32(unsigned taddr, int (&v)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"
        "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];\n"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
          "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
          "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
          "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr)
        : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
}

// Per-CTA state of the Gram role (lives across tiles of a persistent loop).
struct GramI8Ctx {
    unsigned stage_base;   // shared address of stage 0 (1024-byte aligned)
    unsigned bar_done;     // STAGES mbarriers: MMAs that read the stage have completed
    unsigned bar_final;    // all MMAs of the tile have completed
    unsigned tmem;         // TMEM base address (512 columns)
    unsigned done_phase;   // bit s: parity to wait for on bar_done[s]
    unsigned final_phase;
    unsigned used;         // bit s: stage s has outstanding MMAs (bar_done must be waited before refill)
};

// one warp allocates all 512 TMEM columns; every thread learns the base address
__device__ __forceinline__ void gram_i8_setup(GramI8Ctx& ctx, unsigned char* smem_raw) {
    const unsigned raw = smem_u32(smem_raw);
    ctx.stage_base = (raw + 1023u) & ~1023u;
    const unsigned ctl = ctx.stage_base + STAGES * STAGE_BYTES;
    ctx.bar_done = ctl;            // 8 bytes each
    ctx.bar_final = ctl + 8 * STAGES;
    const unsigned tmem_slot = ctl + 8 * STAGES + 8;
    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; ++s) mbar_init(ctx.bar_done + 8 * s, ISSUERS);
        mbar_init(ctx.bar_final, ISSUERS);
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    }
    if (threadIdx.x < 32) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;\n" ::"r"(tmem_slot) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    asm volatile("ld.shared.b32 %0, [%1];\n" : "=r"(ctx.tmem) : "r"(tmem_slot) : "memory");
    ctx.done_phase = 0, ctx.final_phase = 0, ctx.used = 0;
}

__device__ __forceinline__ void gram_i8_teardown(GramI8Ctx& ctx) {
    tc_fence_before();
    __syncthreads();
    if (threadIdx.x < 32) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;\n" ::"r"(ctx.tmem) : "memory");
    }
}

// NC consecutive TMEM columns of this thread's lane (NC = 16 or 32)
template <int NC>
__device__ __forceinline__ void tmem_ld(unsigned taddr, int (&v)[NC]);
template <>
__device__ __forceinline__ void tmem_ld<32>(unsigned taddr, int (&v)[32]) { tmem_ld32(taddr, v); }
template <>
__device__ __forceinline__ void tmem_ld<16>(unsigned taddr, int (&v)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];\n"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
          "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
        : "r"(taddr)
        : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
}

// copy K-block kb of the A (rows m0.., 128 bytes) and B (rows n0.., 64 bytes) digit planes into stage s
// NT = threads of the CTA taking part (256 or 512)
template <int NT>
__device__ __forceinline__ void gram_i8_load(const GramI8Ctx& ctx, const GramI8Args& p, const int8_t* planes, int s,
                                             int kb, int m0, int n0) {
    const int tid = threadIdx.x;
    const unsigned sa = ctx.stage_base + s * STAGE_BYTES;
    const unsigned sb = sa + A_BYTES;
    const long long row0 = (long long)kb * KB;
    {   // A: 256 (row, chunk) slots per plane; thread -> slot tid & 255, planes a0, a0 + NT/256, ...
        constexpr int PSTEP = NT / 256;
        const int c = tid & 7, kk = (tid >> 3) & 31, a0 = tid >> 8;
        const int8_t* src = planes + (row0 + kk) * p.ldp + m0 + c * 16;
        const unsigned dst = sa + kk * 128 + ((c ^ (kk & 7)) << 4);
#pragma unroll
        for (int a = a0; a < NDIG; a += PSTEP) cp16(dst + a * A_PLANE, src + a * p.plane_stride);
    }
    {   // B: 128 slots per plane (4 chunks per row); thread -> slot tid & 127, planes a0, a0 + NT/128, ...
        constexpr int PSTEP = NT / 128;
        const int c = tid & 3, kk = (tid >> 2) & 31, a0 = tid >> 7;
        const int8_t* src = planes + (row0 + kk) * p.ldp + n0 + c * 16;
        const unsigned dst = sb + kk * 64 + ((c ^ ((kk >> 1) & 3)) << 4);
#pragma unroll
        for (int a = a0; a < NDIG; a += PSTEP) cp16(dst + a * B_PLANE, src + a * p.plane_stride);
    }
}

// One 128 x 64 tile: rows m0 = 128*ti, columns n0 = 64*tj of batch entry bz.
template <int NT>
__device__ __forceinline__ void gram_i8_tile(GramI8Ctx& ctx, const GramI8Args& p, int ti, int tj, int bz) {
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int m0 = ti * TM, n0 = tj * TN;
    const int8_t* planes = p.planes + (long long)bz * p.group_stride;
    const int nkb = p.K / KB;

    // prologue: fill up to STAGES-1 stages (stages are free: the previous tile waited for its last MMA)
#pragma unroll
    for (int s = 0; s < STAGES - 1; ++s) {
        if (s < nkb) gram_i8_load<NT>(ctx, p, planes, s, s, m0, n0);
        cp_commit();
    }
    for (int kb = 0; kb < nkb; ++kb) {
        const int s = kb % STAGES;
        cp_wait<STAGES - 2>();   // this thread's copies of K-block kb have landed
        fence_proxy_async();     // ... and are visible to the tensor core's (async proxy) reads
        __syncthreads();
        // Four issuing threads (lane 0 of warps 0..3), each owning the levels a+b = w (mod 4): a single
        // thread cannot issue one N=64 MMA (32 tensor-pipe cycles) every 32 cycles, which capped the
        // pipe at 46 % (profiles/gram_i8_r01_v3.txt).  All MMAs of a level come from one thread, so the
        // non-accumulating first MMA of a level still precedes the others in issue order.
        if (lane == 0 && warp < ISSUERS) {
            tc_fence_after();
            const unsigned sa = ctx.stage_base + s * STAGE_BYTES;
            const unsigned long long da0 = smem_desc(sa, 1024, 2);
            const unsigned long long db0 = smem_desc(sa + A_BYTES, 512, 4);
#pragma unroll
            for (int lv = 0; lv < NDIG; ++lv) {
                if ((lv % ISSUERS) != warp) continue;
#pragma unroll
                for (int a = 0; a <= lv; ++a)
                    mma_i8(ctx.tmem + lv * TN, da0 + (unsigned long long)(a * (A_PLANE >> 4)),
                           db0 + (unsigned long long)((lv - a) * (B_PLANE >> 4)), (kb > 0 || a > 0) ? 1u : 0u);
            }
            mma_commit(ctx.bar_done + 8 * s);
            if (kb == nkb - 1) mma_commit(ctx.bar_final);
        }
        ctx.used |= 1u << s;
        // refill the stage that K-block kb-1 used, once its MMAs are done
        const int nk = kb + STAGES - 1;
        if (nk < nkb) {
            const int sn = nk % STAGES;
            if (ctx.used & (1u << sn)) {
                mbar_wait(ctx.bar_done + 8 * sn, (ctx.done_phase >> sn) & 1u);
                ctx.done_phase ^= 1u << sn;
                ctx.used &= ~(1u << sn);
            }
            gram_i8_load<NT>(ctx, p, planes, sn, nk, m0, n0);
        }
        cp_commit();
    }
    cp_wait<0>();
    // drain: every stage's MMAs have completed once the final commit fires
    mbar_wait(ctx.bar_final, ctx.final_phase);
    ctx.final_phase ^= 1u;
    // keep the per-stage barrier phases in step: each outstanding commit has completed by now
#pragma unroll
    for (int s = 0; s < STAGES; ++s)
        if (ctx.used & (1u << s)) {
            mbar_wait(ctx.bar_done + 8 * s, (ctx.done_phase >> s) & 1u);
            ctx.done_phase ^= 1u << s;
        }
    ctx.used = 0;
    tc_fence_after();

    // epilogue: lane quarter q of TMEM <-> rows m0 + 32q + lane; the NT/128 warps of a quarter split the 64 columns
    constexpr int NC = TN / (NT / 128);
    const int q = warp & 3, part = warp >> 2;
    const int m = m0 + q * 32 + lane;
    double acc[NC];
#pragma unroll
    for (int j = 0; j < NC; ++j) acc[j] = 0.0;
    if (nkb > 0) {
#pragma unroll
        for (int s = NDIG - 1; s >= 0; --s) {  // smallest weights first
            int v[NC];
            tmem_ld<NC>(ctx.tmem + ((unsigned)(q * 32) << 16) + s * TN + part * NC, v);
            const double w = __longlong_as_double((long long)(1023 - 12 - 7 * s) << 52);  // 2^(-12-7s)
#pragma unroll
            for (int j = 0; j < NC; ++j) acc[j] = fma((double)v[j], w, acc[j]);
        }
    }
    if (m < p.N) {
        double* g = p.G + (long long)bz * p.strideG + (long long)m * p.ldg + n0 + part * NC;
        const int nmax = min(NC, min(p.N, m + 1) - (n0 + part * NC));  // columns <= row only (lower triangle)
        if (nmax == NC && ((reinterpret_cast<uintptr_t>(g) & 15) == 0)) {
#pragma unroll
            for (int j = 0; j < NC; j += 2) {
                double2 c = *reinterpret_cast<double2*>(g + j);
                c.x += acc[j], c.y += acc[j + 1];
                *reinterpret_cast<double2*>(g + j) = c;
            }
        } else {
#pragma unroll
            for (int j = 0; j < NC; ++j)
                if (j < nmax) g[j] += acc[j];
        }
    }
    // all TMEM reads of this tile are done before the next tile's first MMA overwrites the accumulators
    tc_fence_before();
    __syncthreads();
}

// ---- fixed-point image of a state: 8 balanced base-128 digits ----
// q = rint(r * 2^55);  q + BIAS has seven non-negative base-128 digits (d_a + 64) and a signed top
// digit, so every digit is a shift and a mask.  Returned packed: w03 = bytes (d0,d1,d2,d3) with d0
// (most significant digit) in the lowest byte, w47 = (d4,d5,d6,d7).
__device__ __forceinline__ void state_digit_words(double r, unsigned& w03, unsigned& w47) {
    // |r| <= 1; non-finite states (diverged runs) are mapped to 0 here, the FP64 ring keeps them
    long long q = (fabs(r) <= 1.0) ? __double2ll_rn(r * 36028797018963968.0 /* 2^55 */) : 0;
    constexpr long long BIAS = 64LL * (1LL + (1LL << 7) + (1LL << 14) + (1LL << 21) + (1LL << 28) + (1LL << 35) + (1LL << 42));
    q += BIAS;
    const unsigned lo = (unsigned)q, hi = (unsigned)((unsigned long long)q >> 32);
    const unsigned e7 = lo & 127u, e6 = (lo >> 7) & 127u, e5 = (lo >> 14) & 127u, e4 = (lo >> 21) & 127u;
    const unsigned e3 = __funnelshift_r(lo, hi, 28) & 127u, e2 = (hi >> 3) & 127u, e1 = (hi >> 10) & 127u;
    const unsigned e0 = (unsigned)(((int)hi) >> 17) + 64u;  // signed top digit, biased like the others
    w03 = __vsub4(e0 | (e1 << 8) | (e2 << 16) | (e3 << 24), 0x40404040u);
    w47 = __vsub4(e4 | (e5 << 8) | (e6 << 16) | (e7 << 24), 0x40404040u);
}

__device__ __forceinline__ void state_digits(double r, signed char (&d)[NDIG]) {
    unsigned w03, w47;
    state_digit_words(r, w03, w47);
#pragma unroll
    for (int a = 0; a < 4; ++a) {
        d[a] = (signed char)((w03 >> (8 * a)) & 255u);
        d[a + 4] = (signed char)((w47 >> (8 * a)) & 255u);
    }
}

}  // namespace gi8
}  // namespace xesn
