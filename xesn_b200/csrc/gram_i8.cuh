// Gram accumulation on the 5th-gen tensor cores (tcgen05, kind::i8) from fixed-point digit planes -- device side.
//
// FP64 has no tcgen05 kind; DMMA tops out at 37 TFLOP/s.  But reservoir states are bounded,
// |r| <= 1 (tanh + convex leak), so they have a fixed-point image
//       q = rint(r * 2^54)  =  sum_{a=0..6} d_a * 256^(6-a),   d_a in [-128, 127]  (balanced base 256)
// and the Gram matrix  sum_t r_i r_j  splits into integer matrix products of the digit planes
//       sum_t r_i(t) r_j(t)  =  sum_{a,b} 2^(-12-8(a+b)) * sum_t d_a,i(t) d_b,j(t).
// Every plane product is an int8 x int8 -> int32 GEMM that the tensor cores compute EXACTLY
// (|d| <= 128, at most 7 pairs per accumulator: K <= 16384 steps per launch keeps the int32 sums below
// 2^31).  Pairs with a+b = s share one TMEM accumulator (same weight 2^(-12-8s)); the 28 pairs with
// a+b <= 6 are kept.  Roundings: the fixed-point image (2^-55 absolute), the 21 dropped pairs (a+b >= 7:
// below 2^-51 per product in the worst case, ~2^-54 typical, zero-mean) and the seven FP64 adds that combine
// the levels.  GUARANTEE: an ABSOLUTE error <= K 2^-50 per Gram entry over K steps, whatever the magnitude of
// the states -- the accuracy of an FP64 DGEMM for |r| ~ 1, a relative accuracy that degrades for states near
// 0 (the ridge system's beta on the diagonal is an absolute scale too); exact when the states live on a
// 2^-42 grid.  Non-finite states are digitised as 0 (the FP64 ring keeps them, so ybar and Wout become NaN).
// 28 int8 MMAs replace the FP64 FMAs: 28 * 2 N^2 K int-ops at 4.5 Pop/s vs 2 N^2 K flops at 37 TFLOP/s.
// (First version: 8 balanced base-128 digits, 36 pairs, 8 levels = all of TMEM; base 256 needs 22 % fewer
// MMAs and one plane less for the same 2^-55..2^-56 resolution.)
//
// Data layout: digit planes [7][rows][ldp] int8, row = time step, ldp >= N rounded up to 128:
// MN-major operands for BOTH sides of R^T R, read straight by tcgen05.mma (a_major = b_major = MN).
// The recurrence writes the digits of a state next to the FP64 value, so no transpose or
// re-quantisation pass exists.
//
// Tile: 128 (M) x 64 (N) outputs, K-block = 32 time steps.  TMEM: 7 levels x 64 columns = 448 of 512.
// Shared memory per stage: A 7 planes x 32 rows x 128 B (SWIZZLE_128B) + B 7 x 32 x 64 B
// (SWIZZLE_64B) = 42 KB, 5 stages.  All threads copy with cp.async (16 B chunks straight into
// the swizzled position), four threads issue the 28 MMAs of a K-block and commit to an mbarrier.
#pragma once
#include "common.cuh"

namespace xesn {
namespace gi8 {

constexpr int TM = 128, TN = 64, KB = 32, NDIG = 7;
constexpr int NDIG_WIDE = 8;  // 8 digits (2^-63 of the row scale): the trailing updates of the ridge solve (cholesky.cu)
constexpr int MAX_K = 16384;  // int32 accumulators: 8 pairs x 2^14 x K < 2^31
constexpr int A_PLANE = KB * 128;              // bytes per plane per stage (A)
constexpr int B_PLANE = KB * 64;               // (B)
// shared-memory ring of an ND-digit pipeline: 7 digits -> 42 KB stages x 5, 8 digits -> 48 KB x 4
template <int ND>
struct Ring {
    static constexpr int A_BYTES = ND * A_PLANE;
    static constexpr int STAGE_BYTES = ND * (A_PLANE + B_PLANE);
    static constexpr int STAGES = ND <= 7 ? 5 : 4;
    static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /*alignment slack*/ + 256 /*barriers + tmem ptr*/;
};
constexpr int SMEM_BYTES = Ring<NDIG_WIDE>::SMEM_BYTES > Ring<NDIG>::SMEM_BYTES ? Ring<NDIG_WIDE>::SMEM_BYTES
                                                                                : Ring<NDIG>::SMEM_BYTES;
constexpr int THREADS = 512;
constexpr int ISSUERS = 1;  // one thread issues the 10 (wide) MMAs of a K-block

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
    int ncols;              // columns of G that exist (0 = N): rows may extend beyond them (solve: appended rhs rows)
    int ndig;               // digit planes per group (NDIG or NDIG_WIDE); 0 = NDIG
    const double* scale;    // nullptr: G += P^T P.  Else the planes hold rows scaled by 1/scale[i] and the
                            // tile update is  G[m][n] -= scale[m] * scale[n] * (P^T P)[m][n]   (SYRK of the solve)
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
// canonical MN-major layout (16-byte units): ((atom,n),(8,k)) : ((1,LBO),(atom,SBO)) -- LBO = distance between
// swizzle atoms along MN (here: between the stacked digit planes of B), SBO = between groups of 8 K rows
__device__ __forceinline__ unsigned long long smem_desc(unsigned addr, unsigned sbo_bytes, unsigned layout_type,
                                                        unsigned lbo_bytes = 16) {
    unsigned long long d = 0;
    d |= (unsigned long long)((addr >> 4) & 0x3FFF);          // start address
    d |= (unsigned long long)((lbo_bytes >> 4) & 0x3FFF) << 16;  // leading byte offset
    d |= (unsigned long long)((sbo_bytes >> 4) & 0x3FFF) << 32;  // stride byte offset: next group of 8 K rows
    d |= (unsigned long long)(1) << 46;                         // descriptor version (Blackwell)
    d |= (unsigned long long)(layout_type & 7) << 61;           // 2 = SWIZZLE_128B, 4 = SWIZZLE_64B
    return d;
}

// kind::i8 instruction descriptor (cute::UMMA::InstrDescriptor): S32 accumulate, signed 8-bit A and B,
// both MN-major, M = 128, N = n (multiple of 16, <= 256)
__host__ __device__ constexpr unsigned idesc_n(int n) {
    return (2u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) | ((unsigned)(n >> 3) << 17) | ((TM >> 4) << 24);
}

__device__ __forceinline__ void mma_i8(unsigned tmem_d, unsigned long long da, unsigned long long db, unsigned idesc,
                                       unsigned accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, {%5, %5, %5, %5}, p;\n\t}\n" ::"r"(tmem_d),
        "l"(da), "l"(db), "r"(idesc), "r"(accumulate), "r"(0u)
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
    unsigned bar_full;     // STAGES mbarriers: the TMA boxes of the stage have landed
    unsigned bar_empty;    // STAGES mbarriers: the MMAs that read the stage have completed
    unsigned bar_tfull;    // all MMAs of the tile have completed: accumulators ready for the epilogue
    unsigned bar_tempty;   // the epilogue has read the accumulators
    unsigned tmem;         // TMEM base address (512 columns)
    // dynamic tile scheduling (gram_i8_run<.., DYN = true>): the producer draws tile numbers from a global counter and
    // hands them to the MMA issuer and the epilogue through a ring of TILE_RING slots, one mbarrier per slot
    unsigned bar_tile;     // TILE_RING mbarriers: the slot holds the number of the next tile (or -1: no more tiles)
    unsigned tile_ids;     // TILE_RING ints
};
constexpr int TILE_RING = 8;   // > STAGES + 2: the producer is never that many tiles ahead of the epilogue

// one warp allocates all 512 TMEM columns; every thread learns the base address
template <int ND = NDIG>
__device__ __forceinline__ void gram_i8_setup(GramI8Ctx& ctx, unsigned char* smem_raw) {
    constexpr int STAGES = Ring<ND>::STAGES, STAGE_BYTES = Ring<ND>::STAGE_BYTES;
    const unsigned raw = smem_u32(smem_raw);
    ctx.stage_base = (raw + 1023u) & ~1023u;
    const unsigned ctl = ctx.stage_base + STAGES * STAGE_BYTES;
    ctx.bar_full = ctl;            // 8 bytes each
    ctx.bar_empty = ctl + 8 * STAGES;
    ctx.bar_tfull = ctl + 16 * STAGES;
    ctx.bar_tempty = ctl + 16 * STAGES + 8;
    const unsigned tmem_slot = ctl + 16 * STAGES + 16;
    ctx.bar_tile = ctl + 16 * STAGES + 24;
    ctx.tile_ids = ctx.bar_tile + 8 * TILE_RING;    // ends at ctl + 16 STAGES + 24 + 96 <= ctl + 256
    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(ctx.bar_full + 8 * s, 1);
            mbar_init(ctx.bar_empty + 8 * s, ISSUERS);
        }
        for (int s = 0; s < TILE_RING; ++s) mbar_init(ctx.bar_tile + 8 * s, 1);
        mbar_init(ctx.bar_tfull, ISSUERS);
        mbar_init(ctx.bar_tempty, 8 /* EPI_WARPS */);
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

// ---- TMA + mbarrier helpers -----------------------------------------------------------------
__device__ __forceinline__ void mbar_expect_tx(unsigned bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(unsigned bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"(bar) : "memory");
}
// one box of the digit-plane tensor (bytes along N, time steps, planes) -> shared memory, swizzled by the TMA unit
__device__ __forceinline__ void tma_load_3d(unsigned dst, const void* tmap, int c0, int c1, int c2, unsigned bar) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];\n" ::"r"(dst),
        "l"(tmap), "r"(c0), "r"(c1), "r"(c2), "r"(bar)
        : "memory");
}

// Persistent Gram role of one CTA (512 threads), warp-specialised:
//   warp 4 lane 0      producer: one TMA box per operand and K-block (A: 7 planes x 32 steps x 128 B, SWIZZLE_128B;
//                      B: 7 x 32 x 64 B, SWIZZLE_64B) into a STAGES-deep ring, completion on full[s]
//   warp 0 lane 0      MMA issuer: wait full[s], issue the 10 MMAs of the K-block, tcgen05.commit -> empty[s];
//                      after the last K-block of a tile commit -> tmem_full
//   warps 8..15        epilogue: wait tmem_full, tcgen05.ld the 7 levels, combine in FP64, release TMEM
//                      (tmem_empty) and only then read-modify-write G, so the next tile's MMAs overlap the
//                      global-memory part of the epilogue
// The first version refilled the ring with cp.async from every thread and ran bulk-synchronously (wait copies,
// __syncthreads, issue, wait MMAs, refill): ncu showed half of all stall samples at that barrier and the
// tensor pipe idle 44 % of the time, because the issuing threads block while the MMA queue is full and
// nothing else overlaps with it.  Here loads, issue and epilogue only meet at mbarriers.
constexpr int EPI_WARP0 = 8, EPI_WARPS = 8, PRODUCER_WARP = 4;

// tile t of the lower triangle: from the host-built list, or (tiles == nullptr) row by row, tile row ti holding
// the column tiles 0 .. 2 ti + 1
__device__ __forceinline__ int2 tile_of(const int2* __restrict__ tiles, long long t) {
    if (tiles) return tiles[t];
    int ti = (int)((sqrt(4.0 * (double)t + 1.0) - 1.0) * 0.5);
    while ((long long)ti * (ti + 1) > t) --ti;
    while ((long long)(ti + 1) * (ti + 2) <= t) ++ti;
    return make_int2(ti, (int)(t - (long long)ti * (ti + 1)));
}

// tile sequence of one role: static (tile worker, worker + nworkers, ...) or drawn by the producer from *counter
template <bool DYN>
struct TileSeq {
    long long t, step, total;
    unsigned n = 0;
    __device__ __forceinline__ TileSeq(long long worker, long long nworkers, long long total_) : t(worker), step(nworkers), total(total_) {}
    // consumer side (issuer, epilogue): number of the next tile, or -1
    __device__ __forceinline__ long long next(const GramI8Ctx& ctx) {
        if (!DYN) {
            const long long r = t < total ? t : -1;
            t += step;
            return r;
        }
        const unsigned slot = n % TILE_RING;
        mbar_wait(ctx.bar_tile + 8 * slot, (n / TILE_RING) & 1u);
        int id;
        asm volatile("ld.shared.b32 %0, [%1];\n" : "=r"(id) : "r"(ctx.tile_ids + 4 * slot) : "memory");
        ++n;
        return id;
    }
    // producer side: draws the next tile and publishes it to the other roles
    __device__ __forceinline__ long long draw(const GramI8Ctx& ctx, int* counter) {
        if (!DYN) return next(ctx);
        const long long got = atomicAdd(counter, 1);
        const int id = got < total ? (int)got : -1;
        const unsigned slot = n % TILE_RING;
        asm volatile("st.shared.b32 [%0], %1;\n" ::"r"(ctx.tile_ids + 4 * slot), "r"(id) : "memory");
        mbar_arrive(ctx.bar_tile + 8 * slot);
        ++n;
        return id;
    }
};

template <int NT, int ND = NDIG, bool DYN = false>
__device__ __forceinline__ void gram_i8_run(GramI8Ctx& ctx, const GramI8Args& p, const void* tmA, const void* tmB,
                                            const int2* __restrict__ tiles, int ntiles, long long worker,
                                            long long nworkers, int* counter = nullptr) {
    static_assert(NT == 512, "the Gram role is laid out for 16 warps");
    constexpr int NDIG = ND;  // shadows the 7-digit default inside this function
    constexpr int STAGES = Ring<ND>::STAGES, STAGE_BYTES = Ring<ND>::STAGE_BYTES, A_BYTES = Ring<ND>::A_BYTES;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int nkb = p.K / KB;
    const long long total = (long long)ntiles * p.batch;

    if (warp == PRODUCER_WARP) {
        if (lane == 0) {
            unsigned it = 0;
            TileSeq<DYN> seq(worker, nworkers, total);
            for (long long t = seq.draw(ctx, counter); t >= 0; t = seq.draw(ctx, counter)) {
                const int2 ij = tile_of(tiles, t % ntiles);
                const int bz = (int)(t / ntiles);
                const int m0 = ij.x * TM, n0 = ij.y * TN;
                for (int kb = 0; kb < nkb; ++kb, ++it) {
                    const unsigned s = it % STAGES, use = it / STAGES;
                    mbar_wait(ctx.bar_empty + 8 * s, (use & 1u) ^ 1u);
                    const unsigned sa = ctx.stage_base + s * STAGE_BYTES;
                    mbar_expect_tx(ctx.bar_full + 8 * s, STAGE_BYTES);
                    tma_load_3d(sa, tmA, m0, kb * KB, bz * NDIG, ctx.bar_full + 8 * s);
                    tma_load_3d(sa + A_BYTES, tmB, n0, kb * KB, bz * NDIG, ctx.bar_full + 8 * s);
                }
            }
        }
        __syncwarp();
    } else if (warp < ISSUERS) {
        if (lane == 0) {
            unsigned it = 0, tc = 0;
            TileSeq<DYN> seq(worker, nworkers, total);
            for (long long t = seq.next(ctx); t >= 0; t = seq.next(ctx), ++tc) {
                mbar_wait(ctx.bar_tempty, (tc & 1u) ^ 1u);  // the epilogue has drained the previous tile's accumulators
                tc_fence_after();
                for (int kb = 0; kb < nkb; ++kb, ++it) {
                    const unsigned s = it % STAGES, use = it / STAGES;
                    mbar_wait(ctx.bar_full + 8 * s, use & 1u);
                    tc_fence_after();
                    const unsigned sa = ctx.stage_base + s * STAGE_BYTES;
                    const unsigned long long da0 = smem_desc(sa, 1024, 2);
                    const unsigned long long db0 = smem_desc(sa + A_BYTES, 512, 4, B_PLANE);
                    // A plane a meets B planes 0..6-a, whose products belong to the CONSECUTIVE levels a..6:
                    // the B planes are stacked along N (descriptor LBO = plane pitch), so one MMA of width
                    // 64*(7-a) (split at the 256-column instruction limit) writes all of them -- 10 MMAs
                    // instead of 28, and the A plane is read from shared memory once or twice instead of 7-a times
#pragma unroll
                    for (int a = 0; a < NDIG; ++a) {
#pragma unroll
                        for (int b0 = 0; b0 < NDIG - a; b0 += 4) {
                            const int nb = (NDIG - a - b0) < 4 ? (NDIG - a - b0) : 4;  // B planes b0 .. b0+nb-1
                            mma_i8(ctx.tmem + (a + b0) * TN, da0 + (unsigned long long)(a * (A_PLANE >> 4)),
                                   db0 + (unsigned long long)(b0 * (B_PLANE >> 4)), idesc_n(nb * TN),
                                   (kb > 0 || a > 0) ? 1u : 0u);
                        }
                    }
                    mma_commit(ctx.bar_empty + 8 * s);  // stage s may be refilled once these MMAs have read it
                }
                mma_commit(ctx.bar_tfull);
            }
        }
        __syncwarp();
    } else if (warp >= EPI_WARP0 && warp < EPI_WARP0 + EPI_WARPS) {
        // lane quarter q of TMEM <-> rows m0 + 32q + lane; the two warps of a quarter split the 64 columns
        constexpr int NC = TN / (EPI_WARPS / 4);
        const int q = warp & 3, part = (warp - EPI_WARP0) >> 2;
        unsigned tc = 0;
        TileSeq<DYN> seq(worker, nworkers, total);
        for (long long t = seq.next(ctx); t >= 0; t = seq.next(ctx), ++tc) {
            const int2 ij = tile_of(tiles, t % ntiles);
            const int bz = (int)(t / ntiles);
            const int m0 = ij.x * TM, n0 = ij.y * TN;
            mbar_wait(ctx.bar_tfull, tc & 1u);
            tc_fence_after();
            double acc[NC];
#pragma unroll
            for (int j = 0; j < NC; ++j) acc[j] = 0.0;
#pragma unroll
            for (int s = NDIG - 1; s >= 0; --s) {  // smallest weights first
                int v[NC];
                tmem_ld<NC>(ctx.tmem + ((unsigned)(q * 32) << 16) + s * TN + part * NC, v);
                const double w = __longlong_as_double((long long)(1023 - 12 - 8 * s) << 52);  // 2^(-12-8s)
#pragma unroll
                for (int j = 0; j < NC; ++j) acc[j] = fma((double)v[j], w, acc[j]);
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(ctx.bar_tempty);  // accumulators may be overwritten by the next tile
            const int m = m0 + q * 32 + lane;
            if (p.scale) {
                const double sm = (m < p.N) ? -p.scale[m] : 0.0;
#pragma unroll
                for (int j = 0; j < NC; ++j) acc[j] *= sm * __ldg(p.scale + n0 + part * NC + j);  // powers of two: exact
            }
            if (m < p.N) {
                double* g = p.G + (long long)bz * p.strideG + (long long)m * p.ldg + n0 + part * NC;
                const int ncols = p.ncols ? p.ncols : p.N;
                const int nmax = min(NC, min(ncols, m + 1) - (n0 + part * NC));  // columns <= row only (lower triangle)
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
        }
    }
}

// ---- fixed-point image of a state: 7 balanced base-256 digits ----
// q = rint(r * 2^54);  q + BIAS has six non-negative base-256 digits (d_a + 128) and a signed top
// digit (|q| <= 2^54 keeps it within +-65), so every digit is a byte of q + BIAS with its top bit flipped.
// Returned packed: w03 = bytes (d0,d1,d2,d3) with d0 (most significant digit) in the lowest byte,
// w47 = (d4,d5,d6,0).
__device__ __forceinline__ void state_digit_words(double r, unsigned& w03, unsigned& w47) {
    // |r| <= 1; non-finite states (diverged runs) are mapped to 0 here, the FP64 ring keeps them
    long long q = (fabs(r) <= 1.0) ? __double2ll_rn(r * 18014398509481984.0 /* 2^54 */) : 0;
    constexpr long long BIAS = 128LL * (1LL + (1LL << 8) + (1LL << 16) + (1LL << 24) + (1LL << 32) + (1LL << 40));
    q += BIAS;
    const unsigned lo = (unsigned)q, hi = (unsigned)((unsigned long long)q >> 32);
    // bytes of q + BIAS, least significant first: lo = (b0,b1,b2,b3), hi = (b4,b5,top,sign); digit d_a = byte 6-a
    const unsigned x03 = __byte_perm(lo, hi, 0x3456);  // (top, b5, b4, b3)
    const unsigned x47 = __byte_perm(lo, hi, 0x7012);  // (b2, b1, b0, sign byte -> cleared below)
    w03 = x03 ^ 0x80808000u;                           // flip the bias bit of b5, b4, b3; the top digit is signed as is
    w47 = (x47 & 0x00FFFFFFu) ^ 0x00808080u;
}

__device__ __forceinline__ void state_digits(double r, signed char (&d)[NDIG]) {
    unsigned w03, w47;
    state_digit_words(r, w03, w47);
#pragma unroll
    for (int a = 0; a < NDIG; ++a) d[a] = (signed char)(((a < 4 ? w03 : w47) >> (8 * (a & 3))) & 255u);
}

}  // namespace gi8
}  // namespace xesn
