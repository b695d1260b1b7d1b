// Device-side body of the persistent teacher-forced recurrence (see recurrence.cu for the design
// notes), shared by the stand-alone cluster kernel and the fused recurrence+Gram chunk kernel.
#pragma once
#include "common.cuh"
#include "kernels.cuh"
#include "gram_i8.cuh"

namespace xesn {
namespace recur {

#ifdef XESN_RECUR_TRACE
// debug builds only: per-CTA cycle sums of the four phases of a step (thread 0's view)
static __device__ long long g_recur_trace[64 * 8];
__device__ __forceinline__ long long trace_clock() {
    long long t;
    asm volatile("mov.u64 %0, %%clock64;\n" : "=l"(t)::"memory");  // ordered against the other volatile asm (barriers, polls)
    return t;
}
#define XESN_TRACE_MARK(i)                         \
    do {                                           \
        if (tid == 0) {                            \
            const long long now_ = trace_clock();  \
            tr_sum[i] += now_ - tr_last;           \
            tr_last = now_;                        \
        }                                          \
    } while (0)
#else
#define XESN_TRACE_MARK(i) \
    do {                   \
    } while (0)
#endif

constexpr int KMAX = 16;  // gathers issued back-to-back before the first FMA

__device__ __forceinline__ double ld_cg(const double* p) {
    double v;
    asm volatile("ld.global.cg.f64 %0, [%1];\n" : "=d"(v) : "l"(p));
    return v;
}
__device__ __forceinline__ unsigned cluster_ctarank() {
    unsigned r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;\n" : "=r"(r));
    return r;
}
__device__ __forceinline__ unsigned cluster_nctarank() {
    unsigned r;
    asm volatile("mov.u32 %0, %%cluster_nctarank;\n" : "=r"(r));
    return r;
}

// sum_k val[k] * r[idx[k]] in CSR order, multiply and add rounded separately like scipy's
// csr_matvec, with all gathers of the row in flight before the first use.
__device__ __forceinline__ double sparse_row_dot(const int* __restrict__ idx, const double* __restrict__ val,
                                                 int len, const double* r_prev) {
    double g[KMAX];
#pragma unroll
    for (int k = 0; k < KMAX; ++k) g[k] = (k < len) ? ld_cg(r_prev + idx[k]) : 0.0;
    double acc = 0.0;
#pragma unroll
    for (int k = 0; k < KMAX; ++k)
        if (k < len) acc = __dadd_rn(acc, __dmul_rn(val[k], g[k]));
    for (int k = KMAX; k < len; ++k) acc = __dadd_rn(acc, __dmul_rn(val[k], ld_cg(r_prev + idx[k])));
    return acc;
}

__device__ __forceinline__ double leaky(double pre, double r_old, double alpha, double one_minus_alpha) {
    // alpha * tanh(p) + (1 - alpha) * r, each product rounded (no FMA contraction) as numpy does
    return __dadd_rn(__dmul_rn(alpha, tanh(pre)), __dmul_rn(one_minus_alpha, r_old));
}

// ---- data-as-flag exchange of the state vector -------------------------------------------------
// Every CTA of the cluster keeps a full copy of r_t in shared memory (random gathers from shared
// memory cost ~6 cycles per warp; the same gathers from L2 cost one L1 wavefront per lane and
// bounded the first version at ~2.5 us per step).  After computing its slice of r_{t+1} a CTA
// stores it to the state ring in global memory -- which the Gram kernel needs anyway -- and
// refreshes its shared copy with COALESCED polling loads of the other slices: the ring is
// pre-filled with an all-ones bit pattern (a NaN no arithmetic produces here) and an aligned
// 8-byte store is single-copy atomic, so a value that is no longer that pattern IS the finished
// value.  No fence and no cluster barrier sit on the per-step critical path (a release at cluster
// scope compiles to MEMBAR.ALL.GPU per warp: 6.4 us per step in the first version).
constexpr long long UNSET_BITS = -1LL;
// 16-byte pieces in flight per thread while refreshing the copy (64 KB per CTA either way)
template <int RC_THREADS, int UPT>
struct PollCfg {
    static constexpr int BATCH = RC_THREADS <= 256 ? 16 : 8;  // measured at 512 threads, N=16000: 4 and 16 are both slower
};

// The first KREG entries of every owned row of W live in REGISTERS for the whole chunk (value and
// byte offset into the shared state copy): they never change, and reading them from the CSR slice in
// shared memory cost more wavefronts (thread-strided -> bank conflicts) than the state gathers themselves:
// ncu, N=16000: 3200 shared wavefronts per step per CTA for the row dot, of which ~800 are the gathers.
template <int UPT, int RC_THREADS>
struct RowRegs {
    static constexpr int K = RC_THREADS <= 256 ? (UPT <= 4 ? 8 : 4) : (UPT <= 2 ? 8 : (UPT == 4 ? 4 : 2));
};

__device__ __forceinline__ double2 ld_flag2(const double* p) {
    double2 v;
    asm volatile("ld.relaxed.gpu.global.v2.f64 {%0, %1}, [%2];\n" : "=d"(v.x), "=d"(v.y) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ double ld_flag(const double* p) {
    double v;
    asm volatile("ld.relaxed.gpu.global.f64 %0, [%1];\n" : "=d"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ double publishable(double v) {
    return (v != v) ? __longlong_as_double(0x7FF8000000000000LL) : v;
}
__device__ __forceinline__ void st_flag(double* p, double v) {
    v = publishable(v);
    asm volatile("st.relaxed.gpu.global.f64 [%0], %1;\n" ::"l"(p), "d"(v) : "memory");
}
// Every published state is finite or the canonical quiet NaN (0x7FF8...), so the high word
// 0xFFFFFFFF can only be the pre-filled "unset" pattern: one 32-bit compare per element.
__device__ __forceinline__ bool unset(double v) { return __double2hiint(v) == -1; }


// Bounded polling, by wall clock.  `spins` is the caller's per-wait scratch (0 before the first retry): the
// first 32 retries (~10-20 us, far longer than any healthy wait) only count; after that it holds the
// %globaltimer tick (~1 us units, top bit set) at which the slow path was entered, so a wait is bounded in
// TIME (status[1] = limit in microseconds, default 2 s, xesn_set_poll_timeout_ms) however the SM is
// time-sliced or profiled.  Returns true once the run has been flagged as failed (status[0] != 0).
__device__ __forceinline__ unsigned poll_tick_us() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return (unsigned)(t >> 10) & 0x7FFFFFFFu;
}
__device__ __forceinline__ bool poll_expired(int& spins, int* status) {
    if (spins >= 0) {
        if (++spins < 32) return false;
        spins = (int)(poll_tick_us() | 0x80000000u);
        return false;
    }
    const unsigned waited = (poll_tick_us() - (unsigned)spins) & 0x7FFFFFFFu;
    const volatile int* st = reinterpret_cast<const volatile int*>(status);
    if (waited > (unsigned)st[1]) atomicExch(status, 1);
    return st[0] != 0;
}

// publish UPT adjacent states (data-as-flag stores; every 8-byte element is single-copy atomic)
template <int UPT>
__device__ __forceinline__ void put_states(double* dst, const double (&r)[UPT], int nvalid) {
    if (UPT >= 2 && nvalid == UPT && ((reinterpret_cast<uintptr_t>(dst) & 15) == 0)) {
#pragma unroll
        for (int j = 0; j < UPT; j += 2) {
            const double a = publishable(r[j]), b = publishable(r[j + 1]);
            asm volatile("st.relaxed.gpu.global.v2.f64 [%0], {%1, %2};\n" ::"l"(dst + j), "d"(a), "d"(b) : "memory");
        }
    } else {
#pragma unroll
        for (int j = 0; j < UPT; ++j)
            if (j < nvalid) st_flag(dst + j, r[j]);
    }
}

// fixed-point digits of UPT adjacent states -> one packed store per digit plane
template <int UPT>
__device__ __forceinline__ void put_digits(signed char* planes, const RecurArgs& p, int row, int ubase,
                                           const double (&r)[UPT], int nvalid) {
    unsigned w[UPT][2];  // w[j][0] = digits 0..3 of unit j (one byte each), w[j][1] = digits 4..7
#pragma unroll
    for (int j = 0; j < UPT; ++j) gi8::state_digit_words(r[j], w[j][0], w[j][1]);
    signed char* dst = planes + (long long)row * p.ldp + ubase;
    if (nvalid == UPT && UPT >= 2) {
#pragma unroll
        for (int a = 0; a < gi8::NDIG; ++a) {
            const int h = a >> 2, sel = a & 3;  // byte `sel` of word `h`
            if (UPT == 2) {
                const unsigned x = __byte_perm(w[0][h], w[1][h], sel | ((4 + sel) << 4));
                *reinterpret_cast<unsigned short*>(dst + a * p.plane_stride) = (unsigned short)x;
            } else {
#pragma unroll
                for (int j0 = 0; j0 < UPT; j0 += 4) {
                    const unsigned x01 = __byte_perm(w[j0][h], w[j0 + 1][h], sel | ((4 + sel) << 4));
                    const unsigned x23 = __byte_perm(w[j0 + 2][h], w[j0 + 3][h], sel | ((4 + sel) << 4));
                    *reinterpret_cast<unsigned*>(dst + a * p.plane_stride + j0) = __byte_perm(x01, x23, 0x5410);
                }
            }
        }
    } else {
#pragma unroll
        for (int j = 0; j < UPT; ++j)
            if (j < nvalid)
#pragma unroll
                for (int a = 0; a < gi8::NDIG; ++a)
                    dst[a * p.plane_stride + j] = (signed char)((w[j][a >> 2] >> (8 * (a & 3))) & 255u);
    }
}

// Refresh of the shared-memory state copy with coalesced polling loads: PB 16-byte pieces in flight per thread and
// round.  PB is chosen per launch (the smallest that covers the state vector in one round, at most PollCfg::BATCH):
// an unrolled batch of 8 costs ~50 predicated instructions per retry even when a thread has one piece to fetch, and
// the recurrence role is instruction-issue bound.  `mine_all`: bit (round * PB + b) = piece b of that round is this
// thread's to fetch (pieces inside the own slice are already in the copy).
template <int PB, int RC_THREADS>
__device__ __forceinline__ void refresh_copy(const double* r_next, double* s_r, int n2, int tid, unsigned mine_all,
                                             bool packed_masks, int u0, int u1, int* status, bool& aborted) {
    int it = 0;
    for (int q0 = tid; q0 < n2; q0 += RC_THREADS * PB, ++it) {
        double2 v[PB];
        unsigned pending = 0;  // bit b: piece b still has an unset half
        if (packed_masks) {
            pending = (mine_all >> (it * PB)) & ((1u << PB) - 1u);
        } else {
#pragma unroll
            for (int b = 0; b < PB; ++b) {
                const int q = q0 + b * RC_THREADS;
                if ((q < n2) && !(2 * q >= u0 && 2 * q + 1 < u1)) pending |= 1u << b;
            }
        }
        const unsigned mine = pending;
        int spins = 0;
        while (pending && !aborted) {
#pragma unroll
            for (int b = 0; b < PB; ++b)
                if (pending & (1u << b)) v[b] = ld_flag2(r_next + 2 * (q0 + b * RC_THREADS));
#pragma unroll
            for (int b = 0; b < PB; ++b)
                if ((pending & (1u << b)) && !unset(v[b].x) && !unset(v[b].y)) pending &= ~(1u << b);
            if (pending) aborted = poll_expired(spins, status);
        }
#pragma unroll
        for (int b = 0; b < PB; ++b)
            if (mine & (1u << b)) *reinterpret_cast<double2*>(s_r + 2 * (q0 + b * RC_THREADS)) = v[b];
    }
}

// The CTA `c` of the `C` CTAs that own reservoir `g` runs the whole chunk.  RC_THREADS = threads
// taking part (the calling CTA's size); all C CTAs must be co-resident (cluster or cooperative launch).
// SUB > 0: the caller is one of several independent RC_THREADS-wide sub-blocks of a wider CTA (fused_train.cu hosts
// two 256-thread team members of DIFFERENT reservoirs per SM so that one computes while the other waits for its
// peers); it passes its thread index inside the sub-block and synchronises on named barrier SUB instead of barrier 0.
template <int SUB>
__device__ __forceinline__ void sub_block_sync(int n_threads) {
    if (SUB == 0) __syncthreads();
    else asm volatile("bar.sync %0, %1;\n" ::"n"(SUB), "r"(n_threads) : "memory");
}

template <int UPT, int RC_THREADS, bool WS, int SUB = 0>
__device__ __forceinline__ void recur_body(const RecurArgs& p, const unsigned c, const unsigned C, const int g,
                                           unsigned char* smem_raw, const int sub_tid = -1) {
    const int tid = SUB == 0 ? (int)threadIdx.x : sub_tid;
    const int N = p.N;
    const int upc = p.units_per_cta;
    const int u0 = c * upc;
    const int u1 = min(N, u0 + upc);
    if (u0 >= N) return;  // nothing to produce, nobody waits for this CTA

    double* s_r = reinterpret_cast<double*>(smem_raw);  // [n_pad] copy of r_t
    // ---- stage this CTA's CSR slice of W in shared memory (or fall back to global) ----
    const int pg = g % max(p.param_groups, 1);
    const double* const w_values = p.values + (long long)pg * p.values_gs;
    const int slice_lo = p.indptr[u0], slice_hi = p.indptr[u1];
    // WS: the slice lives in shared memory with the column indices pre-multiplied to byte offsets
    // into the state copy; else it is read from global memory (huge or dense W)
    double* s_val = s_r + p.n_pad;
    int* s_off = reinterpret_cast<int*>(s_val + p.smem_nnz);
    if (WS) {
        for (int k = slice_lo + tid; k < slice_hi; k += RC_THREADS) {
            s_val[k - slice_lo] = w_values[k];
            s_off[k - slice_lo] = p.indices[k] * (int)sizeof(double);
        }
    }
    double* states = p.states + (long long)g * p.state_group_stride;
    const double* drive = p.drive + (long long)g * p.drive_group_stride;
    const double* carry = p.carry + (long long)g * N;
    signed char* digits = p.planes ? p.planes + (long long)g * p.plane_group_stride : nullptr;
    for (int i = tid; i < N; i += RC_THREADS) s_r[i] = carry[i];

    // each thread owns UPT ADJACENT units (u0 and units_per_cta are multiples of 8), so its FP64
    // states and its int8 digits go out as vector stores
    const int ubase = u0 + tid * UPT;
    int row_lo[UPT], row_len[UPT];
    double r_cur[UPT], d_next[UPT];
    int nvalid = 0;
#pragma unroll
    for (int j = 0; j < UPT; ++j) {
        const int unit = ubase + j;
        const bool ok = unit < u1;
        nvalid += ok ? 1 : 0;
        row_lo[j] = ok ? p.indptr[unit] : 0;
        row_len[j] = ok ? p.indptr[unit + 1] - row_lo[j] : 0;
        r_cur[j] = ok ? carry[unit] : 0.0;
        d_next[j] = (ok && p.steps > 0) ? drive[unit] : 0.0;
    }
    // register-resident head of each owned row.  Padding entries (rows shorter than KREG) carry weight 0 and point at a
    // dedicated slot of the copy that always holds 0: the row dot then needs no predication (0 * 0 added to the
    // running sum changes nothing, bit for bit), which saves three instructions per entry in an issue-bound loop
    constexpr int KREG = RowRegs<UPT, RC_THREADS>::K;
    constexpr int POLL_BATCH = SUB ? 8 : PollCfg<RC_THREADS, UPT>::BATCH;  // sub-blocks live under a 128-register cap
    const unsigned s_r_addr = (unsigned)__cvta_generic_to_shared(s_r);
    const int zero_off = (p.n_pad - 1) * (int)sizeof(double);  // launchers reserve n_pad >= N + 1 entries
    if (tid == 0) s_r[p.n_pad - 1] = 0.0;
    double hv[UPT][KREG];
    int ho[UPT][KREG];
#pragma unroll
    for (int j = 0; j < UPT; ++j)
#pragma unroll
        for (int k = 0; k < KREG; ++k) {
            const bool in = k < row_len[j];
            hv[j][k] = in ? w_values[row_lo[j] + k] : 0.0;
            ho[j][k] = in ? p.indices[row_lo[j] + k] * (int)sizeof(double) : zero_off;
        }
    // pieces of the state copy this thread refreshes per polling round (own slice excluded); constant over
    // the chunk, packed poll_batch bits per round when all rounds fit one word
    const int n2_all = N >> 1;
    // small reservoirs: two (or four) pieces per thread cover the copy in one round
    const int poll_batch = RC_THREADS * 2 >= n2_all ? 2 : ((SUB && RC_THREADS * 4 >= n2_all) ? 4 : POLL_BATCH);
    const int poll_rounds = (n2_all + RC_THREADS * poll_batch - 1) / (RC_THREADS * poll_batch);
    const bool packed_masks = poll_rounds * poll_batch <= 32;
    unsigned mine_all = 0;
    if (packed_masks)
        for (int it = 0; it < poll_rounds; ++it)
            for (int b = 0; b < poll_batch; ++b) {
                const int q = tid + (it * poll_batch + b) * RC_THREADS;
                if ((q < n2_all) && !(2 * q >= u0 && 2 * q + 1 < u1)) mine_all |= 1u << (it * poll_batch + b);
            }
    put_states<UPT>(states + ubase, r_cur, nvalid);  // row 0 = state entering the chunk (read by the Gram kernel)
    if (digits && p.steps > 0) put_digits<UPT>(digits, p, 0, ubase, r_cur, nvalid);
    const double alpha = p.alpha_g ? p.alpha_g[pg] : p.alpha, oma = 1.0 - alpha;
    const bool vec2 = (N % 2 == 0) && ((reinterpret_cast<uintptr_t>(states) & 15) == 0);
    bool aborted = false;
    sub_block_sync<SUB>(RC_THREADS);

#ifdef XESN_RECUR_TRACE
    long long tr_sum[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    long long tr_last = clock64();
#endif
    for (int t = 0; t < p.steps; ++t) {
        double* r_next = states + (long long)(t + 1) * N;
        // ---- my units: W r_t from the shared copy, tanh, leak; publish to the ring ----
        // (branch-free over the UPT units so their dependent chains interleave)
        double pre[UPT];
        // all UPT*KREG gathers are issued back to back (volatile asm keeps them batched: left to itself the
        // compiler serialises load/use pairs to save registers and the phase becomes 16 shared-memory round
        // trips); padding entries read offset 0 and are masked out of the arithmetic
        double gth[UPT][KREG], hsum[UPT];
#pragma unroll
        for (int j = 0; j < UPT; ++j)
#pragma unroll
            for (int k = 0; k < KREG; ++k)
                asm volatile("ld.shared.f64 %0, [%1];\n" : "=d"(gth[j][k]) : "r"(s_r_addr + (unsigned)ho[j][k]));
#pragma unroll
        for (int j = 0; j < UPT; ++j) {
            double acc = 0.0;
#pragma unroll
            for (int k = 0; k < KREG; ++k) acc = __dadd_rn(acc, __dmul_rn(hv[j][k], gth[j][k]));  // padding: + 0 * 0
            hsum[j] = acc;
        }
        // long rows: the next KTAIL entries are staged together (offsets and values from the CSR slice, then
        // the gathers) so the longest row of the warp costs one more round of loads, not one per entry
        bool any_long = false;
#pragma unroll
        for (int j = 0; j < UPT; ++j) any_long |= row_len[j] > KREG;
        if (__any_sync(0xffffffffu, any_long)) {
            constexpr int KTAIL = 4;
            int to[UPT][KTAIL];
            double tv[UPT][KTAIL], tg[UPT][KTAIL];
#pragma unroll
            for (int j = 0; j < UPT; ++j)
#pragma unroll
                for (int k = 0; k < KTAIL; ++k) {
                    const bool in = KREG + k < row_len[j];
                    const int e = row_lo[j] + KREG + k;
                    if (WS) {
                        to[j][k] = in ? s_off[e - slice_lo] : 0;
                        tv[j][k] = in ? s_val[e - slice_lo] : 0.0;
                    } else {
                        to[j][k] = in ? __ldg(p.indices + e) * (int)sizeof(double) : 0;
                        tv[j][k] = in ? __ldg(w_values + e) : 0.0;
                    }
                }
#pragma unroll
            for (int j = 0; j < UPT; ++j)
#pragma unroll
                for (int k = 0; k < KTAIL; ++k)
                    tg[j][k] = (KREG + k < row_len[j])
                                   ? *reinterpret_cast<const double*>(reinterpret_cast<const char*>(s_r) + to[j][k])
                                   : 0.0;
#pragma unroll
            for (int j = 0; j < UPT; ++j) {
                double acc = hsum[j];
#pragma unroll
                for (int k = 0; k < KTAIL; ++k)
                    if (KREG + k < row_len[j]) acc = __dadd_rn(acc, __dmul_rn(tv[j][k], tg[j][k]));
                if (row_len[j] > KREG + KTAIL) {  // very rare: the rest one by one
                    if (WS) {
                        const int* of = s_off + (row_lo[j] - slice_lo);
                        const double* vl = s_val + (row_lo[j] - slice_lo);
                        for (int k = KREG + KTAIL; k < row_len[j]; ++k)
                            acc = __dadd_rn(acc, __dmul_rn(vl[k], *reinterpret_cast<const double*>(
                                                                      reinterpret_cast<const char*>(s_r) + of[k])));
                    } else {
                        const int* ix = p.indices + row_lo[j];
                        const double* vl = w_values + row_lo[j];
                        for (int k = KREG + KTAIL; k < row_len[j]; ++k)
                            acc = __dadd_rn(acc, __dmul_rn(__ldg(vl + k), s_r[__ldg(ix + k)]));
                    }
                }
                hsum[j] = acc;
            }
        }
#pragma unroll
        for (int j = 0; j < UPT; ++j) {
            double acc = hsum[j];
            pre[j] = __dadd_rn(acc, d_next[j]);
        }
        XESN_TRACE_MARK(4);
#pragma unroll
        for (int j = 0; j < UPT; ++j) r_cur[j] = leaky(pre[j], r_cur[j], alpha, oma);
        XESN_TRACE_MARK(5);
        put_states<UPT>(r_next + ubase, r_cur, nvalid);
        XESN_TRACE_MARK(6);
        // fixed-point digit planes for the int8 tensor-core Gram: row t+1 pairs with y(t0+t+1);
        // the state leaving the chunk (row `steps`) belongs to the next chunk's row 0
        if (digits && t + 1 < p.steps) put_digits<UPT>(digits, p, t + 1, ubase, r_cur, nvalid);
        if (t + 1 == p.steps) break;
#pragma unroll
        for (int j = 0; j < UPT; ++j)
            if (j < nvalid) d_next[j] = __ldg(drive + (long long)(t + 1) * N + ubase + j);
        XESN_TRACE_MARK(0);
        sub_block_sync<SUB>(RC_THREADS);  // every gather of r_t is done: the shared copy may be overwritten
        XESN_TRACE_MARK(1);
#pragma unroll
        for (int j = 0; j < UPT; ++j)
            if (j < nvalid) s_r[ubase + j] = r_cur[j];
        // ---- refresh the other CTAs' slices of r_{t+1} with coalesced polling loads ----
        if (C > 1 && !p.debug_no_exchange) {
            if (vec2) {
                const int n2 = N >> 1;
                if (poll_batch <= 2)
                    refresh_copy<2, RC_THREADS>(r_next, s_r, n2, tid, mine_all, packed_masks, u0, u1, p.status, aborted);
                else if (SUB && poll_batch == 4)
                    refresh_copy<4, RC_THREADS>(r_next, s_r, n2, tid, mine_all, packed_masks, u0, u1, p.status, aborted);
                else
                    refresh_copy<POLL_BATCH, RC_THREADS>(r_next, s_r, n2, tid, mine_all, packed_masks, u0, u1, p.status, aborted);
            } else {
                for (int q = tid; q < N; q += RC_THREADS) {
                    if (q >= u0 && q < u1) continue;
                    double v = ld_flag(r_next + q);
                    int spins = 0;
                    while (unset(v) && !aborted) {
                        aborted = poll_expired(spins, p.status);
                        v = ld_flag(r_next + q);
                    }
                    s_r[q] = v;
                }
            }
        }
        XESN_TRACE_MARK(2);
        sub_block_sync<SUB>(RC_THREADS);
        XESN_TRACE_MARK(3);
    }
#ifdef XESN_RECUR_TRACE
    if (tid == 0 && c < 64)
        for (int i = 0; i < 8; ++i) g_recur_trace[c * 8 + i] = tr_sum[i];
#endif
#pragma unroll
    for (int j = 0; j < UPT; ++j)
        if (j < nvalid) p.carry[(long long)g * N + ubase + j] = r_cur[j];
}

}  // namespace recur
}  // namespace xesn
