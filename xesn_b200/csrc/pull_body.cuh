// Pull-model teacher-forced recurrence (device side): every thread owns UPT adjacent units and
// fetches exactly the states its rows of W need -- straight from the state ring in global memory.
//
// The team model (recur_body.cuh) keeps a full copy of r_t in the shared memory of every CTA of a
// team and refreshes it with coalesced polling loads: 128 KB per CTA and step at N=16000, two
// CTA barriers per step, and a team of ~40 SMs taken away from the Gram role.  But W is fixed for
// the whole run and sparse: a unit needs ~5 states per step, not 16000.  Here a thread polls
// only those (ld.relaxed.gpu on the ring, which is pre-filled with the "unset" pattern -- the data
// is its own flag, recur_body.cuh), so
//   * there is no shared-memory copy, no __syncthreads and no per-CTA geometry: the step chain is
//     store -> L2 -> load plus the row dot, tanh and leak of one thread;
//   * the L2 traffic per step is nnz(W) sectors instead of team x N x 8 bytes;
//   * the units spread over ALL SMs (a few warps per CTA), so the recurrence can ride on the idle
//     warps of the Gram role's CTAs instead of owning SMs (fused_train.cu).
// The arithmetic (CSR order, separately rounded multiply and add, FP64 tanh) is the team model's.
#pragma once
#include "recur_body.cuh"

namespace xesn {
namespace recur {

constexpr int PULL_WARPS = 6;                  // warps per CTA that run the recurrence in the fused kernel
constexpr int PULL_THREADS = PULL_WARPS * 32;  // 192

template <int UPT>
struct PullRegs {
    // register-resident head of every owned row (rows of W average 5 entries; P(len > 12) = 0.2 %)
    static constexpr int K = UPT == 1 ? 12 : (UPT == 2 ? 8 : (UPT == 4 ? 4 : 2));
};

// "Everything of row q is out" HINT.  Polling the ring blindly floods the L2: 16000 threads x 5 loads per round
// queue ~800 requests per L2 slice, a round trip grows from ~300 to ~1750 cycles and the stores of the
// producers wait in the same queues (measured: 8 poll rounds = 15000 cycles per step).  So the threads do not
// poll until a counter says the row is probably complete: the last warp of a CTA to store its part of row q
// adds 1 to a global counter (relaxed, NO fence: the counter carries no ordering, the data stays its own
// flag), one warp per CTA polls that single word and releases the others through shared memory, and then
// one round of gathers normally finds every value set.  A value that is still unset is simply polled again.
__device__ __forceinline__ unsigned ld_hint(const unsigned* p) {
    unsigned v;
    asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];\n" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void add_hint(unsigned* p) {
    asm volatile("red.relaxed.gpu.global.add.u32 [%0], 1;\n" ::"l"(p) : "memory");
}
constexpr int HINT_SPIN_LIMIT = 1 << 16;  // then fall through to polling the data itself

// s_sync[0] = warps of this CTA that have stored their part of the rows so far, s_sync[1] = rows known complete
// wmask: the lanes of this warp that own units (the others have left pull_body and may already sit in a CTA
// barrier of the caller, so they must not be named in a warp barrier)
__device__ __forceinline__ void hint_arrive(const RecurArgs& p, int* s_sync, int lane, unsigned wmask, int active_warps,
                                            int q) {
    __syncwarp(wmask);
    if (lane == 0) {
        const int old = atomicAdd_block(&s_sync[0], 1);
        if (old + 1 == active_warps * (q + 1)) add_hint(p.hint);
    }
}
__device__ __forceinline__ bool hint_wait(const RecurArgs& p, int* s_sync, int lane, unsigned wmask, int q) {
    bool timed_out = false;
    if (lane == 0) {
        // every warp polls the counter itself (staggered pollers: the first to see the row complete tells
        // the other warps of the CTA through shared memory)
        volatile int* ready = s_sync + 1;
        const unsigned target = (unsigned)p.n_active_ctas * (unsigned)(q + 1);
        int spins = 0;
        while (*ready < q + 1) {
            if (ld_hint(p.hint) >= target) {
                *ready = q + 1;
                break;
            }
            if (++spins > HINT_SPIN_LIMIT) {
                timed_out = true;
                break;
            }
        }
    }
    __syncwarp(wmask);
    return timed_out;
}

// rtid: index of this thread among the recurrence threads of its CTA (warp rtid/32 must be entirely made of
// recurrence threads); units_per_cta (a multiple of 8) consecutive units of the flattened (group, unit) space
// belong to CTA `cta`.  Requires N % UPT == 0, so the UPT units of a thread never straddle two groups.
// s_sync: two ints of shared memory, zeroed before the first call (with a CTA barrier in between).
template <int UPT>
__device__ __forceinline__ void pull_body(const RecurArgs& p, const int G, const int rtid, const int cta, int* s_sync) {
    const int N = p.N;
    const long long total = (long long)G * N;
    const int upc = p.units_per_cta;
    if (rtid * UPT >= upc) return;
    const long long f0 = (long long)cta * upc + (long long)rtid * UPT;
    if (f0 >= total) return;
    const int g = (int)(f0 / N), i0 = (int)(f0 % N);
    const int lane = rtid & 31;
    const long long cta_units = min((long long)upc, total - (long long)cta * upc);
    const int active_warps = (int)((cta_units + 32 * UPT - 1) / (32 * UPT));
    const long long warp_units = (cta_units - (long long)(rtid & ~31) * UPT + UPT - 1) / UPT;  // lanes of this warp with units
    const unsigned wmask = warp_units >= 32 ? 0xffffffffu : ((1u << (int)warp_units) - 1u);
    double* states = p.states + (long long)g * p.state_group_stride;
    const double* drive = p.drive + (long long)g * p.drive_group_stride;
    double* carry = p.carry + (long long)g * N;
    signed char* digits = p.planes ? p.planes + (long long)g * p.plane_group_stride : nullptr;
    const int pg = g % max(p.param_groups, 1);
    const double* const w_values = p.values + (long long)pg * p.values_gs;

    constexpr int KREG = PullRegs<UPT>::K;
    constexpr int KTAIL = 4;
    int row_lo[UPT], row_len[UPT];
    double r_cur[UPT], d_next[UPT];
    double hv[UPT][KREG];
    int hi[UPT][KREG];
    unsigned head_mask = 0;
    int max_len = 0;
#pragma unroll
    for (int j = 0; j < UPT; ++j) {
        const int unit = i0 + j;
        row_lo[j] = p.indptr[unit];
        row_len[j] = p.indptr[unit + 1] - row_lo[j];
        max_len = max(max_len, row_len[j]);
        r_cur[j] = carry[unit];
        d_next[j] = p.steps > 0 ? drive[unit] : 0.0;
#pragma unroll
        for (int k = 0; k < KREG; ++k) {
            const bool in = k < row_len[j];
            hv[j][k] = in ? w_values[row_lo[j] + k] : 0.0;
            hi[j][k] = in ? p.indices[row_lo[j] + k] : 0;
            if (in) head_mask |= 1u << (j * KREG + k);
        }
    }
    put_states<UPT>(states + i0, r_cur, UPT);  // row 0 = state entering the chunk
    if (digits && p.steps > 0) put_digits<UPT>(digits, p, 0, i0, r_cur, UPT);
    if (p.hint) hint_arrive(p, s_sync, lane, wmask, active_warps, 0);
    const double alpha = p.alpha_g ? p.alpha_g[pg] : p.alpha, oma = 1.0 - alpha;
    bool aborted = false;
#ifdef XESN_RECUR_TRACE
    const int tid = rtid;  // XESN_TRACE_MARK records thread 0 of the CTA
    long long tr_sum[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    long long tr_last = clock64();
#endif

    for (int t = 0; t < p.steps; ++t) {
        const double* r_t = states + (long long)t * N;
        double acc[UPT];
        XESN_TRACE_MARK(0);  // [0] stores, digits, drive prefetch of the previous step
        // metadata of the first tail batch (entries KREG .. KREG+KTAIL-1 of long rows): constant over the steps,
        // L1-resident, and fetched BEFORE the wait so that its gathers go out together with the heads
        double tv[UPT][KTAIL], tg[UPT][KTAIL];
        int ti[UPT][KTAIL];
        unsigned tpend = 0;
        if (max_len > KREG) {
#pragma unroll
            for (int j = 0; j < UPT; ++j)
#pragma unroll
                for (int k = 0; k < KTAIL; ++k) {
                    const bool in = KREG + k < row_len[j];
                    const int e = row_lo[j] + KREG + k;
                    ti[j][k] = in ? __ldg(p.indices + e) : 0;
                    tv[j][k] = in ? __ldg(w_values + e) : 0.0;
                    tg[j][k] = 0.0;
                    if (in) tpend |= 1u << (j * KTAIL + k);
                }
        }
        {
            double gth[UPT][KREG];
            unsigned pend = head_mask;
            int spins = 0;
            if (p.hint) {
                const bool late = hint_wait(p, s_sync, lane, wmask, t);
#ifdef XESN_RECUR_TRACE
                if (tid == 0 && late) {
                    tr_sum[5] += 1;
                    tr_sum[6] = ((long long)ld_hint(p.hint) << 32) | (unsigned)s_sync[0];
                }
#else
                (void)late;
#endif
            }
            XESN_TRACE_MARK(4);  // [4] waiting for the hint
            while (true) {
#pragma unroll
                for (int j = 0; j < UPT; ++j)
#pragma unroll
                    for (int k = 0; k < KREG; ++k)
                        if (pend & (1u << (j * KREG + k))) gth[j][k] = ld_flag(r_t + hi[j][k]);
                if (tpend) {
#pragma unroll
                    for (int j = 0; j < UPT; ++j)
#pragma unroll
                        for (int k = 0; k < KTAIL; ++k)
                            if (tpend & (1u << (j * KTAIL + k))) tg[j][k] = ld_flag(r_t + ti[j][k]);
                }
#pragma unroll
                for (int j = 0; j < UPT; ++j)
#pragma unroll
                    for (int k = 0; k < KREG; ++k)
                        if ((pend & (1u << (j * KREG + k))) && !unset(gth[j][k])) pend &= ~(1u << (j * KREG + k));
                if (tpend) {
#pragma unroll
                    for (int j = 0; j < UPT; ++j)
#pragma unroll
                        for (int k = 0; k < KTAIL; ++k)
                            if ((tpend & (1u << (j * KTAIL + k))) && !unset(tg[j][k])) tpend &= ~(1u << (j * KTAIL + k));
                }
                if (!(pend | tpend) || aborted) break;
                aborted = poll_expired(spins, p.status);
                if (p.poll_sleep_retry) __nanosleep(p.poll_sleep_retry);
            }
#ifdef XESN_RECUR_TRACE
            if (tid == 0) tr_sum[7] += spins;
#endif
            XESN_TRACE_MARK(1);  // [1] gathering the row entries (warp-wide: until the slowest lane is served)
#pragma unroll
            for (int j = 0; j < UPT; ++j) {
                double a = 0.0;
#pragma unroll
                for (int k = 0; k < KREG; ++k)
                    if (k < row_len[j]) a = __dadd_rn(a, __dmul_rn(hv[j][k], gth[j][k]));
                acc[j] = a;
            }
            if (max_len > KREG) {
#pragma unroll
                for (int j = 0; j < UPT; ++j)
#pragma unroll
                    for (int k = 0; k < KTAIL; ++k)
                        if (KREG + k < row_len[j]) acc[j] = __dadd_rn(acc[j], __dmul_rn(tv[j][k], tg[j][k]));
            }
        }
        // rows longer than KREG + KTAIL (one in 10^5 at 5 entries per row): KTAIL entries of every owned row per round
        for (int kb = KREG + KTAIL; kb < max_len; kb += KTAIL) {
            unsigned pend = 0;
#pragma unroll
            for (int j = 0; j < UPT; ++j)
#pragma unroll
                for (int k = 0; k < KTAIL; ++k) {
                    const bool in = kb + k < row_len[j];
                    const int e = row_lo[j] + kb + k;
                    ti[j][k] = in ? __ldg(p.indices + e) : 0;
                    tv[j][k] = in ? __ldg(w_values + e) : 0.0;
                    tg[j][k] = 0.0;
                    if (in) pend |= 1u << (j * KTAIL + k);
                }
            int spins = 0;
            while (true) {
#pragma unroll
                for (int j = 0; j < UPT; ++j)
#pragma unroll
                    for (int k = 0; k < KTAIL; ++k)
                        if (pend & (1u << (j * KTAIL + k))) tg[j][k] = ld_flag(r_t + ti[j][k]);
#pragma unroll
                for (int j = 0; j < UPT; ++j)
#pragma unroll
                    for (int k = 0; k < KTAIL; ++k)
                        if ((pend & (1u << (j * KTAIL + k))) && !unset(tg[j][k])) pend &= ~(1u << (j * KTAIL + k));
                if (!pend || aborted) break;
                aborted = poll_expired(spins, p.status);
            }
#pragma unroll
            for (int j = 0; j < UPT; ++j)
#pragma unroll
                for (int k = 0; k < KTAIL; ++k)
                    if (kb + k < row_len[j]) acc[j] = __dadd_rn(acc[j], __dmul_rn(tv[j][k], tg[j][k]));
        }
        XESN_TRACE_MARK(2);  // [2] head dot + tails
#pragma unroll
        for (int j = 0; j < UPT; ++j) r_cur[j] = leaky(__dadd_rn(acc[j], d_next[j]), r_cur[j], alpha, oma);
        XESN_TRACE_MARK(3);  // [3] tanh + leak
        put_states<UPT>(states + (long long)(t + 1) * N + i0, r_cur, UPT);
        if (p.hint && t + 1 < p.steps) hint_arrive(p, s_sync, lane, wmask, active_warps, t + 1);
        // fixed-point digit planes for the int8 Gram: row t+1 pairs with y(t0+t+1); the state leaving the
        // chunk (row `steps`) belongs to the next chunk's row 0
        if (digits && t + 1 < p.steps) put_digits<UPT>(digits, p, t + 1, i0, r_cur, UPT);
        if (t + 1 < p.steps) {
#pragma unroll
            for (int j = 0; j < UPT; ++j) d_next[j] = __ldg(drive + (long long)(t + 1) * N + i0 + j);
        }
    }
#pragma unroll
    for (int j = 0; j < UPT; ++j) carry[i0 + j] = r_cur[j];
#ifdef XESN_RECUR_TRACE
    if (tid == 0 && cta < 64)
        for (int i = 0; i < 8; ++i) g_recur_trace[cta * 8 + i] = tr_sum[i];
#endif
}

}  // namespace recur
}  // namespace xesn
