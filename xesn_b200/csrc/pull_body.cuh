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
    static constexpr int K = UPT <= 2 ? 8 : (UPT == 4 ? 4 : 2);  // register-resident head of every owned row
};

// rtid: index of this thread among the recurrence threads of its CTA; units_per_cta (a multiple of 8)
// consecutive units of the flattened (group, unit) space belong to CTA `cta`.  Requires N % UPT == 0, so the
// UPT units of a thread never straddle two groups.
template <int UPT>
__device__ __forceinline__ void pull_body(const RecurArgs& p, const int G, const int rtid, const int cta) {
    const int N = p.N;
    const long long total = (long long)G * N;
    const int upc = p.units_per_cta;
    if (rtid * UPT >= upc) return;
    const long long f0 = (long long)cta * upc + (long long)rtid * UPT;
    if (f0 >= total) return;
    const int g = (int)(f0 / N), i0 = (int)(f0 % N);
    double* states = p.states + (long long)g * p.state_group_stride;
    const double* drive = p.drive + (long long)g * p.drive_group_stride;
    double* carry = p.carry + (long long)g * N;
    signed char* digits = p.planes ? p.planes + (long long)g * p.plane_group_stride : nullptr;

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
            hv[j][k] = in ? p.values[row_lo[j] + k] : 0.0;
            hi[j][k] = in ? p.indices[row_lo[j] + k] : 0;
            if (in) head_mask |= 1u << (j * KREG + k);
        }
    }
    put_states<UPT>(states + i0, r_cur, UPT);  // row 0 = state entering the chunk
    if (digits && p.steps > 0) put_digits<UPT>(digits, p, 0, i0, r_cur, UPT);
    const double alpha = p.alpha, oma = 1.0 - p.alpha;
    bool aborted = false;

    for (int t = 0; t < p.steps; ++t) {
        const double* r_t = states + (long long)t * N;
        double acc[UPT];
        {
            double gth[UPT][KREG];
            unsigned pend = head_mask;
            int spins = 0;
            while (true) {
#pragma unroll
                for (int j = 0; j < UPT; ++j)
#pragma unroll
                    for (int k = 0; k < KREG; ++k)
                        if (pend & (1u << (j * KREG + k))) gth[j][k] = ld_flag(r_t + hi[j][k]);
#pragma unroll
                for (int j = 0; j < UPT; ++j)
#pragma unroll
                    for (int k = 0; k < KREG; ++k)
                        if ((pend & (1u << (j * KREG + k))) && !unset(gth[j][k])) pend &= ~(1u << (j * KREG + k));
                if (!pend || aborted) break;
                aborted = poll_expired(spins, p.status);
            }
#pragma unroll
            for (int j = 0; j < UPT; ++j) {
                double a = 0.0;
#pragma unroll
                for (int k = 0; k < KREG; ++k)
                    if (k < row_len[j]) a = __dadd_rn(a, __dmul_rn(hv[j][k], gth[j][k]));
                acc[j] = a;
            }
        }
        // rows longer than the register-resident head: KTAIL entries of every owned row per round
        for (int kb = KREG; kb < max_len; kb += KTAIL) {
            double tv[UPT][KTAIL], tg[UPT][KTAIL];
            int ti[UPT][KTAIL];
            unsigned pend = 0;
#pragma unroll
            for (int j = 0; j < UPT; ++j)
#pragma unroll
                for (int k = 0; k < KTAIL; ++k) {
                    const bool in = kb + k < row_len[j];
                    const int e = row_lo[j] + kb + k;
                    ti[j][k] = in ? __ldg(p.indices + e) : 0;
                    tv[j][k] = in ? __ldg(p.values + e) : 0.0;
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
#pragma unroll
        for (int j = 0; j < UPT; ++j) r_cur[j] = leaky(__dadd_rn(acc[j], d_next[j]), r_cur[j], alpha, oma);
        put_states<UPT>(states + (long long)(t + 1) * N + i0, r_cur, UPT);
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
}

}  // namespace recur
}  // namespace xesn
