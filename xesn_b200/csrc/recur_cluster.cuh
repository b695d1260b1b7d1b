// Teacher-forced recurrence of one reservoir on a THREAD-BLOCK CLUSTER: the state exchange goes through distributed
// shared memory instead of the L2.
//
// recur_body.cuh exchanges the slices of r_{t+1} through the state ring in global memory: a store becomes visible to
// a polling load of another SM after ~800 cycles (three L2 latencies, tools/micro/pingpong.cu) and every retry of
// a poll costs another L2 round trip, so a step of a small reservoir is ~2 us of which ~0.6 us is arithmetic
// (profiles/recur_trace_r02_before.txt: 2150 + 470 of 4485 cycles are the polling phase).  Inside a cluster a CTA can
// store straight into the shared memory of its peers (DSMEM, ~200 cycles) and the arrival can be counted by the
// receiver's mbarrier: `st.async ... mbarrier::complete_tx::bytes` is a remote store that decrements the transaction
// count of an mbarrier in the DESTINATION CTA when its bytes have landed.  So:
//   * every CTA of the team keeps THREE copies of the state vector in shared memory, used round robin:
//     step t gathers from copy t % 3 (complete), the units' new values go into copy (t+1) % 3 of every CTA (a plain
//     store at home, st.async to the peers), and copy (t+2) % 3 rests -- a peer can run at most one step ahead (it
//     needs all of r_{t+1} before it can produce any of r_{t+2}), so what it sends never lands in a copy that is
//     still being read;
//   * one mbarrier per copy: the owner posts "expect (N - own units) x 8 bytes" once per use, every thread waits on it
//     (try_wait sleeps in hardware: no polling loop, no "unset" pattern, no memset of rings), then ONE CTA barrier
//     orders the local stores -- one barrier and no global-memory round trip on the per-step critical path;
//   * inside a copy, state j sits in slot p.slot[j], chosen on the host so that the gathers of a half-warp fall into
//     different banks (bank_layout.cu): the gather phase is bound by shared-memory wavefronts, and the natural
//     placement costs ~1.6x the conflict-free count;
//   * the state ring in global memory is still written (plain stores: FP64 values for Y R and the next chunk, int8
//     digit planes for the Gram role), but nobody waits for it.
// The arithmetic (CSR order, separately rounded multiply and add, FP64 tanh) is recur_body's, bit for bit.
#pragma once
#include "recur_body.cuh"

namespace xesn {
namespace recur {

constexpr int CL_THREADS = 512;

__device__ __forceinline__ unsigned cl_smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ unsigned cl_mapa(unsigned addr, unsigned rank) {
    unsigned r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;\n" : "=r"(r) : "r"(addr), "r"(rank));
    return r;
}
__device__ __forceinline__ void cl_st_async(unsigned remote_addr, double v, unsigned remote_bar) {
    asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.b64 [%0], %1, [%2];\n" ::"r"(remote_addr),
                 "l"(__double_as_longlong(v)), "r"(remote_bar)
                 : "memory");
}
__device__ __forceinline__ void cl_mbar_init(unsigned bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void cl_mbar_expect(unsigned bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void cl_mbar_arrive(unsigned bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"(bar) : "memory");
}
// bounded wait (time-based like poll_expired): false when the run has been flagged as failed
__device__ __forceinline__ bool cl_mbar_wait(unsigned bar, unsigned parity, int* status) {
    int spins = 0;
    while (true) {
        unsigned ok;
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}\n"
            : "=r"(ok)
            : "r"(bar), "r"(parity)
            : "memory");
        if (ok) return true;
        if (poll_expired(spins, status)) return false;
    }
}
__device__ __forceinline__ void cl_cluster_sync() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;\n" ::: "memory");
}

// bytes of shared memory: three state copies, three mbarriers, the tails of the long rows of W
__host__ __device__ inline long long cluster_smem_bytes(int n_pad, int tail_cap) {
    return 3LL * n_pad * 8 + 64 + (long long)(tail_cap + (tail_cap & 1)) * 12;
}

// CTA `c` of the cluster of `C` CTAs that owns reservoir `g`; one unit per thread (units_per_cta <= CL_THREADS).
// tail_ptr: [C][upc + 1] offsets of every owned row's entries beyond CL_KREG inside this CTA's staging area.
__device__ __forceinline__ void recur_cluster_body(const RecurArgs& p, const unsigned c, const unsigned C, const int g,
                                                   unsigned char* smem_raw, const int* __restrict__ tail_ptr_all,
                                                   int tail_cap) {
    const int tid = threadIdx.x;
    const int N = p.N, n_pad = p.n_pad;
    const int upc = p.units_per_cta;
    const int u0 = min(N, (int)c * upc), u1 = min(N, u0 + upc);
    double* s_buf = reinterpret_cast<double*>(smem_raw);                              // [3][n_pad]
    unsigned long long* s_bar = reinterpret_cast<unsigned long long*>(s_buf + 3LL * n_pad);  // [3] (+ padding to 64 B)
    double* s_tval = reinterpret_cast<double*>(s_bar + 8);                              // [tail_cap]
    int* s_toff = reinterpret_cast<int*>(s_tval + tail_cap + (tail_cap & 1));           // [tail_cap] byte offsets
    const unsigned buf_addr = cl_smem_u32(s_buf), bar_addr = cl_smem_u32(s_bar);
    const unsigned copy_bytes = (unsigned)n_pad * 8u;

    const int pg = g % max(p.param_groups, 1);
    const double* const w_values = p.values + (long long)pg * p.values_gs;
    double* states = p.states + (long long)g * p.state_group_stride;
    const double* drive = p.drive + (long long)g * p.drive_group_stride;
    const double* carry = p.carry + (long long)g * N;
    signed char* digits = p.planes ? p.planes + (long long)g * p.plane_group_stride : nullptr;

    // copy 0 <- the state entering the chunk (every CTA reads the whole vector: no exchange for step 0);
    // every copy ends with a slot that always holds 0 (padding entries of short rows point there)
    for (int i = tid; i < N; i += CL_THREADS) s_buf[p.slot[i]] = carry[i];
    if (tid < 3) {
        s_buf[(long long)tid * n_pad + n_pad - 1] = 0.0;
        cl_mbar_init(bar_addr + 8u * tid, 1);
    }
    const int* tail_ptr = tail_ptr_all + (long long)c * (upc + 1);
    const int unit = u0 + tid;
    const bool owner = unit < u1;
    int row_lo = 0, row_len = 0, tail_lo = 0;
    double r_cur = 0.0, d_next = 0.0;
    if (owner) {
        row_lo = p.indptr[unit];
        row_len = p.indptr[unit + 1] - row_lo;
        tail_lo = tail_ptr[tid];
        r_cur = carry[unit];
        d_next = p.steps > 0 ? drive[unit] : 0.0;
        for (int k = CL_KREG; k < row_len; ++k) {
            s_tval[tail_lo + k - CL_KREG] = w_values[row_lo + k];
            s_toff[tail_lo + k - CL_KREG] = p.col_off[row_lo + k];
        }
    }
    const int zero_off = (n_pad - 1) * (int)sizeof(double);
    double hv[CL_KREG];
    int ho[CL_KREG];
#pragma unroll
    for (int k = 0; k < CL_KREG; ++k) {
        const bool in = k < row_len;
        hv[k] = in ? w_values[row_lo + k] : 0.0;
        ho[k] = in ? p.col_off[row_lo + k] : zero_off;
    }
    const int n_tail = max(row_len - CL_KREG, 0);
    const bool warp_long = __any_sync(0xffffffffu, row_len > CL_KHEAD);
    // staggered halves (warp-uniform): see the step loop
    const int owner_warps = (u1 - u0 + 31) >> 5;
    const bool late = p.stagger > 0 && owner_warps >= 4 && (tid >> 5) >= (owner_warps * p.stagger + 99) / 100;
    double d_after = 0.0;
    const unsigned own_off = owner ? (unsigned)p.slot[unit] * 8u : 0u;   // where this unit's state lives in every copy
    if (owner) {
        states[unit] = r_cur;  // row 0 = state entering the chunk (read by the Gram role / Y R of this chunk)
        if (digits && p.steps > 0) {
            double one[1] = {r_cur};
            put_digits<1>(digits, p, 0, unit, one, 1);
        }
    }
    const double alpha = p.alpha_g ? p.alpha_g[pg] : p.alpha, oma = 1.0 - alpha;
    const unsigned expect_bytes = (unsigned)(N - (u1 - u0)) * 8u;   // what the peers send me per step
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    __syncthreads();
    cl_cluster_sync();   // every CTA's barriers are initialised before anybody sends

#ifdef XESN_RECUR_TRACE
    long long tr_sum[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    long long tr_last = clock64();
#endif
    unsigned cur = 0;        // copy holding r_t
    unsigned phase_bits = 0; // bit b: parity of the next completion of barrier b
    bool ok = true;
    for (int t = 0; t < p.steps; ++t) {
        const unsigned nxt = cur == 2 ? 0u : cur + 1u;
        if (tid == 0) cl_mbar_expect(bar_addr + 8u * nxt, expect_bytes);
        const unsigned src = buf_addr + cur * copy_bytes, dst = buf_addr + nxt * copy_bytes;
        double r_new = 0.0;
        if (owner) {
            if (late) {
                // late half of the warps: the bookkeeping of the PREVIOUS step (ring, digit planes, prefetch) comes first,
                // so these warps reach their gathers -- and, at the other end, their sends -- a few hundred cycles after
                // the early half: the early half's states are on the wire while the late half still computes
                if (t > 0) {
                    states[(long long)t * N + unit] = r_cur;
                    if (digits) {
                        double one[1] = {r_cur};
                        put_digits<1>(digits, p, t, unit, one, 1);
                    }
                }
                d_after = t + 1 < p.steps ? __ldg(drive + (long long)(t + 1) * N + unit) : 0.0;
            }
            double gth[CL_KREG];
#pragma unroll
            for (int k = 0; k < CL_KHEAD; ++k)
                asm volatile("ld.shared.f64 %0, [%1];\n" : "=d"(gth[k]) : "r"(src + (unsigned)ho[k]));
            if (warp_long) {   // warp-uniform
#pragma unroll
                for (int k = CL_KHEAD; k < CL_KREG; ++k)
                    asm volatile("ld.shared.f64 %0, [%1];\n" : "=d"(gth[k]) : "r"(src + (unsigned)ho[k]));
            }
            double acc = 0.0;
#pragma unroll
            for (int k = 0; k < CL_KHEAD; ++k) acc = __dadd_rn(acc, __dmul_rn(hv[k], gth[k]));  // padding: + 0 * 0
            if (warp_long) {
#pragma unroll
                for (int k = CL_KHEAD; k < CL_KREG; ++k) acc = __dadd_rn(acc, __dmul_rn(hv[k], gth[k]));
            }
            // long rows: CL_KTAIL more entries per round, staged together (offsets and values, then the gathers) so the
            // longest row of the warp costs one more round of loads, not one per entry; padding -> 0 * 0 again
            for (int k0 = 0; k0 < n_tail; k0 += CL_KTAIL) {
                int to[CL_KTAIL];
                double tv[CL_KTAIL], tg[CL_KTAIL];
#pragma unroll
                for (int k = 0; k < CL_KTAIL; ++k) {
                    const bool in = k0 + k < n_tail;
                    to[k] = in ? s_toff[tail_lo + k0 + k] : zero_off;
                    tv[k] = in ? s_tval[tail_lo + k0 + k] : 0.0;
                }
#pragma unroll
                for (int k = 0; k < CL_KTAIL; ++k)
                    asm volatile("ld.shared.f64 %0, [%1];\n" : "=d"(tg[k]) : "r"(src + (unsigned)to[k]));
#pragma unroll
                for (int k = 0; k < CL_KTAIL; ++k) acc = __dadd_rn(acc, __dmul_rn(tv[k], tg[k]));
            }
            XESN_TRACE_MARK(4);
            r_new = leaky(__dadd_rn(acc, d_next), r_cur, alpha, oma);
            r_cur = r_new;
            XESN_TRACE_MARK(5);
            // home copy, then the peers' copies (each store signs off 8 bytes on the receiver's barrier)
            const unsigned off = own_off;
            asm volatile("st.shared.f64 [%0], %1;\n" ::"r"(dst + off), "d"(r_new) : "memory");
            for (unsigned q = 0; q < C; ++q)
                if (q != c) cl_st_async(cl_mapa(dst + off, q), r_new, cl_mapa(bar_addr + 8u * nxt, q));
            XESN_TRACE_MARK(6);
            // global ring + digit planes: for the Gram role and Y R (nobody polls them)
            if (!late) {
                states[(long long)(t + 1) * N + unit] = r_new;
                if (digits && t + 1 < p.steps) {
                    double one[1] = {r_new};
                    put_digits<1>(digits, p, t + 1, unit, one, 1);
                }
                if (t + 1 < p.steps) d_next = __ldg(drive + (long long)(t + 1) * N + unit);
            } else {
                d_next = d_after;
            }
        }
        XESN_TRACE_MARK(0);
        if (C > 1 && ok) ok = cl_mbar_wait(bar_addr + 8u * nxt, (phase_bits >> nxt) & 1u, p.status);
        phase_bits ^= 1u << nxt;
        XESN_TRACE_MARK(2);
        __syncthreads();   // the home stores of this CTA are in place, and every gather of r_t is done
        XESN_TRACE_MARK(3);
        cur = nxt;
    }
#ifdef XESN_RECUR_TRACE
    if (tid == 0 && c < 64)
        for (int i = 0; i < 8; ++i) g_recur_trace[c * 8 + i] = tr_sum[i];
#endif
    if (owner) {
        if (late && p.steps > 0) states[(long long)p.steps * N + unit] = r_cur;   // the last row of the late half
        p.carry[(long long)g * N + unit] = r_cur;
    }
    cl_cluster_sync();   // nobody leaves while a peer may still store into its shared memory
}

}  // namespace recur
}  // namespace xesn
