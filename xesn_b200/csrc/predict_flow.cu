// Dataflow closed-loop forecast of SEVERAL reservoirs (LazyESN groups, ensemble members) -- sm_100a.
//
//      for n = 1 .. n_steps:   r_g <- update(r_g, gather_g(field));   field[out_map_g] <- Wout_g r_g
//      (reference xesn/lazyesn.py:226-236: _update_nd, _readout and the re-halo `overlap` per step)
//
// predict_loop.cu runs the same loop with two software grid barriers per step (all 148 CTAs meet twice,
// ~8-11 us per step for 16 groups of N=2000) and, across GPUs, stores every predicted cell to every rank
// behind a system-scope fence and a flag handshake.  But a step of group g only depends on
//   (1) the new state of g itself, spread over the `team` CTAs that own g, and
//   (2) the new outputs of the few cells g reads: its own and the halo cells of its neighbours.
// So nothing here is a barrier.  A team of CTAs owns one group; a CTA keeps its rows of Win and its COLUMNS of
// Wout in shared memory, and after updating its units it forms its contribution to every output of the
// group,  part[o][c] = sum_{i in slice c} Wout[o][i] r[i],  and publishes it together with its slice of r.
// Rings in global memory pre-filled with the "unset" pattern carry both (the data is its own flag,
// recur_body.cuh): a consumer polls exactly the values it needs and adds the `team` contributions of an
// output in CTA order (deterministic).  One store -> poll hop per step instead of two grid barriers.
//
// CLUSTER variant (teams of <= 8 CTAs that are all resident as thread-block clusters): the slices of r travel through
// distributed shared memory instead of the state ring -- three copies of r per CTA, `st.async` + mbarrier exactly as in
// recur_cluster.cuh -- so the per-step critical path holds no L2 round trip for the states; the contributions to the
// outputs still go through the partial rings (they are what other groups, other ranks and the trajectory read).
//
// Multi-GPU (one process per GPU, groups split in slabs): a cell that a group on ANOTHER rank reads is
// pushed, contribution by contribution, into that rank's ring with plain 8-byte NVLink P2P stores
// (peer-mapped memory, CUDA IPC) -- only those cells, only to the ranks that read them (the `overlap`-deep
// strips of SURVEY.md section 8e), no fence and no flag: the imported slot is polled like a local one.
#include <algorithm>
#include <climits>

#include "common.cuh"
#include "kernels.cuh"
#include "recur_body.cuh"
#include "recur_cluster.cuh"

namespace xesn {
namespace {

using namespace recur;

constexpr int PF_THREADS = 512;
constexpr int PF_WARPS = PF_THREADS / 32;
constexpr int PF_IO_THREADS = 64;                       // the last two warps poll inputs / trajectory values
constexpr int PF_R_THREADS = PF_THREADS - PF_IO_THREADS;  // the others refresh the state copy
constexpr int PF_KREG = 8, PF_KTAIL = 4;
constexpr int PF_POLL_BATCH = 4;

__device__ __forceinline__ double ld_flag_sys(const double* p) {
    double v;
    asm volatile("ld.relaxed.sys.global.f64 %0, [%1];\n" : "=d"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_flag_sys(double* p, double v) {
    v = publishable(v);
    asm volatile("st.relaxed.sys.global.f64 [%0], %1;\n" ::"l"(p), "d"(v) : "memory");
}

// sum over the team of the contributions to ring slot `slot` at ring row `row`, in CTA order
// (slots written by peer GPUs are polled at system scope, local ones at GPU scope)
__device__ __forceinline__ double team_sum(const PredictFlowArgs& p, long long slot, int row, bool& aborted) {
    const double* q = p.part_ring + (slot * p.ring_steps + row) * p.team;
    const bool imported = slot >= (long long)p.G * p.O;
    double v[XESN_FLOW_MAX_TEAM];
    unsigned pend = (p.team >= 32) ? 0xffffffffu : ((1u << p.team) - 1u);
    int spins = 0;
    while (true) {
#pragma unroll
        for (int c = 0; c < XESN_FLOW_MAX_TEAM; ++c)
            if (pend & (1u << c)) v[c] = imported ? ld_flag_sys(q + c) : ld_flag(q + c);
#pragma unroll
        for (int c = 0; c < XESN_FLOW_MAX_TEAM; ++c)
            if ((pend & (1u << c)) && !unset(v[c])) pend &= ~(1u << c);
        if (!pend || aborted) break;
        aborted = poll_expired(spins, p.status);
    }
    double s = 0.0;
#pragma unroll
    for (int c = 0; c < XESN_FLOW_MAX_TEAM; ++c)
        if (c < p.team) s += v[c];
    return s;
}

// CTA-wide meeting points of the two warp roles (named barriers: the roles run different loops, so they cannot
// share __syncthreads call sites)
__device__ __forceinline__ void bar_inputs_ready() { asm volatile("bar.sync 1, %0;\n" ::"n"(PF_THREADS) : "memory"); }
__device__ __forceinline__ void bar_states_ready() { asm volatile("bar.sync 2, %0;\n" ::"n"(PF_THREADS) : "memory"); }

struct FlowSmem {
    double* s_r;     // [n_pad] copy of r_{n-1}
    double* s_win;   // [upc][ldw]
    double* s_wout;  // [O][upc]
    double* s_u;     // [I + (I & 1)]
    double* s_rnew;  // [upc]
    double* s_tval;  // [tail_cap] entries of W beyond the register-resident head of each owned row ...
    int* s_toff;     // [tail_cap] ... and their columns
    int* s_src;      // [I]
    int ldw;
    // cluster variant: the team's contributions to this group's outputs, three copies like the states ([3][team][O]),
    // and the shared-window address of the three mbarriers
    double* s_part;
    unsigned bar_addr;
};

// phase C: this CTA's contribution to every output of the group (all 16 warps)
template <bool CL>
__device__ __forceinline__ void flow_contributions(const PredictFlowArgs& p, const FlowSmem& sm, int n, int c, long long slot0,
                                                   int warp, int lane, unsigned nxt) {
    const int upc = p.upc;
    for (int o = warp; o < p.O; o += PF_WARPS) {
        const double* w = sm.s_wout + (size_t)o * upc;
        double a = 0.0;
        for (int il = lane; il < upc; il += 32) a = fma(w[il], sm.s_rnew[il], a);
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) a += __shfl_xor_sync(0xffffffffu, a, off);
        if (CL && lane < p.team) {
            // cluster variant: lane q hands the contribution to CTA q of the team (itself included) through distributed
            // shared memory, signed off on the same mbarrier as the states of this step -- the group's own cells reach
            // the next step's inputs without a round trip through the L2
            const unsigned dst = cl_smem_u32(sm.s_part + ((size_t)nxt * p.team + c) * p.O + o);
            cl_st_async(cl_mapa(dst, (unsigned)lane), a, cl_mapa(sm.bar_addr + 8u * nxt, (unsigned)lane));
        }
        if (lane == 0) {
            const long long slot = slot0 + o;
            st_flag(p.part_ring + (slot * p.ring_steps + (n - 1)) * p.team + c, a);
            // cells that groups on other ranks read: the same contribution, into their rings (NVLink P2P store)
            if (p.exp_ptr)
                for (int e = p.exp_ptr[slot]; e < p.exp_ptr[slot + 1]; ++e)
                    st_flag_sys(p.peer_ring[p.exp_peer[e]] + ((long long)p.exp_slot[e] * p.ring_steps + (n - 1)) * p.team + c, a);
        }
    }
}

// warps 14 and 15: the values the step consumes that come through the partial rings
template <bool CL>
__device__ __forceinline__ void flow_io_role(const PredictFlowArgs& p, const FlowSmem& sm, int g, int c, int t /*0..63*/) {
    const int I = p.I, O = p.O;
    unsigned cur = 0, phase_bits = 0;   // CL: copy that holds the contributions of step n-1; parities of the mbarriers
    const int warp = PF_R_THREADS / 32 + (t >> 5), lane = t & 31;
    const bool leader = c == 0;
    const long long slot0 = (long long)g * O;
    bool aborted = false;
#ifdef XESN_RECUR_TRACE
    long long io_sum[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    long long io_last = trace_clock();
#define XESN_IO_MARK(i)                                   \
    do {                                                  \
        if (t == 0) {                                     \
            const long long now_ = trace_clock();         \
            io_sum[i] += now_ - io_last, io_last = now_;  \
        }                                                 \
    } while (0)
#else
#define XESN_IO_MARK(i) \
    do {                \
    } while (0)
#endif
    for (int n = 1; n <= p.steps; ++n) {
        const unsigned nxt = cur == 2 ? 0u : cur + 1u;
        if (t < 32) {
            if (CL && n >= 2 && !aborted) {   // the team's contributions of step n-1 have landed in copy `cur`
                aborted = !cl_mbar_wait(sm.bar_addr + 8u * cur, (phase_bits >> cur) & 1u, p.status);
                phase_bits ^= 1u << cur;
            }
            // inputs of step n: the outputs of step n-1 (ring row n-2), or the field entering the launch
            for (int k = t; k < I; k += 32) {
                const int src = sm.s_src[k];
                double x;
                if (src < 0 && src != INT_MIN) {
                    x = p.fill[-1 - src];
                } else if (n == 1 || src == INT_MIN) {
                    x = p.field[p.in_map[(long long)g * I + k]];
                } else if (CL && src >= slot0 && src < slot0 + O) {
                    // a cell of this group: the contributions are in shared memory, added in CTA order like team_sum
                    const double* q = sm.s_part + (size_t)cur * p.team * O + (src - slot0);
                    x = 0.0;
                    for (int cc = 0; cc < p.team; ++cc) x += q[(size_t)cc * O];
                } else {
                    x = team_sum(p, src, n - 2, aborted);
                }
                sm.s_u[k] = (p.mask_nan && x != x) ? 0.0 : x;
            }
        } else if (leader && n >= 2) {
            // trajectory column of step n-1 (beside the input polls, not after them)
            for (int o = t - 32; o < O; o += 32) {
                const long long cell = p.out_map ? (long long)p.out_map[slot0 + o] : slot0 + o;
                p.pred[cell * p.ldp + p.col0 + n - 1] = team_sum(p, slot0 + o, n - 2, aborted);
            }
        }
        XESN_IO_MARK(0);   // [0] polling the inputs
        bar_inputs_ready();
        XESN_IO_MARK(1);   // [1] waiting for the owners to arrive (they were still busy with the previous step)
        bar_states_ready();
        XESN_IO_MARK(2);   // [2] the owners' update
        flow_contributions<CL>(p, sm, n, c, slot0, warp, lane, nxt);
        XESN_IO_MARK(3);   // [3] contributions
        cur = CL ? nxt : 0u;
    }
#ifdef XESN_RECUR_TRACE
    if (t == 0 && blockIdx.x < 32)
        for (int i = 0; i < 8; ++i) g_recur_trace[(32 + blockIdx.x) * 8 + i] = io_sum[i];
#endif
    // last trajectory column, the field leaving the launch
    if (leader)
        for (int o = t; o < O; o += PF_IO_THREADS) {
            const long long cell = p.out_map ? (long long)p.out_map[slot0 + o] : slot0 + o;
            const double v = team_sum(p, slot0 + o, p.steps - 1, aborted);
            p.pred[cell * p.ldp + p.col0 + p.steps] = v;
            p.field_out[cell] = v;
        }
    if (blockIdx.x == 0)  // imported cells: what this rank's groups read at step 1 of the next launch
        for (int j = t; j < p.n_import; j += PF_IO_THREADS)
            p.field_out[p.import_cell[j]] = team_sum(p, (long long)p.G * O + j, p.steps - 1, aborted);
}

template <bool CL>
__global__ void __launch_bounds__(PF_THREADS, 1) predict_flow_kernel(PredictFlowArgs p) {
    extern __shared__ __align__(16) double smem[];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int g = blockIdx.x / p.team, c = blockIdx.x % p.team;
    const int N = p.N, I = p.I, O = p.O, upc = p.upc;
    const int u0 = c * upc, u1 = min(N, u0 + upc), nu = u1 - u0;
    const int n_pad = (N + 1) & ~1;
    FlowSmem sm;
    sm.ldw = I + 1 + (I & 1);  // odd row pitch (in doubles): conflict-free reads of one row per lane
    sm.s_r = smem;   // CL: three copies of the state vector, then the mbarrier of each copy
    unsigned long long* s_bar = reinterpret_cast<unsigned long long*>(smem + 3 * (size_t)n_pad);
    sm.s_part = smem + 3 * (size_t)n_pad + 8;
    sm.s_win = CL ? sm.s_part + 3 * (size_t)p.team * O : sm.s_r + n_pad;
    sm.s_wout = sm.s_win + (size_t)upc * sm.ldw;
    sm.s_u = sm.s_wout + (size_t)O * upc;
    sm.s_rnew = sm.s_u + I + (I & 1);
    sm.s_tval = sm.s_rnew + upc;
    sm.s_toff = reinterpret_cast<int*>(sm.s_tval + p.tail_cap);
    sm.s_src = sm.s_toff + p.tail_cap + (p.tail_cap & 1);

    const int pg = g % max(p.param_groups, 1);
    {
        const double* win = p.Win + (long long)pg * p.win_gs;
        const double* wout = p.Wout + (long long)g * O * N;
        for (int e = tid; e < nu * I; e += PF_THREADS) sm.s_win[(e / I) * sm.ldw + (e % I)] = win[(long long)(u0 + e / I) * I + e % I];
        for (int e = tid; e < O * upc; e += PF_THREADS) {
            const int o = e / upc, il = e % upc;
            sm.s_wout[e] = il < nu ? wout[(long long)o * N + u0 + il] : 0.0;
        }
        for (int k = tid; k < I; k += PF_THREADS) sm.s_src[k] = p.src[(long long)g * I + k];
    }
    // in/out states are different arrays: a fast CTA may finish before a team mate has loaded
    const double* r_g = p.r_in + (long long)g * N;
    for (int i = tid; i < n_pad; i += PF_THREADS) sm.s_r[i] = i < N ? r_g[i] : 0.0;
    const double* w_values = p.values + (long long)pg * p.values_gs;
    // tails of the long rows of this CTA (entries beyond PF_KREG), packed row after row
    const int* tail_ptr = p.tail_ptr + (long long)c * (upc + 1);
    for (int il = tid; il < nu; il += PF_THREADS) {
        const int lo = p.indptr[u0 + il], len = p.indptr[u0 + il + 1] - lo;
        for (int k = PF_KREG; k < len; ++k) {
            const int dst = tail_ptr[il] + (k - PF_KREG);
            sm.s_tval[dst] = w_values[lo + k];
            sm.s_toff[dst] = p.indices[lo + k];
        }
    }
    const unsigned bar_addr = CL ? cl_smem_u32(s_bar) : 0u;
    sm.bar_addr = bar_addr;
    if (CL) {
        if (tid < 3) cl_mbar_init(bar_addr + 8u * tid, 1);
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    }
    __syncthreads();
    if (CL) cl_cluster_sync();   // every CTA's barriers exist before anybody sends

    if (tid >= PF_R_THREADS) {
        flow_io_role<CL>(p, sm, g, c, tid - PF_R_THREADS);
        if (CL) cl_cluster_sync();   // nobody leaves while a peer may still store into its shared memory
        return;
    }

    // ---- owners: one unit per thread, head of its row of W in registers ----
    const bool owner = tid < nu;
    const int unit = u0 + tid;
    int row_len = 0, tail_lo = 0;
    double r_cur = 0.0, bias = 0.0;
    double hv[PF_KREG];
    int ho[PF_KREG];
    {
        int row_lo = 0;
        if (owner) {
            row_lo = p.indptr[unit];
            row_len = p.indptr[unit + 1] - row_lo;
            tail_lo = tail_ptr[tid];
            r_cur = r_g[unit];
            bias = p.bias[(long long)pg * p.bias_gs + unit];
        }
#pragma unroll
        for (int k = 0; k < PF_KREG; ++k) {
            const bool in = k < row_len;
            hv[k] = in ? w_values[row_lo + k] : 0.0;
            ho[k] = in ? p.indices[row_lo + k] * (int)sizeof(double) : 0;
        }
    }
    const unsigned s_r_addr = (unsigned)__cvta_generic_to_shared(sm.s_r);
    const double alpha = p.alpha_g ? p.alpha_g[pg] : p.alpha, oma = 1.0 - alpha;
    const long long slot0 = (long long)g * O;      // ring slots of this group's outputs
    double* r_ring = p.r_ring + (long long)g * p.ring_steps * N;
    const bool vec2 = (N % 2 == 0) && ((reinterpret_cast<uintptr_t>(r_ring) & 15) == 0);
    bool aborted = false;
#ifdef XESN_RECUR_TRACE
    long long tr_sum[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    long long tr_last = trace_clock();
#endif

    const unsigned copy_bytes = (unsigned)n_pad * 8u;
    // per step: the peers' slices of r and every team member's contributions (mine included) to the group's outputs
    const unsigned expect_bytes = (unsigned)(N - nu) * 8u + (unsigned)(p.team * O) * 8u;
    unsigned cur = 0, phase_bits = 0;                         // CL: copy holding r_{n-1}; parities of the three barriers
    for (int n = 1; n <= p.steps; ++n) {
        const unsigned nxt = cur == 2 ? 0u : cur + 1u;
        if (CL) {
            // ================= phase A (cluster): r_{n-1} has been stored into copy `cur` by its owners =================
            if (tid == 0) cl_mbar_expect(bar_addr + 8u * nxt, expect_bytes);
            if (n >= 2 && p.team > 1 && !aborted) {
                aborted = !cl_mbar_wait(bar_addr + 8u * cur, (phase_bits >> cur) & 1u, p.status);
                phase_bits ^= 1u << cur;
            }
        } else
        // ================= phase A: the other slices of r_{n-1} (ring row n-2), coalesced data-as-flag polling =================
        if (n >= 2) {
            const double* row = r_ring + (long long)(n - 2) * N;
            if (owner) sm.s_r[unit] = r_cur;
            if (vec2) {
                const int n2 = N >> 1;
                for (int q0 = tid; q0 < n2; q0 += PF_R_THREADS * PF_POLL_BATCH) {
                    double2 v[PF_POLL_BATCH];
                    unsigned pend = 0;
#pragma unroll
                    for (int b = 0; b < PF_POLL_BATCH; ++b) {
                        const int q = q0 + b * PF_R_THREADS;
                        if (q < n2 && !(2 * q >= u0 && 2 * q + 1 < u1)) pend |= 1u << b;
                    }
                    const unsigned mine = pend;
                    int spins = 0;
                    while (pend && !aborted) {
#pragma unroll
                        for (int b = 0; b < PF_POLL_BATCH; ++b)
                            if (pend & (1u << b)) v[b] = ld_flag2(row + 2 * (q0 + b * PF_R_THREADS));
#pragma unroll
                        for (int b = 0; b < PF_POLL_BATCH; ++b)
                            if ((pend & (1u << b)) && !unset(v[b].x) && !unset(v[b].y)) pend &= ~(1u << b);
                        if (pend) aborted = poll_expired(spins, p.status);
                    }
#pragma unroll
                    for (int b = 0; b < PF_POLL_BATCH; ++b)
                        if (mine & (1u << b)) *reinterpret_cast<double2*>(sm.s_r + 2 * (q0 + b * PF_R_THREADS)) = v[b];
                }
            } else {
                for (int q = tid; q < N; q += PF_R_THREADS) {
                    if (q >= u0 && q < u1) continue;
                    double v = ld_flag(row + q);
                    int spins = 0;
                    while (unset(v) && !aborted) {
                        aborted = poll_expired(spins, p.status);
                        v = ld_flag(row + q);
                    }
                    sm.s_r[q] = v;
                }
            }
        }
        XESN_TRACE_MARK(0);  // [0] refresh of the state copy (thread 0's pieces)

        // ================= phase B1: W r_{n-1} of the owned units =================
        // needs the states only, not the inputs: done BEFORE the inputs barrier, while the input-polling warps still wait
        // for the contributions of the previous step to come through the L2 (the owners used to idle there for 39 % of a
        // step, profiles/predict_flow_cluster_r02b.txt)
        if (!CL && n >= 2) asm volatile("bar.sync 3, %0;\n" ::"n"(PF_R_THREADS) : "memory");   // every piece of s_r is in place
        double wr = 0.0;
        if (owner) {
            // all gathers of the row head in flight before the first use (volatile asm keeps them batched)
            double gth[PF_KREG];
#pragma unroll
            for (int k = 0; k < PF_KREG; ++k)
                asm volatile("ld.shared.f64 %0, [%1];\n" : "=d"(gth[k]) : "r"(s_r_addr + (CL ? cur * copy_bytes : 0u) + (unsigned)ho[k]));
#pragma unroll
            for (int k = 0; k < PF_KREG; ++k)
                if (k < row_len) wr = __dadd_rn(wr, __dmul_rn(hv[k], gth[k]));
            XESN_TRACE_MARK(2);  // [2] head of the row dot
            for (int k = PF_KREG; k < row_len; ++k)
                wr = __dadd_rn(wr, __dmul_rn(sm.s_tval[tail_lo + k - PF_KREG],
                                             sm.s_r[(CL ? (size_t)cur * n_pad : 0) + sm.s_toff[tail_lo + k - PF_KREG]]));
            XESN_TRACE_MARK(3);  // [3] tail of long rows
        }
        bar_inputs_ready();
        XESN_TRACE_MARK(1);  // [1] waiting for the inputs

        // ================= phase B2: input projection, tanh, publish =================
        if (owner) {
            const double* wrow = sm.s_win + (size_t)tid * sm.ldw;
            double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
            int k = 0;
#pragma unroll 2
            for (; k + 3 < I; k += 4) {
                a0 = fma(wrow[k], sm.s_u[k], a0);
                a1 = fma(wrow[k + 1], sm.s_u[k + 1], a1);
                a2 = fma(wrow[k + 2], sm.s_u[k + 2], a2);
                a3 = fma(wrow[k + 3], sm.s_u[k + 3], a3);
            }
            for (; k < I; ++k) a0 = fma(wrow[k], sm.s_u[k], a0);
            const double pre = __dadd_rn(__dadd_rn(wr, (a0 + a1) + (a2 + a3)), bias);
            XESN_TRACE_MARK(4);  // [4] Win u
            r_cur = leaky(pre, r_cur, alpha, oma);
            XESN_TRACE_MARK(5);  // [5] tanh + leak
            if (CL) {
                // home copy, then the peers' copies (each store signs off 8 bytes on the receiver's barrier)
                const unsigned dst = s_r_addr + nxt * copy_bytes + (unsigned)unit * 8u;
                asm volatile("st.shared.f64 [%0], %1;\n" ::"r"(dst), "d"(r_cur) : "memory");
                for (unsigned q = 0; q < (unsigned)p.team; ++q)
                    if (q != (unsigned)c) cl_st_async(cl_mapa(dst, q), r_cur, cl_mapa(bar_addr + 8u * nxt, q));
            } else {
                st_flag(r_ring + (long long)(n - 1) * N + unit, r_cur);
            }
            sm.s_rnew[tid] = r_cur;
        } else if (tid < upc) {
            sm.s_rnew[tid] = 0.0;
        }
        bar_states_ready();
        XESN_TRACE_MARK(6);  // [6] state store + waiting for the other owners
        flow_contributions<CL>(p, sm, n, c, slot0, warp, lane, nxt);
        XESN_TRACE_MARK(7);  // [7] contributions
        cur = CL ? nxt : 0u;
    }
    // what was sent during the last step must have landed before anybody leaves (nobody else waits for it)
    if (CL && p.steps >= 1 && !aborted) cl_mbar_wait(bar_addr + 8u * cur, (phase_bits >> cur) & 1u, p.status);
#ifdef XESN_RECUR_TRACE
    if (tid == 0 && blockIdx.x < 32)
        for (int i = 0; i < 8; ++i) g_recur_trace[blockIdx.x * 8 + i] = tr_sum[i];
#endif
    if (owner) p.r_out[(long long)g * N + unit] = r_cur;
    if (CL) cl_cluster_sync();
}

// src[g][k]: ring slot that feeds input k of group g (single GPU: every produced cell is local)
__global__ void flow_owner_kernel(const int* __restrict__ out_map, int* __restrict__ owner, long long n_slots) {
    const long long s = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (s < n_slots) owner[out_map ? out_map[s] : s] = (int)s;
}
__global__ void flow_src_kernel(const int* __restrict__ in_map, const int* __restrict__ owner, int* __restrict__ src,
                                long long n) {
    const long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    const int m = in_map[k];
    src[k] = m < 0 ? m : (owner[m] >= 0 ? owner[m] : INT_MIN);
}

}  // namespace

size_t predict_flow_smem(int N, int I, int O, int upc, int tail_cap, int cluster) {
    const int ldw = I + 1 + (I & 1);
    // cluster variant: three copies of the states, the mbarriers, three copies of the team's contributions (team <= 8)
    const size_t n_pad = cluster ? 3 * (size_t)((N + 1) & ~1) + 8 + 3 * 8 * (size_t)O : (size_t)((N + 1) & ~1);
    return sizeof(double) * (n_pad + (size_t)upc * ldw + (size_t)O * upc + (size_t)(I + (I & 1)) + upc + (size_t)tail_cap) +
           sizeof(int) * (size_t)(tail_cap + (tail_cap & 1) + I + 2);
}

// team size / units per CTA of the dataflow forecast; false when the configuration does not fit (caller falls back)
// clusters of `team` CTAs with `smem` bytes each that the device keeps resident at once (0 on error)
static int flow_resident_clusters(int team, size_t smem) {
    if (cudaFuncSetAttribute(predict_flow_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)team), cfg.blockDim = dim3(PF_THREADS), cfg.dynamicSmemBytes = smem;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = (unsigned)team, attr[0].val.clusterDim.y = 1, attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr, cfg.numAttrs = 1;
    int n = 0;
    if (cudaOccupancyMaxActiveClusters(&n, predict_flow_kernel<true>, &cfg) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return n;
}

static bool flow_fits(int N, int I, int O, int team, int cluster, const int* indptr, int* upc_out, int* tail_cap_out) {
    const int upc = ((N + team - 1) / team + 7) & ~7;
    if ((N + upc - 1) / upc != team) return false;   // a team of this size would leave a member without units
    if (upc > PF_R_THREADS) return false;            // one unit per thread of the owner warps
    // entries of W beyond the register-resident head of each row, per CTA slice (shared-memory staging)
    int tail_cap = 0;
    for (int c = 0; c < team; ++c) {
        int t = 0;
        for (int i = c * upc; i < std::min(N, (c + 1) * upc); ++i) t += std::max(0, indptr[i + 1] - indptr[i] - PF_KREG);
        tail_cap = std::max(tail_cap, t);
    }
    if (predict_flow_smem(N, I, O, upc, tail_cap, cluster) > 200 * 1024) return false;
    *upc_out = upc, *tail_cap_out = tail_cap;
    return true;
}

bool predict_flow_geometry(int N, int I, int O, int G, int num_sms, const int* indptr, int* team_out, int* upc_out,
                           int* tail_cap_out, int* cluster_out) {
    if (cluster_out) *cluster_out = 0;
    if (G < 1 || G > num_sms) return false;
    // cluster variant first: the largest team (<= 8 CTAs) whose clusters are ALL resident -- clusters cannot straddle
    // GPCs, so the device holds fewer CTAs in clusters than SMs (measured here: 15 clusters of 8) -- because groups
    // wait for each other's outputs
    const char* e = getenv("XESN_B200_FLOW_CLUSTER");
    if (cluster_out && !(e && e[0] == '0')) {
        for (int team = 8; team >= 2; --team) {   // any size works, not only powers of two
            if ((long long)team * G > num_sms) continue;
            if (!flow_fits(N, I, O, team, 1, indptr, upc_out, tail_cap_out)) continue;
            const int resident = flow_resident_clusters(team, predict_flow_smem(N, I, O, *upc_out, *tail_cap_out, 1));
            if (getenv("XESN_B200_VERBOSE"))
                fprintf(stderr, "xesn_b200: dataflow forecast, clusters of %d: %d resident, %d groups\n", team, resident, G);
            if (resident < G) continue;
            *team_out = team, *cluster_out = 1;
            return true;
        }
    }
    int team = std::min(num_sms / G, XESN_FLOW_MAX_TEAM);
    if (team < 1) return false;
    const int upc = ((N + team - 1) / team + 7) & ~7;
    team = (N + upc - 1) / upc;
    if (!flow_fits(N, I, O, team, 0, indptr, upc_out, tail_cap_out)) return false;
    *team_out = team;
    return true;
}

// tail_ptr[c][il] = offset of unit (c*upc + il)'s tail inside its CTA's staging area, [team][upc+1] ints (host array)
void predict_flow_tail_ptr(int N, int team, int upc, const int* indptr, int* out) {
    for (int c = 0; c < team; ++c) {
        int t = 0;
        for (int il = 0; il <= upc; ++il) {
            out[(long long)c * (upc + 1) + il] = t;
            const int i = c * upc + il;
            if (il < upc && i < N) t += std::max(0, indptr[i + 1] - indptr[i] - PF_KREG);
        }
    }
}

int launch_flow_src(const int* in_map, const int* out_map, int* owner, int* src, long long n_cells, long long n_slots,
                    long long n_inputs, cudaStream_t stream) {
    XESN_CUDA_OK(cudaMemsetAsync(owner, 0xFF, sizeof(int) * (size_t)n_cells, stream));
    flow_owner_kernel<<<(unsigned)((n_slots + 255) / 256), 256, 0, stream>>>(out_map, owner, n_slots);
    flow_src_kernel<<<(unsigned)((n_inputs + 255) / 256), 256, 0, stream>>>(in_map, owner, src, n_inputs);
    XESN_CUDA_OK(cudaGetLastError());
    count_launch(2);
    return 0;
}

int launch_predict_flow(const PredictFlowArgs& a, cudaStream_t stream) {
    if (a.steps <= 0) return 0;
    PredictFlowArgs args = a;
    if (a.cluster) {
        const size_t smem = predict_flow_smem(a.N, a.I, a.O, a.upc, a.tail_cap, 1);
        XESN_CUDA_OK(cudaFuncSetAttribute(predict_flow_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3((unsigned)(a.G * a.team)), cfg.blockDim = dim3(PF_THREADS), cfg.dynamicSmemBytes = smem, cfg.stream = stream;
        // Not a cooperative launch: that bound ignores cluster placement and refuses grids this one needs (and a refused
        // launch is fatal under a profiler).  The geometry was chosen from cudaOccupancyMaxActiveClusters, so every
        // cluster is resident on an otherwise idle device; should the launch be refused all the same, the L2 variant
        // runs with the same team (the state ring is armed for it in either case).
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = (unsigned)a.team, attr[0].val.clusterDim.y = 1, attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr, cfg.numAttrs = 1;
        if (cudaLaunchKernelEx(&cfg, predict_flow_kernel<true>, args) == cudaSuccess) {
            count_launch();
            return 0;
        }
        cudaGetLastError();
        args.cluster = 0;
    }
    {
        XESN_CUDA_OK(cudaFuncSetAttribute(predict_flow_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        const size_t smem = predict_flow_smem(a.N, a.I, a.O, a.upc, a.tail_cap, 0);
        void* kargs[] = {&args};
        XESN_CUDA_OK(cudaLaunchCooperativeKernel((void*)predict_flow_kernel<false>, dim3(a.G * a.team), dim3(PF_THREADS), kargs,
                                                 smem, stream));
    }
    count_launch();
    return 0;
}

}  // namespace xesn

#ifdef XESN_RECUR_TRACE
extern "C" int xesn_debug_trace_flow(long long* out) {
    return (int)cudaMemcpyFromSymbol(out, xesn::recur::g_recur_trace, sizeof(long long) * 64 * 8);
}
#endif
