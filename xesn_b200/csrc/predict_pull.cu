// Barrier-free closed-loop forecast of ONE reservoir (eager ESN, reference xesn/esn.py:264-268):
//
//      for n = 1 .. n_steps:   r <- update(r, v);   v <- Wout r
//
// predict_loop.cu separates the update and the readout of a step by two software grid barriers
// (13 us per step at N=16000).  Here the step has no grid barrier.  The reservoir is spread over all
// SMs (~N/148 units per CTA, one unit per thread).  A CTA keeps its rows of Win and its COLUMNS of Wout in
// shared memory, so after updating its units it forms its contribution to every output,
//      part_c[o] = sum_{i in slice c} Wout[o][i] r[i],
// and publishes it next to its slice of r.  Output o is finished by ONE dedicated reducer warp (CTA
// o % n_ctas) that adds the contributions of all CTAs in a fixed order and publishes v[o].  Everything
// travels through rings in global memory pre-filled with the "unset" pattern (the data is its own flag,
// recur_body.cuh); the many consumers of v are gated by the relaxed hint counter of pull_body.cuh (one
// word polled by one thread per CTA), the few reducer warps poll the contributions themselves.
// Measured on B200 (tools/micro/pingpong.cu): one store -> poll hop between SMs costs ~800 cycles, so a step
// is two hops (contributions -> reducer -> everyone) plus the arithmetic, instead of two grid barriers.
//
// Warp roles (512 threads): warps 0-3 own the units, warps 4-11 compute the contributions, warps 12-15 are
// reducers (free-running: they never join the CTA barriers of the other twelve).
#include <algorithm>

#include "common.cuh"
#include "kernels.cuh"
#include "pull_body.cuh"

namespace xesn {
namespace {

using namespace recur;

constexpr int PP_THREADS = 512;
constexpr int PP_UNITS = 128;         // warps 0..3, one unit per thread
constexpr int PP_CONTRIB_WARP0 = 4, PP_CONTRIB_WARPS = 8;
constexpr int PP_RED_WARP0 = 12, PP_RED_WARPS = 4;
constexpr int PP_MAIN_THREADS = PP_RED_WARP0 * 32;  // 384 threads meet at named barrier 1
constexpr int PP_KREG = 12;
constexpr int PP_KTAIL = 4;

__device__ __forceinline__ void main_barrier() { asm volatile("bar.sync 1, %0;\n" ::"n"(PP_MAIN_THREADS) : "memory"); }

__device__ __forceinline__ double poll_value(const double* p, int* status, bool& aborted) {
    double v = ld_flag(p);
    int spins = 0;
    while (unset(v) && !aborted) {
        aborted = poll_expired(spins, status);
        v = ld_flag(p);
    }
    return v;
}
__device__ __forceinline__ double nan_to_zero(double x) { return (x != x) ? 0.0 : x; }

__global__ void __launch_bounds__(PP_THREADS, 1) predict_pull_kernel(PredictPullArgs p) {
    extern __shared__ __align__(16) double smem[];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int c = blockIdx.x;
    const int N = p.N, I = p.I, O = p.O, upc = p.upc;
    const int u0 = c * upc, u1 = min(N, u0 + upc), nu = u1 - u0;
    const int ldw = I + 1 + (I & 1);  // odd row pitch (in doubles): conflict-free reads of one row per lane
    double* s_win = smem;                          // [upc][ldw]
    double* s_wout = s_win + (size_t)upc * ldw;    // [O][upc]
    double* s_u0 = s_wout + (size_t)O * upc;       // [I] inputs that never change (and all inputs of step 1)
    double* s_u = s_u0 + I;                        // [I]
    double* s_rnew = s_u + I;                      // [upc]
    int* s_src = reinterpret_cast<int*>(s_rnew + upc);  // [I] output that feeds input k, or -1 (static input)

    for (int e = tid; e < nu * I; e += PP_THREADS) s_win[(e / I) * ldw + (e % I)] = p.Win[(long long)(u0 + e / I) * I + e % I];
    for (int e = tid; e < O * upc; e += PP_THREADS) {
        const int o = e / upc, il = e % upc;
        s_wout[e] = il < nu ? p.Wout[(long long)o * N + u0 + il] : 0.0;
    }
    // input k reads cell in_map[k] of the field (or a boundary fill value); the cell is rewritten every step by the
    // output o with out_map[o] == cell, if there is one (NaN -> 0 as in lazyesn.py:362-364)
    for (int k = tid; k < I; k += PP_THREADS) {
        const int m = p.in_map[k];
        int src = -1;
        if (m >= 0)
            for (int o = 0; o < O; ++o)
                if ((p.out_map ? p.out_map[o] : o) == m) src = o;
        s_src[k] = src;
        const double x0 = m >= 0 ? p.field[m] : p.fill[-1 - m];
        s_u0[k] = p.mask_nan ? nan_to_zero(x0) : x0;
    }

    const double alpha = p.alpha, oma = 1.0 - p.alpha;
    bool aborted = false;
    __syncthreads();

    // ================= reducer warps: v_n[o] = sum over the CTAs of part[n][o][cta] =================
    if (warp >= PP_RED_WARP0) {
        const int red_o = c + (warp - PP_RED_WARP0) * p.n_ctas;
        if (red_o >= O) return;
        const long long cell = p.out_map ? (long long)p.out_map[red_o] : (long long)red_o;
        constexpr int RB = 5;  // 160 CTAs at most
        for (int n = 1; n <= p.steps; ++n) {
            const double* row = p.part_ring + ((long long)(n - 1) * O + red_o) * p.n_ctas;
            double v[RB];
            unsigned pend = 0;
#pragma unroll
            for (int b = 0; b < RB; ++b) {
                v[b] = 0.0;
                if (b * 32 + lane < p.n_ctas) pend |= 1u << b;
            }
            int spins = 0;
            while (true) {
#pragma unroll
                for (int b = 0; b < RB; ++b)
                    if (pend & (1u << b)) v[b] = ld_flag(row + b * 32 + lane);
#pragma unroll
                for (int b = 0; b < RB; ++b)
                    if ((pend & (1u << b)) && !unset(v[b])) pend &= ~(1u << b);
                if (!pend || aborted) break;
                aborted = poll_expired(spins, p.status);
            }
            double sum = 0.0;
#pragma unroll
            for (int b = 0; b < RB; ++b) sum += v[b];
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, off);
            if (lane == 0) {
                st_flag(p.v_ring + (long long)(n - 1) * O + red_o, sum);
                add_hint(p.hint);
                p.pred[cell * p.ldp + p.col0 + n] = sum;
                if (n == p.steps) p.field[cell] = sum;
            }
        }
        return;
    }

    // ================= the other twelve warps =================
    const bool owner = tid < PP_UNITS && tid < nu;
    const int unit = u0 + tid;
    int row_lo = 0, row_len = 0;
    double r_cur = 0.0, bias = 0.0;
    double hv[PP_KREG];
    int hi[PP_KREG];
    unsigned head_mask = 0;
    if (owner) {
        row_lo = p.indptr[unit];
        row_len = p.indptr[unit + 1] - row_lo;
        r_cur = p.r[unit];
        bias = p.bias[unit];
    }
#pragma unroll
    for (int k = 0; k < PP_KREG; ++k) {
        const bool in = k < row_len;
        hv[k] = in ? p.values[row_lo + k] : 0.0;
        hi[k] = in ? p.indices[row_lo + k] : 0;
        if (in) head_mask |= 1u << k;
    }
    if (owner) st_flag(p.r_ring + unit, r_cur);  // row 0 = state after spin-up
#ifdef XESN_RECUR_TRACE
    long long tr_sum[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    long long tr_last = trace_clock();
#endif

    for (int n = 1; n <= p.steps; ++n) {
        XESN_TRACE_MARK(4);  // owners: idle while the contributions are formed (added to [4])
        // the hint counts published outputs, O per step: complete => every CTA has published its contribution to
        // every output, which it does after storing its slice of r
        if (n >= 2 && tid == PP_CONTRIB_WARP0 * 32) {
            const unsigned target = (unsigned)O * (unsigned)(n - 1);
            int spins = 0;
#ifdef XESN_RECUR_TRACE
            const long long h0 = trace_clock();
#endif
            while (ld_hint(p.hint) < target)
                if (++spins > HINT_SPIN_LIMIT) break;
#ifdef XESN_RECUR_TRACE
            tr_sum[6] += trace_clock() - h0;
            tr_sum[7] += spins;
#endif
        }
        main_barrier();
        XESN_TRACE_MARK(0);  // [0] hint wait
        const double* r_prev = p.r_ring + (long long)(n - 1) * N;

        // (1) owners: gathers of W r in flight; contribution warps: the input vector from v_{n-1}
        double gth[PP_KREG], tg[PP_KTAIL], tv[PP_KTAIL];
        int ti[PP_KTAIL];
        unsigned pend = owner ? head_mask : 0u, tpend = 0;
        if (pend && row_len > PP_KREG) {
#pragma unroll
            for (int k = 0; k < PP_KTAIL; ++k) {
                const bool in = PP_KREG + k < row_len;
                ti[k] = in ? __ldg(p.indices + row_lo + PP_KREG + k) : 0;
                tv[k] = in ? __ldg(p.values + row_lo + PP_KREG + k) : 0.0;
                tg[k] = 0.0;
                if (in) tpend |= 1u << k;
            }
        }
#pragma unroll
        for (int k = 0; k < PP_KREG; ++k)
            if (pend & (1u << k)) gth[k] = ld_flag(r_prev + hi[k]);
        if (tpend) {
#pragma unroll
            for (int k = 0; k < PP_KTAIL; ++k)
                if (tpend & (1u << k)) tg[k] = ld_flag(r_prev + ti[k]);
        }
        if (warp >= PP_CONTRIB_WARP0) {
#ifdef XESN_RECUR_TRACE
            const long long v0 = trace_clock();
#endif
            for (int k = tid - PP_CONTRIB_WARP0 * 32; k < I; k += PP_CONTRIB_WARPS * 32) {
                const int src = s_src[k];
                double x = s_u0[k];
                if (n > 1 && src >= 0) {
                    x = poll_value(p.v_ring + (long long)(n - 2) * O + src, p.status, aborted);
                    if (p.mask_nan) x = nan_to_zero(x);
                }
                s_u[k] = x;
            }
#ifdef XESN_RECUR_TRACE
            tr_sum[5] += trace_clock() - v0;
#endif
        }
        main_barrier();
        XESN_TRACE_MARK(1);  // [1] issue gathers, input vector

        // (2) update of the owned unit
        if (owner) {
            int spins = 0;
            while (true) {
#pragma unroll
                for (int k = 0; k < PP_KREG; ++k)
                    if ((pend & (1u << k)) && !unset(gth[k])) pend &= ~(1u << k);
                if (tpend) {
#pragma unroll
                    for (int k = 0; k < PP_KTAIL; ++k)
                        if ((tpend & (1u << k)) && !unset(tg[k])) tpend &= ~(1u << k);
                }
                if (!(pend | tpend) || aborted) break;
                aborted = poll_expired(spins, p.status);
#pragma unroll
                for (int k = 0; k < PP_KREG; ++k)
                    if (pend & (1u << k)) gth[k] = ld_flag(r_prev + hi[k]);
#pragma unroll
                for (int k = 0; k < PP_KTAIL; ++k)
                    if (tpend & (1u << k)) tg[k] = ld_flag(r_prev + ti[k]);
            }
            double wr = 0.0;
#pragma unroll
            for (int k = 0; k < PP_KREG; ++k)
                if (k < row_len) wr = __dadd_rn(wr, __dmul_rn(hv[k], gth[k]));
            if (row_len > PP_KREG) {
#pragma unroll
                for (int k = 0; k < PP_KTAIL; ++k)
                    if (PP_KREG + k < row_len) wr = __dadd_rn(wr, __dmul_rn(tv[k], tg[k]));
                for (int k = PP_KREG + PP_KTAIL; k < row_len; ++k)  // one row in 10^5
                    wr = __dadd_rn(wr, __dmul_rn(__ldg(p.values + row_lo + k),
                                                 poll_value(r_prev + __ldg(p.indices + row_lo + k), p.status, aborted)));
            }
            XESN_TRACE_MARK(3);  // [3] gather wait + row dot
            // Win u: four interleaved chains
            const double* wrow = s_win + (size_t)tid * ldw;
            double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
            int k = 0;
#pragma unroll 2
            for (; k + 3 < I; k += 4) {
                a0 = fma(wrow[k], s_u[k], a0);
                a1 = fma(wrow[k + 1], s_u[k + 1], a1);
                a2 = fma(wrow[k + 2], s_u[k + 2], a2);
                a3 = fma(wrow[k + 3], s_u[k + 3], a3);
            }
            for (; k < I; ++k) a0 = fma(wrow[k], s_u[k], a0);
            const double pre = __dadd_rn(__dadd_rn(wr, (a0 + a1) + (a2 + a3)), bias);
            r_cur = leaky(pre, r_cur, alpha, oma);
            st_flag(p.r_ring + (long long)n * N + unit, r_cur);
            s_rnew[tid] = r_cur;
        } else if (tid < upc) {
            s_rnew[tid] = 0.0;
        }
        main_barrier();
        XESN_TRACE_MARK(4);  // [4] Win u + tanh + store + barrier

        // (3) contribution warps: this CTA's share of every output; warp w takes outputs w, w+8, ..., four at a
        // time, lanes stride over the units.  part_ring[step][o][cta]: a reducer reads one contiguous row
        if (warp >= PP_CONTRIB_WARP0) {
            double* part = p.part_ring + (long long)(n - 1) * O * p.n_ctas + c;
            for (int o0 = warp - PP_CONTRIB_WARP0; o0 < O; o0 += 4 * PP_CONTRIB_WARPS) {
                const double* w0 = s_wout + (size_t)o0 * upc;
                const double* w1 = s_wout + (size_t)min(o0 + PP_CONTRIB_WARPS, O - 1) * upc;
                const double* w2 = s_wout + (size_t)min(o0 + 2 * PP_CONTRIB_WARPS, O - 1) * upc;
                const double* w3 = s_wout + (size_t)min(o0 + 3 * PP_CONTRIB_WARPS, O - 1) * upc;
                double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
                for (int il = lane; il < upc; il += 32) {
                    const double rv = s_rnew[il];
                    a0 = fma(w0[il], rv, a0);
                    a1 = fma(w1[il], rv, a1);
                    a2 = fma(w2[il], rv, a2);
                    a3 = fma(w3[il], rv, a3);
                }
#pragma unroll
                for (int off = 16; off > 0; off >>= 1) {
                    a0 += __shfl_xor_sync(0xffffffffu, a0, off);
                    a1 += __shfl_xor_sync(0xffffffffu, a1, off);
                    a2 += __shfl_xor_sync(0xffffffffu, a2, off);
                    a3 += __shfl_xor_sync(0xffffffffu, a3, off);
                }
                if (lane == 0) {
                    st_flag(part + (long long)o0 * p.n_ctas, a0);
                    if (o0 + PP_CONTRIB_WARPS < O) st_flag(part + (long long)(o0 + PP_CONTRIB_WARPS) * p.n_ctas, a1);
                    if (o0 + 2 * PP_CONTRIB_WARPS < O) st_flag(part + (long long)(o0 + 2 * PP_CONTRIB_WARPS) * p.n_ctas, a2);
                    if (o0 + 3 * PP_CONTRIB_WARPS < O) st_flag(part + (long long)(o0 + 3 * PP_CONTRIB_WARPS) * p.n_ctas, a3);
                }
            }
        }
    }
#ifdef XESN_RECUR_TRACE
    if (tid == 0 && c < 64)
        for (int i = 0; i < 5; ++i) g_recur_trace[c * 8 + i] = tr_sum[i];
    if (tid == PP_CONTRIB_WARP0 * 32 && c < 64)
        for (int i = 5; i < 8; ++i) g_recur_trace[c * 8 + i] = tr_sum[i];
#endif
    if (owner) p.r[unit] = r_cur;
}

}  // namespace

size_t predict_pull_smem(int I, int O, int upc, long long n_cells, int n_ctas) {
    const int ldw = I + 1 + (I & 1);
    (void)n_ctas, (void)n_cells;
    return sizeof(double) * ((size_t)upc * ldw + (size_t)O * upc + 2 * (size_t)I + upc) + sizeof(int) * (size_t)(I + 1);
}

// geometry of the pull forecast; returns false when the configuration does not fit (caller falls back)
bool predict_pull_geometry(int N, int I, int O, long long n_cells, int num_sms, int* upc_out, int* n_ctas_out) {
    const int upc = ((N + num_sms - 1) / num_sms + 7) & ~7;
    if (upc > PP_UNITS) return false;
    const int n_ctas = (N + upc - 1) / upc;
    if (O > PP_RED_WARPS * n_ctas || n_ctas > 160) return false;  // one reducer warp per output, five contributions per lane
    if (predict_pull_smem(I, O, upc, n_cells, n_ctas) > 200 * 1024) return false;
    *upc_out = upc, *n_ctas_out = n_ctas;
    return true;
}

int launch_predict_pull(PredictPullArgs a, cudaStream_t stream) {
    if (a.steps <= 0) return 0;
    const size_t smem = predict_pull_smem(a.I, a.O, a.upc, a.n_cells, a.n_ctas);
    XESN_CUDA_OK(cudaFuncSetAttribute(predict_pull_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    XESN_CUDA_OK(cudaMemsetAsync(a.hint, 0, sizeof(unsigned), stream));
    XESN_CUDA_OK(cudaMemsetAsync(a.r_ring, 0xFF, sizeof(double) * (size_t)(a.steps + 1) * a.N, stream));
    XESN_CUDA_OK(cudaMemsetAsync(a.part_ring, 0xFF, sizeof(double) * (size_t)a.steps * a.n_ctas * a.O, stream));
    XESN_CUDA_OK(cudaMemsetAsync(a.v_ring, 0xFF, sizeof(double) * (size_t)a.steps * a.O, stream));
    void* kargs[] = {&a};
    XESN_CUDA_OK(cudaLaunchCooperativeKernel((void*)predict_pull_kernel, dim3(a.n_ctas), dim3(PP_THREADS), kargs, smem, stream));
    count_launch();
    return 0;
}

}  // namespace xesn

#ifdef XESN_RECUR_TRACE
extern "C" int xesn_debug_trace_predict(long long* out) {
    return (int)cudaMemcpyFromSymbol(out, xesn::recur::g_recur_trace, sizeof(long long) * 64 * 8);
}
#endif
