// Barrier-free closed-loop forecast of ONE reservoir (eager ESN, reference xesn/esn.py:264-268):
//
//      for n = 1 .. n_steps:   r <- update(r, v);   v <- Wout r
//
// predict_loop.cu separates the update and the readout of a step by two software grid barriers
// (13 us per step at N=16000).  Here the step has no barrier at all.  The reservoir is spread over all
// SMs (~N/148 units per CTA, one unit per thread).  A CTA keeps its rows of Win and its COLUMNS of Wout in
// shared memory, so after updating its units it can form its contribution to the readout,
//      part_c[o] = sum_{i in slice c} Wout[o][i] r[i],
// right away and publish it next to its slice of r.  Both travel through rings in global memory that are
// pre-filled with the "unset" pattern (the data is its own flag, recur_body.cuh), gated by the relaxed
// hint counter of pull_body.cuh so that nobody polls before the row is probably complete.  Every CTA
// then adds the contributions of all CTAs in the same fixed order (bitwise the same v everywhere) and
// gathers exactly the states its rows of W need.  Per step: one L2 hop instead of two grid barriers.
#include <algorithm>

#include "common.cuh"
#include "kernels.cuh"
#include "pull_body.cuh"

namespace xesn {
namespace {

using namespace recur;

constexpr int PP_THREADS = 512;
constexpr int PP_WARPS = PP_THREADS / 32;
constexpr int PP_UNITS = 128;  // unit threads per CTA (warps 0..3), one unit each; the other warps help with the readout
constexpr int PP_KREG = 12;
constexpr int PP_KTAIL = 4;

__device__ __forceinline__ double poll_value(const double* p, int* status, bool& aborted) {
    double v = ld_flag(p);
    int spins = 0;
    while (unset(v) && !aborted) {
        aborted = poll_expired(spins, status);
        v = ld_flag(p);
    }
    return v;
}

__global__ void __launch_bounds__(PP_THREADS, 1) predict_pull_kernel(PredictPullArgs p) {
    extern __shared__ __align__(16) double smem[];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int c = blockIdx.x;
    const int N = p.N, I = p.I, O = p.O, upc = p.upc;
    const int u0 = c * upc, u1 = min(N, u0 + upc), nu = u1 - u0;
    const int ldw = I + 1 + (I & 1);  // odd row pitch (in doubles): conflict-free reads of one row per lane
    double* s_win = smem;                          // [upc][ldw]
    double* s_wout = s_win + (size_t)upc * ldw;    // [O][upc]
    double* s_field = s_wout + (size_t)O * upc;    // [n_cells]
    double* s_u = s_field + p.n_cells;             // [I]
    double* s_rnew = s_u + I;                      // [upc]

    for (int e = tid; e < nu * I; e += PP_THREADS) s_win[(e / I) * ldw + (e % I)] = p.Win[(long long)(u0 + e / I) * I + e % I];
    for (int e = tid; e < O * upc; e += PP_THREADS) {
        const int o = e / upc, il = e % upc;
        s_wout[e] = il < nu ? p.Wout[(long long)o * N + u0 + il] : 0.0;
    }
    for (int e = tid; e < p.n_cells; e += PP_THREADS) s_field[e] = p.field[e];

    // ---- per-unit constants (warps 0..3) ----
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
    // output o is reduced by the LAST warps of CTA o % n_ctas (the first warps own the units)
    const int red_o = c + (PP_WARPS - 1 - warp) * p.n_ctas;
    const bool reducer = warp >= PP_UNITS / 32 && red_o < O;
    const long long red_cell = reducer ? (p.out_map ? (long long)p.out_map[red_o] : (long long)red_o) : 0;
    const double alpha = p.alpha, oma = 1.0 - p.alpha;
    bool aborted = false;
#ifdef XESN_RECUR_TRACE
    long long tr_sum[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    long long tr_last = clock64();
#endif
    __syncthreads();

    // step n consumes row n-1 of the state ring and v_{n-1} (the field of step 0 is the initial condition); one
    // extra pass (n = steps + 1) only fetches the last readout so that CTA 0 can store the final field
    for (int n = 1; n <= p.steps + 1; ++n) {
        const bool last = n == p.steps + 1;
        if (last && c != 0) break;
        XESN_TRACE_MARK(5);  // [5] readout contribution, publish, (reducer warps: the sum) of the previous step
        // the hint counts the published outputs: O per step.  Complete => every CTA has published its contribution
        // to every output, which it does after storing its slice of r
        if (n >= 2 && tid == 0) {
            const unsigned target = (unsigned)O * (unsigned)(n - 1);
            int spins = 0;
            while (ld_hint(p.hint) < target)
                if (++spins > HINT_SPIN_LIMIT) break;
        }
        __syncthreads();
        XESN_TRACE_MARK(0);  // [0] hint wait
        const double* r_prev = p.r_ring + (long long)(n - 1) * N;

        // (1) gathers of W r in flight first
        double gth[PP_KREG], tg[PP_KTAIL], tv[PP_KTAIL];
        int ti[PP_KTAIL];
        unsigned pend = (owner && !last) ? head_mask : 0u, tpend = 0;
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

        // (2) v_{n-1} into the field
        if (n >= 2) {
            for (int o = tid; o < O; o += PP_THREADS) {
                const double v = poll_value(p.v_ring + (long long)(n - 2) * O + o, p.status, aborted);
                s_field[p.out_map ? p.out_map[o] : o] = v;
            }
            __syncthreads();
        }
        if (last) break;
        XESN_TRACE_MARK(1);  // [1] issue gathers + fetch v

        // (3) input vector of the step: gathered from the field, NaN -> 0 (lazyesn.py:362-364 for masked inputs)
        for (int k = tid; k < I; k += PP_THREADS) {
            const int m = __ldg(p.in_map + k);
            const double x = (m >= 0) ? s_field[m] : __ldg(p.fill + (-1 - m));
            s_u[k] = (x != x) ? 0.0 : x;
        }
        __syncthreads();
        XESN_TRACE_MARK(2);  // [2] input vector

        // (4) update of the owned unit
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
        __syncthreads();
        XESN_TRACE_MARK(4);  // [4] Win u + tanh + store + barrier

        // (5) this CTA's contribution to every output: warp w takes outputs w, w+16, ...; lanes stride over the units.
        // part_ring[step][o][cta]: the reducer of output o reads one contiguous row
        {
            double* part = p.part_ring + (long long)(n - 1) * O * p.n_ctas + c;
            for (int o0 = 0; o0 < O; o0 += PP_WARPS * 4) {
                double acc[4] = {0.0, 0.0, 0.0, 0.0};
                for (int il = lane; il < upc; il += 32) {
                    const double rv = s_rnew[il];
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const int o = o0 + j * PP_WARPS + warp;
                        if (o < O) acc[j] = fma(s_wout[(size_t)o * upc + il], rv, acc[j]);
                    }
                }
#pragma unroll
                for (int off = 16; off > 0; off >>= 1)
#pragma unroll
                    for (int j = 0; j < 4; ++j) acc[j] += __shfl_xor_sync(0xffffffffu, acc[j], off);
                if (lane == 0)
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const int o = o0 + j * PP_WARPS + warp;
                        if (o < O) st_flag(part + (long long)o * p.n_ctas, acc[j]);
                    }
            }
        }
        // (6) reducer warps: v_n[o] = sum over the CTAs, fixed order; few warps poll, so they poll the data itself
        if (reducer) {
            const double* row = p.part_ring + ((long long)(n - 1) * O + red_o) * p.n_ctas;
            double sum = 0.0;
            for (int cb = 0; cb < p.n_ctas; cb += 32 * 8) {
                double v[8];
#pragma unroll
                for (int b = 0; b < 8; ++b) {
                    const int cc = cb + b * 32 + lane;
                    v[b] = cc < p.n_ctas ? ld_flag(row + cc) : 0.0;
                }
#pragma unroll
                for (int b = 0; b < 8; ++b) {
                    const int cc = cb + b * 32 + lane;
                    if (cc < p.n_ctas && unset(v[b])) v[b] = poll_value(row + cc, p.status, aborted);
                    sum += v[b];
                }
            }
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, off);
            if (lane == 0) {
                st_flag(p.v_ring + (long long)(n - 1) * O + red_o, sum);
                add_hint(p.hint);
                p.pred[red_cell * p.ldp + p.col0 + n] = sum;
            }
        }
    }
#ifdef XESN_RECUR_TRACE
    if (tid == 0 && c < 64)
        for (int i = 0; i < 8; ++i) g_recur_trace[c * 8 + i] = tr_sum[i];
#endif
    if (owner) p.r[unit] = r_cur;
    if (c == 0)
        for (int e = tid; e < p.n_cells; e += PP_THREADS) p.field[e] = s_field[e];
}

}  // namespace

size_t predict_pull_smem(int I, int O, int upc, long long n_cells, int n_ctas) {
    const int ldw = I + 1 + (I & 1);
    (void)n_ctas;
    return sizeof(double) * ((size_t)upc * ldw + (size_t)O * upc + (size_t)n_cells + I + upc);
}

// geometry of the pull forecast; returns false when the configuration does not fit (caller falls back)
bool predict_pull_geometry(int N, int I, int O, long long n_cells, int num_sms, int* upc_out, int* n_ctas_out) {
    const int upc = ((N + num_sms - 1) / num_sms + 7) & ~7;
    if (upc > PP_UNITS) return false;
    const int n_ctas = (N + upc - 1) / upc;
    if (O > (PP_WARPS - PP_UNITS / 32) * n_ctas) return false;  // one reducer warp per output
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
