// Reservoir recurrence kernels (sm_100a).
//
//   r_{t+1} = (1 - a) r_t + a tanh(W r_t + Win u_t + b)          reference xesn/esn.py:458-460
//
// K1  recur_forced_kernel : teacher-forced run over a chunk of time steps (training states,
//     spin-up).  PERSISTENT: one thread-block cluster owns one reservoir for the whole chunk;
//     every thread keeps its unit's state in a register and its CSR row of W in shared memory;
//     the per-step all-gather of r_t goes through L2 (the state ring buffer itself is the
//     exchange medium, read back with ld.global.cg) and the per-step synchronisation is the
//     hardware cluster barrier (barrier.cluster release/acquire) instead of a kernel launch.
//     The input projection Win u_t + b of the whole chunk is precomputed by one DMMA GEMM
//     (gemm_f64.cu), so the serial critical path per step is: barrier -> <=16 independent L2
//     gathers -> FP64 tanh -> one coalesced store.
// K4a step_update_kernel  : one step for G reservoirs with the input either projected inline
//     (eager predict: u = previous readout) or taken from a precomputed drive (LazyESN).
// K4b readout_kernel      : v_g = Wout_g r_g (reference xesn/esn.py:268, lazyesn.py:325-327),
//     HBM-streaming batched GEMV, one warp per output row, scattered straight into the field.
// K5  gather kernels      : halo construction (dask overlap, lazyesn.py:166,200,235) as an index
//     gather with NaN->0 masking (lazyesn.py:287-289,362-364).
#include <cstdlib>

#include "common.cuh"
#include "kernels.cuh"

namespace xesn {
namespace {

constexpr int RC_THREADS = 512;
}  // namespace
}  // namespace xesn
#include "recur_body.cuh"
namespace xesn {
namespace {
using namespace recur;

template <int UPT, bool WS>
__global__ void __launch_bounds__(RC_THREADS, 1) recur_forced_kernel(RecurArgs p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const unsigned C = cluster_nctarank();
    recur_body<UPT, RC_THREADS, WS>(p, cluster_ctarank(), C, blockIdx.x / C, smem_raw);
}

// experimental geometry: 256 threads x 4 units with twice the registers per thread
constexpr int RC_THREADS_NARROW = 256;
template <int UPT, bool WS>
__global__ void __launch_bounds__(RC_THREADS_NARROW, 1) recur_forced_narrow_kernel(RecurArgs p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const unsigned C = cluster_nctarank();
    recur_body<UPT, RC_THREADS_NARROW, WS>(p, cluster_ctarank(), C, blockIdx.x / C, smem_raw);
}

// ------------------------------------------------------------------------------------------
// one closed-loop / per-step update for G reservoirs:  r_out = leaky(W r_in + drive, r_in)
// where drive[g][i] is either precomputed (Win u + b) or computed inline from u[g][:]
// ------------------------------------------------------------------------------------------
constexpr int SU_THREADS = 256;

__global__ void __launch_bounds__(SU_THREADS) step_update_kernel(StepArgs p) {
    extern __shared__ __align__(16) double s_u[];
    const int g = blockIdx.y;
    const int i = blockIdx.x * SU_THREADS + threadIdx.x;
    const double* r_in = p.r_in + (long long)g * p.N;
    if (!p.drive) {
        for (int k = threadIdx.x; k < p.I; k += SU_THREADS) {
            double x;
            if (p.in_map) {
                int m = p.in_map[(long long)g * p.I + k];
                x = (m >= 0) ? ld_cg(p.field + m) : p.fill[-1 - m];
            } else {
                x = p.u[(long long)g * p.I + k];
            }
            s_u[k] = (p.mask_nan && x != x) ? 0.0 : x;  // LazyESN masks NaN inputs (lazyesn.py:362-364); eager ESN propagates them
        }
        __syncthreads();
    }
    if (i >= p.N) return;
    const int lo = p.indptr[i], len = p.indptr[i + 1] - lo;
    const int pg = g % max(p.param_groups, 1);
    double wr = sparse_row_dot(p.indices + lo, p.values + (long long)pg * p.values_gs + lo, len, r_in);
    double pre;
    if (p.drive) {
        pre = __dadd_rn(wr, p.drive[(long long)g * p.N + i]);
    } else {
        const double* wrow = p.Win + (long long)pg * p.win_gs + (long long)i * p.I;
        double acc = 0.0;
        if ((p.I & 1) == 0 && ((reinterpret_cast<uintptr_t>(p.Win) & 15) == 0)) {
            const double2* w2 = reinterpret_cast<const double2*>(wrow);
            const double2* u2 = reinterpret_cast<const double2*>(s_u);
            double acc1 = 0.0;
#pragma unroll 10
            for (int k = 0; k < (p.I >> 1); ++k) {
                const double2 a = __ldg(w2 + k);
                const double2 b = u2[k];
                acc = fma(a.x, b.x, acc);
                acc1 = fma(a.y, b.y, acc1);
            }
            acc += acc1;
        } else {
#pragma unroll 4
            for (int k = 0; k < p.I; ++k) acc = fma(__ldg(wrow + k), s_u[k], acc);
        }
        pre = __dadd_rn(__dadd_rn(wr, acc), p.bias[(long long)pg * p.bias_gs + i]);
    }
    double r_old = ld_cg(r_in + i);
    const double alpha = p.alpha_g ? p.alpha_g[pg] : p.alpha;
    p.r_out[(long long)g * p.N + i] = leaky(pre, r_old, alpha, 1.0 - alpha);
}

// ------------------------------------------------------------------------------------------
// batched readout GEMV: v[g][o] = Wout[g][o][:] . r[g][:]
// One warp per (output row, n-slice).  When the batch is too small to fill the GPU the reservoir
// axis is split into S slices; the last CTA to finish a row block sums the S partials in slice
// order (deterministic) -- "last block reduces", no second launch.
// ------------------------------------------------------------------------------------------
constexpr int RO_THREADS = 256;
constexpr int RO_WARPS = RO_THREADS / 32;

__global__ void __launch_bounds__(RO_THREADS) readout_kernel(ReadoutArgs p) {
    __shared__ int s_last;
    const int g = blockIdx.z;
    const int rb = blockIdx.y;
    const int sl = blockIdx.x;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int o = rb * RO_WARPS + warp;
    const int n_lo = sl * p.slice, n_hi = min(p.N, n_lo + p.slice);
    double acc = 0.0;
    if (o < p.O) {
        const double* w = p.Wout + ((long long)g * p.O + o) * p.N;
        const double* r = p.r + (long long)g * p.N;
        double acc0 = 0.0, acc1 = 0.0;
        int n = n_lo;
        const bool vec = ((((uintptr_t)(w + n_lo)) & 15) == 0) && ((((uintptr_t)(r + n_lo)) & 15) == 0);
        if (vec) {
            const double2* w2 = reinterpret_cast<const double2*>(w + n_lo);
            const double2* r2 = reinterpret_cast<const double2*>(r + n_lo);
            const int n2 = (n_hi - n_lo) >> 1;
            for (int k = lane; k < n2; k += 32) {
                double2 a = __ldcs(w2 + k);  // streamed once: evict-first
                double2 b = __ldcg(r2 + k);
                acc0 = fma(a.x, b.x, acc0);
                acc1 = fma(a.y, b.y, acc1);
            }
            n = n_lo + (n2 << 1);
        }
        for (int k = n + lane; k < n_hi; k += 32) acc0 = fma(w[k], __ldcg(r + k), acc0);
        acc = acc0 + acc1;
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, off);
    }
    if (p.n_slices == 1) {
        if (o < p.O && lane == 0) {
            long long dst = p.out_map ? (long long)p.out_map[(long long)g * p.O + o] : (long long)g * p.O + o;
            p.v[dst * p.v_stride] = acc;
            if (p.v_copy) p.v_copy[dst * p.v_copy_stride] = acc;
        }
        return;
    }
    if (o < p.O && lane == 0) p.partial[((long long)g * p.n_slices + sl) * p.O + o] = acc;
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) {
        int* cnt = p.counters + (long long)g * gridDim.y + rb;
        int prev = atomicAdd(cnt, 1);
        s_last = (prev == p.n_slices - 1);
        if (s_last) *cnt = 0;  // re-arm for the next launch
    }
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    {
        // warp w finishes row rb*8+w; lanes fetch different slices in parallel (fixed order)
        double sum = 0.0;
        if (o < p.O)
            for (int k = lane; k < p.n_slices; k += 32) sum += __ldcg(p.partial + ((long long)g * p.n_slices + k) * p.O + o);
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, off);
        if (o < p.O && lane == 0) {
            long long dst = p.out_map ? (long long)p.out_map[(long long)g * p.O + o] : (long long)g * p.O + o;
            p.v[dst * p.v_stride] = sum;
            if (p.v_copy) p.v_copy[dst * p.v_copy_stride] = sum;
        }
    }
}

// ------------------------------------------------------------------------------------------
// halo gathers
// ------------------------------------------------------------------------------------------
// out[g][i] = src[map[g][i] * src_stride]  (map < 0 -> fill[-1-map]); NaN -> 0 when mask_nan
__global__ void gather_vec_kernel(const double* __restrict__ src, long long src_stride, const int* __restrict__ map,
                                  const double* __restrict__ fill, double* __restrict__ out, long long n,
                                  int mask_nan) {
    long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    int m = map[k];
    double x = (m >= 0) ? src[(long long)m * src_stride] : fill[-1 - m];
    if (mask_nan && x != x) x = 0.0;
    out[k] = x;
}

// out[row][t] = src[map[row] * ld_src + t0 + t], t < nt  (time-contiguous rows).
// zero_row[row] != 0 -> whole row is written as 0 (masked cell: NaN at the first time step).
__global__ void gather_series_kernel(const double* __restrict__ src, long long ld_src, long long t0,
                                     const int* __restrict__ map, const double* __restrict__ fill,
                                     const unsigned char* __restrict__ zero_row, double* __restrict__ out,
                                     long long ld_out, int nt) {
    const long long row = blockIdx.y;
    const int m = map[row];
    const bool zero = zero_row && zero_row[row];
    const double* s = (m >= 0) ? src + (long long)m * ld_src + t0 : nullptr;
    const double f = (m >= 0) ? 0.0 : fill[-1 - m];
    double* o = out + row * ld_out;
    for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < nt; t += gridDim.x * blockDim.x) {
        double x = zero ? 0.0 : (s ? s[t] : f);
        o[t] = x;
    }
}

}  // namespace

// ------------------------------------------------------------------------------------------
// launchers
// ------------------------------------------------------------------------------------------
static int recur_geometry(int N, int* C_out, int* upt_out) {
    int need = (N + RC_THREADS - 1) / RC_THREADS;
    int C = 1;
    while (C < need && C < 16) C <<= 1;
    int per = (N + C - 1) / C;
    int upt = (per + RC_THREADS - 1) / RC_THREADS;
    *C_out = C, *upt_out = upt;
    if (upt == 3) upt = 4;
    *upt_out = upt;
    return upt <= 4 ? 0 : -2;
}

int recur_cluster_size(int N) {
    int C, upt;
    recur_geometry(N, &C, &upt);
    return C;
}

int launch_recur_forced(RecurArgs a, int n_groups, int max_slice_nnz, cudaStream_t stream) {
    if (a.steps <= 0 || n_groups <= 0) return 0;
    int C, upt;
    if (recur_geometry(a.N, &C, &upt)) {
        set_error("recur_forced: n_reservoir=%d exceeds the persistent kernel capacity (%d)", a.N, 16 * RC_THREADS * 4);
        return -2;
    }
    a.units_per_cta = recur_units_per_cta(a.N, C);
    a.smem_nnz = max_slice_nnz + (max_slice_nnz & 1);
    a.n_pad = ((a.N + 1) & ~1) + 2;
    // shared memory: full state copy (always) + this CTA's CSR slice of W (if it fits)
    const long long smem_r = (long long)a.n_pad * 8;
    const long long smem_w = (long long)a.smem_nnz * 12 + 16;
    const long long limit = 227 * 1024;
    if (smem_r > limit) {
        set_error("recur_forced: n_reservoir=%d does not fit the shared-memory state copy", a.N);
        return -2;
    }
    a.w_in_smem = (smem_r + smem_w <= limit);
    const size_t smem = (size_t)(smem_r + (a.w_in_smem ? smem_w : 0));

    // every row of the ring the kernel will write/poll starts as the "unset" pattern
    XESN_CUDA_OK(cudaMemset2DAsync(a.states, sizeof(double) * a.state_group_stride, 0xFF,
                                   sizeof(double) * (size_t)(a.steps + 1) * a.N, n_groups, stream));

    void (*kern)(RecurArgs);
    if (a.w_in_smem)
        kern = upt == 1 ? recur_forced_kernel<1, true> : (upt == 2 ? recur_forced_kernel<2, true> : recur_forced_kernel<4, true>);
    else
        kern = upt == 1 ? recur_forced_kernel<1, false> : (upt == 2 ? recur_forced_kernel<2, false> : recur_forced_kernel<4, false>);
    int threads = RC_THREADS;
    if (getenv("XESN_B200_REC_NARROW") && upt == 2) {
        kern = a.w_in_smem ? recur_forced_narrow_kernel<4, true> : recur_forced_narrow_kernel<4, false>;
        threads = RC_THREADS_NARROW;
    }
    XESN_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)limit));
    if (C > 8) XESN_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));

    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)(C * n_groups));
    cfg.blockDim = dim3((unsigned)threads);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = C;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    XESN_CUDA_OK(cudaLaunchKernelEx(&cfg, kern, a));
    count_launch();
    return 0;
}

int launch_step_update(const StepArgs& a, int n_groups, cudaStream_t stream) {
    dim3 grid((a.N + SU_THREADS - 1) / SU_THREADS, n_groups);
    size_t smem = a.drive ? 0 : (size_t)a.I * sizeof(double);
    if (smem > 48 * 1024) {
        // per device, so set on every such launch (not once per process)
        XESN_CUDA_OK(cudaFuncSetAttribute(step_update_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        if (smem > 200 * 1024) {
            set_error("step_update: n_input too large for the inline projection");
            return -2;
        }
    }
    step_update_kernel<<<grid, SU_THREADS, smem, stream>>>(a);
    XESN_CUDA_OK(cudaGetLastError());
    count_launch();
    return 0;
}

int readout_slices(int N, int O, int n_groups) {
    const long long row_blocks = (long long)((O + RO_WARPS - 1) / RO_WARPS) * n_groups;
    long long s = (2 * 148 + row_blocks - 1) / row_blocks;
    const long long max_s = (N + 255) / 256;
    if (s > max_s) s = max_s;
    return (int)(s < 1 ? 1 : s);
}

int launch_readout(ReadoutArgs a, int n_groups, cudaStream_t stream) {
    if (a.n_slices < 1 || !a.partial || !a.counters) a.n_slices = 1;
    a.slice = (a.N + a.n_slices - 1) / a.n_slices;
    a.slice += a.slice & 1;  // even, keeps double2 alignment of the slices
    dim3 grid(a.n_slices, (a.O + RO_WARPS - 1) / RO_WARPS, n_groups);
    readout_kernel<<<grid, RO_THREADS, 0, stream>>>(a);
    XESN_CUDA_OK(cudaGetLastError());
    count_launch();
    return 0;
}

int launch_gather_vec(const double* src, long long src_stride, const int* map, const double* fill, double* out,
                      long long n, int mask_nan, cudaStream_t stream) {
    if (n <= 0) return 0;
    gather_vec_kernel<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(src, src_stride, map, fill, out, n, mask_nan);
    XESN_CUDA_OK(cudaGetLastError());
    count_launch();
    return 0;
}

int launch_gather_series(const double* src, long long ld_src, long long t0, const int* map, const double* fill,
                         const unsigned char* zero_row, double* out, long long ld_out, long long rows, int nt,
                         cudaStream_t stream) {
    if (rows <= 0 || nt <= 0) return 0;
    // grid.y is limited to 65535: split rows over several launches
    for (long long r0 = 0; r0 < rows; r0 += 65535) {
        long long nr = rows - r0 < 65535 ? rows - r0 : 65535;
        dim3 grid((nt + 255) / 256 < 8 ? (nt + 255) / 256 : 8, (unsigned)nr);
        gather_series_kernel<<<grid, 256, 0, stream>>>(src, ld_src, t0, map + r0, fill, zero_row ? zero_row + r0 : nullptr,
                                                       out + r0 * ld_out, ld_out, nt);
        count_launch();
    }
    XESN_CUDA_OK(cudaGetLastError());
    return 0;
}

}  // namespace xesn

#ifdef XESN_RECUR_TRACE
extern "C" int xesn_debug_trace(long long* out) {
    return (int)cudaMemcpyFromSymbol(out, xesn::recur::g_recur_trace, sizeof(long long) * 64 * 8);
}
#endif
