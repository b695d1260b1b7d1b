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
#include "common.cuh"
#include "kernels.cuh"

namespace xesn {
namespace {

constexpr int RC_THREADS = 512;
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
constexpr int SPIN_LIMIT = 1 << 22;  // ~1-2 s of polling, then the run is flagged as failed
constexpr int POLL_BATCH = 16;       // 16-byte pieces in flight per thread while refreshing the copy

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
__device__ __forceinline__ void st_flag(double* p, double v) {
    if (__double_as_longlong(v) == UNSET_BITS) v = __longlong_as_double(0x7FF8000000000000LL);
    asm volatile("st.relaxed.gpu.global.f64 [%0], %1;\n" ::"l"(p), "d"(v) : "memory");
}
__device__ __forceinline__ bool unset(double v) { return __double_as_longlong(v) == UNSET_BITS; }

// bounded spin bookkeeping: returns true once the run has been flagged as failed
__device__ __forceinline__ bool poll_expired(int& spins, int* status) {
    if ((++spins & 255) != 0) return false;
    if (spins >= SPIN_LIMIT) atomicExch(status, 1);
    return *reinterpret_cast<volatile int*>(status) != 0;
}

template <int UPT>
__global__ void __launch_bounds__(RC_THREADS, 1) recur_forced_kernel(RecurArgs p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const unsigned C = cluster_nctarank();
    const unsigned c = cluster_ctarank();
    const int g = blockIdx.x / C;
    const int tid = threadIdx.x;
    const int N = p.N;
    const int upc = p.units_per_cta;
    const int u0 = c * upc;
    const int u1 = min(N, u0 + upc);
    if (u0 >= N) return;  // nothing to produce, nobody waits for this CTA

    double* s_r = reinterpret_cast<double*>(smem_raw);  // [n_pad] copy of r_t
    // ---- stage this CTA's CSR slice of W in shared memory (or fall back to global) ----
    const int slice_lo = p.indptr[u0], slice_hi = p.indptr[u1];
    const int* idx_base = p.indices;
    const double* val_base = p.values;
    if (p.w_in_smem) {
        double* s_val = s_r + p.n_pad;
        int* s_idx = reinterpret_cast<int*>(s_val + p.smem_nnz);
        for (int k = slice_lo + tid; k < slice_hi; k += RC_THREADS) {
            s_val[k - slice_lo] = p.values[k];
            s_idx[k - slice_lo] = p.indices[k];
        }
        idx_base = s_idx - slice_lo;
        val_base = s_val - slice_lo;
    }
    double* states = p.states + (long long)g * p.state_group_stride;
    const double* drive = p.drive + (long long)g * p.drive_group_stride;
    const double* carry = p.carry + (long long)g * N;
    for (int i = tid; i < N; i += RC_THREADS) s_r[i] = carry[i];

    int unit[UPT], row_lo[UPT], row_len[UPT];
    double r_cur[UPT], d_next[UPT];
#pragma unroll
    for (int j = 0; j < UPT; ++j) {
        unit[j] = u0 + tid + j * RC_THREADS;
        const bool ok = unit[j] < u1;
        row_lo[j] = ok ? p.indptr[unit[j]] : 0;
        row_len[j] = ok ? p.indptr[unit[j] + 1] - row_lo[j] : 0;
        r_cur[j] = ok ? carry[unit[j]] : 0.0;
        d_next[j] = (ok && p.steps > 0) ? drive[unit[j]] : 0.0;
        if (ok) st_flag(states + unit[j], r_cur[j]);  // row 0 = state entering the chunk (read by the Gram kernel)
        else unit[j] = -1;
    }
    const double alpha = p.alpha, oma = 1.0 - p.alpha;
    const bool vec2 = (N % 2 == 0) && ((reinterpret_cast<uintptr_t>(states) & 15) == 0);
    bool aborted = false;
    __syncthreads();

    for (int t = 0; t < p.steps; ++t) {
        double* r_next = states + (long long)(t + 1) * N;
        // ---- my units: W r_t from the shared copy, tanh, leak; publish to the ring ----
#pragma unroll
        for (int j = 0; j < UPT; ++j) {
            if (unit[j] >= 0) {
                const int* ix = idx_base + row_lo[j];
                const double* vl = val_base + row_lo[j];
                double acc = 0.0;
#pragma unroll 4
                for (int k = 0; k < row_len[j]; ++k) acc = __dadd_rn(acc, __dmul_rn(vl[k], s_r[ix[k]]));
                const double pre = __dadd_rn(acc, d_next[j]);
                r_cur[j] = leaky(pre, r_cur[j], alpha, oma);
                st_flag(r_next + unit[j], r_cur[j]);
            }
        }
        if (t + 1 == p.steps) break;
#pragma unroll
        for (int j = 0; j < UPT; ++j)
            if (unit[j] >= 0) d_next[j] = __ldg(drive + (long long)(t + 1) * N + unit[j]);
        __syncthreads();  // every gather of r_t is done: the shared copy may be overwritten
#pragma unroll
        for (int j = 0; j < UPT; ++j)
            if (unit[j] >= 0) s_r[unit[j]] = r_cur[j];
        // ---- refresh the other CTAs' slices of r_{t+1} with coalesced polling loads ----
        if (C > 1) {
            if (vec2) {
                const int n2 = N >> 1;
                for (int q0 = tid; q0 < n2; q0 += RC_THREADS * POLL_BATCH) {
                    double2 v[POLL_BATCH];
                    unsigned pending = 0;  // bit b: piece b still has an unset half
#pragma unroll
                    for (int b = 0; b < POLL_BATCH; ++b) {
                        const int q = q0 + b * RC_THREADS;
                        // pieces lying entirely inside my own slice are already in the copy
                        if ((q < n2) && !(2 * q >= u0 && 2 * q + 1 < u1)) pending |= 1u << b;
                    }
                    const unsigned mine = pending;
                    int spins = 0;
                    while (pending && !aborted) {
#pragma unroll
                        for (int b = 0; b < POLL_BATCH; ++b)
                            if (pending & (1u << b)) v[b] = ld_flag2(r_next + 2 * (q0 + b * RC_THREADS));
#pragma unroll
                        for (int b = 0; b < POLL_BATCH; ++b)
                            if ((pending & (1u << b)) && !unset(v[b].x) && !unset(v[b].y)) pending &= ~(1u << b);
                        if (pending) aborted = poll_expired(spins, p.status);
                    }
#pragma unroll
                    for (int b = 0; b < POLL_BATCH; ++b)
                        if (mine & (1u << b)) *reinterpret_cast<double2*>(s_r + 2 * (q0 + b * RC_THREADS)) = v[b];
                }
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
        __syncthreads();
    }
#pragma unroll
    for (int j = 0; j < UPT; ++j)
        if (unit[j] >= 0) p.carry[(long long)g * N + unit[j]] = r_cur[j];
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
            s_u[k] = (x != x) ? 0.0 : x;  // NaN inputs are masked out (lazyesn.py:362-364)
        }
        __syncthreads();
    }
    if (i >= p.N) return;
    const int lo = p.indptr[i], len = p.indptr[i + 1] - lo;
    double wr = sparse_row_dot(p.indices + lo, p.values + lo, len, r_in);
    double pre;
    if (p.drive) {
        pre = __dadd_rn(wr, p.drive[(long long)g * p.N + i]);
    } else {
        const double* wrow = p.Win + (long long)i * p.I;
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
        pre = __dadd_rn(__dadd_rn(wr, acc), p.bias[i]);
    }
    double r_old = ld_cg(r_in + i);
    p.r_out[(long long)g * p.N + i] = leaky(pre, r_old, p.alpha, 1.0 - p.alpha);
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
    if (threadIdx.x < RO_WARPS) {
        const int oo = rb * RO_WARPS + threadIdx.x;
        if (oo < p.O) {
            double sum = 0.0;
            for (int k = 0; k < p.n_slices; ++k) sum += __ldcg(p.partial + ((long long)g * p.n_slices + k) * p.O + oo);
            long long dst = p.out_map ? (long long)p.out_map[(long long)g * p.O + oo] : (long long)g * p.O + oo;
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
    a.units_per_cta = (a.N + C - 1) / C;
    a.smem_nnz = max_slice_nnz + (max_slice_nnz & 1);
    a.n_pad = (a.N + 1) & ~1;
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

    void (*kern)(RecurArgs) =
        upt == 1 ? recur_forced_kernel<1> : (upt == 2 ? recur_forced_kernel<2> : recur_forced_kernel<4>);
    XESN_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)limit));
    if (C > 8) XESN_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));

    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)(C * n_groups));
    cfg.blockDim = dim3(RC_THREADS);
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
        static bool done = false;
        if (!done) {
            XESN_CUDA_OK(cudaFuncSetAttribute(step_update_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
            done = true;
        }
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
