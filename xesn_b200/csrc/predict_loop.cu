// Persistent closed-loop forecast kernel (sm_100a).
//
//   for n = 1 .. n_steps:   r <- update(r, gather(field));   field <- scatter(Wout r)
//   (reference xesn/esn.py:264-268 and xesn/lazyesn.py:226-236: update, readout, re-halo per step)
//
// The per-step launches of the first version (update kernel + readout kernel, 18 us per step at
// N=16000) spend their time in launch gaps and cold caches.  Here ONE cooperative kernel runs the
// whole forecast: every CTA keeps the same work items in every step, so its rows of W, Win and
// Wout stay hot in its L1/L2, and the two phase changes per step are software grid barriers
// (one atomic per CTA).  Input projection is inline (the kernel is used when G*N*I is small; big
// LazyESN configurations are bandwidth-bound and keep the GEMM-based per-step path).
//
// Multi-GPU (one process per GPU): the readout epilogue stores each predicted cell into the field
// buffer of EVERY rank through peer-mapped pointers (NVLink P2P stores), and the second grid barrier
// of the step is extended by a cross-GPU handshake on peer-mapped flags -- the per-step neighbour
// exchange that dask's `overlap` performs in the reference (lazyesn.py:235), without leaving the kernel.
#include <algorithm>

#include "common.cuh"
#include "kernels.cuh"
#include "recur_body.cuh"

namespace xesn {
namespace {

using namespace recur;

constexpr int PL_THREADS = 256;
constexpr int PL_WARPS = PL_THREADS / 32;

__device__ __forceinline__ unsigned ld_acquire_u32(const unsigned* p) {
    unsigned v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];\n" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_relaxed_sys_u32(unsigned* p, unsigned v) {
    asm volatile("st.relaxed.sys.global.u32 [%0], %1;\n" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned ld_relaxed_sys_u32(const unsigned* p) {
    unsigned v;
    asm volatile("ld.relaxed.sys.global.u32 %0, [%1];\n" : "=r"(v) : "l"(p) : "memory");
    return v;
}

// Software grid barrier (all CTAs are co-resident: cooperative launch).  bar[0] = arrival counter,
// bar[1] = generation.  If `cross` is set the last-arriving CTA also performs the cross-GPU handshake
// before it releases the local generation.
__device__ __forceinline__ void grid_barrier(const PredictLoopArgs& p, unsigned& gen, bool cross, unsigned step) {
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned target = gen + 1;
        __threadfence();
        const unsigned prev = atomicAdd(&p.bar[0], 1u);
        if (prev == gridDim.x - 1) {
            if (cross && p.world > 1) {
                // every local CTA has issued its peer stores: publish "rank `rank` finished step" to all ranks,
                // then wait until every rank has published it
                // ONE system-scope fence releases all peer stores of this rank; the flags themselves go out relaxed
                // (a st.release.sys per peer is a fence per peer: measured 16 us per step at 8 ranks)
                __threadfence_system();
                for (int k = 0; k < p.world; ++k) st_relaxed_sys_u32(p.peer_flags[k] + p.rank, step);
                for (int k = 0; k < p.world; ++k) {
                    int spins = 0;
                    while (ld_relaxed_sys_u32(p.peer_flags[p.rank] + k) < step)
                        if (poll_expired(spins, p.status)) break;
                }
                __threadfence_system();  // acquire side: the peers' field stores are visible before the local release
            }
            p.bar[0] = 0;
            __threadfence();
            atomicExch(&p.bar[1], target);
        } else {
            int spins = 0;
            while (ld_acquire_u32(&p.bar[1]) != target)
                if (poll_expired(spins, p.status)) break;
        }
        gen = target;
    }
    __syncthreads();
}

__global__ void __launch_bounds__(PL_THREADS, 1) predict_loop_kernel(PredictLoopArgs p) {
    extern __shared__ __align__(16) double s_u[];  // [I] input vector of the current item
    __shared__ int s_last;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int nblk = (p.N + PL_THREADS - 1) / PL_THREADS;
    const int row_blocks = (p.O + PL_WARPS - 1) / PL_WARPS;
    const long long items_u = (long long)p.G * nblk;
    const long long items_r = (long long)p.G * row_blocks * p.n_slices;
    unsigned gen = 0;
    if (tid == 0) gen = ld_acquire_u32(&p.bar[1]);

    for (int n = 1; n <= p.n_steps; ++n) {
        const double* field_cur = p.field[(n - 1) & 1];
        double* field_nxt = p.field[n & 1];
        const double* r_cur = p.r[(n - 1) & 1];
        double* r_nxt = p.r[n & 1];

        // ---- phase U: r_nxt = leaky(W r_cur + Win u + b, r_cur) ----
        for (long long it = blockIdx.x; it < items_u; it += gridDim.x) {
            const int g = (int)(it / nblk), ib = (int)(it % nblk);
            __syncthreads();  // previous item's readers of s_u are done
            for (int k = tid; k < p.I; k += PL_THREADS) {
                const int m = p.in_map[(long long)g * p.I + k];
                double x = (m >= 0) ? __ldcg(field_cur + m) : p.fill[-1 - m];
                s_u[k] = (p.mask_nan && x != x) ? 0.0 : x;  // LazyESN masks NaN inputs (lazyesn.py:362-364)
            }
            __syncthreads();
            const int i = ib * PL_THREADS + tid;
            if (i < p.N) {
                const double* rg = r_cur + (long long)g * p.N;
                const int lo = __ldg(p.indptr + i), len = __ldg(p.indptr + i + 1) - lo;
                const int pg = g % max(p.param_groups, 1);
                const double wr = sparse_row_dot(p.indices + lo, p.values + (long long)pg * p.values_gs + lo, len, rg);
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
                const double pre = __dadd_rn(__dadd_rn(wr, acc), __ldg(p.bias + (long long)pg * p.bias_gs + i));
                const double alpha = p.alpha_g ? p.alpha_g[pg] : p.alpha;
                r_nxt[(long long)g * p.N + i] = leaky(pre, ld_cg(rg + i), alpha, 1.0 - alpha);
            }
        }
        grid_barrier(p, gen, false, 0);

        // ---- phase R: v = Wout r_nxt, scattered into the field (all ranks) and the trajectory ----
        for (long long it = blockIdx.x; it < items_r; it += gridDim.x) {
            const int sl = (int)(it % p.n_slices);
            const int rb = (int)((it / p.n_slices) % row_blocks);
            const int g = (int)(it / ((long long)p.n_slices * row_blocks));
            const int o = rb * PL_WARPS + warp;
            const int n_lo = sl * p.slice, n_hi = min(p.N, n_lo + p.slice);
            double acc = 0.0;
            if (o < p.O) {
                const double* w = p.Wout + ((long long)g * p.O + o) * p.N;
                const double* r = r_nxt + (long long)g * p.N;
                double acc0 = 0.0, acc1 = 0.0;
                int k0 = n_lo;
                if (((reinterpret_cast<uintptr_t>(w + n_lo) & 15) == 0) && ((reinterpret_cast<uintptr_t>(r + n_lo) & 15) == 0)) {
                    const double2* w2 = reinterpret_cast<const double2*>(w + n_lo);
                    const double2* r2 = reinterpret_cast<const double2*>(r + n_lo);
                    const int n2 = (n_hi - n_lo) >> 1;
                    for (int k = lane; k < n2; k += 32) {
                        const double2 a = __ldg(w2 + k);  // same rows every step: keep them cached
                        const double2 b = __ldcg(r2 + k);
                        acc0 = fma(a.x, b.x, acc0);
                        acc1 = fma(a.y, b.y, acc1);
                    }
                    k0 = n_lo + (n2 << 1);
                }
                for (int k = k0 + lane; k < n_hi; k += 32) acc0 = fma(__ldg(w + k), __ldcg(r + k), acc0);
                acc = acc0 + acc1;
#pragma unroll
                for (int off = 16; off > 0; off >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, off);
            }
            bool finish = (p.n_slices == 1) && lane == 0;
            double v = acc;
            if (p.n_slices > 1) {
                if (o < p.O && lane == 0) p.partial[((long long)g * p.n_slices + sl) * p.O + o] = acc;
                __threadfence();
                __syncthreads();
                if (tid == 0) {
                    int* cnt = p.counters + (long long)g * row_blocks + rb;
                    const int prev = atomicAdd(cnt, 1);
                    s_last = (prev == p.n_slices - 1);
                    if (s_last) *cnt = 0;
                }
                __syncthreads();
                if (s_last) {
                    // the last CTA to arrive combines the slices: warp w finishes row rb*8+w, the lanes
                    // fetch different slices in parallel (fixed order -> deterministic)
                    __threadfence();
                    double sum = 0.0;
                    if (o < p.O)
                        for (int k = lane; k < p.n_slices; k += 32)
                            sum += __ldcg(p.partial + ((long long)g * p.n_slices + k) * p.O + o);
#pragma unroll
                    for (int off = 16; off > 0; off >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, off);
                    v = sum;
                    finish = (lane == 0);
                }
            }
            if (finish && o < p.O) {
                const long long cell = p.out_map ? (long long)p.out_map[(long long)g * p.O + o] : (long long)g * p.O + o;
                field_nxt[cell] = v;
                p.pred[cell * p.ldp + n] = v;
                for (int k = 0; k < p.world; ++k)
                    if (k != p.rank) p.peer_field[(n & 1) * p.world + k][cell] = v;  // NVLink P2P store
            }
        }
        grid_barrier(p, gen, true, (unsigned)(p.step_base + n));
    }
}

}  // namespace

int launch_predict_loop(const PredictLoopArgs& a, int num_sms, cudaStream_t stream) {
    if (a.n_steps <= 0) return 0;
    const int nblk = (a.N + PL_THREADS - 1) / PL_THREADS;
    const int row_blocks = (a.O + PL_WARPS - 1) / PL_WARPS;
    long long items = std::max<long long>((long long)a.G * nblk, (long long)a.G * row_blocks * a.n_slices);
    int grid = (int)std::min<long long>(items, num_sms);
    size_t smem = (size_t)((a.I + 1) & ~1) * sizeof(double);
    if (smem > 200 * 1024) {
        set_error("predict_loop: n_input too large");
        return -2;
    }
    XESN_CUDA_OK(cudaFuncSetAttribute(predict_loop_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    PredictLoopArgs args = a;
    void* kargs[] = {&args};
    XESN_CUDA_OK(cudaLaunchCooperativeKernel((void*)predict_loop_kernel, dim3(grid), dim3(PL_THREADS), kargs, smem, stream));
    count_launch();
    return 0;
}

}  // namespace xesn
