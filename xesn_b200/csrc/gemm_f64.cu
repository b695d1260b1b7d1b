// FP64 GEMM / SYRK on the tensor pipe (DMMA m8n8k4) for sm_100a.
//
// This is the one dense contraction engine of the ESN hot path.  It replaces the numpy
// `@` / cuBLAS dgemm calls of the reference at
//   * xesn/esn.py:510  `rbar += rT.T @ rT`        (Gram, symmetric: lower tiles only)
//   * xesn/esn.py:509  `ybar += y[:, i0:i1] @ rT`
//   * xesn/esn.py:459  `Win @ u` for a whole chunk of time steps (the "drive")
// and is re-used by the blocked Cholesky / triangular solves that replace
// `scipy.linalg.solve(..., assume_a="sym")` (xesn/esn.py:516-520).
//
// tcgen05 has no FP64 kind, so FP64 tensor work on Blackwell is `mma.sync ... f64`
// (SASS: DMMA.8x8x4).  Operand tiles are staged with cp.async (LDGSTS) through a 4-stage
// shared-memory ring; each operand may be K-contiguous or MN-contiguous in global memory,
// so the time-last (n_vars, n_time) arrays of the reference are consumed without a transpose.
//
// CTA tile 128x128x16, 256 threads = 8 warps (2 x 4), warp tile 64x32 -> 32 DMMA per k4 step
// against 12 LDS.64 (shared memory padded so fragment loads are bank-conflict free).
#include <algorithm>

#include "common.cuh"
#include "gemm_tile.cuh"
#include "yr_skinny.cuh"

namespace xesn {
namespace {

using namespace gemm;

template <bool A_K, bool B_K, bool VEC>
__global__ void __launch_bounds__(NTHREADS, 1) gemm_f64_kernel(GemmArgs p, int tiles_n) {
    extern __shared__ __align__(16) double smem[];
    int ti, tj;
    if (p.flags & GEMM_LOWER) {
        lower_tile(blockIdx.x, ti, tj);
    } else {
        ti = blockIdx.x / tiles_n;
        tj = blockIdx.x % tiles_n;
    }
    gemm_tile<A_K, B_K, VEC>(p, ti, tj, blockIdx.y, smem);
}

inline bool aligned16(const void* p) { return ((uintptr_t)p & 15) == 0; }

constexpr int YR_THREADS = 256;
__global__ void __launch_bounds__(YR_THREADS) yr_skinny_kernel(GemmArgs p) {
    extern __shared__ __align__(16) double smem[];
    const long long items = yr::n_items(p);
    for (long long it = blockIdx.x; it < items; it += gridDim.x) yr::yr_item<YR_THREADS>(p, it, smem);
}

}  // namespace

// C += alpha * A B for a skinny A (few target rows, time-contiguous) and B = state ring rows; see yr_skinny.cuh
int launch_yr_skinny(const GemmArgs& a, cudaStream_t stream) {
    if (a.M <= 0 || a.N <= 0 || a.K <= 0 || a.batch <= 0) return 0;
    const int smem = yr::smem_bytes<YR_THREADS>();
    XESN_CUDA_OK(cudaFuncSetAttribute(yr_skinny_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    const long long items = yr::n_items(a);
    yr_skinny_kernel<<<(unsigned)std::min<long long>(items, 4 * 148), YR_THREADS, smem, stream>>>(a);
    count_launch();
    XESN_CUDA_OK(cudaGetLastError());
    return 0;
}

int launch_gemm_f64(const GemmArgs& a, cudaStream_t stream) {
    if (a.M <= 0 || a.N <= 0 || a.batch <= 0) return 0;
    const int tiles_m = (a.M + BM - 1) / BM, tiles_n = (a.N + BN - 1) / BN;
    long long ntiles;
    if (a.flags & GEMM_LOWER) {
        if (a.M != a.N) {
            set_error("gemm: GEMM_LOWER needs a square output");
            return -1;
        }
        ntiles = (long long)tiles_m * (tiles_m + 1) / 2;
    } else {
        ntiles = (long long)tiles_m * tiles_n;
    }
    const bool vec = aligned16(a.A) && (a.lda % 2 == 0) && (a.strideA % 2 == 0) && aligned16(a.B) &&
                     (a.ldb % 2 == 0) && (a.strideB % 2 == 0);
    dim3 grid((unsigned)ntiles, (unsigned)a.batch);
    const bool ak = a.flags & GEMM_A_KMAJOR, bk = a.flags & GEMM_B_KMAJOR;
    void (*table[8])(GemmArgs, int) = {
        gemm_f64_kernel<false, false, false>, gemm_f64_kernel<false, false, true>,
        gemm_f64_kernel<false, true, false>,  gemm_f64_kernel<false, true, true>,
        gemm_f64_kernel<true, false, false>,  gemm_f64_kernel<true, false, true>,
        gemm_f64_kernel<true, true, false>,   gemm_f64_kernel<true, true, true>,
    };
    const int which = (ak ? 4 : 0) + (bk ? 2 : 0) + (vec ? 1 : 0);
    void (*kern)(GemmArgs, int) = table[which];
    // per device, so on every launch rather than once per process (two handles may live on two GPUs)
    XESN_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
    kern<<<grid, NTHREADS, SMEM_BYTES, stream>>>(a, tiles_n);
    count_launch();
    XESN_CUDA_OK(cudaGetLastError());
    return 0;
}

}  // namespace xesn
