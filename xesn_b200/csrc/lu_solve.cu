// Pivoted fallback of the ridge solve (sm_100a): blocked LU with partial pivoting.
//
// The reference solves  Wout (R R^T + beta I) = Y R^T  with `scipy.linalg.solve(..., assume_a="sym")`, i.e. LAPACK's
// pivoted LDL^T (dsysv) on the CPU and LU (getrf/getrs) under cupy (xesn/esn.py:516-520).  Both survive systems that
// are numerically indefinite (tiny beta, rank-deficient Gram); a Cholesky factorisation does not.  So when the
// fast path (cholesky.cu) reports a non-positive pivot, that system is solved again here from the re-accumulated
// Gram matrix -- like the reference's own GPU path, by LU with partial pivoting, which is backward stable.
//
// Memory view: A is symmetric and row-major, so memory row j is COLUMN j of the column-major view M = A.  The
// right-hand sides (the rows of Y R^T, one per output) are further columns of the augmented matrix [M | B^T].
//   pivot search down view-column j      = scan along memory row j (contiguous)
//   swap of view-rows j and p            = swap of memory columns j and p in every row (A and B)
//   panel (nb view-columns)              = nb memory rows; rank-1 updates touch only those rows
//   U12 = L11^-1 A12, A22 -= L21 U12     = per memory row of the trailing part: a 32-long forward substitution, then one
//                                          FP64 tensor-pipe GEMM (gemm_f64.cu) with K = nb
// which leaves B = L^-1 P B^T; the backward substitution with U reuses the blocked sweep of the Cholesky path
// (U's diagonal blocks are the lower-triangular memory blocks, inverted explicitly 128 at a time).
// This path is for correctness, not speed: ~3 small launches per column.
#include <algorithm>

#include "common.cuh"
#include "kernels.cuh"

namespace xesn {
namespace {

constexpr int LU_NB = 32;    // panel width
constexpr int LU_INV = 128;  // diagonal-block size of the backward sweep (= cholesky.cu's NB)

// upper triangle <- lower triangle
__global__ void mirror_lower_kernel(double* A, long long lda, int n) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;  // column
    const int i = blockIdx.y;                             // row
    if (j < n && j > i) A[(long long)i * lda + j] = A[(long long)j * lda + i];
}

// piv[j] = argmax_{i >= j} |A[j][i]| (first one on ties); info = j+1 if that maximum is 0 or not finite
__global__ void __launch_bounds__(1024) lu_pivot_kernel(const double* __restrict__ A, long long lda, int n, int j, int* piv,
                                                        int* info) {
    __shared__ double s_val[32];
    __shared__ int s_idx[32];
    const double* row = A + (long long)j * lda;
    double best = -1.0;
    int at = j;
    for (int i = j + threadIdx.x; i < n; i += blockDim.x) {
        const double v = fabs(row[i]);
        if (v > best || !(v == v)) {  // a NaN wins and is reported
            best = (v == v) ? v : 1.0 / 0.0;
            at = i;
            if (!(v == v)) break;
        }
    }
    for (int off = 16; off > 0; off >>= 1) {
        const double ob = __shfl_xor_sync(0xffffffffu, best, off);
        const int oa = __shfl_xor_sync(0xffffffffu, at, off);
        if (ob > best || (ob == best && oa < at)) best = ob, at = oa;
    }
    if ((threadIdx.x & 31) == 0) s_val[threadIdx.x >> 5] = best, s_idx[threadIdx.x >> 5] = at;
    __syncthreads();
    if (threadIdx.x < 32) {
        best = threadIdx.x < (blockDim.x >> 5) ? s_val[threadIdx.x] : -1.0;
        at = threadIdx.x < (blockDim.x >> 5) ? s_idx[threadIdx.x] : j;
        for (int off = 16; off > 0; off >>= 1) {
            const double ob = __shfl_xor_sync(0xffffffffu, best, off);
            const int oa = __shfl_xor_sync(0xffffffffu, at, off);
            if (ob > best || (ob == best && oa < at)) best = ob, at = oa;
        }
        if (threadIdx.x == 0) {
            piv[j] = at;
            if (!(best > 0.0) || !(best < 1.7e308)) atomicCAS(info, 0, j + 1);
        }
    }
}

// swap memory columns j and piv[j] in every row of A (n rows) and of B (nrhs rows)
__global__ void lu_swap_kernel(double* A, long long lda, int n, double* B, long long ldb, int nrhs, int j, const int* piv) {
    const int p = piv[j];
    if (p == j) return;
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n + nrhs) return;
    double* row = r < n ? A + (long long)r * lda : B + (long long)(r - n) * ldb;
    const double a = row[j], b = row[p];
    row[j] = b, row[p] = a;
}

// l_i = A[j][i] / A[j][j] for i > j (stored in place), then the rank-1 update of the rest of the panel:
// A[c][i] -= A[c][j] * l_i for the panel rows j < c < k1.  One thread per column i.
__global__ void lu_scale_update_kernel(double* A, long long lda, int n, int j, int k1) {
    const int i = j + 1 + blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double* rowj = A + (long long)j * lda;
    const double l = rowj[i] / rowj[j];
    rowj[i] = l;
    for (int c = j + 1; c < k1; ++c) {
        double* rowc = A + (long long)c * lda;
        rowc[i] = fma(-rowc[j], l, rowc[i]);
    }
}

// v <- L11^-1 v for v = row[k0 .. k1) of every trailing memory row of A (rows >= k1) and of every row of B;
// L11(a, b) = A[k0 + b][k0 + a] for a > b (unit lower triangular)
__global__ void __launch_bounds__(128) lu_panel_trsm_kernel(double* A, long long lda, int n, double* B, long long ldb,
                                                            int nrhs, int k0, int k1) {
    __shared__ double s_l[LU_NB][LU_NB + 1];  // s_l[b][a] = L11(a, b)
    const int nb = k1 - k0;
    for (int e = threadIdx.x; e < nb * nb; e += blockDim.x) {
        const int b = e / nb, a = e % nb;
        s_l[b][a] = a > b ? A[(long long)(k0 + b) * lda + k0 + a] : 0.0;
    }
    __syncthreads();
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    const int rows = (n - k1) + nrhs;
    if (r >= rows) return;
    double* row = (r < n - k1 ? A + (long long)(k1 + r) * lda : B + (long long)(r - (n - k1)) * ldb) + k0;
    double v[LU_NB];
#pragma unroll
    for (int a = 0; a < LU_NB; ++a) v[a] = a < nb ? row[a] : 0.0;
#pragma unroll
    for (int b = 0; b < LU_NB; ++b)
#pragma unroll
        for (int a = b + 1; a < LU_NB; ++a)
            if (a < nb) v[a] = fma(-s_l[b][a], v[b], v[a]);
#pragma unroll
    for (int a = 0; a < LU_NB; ++a)
        if (a < nb) row[a] = v[a];
}

// explicit inverse of the lower-triangular (non-unit) diagonal block T = A[j0 .. j0+nbj)^2: Linv (128 x 128 row-major,
// identity in the padding).  Column t of the inverse by forward substitution, one thread per column.
__global__ void __launch_bounds__(LU_INV) lu_trinv_kernel(const double* __restrict__ A, long long lda, int n, double* Linv) {
    extern __shared__ double s_t[];  // [LU_INV][LU_INV + 1]
    const int jb = blockIdx.x, j0 = jb * LU_INV, nbj = min(LU_INV, n - j0);
    const int LD = LU_INV + 1;
    for (int e = threadIdx.x; e < LU_INV * LU_INV; e += blockDim.x) {
        const int i = e / LU_INV, k = e % LU_INV;
        double v = (i == k) ? 1.0 : 0.0;
        if (i < nbj && k <= i) v = A[(long long)(j0 + i) * lda + j0 + k];
        s_t[i * LD + k] = v;
    }
    __syncthreads();
    const int t = threadIdx.x;
    double* out = Linv + (long long)jb * LU_INV * LU_INV;
    // rows above t are zero; x_i = (delta_it - sum_{t <= k < i} T[i][k] x_k) / T[i][i], kept in the output array
    for (int i = 0; i < LU_INV; ++i) {
        double acc = (i == t) ? 1.0 : 0.0;
        if (i >= t) {
            for (int k = t; k < i; ++k) acc = fma(-s_t[i * LD + k], out[k * LU_INV + t], acc);
            acc /= s_t[i * LD + i];
        } else {
            acc = 0.0;
        }
        out[i * LU_INV + t] = acc;
    }
}

}  // namespace

long long lu_workspace_doubles(int n) {
    const long long nblk = (n + LU_INV - 1) / LU_INV;
    return nblk * LU_INV * LU_INV + (n + 1) / 2 + 8;
}

int backward_sweep_128(double* A, long long lda, long long strideA, double* B, long long ldb, long long strideB, int n,
                       int nrhs, int batch, const double* linv, long long strideLinv, int jb_first, cudaStream_t stream,
                       int stop_col);

// A: n x n, lower triangle holds the symmetric matrix (beta already on the diagonal); B: nrhs x n right-hand sides,
// overwritten with the solution X of X A = B.  info (device int): 0, or j+1 if no usable pivot exists in column j.
int launch_lu_solve(double* A, long long lda, double* B, long long ldb, int n, int nrhs, int* info, double* work,
                    cudaStream_t stream) {
    XESN_CUDA_OK(cudaMemsetAsync(info, 0, sizeof(int), stream));
    const int nblk = (n + LU_INV - 1) / LU_INV;
    double* linv = work;
    int* piv = reinterpret_cast<int*>(work + (long long)nblk * LU_INV * LU_INV);
    mirror_lower_kernel<<<dim3((n + 255) / 256, n), 256, 0, stream>>>(A, lda, n);
    count_launch();
    for (int k0 = 0; k0 < n; k0 += LU_NB) {
        const int k1 = std::min(k0 + LU_NB, n);
        for (int j = k0; j < k1; ++j) {
            lu_pivot_kernel<<<1, 1024, 0, stream>>>(A, lda, n, j, piv, info);
            lu_swap_kernel<<<(n + nrhs + 255) / 256, 256, 0, stream>>>(A, lda, n, B, ldb, nrhs, j, piv);
            if (j + 1 < n) lu_scale_update_kernel<<<(n - j - 1 + 255) / 256, 256, 0, stream>>>(A, lda, n, j, k1);
            count_launch(3);
        }
        XESN_CUDA_OK(cudaGetLastError());
        const int rest = n - k1;
        lu_panel_trsm_kernel<<<(rest + nrhs + 127) / 128, 128, 0, stream>>>(A, lda, n, B, ldb, nrhs, k0, k1);
        count_launch();
        if (rest > 0) {
            GemmArgs g = {};  // A22 -= L21 U12 in the memory view: C[c][i] -= sum_j A[c][k0+j] * A[k0+j][k1+i]
            g.A = A + (long long)k1 * lda + k0, g.lda = lda;
            g.B = A + (long long)k0 * lda + k1, g.ldb = lda;
            g.C = A + (long long)k1 * lda + k1, g.ldc = lda;
            g.M = rest, g.N = rest, g.K = k1 - k0, g.batch = 1, g.alpha = -1.0, g.beta = 1.0, g.flags = GEMM_A_KMAJOR;
            if (int rc = launch_gemm_f64(g, stream)) return rc;
            GemmArgs b = {};  // the right-hand sides ride along as further columns of the view
            b.A = B + k0, b.lda = ldb;
            b.B = A + (long long)k0 * lda + k1, b.ldb = lda;
            b.C = B + k1, b.ldc = ldb;
            b.M = nrhs, b.N = rest, b.K = k1 - k0, b.batch = 1, b.alpha = -1.0, b.beta = 1.0, b.flags = GEMM_A_KMAJOR;
            if (int rc = launch_gemm_f64(b, stream)) return rc;
        }
    }
    // backward substitution U x = y: U's diagonal blocks are the lower-triangular memory blocks
    const size_t smem = sizeof(double) * LU_INV * (LU_INV + 1);
    XESN_CUDA_OK(cudaFuncSetAttribute(lu_trinv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    lu_trinv_kernel<<<nblk, LU_INV, smem, stream>>>(A, lda, n, linv);
    XESN_CUDA_OK(cudaGetLastError());
    count_launch();
    return backward_sweep_128(A, lda, 0, B, ldb, 0, n, nrhs, 1, linv, 0, nblk - 1, stream, 0);
}

}  // namespace xesn
