// Batched blocked Cholesky + triangular solves for the ridge system (sm_100a).
//
// Replaces `scipy.linalg.solve(rbar.T, ybar.T, assume_a="sym").T` (reference
// xesn/esn.py:516-520): Wout (O x N) with  Wout (R R^T + beta I) = Y R^T.
// R R^T + beta I is symmetric positive definite for beta > 0, so a Cholesky factorisation
// A = L L^T is used instead of LAPACK's Bunch-Kaufman LDL^T; a non-positive pivot is reported
// through `info` (LAPACK convention) instead of being pivoted around.
//
// Right-looking, block size 128.  Per block column j:
//   1. potrf_diag_kernel : factor the 128x128 diagonal block in shared memory and form its
//                          explicit inverse (one CTA per matrix of the batch);
//   2. panel             : L_ij = A_ij inv(L_jj)^T           -> DMMA GEMM (gemm_f64.cu)
//   3. trailing update   : A_22 -= L_21 L_21^T (lower tiles) -> DMMA GEMM
// The two triangular solves with the (O x N) right-hand side are also pure GEMM sweeps that
// use the stored diagonal-block inverses, so >99% of the flops run on the tensor pipe.
#include "common.cuh"
#include "kernels.cuh"

namespace xesn {
namespace {

constexpr int NB = 128;
constexpr int LDS = NB + 1;
constexpr int PD_THREADS = 256;
constexpr int PD_SMEM = (NB * LDS + NB * (NB + 1) / 2) * (int)sizeof(double) + 16;

__device__ __forceinline__ int tri(int i, int j) { return i * (i + 1) / 2 + j; }

// A: base of matrix 0; block (j0, j0) of size nbj is factored in place; inv(L_jj) is written to
// Linv (NB x NB row-major per matrix, zero above the diagonal, identity in the padding).
__global__ void __launch_bounds__(PD_THREADS, 1)
potrf_diag_kernel(double* A, long long lda, long long strideA, int j0, int nbj, double* Linv, long long strideLinv,
                  int* info) {
    extern __shared__ __align__(16) double sm[];
    double* S = sm;             // [NB][LDS]
    double* X = sm + NB * LDS;  // packed lower triangle
    __shared__ int s_fail;
    const int tid = threadIdx.x;
    const int b = blockIdx.x;
    double* Ab = A + (long long)b * strideA + (long long)j0 * lda + j0;
    double* Lb = Linv + (long long)b * strideLinv;
    if (tid == 0) s_fail = 0;

    for (int e = tid; e < NB * NB; e += PD_THREADS) {
        int i = e / NB, j = e % NB;
        double v = 0.0;
        if (i < nbj && j <= i) v = Ab[(long long)i * lda + j];
        else if (i >= nbj && i == j) v = 1.0;
        S[i * LDS + j] = v;
    }
    __syncthreads();

    for (int k = 0; k < NB; ++k) {
        double piv = S[k * LDS + k];
        if (!(piv > 0.0) || !(piv < 1.7e308)) {  // <= 0, NaN or inf: not positive definite
            if (tid == 0 && !s_fail) {
                s_fail = 1;
                atomicCAS(&info[b], 0, j0 + k + 1);
            }
        }
        double d = sqrt(piv);
        double rinv = 1.0 / d;
        __syncthreads();
        if (tid == 0) S[k * LDS + k] = d;
        for (int i = k + 1 + tid; i < NB; i += PD_THREADS) S[i * LDS + k] *= rinv;
        __syncthreads();
        const int w = NB - 1 - k;
        for (int e = tid; e < w * w; e += PD_THREADS) {
            int i = k + 1 + e / w, j = k + 1 + e % w;
            if (j <= i) S[i * LDS + j] -= S[i * LDS + k] * S[j * LDS + k];
        }
        __syncthreads();
    }

    // inverse: thread j solves L x = e_j by forward substitution (column j of inv(L))
    if (tid < NB) {
        const int j = tid;
        X[tri(j, j)] = 1.0 / S[j * LDS + j];
        for (int i = j + 1; i < NB; ++i) {
            double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
            int m = j;
            for (; m + 4 <= i; m += 4) {
                a0 = fma(S[i * LDS + m], X[tri(m, j)], a0);
                a1 = fma(S[i * LDS + m + 1], X[tri(m + 1, j)], a1);
                a2 = fma(S[i * LDS + m + 2], X[tri(m + 2, j)], a2);
                a3 = fma(S[i * LDS + m + 3], X[tri(m + 3, j)], a3);
            }
            for (; m < i; ++m) a0 = fma(S[i * LDS + m], X[tri(m, j)], a0);
            X[tri(i, j)] = -((a0 + a1) + (a2 + a3)) / S[i * LDS + i];
        }
    }
    __syncthreads();

    for (int e = tid; e < NB * NB; e += PD_THREADS) {
        int i = e / NB, j = e % NB;
        if (i < nbj && j <= i) Ab[(long long)i * lda + j] = S[i * LDS + j];
        Lb[e] = (j <= i) ? X[tri(i, j)] : 0.0;
    }
}

__global__ void add_diag_kernel(double* A, long long lda, long long strideA, int n, const double* beta, double beta_scalar) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double bval = beta ? beta[blockIdx.y] : beta_scalar;
    A[(long long)blockIdx.y * strideA + (long long)i * lda + i] += bval;
}

}  // namespace

int launch_add_diag(double* A, long long lda, long long strideA, int n, int batch, const double* beta_dev,
                    double beta_scalar, cudaStream_t stream) {
    dim3 grid((n + 255) / 256, batch);
    add_diag_kernel<<<grid, 256, 0, stream>>>(A, lda, strideA, n, beta_dev, beta_scalar);
    XESN_CUDA_OK(cudaGetLastError());
    count_launch();
    return 0;
}

// workspace: Linv blocks, nblk * NB*NB doubles per matrix
long long cholesky_workspace_doubles(int n, int batch) {
    long long nblk = (n + NB - 1) / NB;
    return nblk * NB * NB * (long long)batch;
}

int launch_cholesky_solve_ws(double* A, long long lda, long long strideA, double* B, long long ldb, long long strideB,
                             int n, int nrhs, int batch, int* info, double* work, cudaStream_t stream) {
    static bool attr = false;
    if (!attr) {
        XESN_CUDA_OK(cudaFuncSetAttribute(potrf_diag_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, PD_SMEM));
        attr = true;
    }
    const int nblk = (n + NB - 1) / NB;
    const long long strideLinv = (long long)nblk * NB * NB;
    XESN_CUDA_OK(cudaMemsetAsync(info, 0, sizeof(int) * batch, stream));

    // ---- factorisation ----
    for (int jb = 0; jb < nblk; ++jb) {
        const int j0 = jb * NB;
        const int nbj = min(NB, n - j0);
        double* Linv_j = work + (long long)jb * NB * NB;
        potrf_diag_kernel<<<batch, PD_THREADS, PD_SMEM, stream>>>(A, lda, strideA, j0, nbj, Linv_j, strideLinv, info);
        XESN_CUDA_OK(cudaGetLastError());
        count_launch();
        const int rest = n - j0 - nbj;
        if (rest <= 0) break;
        GemmArgs g = {};
        // panel: L_ij = A_ij * inv(L_jj)^T   (in place: a tile reads only the rows it writes)
        g.A = A + (long long)(j0 + nbj) * lda + j0;
        g.lda = lda, g.strideA = strideA;
        g.B = Linv_j, g.ldb = NB, g.strideB = strideLinv;
        g.C = A + (long long)(j0 + nbj) * lda + j0;
        g.ldc = lda, g.strideC = strideA;
        g.M = rest, g.N = nbj, g.K = nbj;
        g.batch = batch, g.alpha = 1.0, g.beta = 0.0;
        g.flags = GEMM_A_KMAJOR | GEMM_B_KMAJOR;
        int rc = launch_gemm_f64(g, stream);
        if (rc) return rc;
        // trailing: A_22 -= L_21 L_21^T (lower tiles only)
        GemmArgs t = {};
        t.A = A + (long long)(j0 + nbj) * lda + j0;
        t.lda = lda, t.strideA = strideA;
        t.B = t.A, t.ldb = lda, t.strideB = strideA;
        t.C = A + (long long)(j0 + nbj) * lda + (j0 + nbj);
        t.ldc = lda, t.strideC = strideA;
        t.M = rest, t.N = rest, t.K = nbj;
        t.batch = batch, t.alpha = -1.0, t.beta = 1.0;
        t.flags = GEMM_A_KMAJOR | GEMM_B_KMAJOR | GEMM_LOWER;
        rc = launch_gemm_f64(t, stream);
        if (rc) return rc;
    }

    // ---- forward sweep:  Z L^T = B   (B is nrhs x n, row-major: every row is one right-hand side) ----
    for (int jb = 0; jb < nblk; ++jb) {
        const int j0 = jb * NB;
        const int nbj = min(NB, n - j0);
        double* Linv_j = work + (long long)jb * NB * NB;
        GemmArgs g = {};
        g.A = B + j0, g.lda = ldb, g.strideA = strideB;
        g.B = Linv_j, g.ldb = NB, g.strideB = strideLinv;  // B(k,n) = Linv[n][k]
        g.C = B + j0, g.ldc = ldb, g.strideC = strideB;
        g.M = nrhs, g.N = nbj, g.K = nbj;
        g.batch = batch, g.alpha = 1.0, g.beta = 0.0;
        g.flags = GEMM_A_KMAJOR | GEMM_B_KMAJOR;
        int rc = launch_gemm_f64(g, stream);
        if (rc) return rc;
        const int rest = n - j0 - nbj;
        if (rest <= 0) break;
        GemmArgs u = {};
        u.A = B + j0, u.lda = ldb, u.strideA = strideB;
        u.B = A + (long long)(j0 + nbj) * lda + j0, u.ldb = lda, u.strideB = strideA;  // B(k,n) = L[j0+nbj+n][j0+k]
        u.C = B + j0 + nbj, u.ldc = ldb, u.strideC = strideB;
        u.M = nrhs, u.N = rest, u.K = nbj;
        u.batch = batch, u.alpha = -1.0, u.beta = 1.0;
        u.flags = GEMM_A_KMAJOR | GEMM_B_KMAJOR;
        rc = launch_gemm_f64(u, stream);
        if (rc) return rc;
    }

    // ---- backward sweep:  X L = Z ----
    for (int jb = nblk - 1; jb >= 0; --jb) {
        const int j0 = jb * NB;
        const int nbj = min(NB, n - j0);
        double* Linv_j = work + (long long)jb * NB * NB;
        GemmArgs g = {};
        g.A = B + j0, g.lda = ldb, g.strideA = strideB;
        g.B = Linv_j, g.ldb = NB, g.strideB = strideLinv;  // B(k,n) = Linv[k][n]
        g.C = B + j0, g.ldc = ldb, g.strideC = strideB;
        g.M = nrhs, g.N = nbj, g.K = nbj;
        g.batch = batch, g.alpha = 1.0, g.beta = 0.0;
        g.flags = GEMM_A_KMAJOR;
        int rc = launch_gemm_f64(g, stream);
        if (rc) return rc;
        if (j0 == 0) break;
        GemmArgs u = {};
        u.A = B + j0, u.lda = ldb, u.strideA = strideB;
        u.B = A + (long long)j0 * lda, u.ldb = lda, u.strideB = strideA;  // B(k,n) = L[j0+k][n]
        u.C = B, u.ldc = ldb, u.strideC = strideB;
        u.M = nrhs, u.N = j0, u.K = nbj;
        u.batch = batch, u.alpha = -1.0, u.beta = 1.0;
        u.flags = GEMM_A_KMAJOR;
        rc = launch_gemm_f64(u, stream);
        if (rc) return rc;
    }
    return 0;
}

}  // namespace xesn
