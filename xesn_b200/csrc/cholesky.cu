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
#include <algorithm>
#include <cstdlib>

#include "common.cuh"
#include "kernels.cuh"

namespace xesn {
namespace {

constexpr int NB = 128;
constexpr int I8_MIN_TRAIL = 1024;  // smaller trailing matrices stay on DMMA
constexpr int FOLD_MAX_RHS = 256;   // right-hand-side rows that may ride along with the factorisation
constexpr int WB = 512;             // width of the explicit inverses used by the backward sweep
constexpr int LDS = NB + 1;
constexpr int PD_THREADS = 256;  // 16 x 16 threads, thread (ty,tx) owns entries (ty+16a, tx+16b), a,b < 8
constexpr int PD_SMEM = (NB * LDS + 4 * NB) * (int)sizeof(double) + 16;

// A: base of matrix 0; block (j0, j0) of size nbj is factored in place; inv(L_jj) is written to
// Linv (NB x NB row-major per matrix, zero above the diagonal, identity in the padding).
// The block lives in REGISTERS (8x8 per thread, cyclic distribution so every thread stays busy as
// the factorisation front advances); shared memory only carries the current column/row.
__global__ void __launch_bounds__(PD_THREADS, 1)
potrf_diag_kernel(double* A, long long lda, long long strideA, int j0, int nbj, double* Linv, long long strideLinv,
                  int* info) {
    extern __shared__ __align__(16) double sm[];
    double* S = sm;                      // [NB][LDS] the factor, for the inversion phase
    double* colbuf = sm + NB * LDS;      // [NB]
    double* rowbuf = colbuf + NB;        // [2][NB]
    const int tid = threadIdx.x;
    const int ty = tid >> 4, tx = tid & 15;
    const int b = blockIdx.x;
    double* Ab = A + (long long)b * strideA + (long long)j0 * lda + j0;
    double* Lb = Linv + (long long)b * strideLinv;

    double a[8][8];
#pragma unroll
    for (int x = 0; x < 8; ++x)
#pragma unroll
        for (int y = 0; y < 8; ++y) {
            const int i = ty + 16 * x, j = tx + 16 * y;
            double v = 0.0;
            if (i < nbj && j <= i) v = Ab[(long long)i * lda + j];
            else if (i >= nbj && i == j) v = 1.0;
            a[x][y] = v;
        }

    // ---- factorisation (right-looking, one column per step, ONE CTA barrier per step) ----
    // After the update of step k-1 the owners of column k (tx == k % 16) publish it RAW -- pivot included -- and every
    // thread scales it itself with rsqrt(pivot) computed redundantly: no owner-only sqrt/divide between two barriers
    // (the first version spent ~1400 cycles per column on sqrt + divide + two barriers; the chain of 125 such blocks
    // is what the look-ahead of the blocked factorisation has to hide).
    // column buffers, double-buffered: colbuf for even k, the first half of rowbuf (free in this phase) for odd k
    if (tx == 0) {
#pragma unroll
        for (int x = 0; x < 8; ++x) colbuf[ty + 16 * x] = a[x][0];
    }
    __syncthreads();
#pragma unroll
    for (int ka = 0; ka < 8; ++ka) {
        for (int kx = 0; kx < 16; ++kx) {
            const int k = ka * 16 + kx;
            const double* col = (k & 1) ? rowbuf : colbuf;
            const double piv = col[k];
            if (tid == 0 && (!(piv > 0.0) || !(piv < 1.7e308))) atomicCAS(&info[b], 0, j0 + k + 1);  // not positive definite
            const double rinv = rsqrt(piv);
            const double d = piv * rinv;
            // scaled column entries this thread needs: rows of its register blocks and columns of its register blocks
            double li[8], lj[8];
#pragma unroll
            for (int x = ka; x < 8; ++x) li[x] = col[ty + 16 * x] * rinv;
#pragma unroll
            for (int y = ka; y < 8; ++y) lj[y] = col[tx + 16 * y] * rinv;
            if (tx == kx) {   // owners of column k keep the final values
#pragma unroll
                for (int x = ka; x < 8; ++x) {
                    const int i = ty + 16 * x;
                    if (i > k) a[x][ka] = li[x];
                    else if (i == k) a[x][ka] = d;
                }
            }
            // only register blocks that can still change: x >= ka (rows > k), ka <= y <= x
#pragma unroll
            for (int x = ka; x < 8; ++x)
#pragma unroll
                for (int y = ka; y <= x; ++y) {
                    const int i = ty + 16 * x, j = tx + 16 * y;
                    if (j > k && j <= i) a[x][y] = fma(-li[x], lj[y], a[x][y]);
                }
            // publish column k+1 (final after this update) for the next step
            // (register arrays need compile-time indices: the two cases of "next column" are spelled out)
            double* nxt = (k & 1) ? colbuf : rowbuf;
            if (kx < 15) {
                if (tx == kx + 1) {
#pragma unroll
                    for (int x = ka; x < 8; ++x) nxt[ty + 16 * x] = a[x][ka];
                }
            } else if (ka + 1 < 8) {
                if (tx == 0) {
#pragma unroll
                    for (int x = ka + 1; x < 8; ++x) nxt[ty + 16 * x] = a[x][(ka + 1) & 7];
                }
            }
            __syncthreads();
        }
    }
#pragma unroll
    for (int x = 0; x < 8; ++x)
#pragma unroll
        for (int y = 0; y < 8; ++y) {
            const int i = ty + 16 * x, j = tx + 16 * y;
            const double v = (j <= i) ? a[x][y] : 0.0;
            S[i * LDS + j] = v;
            if (i < nbj && j <= i) Ab[(long long)i * lda + j] = v;
        }
    __syncthreads();
    // reciprocals of the diagonal, once, instead of a divide inside every step of the substitution below
    double* dinv_all = colbuf;
    if (tid < NB) dinv_all[tid] = 1.0 / S[tid * LDS + tid];
    __syncthreads();

    // ---- inverse: L X = I by forward substitution, one row per step ----
#pragma unroll
    for (int x = 0; x < 8; ++x)
#pragma unroll
        for (int y = 0; y < 8; ++y) a[x][y] = (ty + 16 * x == tx + 16 * y) ? 1.0 : 0.0;
    double* rowbufs = rowbuf;  // [2][NB]
#pragma unroll
    for (int ka = 0; ka < 8; ++ka) {
        for (int kx = 0; kx < 16; ++kx) {
            const int k = ka * 16 + kx;
            double* rb = rowbufs + (k & 1) * NB;
            if (ty == kx) {
                const double dinv = dinv_all[k];
#pragma unroll
                for (int y = 0; y < 8; ++y) {
                    const int j = tx + 16 * y;
                    if (j <= k) {
                        a[ka][y] *= dinv;
                        rb[j] = a[ka][y];
                    } else {
                        rb[j] = 0.0;
                    }
                }
            }
            __syncthreads();
            // rows i > k (x >= ka) and columns j <= k (y <= ka) are the only ones touched
            double lik[8], xk[8];
#pragma unroll
            for (int x = ka; x < 8; ++x) lik[x] = S[(ty + 16 * x) * LDS + k];
#pragma unroll
            for (int y = 0; y <= ka; ++y) xk[y] = rb[tx + 16 * y];
#pragma unroll
            for (int x = ka; x < 8; ++x)
#pragma unroll
                for (int y = 0; y <= ka; ++y) {
                    const int i = ty + 16 * x;
                    if (i > k) a[x][y] = fma(-lik[x], xk[y], a[x][y]);
                }
        }
    }
#pragma unroll
    for (int x = 0; x < 8; ++x)
#pragma unroll
        for (int y = 0; y < 8; ++y) {
            const int i = ty + 16 * x, j = tx + 16 * y;
            Lb[i * NB + j] = (j <= i) ? a[x][y] : 0.0;
        }
}

// the four 128 x 128 inverses of a full 512 block -> its diagonal positions in the 512 x 512 inverse
__global__ void copy_diag_blocks_kernel(const double* __restrict__ linv, double* __restrict__ winv) {
    const int jb = blockIdx.x;  // 128-block index; belongs to wide block jb / 4, slot jb % 4
    const double* src = linv + (long long)jb * NB * NB;
    double* dst = winv + (long long)(jb >> 2) * WB * WB + (long long)(jb & 3) * NB * (WB + 1);
    for (int e = threadIdx.x; e < NB * NB; e += blockDim.x) dst[(e >> 7) * WB + (e & 127)] = src[e];
}

__global__ void add_diag_kernel(double* A, long long lda, long long strideA, int n, const double* beta, double beta_scalar) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double bval = beta ? beta[blockIdx.y] : beta_scalar;
    A[(long long)blockIdx.y * strideA + (long long)i * lda + i] += bval;
}

}  // namespace

// Backward sweep  X T = Z  for the block-lower-triangular T stored in A (row-major, 128-wide blocks whose explicit
// inverses are in linv, [block][128][128] per matrix), Z = B (nrhs x n rows) overwritten with X: block columns
// jb_first down to the one starting at `stop_col`.  Shared by the Cholesky path (T = L) and the pivoted LU fallback
// (T = U^T: the upper factor in the column-major view of the symmetric matrix, lu_solve.cu).
int backward_sweep_128(double* A, long long lda, long long strideA, double* B, long long ldb, long long strideB, int n,
                       int nrhs, int batch, const double* linv, long long strideLinv, int jb_first, cudaStream_t stream,
                       int stop_col = 0) {
    for (int jb = jb_first; jb >= 0 && jb * NB >= stop_col; --jb) {
        const int j0 = jb * NB;
        const int nbj = min(NB, n - j0);
        const double* Linv_j = linv + (long long)jb * NB * NB;
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
        u.B = A + (long long)j0 * lda, u.ldb = lda, u.strideB = strideA;  // B(k,n) = T[j0+k][n]
        u.C = B, u.ldc = ldb, u.strideC = strideB;
        u.M = nrhs, u.N = j0, u.K = nbj;
        u.batch = batch, u.alpha = -1.0, u.beta = 1.0;
        u.flags = GEMM_A_KMAJOR;
        rc = launch_gemm_f64(u, stream);
        if (rc) return rc;
    }
    return 0;
}

int launch_add_diag(double* A, long long lda, long long strideA, int n, int batch, const double* beta_dev,
                    double beta_scalar, cudaStream_t stream) {
    dim3 grid((n + 255) / 256, batch);
    add_diag_kernel<<<grid, 256, 0, stream>>>(A, lda, strideA, n, beta_dev, beta_scalar);
    XESN_CUDA_OK(cudaGetLastError());
    count_launch();
    return 0;
}

long long syrk_i8_workspace_bytes(int n, int K);
int launch_syrk_i8(const double* L, long long ldl, double* C, long long ldc, int n, int ncols, int K, void* work,
                   int num_sms, cudaStream_t stream);

// The big trailing updates run on the int8 tensor cores (row-scaled 8-digit fixed point, gram_i8.cu) when there
// is one large matrix; XESN_B200_SOLVE_I8=0 keeps them on the FP64 DMMA GEMM.
static bool solve_uses_i8(int n, int batch) {
    const char* e = getenv("XESN_B200_SOLVE_I8");
    const bool on = !(e && e[0] == '0');
    if (batch > 1) {   // several large systems: one SYRK per system and panel (XESN_B200_SOLVE_I8_BATCH=0: FP64 DMMA)
        const char* b = getenv("XESN_B200_SOLVE_I8_BATCH");
        if (b && b[0] == '0') return false;
    }
    return on && n >= 2048;
}
// inverses of the full WB x WB diagonal blocks + one 256 x 256 temporary per block
static long long wide_inverse_doubles(int n) { return (long long)(n / WB) * (WB * WB + 256 * 256) + 64; }

// workspace: Linv blocks, nblk * NB*NB doubles per matrix (+ the digit planes of one outer panel)
long long cholesky_workspace_doubles(int n, int nrhs, int batch) {
    long long nblk = (n + NB - 1) / NB;
    long long d = nblk * NB * NB * (long long)batch;
    if (solve_uses_i8(n, batch)) d += 32 + (syrk_i8_workspace_bytes(n + FOLD_MAX_RHS, 512) + 7) / 8;
    if (batch == 1 && n >= 4096) d += wide_inverse_doubles(n) + (long long)nrhs * WB;
    return d;
}

void solve_ctx_destroy(SolveCtx& c) {
    if (c.side) cudaStreamDestroy(c.side);
    for (int k = 0; k < 2; ++k) {
        if (c.ev_panel[k]) cudaEventDestroy(c.ev_panel[k]);
        if (c.ev_trail[k]) cudaEventDestroy(c.ev_trail[k]);
    }
    if (c.ev_fork) cudaEventDestroy(c.ev_fork);
    if (c.ev_join) cudaEventDestroy(c.ev_join);
    c = SolveCtx();
}

static int solve_ctx_init(SolveCtx& c) {
    if (c.side) return 0;
    int lo = 0, hi = 0, dev = 0;
    XESN_CUDA_OK(cudaGetDevice(&dev));
    XESN_CUDA_OK(cudaDeviceGetAttribute(&c.num_sms, cudaDevAttrMultiProcessorCount, dev));
    XESN_CUDA_OK(cudaDeviceGetStreamPriorityRange(&lo, &hi));
    XESN_CUDA_OK(cudaStreamCreateWithPriority(&c.side, cudaStreamNonBlocking, hi));
    for (int k = 0; k < 2; ++k) {
        XESN_CUDA_OK(cudaEventCreateWithFlags(&c.ev_panel[k], cudaEventDisableTiming));
        XESN_CUDA_OK(cudaEventCreateWithFlags(&c.ev_trail[k], cudaEventDisableTiming));
    }
    XESN_CUDA_OK(cudaEventCreateWithFlags(&c.ev_fork, cudaEventDisableTiming));
    XESN_CUDA_OK(cudaEventCreateWithFlags(&c.ev_join, cudaEventDisableTiming));
    return 0;
}

static int cholesky_solve_body(SolveCtx& ctx, double* A, long long lda, long long strideA, double* B, long long ldb,
                               long long strideB, int n, int nrhs, int batch, int* info, double* work,
                               cudaStream_t stream);

int launch_cholesky_solve_ws(SolveCtx& ctx, double* A, long long lda, long long strideA, double* B, long long ldb,
                             long long strideB, int n, int nrhs, int batch, int* info, double* work,
                             cudaStream_t stream) {
    if (int rc = solve_ctx_init(ctx)) return rc;
    // the attribute is per device: set it on every call (cheap) rather than once per process
    XESN_CUDA_OK(cudaFuncSetAttribute(potrf_diag_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, PD_SMEM));
    XESN_CUDA_OK(cudaEventRecord(ctx.ev_fork, stream));
    XESN_CUDA_OK(cudaStreamWaitEvent(ctx.side, ctx.ev_fork, 0));
    const int rc = cholesky_solve_body(ctx, A, lda, strideA, B, ldb, strideB, n, nrhs, batch, info, work, stream);
    // join the side stream on every path, errors included: nothing of this solve may still be running on it when
    // the caller's stream moves on (or the caller frees the buffers)
    cudaEventRecord(ctx.ev_join, ctx.side);
    cudaStreamWaitEvent(stream, ctx.ev_join, 0);
    return rc;
}

static int cholesky_solve_body(SolveCtx& ctx, double* A, long long lda, long long strideA, double* B, long long ldb,
                               long long strideB, int n, int nrhs, int batch, int* info, double* work,
                               cudaStream_t stream) {
    const int nblk = (n + NB - 1) / NB;
    const long long strideLinv = (long long)nblk * NB * NB;
    XESN_CUDA_OK(cudaMemsetAsync(info, 0, sizeof(int) * batch, stream));

    // ---- factorisation: outer panels of OUTER columns, 128-wide sub-blocks inside a panel ----
    // Inside a panel the updates have K = 128; the trailing matrix is updated once per panel with
    // K = OUTER, which keeps the DMMA GEMM in its efficient regime and quarters the read-modify-write
    // traffic of the trailing matrix.  LOOK-AHEAD: the update of panel J is split into the narrow
    // part that touches the next panel's columns (N) and the rest (T); the latency-bound chain of small
    // kernels that factors panel J+1 runs on a second, high-priority stream while T(J) still occupies
    // the GPU on the caller's stream.
    const int OUTER = n >= 4096 ? 512 : (n >= 1024 ? 256 : NB);
    // FOLD: when the right-hand sides are stored as the rows n .. n+nrhs-1 of the same array (B = A + n*lda),
    // every panel operation simply runs over nrhs more rows, which turns B into Z = B L^-T on the way:
    // the forward sweep (125 dependent pairs of small launches at n = 16000) disappears.
    const bool fold = nrhs <= FOLD_MAX_RHS && ldb == lda && B == A + (long long)n * lda && (batch == 1 || strideB == strideA) &&
                      !getenv("XESN_B200_NO_FOLD");
    const int xr = fold ? nrhs : 0;
    cudaStream_t side = ctx.side;
    cudaEvent_t* ev_panel = ctx.ev_panel;
    cudaEvent_t* ev_trail = ctx.ev_trail;

    auto factor_panel = [&](int J, int Jend, cudaStream_t s) -> int {
        for (int j0 = J; j0 < Jend; j0 += NB) {
            const int jb = j0 / NB;
            const int nbj = min(NB, n - j0);
            double* Linv_j = work + (long long)jb * NB * NB;
            potrf_diag_kernel<<<batch, PD_THREADS, PD_SMEM, s>>>(A, lda, strideA, j0, nbj, Linv_j, strideLinv, info);
            XESN_CUDA_OK(cudaGetLastError());
            count_launch();
            const int rest = n - j0 - nbj + xr;
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
            int rc = launch_gemm_f64(g, s);
            if (rc) return rc;
            // remaining columns of this panel: A[j1:, j1:Jend] -= L[j1:, j0:j1] L[j1:Jend, j0:j1]^T
            const int j1 = j0 + nbj;
            const int pw = Jend - j1;
            if (pw > 0) {
                GemmArgs u = {};
                u.A = A + (long long)j1 * lda + j0, u.lda = lda, u.strideA = strideA;
                u.B = A + (long long)j1 * lda + j0, u.ldb = lda, u.strideB = strideA;
                u.C = A + (long long)j1 * lda + j1, u.ldc = lda, u.strideC = strideA;
                u.M = n - j1 + xr, u.N = pw, u.K = nbj;
                u.batch = batch, u.alpha = -1.0, u.beta = 1.0;
                u.flags = GEMM_A_KMAJOR | GEMM_B_KMAJOR;
                rc = launch_gemm_f64(u, s);
                if (rc) return rc;
            }
        }
        return 0;
    };

    int rc = factor_panel(0, min(OUTER, n), side);
    if (rc) return rc;
    int it = 0;
    for (int J = 0; J < n; J += OUTER, ++it) {
        const int Jend = min(J + OUTER, n);
        const int Jend2 = min(Jend + OUTER, n);
        XESN_CUDA_OK(cudaEventRecord(ev_panel[it & 1], side));            // panel J is final
        if (Jend >= n) break;
        // N(J) touches the same elements as T(J-1) (both update the columns of panel J+1): order them
        if (it >= 1) XESN_CUDA_OK(cudaStreamWaitEvent(side, ev_trail[(it - 1) & 1], 0));
        // N(J) on the side stream: columns of the next panel, A[Jend:, Jend:Jend2] -= L[Jend:, J:Jend] L[Jend:Jend2, J:Jend]^T
        {
            GemmArgs u = {};
            u.A = A + (long long)Jend * lda + J, u.lda = lda, u.strideA = strideA;
            u.B = A + (long long)Jend * lda + J, u.ldb = lda, u.strideB = strideA;
            u.C = A + (long long)Jend * lda + Jend, u.ldc = lda, u.strideC = strideA;
            u.M = n - Jend + xr, u.N = Jend2 - Jend, u.K = Jend - J;
            u.batch = batch, u.alpha = -1.0, u.beta = 1.0;
            u.flags = GEMM_A_KMAJOR | GEMM_B_KMAJOR;
            if ((rc = launch_gemm_f64(u, side))) return rc;
        }
        // T(J) on the caller's stream: everything right of the next panel (lower tiles)
        XESN_CUDA_OK(cudaStreamWaitEvent(stream, ev_panel[it & 1], 0));
        if (n - Jend2 >= I8_MIN_TRAIL && solve_uses_i8(n, batch) && (Jend - J) % 32 == 0) {
            const int num_sms = ctx.num_sms;
            // the persistent SYRK kernel would otherwise hold every SM until it ends and starve the look-ahead
            // stream: leave some SMs to the panel factorisation (measured at N=16000, ms per solve: 0 free 42.8,
            // 8: 41.9, 16: 41.0, 32: 38.8, 48: 40.3, 64: 43.0)
            // (several systems, one SYRK after the other: 16 -- cfg 4 share 170.5 / 173.5 / 177.3 / 180.6 ms at 16 / 32 / 48 / 64)
            int free_sms = batch == 1 ? 32 : 16;
            if (const char* e = getenv("XESN_B200_SYRK_FREE")) free_sms = atoi(e);
            const int syrk_sms = std::max(1, num_sms - free_sms);
            double* w8 = work + strideLinv * batch;
            w8 = reinterpret_cast<double*>((reinterpret_cast<uintptr_t>(w8) + 255) & ~(uintptr_t)255);
            for (int b = 0; b < batch; ++b) {   // the systems share the digit-plane workspace: one after the other
                double* Ab = A + (long long)b * strideA;
                if ((rc = launch_syrk_i8(Ab + (long long)Jend2 * lda + J, lda, Ab + (long long)Jend2 * lda + Jend2, lda,
                                         n - Jend2 + xr, n - Jend2, Jend - J, w8, syrk_sms, stream)))
                    return rc;
            }
        } else if (n - Jend2 > 0) {
            if (xr) {  // right-hand-side rows: B[:, Jend2:] -= Z[:, J:Jend] L[Jend2:, J:Jend]^T
                GemmArgs t = {};
                t.A = A + (long long)n * lda + J, t.lda = lda, t.strideA = strideA;
                t.B = A + (long long)Jend2 * lda + J, t.ldb = lda, t.strideB = strideA;
                t.C = A + (long long)n * lda + Jend2, t.ldc = lda, t.strideC = strideA;
                t.M = xr, t.N = n - Jend2, t.K = Jend - J;
                t.batch = batch, t.alpha = -1.0, t.beta = 1.0;
                t.flags = GEMM_A_KMAJOR | GEMM_B_KMAJOR;
                if ((rc = launch_gemm_f64(t, stream))) return rc;
            }
            GemmArgs t = {};
            t.A = A + (long long)Jend2 * lda + J, t.lda = lda, t.strideA = strideA;
            t.B = t.A, t.ldb = lda, t.strideB = strideA;
            t.C = A + (long long)Jend2 * lda + Jend2, t.ldc = lda, t.strideC = strideA;
            t.M = n - Jend2, t.N = n - Jend2, t.K = Jend - J;
            t.batch = batch, t.alpha = -1.0, t.beta = 1.0;
            t.flags = GEMM_A_KMAJOR | GEMM_B_KMAJOR | GEMM_LOWER;
            if ((rc = launch_gemm_f64(t, stream))) return rc;
        }
        XESN_CUDA_OK(cudaEventRecord(ev_trail[it & 1], stream));
        // panel J+1 needs N(J) and T(J-1), both ordered before this point on the side stream;
        // it runs while T(J) occupies the caller's stream
        if ((rc = factor_panel(Jend, Jend2, side))) return rc;
    }
    // join: the sweeps below run on the caller's stream
    XESN_CUDA_OK(cudaStreamWaitEvent(stream, ev_panel[it & 1], 0));

    // ---- forward sweep:  Z L^T = B   (B is nrhs x n, row-major: every row is one right-hand side) ----
    for (int jb = 0; jb < nblk && !fold; ++jb) {
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
    // Big single systems: explicit inverses of the full 512 x 512 diagonal blocks are assembled from the 128-wide
    // ones (inv [[P,0],[C,Q]] = [[P^-1,0],[-Q^-1 C P^-1, Q^-1]], two batched levels), so the sweep is 2 launches per
    // 512 columns instead of 2 per 128 -- the chain of dependent small launches was 6 of the solve's 54 ms.
    const int n_wide = (batch == 1 && n >= 4096) ? n / WB : 0;  // full WB blocks [0, n_wide*WB)
    double *Winv = nullptr, *Xtmp = nullptr;
    if (n_wide > 0) {
        double* wbase = work + strideLinv * batch;
        if (solve_uses_i8(n, batch)) wbase += 32 + (syrk_i8_workspace_bytes(n + FOLD_MAX_RHS, 512) + 7) / 8;
        Winv = reinterpret_cast<double*>((reinterpret_cast<uintptr_t>(wbase) + 255) & ~(uintptr_t)255);
        double* Tmp = Winv + (long long)n_wide * WB * WB;
        Xtmp = Tmp + (long long)n_wide * 256 * 256;
        XESN_CUDA_OK(cudaMemsetAsync(Winv, 0, sizeof(double) * (size_t)n_wide * WB * WB, stream));
        // diagonal 128-blocks
        copy_diag_blocks_kernel<<<dim3(4 * n_wide), 256, 0, stream>>>(work, Winv);
        XESN_CUDA_OK(cudaGetLastError());
        count_launch();
        auto level = [&](int h, int slot_count) -> int {
            // h = half size (128 or 256); slot s places the pair at rows/cols s*2h inside the 512 block
            for (int sl = 0; sl < slot_count; ++sl) {
                const long long o = (long long)sl * 2 * h;  // offset of the pair inside the block
                GemmArgs g = {};  // Tmp = C * inv(P)
                g.A = A + (o + h) * lda + o, g.lda = lda, g.strideA = (long long)WB * (lda + 1);
                g.B = Winv + o * WB + o, g.ldb = WB, g.strideB = (long long)WB * WB;
                g.C = Tmp, g.ldc = h, g.strideC = 256 * 256;
                g.M = h, g.N = h, g.K = h, g.batch = n_wide, g.alpha = 1.0, g.beta = 0.0;
                g.flags = GEMM_A_KMAJOR;
                int rc = launch_gemm_f64(g, stream);
                if (rc) return rc;
                GemmArgs w = {};  // W = -inv(Q) * Tmp
                w.A = Winv + (o + h) * WB + (o + h), w.lda = WB, w.strideA = (long long)WB * WB;
                w.B = Tmp, w.ldb = h, w.strideB = 256 * 256;
                w.C = Winv + (o + h) * WB + o, w.ldc = WB, w.strideC = (long long)WB * WB;
                w.M = h, w.N = h, w.K = h, w.batch = n_wide, w.alpha = -1.0, w.beta = 0.0;
                w.flags = GEMM_A_KMAJOR;
                rc = launch_gemm_f64(w, stream);
                if (rc) return rc;
            }
            return 0;
        };
        if ((rc = level(NB, 2))) return rc;
        if ((rc = level(2 * NB, 1))) return rc;
    }
    // columns beyond the full wide blocks (and everything when there are none): 128-wide steps
    if ((rc = backward_sweep_128(A, lda, strideA, B, ldb, strideB, n, nrhs, batch, work, strideLinv, nblk - 1, stream,
                                 n_wide * WB)))
        return rc;
    for (int J = n_wide - 1; J >= 0; --J) {
        const int j0 = J * WB;
        GemmArgs g = {};  // X_J = Z_J * inv(L_JJ)   (four column tiles read all of Z_J: not in place)
        g.A = B + j0, g.lda = ldb, g.strideA = strideB;
        g.B = Winv + (long long)J * WB * WB, g.ldb = WB, g.strideB = 0;
        g.C = Xtmp, g.ldc = WB, g.strideC = 0;
        g.M = nrhs, g.N = WB, g.K = WB, g.batch = 1, g.alpha = 1.0, g.beta = 0.0;
        g.flags = GEMM_A_KMAJOR;
        int rc = launch_gemm_f64(g, stream);
        if (rc) return rc;
        XESN_CUDA_OK(cudaMemcpy2DAsync(B + j0, sizeof(double) * ldb, Xtmp, sizeof(double) * WB, sizeof(double) * WB, nrhs,
                                       cudaMemcpyDeviceToDevice, stream));
        if (j0 == 0) break;
        GemmArgs u = {};  // Z[:, :j0] -= X_J * L[j0:j0+WB, :j0]
        u.A = Xtmp, u.lda = WB, u.strideA = 0;
        u.B = A + (long long)j0 * lda, u.ldb = lda, u.strideB = strideA;
        u.C = B, u.ldc = ldb, u.strideC = strideB;
        u.M = nrhs, u.N = j0, u.K = WB, u.batch = 1, u.alpha = -1.0, u.beta = 1.0;
        u.flags = GEMM_A_KMAJOR;
        rc = launch_gemm_f64(u, stream);
        if (rc) return rc;
    }
    return 0;
}

}  // namespace xesn
