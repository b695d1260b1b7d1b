// Shared helpers for the xesn_b200 CUDA library (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>

namespace xesn {

// last error string, returned through the C ABI (xesn_last_error)
void set_error(const char* fmt, ...);
// kernel-launch counter exported as xesn_launch_count()
void count_launch(int n = 1);

#define XESN_CUDA_OK(expr)                                                              \
    do {                                                                                \
        cudaError_t _e = (expr);                                                        \
        if (_e != cudaSuccess) {                                                        \
            ::xesn::set_error("%s failed at %s:%d: %s", #expr, __FILE__, __LINE__,      \
                              cudaGetErrorString(_e));                                  \
            return (int)_e;                                                             \
        }                                                                               \
    } while (0)

#define XESN_REQUIRE(cond, msg)                                                         \
    do {                                                                                \
        if (!(cond)) {                                                                  \
            ::xesn::set_error("invalid argument: %s (%s:%d)", msg, __FILE__, __LINE__); \
            return -1;                                                                  \
        }                                                                               \
    } while (0)

// optional per-phase device timers (xesn_profile_enable / xesn_profile_read)
enum ProfClass : int { PROF_DRIVE = 0, PROF_RECUR, PROF_GRAM, PROF_YR, PROF_SOLVE, PROF_PREDICT, PROF_GATHER, PROF_FUSED, PROF_N };
void prof_begin(int cls, cudaStream_t st);
void prof_end(int cls, cudaStream_t st);
struct ProfScope {
    int cls;
    cudaStream_t st;
    ProfScope(int c, cudaStream_t s) : cls(c), st(s) { prof_begin(cls, st); }
    ~ProfScope() { prof_end(cls, st); }
};

// ---- internal launchers shared between translation units -------------------------------------
enum GemmFlags : int {
    GEMM_A_KMAJOR = 1,   // A[m*lda + k]   (else A[k*lda + m])
    GEMM_B_KMAJOR = 2,   // B[n*ldb + k]   (else B[k*ldb + n])
    GEMM_LOWER = 4,      // only tiles touching the lower triangle (tile_n <= tile_m) are computed
};

struct GemmArgs {
    const double* A;
    const double* B;
    double* C;
    const double* bias;  // optional, length N, added to every row of C
    long long strideBias;  // per batch entry (0 = shared)
    int modB;              // > 0: B and bias of batch entry bz are those of entry bz % modB (parameter sets repeated)
    int M, N, K;
    long long lda, ldb, ldc;
    long long strideA, strideB, strideC;  // per batch entry (0 = shared)
    int batch;
    double alpha, beta;
    int flags;
};

// C = alpha * op(A) op(B) + beta * C (+ bias), FP64 on the tensor pipe (DMMA m8n8k4)
int launch_gemm_f64(const GemmArgs& a, cudaStream_t stream);
// the same contraction for a skinny A (n_output <= 64 target rows, time-contiguous) and B[k*ldb + n]: yr_skinny.cuh
int launch_yr_skinny(const GemmArgs& a, cudaStream_t stream);
constexpr int YR_SKINNY_MAX_ROWS = 64;

}  // namespace xesn
