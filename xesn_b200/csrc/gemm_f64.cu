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
#include "common.cuh"

namespace xesn {
namespace {

constexpr int BM = 128, BN = 128, BK = 16;
constexpr int NTHREADS = 256;
constexpr int STAGES = 4;
constexpr int LD_MN = BM + 4;  // tile stored [BK][LD_MN]  (operand MN-contiguous in global)
constexpr int LD_K = BK + 4;   // tile stored [BM][LD_K]   (operand K-contiguous in global)
constexpr int TILE_DOUBLES = BM * LD_K;  // 2560 >= BK*LD_MN = 2112
constexpr int STAGE_DOUBLES = 2 * TILE_DOUBLES;
constexpr int SMEM_BYTES = STAGES * STAGE_DOUBLES * (int)sizeof(double);  // 160 KB

__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
    asm volatile("cp.async.wait_group %0;\n" ::"n"(N) : "memory");
}

__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(c0), "+d"(c1)
                 : "d"(a), "d"(b));
}

// Per-thread state for copying one operand tile (128 along MN, 16 along K) per k-tile into shared
// memory, zero-filling outside [0,MN) x [0,K).  Everything that does not depend on the k-tile
// (source pointers, shared-memory offsets, MN-edge byte counts) is computed once; inside the main
// loop a copy is one add + one LDGSTS.  (The first version recomputed ~900 integer instructions per
// k-tile per warp, which left the DMMA pipe idle 22 % of the time -- profiles/gram_syrk_r01_v1.txt.)
template <bool KMAJOR, bool VEC>
struct TileLoader {
    const char* src[4];
    const char* base;
    unsigned dst[4];
    int mn_bytes[4];
    int kpos[4];
    long long step;

    __device__ __forceinline__ void init(const double* g, long long ld, int mn0, int MN, int tid) {
        base = reinterpret_cast<const char*>(g);
#pragma unroll
        for (int p = 0; p < 4; ++p) {
            if (!KMAJOR) {
                const int row = p * 4 + (tid >> 6), v = tid & 63;
                const int mn = mn0 + 2 * v;
                mn_bytes[p] = min(max(MN - mn, 0), 2) * 8;
                src[p] = mn_bytes[p] ? reinterpret_cast<const char*>(g + (long long)row * ld + mn) : base;
                dst[p] = (unsigned)((row * LD_MN + 2 * v) * sizeof(double));
                kpos[p] = row;
            } else {
                const int row = p * 32 + (tid >> 3), v = tid & 7;
                const int mn = mn0 + row;
                mn_bytes[p] = (mn < MN) ? 16 : 0;
                src[p] = mn_bytes[p] ? reinterpret_cast<const char*>(g + (long long)mn * ld + 2 * v) : base;
                dst[p] = (unsigned)((row * LD_K + 2 * v) * sizeof(double));
                kpos[p] = 2 * v;
            }
        }
        step = KMAJOR ? (long long)BK * sizeof(double) : (long long)BK * ld * sizeof(double);
    }

    // `smem_tile`: shared-space byte address of this operand's tile in the target stage
    template <bool FULL>
    __device__ __forceinline__ void load(unsigned smem_tile, int kt, int K) const {
        const long long off = (long long)kt * step;
#pragma unroll
        for (int p = 0; p < 4; ++p) {
            int bytes = mn_bytes[p];
            if (!FULL) {
                const int k = kt * BK + kpos[p];
                if (!KMAJOR) bytes = (k < K) ? bytes : 0;
                else bytes = min(bytes, max(K - k, 0) * 8);
            }
            const char* sp = bytes ? src[p] + off : base;
            const unsigned d = smem_tile + dst[p];
            if (VEC) {
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(d), "l"(sp), "r"(bytes) : "memory");
            } else {
                const int b0 = min(bytes, 8), b1 = max(bytes - 8, 0);
                asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;\n" ::"r"(d), "l"(sp), "r"(b0) : "memory");
                asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;\n" ::"r"(d + 8), "l"(b1 ? sp + 8 : base),
                             "r"(b1)
                             : "memory");
            }
        }
    }
};

template <bool KMAJOR>
__device__ __forceinline__ double frag(const double* s, int mn, int k) {
    return KMAJOR ? s[mn * LD_K + k] : s[k * LD_MN + mn];
}

template <bool A_K, bool B_K, bool VEC>
__global__ void __launch_bounds__(NTHREADS, 1) gemm_f64_kernel(GemmArgs p, int tiles_n) {
    extern __shared__ __align__(16) double smem[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int wm = warp >> 2, wn = warp & 3;

    int ti, tj;
    if (p.flags & GEMM_LOWER) {
        // linear index -> (ti >= tj) of the lower triangle
        long long x = blockIdx.x;
        int i = (int)((sqrt(8.0 * (double)x + 1.0) - 1.0) * 0.5);
        while ((long long)(i + 1) * (i + 2) / 2 <= x) ++i;
        while ((long long)i * (i + 1) / 2 > x) --i;
        ti = i;
        tj = (int)(x - (long long)i * (i + 1) / 2);
    } else {
        ti = blockIdx.x / tiles_n;
        tj = blockIdx.x % tiles_n;
    }
    const int m0 = ti * BM, n0 = tj * BN;
    const int bz = blockIdx.y;
    double* C = p.C + (long long)bz * p.strideC;

    TileLoader<A_K, VEC> la;
    TileLoader<B_K, VEC> lb;
    la.init(p.A + (long long)bz * p.strideA, p.lda, m0, p.M, tid);
    lb.init(p.B + (long long)bz * p.strideB, p.ldb, n0, p.N, tid);
    const unsigned smem_base = (unsigned)__cvta_generic_to_shared(smem);
    constexpr unsigned STAGE_BYTES = STAGE_DOUBLES * sizeof(double);
    constexpr unsigned TILE_BYTES = TILE_DOUBLES * sizeof(double);

    double acc[8][4][2];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

    const int ktiles = (p.K + BK - 1) / BK;
    const int kfull = p.K / BK;  // k-tiles that need no K-edge handling
#pragma unroll
    for (int s = 0; s < STAGES - 1; ++s) {
        if (s < ktiles) {
            if (s < kfull) {
                la.template load<true>(smem_base + s * STAGE_BYTES, s, p.K);
                lb.template load<true>(smem_base + s * STAGE_BYTES + TILE_BYTES, s, p.K);
            } else {
                la.template load<false>(smem_base + s * STAGE_BYTES, s, p.K);
                lb.template load<false>(smem_base + s * STAGE_BYTES + TILE_BYTES, s, p.K);
            }
        }
        cp_async_commit();
    }

    const int fr = lane >> 2, fk = lane & 3;
    for (int kt = 0; kt < ktiles; ++kt) {
        cp_async_wait<STAGES - 2>();
        __syncthreads();
        {
            const int nk = kt + STAGES - 1;
            if (nk < ktiles) {
                const unsigned sb = smem_base + (nk % STAGES) * STAGE_BYTES;
                if (nk < kfull) {
                    la.template load<true>(sb, nk, p.K);
                    lb.template load<true>(sb + TILE_BYTES, nk, p.K);
                } else {
                    la.template load<false>(sb, nk, p.K);
                    lb.template load<false>(sb + TILE_BYTES, nk, p.K);
                }
            }
            cp_async_commit();
        }
        const double* As = smem + (kt % STAGES) * STAGE_DOUBLES;
        const double* Bs = As + TILE_DOUBLES;
#pragma unroll
        for (int kk = 0; kk < BK / 4; ++kk) {
            double a[8], b[4];
#pragma unroll
            for (int i = 0; i < 8; ++i) a[i] = frag<A_K>(As, wm * 64 + i * 8 + fr, kk * 4 + fk);
#pragma unroll
            for (int j = 0; j < 4; ++j) b[j] = frag<B_K>(Bs, wn * 32 + j * 8 + fr, kk * 4 + fk);
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) dmma884(acc[i][j][0], acc[i][j][1], a[i], b[j]);
        }
    }
    cp_async_wait<0>();

    // epilogue: C = alpha*acc + beta*C (+ bias[n])
    const double alpha = p.alpha, beta = p.beta;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        int m = m0 + wm * 64 + i * 8 + fr;
        if (m >= p.M) continue;
        double* crow = C + (long long)m * p.ldc;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            int n = n0 + wn * 32 + j * 8 + fk * 2;
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                if (n + e < p.N) {
                    double v = alpha * acc[i][j][e];
                    if (beta != 0.0) v += beta * crow[n + e];
                    if (p.bias) v += p.bias[n + e];
                    crow[n + e] = v;
                }
            }
        }
    }
}

inline bool aligned16(const void* p) { return ((uintptr_t)p & 15) == 0; }

}  // namespace

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
    static bool attr_done[8] = {false, false, false, false, false, false, false, false};
    if (!attr_done[which]) {
        XESN_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
        attr_done[which] = true;
    }
    kern<<<grid, NTHREADS, SMEM_BYTES, stream>>>(a, tiles_n);
    count_launch();
    XESN_CUDA_OK(cudaGetLastError());
    return 0;
}

}  // namespace xesn
