// ybar += Y R for a SKINNY Y (the n_output targets of a reservoir: 16 ... 64 rows) -- xesn/esn.py:509.
//
// The DMMA GEMM of gemm_tile.cuh works on 128 x 128 tiles: with 16 target rows 7/8 of every tile is padding, and the
// launch (0.28 ms per 2048-step chunk for 16 groups of N=2000) is 3x over its memory bound -- the state ring R is
// read once (K x N x 8 bytes per reservoir), everything else is tiny.  Here a work item is (reservoir, slab of 64
// columns): the warps split the time steps, a lane owns two adjacent columns and keeps 16 x 2 accumulators in
// registers; R is streamed with 16-byte loads (512 bytes per warp and step, coalesced), Y comes from a shared-memory
// tile by broadcast.  The warps' partial sums are combined in a fixed order.  Used stand-alone and by the idle Gram
// CTAs of the fused training kernel (fused_train.cu), which hides it completely in recurrence-bound configurations.
//   C[m][n] += sum_k A[m*lda + k] * B[k*ldb + n]      (A = targets, time-contiguous; B = state ring rows)
#pragma once
#include "common.cuh"

namespace xesn {
namespace yr {

constexpr int TN = 64;    // columns per work item
constexpr int TM = 16;    // target rows per pass
constexpr int KT = 256;   // time steps of Y staged per tile

template <int NTHREADS>
constexpr int smem_bytes() {
    return (TM * KT + (NTHREADS / 32) * TM * TN) * (int)sizeof(double);
}

__host__ __device__ inline long long n_items(const GemmArgs& p) { return (long long)p.batch * ((p.N + TN - 1) / TN); }

template <int NTHREADS>
__device__ __forceinline__ void yr_item(const GemmArgs& p, long long item, double* smem) {
    constexpr int NW = NTHREADS / 32;
    const int slabs = (p.N + TN - 1) / TN;
    const int bz = (int)(item / slabs), n0 = (int)(item % slabs) * TN;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const double* A = p.A + (long long)bz * p.strideA;
    const double* B = p.B + (long long)bz * p.strideB;
    double* C = p.C + (long long)bz * p.strideC;
    double* s_y = smem;                 // [TM][KT]
    double* s_red = smem + TM * KT;     // [NW][TM][TN]
    const int col = n0 + 2 * lane;
    const bool vec = (p.ldb % 2 == 0) && ((reinterpret_cast<uintptr_t>(B) & 15) == 0) && col + 1 < p.N;
    for (int m0 = 0; m0 < p.M; m0 += TM) {
        double acc[TM][2];
#pragma unroll
        for (int o = 0; o < TM; ++o) acc[o][0] = acc[o][1] = 0.0;
        for (int k0 = 0; k0 < p.K; k0 += KT) {
            const int kt = min(KT, p.K - k0);
            __syncthreads();  // the previous tile's readers are done
            for (int e = tid; e < TM * KT; e += NTHREADS) {
                const int o = e / KT, k = e % KT;
                s_y[e] = (m0 + o < p.M && k < kt) ? A[(long long)(m0 + o) * p.lda + k0 + k] : 0.0;
            }
            __syncthreads();
            for (int k = warp; k < kt; k += NW) {
                double r0 = 0.0, r1 = 0.0;
                const double* row = B + (long long)(k0 + k) * p.ldb + col;
                if (vec) {
                    const double2 v = *reinterpret_cast<const double2*>(row);
                    r0 = v.x, r1 = v.y;
                } else {
                    if (col < p.N) r0 = row[0];
                    if (col + 1 < p.N) r1 = row[1];
                }
#pragma unroll
                for (int o = 0; o < TM; ++o) {
                    const double y = s_y[o * KT + k];
                    acc[o][0] = fma(y, r0, acc[o][0]);
                    acc[o][1] = fma(y, r1, acc[o][1]);
                }
            }
        }
#pragma unroll
        for (int o = 0; o < TM; ++o) {
            s_red[(warp * TM + o) * TN + 2 * lane] = acc[o][0];
            s_red[(warp * TM + o) * TN + 2 * lane + 1] = acc[o][1];
        }
        __syncthreads();
        for (int e = tid; e < TM * TN; e += NTHREADS) {
            const int o = e / TN, n = e % TN;
            if (m0 + o < p.M && n0 + n < p.N) {
                double s = 0.0;
#pragma unroll
                for (int w = 0; w < NW; ++w) s += s_red[(w * TM + o) * TN + n];
                double* c = C + (long long)(m0 + o) * p.ldc + n0 + n;
                *c = p.beta * *c + p.alpha * s;
            }
        }
    }
    __syncthreads();
}

}  // namespace yr
}  // namespace xesn

// ---- drive of a chunk, D[g][t][n] = sum_i U_g[i][t] Win[n][i] + b[n], for a SMALL number of inputs -----------------
// (xesn/esn.py:459 `Win @ u` for every step of the chunk).  With 18 ... 64 inputs this is an outer-product-like pass
// bound by writing D (m x N x 8 bytes per reservoir), not a GEMM worth 128 x 128 tiles.  A work item is (reservoir,
// block of 16 time steps): the inputs of the block sit in shared memory, a thread owns a column n and keeps the 16
// sums in registers, streaming its row of Win once; stores are coalesced along n.  Run by the Gram CTAs of the fused
// training kernel for chunk c+1 while the recurrence role works on chunk c.
namespace xesn {
namespace drv {

constexpr int TB = 16;          // time steps per work item (16 sums in registers: the fused kernel lives under a 128-register cap)
constexpr int MAX_INPUTS = 64;  // shared-memory tile [I][TB]

struct DriveArgs {
    const double* U;      // row i of reservoir g, time t: U[g*ugs + i*ldu + t]
    long long ldu, ugs;
    const double* Win;    // [N][I] per parameter set
    const double* bias;   // [N]
    long long win_gs, bias_gs;
    int param_groups;     // reservoir g uses parameter set g % param_groups (0 / 1: shared)
    double* D;            // [G][m][N]
    int m, N, I, G;
};

__host__ __device__ inline long long n_items(const DriveArgs& p) { return (long long)p.G * ((p.m + TB - 1) / TB); }

template <int NTHREADS>
__device__ __forceinline__ void drive_item(const DriveArgs& p, long long item, double* smem) {
    const int blocks = (p.m + TB - 1) / TB;
    const int g = (int)(item / blocks), t0 = (int)(item % blocks) * TB;
    const int nt = min(TB, p.m - t0);
    const int pg = g % max(p.param_groups, 1);
    const double* U = p.U + (long long)g * p.ugs + t0;
    const double* Win = p.Win + (long long)pg * p.win_gs;
    const double* bias = p.bias + (long long)pg * p.bias_gs;
    double* D = p.D + ((long long)g * p.m + t0) * p.N;
    __syncthreads();
    for (int e = threadIdx.x; e < p.I * TB; e += NTHREADS) {
        const int i = e / TB, t = e % TB;
        smem[e] = t < nt ? U[(long long)i * p.ldu + t] : 0.0;
    }
    __syncthreads();
    for (int n = threadIdx.x; n < p.N; n += NTHREADS) {
        double acc[TB];
#pragma unroll
        for (int t = 0; t < TB; ++t) acc[t] = 0.0;
        const double* w = Win + (long long)n * p.I;
        for (int i = 0; i < p.I; ++i) {
            const double wi = __ldg(w + i);
            const double2* u2 = reinterpret_cast<const double2*>(smem + i * TB);
#pragma unroll
            for (int t = 0; t < TB / 2; ++t) {
                const double2 u = u2[t];
                acc[2 * t] = fma(u.x, wi, acc[2 * t]);
                acc[2 * t + 1] = fma(u.y, wi, acc[2 * t + 1]);
            }
        }
        const double b = __ldg(bias + n);
#pragma unroll
        for (int t = 0; t < TB; ++t)
            if (t < nt) D[(long long)t * p.N + n] = acc[t] + b;
    }
}

}  // namespace drv
}  // namespace xesn
