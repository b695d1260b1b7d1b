// Device-side FP64 DMMA tile engine shared by the stand-alone GEMM kernel (gemm_f64.cu) and the
// fused recurrence+Gram chunk kernel (fused_train.cu).  See gemm_f64.cu for the design notes.
#pragma once
#include "common.cuh"

namespace xesn {
namespace gemm {

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

// One 128x128 output tile (ti, tj) of batch entry bz, computed by a 256-thread CTA (or the 256
// threads of a role inside a larger persistent kernel).  `smem` must hold SMEM_BYTES.
template <bool A_K, bool B_K, bool VEC>
__device__ __forceinline__ void gemm_tile(const GemmArgs& p, int ti, int tj, int bz, double* smem) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int wm = warp >> 2, wn = warp & 3;

    const int m0 = ti * BM, n0 = tj * BN;
    double* C = p.C + (long long)bz * p.strideC;
    const int bzB = p.modB > 0 ? bz % p.modB : bz;
    const double* bias = p.bias ? p.bias + (long long)bzB * p.strideBias : nullptr;
    __syncthreads();  // persistent callers: every warp has left the previous tile's shared-memory stages

    TileLoader<A_K, VEC> la;
    TileLoader<B_K, VEC> lb;
    la.init(p.A + (long long)bz * p.strideA, p.lda, m0, p.M, tid);
    lb.init(p.B + (long long)bzB * p.strideB, p.ldb, n0, p.N, tid);
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
        const double* As = smem + (kt % STAGES) * STAGE_DOUBLES;
        const double* Bs = As + TILE_DOUBLES;
        // fragments are double-buffered in registers: the LDS of k4-step kk+1 are issued before the
        // 32 DMMA of step kk, so their latency hides behind the tensor pipe (with 2 warps per
        // scheduler nothing else would hide it: ncu showed every first DMMA after an LDS group
        // waiting ~80 cycles).
        double a[2][8], b[2][4];
#pragma unroll
        for (int i = 0; i < 8; ++i) a[0][i] = frag<A_K>(As, wm * 64 + i * 8 + fr, fk);
#pragma unroll
        for (int j = 0; j < 4; ++j) b[0][j] = frag<B_K>(Bs, wn * 32 + j * 8 + fr, fk);
#pragma unroll
        for (int kk = 0; kk < BK / 4; ++kk) {
            const int cur = kk & 1, nxt = cur ^ 1;
            if (kk + 1 < BK / 4) {
#pragma unroll
                for (int i = 0; i < 8; ++i) a[nxt][i] = frag<A_K>(As, wm * 64 + i * 8 + fr, (kk + 1) * 4 + fk);
#pragma unroll
                for (int j = 0; j < 4; ++j) b[nxt][j] = frag<B_K>(Bs, wn * 32 + j * 8 + fr, (kk + 1) * 4 + fk);
            }
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) dmma884(acc[i][j][0], acc[i][j][1], a[cur][i], b[cur][j]);
            if (kk == 0) {
                // refill the stage consumed by the previous k-tile (all its reads finished before
                // this iteration's barrier) only now, so that after the barrier the tensor pipe
                // restarts at once instead of waiting for the copy address arithmetic
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
        }
    }
    cp_async_wait<0>();

    // epilogue: C = alpha*acc + beta*C (+ bias[n]).  The C tile is read in batches of 8 x 16 bytes
    // per thread (two accumulator rows) so a tile pays 4 memory round trips instead of 32.
    const double alpha = p.alpha, beta = p.beta;
    const bool cvec = ((reinterpret_cast<uintptr_t>(C) & 15) == 0) && (p.ldc % 2 == 0);
#pragma unroll
    for (int ib = 0; ib < 4; ++ib) {
        double2 cold[2][4];
#pragma unroll
        for (int ii = 0; ii < 2; ++ii) {
            const int m = m0 + wm * 64 + (2 * ib + ii) * 8 + fr;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int n = n0 + wn * 32 + j * 8 + fk * 2;
                cold[ii][j] = make_double2(0.0, 0.0);
                if (beta != 0.0 && m < p.M) {
                    const double* cp = C + (long long)m * p.ldc + n;
                    if (cvec && n + 1 < p.N) {
                        cold[ii][j] = *reinterpret_cast<const double2*>(cp);
                    } else {
                        if (n < p.N) cold[ii][j].x = cp[0];
                        if (n + 1 < p.N) cold[ii][j].y = cp[1];
                    }
                }
            }
        }
#pragma unroll
        for (int ii = 0; ii < 2; ++ii) {
            const int i = 2 * ib + ii;
            const int m = m0 + wm * 64 + i * 8 + fr;
            if (m >= p.M) continue;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int n = n0 + wn * 32 + j * 8 + fk * 2;
                double2 v;
                v.x = alpha * acc[i][j][0];
                v.y = alpha * acc[i][j][1];
                if (beta != 0.0) {
                    v.x += beta * cold[ii][j].x;
                    v.y += beta * cold[ii][j].y;
                }
                if (bias) {
                    if (n < p.N) v.x += bias[n];
                    if (n + 1 < p.N) v.y += bias[n + 1];
                }
                double* cp = C + (long long)m * p.ldc + n;
                if (cvec && n + 1 < p.N) {
                    *reinterpret_cast<double2*>(cp) = v;
                } else {
                    if (n < p.N) cp[0] = v.x;
                    if (n + 1 < p.N) cp[1] = v.y;
                }
            }
        }
    }
}

// linear index -> (ti >= tj) of the lower triangle, row by row
__device__ __forceinline__ void lower_tile(long long x, int& ti, int& tj) {
    int i = (int)((sqrt(8.0 * (double)x + 1.0) - 1.0) * 0.5);
    while ((long long)(i + 1) * (i + 2) / 2 <= x) ++i;
    while ((long long)i * (i + 1) / 2 > x) --i;
    ti = i;
    tj = (int)(x - (long long)i * (i + 1) / 2);
}

}  // namespace gemm
}  // namespace xesn
