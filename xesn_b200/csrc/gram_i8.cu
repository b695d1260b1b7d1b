// Stand-alone launchers of the int8 (tcgen05) Gram accumulation: digitiser + persistent Gram kernel.
// The production path runs the same device code as a role of the fused chunk kernel (fused_train.cu);
// these entry points serve parity tests, A/B measurements and the non-fused fallback.
#include <cuda.h>

#include <vector>

#include "common.cuh"
#include "gram_i8.cuh"
#include "kernels.cuh"

namespace xesn {
namespace {

using namespace gi8;

template <int ND>
__global__ void __launch_bounds__(THREADS, 1)
gram_i8_kernel(GramI8Args p, const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
               const int2* __restrict__ tiles, int ntiles) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    GramI8Ctx ctx;
    gram_i8_setup<ND>(ctx, smem_raw);
    gram_i8_run<THREADS, ND>(ctx, p, &tmA, &tmB, tiles, ntiles, blockIdx.x, gridDim.x);
    gram_i8_teardown(ctx);
}

// ---- row-scaled 8-digit image of a panel (the SYRK operand of the blocked Cholesky) -------------------
// scale[i] = 2^e >= max_k |L[i][k]| (1 for an all-zero or non-finite row)
__global__ void rowscale_kernel(const double* __restrict__ L, long long ldl, int n_rows, int K, double* __restrict__ scale) {
    const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (row >= n_rows) return;
    const double* x = L + (long long)row * ldl;
    double m = 0.0;
    for (int k = lane; k < K; k += 32) m = fmax(m, fabs(x[k]));
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) m = fmax(m, __shfl_xor_sync(0xffffffffu, m, off));
    if (lane == 0) {
        double s = 1.0;
        if (m > 0.0 && m < 1.7e308) {
            int e;
            frexp(m, &e);        // m = f * 2^e, 0.5 <= f < 1
            s = ldexp(1.0, e);
        }
        scale[row] = s;
    }
}

// q = rint(x / scale * 2^62) = sum_a d_a 256^(7-a), balanced digits; returned as the 8 bytes d0..d7
__device__ __forceinline__ void wide_digits(double x, double inv_scale_2p62, unsigned& w03, unsigned& w47) {
    const double y = x * inv_scale_2p62;
    long long q = (fabs(y) <= 4611686018427387904.0 /* 2^62 */) ? __double2ll_rn(y) : 0;
    constexpr long long BIAS = 128LL * (1LL + (1LL << 8) + (1LL << 16) + (1LL << 24) + (1LL << 32) + (1LL << 40) + (1LL << 48));
    q += BIAS;
    const unsigned lo = (unsigned)q, hi = (unsigned)((unsigned long long)q >> 32);
    w03 = __byte_perm(lo, hi, 0x4567) ^ 0x80808000u;  // (top, b6, b5, b4): the top digit is signed as is
    w47 = __byte_perm(lo, hi, 0x0123) ^ 0x80808080u;  // (b3, b2, b1, b0)
}

// L [n_rows][K] (row pitch ldl) -> planes [8][K][ldp] with planes[a][k][i] = digit a of L[i][k] / scale[i]
// block: 64 rows x 32 columns of L, transposed through shared memory so that both sides are coalesced
__global__ void __launch_bounds__(256) digitize_panel_kernel(const double* __restrict__ L, long long ldl, int n_rows, int K,
                                                             const double* __restrict__ scale, int8_t* __restrict__ planes,
                                                             long long ldp, long long plane_stride) {
    __shared__ __align__(16) signed char dig[NDIG_WIDE][32][64 + 16];
    const int i0 = blockIdx.x * 64, k0 = blockIdx.y * 32;
    const int t = threadIdx.x;
    {
        const int il = t >> 2, kq = (t & 3) * 8;
        const int i = i0 + il;
        double x[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) x[j] = (i < n_rows && k0 + kq + j < K) ? L[(long long)i * ldl + k0 + kq + j] : 0.0;
        const double inv = (i < n_rows) ? 4611686018427387904.0 / scale[i] : 0.0;  // 2^62 / 2^e: exact
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            unsigned w03, w47;
            wide_digits(x[j], inv, w03, w47);
#pragma unroll
            for (int a = 0; a < 4; ++a) {
                dig[a][kq + j][il] = (signed char)((w03 >> (8 * a)) & 255u);
                dig[a + 4][kq + j][il] = (signed char)((w47 >> (8 * a)) & 255u);
            }
        }
    }
    __syncthreads();
    {
        const int a = t >> 5, k = t & 31;
        if (k0 + k < K) {
            int8_t* dst = planes + a * plane_stride + (long long)(k0 + k) * ldp + i0;
            const int4* src = reinterpret_cast<const int4*>(&dig[a][k][0]);
#pragma unroll
            for (int c = 0; c < 4; ++c)
                if (i0 + c * 16 < ldp) reinterpret_cast<int4*>(dst)[c] = src[c];
        }
    }
}

// states [rows][N] f64 (row stride lds) -> planes [NDIG][rows_pad][ldp]; rows >= n_rows and columns >= N are zero
__global__ void fill_kernel(double* x, int n, double v) {
    for (int i = threadIdx.x; i < n; i += blockDim.x) x[i] = v;
}

__global__ void digitize_kernel(const double* __restrict__ states, long long lds, int n_rows, int N, int8_t* planes,
                                long long ldp, long long plane_stride, int rows_pad) {
    const long long col = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const int row = blockIdx.y;
    if (col >= ldp || row >= rows_pad) return;
    signed char d[NDIG];
    if (row < n_rows && col < N) {
        state_digits(states[(long long)row * lds + col], d);
    } else {
#pragma unroll
        for (int a = 0; a < NDIG; ++a) d[a] = 0;
    }
#pragma unroll
    for (int a = 0; a < NDIG; ++a) planes[a * plane_stride + (long long)row * ldp + col] = d[a];
}

}  // namespace

// lower-triangle list of 128 x 64 tiles (ti: 128-row block, tj: 64-column block) in super-rows of
// 12 tile rows, column by column, so that co-running CTAs share operand panels in L2
void gram_i8_tile_list(int N, std::vector<int2>& tiles) {
    const int TI = (N + TM - 1) / TM, H = 12;
    tiles.clear();
    for (int sb = 0; sb < TI; sb += H) {
        const int ti_hi = sb + H < TI ? sb + H : TI;
        const int tj_hi = (ti_hi * TM + TN - 1) / TN;  // columns needed by the last row block of this super-row
        for (int tj = 0; tj < tj_hi; ++tj)
            for (int ti = sb; ti < ti_hi; ++ti)
                if ((long long)tj * TN <= (long long)ti * TM + TM - 1 && tj * TN < N) tiles.push_back(make_int2(ti, tj));
    }
}

// TMA descriptors of the digit planes seen as a 3-D byte tensor (N bytes, time steps, planes of all groups):
// box A = 128 B x 32 steps x 7 planes (SWIZZLE_128B), box B = 64 B x 32 x 7 (SWIZZLE_64B)
int gram_i8_make_maps(const gi8::GramI8Args& a, void* map_a, void* map_b) {
    typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                 const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                 CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    static EncodeFn encode = nullptr;
    if (!encode) {
        void* fn = nullptr;
        cudaDriverEntryPointQueryResult qres;
        XESN_CUDA_OK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
        if (!fn || qres != cudaDriverEntryPointSuccess) {
            set_error("gram_i8: cuTensorMapEncodeTiled is not available in this driver");
            return -3;
        }
        encode = (EncodeFn)fn;
    }
    const int nd = a.ndig ? a.ndig : NDIG;
    if (a.batch > 1 && a.group_stride != (long long)nd * a.plane_stride) {
        set_error("gram_i8: groups must be stored back to back (group stride = %d planes)", nd);
        return -1;
    }
    if ((a.ldp % 16) || (a.plane_stride % 16) || (reinterpret_cast<uintptr_t>(a.planes) & 15)) {
        set_error("gram_i8: digit planes must be 16-byte aligned with row and plane pitches that are multiples of 16");
        return -1;
    }
    const cuuint64_t gdim[3] = {(cuuint64_t)a.ldp, (cuuint64_t)a.K, (cuuint64_t)nd * (cuuint64_t)a.batch};
    const cuuint64_t gstride[2] = {(cuuint64_t)a.ldp, (cuuint64_t)a.plane_stride};
    const cuuint32_t estride[3] = {1, 1, 1};
    const cuuint32_t box_a[3] = {128, (cuuint32_t)KB, (cuuint32_t)nd};
    const cuuint32_t box_b[3] = {64, (cuuint32_t)KB, (cuuint32_t)nd};
    void* base = const_cast<int8_t*>(a.planes);
    CUresult r1 = encode((CUtensorMap*)map_a, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, base, gdim, gstride, box_a, estride,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    CUresult r2 = encode((CUtensorMap*)map_b, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, base, gdim, gstride, box_b, estride,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r1 != CUDA_SUCCESS || r2 != CUDA_SUCCESS) {
        set_error("gram_i8: cuTensorMapEncodeTiled failed (%d, %d)", (int)r1, (int)r2);
        return -3;
    }
    return 0;
}

int launch_digitize(const double* states, long long lds, int n_rows, int N, int8_t* planes, long long ldp,
                    long long plane_stride, int rows_pad, cudaStream_t stream) {
    dim3 grid((unsigned)((ldp + 255) / 256), (unsigned)rows_pad);
    digitize_kernel<<<grid, 256, 0, stream>>>(states, lds, n_rows, N, planes, ldp, plane_stride, rows_pad);
    XESN_CUDA_OK(cudaGetLastError());
    count_launch();
    return 0;
}

int launch_gram_i8(const gi8::GramI8Args& a, const int2* tiles, int ntiles, int num_sms, cudaStream_t stream) {
    if (ntiles <= 0 || a.batch <= 0 || a.K <= 0) return 0;
    if (a.K > MAX_K) {
        set_error("gram_i8: K=%d exceeds %d steps per launch (int32 accumulators)", a.K, MAX_K);
        return -1;
    }
    if (a.K % KB) {
        set_error("gram_i8: K=%d must be a multiple of %d (zero-padded rows)", a.K, KB);
        return -1;
    }
    static bool attr = false;
    if (!attr) {
        XESN_CUDA_OK(cudaFuncSetAttribute(gram_i8_kernel<NDIG>, cudaFuncAttributeMaxDynamicSharedMemorySize, Ring<NDIG>::SMEM_BYTES));
        XESN_CUDA_OK(cudaFuncSetAttribute(gram_i8_kernel<NDIG_WIDE>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                          Ring<NDIG_WIDE>::SMEM_BYTES));
        attr = true;
    }
    alignas(64) CUtensorMap tm_a, tm_b;
    if (int rc = gram_i8_make_maps(a, &tm_a, &tm_b)) return rc;
    long long total = (long long)ntiles * a.batch;
    int grid = (int)(total < num_sms ? total : num_sms);
    if (a.ndig == NDIG_WIDE)
        gram_i8_kernel<NDIG_WIDE><<<grid, THREADS, Ring<NDIG_WIDE>::SMEM_BYTES, stream>>>(a, tm_a, tm_b, tiles, ntiles);
    else
        gram_i8_kernel<NDIG><<<grid, THREADS, Ring<NDIG>::SMEM_BYTES, stream>>>(a, tm_a, tm_b, tiles, ntiles);
    XESN_CUDA_OK(cudaGetLastError());
    count_launch();
    return 0;
}


// C (n x n, lower tiles, row pitch ldc) -= L L^T for the panel L (n x K, row pitch ldl), on the int8 tensor cores:
// rows are scaled to |x| <= 1 by a power of two, cut into 8 balanced base-256 digits (2^-63 of the row scale,
// i.e. finer than the FP64 rounding of the products themselves) and multiplied plane by plane.
// work: planes (8 * round_up(K,32) * ldp bytes, ldp = round_up(n,128)) followed by n doubles of scales.
long long syrk_i8_workspace_bytes(int n, int K) {
    const long long ldp = (n + 127) / 128 * 128, kp = (K + KB - 1) / KB * KB;
    return NDIG_WIDE * kp * ldp + 256 + 8 * ldp;
}

// ncols < n: only the first ncols columns of C exist (the rows beyond them are right-hand-side rows carried along).
int launch_syrk_i8(const double* L, long long ldl, double* C, long long ldc, int n, int ncols, int K, void* work,
                   int num_sms, cudaStream_t stream) {
    if (n <= 0 || K <= 0) return 0;
    const long long ldp = (n + 127) / 128 * 128, kp = (K + KB - 1) / KB * KB;
    int8_t* planes = reinterpret_cast<int8_t*>(work);
    double* scale = reinterpret_cast<double*>(planes + ((NDIG_WIDE * kp * ldp + 255) / 256) * 256);
    if (ldp > n) {  // scales of the padding columns (read by the epilogue of the last column tile)
        fill_kernel<<<1, 128, 0, stream>>>(scale + n, (int)(ldp - n), 1.0);
        XESN_CUDA_OK(cudaGetLastError());
        count_launch();
    }
    rowscale_kernel<<<(n + 7) / 8, 256, 0, stream>>>(L, ldl, n, K, scale);
    XESN_CUDA_OK(cudaGetLastError());
    count_launch();
    dim3 grid((unsigned)(ldp / 64), (unsigned)(kp / 32));
    digitize_panel_kernel<<<grid, 256, 0, stream>>>(L, ldl, n, K, scale, planes, ldp, ldp * kp);
    XESN_CUDA_OK(cudaGetLastError());
    count_launch();
    GramI8Args a = {};
    a.planes = planes, a.ldp = ldp, a.plane_stride = ldp * kp, a.group_stride = 0;
    a.K = (int)kp, a.N = n, a.G = C, a.ldg = ldc, a.strideG = 0, a.batch = 1;
    a.ndig = NDIG_WIDE, a.scale = scale, a.ncols = ncols;
    const int TI = (n + TM - 1) / TM;
    return launch_gram_i8(a, nullptr, TI * (TI + 1), num_sms, stream);
}

}  // namespace xesn
