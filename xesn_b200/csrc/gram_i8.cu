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

__global__ void __launch_bounds__(THREADS, 1)
gram_i8_kernel(GramI8Args p, const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
               const int2* __restrict__ tiles, int ntiles) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    GramI8Ctx ctx;
    gram_i8_setup(ctx, smem_raw);
    gram_i8_run<THREADS>(ctx, p, &tmA, &tmB, tiles, ntiles, blockIdx.x, gridDim.x);
    gram_i8_teardown(ctx);
}

// states [rows][N] f64 (row stride lds) -> planes [NDIG][rows_pad][ldp]; rows >= n_rows and columns >= N are zero
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
    if (a.batch > 1 && a.group_stride != (long long)NDIG * a.plane_stride) {
        set_error("gram_i8: groups must be stored back to back (group stride = %d planes)", NDIG);
        return -1;
    }
    if ((a.ldp % 16) || (a.plane_stride % 16) || (reinterpret_cast<uintptr_t>(a.planes) & 15)) {
        set_error("gram_i8: digit planes must be 16-byte aligned with row and plane pitches that are multiples of 16");
        return -1;
    }
    const cuuint64_t gdim[3] = {(cuuint64_t)a.ldp, (cuuint64_t)a.K, (cuuint64_t)NDIG * (cuuint64_t)a.batch};
    const cuuint64_t gstride[2] = {(cuuint64_t)a.ldp, (cuuint64_t)a.plane_stride};
    const cuuint32_t estride[3] = {1, 1, 1};
    const cuuint32_t box_a[3] = {128, (cuuint32_t)KB, (cuuint32_t)NDIG};
    const cuuint32_t box_b[3] = {64, (cuuint32_t)KB, (cuuint32_t)NDIG};
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
        XESN_CUDA_OK(cudaFuncSetAttribute(gram_i8_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
        attr = true;
    }
    alignas(64) CUtensorMap tm_a, tm_b;
    if (int rc = gram_i8_make_maps(a, &tm_a, &tm_b)) return rc;
    long long total = (long long)ntiles * a.batch;
    int grid = (int)(total < num_sms ? total : num_sms);
    gram_i8_kernel<<<grid, THREADS, SMEM_BYTES, stream>>>(a, tm_a, tm_b, tiles, ntiles);
    XESN_CUDA_OK(cudaGetLastError());
    count_launch();
    return 0;
}

}  // namespace xesn
