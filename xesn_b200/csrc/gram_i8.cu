// Stand-alone launchers of the int8 (tcgen05) Gram accumulation: digitiser + persistent Gram kernel.
// The production path runs the same device code as a role of the fused chunk kernel (fused_train.cu);
// these entry points serve parity tests, A/B measurements and the non-fused fallback.
#include <vector>

#include "common.cuh"
#include "gram_i8.cuh"
#include "kernels.cuh"

namespace xesn {
namespace {

using namespace gi8;

__global__ void __launch_bounds__(THREADS, 1)
gram_i8_kernel(GramI8Args p, const int2* __restrict__ tiles, int ntiles) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    GramI8Ctx ctx;
    gram_i8_setup(ctx, smem_raw);
    const long long total = (long long)ntiles * p.batch;
    for (long long t = blockIdx.x; t < total; t += gridDim.x) {
        const int2 ij = tiles[t % ntiles];
        gram_i8_tile<THREADS>(ctx, p, ij.x, ij.y, (int)(t / ntiles));
    }
    gram_i8_teardown(ctx);
}

// states [rows][N] f64 (row stride lds) -> planes [8][rows_pad][ldp]; rows >= n_rows and columns >= N are zero
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
    if (a.K % KB) {
        set_error("gram_i8: K=%d must be a multiple of %d (zero-padded rows)", a.K, KB);
        return -1;
    }
    static bool attr = false;
    if (!attr) {
        XESN_CUDA_OK(cudaFuncSetAttribute(gram_i8_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
        attr = true;
    }
    long long total = (long long)ntiles * a.batch;
    int grid = (int)(total < num_sms ? total : num_sms);
    gram_i8_kernel<<<grid, THREADS, SMEM_BYTES, stream>>>(a, tiles, ntiles);
    XESN_CUDA_OK(cudaGetLastError());
    count_launch();
    return 0;
}

}  // namespace xesn
