// Device-side cost of a batch of forecasts: the NRMSE term of the reference's macro-training cost
// (xesn/cost.py:227-254 `nrmse`, :201-225 the per-candidate average and the clamp).
//
//   nrmse(f, g) = sqrt( mean_{o, s} ( (pred[f][g][o][s] - truth[f][o][s]) / std_s(truth[f][o][:]) )^2 )
//   cost(g)     = mean_f nrmse(f, g);   NaN, inf or > upper_bound  ->  upper_bound
//
// with the standard deviation over the forecast times of one variable (ddof 0, numpy's default, which
// xarray's `.std("ftime")` uses).  The forecasts of an ESNEnsemble never leave the device: only the G costs do.
// HBM-bound streaming of the forecasts (8 bytes per value, read once); one CTA per (window, candidate).
#include "common.cuh"

namespace xesn {
namespace {

constexpr int CT = 256;

__device__ __forceinline__ double block_sum(double v, double* s_red) {
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    __syncthreads();
    if (lane == 0) s_red[warp] = v;
    __syncthreads();
    double t = 0.0;
    for (int w = 0; w < CT / 32; ++w) t += s_red[w];  // same order in every thread: deterministic
    return t;
}

// pred: [F][G][O][S] contiguous; truth row (f, o) = truth[(f*O + o) * ld + t0 + s]
__global__ void __launch_bounds__(CT) nrmse_kernel(const double* __restrict__ pred, const double* __restrict__ truth,
                                                   long long ld, long long t0, int G, int O, int S,
                                                   double* __restrict__ out_fg) {
    __shared__ double s_red[CT / 32];
    const int f = blockIdx.x / G, g = blockIdx.x % G;
    double total = 0.0;
    for (int o = 0; o < O; ++o) {
        const double* tr = truth + ((long long)f * O + o) * ld + t0;
        const double* pr = pred + (((long long)f * G + g) * O + o) * S;
        double a = 0.0;
        for (int s = threadIdx.x; s < S; s += CT) a += tr[s];
        const double mean = block_sum(a, s_red) / S;
        double v = 0.0, e = 0.0;
        for (int s = threadIdx.x; s < S; s += CT) {
            const double d = tr[s] - mean, q = pr[s] - tr[s];
            v += d * d;
            e += q * q;
        }
        const double var = block_sum(v, s_red) / S;
        const double err = block_sum(e, s_red);
        total += err / var;  // sum_s ((pred - truth) / std)^2
    }
    if (threadIdx.x == 0) out_fg[blockIdx.x] = sqrt(total / ((double)O * S));
}

// NRMSE of the 1-D power spectral density (xesn/cost.py:257-272 `psd_nrmse` over xesn/psd.py:12-65 `psd` for a field
// with ONE spatial axis, e.g. Lorenz-96): for every time, |FFT over the O variables|^2 at the integer wavenumbers
// k = 1 .. O/2 times the annulus factor pi((k+.5)^2 - (k-.5)^2); then the NRMSE over (k, time) with the temporal
// standard deviation of the truth's spectrum per wavenumber.  For even O the reference's last bin (k = O/2) is empty
// (scipy's fftfreq files that wavenumber as -O/2, outside the bins) and xarray's skipna mean ignores it: skipped here.
// One CTA per (window, candidate); the O-point transform is a direct sum with an exact-argument twiddle table.
__global__ void __launch_bounds__(CT) psd_nrmse_kernel(const double* __restrict__ pred, const double* __restrict__ truth,
                                                       long long ld, long long t0, int G, int O, int S,
                                                       double* __restrict__ out_fg) {
    extern __shared__ double s_dyn[];  // [O] cos, [O] sin, [S] truth spectrum, [S] forecast spectrum
    __shared__ double s_red[CT / 32];
    double* s_cos = s_dyn;
    double* s_sin = s_cos + O;
    double* s_pt = s_sin + O;
    double* s_pp = s_pt + S;
    const int f = blockIdx.x / G, g = blockIdx.x % G;
    for (int m = threadIdx.x; m < O; m += CT) sincospi(2.0 * m / O, &s_sin[m], &s_cos[m]);
    __syncthreads();
    const int kmax = (O % 2 == 0) ? O / 2 - 1 : O / 2;  // wavenumbers that own a non-empty bin
    double total = 0.0;
    for (int k = 1; k <= kmax; ++k) {
        const double annulus = 3.14159265358979323846 * ((k + 0.5) * (k + 0.5) - (k - 0.5) * (k - 0.5));
        for (int s = threadIdx.x; s < S; s += CT) {
            double tre = 0.0, tim = 0.0, pre = 0.0, pim = 0.0;
            int m = 0;  // (k * o) mod O
            for (int o = 0; o < O; ++o) {
                const double xt = truth[((long long)f * O + o) * ld + t0 + s];
                const double xp = pred[(((long long)f * G + g) * O + o) * S + s];
                tre = fma(xt, s_cos[m], tre), tim = fma(-xt, s_sin[m], tim);
                pre = fma(xp, s_cos[m], pre), pim = fma(-xp, s_sin[m], pim);
                m += k;
                if (m >= O) m -= O;
            }
            s_pt[s] = (tre * tre + tim * tim) * annulus;
            s_pp[s] = (pre * pre + pim * pim) * annulus;
        }
        __syncthreads();
        double a = 0.0;
        for (int s = threadIdx.x; s < S; s += CT) a += s_pt[s];
        const double mean = block_sum(a, s_red) / S;
        double v = 0.0, e = 0.0;
        for (int s = threadIdx.x; s < S; s += CT) {
            const double d = s_pt[s] - mean, q = s_pp[s] - s_pt[s];
            v += d * d;
            e += q * q;
        }
        const double var = block_sum(v, s_red) / S;
        const double err = block_sum(e, s_red);
        total += err / var;
        __syncthreads();
    }
    if (threadIdx.x == 0) out_fg[blockIdx.x] = kmax > 0 ? sqrt(total / ((double)kmax * S)) : 0.0;
}

// Radially binned power spectrum of fields with two or three spatial axes (xesn/psd.py:42-65): for image `img` (one
// forecast time of one field) out[img][b] = factor[b] * mean over the spectral cells p of annulus b of
// (mean over z of |spec[img][z][p]|^2); an annulus without cells gives NaN like scipy's binned_statistic.  The cell
// lists per annulus come from the host (they depend only on the grid size); a warp owns an annulus and adds its
// cells in a fixed order, so the result is deterministic.  spec: interleaved (re, im) pairs.
__global__ void __launch_bounds__(CT) psd_radial_kernel(const double2* __restrict__ spec, int n_z, int n_pix,
                                                        const int* __restrict__ bin_ptr, const int* __restrict__ pix_idx,
                                                        const double* __restrict__ factor, int n_bins,
                                                        double* __restrict__ out) {
    const long long img = blockIdx.x;
    const double2* base = spec + img * (long long)n_z * n_pix;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int b = warp; b < n_bins; b += CT / 32) {
        const int lo = bin_ptr[b], hi = bin_ptr[b + 1];
        double acc = 0.0;
        for (int e = lo + lane; e < hi; e += 32) {
            const int p = pix_idx[e];
            double zsum = 0.0;
            for (int z = 0; z < n_z; ++z) {
                const double2 c = base[(long long)z * n_pix + p];
                zsum += c.x * c.x + c.y * c.y;
            }
            acc += zsum / n_z;
        }
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, off);
        if (lane == 0) out[img * n_bins + b] = hi > lo ? acc / (hi - lo) * factor[b] : __longlong_as_double(0x7FF8000000000000LL);
    }
}

// NRMSE between two sets of rows with xarray's skip-NaN reductions (xesn/cost.py:227-254 applied to spectra,
// :257-272): per row k the truth's temporal mean / variance over its non-NaN samples, the squared normalised error
// over the samples where both are numbers, then the mean over every such (k, s) and the square root.
// pred [F][G][K][S], truth [F][K][S]; one CTA per (f, g).
__global__ void __launch_bounds__(CT) nrmse_skipnan_kernel(const double* __restrict__ pred, const double* __restrict__ truth,
                                                           int G, int K, int S, double* __restrict__ out_fg) {
    __shared__ double s_red[CT / 32];
    const int f = blockIdx.x / G, g = blockIdx.x % G;
    double total = 0.0, count = 0.0;
    for (int k = 0; k < K; ++k) {
        const double* tr = truth + ((long long)f * K + k) * S;
        const double* pr = pred + (((long long)f * G + g) * K + k) * S;
        double a = 0.0, n = 0.0;
        for (int s = threadIdx.x; s < S; s += CT) {
            const double t = tr[s];
            if (t == t) a += t, n += 1.0;
        }
        const double nt = block_sum(n, s_red);
        const double mean = block_sum(a, s_red) / nt;
        double v = 0.0, e = 0.0, m = 0.0;
        for (int s = threadIdx.x; s < S; s += CT) {
            const double t = tr[s], q = pr[s];
            if (t == t) {
                v += (t - mean) * (t - mean);
                if (q == q) e += (q - t) * (q - t), m += 1.0;
            }
        }
        const double var = block_sum(v, s_red) / nt;
        const double err = block_sum(e, s_red);
        const double cnt = block_sum(m, s_red);
        if (nt > 0.0 && cnt > 0.0) total += err / var, count += cnt;   // an all-NaN row has NaN weights: skipped
    }
    if (threadIdx.x == 0) out_fg[blockIdx.x] = sqrt(total / count);
}

__global__ void cost_mean_kernel(const double* __restrict__ fg, int F, int G, double upper, double* __restrict__ cost) {
    const int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= G) return;
    double a = 0.0;
    for (int f = 0; f < F; ++f) a += fg[(long long)f * G + g];
    a /= F;
    if (!(a <= upper) || isinf(a)) a = upper;  // NaN fails the comparison
    cost[g] = a;
}

}  // namespace
}  // namespace xesn

using namespace xesn;

extern "C" int xesn_cost_nrmse(const double* pred_dev, const double* truth_dev, int64_t truth_ld, int64_t t0,
                               int32_t n_windows, int32_t n_groups, int32_t n_output, int32_t n_times,
                               double upper_bound, double* nrmse_fg_dev, double* cost_dev, void* stream) {
    XESN_REQUIRE(pred_dev && truth_dev && nrmse_fg_dev && cost_dev, "null pointer");
    XESN_REQUIRE(n_windows > 0 && n_groups > 0 && n_output > 0 && n_times > 0 && truth_ld >= t0 + n_times, "bad sizes");
    cudaStream_t st = (cudaStream_t)stream;
    nrmse_kernel<<<(unsigned)(n_windows * n_groups), CT, 0, st>>>(pred_dev, truth_dev, truth_ld, t0, n_groups, n_output,
                                                                n_times, nrmse_fg_dev);
    XESN_CUDA_OK(cudaGetLastError());
    count_launch();
    cost_mean_kernel<<<(unsigned)((n_groups + 127) / 128), 128, 0, st>>>(nrmse_fg_dev, n_windows, n_groups, upper_bound,
                                                                       cost_dev);
    XESN_CUDA_OK(cudaGetLastError());
    count_launch();
    return 0;
}

extern "C" int xesn_cost_psd_nrmse(const double* pred_dev, const double* truth_dev, int64_t truth_ld, int64_t t0,
                                   int32_t n_windows, int32_t n_groups, int32_t n_output, int32_t n_times,
                                   double* psd_nrmse_fg_dev, void* stream) {
    XESN_REQUIRE(pred_dev && truth_dev && psd_nrmse_fg_dev, "null pointer");
    XESN_REQUIRE(n_windows > 0 && n_groups > 0 && n_output > 1 && n_times > 0 && truth_ld >= t0 + n_times, "bad sizes");
    const size_t smem = sizeof(double) * (2 * (size_t)n_output + 2 * (size_t)n_times);
    XESN_REQUIRE(smem <= 200 * 1024, "forecast too long for the spectrum buffers (n_times <= ~12000)");
    XESN_CUDA_OK(cudaFuncSetAttribute(psd_nrmse_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    psd_nrmse_kernel<<<(unsigned)(n_windows * n_groups), CT, smem, (cudaStream_t)stream>>>(
        pred_dev, truth_dev, truth_ld, t0, n_groups, n_output, n_times, psd_nrmse_fg_dev);
    XESN_CUDA_OK(cudaGetLastError());
    count_launch();
    return 0;
}

extern "C" int xesn_psd_radial(const double* spec_dev, int64_t n_images, int32_t n_z, int32_t n_pix, const int32_t* bin_ptr_dev,
                               const int32_t* pix_idx_dev, const double* factor_dev, int32_t n_bins, double* out_dev,
                               void* stream) {
    XESN_REQUIRE(spec_dev && bin_ptr_dev && pix_idx_dev && factor_dev && out_dev, "null pointer");
    XESN_REQUIRE(n_images > 0 && n_z > 0 && n_pix > 0 && n_bins > 0 && n_images < (1LL << 31), "bad sizes");
    psd_radial_kernel<<<(unsigned)n_images, CT, 0, (cudaStream_t)stream>>>(reinterpret_cast<const double2*>(spec_dev), n_z, n_pix,
                                                                          bin_ptr_dev, pix_idx_dev, factor_dev, n_bins, out_dev);
    XESN_CUDA_OK(cudaGetLastError());
    count_launch();
    return 0;
}

extern "C" int xesn_nrmse_skipnan(const double* pred_dev, const double* truth_dev, int32_t n_windows, int32_t n_groups,
                                  int32_t n_rows, int32_t n_times, double* nrmse_fg_dev, void* stream) {
    XESN_REQUIRE(pred_dev && truth_dev && nrmse_fg_dev, "null pointer");
    XESN_REQUIRE(n_windows > 0 && n_groups > 0 && n_rows > 0 && n_times > 0, "bad sizes");
    nrmse_skipnan_kernel<<<(unsigned)(n_windows * n_groups), CT, 0, (cudaStream_t)stream>>>(pred_dev, truth_dev, n_groups, n_rows,
                                                                                           n_times, nrmse_fg_dev);
    XESN_CUDA_OK(cudaGetLastError());
    count_launch();
    return 0;
}
