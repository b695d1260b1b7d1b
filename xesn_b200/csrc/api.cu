// extern "C" entry points of libxesn_b200.so (declared in include/xesn_b200.h).
#include <cstdarg>
#include <cstdlib>
#include <cstring>
#include <atomic>
#include <mutex>
#include <string>
#include <algorithm>
#include <vector>

#include "../../include/xesn_b200.h"
#include "common.cuh"
#include "kernels.cuh"
#include "gram_i8.cuh"
#include "yr_skinny.cuh"

namespace xesn {

void gram_i8_tile_list(int N, std::vector<int2>& tiles);
int launch_digitize(const double* states, long long lds, int n_rows, int N, int8_t* planes, long long ldp,
                    long long plane_stride, int rows_pad, cudaStream_t stream);
int launch_gram_i8(const gi8::GramI8Args& a, const int2* tiles, int ntiles, int num_sms, cudaStream_t stream);

static thread_local std::string g_error;
static std::atomic<long long> g_launches{0};

void set_error(const char* fmt, ...) {
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    g_error = buf;
}
void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

// ---- per-phase device timers: CUDA events recorded on the launching stream ----
struct ProfRec {
    int cls;
    cudaEvent_t a, b;
};
static bool g_prof_on = false;
static std::vector<ProfRec> g_prof;
static std::vector<cudaEvent_t> g_prof_pool;
static cudaEvent_t prof_event() {
    cudaEvent_t e;
    if (!g_prof_pool.empty()) {
        e = g_prof_pool.back();
        g_prof_pool.pop_back();
    } else {
        cudaEventCreate(&e);
    }
    return e;
}
void prof_begin(int cls, cudaStream_t st) {
    if (!g_prof_on) return;
    ProfRec r{cls, prof_event(), prof_event()};
    cudaEventRecord(r.a, st);
    g_prof.push_back(r);
}
void prof_end(int cls, cudaStream_t st) {
    if (!g_prof_on) return;
    for (size_t i = g_prof.size(); i-- > 0;)
        if (g_prof[i].cls == cls) {
            cudaEventRecord(g_prof[i].b, st);
            return;
        }
}

namespace {
__global__ void strided_copy_kernel(const double* __restrict__ src, double* __restrict__ dst, long long dst_stride,
                                    long long n) {
    long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (k < n) dst[k * dst_stride] = src[k];
}
inline long long round_up(long long x, long long m) { return (x + m - 1) / m * m; }
// makes `device` current for the scope and restores the caller's device afterwards
struct DeviceGuard {
    int prev = -1;
    explicit DeviceGuard(int device) {
        if (cudaGetDevice(&prev) != cudaSuccess) prev = -1;
        if (prev != device) cudaSetDevice(device);
        else prev = -1;
    }
    ~DeviceGuard() {
        if (prev >= 0) cudaSetDevice(prev);
    }
};
}  // namespace

}  // namespace xesn

using namespace xesn;

struct xesn_reservoir {
    int device = 0;
    int N = 0, I = 0;
    long long nnz = 0;
    double alpha = 0.0;
    int* d_indptr = nullptr;
    int* d_indices = nullptr;
    double* d_values = nullptr;
    double* d_win = nullptr;
    double* d_bias = nullptr;
    // per-group parameters (xesn_reservoir_set_group_params): d_values / d_win / d_bias then hold n_param_groups copies
    int n_param_groups = 0;
    long long values_gs = 0, win_gs = 0, bias_gs = 0;
    double* d_alpha_g = nullptr;
    int* d_status = nullptr;
    unsigned* d_bar = nullptr;  // [0..1] grid barrier words of the persistent forecast kernel, [2] hint counter of the pull recurrence
    int persistent_predict = 1;
    double* d_ro_partial = nullptr;
    int* d_ro_counters = nullptr;
    long long ro_partial_cap = 0, ro_counter_cap = 0;
    int cluster = 1;
    int max_slice_nnz = 0;
    int num_sms = 0;
    std::vector<int> h_indptr, h_indices;
    int2* d_tiles = nullptr;  // Gram tile order of the fused chunk kernel
    int n_tiles = 0;
    int2* d_tiles_i8 = nullptr;  // 128 x 64 tiles of the int8 tensor-core Gram role
    int n_tiles_i8 = 0;
    int gram_mode = 1;  // 0 = FP64 DMMA, 1 = fixed-point int8 digit planes on tcgen05 (absolute error <= K 2^-50 per entry)
    int fused_enabled = 1;
    // recurrence engine: 0 = team model everywhere (recur_body.cuh), 1 = pull model (pull_body.cuh) for forced runs and
    // spin-up, team model inside the fused training kernel (default: measured fastest), 2 = pull model everywhere
    int recur_pull = 1;
    // leak rates all inside [0, 1]: the recurrence is then a convex combination and |r| <= 1 holds, which the
    // fixed-point digit planes of the int8 Gram role rely on; otherwise the FP64 DMMA Gram is used
    bool leak_unit_range = true;
    // closed loop: NaN inputs read as 0 (LazyESN, xesn/lazyesn.py:362-364) instead of propagating (eager ESN)
    int mask_nan_inputs = 0;
    SolveCtx solve;  // side stream + events of the look-ahead Cholesky, per handle (= per device)
    int* d_flow_tail_ptr = nullptr;  // predict_flow.cu: offsets of the rows' tails, cached for one (team, upc)
    int flow_tail_team = 0, flow_tail_upc = 0;
    // recur_cluster.cuh: per team size, the offsets of the rows' tails and the shared-memory placement of the states
    // (bank_layout.cu: slot of every state, byte offset of the slot of every entry of W)
    struct ClusterPlan {
        int* d_tail_ptr = nullptr;
        int* d_slot = nullptr;
        int* d_col_off = nullptr;
    };
    ClusterPlan cl_plan[9];   // index = team size (1..8)
    // low-priority stream + fork/join events of the worker grid that runs beside the recurrence clusters
    cudaStream_t worker_stream = nullptr;
    cudaEvent_t ev_wfork = nullptr, ev_wjoin = nullptr;
    int* d_tile_counter = nullptr;   // Gram tiles of a chunk launch are drawn from this counter (gram_i8.cuh, DYN)
};
static inline int gram_mode_effective(const xesn_reservoir* h) { return (h->gram_mode == 1 && h->leak_unit_range) ? 1 : 0; }

extern "C" {

const char* xesn_last_error(void) { return g_error.c_str(); }
int xesn_abi_version(void) { return 1; }
int64_t xesn_launch_count(void) { return (int64_t)g_launches.load(); }

int64_t xesn_bank_layout(int32_t n_reservoir, const int32_t* indptr, const int32_t* indices, int32_t units_per_cta,
                         int32_t n_slots, int32_t* slot_out, int64_t* natural_out) {
    XESN_REQUIRE(indptr && slot_out && n_reservoir > 0 && units_per_cta > 0, "bad arguments");
    XESN_REQUIRE(n_slots >= n_reservoir && n_slots % 16 == 0, "n_slots: a multiple of 16, at least n_reservoir");
    XESN_REQUIRE(indptr[n_reservoir] == 0 || indices, "null CSR indices");
    long long before = 0;
    const long long after = plan_bank_layout(n_reservoir, indptr, indices, units_per_cta, recur::CL_KHEAD, recur::CL_KTAIL,
                                             n_slots, n_slots + 1, slot_out, &before);
    if (natural_out) *natural_out = before;
    return after;
}

int xesn_profile_enable(int on) {
    for (auto& r : g_prof) {
        g_prof_pool.push_back(r.a);
        g_prof_pool.push_back(r.b);
    }
    g_prof.clear();
    g_prof_on = on != 0;
    return 0;
}

int xesn_profile_read(int cls, int64_t* n_regions, double* total_ms) {
    XESN_REQUIRE(n_regions && total_ms && cls >= 0 && cls < PROF_N, "bad arguments");
    long long n = 0;
    double ms = 0.0;
    for (auto& r : g_prof) {
        if (r.cls != cls) continue;
        XESN_CUDA_OK(cudaEventSynchronize(r.b));
        float t = 0.f;
        XESN_CUDA_OK(cudaEventElapsedTime(&t, r.a, r.b));
        ms += t, ++n;
    }
    *n_regions = n, *total_ms = ms;
    return 0;
}

int64_t xesn_profile_regions(int32_t cls, double* ms_out, int64_t capacity) {
    XESN_REQUIRE(cls >= 0 && cls < PROF_N && (ms_out || capacity == 0), "bad arguments");
    long long n = 0;
    for (auto& r : g_prof) {
        if (r.cls != cls) continue;
        if (n < capacity) {
            XESN_CUDA_OK(cudaEventSynchronize(r.b));
            float t = 0.f;
            XESN_CUDA_OK(cudaEventElapsedTime(&t, r.a, r.b));
            ms_out[n] = t;
        }
        ++n;
    }
    return n;
}

int xesn_reservoir_create(xesn_reservoir** out, int device, int32_t N, int32_t I, const int32_t* indptr,
                          const int32_t* indices, const double* values, const double* win, const double* bias,
                          double leak_rate) {
    XESN_REQUIRE(out && indptr && win && bias, "null pointer");
    XESN_REQUIRE(N > 0 && I > 0, "n_reservoir and n_input must be positive");
    XESN_CUDA_OK(cudaSetDevice(device));
    cudaDeviceProp prop;
    XESN_CUDA_OK(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10) {
        set_error("xesn_b200 is built for sm_100a (Blackwell B200) only; device %d is sm_%d%d", device, prop.major,
                  prop.minor);
        return -2;
    }
    const long long nnz = indptr[N];
    XESN_REQUIRE(nnz >= 0 && indptr[0] == 0, "malformed CSR indptr");
    XESN_REQUIRE(nnz == 0 || (indices && values), "null CSR arrays");
    for (int i = 0; i < N; ++i) XESN_REQUIRE(indptr[i + 1] >= indptr[i], "CSR indptr must be non-decreasing");
    for (long long k = 0; k < nnz; ++k) XESN_REQUIRE(indices[k] >= 0 && indices[k] < N, "CSR column index out of range");

    auto* h = new xesn_reservoir();
    h->device = device, h->N = N, h->I = I, h->nnz = nnz, h->alpha = leak_rate;
    h->leak_unit_range = leak_rate >= 0.0 && leak_rate <= 1.0;
    XESN_CUDA_OK(cudaMalloc(&h->d_indptr, sizeof(int) * (N + 1)));
    XESN_CUDA_OK(cudaMalloc(&h->d_indices, sizeof(int) * (nnz > 0 ? nnz : 1)));
    XESN_CUDA_OK(cudaMalloc(&h->d_values, sizeof(double) * (nnz > 0 ? nnz : 1)));
    XESN_CUDA_OK(cudaMalloc(&h->d_win, sizeof(double) * (size_t)N * I));
    XESN_CUDA_OK(cudaMalloc(&h->d_bias, sizeof(double) * N));
    // d_status[0] = flag (0 ok, 1 poll timed out, 2 grid/peer barrier timed out), d_status[1] = poll limit in microseconds
    XESN_CUDA_OK(cudaMalloc(&h->d_status, 2 * sizeof(int)));
    {
        int init[2] = {0, 2000000};
        if (const char* e = getenv("XESN_B200_POLL_TIMEOUT_MS")) init[1] = std::max(1, atoi(e)) * 1000;
        XESN_CUDA_OK(cudaMemcpy(h->d_status, init, sizeof(init), cudaMemcpyHostToDevice));
    }
    XESN_CUDA_OK(cudaMalloc(&h->d_bar, 4 * sizeof(unsigned)));
    XESN_CUDA_OK(cudaMemset(h->d_bar, 0, 4 * sizeof(unsigned)));
    // forecast engine: 1 = barrier-free pull kernel for a single reservoir (else the barrier kernel), 2 = always the
    // persistent kernel with grid barriers, 0 = one launch per phase and step
    if (const char* e = getenv("XESN_B200_PREDICT"))
        h->persistent_predict = strcmp(e, "steps") == 0 ? 0 : (strcmp(e, "barrier") == 0 ? 2 : 1);
    if (const char* e = getenv("XESN_B200_RECUR"))
        h->recur_pull = strcmp(e, "pull") == 0 ? 2 : (strcmp(e, "team") == 0 ? 0 : 1);
    XESN_CUDA_OK(cudaMemcpy(h->d_indptr, indptr, sizeof(int) * (N + 1), cudaMemcpyHostToDevice));
    if (nnz) {
        XESN_CUDA_OK(cudaMemcpy(h->d_indices, indices, sizeof(int) * nnz, cudaMemcpyHostToDevice));
        XESN_CUDA_OK(cudaMemcpy(h->d_values, values, sizeof(double) * nnz, cudaMemcpyHostToDevice));
    }
    XESN_CUDA_OK(cudaMemcpy(h->d_win, win, sizeof(double) * (size_t)N * I, cudaMemcpyHostToDevice));
    XESN_CUDA_OK(cudaMemcpy(h->d_bias, bias, sizeof(double) * N, cudaMemcpyHostToDevice));

    h->cluster = recur_cluster_size(N);
    const int upc = recur_units_per_cta(N, h->cluster);
    int mx = 0;
    for (int c = 0; c < h->cluster; ++c) {
        int lo = c * upc < N ? c * upc : N, hi = (c + 1) * upc < N ? (c + 1) * upc : N;
        int s = indptr[hi] - indptr[lo];
        if (s > mx) mx = s;
    }
    h->max_slice_nnz = mx;
    h->num_sms = prop.multiProcessorCount;
    h->h_indptr.assign(indptr, indptr + N + 1);
    h->h_indices.assign(indices, indices + nnz);
    {
        // lower-triangle tile list in super-rows of 12 tile rows, column by column (L2 locality)
        const int T = (N + 127) / 128, H = 12;
        std::vector<int2> tiles;
        tiles.reserve((size_t)T * (T + 1) / 2);
        for (int sb = 0; sb < T; sb += H)
            for (int tj = 0; tj < (sb + H < T ? sb + H : T); ++tj)
                for (int ti = (sb > tj ? sb : tj); ti < (sb + H < T ? sb + H : T); ++ti) tiles.push_back(make_int2(ti, tj));
        h->n_tiles = (int)tiles.size();
        XESN_CUDA_OK(cudaMalloc(&h->d_tiles, sizeof(int2) * tiles.size()));
        XESN_CUDA_OK(cudaMemcpy(h->d_tiles, tiles.data(), sizeof(int2) * tiles.size(), cudaMemcpyHostToDevice));
        std::vector<int2> t8;
        gram_i8_tile_list(N, t8);
        h->n_tiles_i8 = (int)t8.size();
        XESN_CUDA_OK(cudaMalloc(&h->d_tiles_i8, sizeof(int2) * t8.size()));
        XESN_CUDA_OK(cudaMemcpy(h->d_tiles_i8, t8.data(), sizeof(int2) * t8.size(), cudaMemcpyHostToDevice));
    }
    *out = h;
    return 0;
}

int xesn_reservoir_destroy(xesn_reservoir* h) {
    if (!h) return 0;
    cudaSetDevice(h->device);
    cudaFree(h->d_indptr);
    cudaFree(h->d_indices);
    cudaFree(h->d_values);
    cudaFree(h->d_win);
    cudaFree(h->d_bias);
    cudaFree(h->d_alpha_g);
    cudaFree(h->d_status);
    cudaFree(h->d_bar);
    cudaFree(h->d_tiles);
    cudaFree(h->d_tiles_i8);
    cudaFree(h->d_ro_partial);
    cudaFree(h->d_ro_counters);
    solve_ctx_destroy(h->solve);
    cudaFree(h->d_flow_tail_ptr);
    for (auto& pl : h->cl_plan) cudaFree(pl.d_tail_ptr), cudaFree(pl.d_slot), cudaFree(pl.d_col_off);
    cudaFree(h->d_tile_counter);
    if (h->worker_stream) cudaStreamDestroy(h->worker_stream);
    if (h->ev_wfork) cudaEventDestroy(h->ev_wfork);
    if (h->ev_wjoin) cudaEventDestroy(h->ev_wjoin);
    delete h;
    return 0;
}

// Per-group parameters: G independent reservoirs that share the sparsity pattern of W but have their own values of W,
// Win, bias and leak rate (candidates of a hyper-parameter search, reference cost.py:190-203: the same seeds, other
// scaling factors).  Host arrays values[G][nnz], win[G][N*I], bias[G][N], leak[G], G >= 1; they replace the arrays given
// to xesn_reservoir_create.  Afterwards n_groups of every entry point must be a multiple of G: group g uses set g % G.
int xesn_reservoir_set_group_params(xesn_reservoir* h, int32_t G, const double* values, const double* win,
                                    const double* bias, const double* leak) {
    XESN_REQUIRE(h && G >= 1 && values && win && bias && leak, "bad arguments");
    XESN_CUDA_OK(cudaSetDevice(h->device));
    const size_t nnz = (size_t)(h->nnz > 0 ? h->nnz : 1), nw = (size_t)h->N * h->I, nb = (size_t)h->N;
    cudaFree(h->d_values), cudaFree(h->d_win), cudaFree(h->d_bias), cudaFree(h->d_alpha_g);
    h->d_values = h->d_win = h->d_bias = h->d_alpha_g = nullptr;
    XESN_CUDA_OK(cudaMalloc(&h->d_values, sizeof(double) * nnz * G));
    XESN_CUDA_OK(cudaMalloc(&h->d_win, sizeof(double) * nw * G));
    XESN_CUDA_OK(cudaMalloc(&h->d_bias, sizeof(double) * nb * G));
    XESN_CUDA_OK(cudaMalloc(&h->d_alpha_g, sizeof(double) * G));
    if (h->nnz > 0) XESN_CUDA_OK(cudaMemcpy(h->d_values, values, sizeof(double) * (size_t)h->nnz * G, cudaMemcpyHostToDevice));
    XESN_CUDA_OK(cudaMemcpy(h->d_win, win, sizeof(double) * nw * G, cudaMemcpyHostToDevice));
    XESN_CUDA_OK(cudaMemcpy(h->d_bias, bias, sizeof(double) * nb * G, cudaMemcpyHostToDevice));
    XESN_CUDA_OK(cudaMemcpy(h->d_alpha_g, leak, sizeof(double) * G, cudaMemcpyHostToDevice));
    h->n_param_groups = G;
    h->values_gs = (long long)h->nnz, h->win_gs = (long long)nw, h->bias_gs = (long long)nb;
    h->alpha = leak[0];
    h->leak_unit_range = true;
    for (int g = 0; g < G; ++g) h->leak_unit_range = h->leak_unit_range && leak[g] >= 0.0 && leak[g] <= 1.0;
    return 0;
}
#define XESN_REQUIRE_GROUPS(h, G) \
    XESN_REQUIRE((h)->n_param_groups == 0 || (G) % (h)->n_param_groups == 0, "n_groups must be a multiple of the number of parameter sets")

int xesn_reservoir_status(xesn_reservoir* h, int32_t* status_host) {
    XESN_REQUIRE(h && status_host, "null pointer");
    DeviceGuard guard(h->device);
    XESN_CUDA_OK(cudaMemcpy(status_host, h->d_status, sizeof(int), cudaMemcpyDeviceToHost));
    // read-and-clear: a timed-out run is reported once; later runs on this handle start clean
    if (*status_host) XESN_CUDA_OK(cudaMemset(h->d_status, 0, sizeof(int)));
    return 0;
}

int xesn_set_poll_timeout_ms(xesn_reservoir* h, int32_t ms) {
    XESN_REQUIRE(h && ms > 0, "bad arguments");
    DeviceGuard guard(h->device);
    const int us = ms > 2000000 ? 2000000000 : ms * 1000;
    XESN_CUDA_OK(cudaMemcpy(h->d_status + 1, &us, sizeof(int), cudaMemcpyHostToDevice));
    return 0;
}

int xesn_set_nan_input_mask(xesn_reservoir* h, int32_t on) {
    XESN_REQUIRE(h, "null pointer");
    h->mask_nan_inputs = on != 0;
    return 0;
}

int xesn_get_gram_mode(xesn_reservoir* h, int32_t* mode_out) {
    XESN_REQUIRE(h && mode_out, "null pointer");
    *mode_out = gram_mode_effective(h);
    return 0;
}

int xesn_update(xesn_reservoir* h, const double* r_in, const double* u, double* r_out, int32_t G, void* stream) {
    XESN_REQUIRE(h && r_in && u && r_out && G > 0, "bad arguments");
    XESN_REQUIRE(r_in != r_out, "r_in and r_out must not alias");
    XESN_REQUIRE_GROUPS(h, G);
    StepArgs a = {};
    a.indptr = h->d_indptr, a.indices = h->d_indices, a.values = h->d_values;
    a.Win = h->d_win, a.bias = h->d_bias, a.u = u, a.drive = nullptr;
    a.values_gs = h->values_gs, a.win_gs = h->win_gs, a.bias_gs = h->bias_gs, a.alpha_g = h->d_alpha_g;
    a.param_groups = h->n_param_groups;
    a.r_in = r_in, a.r_out = r_out, a.N = h->N, a.I = h->I, a.alpha = h->alpha;
    a.mask_nan = h->mask_nan_inputs;
    return launch_step_update(a, G, (cudaStream_t)stream);
}

// ---------------------------------------------------------------------------------------------
// drive for a chunk: D[g][t][n] = sum_i U_g[i][t0+t] Win[n][i] + b[n]
// ---------------------------------------------------------------------------------------------
static int chunk_inputs(xesn_reservoir* h, const xesn_series* s, int rows, int G, long long t0, int m, double* stage,
                        long long stage_ld, const double** base, long long* ld, long long* gstride,
                        cudaStream_t st) {
    if (s->map_dev || s->zero_row_dev) {
        ProfScope prof(PROF_GATHER, st);
        int rc = launch_gather_series(s->data_dev, s->ld, t0, s->map_dev, s->fill_dev, s->zero_row_dev, stage, stage_ld,
                                      (long long)G * rows, m, st);
        if (rc) return rc;
        *base = stage, *ld = stage_ld, *gstride = (long long)rows * stage_ld;
    } else {
        *base = s->data_dev + t0, *ld = s->ld, *gstride = (long long)rows * s->ld;
    }
    return 0;
}

static int drive_gemm(xesn_reservoir* h, const double* ubase, long long ldu, long long ugstride, int G, int m, double* D,
                      cudaStream_t st) {
    ProfScope prof(PROF_DRIVE, st);
    GemmArgs g = {};
    g.A = ubase, g.lda = ldu, g.strideA = ugstride;           // A(t,i) = U[i*ld + t]   (MN-contiguous)
    g.B = h->d_win, g.ldb = h->I, g.strideB = h->win_gs;      // B(i,n) = Win[n*I + i]  (K-contiguous)
    g.C = D, g.ldc = h->N, g.strideC = (long long)m * h->N;
    g.bias = h->d_bias, g.strideBias = h->bias_gs, g.modB = h->n_param_groups;
    g.M = m, g.N = h->N, g.K = h->I, g.batch = G;
    g.alpha = 1.0, g.beta = 0.0, g.flags = GEMM_B_KMAJOR;
    return launch_gemm_f64(g, st);
}

// tables of the cluster recurrence for one team size, built on first use
static int ensure_cluster_plan(xesn_reservoir* h, int team) {
    XESN_REQUIRE(team >= 1 && team <= 8, "cluster team size");
    auto& pl = h->cl_plan[team];
    if (pl.d_tail_ptr) return 0;
    const int N = h->N, upc = recur_units_per_cta(N, team);
    std::vector<int> tp((size_t)team * (upc + 1));
    cluster_tail_ptr(N, team, h->h_indptr.data(), tp.data());
    // placement of the state copies: banks chosen against the gathers of this geometry
    std::vector<int> slot(N), col_off((size_t)std::max<long long>(h->nnz, 1));
    const int n_slots = cluster_state_slots(N);
    if (getenv("XESN_B200_NATURAL_SLOTS")) {   // A/B measurements
        for (int j = 0; j < N; ++j) slot[j] = j;
    } else {
        plan_bank_layout(N, h->h_indptr.data(), h->h_indices.data(), upc, recur::CL_KHEAD, recur::CL_KTAIL, n_slots,
                         n_slots + 1, slot.data(), nullptr);
    }
    for (long long k = 0; k < h->nnz; ++k) col_off[k] = slot[h->h_indices[k]] * (int)sizeof(double);
    XESN_CUDA_OK(cudaMalloc(&pl.d_slot, sizeof(int) * slot.size()));
    XESN_CUDA_OK(cudaMalloc(&pl.d_col_off, sizeof(int) * col_off.size()));
    XESN_CUDA_OK(cudaMemcpy(pl.d_slot, slot.data(), sizeof(int) * slot.size(), cudaMemcpyHostToDevice));
    XESN_CUDA_OK(cudaMemcpy(pl.d_col_off, col_off.data(), sizeof(int) * col_off.size(), cudaMemcpyHostToDevice));
    XESN_CUDA_OK(cudaMalloc(&pl.d_tail_ptr, sizeof(int) * tp.size()));
    XESN_CUDA_OK(cudaMemcpy(pl.d_tail_ptr, tp.data(), sizeof(int) * tp.size(), cudaMemcpyHostToDevice));
    return 0;
}

// teacher-forced recurrence without a Gram role (spin-up, forced runs)
static int forced_launch(xesn_reservoir* h, const RecurArgs& r, int G, cudaStream_t st) {
    // small reservoirs: clusters exchanging states through distributed shared memory, as thin as the machine allows
    // (no Gram role to leave SMs to); needs |r| bookkeeping of nothing but the ring, so every Gram mode can use it
    int cteam = 0, ctail = 0;
    if (h->recur_pull != 2 && cluster_forced_geometry(h->N, G, h->h_indptr.data(), h->num_sms, &cteam, &ctail)) {
        if (int rc = ensure_cluster_plan(h, cteam)) return rc;
        RecurArgs rc_args = r;
        rc_args.slot = h->cl_plan[cteam].d_slot, rc_args.col_off = h->cl_plan[cteam].d_col_off;
        rc_args.planes = nullptr;
        return launch_fused_cluster_i8(rc_args, G, cteam, h->cl_plan[cteam].d_tail_ptr, ctail, gi8::GramI8Args{}, nullptr, 0,
                                       GemmArgs{}, drv::DriveArgs{}, h->num_sms, st);
    }
    int upc = 0, upt = 0;
    if (h->recur_pull && pull_geometry(h->N, G, h->num_sms, &upc, &upt) == 0) return launch_pull_forced(r, G, h->num_sms, st);
    return launch_recur_forced(r, G, h->max_slice_nnz, st);
}

static RecurArgs recur_args(xesn_reservoir* h) {
    RecurArgs r = {};
    r.indptr = h->d_indptr, r.indices = h->d_indices, r.values = h->d_values;
    r.values_gs = h->values_gs, r.alpha_g = h->d_alpha_g, r.param_groups = h->n_param_groups;
    r.N = h->N, r.alpha = h->alpha, r.smem_nnz = h->max_slice_nnz, r.status = h->d_status;
    r.stagger = getenv("XESN_B200_NO_STAGGER") ? 0 : 50;
    if (const char* e = getenv("XESN_B200_STAGGER_PCT")) r.stagger = atoi(e);   // experiments
    r.debug_no_exchange = getenv("XESN_B200_DEBUG_NO_EXCHANGE") != nullptr;
    r.hint = getenv("XESN_B200_NO_HINT") ? nullptr : h->d_bar + 2;
    if (const char* e = getenv("XESN_B200_POLL_SLEEP1")) r.poll_sleep_retry = atoi(e);
    return r;
}

constexpr int FORCED_CHUNK = 512;

int64_t xesn_forced_scratch_bytes(xesn_reservoir* h, int32_t G, int32_t n_steps, int32_t gathered) {
    if (!h) return -1;
    long long c = n_steps < FORCED_CHUNK ? (n_steps > 0 ? n_steps : 1) : FORCED_CHUNK;
    long long cld = round_up(c, 2);
    long long d = (long long)G * c * h->N + (long long)G * (c + 1) * h->N;
    if (gathered) d += (long long)G * h->I * cld;
    return d * 8 + 256;
}

int xesn_forced_run(xesn_reservoir* h, const xesn_series* u, int64_t t0, int32_t n_steps, int32_t G, double* carry,
                    double* states, void* scratch, int64_t scratch_bytes, void* stream) {
    XESN_REQUIRE(h && u && carry && G > 0 && n_steps >= 0, "bad arguments");
    XESN_REQUIRE_GROUPS(h, G);
    const int gathered = (u->map_dev || u->zero_row_dev) ? 1 : 0;
    XESN_REQUIRE(scratch && scratch_bytes >= xesn_forced_scratch_bytes(h, G, n_steps, gathered), "scratch too small");
    cudaStream_t st = (cudaStream_t)stream;
    const int N = h->N;
    const long long c = n_steps < FORCED_CHUNK ? (n_steps > 0 ? n_steps : 1) : FORCED_CHUNK;
    const long long cld = round_up(c, 2);
    double* D = (double*)scratch;
    double* pp = D + (long long)G * c * N;
    double* Uc = pp + (long long)G * (c + 1) * N;
    if (states && n_steps == 0) {
        XESN_CUDA_OK(cudaMemcpyAsync(states, carry, sizeof(double) * (size_t)G * N, cudaMemcpyDeviceToDevice, st));
    }
    for (long long s0 = 0; s0 < n_steps; s0 += c) {
        const int m = (int)(n_steps - s0 < c ? n_steps - s0 : c);
        const double* ub;
        long long ldu, ugs;
        int rc = chunk_inputs(h, u, h->I, G, t0 + s0, m, Uc, cld, &ub, &ldu, &ugs, st);
        if (rc) return rc;
        rc = drive_gemm(h, ub, ldu, ugs, G, m, D, st);
        if (rc) return rc;
        RecurArgs r = recur_args(h);
        r.drive = D, r.drive_group_stride = (long long)m * N;
        r.steps = m;
        r.carry = carry;
        if (states) {
            r.states = states + s0 * N, r.state_group_stride = (long long)(n_steps + 1) * N;
        } else {
            r.states = pp, r.state_group_stride = (long long)(m + 1) * N;
        }
        rc = forced_launch(h, r, G, st);
        if (rc) return rc;
    }
    return 0;
}

// ---------------------------------------------------------------------------------------------
// training
// ---------------------------------------------------------------------------------------------
struct TrainLayout {
    long long cld;
    long long carry, Uc, Yc, D, R, D2, R2, P1, P2, total;  // offsets in doubles
    long long ldp, plane_stride;                          // digit planes (bytes)
};
static TrainLayout train_layout(xesn_reservoir* h, int O, int G, int chunk, int gu, int gy) {
    TrainLayout L;
    L.cld = round_up(chunk, 2);
    long long off = 0;
    L.carry = off, off += round_up((long long)G * h->N, 2);
    L.Uc = off, off += gu ? (long long)G * h->I * L.cld : 0;
    L.Yc = off, off += gy ? (long long)G * O * L.cld : 0;
    L.D = off, off += round_up((long long)G * chunk * h->N, 2);
    L.R = off, off += round_up((long long)G * (chunk + 1) * h->N, 2);
    // second drive / state-ring buffers: the fused chunk kernel runs recurrence(c) beside Gram(c-1)
    L.D2 = off, off += round_up((long long)G * chunk * h->N, 2);
    L.R2 = off, off += round_up((long long)G * (chunk + 1) * h->N, 2);
    // int8 digit planes of the states, [G][7][round_up(chunk,32)][round_up(N,128)] bytes, double-buffered
    L.ldp = round_up(h->N, 128);
    L.plane_stride = L.ldp * round_up(chunk, gi8::KB);
    const long long plane_doubles = round_up((long long)G * gi8::NDIG * L.plane_stride, 16) / 8;
    L.P1 = off, off += plane_doubles;
    L.P2 = off, off += plane_doubles;
    L.total = off;
    return L;
}

int64_t xesn_train_scratch_bytes(xesn_reservoir* h, int32_t O, int32_t G, int32_t chunk, int32_t gu, int32_t gy) {
    if (!h || chunk <= 0) return -1;
    long long train = train_layout(h, O, G, chunk, gu, gy).total;
    long long solve = std::max(cholesky_workspace_doubles(h->N, O, G), lu_workspace_doubles(h->N));
    return (train > solve ? train : solve) * 8 + 256;
}

// Group strides of the Gram matrices and readouts.  Separate arrays: [G][N][N] and [G][O][N].  When the readout rows
// start right below the first Gram matrix (wout == gram + N*N) the two are ONE array [G][N+O][N]: the solve then carries
// the right-hand sides through the factorisation as extra rows and needs no forward substitution (cholesky.cu, FOLD).
static inline bool gram_interleaved(const double* gram, const double* wout, int N) { return wout == gram + (long long)N * N; }

int xesn_train_accumulate(xesn_reservoir* h, const xesn_series* u, const xesn_series* y, int32_t O, int32_t G,
                          int64_t T, int64_t n_spinup, int32_t chunk, double* gram, double* wout, void* scratch,
                          int64_t scratch_bytes, void* stream) {
    XESN_REQUIRE(h && u && y && gram && wout && scratch, "null pointer");
    XESN_REQUIRE(O > 0 && G > 0 && chunk > 0 && T >= 0 && n_spinup >= 0 && n_spinup <= T, "bad sizes");
    XESN_REQUIRE_GROUPS(h, G);
    const int gu = (u->map_dev || u->zero_row_dev) ? 1 : 0, gy = (y->map_dev || y->zero_row_dev) ? 1 : 0;
    XESN_REQUIRE(scratch_bytes >= xesn_train_scratch_bytes(h, O, G, chunk, gu, gy), "scratch too small");
    XESN_REQUIRE(chunk <= gi8::MAX_K, "chunk: at most 16384 time steps per chunk (int32 accumulators of the int8 Gram)");
    cudaStream_t st = (cudaStream_t)stream;
    const int N = h->N;
    const TrainLayout L = train_layout(h, O, G, chunk, gu, gy);
    double* base = (double*)scratch;
    double *carry = base + L.carry, *Uc = base + L.Uc, *Yc = base + L.Yc, *D = base + L.D, *R = base + L.R;

    XESN_CUDA_OK(cudaMemsetAsync(carry, 0, sizeof(double) * (size_t)G * N, st));
    const bool inter = gram_interleaved(gram, wout, N);
    const long long gram_gs = inter ? (long long)(N + O) * N : (long long)N * N;
    const long long wout_gs = inter ? (long long)(N + O) * N : (long long)O * N;
    if (inter) {
        XESN_CUDA_OK(cudaMemsetAsync(gram, 0, sizeof(double) * (size_t)G * gram_gs, st));
    } else {
        XESN_CUDA_OK(cudaMemsetAsync(gram, 0, sizeof(double) * (size_t)G * N * N, st));
        XESN_CUDA_OK(cudaMemsetAsync(wout, 0, sizeof(double) * (size_t)G * O * N, st));
    }

    // spin-up (states not stored): xesn/esn.py:498-499
    for (long long t0 = 0; t0 < n_spinup; t0 += chunk) {
        const int m = (int)(n_spinup - t0 < chunk ? n_spinup - t0 : chunk);
        const double* ub;
        long long ldu, ugs;
        int rc = chunk_inputs(h, u, h->I, G, t0, m, Uc, L.cld, &ub, &ldu, &ugs, st);
        if (rc) return rc;
        if ((rc = drive_gemm(h, ub, ldu, ugs, G, m, D, st))) return rc;
        RecurArgs r = recur_args(h);
        r.drive = D, r.drive_group_stride = (long long)m * N, r.steps = m, r.carry = carry;
        r.states = R, r.state_group_stride = (long long)(m + 1) * N;
        if ((rc = forced_launch(h, r, G, st))) return rc;
    }
    // accumulate: xesn/esn.py:502-513
    int team = 0, upt = 0, halves = 1;
    int pull_upc = 0, pull_upt = 0;
    const bool pull = h->recur_pull == 2 && gram_mode_effective(h) == 1 && pull_geometry(N, G, h->num_sms, &pull_upc, &pull_upt) == 0;
    // small reservoirs: the team of a reservoir is a thread-block cluster that exchanges states through distributed
    // shared memory (recur_cluster.cuh); XESN_B200_REC_CLUSTER=0 keeps the L2-polling teams
    int cteam = 0, ctail = 0;
    const bool cluster = !pull && gram_mode_effective(h) == 1 &&
                         cluster_recur_geometry(N, G, h->h_indptr.data(), h->num_sms, &cteam, &ctail);
    if (cluster)
        if (int rc = ensure_cluster_plan(h, cteam)) return rc;
    if (!h->d_tile_counter) XESN_CUDA_OK(cudaMalloc(&h->d_tile_counter, 64));
    const bool split = cluster && !getenv("XESN_B200_CLUSTER_SINGLE");
    if (split && !h->worker_stream) {
        int lo = 0, hi = 0;
        XESN_CUDA_OK(cudaDeviceGetStreamPriorityRange(&lo, &hi));
        XESN_CUDA_OK(cudaStreamCreateWithPriority(&h->worker_stream, cudaStreamNonBlocking, lo));
        XESN_CUDA_OK(cudaEventCreateWithFlags(&h->ev_wfork, cudaEventDisableTiming));
        XESN_CUDA_OK(cudaEventCreateWithFlags(&h->ev_wjoin, cudaEventDisableTiming));
    }
    const bool fuse = h->fused_enabled && (N % 2 == 0) && T > n_spinup &&
                      (pull || cluster || (fused_recur_geometry(N, G, fused_threads(gram_mode_effective(h) == 1), &team, &upt,
                                                     gram_mode_effective(h) == 1 ? &halves : nullptr) == 0 &&
                                (long long)team * G / halves < h->num_sms));
    if (fuse) {
        // software pipeline over chunks: launch c runs recurrence(c) || Gram(c-1) in one cooperative kernel
        const int upc = (pull || cluster) ? N : recur_units_per_cta(N, team);
        int slice_nnz = 0;
        for (int c = 0; c < ((pull || cluster) ? 0 : team); ++c) {
            const int lo = c * upc < N ? c * upc : N, hi = (c + 1) * upc < N ? (c + 1) * upc : N;
            slice_nnz = std::max(slice_nnz, h->h_indptr[hi] - h->h_indptr[lo]);
        }
        double* Dbuf[2] = {D, base + L.D2};
        double* Rbuf[2] = {R, base + L.R2};
        const bool i8 = gram_mode_effective(h) == 1;
        signed char* Pbuf[2] = {reinterpret_cast<signed char*>(base + L.P1), reinterpret_cast<signed char*>(base + L.P2)};
        const long long plane_group = (long long)gi8::NDIG * L.plane_stride;
        if (i8 && L.ldp > N) {  // padding columns must read as zero digits (rows of every plane and group are ldp apart)
            const size_t rows = (size_t)(G * plane_group / L.ldp);
            XESN_CUDA_OK(cudaMemset2DAsync(Pbuf[0] + N, (size_t)L.ldp, 0, (size_t)(L.ldp - N), rows, st));
            XESN_CUDA_OK(cudaMemset2DAsync(Pbuf[1] + N, (size_t)L.ldp, 0, (size_t)(L.ldp - N), rows, st));
        }
        // chunks of equal length (a multiple of the Gram role's K-block): a short last chunk would still cost a launch
        // as long as the Gram pass of the full chunk before it.  (A ramp of short first chunks was measured too: the
        // first launch has no Gram pass beside it, but the recurrence runs in it all the same -- no gain.)
        std::vector<long long> sched_t0;
        std::vector<int> sched_m;
        {
            long long t = n_spinup, left = T - n_spinup;
            const long long n_eq = (left + chunk - 1) / chunk;
            const long long m_eq = n_eq ? std::min<long long>(chunk, round_up((left + n_eq - 1) / n_eq, gi8::KB)) : 0;
            for (long long k = 0; k < n_eq && left > 0; ++k) {
                const long long m = std::min(m_eq, left);
                sched_t0.push_back(t), sched_m.push_back((int)m), t += m, left -= m;
            }
        }
        const long long nch = (long long)sched_m.size();
        auto chunk_t0 = [&](long long c) { return sched_t0[(size_t)c]; };
        auto chunk_m = [&](long long c) { return sched_m[(size_t)c]; };
        auto drive_for = [&](long long c) -> int {
            const double* ub;
            long long ldu, ugs;
            int rc = chunk_inputs(h, u, h->I, G, chunk_t0(c), chunk_m(c), Uc, L.cld, &ub, &ldu, &ugs, st);
            if (rc) return rc;
            return drive_gemm(h, ub, ldu, ugs, G, chunk_m(c), Dbuf[c & 1], st);
        };
        int rc = drive_for(0);
        if (rc) return rc;
        // few inputs: the drive of chunk c+1 is computed by the Gram CTAs of launch c (yr_skinny.cuh) instead of a
        // GEMM launch of its own between the fused launches
        // ... when the Gram role has time to spare: small reservoirs, whose chunk is bounded by the serial recurrence
        // (measured: 16 x N=2000 gains 3 ms of 57; N=16000, bounded by the Gram role, loses 6 of 125)
        const bool drive_in_kernel = i8 && !pull && (cluster || upt == 1) && h->I <= drv::MAX_INPUTS && (double)G * N * N <= 1.0e8 &&
                                     !getenv("XESN_B200_DRIVE_SEPARATE");
        for (long long c = 0; c <= nch; ++c) {
            drv::DriveArgs dv = {};
            if (c + 1 < nch) {
                if (drive_in_kernel) {
                    const double* ub;
                    long long ldu, ugs;
                    if ((rc = chunk_inputs(h, u, h->I, G, chunk_t0(c + 1), chunk_m(c + 1), Uc, L.cld, &ub, &ldu, &ugs, st))) return rc;
                    dv.U = ub, dv.ldu = ldu, dv.ugs = ugs;
                    dv.Win = h->d_win, dv.bias = h->d_bias, dv.win_gs = h->win_gs, dv.bias_gs = h->bias_gs;
                    dv.param_groups = h->n_param_groups, dv.D = Dbuf[(c + 1) & 1];
                    dv.m = chunk_m(c + 1), dv.N = N, dv.I = h->I, dv.G = G;
                } else if ((rc = drive_for(c + 1))) {
                    return rc;
                }
            }
            RecurArgs r = recur_args(h);
            if (cluster) r.slot = h->cl_plan[cteam].d_slot, r.col_off = h->cl_plan[cteam].d_col_off;
            if (c < nch) {
                const int m = chunk_m(c);
                r.drive = Dbuf[c & 1], r.drive_group_stride = (long long)m * N, r.steps = m, r.carry = carry;
                r.states = Rbuf[c & 1], r.state_group_stride = (long long)(m + 1) * N;
            }
            if (i8 && c < nch) {
                const int m = chunk_m(c);
                const int mpad = (int)round_up(m, gi8::KB);
                r.planes = Pbuf[c & 1], r.ldp = L.ldp, r.plane_stride = L.plane_stride, r.plane_group_stride = plane_group;
                if (mpad > m)  // rows of the last K-block beyond the chunk read as zero
                    XESN_CUDA_OK(cudaMemset2DAsync(Pbuf[c & 1] + (long long)m * L.ldp, (size_t)L.plane_stride, 0,
                                                   (size_t)(mpad - m) * L.ldp, (size_t)gi8::NDIG * G, st));
            }
            GemmArgs g = {};
            gi8::GramI8Args g8 = {};
            int ntiles = 0;
            if (c >= 1) {
                const int mp = chunk_m(c - 1);
                const long long rgs = (long long)(mp + 1) * N;
                double* Rp = Rbuf[(c - 1) & 1];
                g.A = Rp, g.lda = N, g.strideA = rgs;
                g.B = Rp, g.ldb = N, g.strideB = rgs;
                g.C = gram, g.ldc = N, g.strideC = gram_gs;
                g.M = N, g.N = N, g.K = mp, g.batch = G, g.alpha = 1.0, g.beta = 1.0, g.flags = GEMM_LOWER;
                g8.planes = reinterpret_cast<const int8_t*>(Pbuf[(c - 1) & 1]);
                g8.ldp = L.ldp, g8.plane_stride = L.plane_stride, g8.group_stride = plane_group;
                g8.K = (int)round_up(mp, gi8::KB), g8.N = N, g8.G = gram, g8.ldg = N, g8.strideG = gram_gs;
                g8.batch = G;
                ntiles = i8 ? h->n_tiles_i8 : h->n_tiles;
            }
            // ybar += Y R of chunk c-1: skinny targets ride in the Gram role of the int8 fused kernel, else a launch of
            // their own after it
            GemmArgs q = {};
            bool yr_in_kernel = false;
            if (c >= 1) {
                const int mp = chunk_m(c - 1);
                const double* yb;
                long long ldy, ygs;
                if ((rc = chunk_inputs(h, y, O, G, chunk_t0(c - 1), mp, Yc, L.cld, &yb, &ldy, &ygs, st))) return rc;
                q.A = yb, q.lda = ldy, q.strideA = ygs;
                q.B = Rbuf[(c - 1) & 1], q.ldb = N, q.strideB = (long long)(mp + 1) * N;
                q.C = wout, q.ldc = N, q.strideC = wout_gs;
                q.M = O, q.N = N, q.K = mp, q.batch = G, q.alpha = 1.0, q.beta = 1.0, q.flags = GEMM_A_KMAJOR;
                yr_in_kernel = i8 && !pull && O <= YR_SKINNY_MAX_ROWS && !getenv("XESN_B200_YR_SEPARATE");
            }
            GemmArgs q_none = {};
            {
                ProfScope prof(PROF_FUSED, st);
                rc = pull ? launch_fused_pull_i8(r, G, g8, h->d_tiles_i8, ntiles, h->num_sms, st)
                     : split ? launch_fused_cluster_split_i8(r, G, cteam, h->cl_plan[cteam].d_tail_ptr, ctail, g8, h->d_tiles_i8,
                                                             ntiles, yr_in_kernel ? q : q_none, dv, h->num_sms, st,
                                                             h->worker_stream, h->ev_wfork, h->ev_wjoin, h->d_tile_counter)
                     : cluster ? launch_fused_cluster_i8(r, G, cteam, h->cl_plan[cteam].d_tail_ptr, ctail, g8, h->d_tiles_i8, ntiles,
                                                         yr_in_kernel ? q : q_none, dv, h->num_sms, st, h->d_tile_counter)
                     : i8 ? launch_fused_chunk_i8(r, G, team, upt, halves, slice_nnz, g8, h->d_tiles_i8, ntiles,
                                                  yr_in_kernel ? q : q_none, dv, h->num_sms, st, h->d_tile_counter)
                          : launch_fused_chunk(r, G, team, upt, slice_nnz, g, h->d_tiles, ntiles, h->num_sms, st);
                if (rc) return rc;
            }
            if (c >= 1 && !yr_in_kernel) {
                ProfScope prof(PROF_YR, st);
                if ((rc = (O <= YR_SKINNY_MAX_ROWS ? launch_yr_skinny(q, st) : launch_gemm_f64(q, st)))) return rc;
            }
        }
        return 0;
    }
    for (long long t0 = n_spinup; t0 < T; t0 += chunk) {
        const int m = (int)(T - t0 < chunk ? T - t0 : chunk);
        const double *ub, *yb;
        long long ldu, ugs, ldy, ygs;
        int rc = chunk_inputs(h, u, h->I, G, t0, m, Uc, L.cld, &ub, &ldu, &ugs, st);
        if (rc) return rc;
        if ((rc = chunk_inputs(h, y, O, G, t0, m, Yc, L.cld, &yb, &ldy, &ygs, st))) return rc;
        if ((rc = drive_gemm(h, ub, ldu, ugs, G, m, D, st))) return rc;
        RecurArgs r = recur_args(h);
        const long long rgs = (long long)(m + 1) * N;
        r.drive = D, r.drive_group_stride = (long long)m * N, r.steps = m, r.carry = carry;
        r.states = R, r.state_group_stride = rgs;
        {
            ProfScope prof(PROF_RECUR, st);
            if ((rc = forced_launch(h, r, G, st))) return rc;
        }
        // rbar += R^T R  (rows 0..m-1: the state that has seen u(..t-1) pairs with y(t))
        GemmArgs g = {};
        g.A = R, g.lda = N, g.strideA = rgs;
        g.B = R, g.ldb = N, g.strideB = rgs;
        g.C = gram, g.ldc = N, g.strideC = gram_gs;
        g.M = N, g.N = N, g.K = m, g.batch = G, g.alpha = 1.0, g.beta = 1.0, g.flags = GEMM_LOWER;
        {
            ProfScope prof(PROF_GRAM, st);
            if ((rc = launch_gemm_f64(g, st))) return rc;
        }
        // ybar += Y R
        GemmArgs q = {};
        q.A = yb, q.lda = ldy, q.strideA = ygs;
        q.B = R, q.ldb = N, q.strideB = rgs;
        q.C = wout, q.ldc = N, q.strideC = wout_gs;
        q.M = O, q.N = N, q.K = m, q.batch = G, q.alpha = 1.0, q.beta = 1.0, q.flags = GEMM_A_KMAJOR;
        {
            ProfScope prof(PROF_YR, st);
            if ((rc = (O <= YR_SKINNY_MAX_ROWS ? launch_yr_skinny(q, st) : launch_gemm_f64(q, st)))) return rc;
        }
    }
    return 0;
}

// 0 = FP64 DMMA Gram, 1 = fixed-point int8 tensor-core Gram (default; fused pipeline only)
int xesn_set_gram_mode(xesn_reservoir* h, int32_t mode) {
    XESN_REQUIRE(h && (mode == 0 || mode == 1), "bad arguments");
    h->gram_mode = mode;
    return 0;
}

// 0 = team-model recurrence (shared-memory state copy per CTA), 1 = pull model (every thread polls the states
// its rows need) for forced runs / spin-up only, 2 = pull model also inside the fused training kernel (the
// recurrence rides on the idle warps of the Gram CTAs)
int xesn_set_recur_mode(xesn_reservoir* h, int32_t mode) {
    XESN_REQUIRE(h && mode >= 0 && mode <= 2, "bad arguments");
    h->recur_pull = mode;
    return 0;
}

// switch the fused recurrence+Gram chunk kernel on/off (A/B measurements, debugging)
int xesn_set_fused(xesn_reservoir* h, int32_t on) {
    XESN_REQUIRE(h, "null pointer");
    h->fused_enabled = on != 0;
    return 0;
}

int xesn_ridge_solve(xesn_reservoir* h, int32_t O, int32_t G, const double* beta_dev, double beta_scalar, double* gram,
                     double* wout, int32_t* info, void* scratch, int64_t scratch_bytes, void* stream) {
    XESN_REQUIRE(h && gram && wout && info && scratch, "null pointer");
    XESN_REQUIRE(scratch_bytes >= cholesky_workspace_doubles(h->N, O, G) * 8, "scratch too small");
    cudaStream_t st = (cudaStream_t)stream;
    const int N = h->N;
    ProfScope prof(PROF_SOLVE, st);
    const bool inter = gram_interleaved(gram, wout, N);
    const long long gram_gs = inter ? (long long)(N + O) * N : (long long)N * N;
    const long long wout_gs = inter ? (long long)(N + O) * N : (long long)O * N;
    int rc = launch_add_diag(gram, N, gram_gs, N, G, beta_dev, beta_scalar, st);
    if (rc) return rc;
    return launch_cholesky_solve_ws(h->solve, gram, N, gram_gs, wout, N, wout_gs, N, O, G, info, (double*)scratch, st);
}

// Pivoted fallback for ONE system whose Cholesky factorisation failed (info > 0 from xesn_ridge_solve): the caller
// re-accumulates gram/wout of that group (xesn_train_accumulate) and hands them here.  LU with partial pivoting,
// the factorisation the reference's own GPU path uses (cupy.linalg.solve; scipy's assume_a="sym" is pivoted LDL^T).
int xesn_ridge_solve_pivoted(xesn_reservoir* h, int32_t O, double beta, double* gram, double* wout, int32_t* info,
                             void* scratch, int64_t scratch_bytes, void* stream) {
    XESN_REQUIRE(h && gram && wout && info && scratch, "null pointer");
    XESN_REQUIRE(scratch_bytes >= lu_workspace_doubles(h->N) * 8, "scratch too small");
    cudaStream_t st = (cudaStream_t)stream;
    const int N = h->N;
    ProfScope prof(PROF_SOLVE, st);
    if (int rc = launch_add_diag(gram, N, (long long)N * N, N, 1, nullptr, beta, st)) return rc;
    return launch_lu_solve(gram, N, wout, N, N, O, info, (double*)scratch, st);
}

// ---------------------------------------------------------------------------------------------
// closed loop
// ---------------------------------------------------------------------------------------------
constexpr int INLINE_PROJECTION_MAX_GROUPS = 4;

// readout scratch owned by the handle: partial sums of the split reservoir axis + arrival counters
static int ensure_readout_scratch(xesn_reservoir* h, int O, int G, int* n_slices) {
    const int S = readout_slices(h->N, O, G);
    *n_slices = S;
    if (S == 1) return 0;
    const long long need_p = (long long)G * S * O, need_c = (long long)G * ((O + 7) / 8);
    if (need_p > h->ro_partial_cap) {
        if (h->d_ro_partial) XESN_CUDA_OK(cudaFree(h->d_ro_partial));
        XESN_CUDA_OK(cudaMalloc(&h->d_ro_partial, sizeof(double) * need_p));
        h->ro_partial_cap = need_p;
    }
    if (need_c > h->ro_counter_cap) {
        if (h->d_ro_counters) XESN_CUDA_OK(cudaFree(h->d_ro_counters));
        XESN_CUDA_OK(cudaMalloc(&h->d_ro_counters, sizeof(int) * need_c));
        XESN_CUDA_OK(cudaMemset(h->d_ro_counters, 0, sizeof(int) * need_c));
        h->ro_counter_cap = need_c;
    }
    return 0;
}

// r_out = update(r_in, gather(field_in));  field_out[out_map] = Wout r_out  (+ optional copy)
static int closed_loop_step(xesn_reservoir* h, const double* wout, int O, int G, const int32_t* in_map,
                            const double* fill, const int32_t* out_map, const double* field_in, double* field_out,
                            const double* r_in, double* r_out, double* pred_col, long long pred_stride, double* Ug,
                            double* drive, int n_slices, cudaStream_t st) {
    const int N = h->N, I = h->I;
    int rc;
    StepArgs a = {};
    a.indptr = h->d_indptr, a.indices = h->d_indices, a.values = h->d_values;
    a.Win = h->d_win, a.bias = h->d_bias, a.N = N, a.I = I, a.alpha = h->alpha;
    a.values_gs = h->values_gs, a.win_gs = h->win_gs, a.bias_gs = h->bias_gs, a.alpha_g = h->d_alpha_g;
    a.param_groups = h->n_param_groups;
    a.r_in = r_in, a.r_out = r_out, a.mask_nan = h->mask_nan_inputs;
    if (G > INLINE_PROJECTION_MAX_GROUPS) {
        if ((rc = launch_gather_vec(field_in, 1, in_map, fill, Ug, (long long)G * I, h->mask_nan_inputs, st))) return rc;
        GemmArgs g = {};
        g.A = Ug, g.lda = I, g.strideA = 0;
        g.B = h->d_win, g.ldb = I, g.strideB = h->win_gs;
        g.C = drive, g.ldc = N, g.strideC = 0;
        g.bias = h->d_bias, g.strideBias = h->bias_gs, g.modB = h->n_param_groups;
        g.M = G, g.N = N, g.K = I, g.batch = 1, g.alpha = 1.0, g.beta = 0.0;
        g.flags = GEMM_A_KMAJOR | GEMM_B_KMAJOR;
        if ((rc = launch_gemm_f64(g, st))) return rc;
        a.drive = drive;
    } else {
        a.field = field_in, a.in_map = in_map, a.fill = fill;
    }
    if ((rc = launch_step_update(a, G, st))) return rc;
    ReadoutArgs q = {};
    q.Wout = wout, q.r = r_out, q.v = field_out, q.v_stride = 1;
    q.v_copy = pred_col, q.v_copy_stride = pred_stride, q.out_map = out_map, q.N = N, q.O = O;
    q.n_slices = n_slices, q.partial = h->d_ro_partial, q.counters = h->d_ro_counters;
    return launch_readout(q, G, st);
}

// the persistent forecast kernel projects the input inline: use it while the per-step Win traffic
// (every group reads its rows of Win) stays cache-sized; bandwidth-bound LazyESN shapes keep the
// GEMM-based per-step path
static bool predict_loop_eligible(xesn_reservoir* h, int G) {
    return h->persistent_predict && (long long)G * h->N * h->I * 8 <= (64LL << 20) && (long long)h->I * 8 <= 200 * 1024;
}

static PredictLoopArgs predict_loop_args(xesn_reservoir* h, const double* wout, int O, int G, const int32_t* in_map,
                                         const double* fill, const int32_t* out_map, double* const* fbuf,
                                         double* const* rbuf, double* pred, long long ldp, int n_steps, int n_slices) {
    PredictLoopArgs a = {};
    a.indptr = h->d_indptr, a.indices = h->d_indices, a.values = h->d_values;
    a.Win = h->d_win, a.bias = h->d_bias, a.Wout = wout;
    a.values_gs = h->values_gs, a.win_gs = h->win_gs, a.bias_gs = h->bias_gs, a.alpha_g = h->d_alpha_g;
    a.param_groups = h->n_param_groups;
    a.in_map = in_map, a.fill = fill, a.out_map = out_map;
    a.field[0] = fbuf[0], a.field[1] = fbuf[1], a.r[0] = rbuf[0], a.r[1] = rbuf[1];
    a.pred = pred, a.ldp = ldp;
    a.N = h->N, a.I = h->I, a.O = O, a.G = G, a.n_steps = n_steps, a.alpha = h->alpha;
    // one readout item per CTA if possible: the phase then costs one pass, not several
    {
        const long long rows = (long long)G * ((O + 7) / 8);
        const long long fit = rows > 0 ? h->num_sms / rows : 1;
        if (n_slices > fit) n_slices = (int)(fit < 1 ? 1 : fit);
    }
    a.n_slices = n_slices;
    a.slice = (h->N + n_slices - 1) / n_slices;
    a.slice += a.slice & 1;
    a.partial = h->d_ro_partial, a.counters = h->d_ro_counters;
    a.bar = h->d_bar, a.status = h->d_status, a.mask_nan = h->mask_nan_inputs;
    a.world = 1, a.rank = 0, a.step_base = 0;
    return a;
}

// steps per launch of the barrier-free forecast kernel (bounds its rings: (257 N + 256 x 148 x O) x 8 bytes)
constexpr int PULL_PREDICT_CHUNK = 256;

static bool predict_pull_eligible(xesn_reservoir* h, int O, int G, long long n_cells, int* upc, int* n_ctas) {
    return h->persistent_predict == 1 && G == 1 && h->n_param_groups == 0 && n_cells > 0 &&
           predict_pull_geometry(h->N, h->I, O, n_cells, h->num_sms, upc, n_ctas);
}

// dataflow forecast (predict_flow.cu): several reservoirs, no grid barrier
static bool predict_flow_eligible(xesn_reservoir* h, int O, int G, int* team, int* upc, int* tail_cap = nullptr,
                                  int* cluster = nullptr) {
    int tc = 0, cl = 0;
    bool ok = h->persistent_predict == 1 && G >= 1 &&
              predict_flow_geometry(h->N, h->I, O, G, h->num_sms, h->h_indptr.data(), team, upc, &tc, &cl);
    // one reservoir: only the cluster variant (everything a step needs travels through distributed shared memory); the
    // L2 variant would lose to the barrier-free kernel of predict_pull.cu.  XESN_B200_FLOW_SINGLE=0 keeps that kernel.
    if (ok && G == 1) {
        const char* e = getenv("XESN_B200_FLOW_SINGLE");
        ok = cl == 1 && !(e && e[0] == '0');
    }
    if (ok && tail_cap) *tail_cap = tc;
    if (ok && cluster) *cluster = cl;
    return ok;
}
struct FlowLayout {
    long long r_ring, part_ring, src, owner, total;  // offsets in doubles from the start of the flow area
};
static FlowLayout flow_layout(xesn_reservoir* h, int O, int G, long long n_cells, int team, bool own_ring) {
    FlowLayout L;
    long long off = 0;
    L.r_ring = off, off += round_up((long long)G * XESN_FLOW_CHUNK * h->N, 2);
    L.part_ring = off, off += own_ring ? round_up((long long)G * O * XESN_FLOW_CHUNK * team, 2) : 0;
    L.src = off, off += own_ring ? round_up(((long long)G * h->I + 1) / 2, 2) : 0;
    L.owner = off, off += own_ring ? round_up((n_cells + 1) / 2, 2) : 0;
    L.total = off;
    return L;
}

int64_t xesn_predict_scratch_bytes(xesn_reservoir* h, int32_t O, int32_t G, int64_t n_cells) {
    if (!h) return -1;
    long long d = round_up((long long)G * h->I, 2) + 2 * round_up((long long)G * h->N, 2) + round_up(n_cells, 2);
    int upc = 0, n_ctas = 0, team = 0;
    if (predict_pull_eligible(h, O, G, n_cells, &upc, &n_ctas))
        d += (long long)(PULL_PREDICT_CHUNK + 1) * h->N + (long long)PULL_PREDICT_CHUNK * (n_ctas + 1) * O;
    if (predict_flow_eligible(h, O, G, &team, &upc)) d += flow_layout(h, O, G, n_cells, team, true).total;
    return d * 8 + 256;
}

// one forecast through the dataflow kernel: launches of <= XESN_FLOW_CHUNK steps; the partial ring of parity p is
// re-armed ("unset") right after the launch that used it, so a peer that runs ahead into its next-but-one launch
// finds it clean (see INTEGRATION/DESIGN: only ranks that exchange cells with each other can drift, by one launch)
struct FlowRun {
    const double* wout;
    int O, G;
    long long n_cells;
    const int32_t *in_map, *out_map, *src, *import_cell, *exp_ptr, *exp_peer, *exp_slot;
    const double* fill;
    int n_import, team, upc, tail_cap, cluster;
    double *field, *field_alt, *r, *r_alt, *pred, *r_ring;
    long long ldp;
    int n_steps;
    double* ring[2];                             // this rank's partial rings (two parities; equal when there are no peers)
    int world;
    double* peer_ring[2][XESN_MAX_WORLD];        // the peers' rings as mapped here
};
static int flow_run(xesn_reservoir* h, const FlowRun& f, uint32_t* launches_done, cudaStream_t st) {
    const int N = h->N;
    if (h->flow_tail_team != f.team || h->flow_tail_upc != f.upc) {
        std::vector<int> tp((size_t)f.team * (f.upc + 1));
        predict_flow_tail_ptr(N, f.team, f.upc, h->h_indptr.data(), tp.data());
        if (h->d_flow_tail_ptr) XESN_CUDA_OK(cudaFree(h->d_flow_tail_ptr));
        h->d_flow_tail_ptr = nullptr;
        XESN_CUDA_OK(cudaMalloc(&h->d_flow_tail_ptr, sizeof(int) * tp.size()));
        XESN_CUDA_OK(cudaMemcpy(h->d_flow_tail_ptr, tp.data(), sizeof(int) * tp.size(), cudaMemcpyHostToDevice));
        h->flow_tail_team = f.team, h->flow_tail_upc = f.upc;
    }
    PredictFlowArgs a = {};
    a.tail_cap = f.tail_cap, a.tail_ptr = h->d_flow_tail_ptr;
    a.indptr = h->d_indptr, a.indices = h->d_indices, a.values = h->d_values, a.Win = h->d_win, a.bias = h->d_bias;
    a.values_gs = h->values_gs, a.win_gs = h->win_gs, a.bias_gs = h->bias_gs, a.param_groups = h->n_param_groups;
    a.alpha_g = h->d_alpha_g, a.alpha = h->alpha, a.Wout = f.wout;
    a.N = N, a.I = h->I, a.O = f.O, a.G = f.G, a.team = f.team, a.upc = f.upc, a.ring_steps = XESN_FLOW_CHUNK;
    a.cluster = f.cluster;
    a.src = f.src, a.in_map = f.in_map, a.fill = f.fill, a.out_map = f.out_map;
    a.import_cell = f.import_cell, a.n_import = f.n_import;
    a.r_ring = f.r_ring, a.pred = f.pred, a.ldp = f.ldp;
    a.exp_ptr = f.exp_ptr, a.exp_peer = f.exp_peer, a.exp_slot = f.exp_slot;
    a.mask_nan = h->mask_nan_inputs, a.status = h->d_status;
    const size_t ring_bytes = sizeof(double) * (size_t)((long long)f.G * f.O + f.n_import) * XESN_FLOW_CHUNK * f.team;
    double* fbuf[2] = {f.field, f.field_alt};
    double* rbuf[2] = {f.r, f.r_alt};
    int cur = 0;
    uint32_t local_count = 0;
    uint32_t* count = launches_done ? launches_done : &local_count;
    const bool shared_ring = f.ring[0] == f.ring[1];
    for (int s0 = 0; s0 < f.n_steps; s0 += XESN_FLOW_CHUNK, cur ^= 1) {
        const int steps = std::min(XESN_FLOW_CHUNK, f.n_steps - s0);
        const int par = (int)(*count & 1u);
        if (shared_ring) XESN_CUDA_OK(cudaMemsetAsync(f.ring[0], 0xFF, ring_bytes, st));
        // (the cluster variant moves the states through distributed shared memory, but its fallback polls this ring)
        XESN_CUDA_OK(cudaMemset2DAsync(f.r_ring, sizeof(double) * (size_t)XESN_FLOW_CHUNK * N, 0xFF,
                                       sizeof(double) * (size_t)steps * N, f.G, st));
        // cells nobody produces keep their value in the field leaving the launch
        XESN_CUDA_OK(cudaMemcpyAsync(fbuf[cur ^ 1], fbuf[cur], sizeof(double) * (size_t)f.n_cells, cudaMemcpyDeviceToDevice, st));
        a.steps = steps, a.col0 = s0;
        a.field = fbuf[cur], a.field_out = fbuf[cur ^ 1], a.r_in = rbuf[cur], a.r_out = rbuf[cur ^ 1];
        a.part_ring = f.ring[par];
        for (int k = 0; k < f.world && k < XESN_MAX_WORLD; ++k) a.peer_ring[k] = f.peer_ring[par][k];
        if (int rc = launch_predict_flow(a, st)) return rc;
        if (!shared_ring) XESN_CUDA_OK(cudaMemsetAsync(f.ring[par], 0xFF, ring_bytes, st));
        ++*count;
    }
    if (cur) {  // results live in the alternates
        XESN_CUDA_OK(cudaMemcpyAsync(f.r, f.r_alt, sizeof(double) * (size_t)f.G * N, cudaMemcpyDeviceToDevice, st));
        XESN_CUDA_OK(cudaMemcpyAsync(f.field, f.field_alt, sizeof(double) * (size_t)f.n_cells, cudaMemcpyDeviceToDevice, st));
    }
    return 0;
}

int xesn_predict_step(xesn_reservoir* h, const double* wout, int32_t O, int32_t G, const int32_t* in_map,
                      const double* fill, const int32_t* out_map, const double* field_in, double* field_out, double* r,
                      double* pred_col, int64_t pred_stride, void* scratch, int64_t scratch_bytes, void* stream) {
    XESN_REQUIRE(h && wout && in_map && field_in && field_out && r && scratch, "null pointer");
    XESN_REQUIRE(scratch_bytes >= xesn_predict_scratch_bytes(h, O, G, 0), "scratch too small");
    cudaStream_t st = (cudaStream_t)stream;
    const int N = h->N, I = h->I;
    double* Ug = (double*)scratch;
    double* drive = Ug + round_up((long long)G * I, 2);
    double* r_new = drive + round_up((long long)G * N, 2);
    int S, rc;
    if ((rc = ensure_readout_scratch(h, O, G, &S))) return rc;
    if ((rc = closed_loop_step(h, wout, O, G, in_map, fill, out_map, field_in, field_out, r, r_new, pred_col,
                               pred_stride, Ug, drive, S, st)))
        return rc;
    XESN_CUDA_OK(cudaMemcpyAsync(r, r_new, sizeof(double) * (size_t)G * N, cudaMemcpyDeviceToDevice, st));
    return 0;
}

int xesn_predict(xesn_reservoir* h, const double* wout, int32_t O, int32_t G, int64_t n_cells, const int32_t* in_map,
                 const double* fill, const int32_t* out_map, double* field, double* r, int32_t n_steps, double* pred,
                 void* scratch, int64_t scratch_bytes, void* stream) {
    XESN_REQUIRE(h && wout && in_map && field && r && pred && scratch, "null pointer");
    XESN_REQUIRE(n_steps >= 0 && n_cells > 0, "bad sizes");
    XESN_REQUIRE_GROUPS(h, G);
    XESN_REQUIRE(scratch_bytes >= xesn_predict_scratch_bytes(h, O, G, n_cells), "scratch too small");
    cudaStream_t st = (cudaStream_t)stream;
    const int N = h->N, I = h->I;
    double* Ug = (double*)scratch;
    double* drive = Ug + round_up((long long)G * I, 2);
    double* r_alt = drive + round_up((long long)G * N, 2);
    double* field_alt = r_alt + round_up((long long)G * N, 2);
    const long long ldp = (long long)n_steps + 1;
    int S, rc;
    if ((rc = ensure_readout_scratch(h, O, G, &S))) return rc;
    ProfScope prof(PROF_PREDICT, st);

    strided_copy_kernel<<<(unsigned)((n_cells + 255) / 256), 256, 0, st>>>(field, pred, ldp, n_cells);
    XESN_CUDA_OK(cudaGetLastError());
    count_launch();
    // cells not covered by out_map keep their previous value in both field buffers
    XESN_CUDA_OK(cudaMemcpyAsync(field_alt, field, sizeof(double) * n_cells, cudaMemcpyDeviceToDevice, st));

    double* rbuf[2] = {r, r_alt};
    double* fbuf[2] = {field, field_alt};
    int pp_upc = 0, pp_ctas = 0;
    int fteam = 0, fupc = 0, ftail = 0, fcluster = 0;
    const bool flow_ok = predict_flow_eligible(h, O, G, &fteam, &fupc, &ftail, &fcluster);
    if (!(G == 1 && flow_ok) && predict_pull_eligible(h, O, G, n_cells, &pp_upc, &pp_ctas)) {
        // one reservoir: barrier-free forecast kernel, PULL_PREDICT_CHUNK steps per launch
        PredictPullArgs a = {};
        a.indptr = h->d_indptr, a.indices = h->d_indices, a.values = h->d_values;
        a.Win = h->d_win, a.bias = h->d_bias, a.Wout = wout;
        a.in_map = in_map, a.fill = fill, a.out_map = out_map;
        a.field = field, a.r = r;
        a.r_ring = field_alt + round_up(n_cells, 2);
        a.part_ring = a.r_ring + (long long)(PULL_PREDICT_CHUNK + 1) * N;
        a.v_ring = a.part_ring + (long long)PULL_PREDICT_CHUNK * pp_ctas * O;
        a.pred = pred, a.ldp = ldp, a.n_cells = n_cells;
        a.N = N, a.I = I, a.O = O, a.alpha = h->alpha, a.upc = pp_upc, a.n_ctas = pp_ctas;
        a.hint = h->d_bar + 2, a.status = h->d_status, a.mask_nan = h->mask_nan_inputs;
        for (int s0 = 0; s0 < n_steps; s0 += PULL_PREDICT_CHUNK) {
            a.col0 = s0, a.steps = std::min(PULL_PREDICT_CHUNK, n_steps - s0);
            if ((rc = launch_predict_pull(a, st))) return rc;
        }
        return 0;
    }
    if (flow_ok) {
        // several reservoirs on one GPU (or one small reservoir on a cluster): dataflow kernel, every produced cell is local (no imports / exports)
        const FlowLayout L = flow_layout(h, O, G, n_cells, fteam, true);
        double* area = field_alt + round_up(n_cells, 2);
        int* src = reinterpret_cast<int*>(area + L.src);
        int* owner = reinterpret_cast<int*>(area + L.owner);
        if ((rc = launch_flow_src(in_map, out_map, owner, src, n_cells, (long long)G * O, (long long)G * I, st))) return rc;
        FlowRun f = {};
        f.wout = wout, f.O = O, f.G = G, f.n_cells = n_cells, f.in_map = in_map, f.fill = fill, f.out_map = out_map;
        f.src = src, f.team = fteam, f.upc = fupc, f.tail_cap = ftail, f.cluster = fcluster;
        f.field = field, f.field_alt = field_alt, f.r = r, f.r_alt = r_alt, f.n_steps = n_steps, f.pred = pred, f.ldp = ldp;
        f.r_ring = area + L.r_ring, f.ring[0] = f.ring[1] = area + L.part_ring, f.world = 1;
        return flow_run(h, f, nullptr, st);
    }
    if (predict_loop_eligible(h, G)) {
        PredictLoopArgs a = predict_loop_args(h, wout, O, G, in_map, fill, out_map, fbuf, rbuf, pred, ldp, n_steps, S);
        if ((rc = launch_predict_loop(a, h->num_sms, st))) return rc;
        if (n_steps & 1) {
            XESN_CUDA_OK(cudaMemcpyAsync(r, r_alt, sizeof(double) * (size_t)G * N, cudaMemcpyDeviceToDevice, st));
            XESN_CUDA_OK(cudaMemcpyAsync(field, field_alt, sizeof(double) * n_cells, cudaMemcpyDeviceToDevice, st));
        }
        return 0;
    }
    for (int n = 1; n <= n_steps; ++n) {
        const int cur = (n - 1) & 1, nxt = n & 1;
        if ((rc = closed_loop_step(h, wout, O, G, in_map, fill, out_map, fbuf[cur], fbuf[nxt], rbuf[cur], rbuf[nxt],
                                   pred + n, ldp, Ug, drive, S, st)))
            return rc;
    }
    if (n_steps & 1) {
        XESN_CUDA_OK(cudaMemcpyAsync(r, r_alt, sizeof(double) * (size_t)G * N, cudaMemcpyDeviceToDevice, st));
        XESN_CUDA_OK(cudaMemcpyAsync(field, field_alt, sizeof(double) * n_cells, cudaMemcpyDeviceToDevice, st));
    }
    return 0;
}

// ---- peer-mapped buffers for the in-kernel multi-GPU exchange (CUDA IPC, one process per GPU) ----
int xesn_p2p_alloc(int64_t bytes, void** ptr_out, uint8_t* handle64) {
    XESN_REQUIRE(bytes > 0 && ptr_out && handle64, "bad arguments");
    void* p = nullptr;
    XESN_CUDA_OK(cudaMalloc(&p, (size_t)bytes));
    XESN_CUDA_OK(cudaMemset(p, 0, (size_t)bytes));
    cudaIpcMemHandle_t hd;
    XESN_CUDA_OK(cudaIpcGetMemHandle(&hd, p));
    static_assert(sizeof(hd) == 64, "cudaIpcMemHandle_t is 64 bytes");
    memcpy(handle64, &hd, 64);
    *ptr_out = p;
    return 0;
}
int xesn_p2p_open(const uint8_t* handle64, void** ptr_out) {
    XESN_REQUIRE(handle64 && ptr_out, "bad arguments");
    cudaIpcMemHandle_t hd;
    memcpy(&hd, handle64, 64);
    XESN_CUDA_OK(cudaIpcOpenMemHandle(ptr_out, hd, cudaIpcMemLazyEnablePeerAccess));
    return 0;
}
int xesn_memcpy_async(void* dst, const void* src, int64_t bytes, void* stream) {
    XESN_CUDA_OK(cudaMemcpyAsync(dst, src, (size_t)bytes, cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
    return 0;
}
int xesn_p2p_close(void* ptr) {
    if (ptr) XESN_CUDA_OK(cudaIpcCloseMemHandle(ptr));
    return 0;
}
int xesn_p2p_free(void* ptr) {
    if (ptr) XESN_CUDA_OK(cudaFree(ptr));
    return 0;
}

// Multi-GPU closed loop inside ONE persistent kernel per rank: field_ptrs[parity*world + k] is rank k's
// field buffer of that parity as mapped in THIS process (k == rank: the local buffer), flag_ptrs[k] rank
// k's flag array (world uint32).  Every rank must call this with the same n_steps / step_base.
int xesn_predict_p2p(xesn_reservoir* h, const double* wout, int32_t O, int32_t G, int64_t n_cells,
                     const int32_t* in_map, const double* fill, const int32_t* out_map, double* r, int32_t n_steps,
                     double* pred, int32_t world, int32_t rank, uint32_t step_base, const uint64_t* field_ptrs,
                     const uint64_t* flag_ptrs, void* scratch, int64_t scratch_bytes, void* stream) {
    XESN_REQUIRE(h && wout && in_map && r && pred && field_ptrs && flag_ptrs && scratch, "null pointer");
    XESN_REQUIRE(world >= 1 && world <= XESN_MAX_WORLD && rank >= 0 && rank < world, "bad world/rank");
    XESN_REQUIRE(scratch_bytes >= xesn_predict_scratch_bytes(h, O, G, n_cells), "scratch too small");
    XESN_REQUIRE((long long)h->I * 8 <= 200 * 1024, "n_input too large for the persistent forecast kernel");
    cudaStream_t st = (cudaStream_t)stream;
    const int N = h->N, I = h->I;
    double* Ug = (double*)scratch;
    double* drive = Ug + round_up((long long)G * I, 2);
    double* r_alt = drive + round_up((long long)G * N, 2);
    int S, rc;
    if ((rc = ensure_readout_scratch(h, O, G, &S))) return rc;
    double* fbuf[2] = {reinterpret_cast<double*>(field_ptrs[0 * world + rank]),
                       reinterpret_cast<double*>(field_ptrs[1 * world + rank])};
    double* rbuf[2] = {r, r_alt};
    const long long ldp = (long long)n_steps + 1;
    ProfScope prof(PROF_PREDICT, st);
    PredictLoopArgs a = predict_loop_args(h, wout, O, G, in_map, fill, out_map, fbuf, rbuf, pred, ldp, n_steps, S);
    a.world = world, a.rank = rank, a.step_base = step_base;
    for (int k = 0; k < world; ++k) {
        a.peer_field[k] = reinterpret_cast<double*>(field_ptrs[k]);
        a.peer_field[world + k] = reinterpret_cast<double*>(field_ptrs[world + k]);
        a.peer_flags[k] = reinterpret_cast<unsigned*>(flag_ptrs[k]);
    }
    if ((rc = launch_predict_loop(a, h->num_sms, st))) return rc;
    if (n_steps & 1) XESN_CUDA_OK(cudaMemcpyAsync(r, r_alt, sizeof(double) * (size_t)G * N, cudaMemcpyDeviceToDevice, st));
    return 0;
}

// ---- dataflow forecast with neighbour-only exchange (predict_flow.cu) ----
int xesn_flow_geometry(xesn_reservoir* h, int32_t O, int32_t G, int32_t* team_out, int32_t* ring_steps_out) {
    XESN_REQUIRE(h && team_out && ring_steps_out, "null pointer");
    int team = 0, upc = 0;
    if (!predict_flow_eligible(h, O, G, &team, &upc)) {
        set_error("dataflow forecast: %d groups of n_reservoir=%d, n_input=%d, n_output=%d do not fit (use xesn_predict_p2p)", G,
                  h->N, h->I, O);
        return -2;
    }
    *team_out = team, *ring_steps_out = XESN_FLOW_CHUNK;
    return 0;
}

int xesn_predict_flow(xesn_reservoir* h, xesn_flow_plan* plan, const double* wout, int32_t O, int32_t G, int64_t n_cells,
                      const int32_t* in_map, const double* fill, const int32_t* out_map, double* field, double* r,
                      int32_t n_steps, double* pred, void* scratch, int64_t scratch_bytes, void* stream) {
    XESN_REQUIRE(h && plan && wout && in_map && field && r && pred && scratch, "null pointer");
    XESN_REQUIRE(plan->world >= 1 && plan->world <= XESN_MAX_WORLD && plan->rank >= 0 && plan->rank < plan->world, "bad world/rank");
    XESN_REQUIRE(plan->src_dev && (plan->n_import == 0 || plan->import_cell_dev), "plan lacks tables");
    XESN_REQUIRE(scratch_bytes >= xesn_predict_scratch_bytes(h, O, G, n_cells), "scratch too small");
    XESN_REQUIRE_GROUPS(h, G);
    int team = 0, upc = 0, tail_cap = 0, cluster = 0;
    XESN_REQUIRE(predict_flow_eligible(h, O, G, &team, &upc, &tail_cap, &cluster), "configuration does not fit the dataflow forecast");
    cudaStream_t st = (cudaStream_t)stream;
    const int N = h->N, I = h->I;
    double* Ug = (double*)scratch;
    double* drive = Ug + round_up((long long)G * I, 2);
    double* r_alt = drive + round_up((long long)G * N, 2);
    double* field_alt = r_alt + round_up((long long)G * N, 2);
    double* area = field_alt + round_up(n_cells, 2);
    const FlowLayout L = flow_layout(h, O, G, n_cells, team, true);
    const long long ldp = (long long)n_steps + 1;
    ProfScope prof(PROF_PREDICT, st);
    strided_copy_kernel<<<(unsigned)((n_cells + 255) / 256), 256, 0, st>>>(field, pred, ldp, n_cells);
    XESN_CUDA_OK(cudaGetLastError());
    count_launch();
    FlowRun f = {};
    f.wout = wout, f.O = O, f.G = G, f.n_cells = n_cells, f.in_map = in_map, f.fill = fill, f.out_map = out_map;
    f.src = plan->src_dev, f.import_cell = plan->import_cell_dev, f.n_import = plan->n_import;
    f.exp_ptr = plan->exp_ptr_dev, f.exp_peer = plan->exp_peer_dev, f.exp_slot = plan->exp_slot_dev;
    f.team = team, f.upc = upc, f.tail_cap = tail_cap, f.cluster = cluster;
    f.field = field, f.field_alt = field_alt, f.r = r, f.r_alt = r_alt, f.n_steps = n_steps, f.pred = pred, f.ldp = ldp;
    f.r_ring = area + L.r_ring, f.world = plan->world;
    for (int par = 0; par < 2; ++par) {
        f.ring[par] = reinterpret_cast<double*>(plan->ring_ptrs[par * plan->world + plan->rank]);
        XESN_REQUIRE(f.ring[par], "plan lacks this rank's ring");
        for (int k = 0; k < plan->world; ++k) f.peer_ring[par][k] = reinterpret_cast<double*>(plan->ring_ptrs[par * plan->world + k]);
    }
    XESN_REQUIRE(f.ring[0] != f.ring[1], "the two ring parities must be different buffers");
    return flow_run(h, f, &plan->launches_done, st);
}

int xesn_memset_async(void* dst, int32_t byte_value, int64_t bytes, void* stream) {
    XESN_CUDA_OK(cudaMemsetAsync(dst, byte_value, (size_t)bytes, (cudaStream_t)stream));
    return 0;
}

// test / A-B entry: G (N x N, lower triangle) += S^T S for S = states [n_rows][N] through the int8
// tensor-core path (digitise -> tcgen05 kind::i8 Gram).  Allocates its own scratch and synchronises.
int xesn_gram_i8(const double* states, int64_t lds, int32_t n_rows, int32_t N, double* gram, int64_t ldg, void* stream) {
    XESN_REQUIRE(states && gram && n_rows > 0 && N > 0, "bad arguments");
    XESN_REQUIRE(n_rows <= gi8::MAX_K, "gram_i8: at most 16384 rows per call (int32 accumulators); accumulate in pieces");
    cudaStream_t st = (cudaStream_t)stream;
    int dev = 0;
    XESN_CUDA_OK(cudaGetDevice(&dev));
    cudaDeviceProp prop;
    XESN_CUDA_OK(cudaGetDeviceProperties(&prop, dev));
    const long long ldp = round_up(N, 256);
    const int rows_pad = (int)round_up(n_rows, gi8::KB);
    const long long plane_stride = ldp * rows_pad;
    int8_t* planes = nullptr;
    int2* d_tiles = nullptr;
    XESN_CUDA_OK(cudaMalloc(&planes, (size_t)plane_stride * gi8::NDIG));
    std::vector<int2> tiles;
    gram_i8_tile_list(N, tiles);
    XESN_CUDA_OK(cudaMalloc(&d_tiles, sizeof(int2) * tiles.size()));
    XESN_CUDA_OK(cudaMemcpyAsync(d_tiles, tiles.data(), sizeof(int2) * tiles.size(), cudaMemcpyHostToDevice, st));
    int rc = launch_digitize(states, lds, n_rows, N, planes, ldp, plane_stride, rows_pad, st);
    gi8::GramI8Args a = {};
    a.planes = planes, a.ldp = ldp, a.plane_stride = plane_stride, a.group_stride = 0;
    a.K = rows_pad, a.N = N, a.G = gram, a.ldg = ldg, a.strideG = 0, a.batch = 1;
    if (!rc) rc = launch_gram_i8(a, d_tiles, (int)tiles.size(), prop.multiProcessorCount, st);
    cudaError_t e = cudaStreamSynchronize(st);
    cudaFree(planes);
    cudaFree(d_tiles);
    if (rc) return rc;
    XESN_CUDA_OK(e);
    return 0;
}

int xesn_gemm_f64(const double* a, int64_t lda, const double* b, int64_t ldb, double* c, int64_t ldc, int32_t m,
                  int32_t n, int32_t k, double alpha, double beta, int32_t flags, void* stream) {
    XESN_REQUIRE(a && b && c, "null pointer");
    GemmArgs g = {};
    g.A = a, g.lda = lda, g.B = b, g.ldb = ldb, g.C = c, g.ldc = ldc;
    g.M = m, g.N = n, g.K = k, g.batch = 1, g.alpha = alpha, g.beta = beta, g.flags = flags;
    return launch_gemm_f64(g, (cudaStream_t)stream);
}

int xesn_readout(const double* wout, const double* r, double* v, int32_t N, int32_t O, int32_t G, void* stream) {
    XESN_REQUIRE(wout && r && v, "null pointer");
    ReadoutArgs q = {};
    q.Wout = wout, q.r = r, q.v = v, q.v_stride = 1, q.N = N, q.O = O, q.n_slices = 1;
    return launch_readout(q, G, (cudaStream_t)stream);
}

}  // extern "C"
