// Fused training-chunk kernel (sm_100a): SM-specialised persistent kernel.
//
//   CTAs [0, rec_ctas)      : the teacher-forced recurrence of time chunk c     (recur_body.cuh)
//   CTAs [rec_ctas, grid)   : the Gram SYRK tiles  rbar += R^T R  of chunk c-1  (gemm_tile.cuh)
//
// The recurrence is a latency-bound serial chain that can use at most a handful of SMs
// (~4.5 us per step at N=16000 on 16 SMs), the Gram accumulation is a throughput-bound FP64
// tensor-pipe job for everything else.  Run back to back they cost t_rec + t_gram per chunk; run
// side by side in ONE cooperative launch (one CTA per SM, so the recurrence CTAs are guaranteed
// co-resident and the scheduler cannot scatter them) the chunk costs max(t_rec, t_gram * 148/132).
// The two roles only share the state ring: the recurrence writes ring c while the Gram reads
// ring c-1 (double-buffered by the caller).
//
// Tile order of the Gram role: a host-built list that walks the lower triangle in super-rows of
// 12 tile rows, column by column, so the ~132 tiles in flight at any time touch about 12 + 11
// operand panels (L2-resident) instead of ~130 distinct ones.
#include <cuda.h>

#include <algorithm>
#include <cstdlib>
#include <cstring>

#include "common.cuh"
#include "gemm_tile.cuh"
#include "gram_i8.cuh"
#include "kernels.cuh"
#include "recur_body.cuh"
#include "pull_body.cuh"
#include "yr_skinny.cuh"
#include "recur_cluster.cuh"

namespace xesn {
namespace {

constexpr int FUSED_THREADS = 256;

template <int UPT, bool WS>
__global__ void __launch_bounds__(FUSED_THREADS, 1)
fused_chunk_kernel(RecurArgs rp, int rec_ctas, int team, GemmArgs gp, const int2* __restrict__ tiles, int ntiles) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    if ((int)blockIdx.x < rec_ctas) {
        recur::recur_body<UPT, FUSED_THREADS, WS>(rp, blockIdx.x % team, team, blockIdx.x / team, smem_raw);
        return;
    }
    const int worker = blockIdx.x - rec_ctas, nworkers = gridDim.x - rec_ctas;
    const long long total = (long long)ntiles * gp.batch;
    for (long long t = worker; t < total; t += nworkers) {
        const int2 ij = tiles[t % ntiles];
        gemm::gemm_tile<false, false, true>(gp, ij.x, ij.y, (int)(t / ntiles), reinterpret_cast<double*>(smem_raw));
    }
}

// Gram role on the int8 tensor cores: tcgen05.mma kind::i8 on the digit planes written by the
// recurrence role of the PREVIOUS launch (gram_i8.cuh).
constexpr int FUSED_I8_THREADS = 512;  // 16 warps: twice the latency hiding for the recurrence role

#ifdef XESN_RECUR_TRACE
static __device__ unsigned long long g_cta_ns[2 * 160 * 2];  // debug builds: start/end %globaltimer of every CTA
__device__ __forceinline__ unsigned long long globaltimer_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
#define XESN_CTA_STAMP(i)                                                              \
    do {                                                                               \
        __syncthreads();                                                               \
        if (threadIdx.x == 0 && blockIdx.x < 160 && ntiles > 0) g_cta_ns[(rec_ctas > 0 ? 0 : 320) + blockIdx.x * 2 + (i)] = globaltimer_ns(); \
    } while (0)
#else
#define XESN_CTA_STAMP(i) \
    do {                  \
    } while (0)
#endif

// member_bytes > 0: every recurrence CTA hosts TWO 256-thread team members of different reservoirs (virtual member
// v = half * rec_ctas + blockIdx.x; reservoirs [0, G/2) in half 0, [G/2, G) in half 1).  A member spends more than
// half of a step waiting for its peers' states; the other member computes meanwhile, so each reservoir's chain sees
// a half-size slice (shorter arithmetic, less skew between its warps) on the same number of SMs.
// DRV: the variant whose Gram CTAs can also compute the next chunk's drive (kept out of the plain variant: its
// registers would spill in the Gram role, which bounds the large-reservoir configurations)
template <int UPT, bool WS, bool DRV>
__global__ void __launch_bounds__(FUSED_I8_THREADS, 1)
fused_chunk_i8_kernel(RecurArgs rp, int rec_ctas, int team, int member_bytes, gi8::GramI8Args gp,
                      const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                      const int2* __restrict__ tiles, int ntiles, GemmArgs yrp, drv::DriveArgs dvp, int* tile_counter) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    XESN_CTA_STAMP(0);
    const long long total = (long long)ntiles * gp.batch;
    if ((int)blockIdx.x < rec_ctas) {
        bool done = false;
        if constexpr (UPT == 1) {   // the launcher pairs members only for one unit per thread
            if (member_bytes > 0) {
                const int half = threadIdx.x >> 8, v = half * rec_ctas + blockIdx.x;
                if (half == 0)
                    recur::recur_body<UPT, FUSED_I8_THREADS / 2, WS, 1>(rp, v % team, team, v / team, smem_raw, threadIdx.x & 255);
                else
                    recur::recur_body<UPT, FUSED_I8_THREADS / 2, WS, 2>(rp, v % team, team, v / team, smem_raw + member_bytes,
                                                                        threadIdx.x & 255);
                done = true;
            }
        }
        if (!done) recur::recur_body<UPT, FUSED_I8_THREADS, WS>(rp, blockIdx.x % team, team, blockIdx.x / team, smem_raw);
        XESN_CTA_STAMP(1);
        // the recurrence of this CTA is done: it joins the Gram role for whatever tiles are left (the tiles are drawn from a
        // counter, so a chunk costs (recurrence + Gram work) / SMs instead of max(recurrence, Gram on the other SMs))
        if (total > 0) {
            __syncthreads();
            gi8::GramI8Ctx ctx;
            gi8::gram_i8_setup(ctx, smem_raw);
            gi8::gram_i8_run<FUSED_I8_THREADS, gi8::NDIG, true>(ctx, gp, &tmA, &tmB, tiles, ntiles, 0, 1, tile_counter);
            gi8::gram_i8_teardown(ctx);
        }
        return;
    }
    const int worker = blockIdx.x - rec_ctas, nworkers = gridDim.x - rec_ctas;
    // The statically dealt work comes FIRST -- ybar += Y R of the previous chunk (skinny work items) and Win u + b of the
    // NEXT chunk (few inputs) -- and the Gram tiles, which are drawn from a counter, last: whatever is left when the
    // recurrence CTAs finish is then work they can share (measured at cfg 3: the launches that carry all three were
    // bounded by the workers, 3.17 ms against 2.98 for the recurrence).
    if (yrp.M > 0) {
        const long long items = yr::n_items(yrp);
        for (long long it = worker; it < items; it += nworkers)
            yr::yr_item<FUSED_I8_THREADS>(yrp, it, reinterpret_cast<double*>(smem_raw));
    }
    if constexpr (DRV) {
        if (dvp.m > 0) {
            const long long items = drv::n_items(dvp);
            for (long long it = worker; it < items; it += nworkers)
                drv::drive_item<FUSED_I8_THREADS>(dvp, it, reinterpret_cast<double*>(smem_raw));
        }
    }
    if (total > 0) {
        __syncthreads();
        gi8::GramI8Ctx ctx;
        gi8::gram_i8_setup(ctx, smem_raw);
        gi8::gram_i8_run<FUSED_I8_THREADS, gi8::NDIG, true>(ctx, gp, &tmA, &tmB, tiles, ntiles, worker, nworkers, tile_counter);
        gi8::gram_i8_teardown(ctx);
    }
    XESN_CTA_STAMP(1);
}


// The same SM-specialised chunk kernel with the recurrence role on THREAD-BLOCK CLUSTERS (recur_cluster.cuh): the
// team of a reservoir is one cluster and exchanges its states through distributed shared memory (st.async +
// mbarrier) instead of polling the L2.  The grid is a whole number of clusters; clusters [0, G) run the recurrence,
// the CTAs of the others are Gram / Y R / drive workers exactly as in fused_chunk_i8_kernel.  Not a cooperative
// launch: a cluster is co-scheduled by the hardware and nothing else in the grid waits for anything.
template <bool DRV>
__global__ void __launch_bounds__(FUSED_I8_THREADS, 1)
fused_cluster_i8_kernel(RecurArgs rp, int rec_ctas, int team, const int* __restrict__ tail_ptr, int tail_cap,
                        gi8::GramI8Args gp, const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                        const int2* __restrict__ tiles, int ntiles, GemmArgs yrp, drv::DriveArgs dvp, int* tile_counter) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const long long total = (long long)ntiles * gp.batch;
    if ((int)blockIdx.x < rec_ctas) {
        recur::recur_cluster_body(rp, recur::cluster_ctarank(), team, blockIdx.x / team, smem_raw, tail_ptr, tail_cap);
        if (total > 0) {   // done with the recurrence: help with the Gram tiles that are left (see fused_chunk_i8_kernel)
            __syncthreads();
            gi8::GramI8Ctx ctx;
            gi8::gram_i8_setup(ctx, smem_raw);
            gi8::gram_i8_run<FUSED_I8_THREADS, gi8::NDIG, true>(ctx, gp, &tmA, &tmB, tiles, ntiles, 0, 1, tile_counter);
            gi8::gram_i8_teardown(ctx);
        }
        return;
    }
    const int worker = blockIdx.x - rec_ctas, nworkers = gridDim.x - rec_ctas;
    if (yrp.M > 0) {   // statically dealt work first, the Gram tiles (drawn from a counter) last: see fused_chunk_i8_kernel
        const long long items = yr::n_items(yrp);
        for (long long it = worker; it < items; it += nworkers)
            yr::yr_item<FUSED_I8_THREADS>(yrp, it, reinterpret_cast<double*>(smem_raw));
    }
    if constexpr (DRV) {
        if (dvp.m > 0) {
            const long long items = drv::n_items(dvp);
            for (long long it = worker; it < items; it += nworkers)
                drv::drive_item<FUSED_I8_THREADS>(dvp, it, reinterpret_cast<double*>(smem_raw));
        }
    }
    if (total > 0) {
        __syncthreads();
        gi8::GramI8Ctx ctx;
        gi8::gram_i8_setup(ctx, smem_raw);
        gi8::gram_i8_run<FUSED_I8_THREADS, gi8::NDIG, true>(ctx, gp, &tmA, &tmB, tiles, ntiles, worker, nworkers, tile_counter);
        gi8::gram_i8_teardown(ctx);
    }
}

// Pull-model fusion: no SM specialisation.  EVERY CTA runs the Gram role (producer warp 4, issuer warp 0,
// epilogue warps 8..15 -- gram_i8.cuh) and its otherwise idle warps 1-3 and 5-7 run the recurrence of
// ~(G N / 148) units in the pull model (pull_body.cuh): the Gram role gets all 148 tensor cores, the
// recurrence all 148 L1s, and neither needs a barrier with the other.
template <int UPT>
__global__ void __launch_bounds__(FUSED_I8_THREADS, 1)
fused_pull_i8_kernel(RecurArgs rp, int n_groups, gi8::GramI8Args gp, const __grid_constant__ CUtensorMap tmA,
                     const __grid_constant__ CUtensorMap tmB, const int2* __restrict__ tiles, int ntiles) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ int s_sync[2];
    if (threadIdx.x < 2) s_sync[threadIdx.x] = 0;
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long long total = (long long)ntiles * gp.batch;
    const bool has_gram = (long long)blockIdx.x < total;  // uniform per CTA
    gi8::GramI8Ctx ctx;
    if (has_gram) gi8::gram_i8_setup(ctx, smem_raw);
    const bool rec_warp = (warp >= 1 && warp <= 3) || (warp >= 5 && warp <= 7);
    if (rec_warp) {
        if (rp.steps > 0) recur::pull_body<UPT>(rp, n_groups, (warp < 4 ? warp - 1 : warp - 2) * 32 + lane, blockIdx.x, s_sync);
    } else if (has_gram) {
        gi8::gram_i8_run<FUSED_I8_THREADS>(ctx, gp, &tmA, &tmB, tiles, ntiles, blockIdx.x, gridDim.x);
    }
    if (has_gram) gi8::gram_i8_teardown(ctx);
}

// stand-alone pull recurrence (spin-up, forced runs): PULL_THREADS-wide CTAs, one per SM, all co-resident
template <int UPT>
__global__ void __launch_bounds__(recur::PULL_THREADS, 1) pull_forced_kernel(RecurArgs rp, int n_groups) {
    __shared__ int s_sync[2];
    if (threadIdx.x < 2) s_sync[threadIdx.x] = 0;
    __syncthreads();
    recur::pull_body<UPT>(rp, n_groups, threadIdx.x, blockIdx.x, s_sync);
}

}  // namespace

// team size / units per thread of the recurrence role for FUSED_THREADS-wide CTAs
int fused_recur_geometry(int N, int n_groups, int threads, int* team_out, int* upt_out, int* halves_out) {
    if (halves_out) *halves_out = 1;
    // The recurrence role wants many thin CTAs (its per-step cost grows with the units per CTA: the
    // shared-memory gathers of W r are bank-conflict bound), the Gram role wants the SMs.  Default: `threads`
    // x UPT units per CTA with the smallest UPT whose total CTA count stays within the cap (48 CTAs for one
    // big reservoir, 64 when many small reservoirs run side by side).  Measured at N=16000 with the first TMA
    // Gram role (ms per 2048-step chunk, recurrence / Gram role): 16 CTAs 10.7 / 6.8, 32 CTAs 6.9 / 7.7.
    // XESN_B200_REC_UPT / _REC_CAP / _REC_TEAM override.
    int cap = n_groups > 1 ? 64 : 48;
    int force_upt = 0;
    if (const char* e = getenv("XESN_B200_REC_CAP")) cap = atoi(e);
    if (const char* e = getenv("XESN_B200_REC_UPT")) force_upt = atoi(e);
    int best_team = -1, best_upt = 0;
    for (int upt = 1; upt <= 8; upt <<= 1) {
        if (force_upt && upt != force_upt) continue;
        int team = (N + threads * upt - 1) / (threads * upt);
        if ((long long)team * n_groups > cap) continue;
        best_team = team, best_upt = upt;
        break;
    }
    if (best_team < 0) return -2;
    // one big reservoir: ~400 units per CTA instead of one per thread (40 CTAs at N=16000: measured 67.7 ms of
    // fused kernels per 20000 steps against 71.5 with 32 CTAs and 66.8 with 48)
    int want_team = 0;
    if (n_groups == 1 && best_upt == 1 && !force_upt) want_team = std::min(cap, (N + 399) / 400);
    // one small reservoir is latency-bound and its Gram role needs few SMs: ~64 units per CTA (two warps) keeps the
    // arithmetic of a step short (measured at N=1000, ms of fused kernels per 20000 steps: 2 CTAs 36.5, 8: 32.3,
    // 16: 32.0, 32: 31.7)
    if (n_groups == 1 && !force_upt && N <= 4096) want_team = std::max(want_team, std::min(32, N / 64));
    if (const char* e = getenv("XESN_B200_REC_TEAM")) want_team = atoi(e);  // experiments: any team size
    if (want_team > best_team && (long long)want_team * n_groups <= 96) {
        const int upc = recur_units_per_cta(N, want_team);
        int upt = (upc + threads - 1) / threads;
        upt = upt <= 1 ? 1 : (upt <= 2 ? 2 : (upt <= 4 ? 4 : 8));
        if ((long long)threads * upt >= upc) best_team = want_team, best_upt = upt;
    }
    *team_out = best_team, *upt_out = best_upt;
    // several small reservoirs on the int8 pipeline: two 256-thread members of different reservoirs per SM
    // (fused_chunk_i8_kernel); same SM count, twice the team size.  XESN_B200_REC_SPLIT=0 keeps whole-CTA members.
    const char* e = getenv("XESN_B200_REC_SPLIT");
    if (halves_out && threads == 512 && n_groups >= 2 && n_groups % 2 == 0 && best_upt == 1 && !(e && e[0] == '0')) {
        int team_v = (N + 255) / 256, cap_v = cap;
        if (const char* t = getenv("XESN_B200_REC_TEAM_V")) team_v = atoi(t), cap_v = 100;   // experiments: thinner members
        const int upc_v = recur_units_per_cta(N, team_v);
        if (team_v >= 1 && upc_v <= 256 && (long long)team_v * n_groups / 2 <= cap_v && (N + upc_v - 1) / upc_v == team_v)
            *team_out = team_v, *halves_out = 2;
    }
    return 0;
}

int fused_smem_bytes(const RecurArgs& a, int team, bool* w_in_smem) {
    const int upc = (a.N + team - 1) / team;
    (void)upc;
    const long long n_pad = ((a.N + 1) & ~1) + 2;
    const long long smem_r = n_pad * 8;
    const long long smem_w = (long long)(a.smem_nnz + (a.smem_nnz & 1)) * 12 + 16;
    const long long limit = 227 * 1024;
    if (smem_r > limit) return -1;
    *w_in_smem = (smem_r + smem_w <= limit);
    long long need = smem_r + (*w_in_smem ? smem_w : 0);
    if (need < gemm::SMEM_BYTES) need = gemm::SMEM_BYTES;
    if (need < gi8::SMEM_BYTES) need = gi8::SMEM_BYTES;
    return (int)need;
}

// rp.steps == 0 -> no recurrence role; ntiles == 0 -> no Gram role.
int launch_fused_chunk(RecurArgs rp, int n_groups, int team, int upt, int max_slice_nnz, const GemmArgs& gp,
                       const int2* tiles, int ntiles, int num_sms, cudaStream_t stream) {
    const bool has_rec = rp.steps > 0;
    rp.units_per_cta = recur_units_per_cta(rp.N, team);
    rp.smem_nnz = max_slice_nnz + (max_slice_nnz & 1);
    rp.n_pad = ((rp.N + 1) & ~1) + 2;
    bool w_in_smem = false;
    const int smem = fused_smem_bytes(rp, team, &w_in_smem);
    if (smem < 0) {
        set_error("fused chunk: n_reservoir=%d does not fit the shared-memory state copy", rp.N);
        return -2;
    }
    rp.w_in_smem = w_in_smem;
    int rec_ctas = has_rec ? team * n_groups : 0;
    if (has_rec) {
        XESN_CUDA_OK(cudaMemset2DAsync(rp.states, sizeof(double) * rp.state_group_stride, 0xFF,
                                       sizeof(double) * (size_t)(rp.steps + 1) * rp.N, n_groups, stream));
    }
    int grid = num_sms;
    if (ntiles == 0) grid = rec_ctas;
    if (grid <= 0) return 0;
    if (rec_ctas >= grid && ntiles > 0) {
        set_error("fused chunk: recurrence role needs %d CTAs, device has %d SMs", rec_ctas, num_sms);
        return -2;
    }
    void (*kern)(RecurArgs, int, int, GemmArgs, const int2*, int);
    if (w_in_smem)
        kern = upt == 1 ? fused_chunk_kernel<1, true>
                        : (upt == 2 ? fused_chunk_kernel<2, true>
                                    : (upt == 4 ? fused_chunk_kernel<4, true> : fused_chunk_kernel<8, true>));
    else
        kern = upt == 1 ? fused_chunk_kernel<1, false>
                        : (upt == 2 ? fused_chunk_kernel<2, false>
                                    : (upt == 4 ? fused_chunk_kernel<4, false> : fused_chunk_kernel<8, false>));
    XESN_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    GemmArgs g = gp;
    void* args[] = {&rp, &rec_ctas, &team, &g, (void*)&tiles, &ntiles};
    XESN_CUDA_OK(cudaLaunchCooperativeKernel((void*)kern, dim3(grid), dim3(FUSED_THREADS), args, (size_t)smem, stream));
    count_launch();
    return 0;
}

}  // namespace xesn

namespace xesn {

int fused_threads(int i8) { return i8 ? FUSED_I8_THREADS : FUSED_THREADS; }


int launch_fused_chunk_i8(RecurArgs rp, int n_groups, int team, int upt, int halves, int max_slice_nnz,
                          const gi8::GramI8Args& gp, const int2* tiles, int ntiles, const GemmArgs& yr_args,
                          const drv::DriveArgs& drive_args, int num_sms, cudaStream_t stream, int* tile_counter) {
    const bool has_rec = rp.steps > 0;
    rp.units_per_cta = recur_units_per_cta(rp.N, team);
    rp.smem_nnz = max_slice_nnz + (max_slice_nnz & 1);
    rp.n_pad = ((rp.N + 1) & ~1) + 2;
    bool w_in_smem = false;
    int smem = fused_smem_bytes(rp, team, &w_in_smem);
    if (smem < 0) {
        set_error("fused chunk: n_reservoir=%d does not fit the shared-memory state copy", rp.N);
        return -2;
    }
    int member_bytes = 0;
    if (halves == 2) {
        // two members per CTA: each gets its own state copy (+ slice of W if both fit)
        const long long smem_r = (long long)rp.n_pad * 8;
        const long long smem_w = (long long)rp.smem_nnz * 12 + 16;
        w_in_smem = 2 * (smem_r + smem_w) + 32 <= 227 * 1024;
        member_bytes = (int)(((w_in_smem ? smem_r + smem_w : smem_r) + 15) & ~15LL);
        if (2LL * member_bytes > 227 * 1024) {
            set_error("fused chunk: two members of n_reservoir=%d do not fit one SM", rp.N);
            return -2;
        }
        smem = std::max(smem, 2 * member_bytes);
    }
    rp.w_in_smem = w_in_smem;
    int rec_ctas = has_rec ? team * n_groups / halves : 0;
    if (has_rec) {
        XESN_CUDA_OK(cudaMemset2DAsync(rp.states, sizeof(double) * rp.state_group_stride, 0xFF,
                                       sizeof(double) * (size_t)(rp.steps + 1) * rp.N, n_groups, stream));
    }
    int grid = num_sms;
    if (ntiles == 0 && drive_args.m <= 0) grid = rec_ctas;
    if (grid <= 0) return 0;
    if (rec_ctas >= grid && (ntiles > 0 || drive_args.m > 0)) {
        set_error("fused chunk: recurrence role needs %d CTAs, device has %d SMs", rec_ctas, num_sms);
        return -2;
    }
    void (*kern)(RecurArgs, int, int, int, gi8::GramI8Args, const CUtensorMap, const CUtensorMap, const int2*, int, GemmArgs,
                 drv::DriveArgs, int*);
    const bool drv_on = drive_args.m > 0;
    if (drv_on && upt != 1) {
        set_error("fused chunk: the in-kernel drive is built for one unit per thread");
        return -2;
    }
    if (w_in_smem)
        kern = upt == 1 ? (drv_on ? fused_chunk_i8_kernel<1, true, true> : fused_chunk_i8_kernel<1, true, false>)
                        : (upt == 2 ? fused_chunk_i8_kernel<2, true, false>
                                    : (upt == 4 ? fused_chunk_i8_kernel<4, true, false> : fused_chunk_i8_kernel<8, true, false>));
    else
        kern = upt == 1 ? (drv_on ? fused_chunk_i8_kernel<1, false, true> : fused_chunk_i8_kernel<1, false, false>)
                        : (upt == 2 ? fused_chunk_i8_kernel<2, false, false>
                                    : (upt == 4 ? fused_chunk_i8_kernel<4, false, false> : fused_chunk_i8_kernel<8, false, false>));
    XESN_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    gi8::GramI8Args g = gp;
    alignas(64) CUtensorMap tm_a, tm_b;
    memset(&tm_a, 0, sizeof(tm_a));
    memset(&tm_b, 0, sizeof(tm_b));
    if (ntiles > 0)
        if (int rc = gram_i8_make_maps(g, &tm_a, &tm_b)) return rc;
    GemmArgs yr_copy = yr_args;
    if (ntiles == 0) yr_copy.M = 0;   // no Gram CTAs in this launch: the caller runs the Y R pass itself
    smem = std::max(smem, yr::smem_bytes<FUSED_I8_THREADS>());
    drv::DriveArgs dv = drive_args;
    if (ntiles > 0) {
        XESN_REQUIRE(tile_counter, "fused chunk: no tile counter");
        XESN_CUDA_OK(cudaMemsetAsync(tile_counter, 0, sizeof(int), stream));
    }
    void* args[] = {&rp, &rec_ctas, &team, &member_bytes, &g, &tm_a, &tm_b, (void*)&tiles, &ntiles, &yr_copy, &dv, &tile_counter};
    XESN_CUDA_OK(cudaLaunchCooperativeKernel((void*)kern, dim3(grid), dim3(FUSED_I8_THREADS), args, (size_t)smem, stream));
    count_launch();
    return 0;
}


// cluster geometry of the recurrence role: team (= cluster size, <= 8) and tail staging capacity; false if the
// reservoir does not fit (one unit per thread, three state copies in shared memory)
bool cluster_recur_geometry(int N, int n_groups, const int* indptr, int num_sms, int* team_out, int* tail_cap_out) {
    const char* e = getenv("XESN_B200_REC_CLUSTER");
    if (e && e[0] == '0') return false;
    // the thinner the members the shorter a step (less arithmetic per CTA between two exchanges), within what is left
    // of the machine after the Gram role's share: at most 8 CTAs per cluster (portable size), at most ~64 SMs in all,
    // at least a warp of units per member
    int team = std::min(8, std::max(1, 64 / std::max(n_groups, 1)));
    team = std::min(team, std::max(1, (N + 31) / 32));
    team = std::max(team, (N + recur::CL_THREADS - 1) / recur::CL_THREADS);
    if (const char* t = getenv("XESN_B200_REC_CLUSTER_TEAM")) team = atoi(t);
    if (team < 1 || team > 8) return false;
    // only where the chunk is bounded by the recurrence rather than by the Gram role.  (With the Gram workers in the
    // same cluster launch the clusters also cost resident CTAs -- 8 x N=4000 lost 23 ms of 50; as two kernels side by
    // side they do not, and that shape gains 2 %: bench macro 98.2 -> 96.1 ms.)
    const double limit = getenv("XESN_B200_CLUSTER_SINGLE") ? 1.0e8 : 1.5e8;
    if ((double)n_groups * N * N > limit && !getenv("XESN_B200_REC_CLUSTER_TEAM")) return false;
    const int upc = recur_units_per_cta(N, team);
    if (upc > recur::CL_THREADS) return false;
    if ((long long)team * n_groups > num_sms - team) return false;   // at least one cluster of workers must remain
    int tail_cap = 0;
    for (int c = 0; c < team; ++c) {
        int t = 0;
        for (int i = c * upc; i < std::min(N, (c + 1) * upc); ++i) t += std::max(0, indptr[i + 1] - indptr[i] - recur::CL_KREG);
        tail_cap = std::max(tail_cap, t);
    }
    const long long n_pad = cluster_state_slots(N) + 2;
    if (recur::cluster_smem_bytes((int)n_pad, tail_cap) > 227 * 1024) return false;
    *team_out = team, *tail_cap_out = tail_cap;
    return true;
}

static int cluster_tail_cap(int N, int team, const int* indptr) {
    const int upc = recur_units_per_cta(N, team);
    int tail_cap = 0;
    for (int c = 0; c < team; ++c) {
        int t = 0;
        for (int i = c * upc; i < std::min(N, (c + 1) * upc); ++i) t += std::max(0, indptr[i + 1] - indptr[i] - recur::CL_KREG);
        tail_cap = std::max(tail_cap, t);
    }
    return tail_cap;
}

bool cluster_forced_geometry(int N, int n_groups, const int* indptr, int num_sms, int* team_out, int* tail_cap_out) {
    const char* e = getenv("XESN_B200_REC_CLUSTER");
    if (e && e[0] == '0') return false;
    // clusters of 8 keep 128 CTAs resident on a B200, clusters of 4 132 (GPC boundaries); more clusters than that would
    // simply queue (nothing waits across clusters), but then the pull model over all SMs is the better engine
    int first = 8;
    while (first > 1 && first * n_groups > 64) first >>= 1;   // as in training: at most ~64 SMs of clusters side by side
    if (const char* t = getenv("XESN_B200_FORCED_TEAM")) first = atoi(t);
    for (int team = first; team >= 1; team >>= 1) {
        if (team > 1 && (N + team - 1) / team < 32) continue;   // at least a warp of units per member
        const int upc = recur_units_per_cta(N, team);
        if (upc > recur::CL_THREADS) return false;               // thinner teams only get fatter members
        if ((long long)team * n_groups > (team >= 8 ? 128 : (team >= 4 ? 132 : num_sms))) continue;
        const int tail_cap = cluster_tail_cap(N, team, indptr);
        if (recur::cluster_smem_bytes(cluster_state_slots(N) + 2, tail_cap) > 227 * 1024) return false;
        *team_out = team, *tail_cap_out = tail_cap;
        return true;
    }
    return false;
}

void cluster_tail_ptr(int N, int team, const int* indptr, int* out) {
    const int upc = recur_units_per_cta(N, team);
    for (int c = 0; c < team; ++c) {
        int t = 0;
        for (int il = 0; il <= upc; ++il) {
            out[(long long)c * (upc + 1) + il] = t;
            const int i = c * upc + il;
            if (il < upc && i < N) t += std::max(0, indptr[i + 1] - indptr[i] - recur::CL_KREG);
        }
    }
}

// rp.steps == 0 -> no recurrence role; ntiles == 0 -> no Gram role
int launch_fused_cluster_i8(RecurArgs rp, int n_groups, int team, const int* tail_ptr_dev, int tail_cap,
                            const gi8::GramI8Args& gp, const int2* tiles, int ntiles, const GemmArgs& yr_args,
                            const drv::DriveArgs& drive_args, int num_sms, cudaStream_t stream, int* tile_counter,
                            bool rec_grid_only, bool zero_counter) {
    const bool has_rec = rp.steps > 0;
    rp.units_per_cta = recur_units_per_cta(rp.N, team);
    rp.n_pad = cluster_state_slots(rp.N) + 2;   // the last slot of a copy always holds 0
    if (has_rec && (!rp.slot || !rp.col_off)) {
        set_error("fused cluster chunk: no shared-memory placement");
        return -1;
    }
    int smem = (int)recur::cluster_smem_bytes(rp.n_pad, tail_cap);
    smem = std::max(smem, gi8::SMEM_BYTES);
    smem = std::max(smem, yr::smem_bytes<FUSED_I8_THREADS>());
    const bool drv_on = drive_args.m > 0 && !rec_grid_only;
    auto kern = drv_on ? fused_cluster_i8_kernel<true> : fused_cluster_i8_kernel<false>;
    XESN_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    int rec_ctas = has_rec ? team * n_groups : 0;
    // how many clusters of this size the device holds at once (GPC boundaries: fewer than num_sms / team)
    cudaLaunchConfig_t cfg = {};
    cfg.blockDim = dim3(FUSED_I8_THREADS);
    cfg.dynamicSmemBytes = (size_t)smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = team, attr[0].val.clusterDim.y = 1, attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr, cfg.numAttrs = 1;
    cfg.gridDim = dim3((unsigned)(num_sms / team * team));
    int max_clusters = 0;
    XESN_CUDA_OK(cudaOccupancyMaxActiveClusters(&max_clusters, kern, &cfg));
    if (!rec_grid_only && max_clusters * team <= rec_ctas && (ntiles > 0 || drv_on)) {
        set_error("fused cluster chunk: %d clusters of %d fit, the recurrence role needs %d", max_clusters, team, n_groups);
        return -2;
    }
    int grid = max_clusters * team;
    // rec_grid_only: the workers are a launch of their own (launch_fused_cluster_split_i8); the Gram arguments are only
    // there for the recurrence CTAs to help with the tiles once their chunk is done
    if ((ntiles == 0 && !drv_on) || rec_grid_only) grid = rec_ctas;
    if (grid <= 0) return 0;
    if (ntiles > 0) {
        XESN_REQUIRE(tile_counter, "fused cluster chunk: no tile counter");
        if (zero_counter) XESN_CUDA_OK(cudaMemsetAsync(tile_counter, 0, sizeof(int), stream));
    }
    cfg.gridDim = dim3((unsigned)grid);
    gi8::GramI8Args g = gp;
    alignas(64) CUtensorMap tm_a, tm_b;
    memset(&tm_a, 0, sizeof(tm_a));
    memset(&tm_b, 0, sizeof(tm_b));
    if (ntiles > 0)
        if (int rc = gram_i8_make_maps(g, &tm_a, &tm_b)) return rc;
    GemmArgs yr_copy = yr_args;
    if (ntiles == 0) yr_copy.M = 0;
    drv::DriveArgs dv = drive_args;
    if (rec_grid_only) yr_copy.M = 0, dv.m = 0;
    XESN_CUDA_OK(cudaLaunchKernelEx(&cfg, kern, rp, rec_ctas, team, tail_ptr_dev, tail_cap, g, tm_a, tm_b, tiles, ntiles, yr_copy, dv,
                                    tile_counter));
    count_launch();
    return 0;
}

// Gram / Y R / drive workers alone, as an ordinary (non-cluster, non-cooperative) launch of `n_workers` CTAs: the
// companion of a recurrence-only cluster launch (below), and the last launch of a training pass
static int launch_gram_workers_i8(const gi8::GramI8Args& gp, const int2* tiles, int ntiles, const GemmArgs& yr_args,
                                  const drv::DriveArgs& drive_args, int n_workers, cudaStream_t stream, int* tile_counter,
                                  bool zero_counter) {
    const bool drv_on = drive_args.m > 0;
    if (n_workers <= 0 || (ntiles == 0 && !drv_on)) return 0;
    auto kern = drv_on ? fused_chunk_i8_kernel<1, false, true> : fused_chunk_i8_kernel<1, false, false>;
    XESN_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    gi8::GramI8Args g = gp;
    alignas(64) CUtensorMap tm_a, tm_b;
    memset(&tm_a, 0, sizeof(tm_a));
    memset(&tm_b, 0, sizeof(tm_b));
    if (ntiles > 0)
        if (int rc = gram_i8_make_maps(g, &tm_a, &tm_b)) return rc;
    GemmArgs yr_copy = yr_args;
    if (ntiles == 0) yr_copy.M = 0;
    drv::DriveArgs dv = drive_args;
    RecurArgs none = {};
    int rec_ctas = 0, team = 1, member_bytes = 0;
    const int smem = std::max(gi8::SMEM_BYTES, yr::smem_bytes<FUSED_I8_THREADS>());
    if (ntiles > 0) {
        XESN_REQUIRE(tile_counter, "Gram workers: no tile counter");
        if (zero_counter) XESN_CUDA_OK(cudaMemsetAsync(tile_counter, 0, sizeof(int), stream));
    }
    void* args[] = {&none, &rec_ctas, &team, &member_bytes, &g, &tm_a, &tm_b, (void*)&tiles, &ntiles, &yr_copy, &dv, &tile_counter};
    XESN_CUDA_OK(cudaLaunchKernel((void*)kern, dim3(n_workers), dim3(FUSED_I8_THREADS), args, (size_t)smem, stream));
    count_launch();
    return 0;
}

// One chunk launch of the cluster engine as TWO kernels side by side: the recurrence clusters on `stream`, the
// workers as a plain grid on `side`.  A single cluster launch keeps only 132 (cluster size 4) or 128 (size 8) of the 148
// SMs busy -- clusters cannot straddle GPCs -- which costs the Gram role a fifth of its SMs where it shares the chunk
// with many recurrence clusters; a plain grid has no such limit, and nothing in either kernel waits for the other.
// The recurrence kernel is enqueued first and `side` has the lower priority, so the clusters are placed before the
// workers fill the remaining SMs.
int launch_fused_cluster_split_i8(RecurArgs rp, int n_groups, int team, const int* tail_ptr_dev, int tail_cap,
                                  const gi8::GramI8Args& gp, const int2* tiles, int ntiles, const GemmArgs& yr_args,
                                  const drv::DriveArgs& drive_args, int num_sms, cudaStream_t stream, cudaStream_t side,
                                  cudaEvent_t ev_fork, cudaEvent_t ev_join, int* tile_counter) {
    const bool has_rec = rp.steps > 0, has_work = ntiles > 0 || drive_args.m > 0;
    if (!has_rec) return launch_gram_workers_i8(gp, tiles, ntiles, yr_args, drive_args, num_sms, stream, tile_counter, true);
    const int rec_ctas = team * n_groups;
    if (!has_work || rec_ctas >= num_sms)
        return launch_fused_cluster_i8(rp, n_groups, team, tail_ptr_dev, tail_cap, gp, tiles, ntiles, yr_args, drive_args, num_sms,
                                       stream, tile_counter, false, true);
    // the tile counter is shared by both kernels: zeroed once, before the fork
    if (ntiles > 0) {
        XESN_REQUIRE(tile_counter, "fused cluster chunk: no tile counter");
        XESN_CUDA_OK(cudaMemsetAsync(tile_counter, 0, sizeof(int), stream));
    }
    XESN_CUDA_OK(cudaEventRecord(ev_fork, stream));
    XESN_CUDA_OK(cudaStreamWaitEvent(side, ev_fork, 0));
    // XESN_B200_NO_JOIN=1: the recurrence CTAs do not help with the Gram tiles.  For profiler captures: ncu serialises the
    // two kernels, so the recurrence kernel -- run first, alone -- would otherwise take every tile and the worker grid none.
    const bool join = !getenv("XESN_B200_NO_JOIN");
    if (int rc = launch_fused_cluster_i8(rp, n_groups, team, tail_ptr_dev, tail_cap, join ? gp : gi8::GramI8Args{},
                                         join ? tiles : nullptr, join ? ntiles : 0, GemmArgs{}, drv::DriveArgs{}, num_sms, stream,
                                         tile_counter, true, false))
        return rc;
    int rc = launch_gram_workers_i8(gp, tiles, ntiles, yr_args, drive_args, num_sms - rec_ctas, side, tile_counter, false);
    XESN_CUDA_OK(cudaEventRecord(ev_join, side));
    XESN_CUDA_OK(cudaStreamWaitEvent(stream, ev_join, 0));
    return rc;
}

// units per CTA (multiple of 8) and units per thread of the pull-model recurrence spread over n_ctas CTAs
int pull_geometry(int N, int n_groups, int n_ctas, int* upc_out, int* upt_out) {
    const long long total = (long long)N * n_groups;
    const long long upc = (((total + n_ctas - 1) / n_ctas) + 7) & ~7LL;
    for (int upt = 1; upt <= 8; upt <<= 1) {
        if (N % upt) break;
        if ((long long)recur::PULL_THREADS * upt >= upc) {
            *upc_out = (int)upc, *upt_out = upt;
            return 0;
        }
    }
    return -2;
}

static int prefill_ring(RecurArgs& rp, int n_groups, cudaStream_t stream) {
    const long long total = (long long)rp.N * n_groups;
    rp.n_active_ctas = (int)((total + rp.units_per_cta - 1) / rp.units_per_cta);
    if (rp.hint) XESN_CUDA_OK(cudaMemsetAsync(rp.hint, 0, sizeof(unsigned), stream));
    XESN_CUDA_OK(cudaMemset2DAsync(rp.states, sizeof(double) * rp.state_group_stride, 0xFF,
                                   sizeof(double) * (size_t)(rp.steps + 1) * rp.N, n_groups, stream));
    return 0;
}

int launch_pull_forced(RecurArgs rp, int n_groups, int num_sms, cudaStream_t stream) {
    if (rp.steps <= 0) return 0;
    int upc = 0, upt = 0;
    if (const char* e = getenv("XESN_B200_PULL_CTAS")) num_sms = std::min(num_sms, std::max(1, atoi(e)));  // experiments
    if (pull_geometry(rp.N, n_groups, num_sms, &upc, &upt)) {
        set_error("pull recurrence: %d x %d units do not fit %d CTAs", n_groups, rp.N, num_sms);
        return -2;
    }
    rp.units_per_cta = upc;
    if (int rc = prefill_ring(rp, n_groups, stream)) return rc;
    const long long total = (long long)rp.N * n_groups;
    int grid = (int)((total + upc - 1) / upc);
    void (*kern)(RecurArgs, int) = upt == 1 ? pull_forced_kernel<1>
                                            : (upt == 2 ? pull_forced_kernel<2>
                                                        : (upt == 4 ? pull_forced_kernel<4> : pull_forced_kernel<8>));
    void* args[] = {&rp, &n_groups};
    XESN_CUDA_OK(cudaLaunchCooperativeKernel((void*)kern, dim3(grid), dim3(recur::PULL_THREADS), args, 0, stream));
    count_launch();
    return 0;
}

// rp.steps == 0 -> no recurrence; ntiles == 0 -> no Gram role
int launch_fused_pull_i8(RecurArgs rp, int n_groups, const gi8::GramI8Args& gp, const int2* tiles, int ntiles,
                         int num_sms, cudaStream_t stream) {
    int upc = 0, upt = 0;
    if (pull_geometry(rp.N, n_groups, num_sms, &upc, &upt)) {
        set_error("pull recurrence: %d x %d units do not fit %d CTAs", n_groups, rp.N, num_sms);
        return -2;
    }
    rp.units_per_cta = upc;
    if (rp.steps > 0)
        if (int rc = prefill_ring(rp, n_groups, stream)) return rc;
    if (rp.steps <= 0 && ntiles == 0) return 0;
    void (*kern)(RecurArgs, int, gi8::GramI8Args, const CUtensorMap, const CUtensorMap, const int2*, int) =
        upt == 1 ? fused_pull_i8_kernel<1>
                 : (upt == 2 ? fused_pull_i8_kernel<2> : (upt == 4 ? fused_pull_i8_kernel<4> : fused_pull_i8_kernel<8>));
    XESN_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, gi8::SMEM_BYTES));
    gi8::GramI8Args g = gp;
    alignas(64) CUtensorMap tm_a, tm_b;
    memset(&tm_a, 0, sizeof(tm_a));
    memset(&tm_b, 0, sizeof(tm_b));
    if (ntiles > 0)
        if (int rc = gram_i8_make_maps(g, &tm_a, &tm_b)) return rc;
    void* args[] = {&rp, &n_groups, &g, &tm_a, &tm_b, (void*)&tiles, &ntiles};
    XESN_CUDA_OK(cudaLaunchCooperativeKernel((void*)kern, dim3(num_sms), dim3(FUSED_I8_THREADS), args,
                                             (size_t)gi8::SMEM_BYTES, stream));
    count_launch();
    return 0;
}

}  // namespace xesn

#ifdef XESN_RECUR_TRACE
extern "C" int xesn_debug_cta_ns(unsigned long long* out) {
    return (int)cudaMemcpyFromSymbol(out, xesn::g_cta_ns, sizeof(unsigned long long) * 2 * 160 * 2);
}
extern "C" int xesn_debug_trace_fused(long long* out) {
    return (int)cudaMemcpyFromSymbol(out, xesn::recur::g_recur_trace, sizeof(long long) * 64 * 8);
}
#endif
