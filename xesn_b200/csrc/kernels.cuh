// Argument blocks + launchers of the non-GEMM kernels (internal to the library).
#pragma once
#include "common.cuh"

namespace xesn {

struct RecurArgs {
    const int* indptr;       // CSR of W, shared by every group
    const int* indices;
    const double* values;
    long long values_gs;     // stride between the groups' value arrays (0: all groups share W)
    int param_groups;        // >= 1: group g uses parameter set g % param_groups
    const double* alpha_g;   // optional per-group leak rates (else `alpha`)
    const double* drive;     // [G][steps][N] : Win u_t + b for every step of the chunk
    long long drive_group_stride;
    double* states;          // [G][steps+1][N] ring: row 0 <- carry, rows 1..steps written (pre-filled "unset")
    long long state_group_stride;
    signed char* planes;     // optional [G][7][rows][ldp] int8 digit planes of the states (gram_i8.cuh); rows 0..steps-1
    long long ldp, plane_stride, plane_group_stride;
    int* status;             // device int, set to 1 if a poll timed out (co-residency lost: never expected)
    double* carry;           // [G][N] state entering the chunk; overwritten with the state leaving it
    int N;
    int steps;
    double alpha;
    // filled by the launcher
    int units_per_cta;
    int smem_nnz;            // capacity (entries) of the shared-memory CSR slice
    int w_in_smem;
    int n_pad;               // N rounded up to even: size of the shared-memory state copy
    int debug_no_exchange;   // timing experiments only: skip the refresh of the other CTAs' slices (wrong results)
    int poll_sleep_retry;    // pull model, experiments: nanoseconds to sleep between polls
    unsigned* hint;          // pull model: device counter of CTAs that have published their part of a row (zeroed per launch)
    // cluster engine: conflict-avoiding placement of the state copies in shared memory (bank_layout.cu)
    const int* slot;         // [N] 8-byte slot of state j
    const int* col_off;      // [nnz] byte offset of the slot of indices[k]
    int stagger;             // cluster engine: > 0: percentage of a CTA's warps in the early group; the others run a few
                             // hundred cycles behind (0: no staggering)
    int n_active_ctas;       // filled by the launcher
};

struct StepArgs {
    const int* indptr;
    const int* indices;
    const double* values;
    long long values_gs, win_gs, bias_gs;  // strides between parameter sets (0: shared parameters)
    int param_groups;                      // >= 1: group g uses parameter set g % param_groups
    const double* alpha_g;                 // optional per-set leak rates
    const double* Win;    // [N][I] (inline projection) -- used when drive == nullptr
    const double* bias;   // [N]
    const double* u;      // [G][I] dense inputs, or (when in_map != nullptr) gathered from `field`:
    const double* field;  //   u[g][k] = in_map[g*I+k] >= 0 ? field[in_map[..]] : fill[-1-in_map[..]]
    const int* in_map;
    const double* fill;
    const double* drive;  // [G][N] precomputed Win u + b, or nullptr
    const double* r_in;   // [G][N]
    double* r_out;        // [G][N]
    int N, I;
    double alpha;
    int mask_nan;         // NaN inputs read as 0 (LazyESN: xesn/lazyesn.py:362-364); 0 = NaN propagates (eager ESN)
};

struct ReadoutArgs {
    const double* Wout;   // [G][O][N]
    const double* r;      // [G][N]
    double* v;            // destination; element (g,o) lands at v[dst * v_stride]
    long long v_stride;
    double* v_copy;       // optional second destination (trajectory output)
    long long v_copy_stride;
    const int* out_map;   // optional [G][O] -> dst index; identity (g*O+o) when null
    int N, O;
    // split of the reservoir axis (filled by the caller from readout_slices(); 1 = no split)
    int n_slices;
    double* partial;      // [G][n_slices][O] scratch
    int* counters;        // [G][ceil(O/8)] zero-initialised once; the kernel re-arms them
    int slice;            // filled by the launcher
};

int launch_recur_forced(RecurArgs a, int n_groups, int max_slice_nnz, cudaStream_t stream);
int recur_cluster_size(int N);
// units owned by one CTA of a team: ceil(N / team) rounded up to a multiple of 8 (vector stores)
inline int recur_units_per_cta(int N, int team) { return ((N + team - 1) / team + 7) & ~7; }
// fused recurrence(c) || Gram(c-1) chunk kernel (fused_train.cu)
int fused_threads(int i8);  // CTA width of the fused chunk kernel (256 FP64-Gram, 512 int8-Gram)
int fused_recur_geometry(int N, int n_groups, int threads, int* team_out, int* upt_out, int* halves_out = nullptr);
int launch_fused_chunk(RecurArgs rp, int n_groups, int team, int upt, int max_slice_nnz, const GemmArgs& gp,
                       const int2* tiles, int ntiles, int num_sms, cudaStream_t stream);
namespace gi8 { struct GramI8Args; }
// same, with the Gram role on the int8 tensor cores (tcgen05 kind::i8) reading digit planes
int gram_i8_make_maps(const gi8::GramI8Args& a, void* map_a, void* map_b);  // two CUtensorMap (64-byte aligned)
// yr_args.M > 0: the Gram CTAs also run ybar += Y R of the previous chunk (skinny targets, yr_skinny.cuh);
// drive_args.m > 0: and the input projection Win u + b of the next chunk (few inputs)
namespace drv { struct DriveArgs; }
// tile_counter: device int from which the Gram CTAs -- and the recurrence CTAs once their chunk is done -- draw tiles
int launch_fused_chunk_i8(RecurArgs rp, int n_groups, int team, int upt, int halves, int max_slice_nnz,
                          const gi8::GramI8Args& gp, const int2* tiles, int ntiles, const GemmArgs& yr_args,
                          const drv::DriveArgs& drive_args, int num_sms, cudaStream_t stream, int* tile_counter);
// recurrence role on thread-block clusters (recur_cluster.cuh): state exchange through distributed shared memory
bool cluster_recur_geometry(int N, int n_groups, const int* indptr, int num_sms, int* team_out, int* tail_cap_out);
// the same engine for runs without a Gram role (spin-up, forced runs): the thinnest clusters that are all resident
bool cluster_forced_geometry(int N, int n_groups, const int* indptr, int num_sms, int* team_out, int* tail_cap_out);
void cluster_tail_ptr(int N, int team, const int* indptr, int* out);   // host table [team][units_per_cta + 1]
int launch_fused_cluster_i8(RecurArgs rp, int n_groups, int team, const int* tail_ptr_dev, int tail_cap,
                            const gi8::GramI8Args& gp, const int2* tiles, int ntiles, const GemmArgs& yr_args,
                            const drv::DriveArgs& drive_args, int num_sms, cudaStream_t stream, int* tile_counter = nullptr,
                            bool rec_grid_only = false, bool zero_counter = true);
// placement of the states in shared memory that avoids bank conflicts of the row gathers (bank_layout.cu): slot_out[N]
// in [0, n_slots), n_slots % 16 == 0; returns the gather wavefronts per step over all CTAs, *before = natural placement
long long plan_bank_layout(int N, const int* indptr, const int* indices, int upc, int kreg, int ktail, int n_slots,
                           int zero_slot, int* slot_out, long long* before);
namespace recur {
constexpr int CL_KREG = 12;   // entries of a row of W held in registers by the cluster engine ...
constexpr int CL_KHEAD = 8;   // ... of which the last CL_KREG - CL_KHEAD are skipped by warps whose rows are all shorter
constexpr int CL_KTAIL = 4;   // entries per round of the rows' tails (shared memory)
}
inline int cluster_state_slots(int N) { return (N + 15) / 16 * 16 + 32; }   // two spare slots per bank pair
// the same chunk as two kernels side by side (recurrence clusters on `stream`, plain worker grid on `side`)
int launch_fused_cluster_split_i8(RecurArgs rp, int n_groups, int team, const int* tail_ptr_dev, int tail_cap,
                                  const gi8::GramI8Args& gp, const int2* tiles, int ntiles, const GemmArgs& yr_args,
                                  const drv::DriveArgs& drive_args, int num_sms, cudaStream_t stream, cudaStream_t side,
                                  cudaEvent_t ev_fork, cudaEvent_t ev_join, int* tile_counter);
// pull-model recurrence (pull_body.cuh): spread over all SMs, beside the Gram role in every CTA
int pull_geometry(int N, int n_groups, int n_ctas, int* upc_out, int* upt_out);
int launch_pull_forced(RecurArgs rp, int n_groups, int num_sms, cudaStream_t stream);
int launch_fused_pull_i8(RecurArgs rp, int n_groups, const gi8::GramI8Args& gp, const int2* tiles, int ntiles,
                         int num_sms, cudaStream_t stream);
int launch_step_update(const StepArgs& a, int n_groups, cudaStream_t stream);
int readout_slices(int N, int O, int n_groups);
int launch_readout(ReadoutArgs a, int n_groups, cudaStream_t stream);
int launch_gather_vec(const double* src, long long src_stride, const int* map, const double* fill, double* out,
                      long long n, int mask_nan, cudaStream_t stream);
int launch_gather_series(const double* src, long long ld_src, long long t0, const int* map, const double* fill,
                         const unsigned char* zero_row, double* out, long long ld_out, long long rows, int nt,
                         cudaStream_t stream);

// ---- persistent closed-loop forecast (predict_loop.cu) ----------------------------------------
constexpr int XESN_MAX_WORLD = 8;
struct PredictLoopArgs {
    const int* indptr;
    const int* indices;
    const double* values;
    const double* Win;
    const double* bias;
    long long values_gs, win_gs, bias_gs;  // strides between parameter sets (0: shared parameters)
    int param_groups;                      // >= 1: group g uses parameter set g % param_groups
    const double* alpha_g;                 // optional per-set leak rates
    const double* Wout;     // [G][O][N]
    const int* in_map;      // [G][I] cell index or -1-k -> fill[k]
    const double* fill;
    const int* out_map;     // [G][O] cell index (nullptr: g*O+o)
    double* field[2];       // [n_cells] ping-pong; field[0] holds the initial condition
    double* r[2];           // [G][N] ping-pong; r[0] holds the state after spin-up
    double* pred;           // [n_cells][ldp] trajectory, column n written at step n
    long long ldp;
    int N, I, O, G, n_steps;
    double alpha;
    int n_slices, slice;    // split of the reservoir axis in the readout
    double* partial;        // [G][n_slices][O]
    int* counters;          // [G][ceil(O/8)] zero-initialised, re-armed by the kernel
    unsigned* bar;          // [2] grid barrier: arrival counter, generation
    int* status;
    int mask_nan;           // NaN inputs read as 0 (LazyESN) instead of propagating
    // multi-GPU exchange through peer-mapped memory (world == 1: unused)
    int world, rank;
    unsigned step_base;                      // flags count steps monotonically across launches
    double* peer_field[2 * XESN_MAX_WORLD];  // [parity][rank] -> that rank's field[parity]
    unsigned* peer_flags[XESN_MAX_WORLD];    // [rank] -> that rank's flag array (one entry per source rank)
};
int launch_predict_loop(const PredictLoopArgs& a, int num_sms, cudaStream_t stream);

// ---- dataflow closed loop of several reservoirs (predict_flow.cu) ------------------------------
constexpr int XESN_FLOW_MAX_TEAM = 16;   // CTAs per reservoir
constexpr int XESN_FLOW_CHUNK = 256;     // steps per launch = rows of the rings
struct PredictFlowArgs {
    const int* indptr;
    const int* indices;
    const double* values;
    const double* Win;
    const double* bias;
    long long values_gs, win_gs, bias_gs;
    int param_groups;
    const double* alpha_g;
    double alpha;
    const double* Wout;       // [G][O][N]
    int N, I, O, G;
    int team, upc;            // CTAs per group, units per CTA
    int cluster;              // the team is a thread-block cluster: states travel through distributed shared memory
    int steps, ring_steps;    // steps of this launch (<= ring_steps)
    int col0;                 // step n of this launch is trajectory column col0 + n
    const int* src;           // [G][I] ring slot feeding the input | -1-k: fill[k] | INT_MIN: static cell of `field`
    const int* in_map;        // [G][I] cell index (step 1 of a launch reads `field`)
    const double* fill;
    const double* field;      // [n_cells] field entering the launch
    double* field_out;        // [n_cells] field leaving it (local outputs + imported cells written)
    const int* out_map;       // [G][O] cell of every local output (nullptr: g*O+o)
    const int* import_cell;   // [n_import] cell of every imported ring slot (slots G*O ..)
    int n_import;
    const double* r_in;       // [G][N] states entering the launch
    double* r_out;            // [G][N] states leaving it
    double* r_ring;           // [G][ring_steps][N]              pre-filled "unset"
    double* part_ring;        // [G*O + n_import][ring_steps][team] pre-filled "unset"
    double* pred;             // [n_cells][ldp]
    long long ldp;
    const int* exp_ptr;       // [G*O + 1] exports of every local output (nullptr: none)
    const int* exp_peer;      // peer index of an export
    const int* exp_slot;      // ring slot in that peer's ring
    double* peer_ring[XESN_MAX_WORLD];  // the peers' partial rings (current parity) as mapped in this process
    int mask_nan;
    int* status;
    int tail_cap;             // shared-memory capacity (entries) for the rows' tails beyond the register-resident head
    const int* tail_ptr;      // [team][upc+1] offsets of every owned row's tail in that staging area
};
size_t predict_flow_smem(int N, int I, int O, int upc, int tail_cap, int cluster);
bool predict_flow_geometry(int N, int I, int O, int G, int num_sms, const int* indptr, int* team_out, int* upc_out,
                           int* tail_cap_out, int* cluster_out);
void predict_flow_tail_ptr(int N, int team, int upc, const int* indptr, int* out);
int launch_flow_src(const int* in_map, const int* out_map, int* owner, int* src, long long n_cells, long long n_slots,
                    long long n_inputs, cudaStream_t stream);
int launch_predict_flow(const PredictFlowArgs& a, cudaStream_t stream);

// ---- barrier-free closed loop of one reservoir (predict_pull.cu) --------------------------------
struct PredictPullArgs {
    const int* indptr;
    const int* indices;
    const double* values;
    const double* Win;      // [N][I]
    const double* bias;
    const double* Wout;     // [O][N]
    const int* in_map;      // [I] cell index or -1-k -> fill[k]
    const double* fill;
    const int* out_map;     // [O] cell index (nullptr: o)
    double* field;          // [n_cells] in: field entering the launch, out: field after `steps` steps
    double* r;              // [N] in/out
    double* r_ring;         // [steps+1][N]        pre-filled "unset" by the launcher
    double* part_ring;      // [steps][O][n_ctas]  pre-filled "unset" by the launcher
    double* v_ring;         // [steps][O]          pre-filled "unset" by the launcher
    double* pred;           // [n_cells][ldp]; column col0 + n written for n = 1..steps
    long long ldp;
    long long n_cells;
    int col0;
    int N, I, O, steps;
    double alpha;
    int upc, n_ctas;        // units per CTA, CTAs
    unsigned* hint;
    int* status;
    int mask_nan;           // NaN inputs read as 0 (LazyESN) instead of propagating
};
size_t predict_pull_smem(int I, int O, int upc, long long n_cells, int n_ctas);
bool predict_pull_geometry(int N, int I, int O, long long n_cells, int num_sms, int* upc_out, int* n_ctas_out);
int launch_predict_pull(PredictPullArgs a, cudaStream_t stream);

// ---- ridge solve (cholesky.cu) ---------------------------------------------------------------
// In-place blocked Cholesky of the lower triangle of `batch` SPD matrices A[b] (n x n, row-major,
// leading dimension lda), then X = A^{-1} B for B[b] (n x nrhs, row-major, ldb) in place.
// info[b] = 0 on success, k>0 if the leading minor of order k is not positive definite.
long long cholesky_workspace_doubles(int n, int nrhs, int batch);
// streams/events of the look-ahead factorisation: owned by a reservoir handle, created on first use on the
// device that is current then (one handle lives on one device)
struct SolveCtx {
    cudaStream_t side = nullptr;
    cudaEvent_t ev_panel[2] = {nullptr, nullptr}, ev_trail[2] = {nullptr, nullptr}, ev_fork = nullptr, ev_join = nullptr;
    int num_sms = 0;
};
void solve_ctx_destroy(SolveCtx& c);
int launch_cholesky_solve_ws(SolveCtx& ctx, double* A, long long lda, long long strideA, double* B, long long ldb, long long strideB,
                             int n, int nrhs, int batch, int* info, double* work, cudaStream_t stream);
// pivoted fallback (lu_solve.cu): one system, lower triangle of A valid (beta on the diagonal), B = nrhs x n rows
long long lu_workspace_doubles(int n);
int launch_lu_solve(double* A, long long lda, double* B, long long ldb, int n, int nrhs, int* info, double* work,
                    cudaStream_t stream);
int launch_add_diag(double* A, long long lda, long long strideA, int n, int batch, const double* beta_dev,
                    double beta_scalar, cudaStream_t stream);

}  // namespace xesn
