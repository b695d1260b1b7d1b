/* xesn_b200 -- C ABI of the B200-native ESN hot path.
 *
 * Drop-in boundary for the numerical layer of timothyas/xesn (v0.2.2).  The reference has no
 * FFI of its own (it is pure Python over numpy/cupy); each entry point below replaces one of
 * its hot-path functions, cited as <file>:<lines> relative to the upstream tree.  The Python
 * classes `xesn_b200.ESN` / `xesn_b200.LazyESN` bind these through ctypes.
 *
 * Conventions
 *   - every function returns 0 on success, a CUDA error code (>0) or -1/-2 (bad argument /
 *     unsupported size) on failure; `xesn_last_error()` describes the last failure;
 *   - all `*_dev` pointers are DEVICE pointers owned by the caller (e.g. torch tensors);
 *     `stream` is a `cudaStream_t` passed as `void*` (NULL = default stream); calls are
 *     asynchronous with respect to the host unless stated otherwise;
 *   - all floating-point data is IEEE binary64, row-major; "series" arrays are time-LAST like the
 *     reference's `(n_vars, n_time)` DataArrays;
 *   - G = number of independent reservoirs ("groups": LazyESN chunks, the candidates of an ESNEnsemble, or 1 for
 *     the eager ESN); all groups share one (W, Win, b) as in xesn/lazyesn.py:175-177 unless
 *     xesn_reservoir_set_group_params installed per-group parameter sets;
 *   - handles are not thread safe; one handle lives on one device.
 */
#ifndef XESN_B200_H
#define XESN_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct xesn_reservoir xesn_reservoir;

/* Rows of time series, optionally gathered through an index map (the halo of a LazyESN group,
 * xesn/lazyesn.py:165-167, or the identity for the eager ESN).
 *   row r of group g, time t  =  data[src(g,r) * ld + t]
 *   src(g,r) = map ? map[g*rows + r] : g*rows + r ;  map value m < 0 -> constant fill[-1-m]
 *   zero_row[g*rows + r] != 0 -> the row reads as 0 for all t (cell masked because it is NaN at
 *   the first time step: xesn/lazyesn.py:287-289, :362-364). */
typedef struct {
    const double* data_dev;
    int64_t ld;
    const int32_t* map_dev;          /* may be NULL */
    const double* fill_dev;          /* may be NULL when map has no negative entries */
    const uint8_t* zero_row_dev;     /* may be NULL */
} xesn_series;

const char* xesn_last_error(void);
int xesn_abi_version(void);

/* ---- reservoir handle: uploads W (CSR, HOST arrays, column indices sorted within rows),
 *      Win (HOST, [N][I] row-major) and the bias vector (HOST, [N]).
 *      Replaces the attributes set by ESN.build, xesn/esn.py:147-187. */
int xesn_reservoir_create(xesn_reservoir** out, int device, int32_t n_reservoir, int32_t n_input,
                          const int32_t* indptr, const int32_t* indices, const double* values,
                          const double* win, const double* bias, double leak_rate);
int xesn_reservoir_destroy(xesn_reservoir* h);
/* Per-group parameters: n_groups independent reservoirs that share the sparsity pattern of W (indptr, indices) but
 * have their own values of W, Win, bias and leak rate -- the candidates of a hyper-parameter search, which the
 * reference rebuilds one by one from the same seeds with other scaling factors (xesn/cost.py:190-203).  Host arrays
 * values[n_groups][nnz], win[n_groups][N*I], bias[n_groups][N], leak[n_groups] are copied.  Afterwards every entry
 * point must be called with a multiple of this n_groups: group g uses parameter set g % n_groups (several forecast
 * windows of every candidate in one loop).  The per-group Tikhonov parameters go to xesn_ridge_solve. */
int xesn_reservoir_set_group_params(xesn_reservoir* h, int32_t n_groups, const double* values, const double* win,
                                    const double* bias, const double* leak);
/* Synchronous health check, read-and-clear: *status_host = 0 if every persistent-kernel poll completed since the
 * last call, 1 if one timed out (the CTAs of a team lost co-residency or a peer rank died; results of that run
 * are invalid).  Reporting clears the flag, so the next run on the handle starts clean. */
int xesn_reservoir_status(xesn_reservoir* h, int32_t* status_host);
/* Upper bound of any single in-kernel wait, by wall clock (%globaltimer), default 2000 ms
 * (XESN_B200_POLL_TIMEOUT_MS at creation).  Raise it under time-slicing, MPS or a profiler. */
int xesn_set_poll_timeout_ms(xesn_reservoir* h, int32_t milliseconds);
/* Closed-loop / xesn_update input handling: on = NaN inputs read as 0, the LazyESN rule
 * (xesn/lazyesn.py:357-365 _prepare_1d_inputs); off (default) = NaN propagates like the eager
 * ESN.predict / _update of the reference (xesn/esn.py:251-268, :458-460). */
int xesn_set_nan_input_mask(xesn_reservoir* h, int32_t on);

/* ---- a1: one update, r' = (1-a) r + a tanh(W r + Win u + b)      xesn/esn.py:458-460 (_update),
 *      xesn/lazyesn.py:319-322 (_update_nd).  r_in/r_out: [G][N]; u: [G][I]; NaN inputs propagate unless
 *      xesn_set_nan_input_mask(h, 1). */
int xesn_update(xesn_reservoir* h, const double* r_in_dev, const double* u_dev, double* r_out_dev,
                int32_t n_groups, void* stream);

/* ---- teacher-forced run of `n_steps` updates from `r_carry_dev` ([G][N], updated in place)
 *      on inputs u(t0 .. t0+n_steps-1).  xesn/esn.py:498-499 (spin-up), :506-507 (training
 *      states), xesn/lazyesn.py:301-316 (_spinup).  If states_dev != NULL it receives
 *      [G][n_steps+1][N]: row k = state before consuming u(t0+k) (the reference's rT buffer).
 *      scratch_dev: xesn_forced_scratch_bytes() bytes. */
int64_t xesn_forced_scratch_bytes(xesn_reservoir* h, int32_t n_groups, int32_t n_steps, int32_t gathered);
int xesn_forced_run(xesn_reservoir* h, const xesn_series* u, int64_t t0, int32_t n_steps, int32_t n_groups,
                    double* r_carry_dev, double* states_dev, void* scratch_dev, int64_t scratch_bytes,
                    void* stream);

/* ---- a2/a5: ridge training for G groups                     xesn/esn.py:463-520 (_train_1d),
 *      xesn/lazyesn.py:264-298 (_train_nd).  Streams the time axis in chunks of `chunk` steps:
 *      drive GEMM -> persistent recurrence -> Gram (R^T R, lower triangle) and Y R accumulation on
 *      the FP64 tensor pipe; the state history is never materialised beyond one chunk.
 *      y rows are paired with the state that has seen u(0..t-1) exactly as the reference does.
 *        gram_dev : [G][N][N] workspace (lower triangle = R R^T + beta I on return from
 *                   xesn_train_accumulate; destroyed by xesn_ridge_solve)
 *        wout_dev : [G][O][N]; holds Y R^T after accumulate, Wout after solve
 *        info_dev : [G] ints, 0 = ok, k>0 = leading minor k not positive definite
 *      INTERLEAVED LAYOUT: when wout_dev == gram_dev + N*N (in doubles) the two are taken as ONE array [G][N+O][N], the
 *      O readout rows of a group right below its Gram matrix (group stride (N+O)*N for both).  Pass the same two
 *      pointers to xesn_train_accumulate and xesn_ridge_solve. */
int64_t xesn_train_scratch_bytes(xesn_reservoir* h, int32_t n_output, int32_t n_groups, int32_t chunk,
                                 int32_t gathered_u, int32_t gathered_y);
int xesn_train_accumulate(xesn_reservoir* h, const xesn_series* u, const xesn_series* y, int32_t n_output,
                          int32_t n_groups, int64_t n_time, int64_t n_spinup, int32_t chunk,
                          double* gram_dev, double* wout_dev, void* scratch_dev, int64_t scratch_bytes,
                          void* stream);
/* (R R^T + beta I) solve: xesn/esn.py:516-520.  beta_dev: [G] or NULL (then beta_scalar is used).
 *      Blocked Cholesky; for one large system the trailing updates run on the int8 tensor cores (row-scaled
 *      8-digit fixed point, XESN_B200_SOLVE_I8=0 keeps FP64 DMMA).  With the interleaved layout (above) and
 *      n_output <= 256 the right-hand sides ride through the factorisation as extra rows: no forward substitution. */
int xesn_ridge_solve(xesn_reservoir* h, int32_t n_output, int32_t n_groups, const double* beta_dev,
                     double beta_scalar, double* gram_dev, double* wout_dev, int32_t* info_dev,
                     void* scratch_dev, int64_t scratch_bytes, void* stream);

/* Pivoted fallback for ONE system (n_groups == 1 layout: gram_dev [N][N] with the lower triangle valid and WITHOUT beta,
 *      wout_dev [O][N] = Y R^T) whose Cholesky factorisation reported info > 0: blocked LU with partial pivoting, the
 *      factorisation the reference's GPU path uses (cupy getrf/getrs; scipy's assume_a="sym" is the pivoted LDL^T),
 *      xesn/esn.py:516-520.  The caller re-runs xesn_train_accumulate for that group first (the failed factorisation
 *      destroyed the Gram matrix).  info_dev: one int, 0 or j+1 when column j has no usable pivot (exactly singular). */
int xesn_ridge_solve_pivoted(xesn_reservoir* h, int32_t n_output, double beta, double* gram_dev, double* wout_dev,
                             int32_t* info_dev, void* scratch_dev, int64_t scratch_bytes, void* stream);

/* ---- a3/a6: closed-loop forecast                            xesn/esn.py:264-268 (ESN.predict),
 *      xesn/lazyesn.py:226-236 (+ _update_nd, _readout, overlap per step).
 *      The domain is a flat vector of `n_cells` values.  in_map [G][I] gathers every group's
 *      (halo'd) input from it (negative -> fill constant; NaN -> 0 with xesn_set_nan_input_mask); out_map [G][O] scatters the
 *      readout back.  Eager ESN: G=1, identity maps.
 *        field_dev  : [n_cells] initial condition (truth at n_spinup); overwritten with the last step
 *        r_dev      : [G][N] reservoir states after spin-up; updated in place
 *        pred_dev   : [n_cells][n_steps+1] (time last); column 0 is written with the initial field
 *      scratch: xesn_predict_scratch_bytes().
 *      Engines (same results to rounding of the readout sums): one reservoir (n_groups == 1, shared parameters)
 *      runs the barrier-free kernel of predict_pull.cu (every CTA publishes its slice of r and its contribution
 *      to each output, one reducer warp per output); several reservoirs run the dataflow kernel of predict_flow.cu
 *      (a team of CTAs per reservoir, consumers poll the values they read, no grid barrier) when a team's slices of
 *      Win / Wout fit shared memory, else the persistent kernel with software grid barriers (predict_loop.cu);
 *      configurations whose input projection is too large to inline fall back to one update + one readout launch
 *      per step.  XESN_B200_PREDICT=barrier|steps forces the latter two. */
int64_t xesn_predict_scratch_bytes(xesn_reservoir* h, int32_t n_output, int32_t n_groups, int64_t n_cells);
int xesn_predict(xesn_reservoir* h, const double* wout_dev, int32_t n_output, int32_t n_groups,
                 int64_t n_cells, const int32_t* in_map_dev, const double* fill_dev,
                 const int32_t* out_map_dev, double* field_dev, double* r_dev, int32_t n_steps,
                 double* pred_dev, void* scratch_dev, int64_t scratch_bytes, void* stream);
/* one closed-loop step (used by the multi-GPU LazyESN driver, which exchanges halos between
 * steps): r <- update(r, gather(field_in)); field_out[out_map] = Wout r; optionally also
 * pred_col_dev[cell * pred_stride]. */
int xesn_predict_step(xesn_reservoir* h, const double* wout_dev, int32_t n_output, int32_t n_groups,
                      const int32_t* in_map_dev, const double* fill_dev, const int32_t* out_map_dev,
                      const double* field_in_dev, double* field_out_dev, double* r_dev,
                      double* pred_col_dev, int64_t pred_stride, void* scratch_dev, int64_t scratch_bytes,
                      void* stream);

/* ---- multi-GPU closed loop with the exchange inside the kernel (one process per GPU) ----
 * xesn_p2p_alloc / _open / _close / _free: device buffers that peers map through CUDA IPC
 * (handle64 = the 64-byte cudaIpcMemHandle_t, exchanged by the caller, e.g. torch.distributed).
 * xesn_predict_p2p: like xesn_predict, but every rank runs ONE persistent kernel; the readout stores
 * each predicted cell into the field buffer of every rank (NVLink P2P stores) and the ranks
 * synchronise once per step on peer-mapped flags.  field_ptrs[parity*world + k]: rank k's field
 * buffer of that parity as mapped in this process (n_cells doubles; parity 0 holds the initial
 * condition on every rank), flag_ptrs[k]: rank k's flag array (world uint32, zero at allocation).
 * step_base: number of forecast steps already run on these flags.  The last field is in
 * parity n_steps&1.  Replaces the per-step dask overlap of xesn/lazyesn.py:235. */
int xesn_p2p_alloc(int64_t bytes, void** ptr_out, uint8_t* handle64);
int xesn_p2p_open(const uint8_t* handle64, void** ptr_out);
int xesn_p2p_close(void* ptr);
int xesn_memcpy_async(void* dst_dev, const void* src_dev, int64_t bytes, void* stream); /* device-to-device */
int xesn_p2p_free(void* ptr);
int xesn_predict_p2p(xesn_reservoir* h, const double* wout_dev, int32_t n_output, int32_t n_groups, int64_t n_cells,
                     const int32_t* in_map_dev, const double* fill_dev, const int32_t* out_map_dev, double* r_dev,
                     int32_t n_steps, double* pred_dev, int32_t world, int32_t rank, uint32_t step_base,
                     const uint64_t* field_ptrs, const uint64_t* flag_ptrs, void* scratch_dev, int64_t scratch_bytes,
                     void* stream);

/* ---- multi-GPU closed loop as a dataflow with neighbour-only exchange (predict_flow.cu) ----
 * Replaces the per-step `overlap` of xesn/lazyesn.py:235 (and :226-236 around it) for several reservoirs per rank.
 * A team of CTAs owns a group; per step every CTA publishes its slice of the state and its contribution to each
 * output of the group into rings in global memory (the data is its own ready flag), and consumers poll exactly the
 * values they read.  Cells that groups of OTHER ranks read are pushed into those ranks' rings with NVLink P2P
 * stores -- only those cells, only to those ranks.  xesn_predict uses the same kernel for n_groups > 1 on one GPU.
 *   ring slot s < G*O: local output (g, o) = (s / O, s % O); slot G*O + j: imported cell import_cell[j].
 *   src[g][k]: slot feeding input k of local group g | -1-f: constant fill[f] | INT32_MIN: cell nobody produces.
 *   exports of local slot s: entries exp_ptr[s] .. exp_ptr[s+1]-1 -> (exp_peer[e] = rank, exp_slot[e] = slot there).
 *   ring_ptrs[parity*world + k]: rank k's partial ring of that parity as mapped in this process (k == rank: local;
 *   ranks this one exports nothing to may be 0).  One ring = (G*O + n_import) * ring_steps * team doubles
 *   (xesn_flow_geometry), allocated with xesn_p2p_alloc and filled with 0xFF bytes (xesn_memset_async) on every
 *   rank BEFORE any rank's first call; the library re-arms a ring right after each launch that used it.
 *   launches_done: zero at plan creation, advanced by the library; every rank makes the same calls. */
typedef struct {
    int32_t n_import;
    const int32_t* src_dev;
    const int32_t* import_cell_dev;
    const int32_t* exp_ptr_dev;      /* may be NULL: nothing exported */
    const int32_t* exp_peer_dev;
    const int32_t* exp_slot_dev;
    int32_t world, rank;
    uint64_t ring_ptrs[16];
    uint32_t launches_done;
} xesn_flow_plan;
/* 0 and (team, ring_steps) if n_groups reservoirs fit the dataflow kernel on this device, else -2 */
int xesn_flow_geometry(xesn_reservoir* h, int32_t n_output, int32_t n_groups, int32_t* team_out, int32_t* ring_steps_out);
/* arguments as xesn_predict (in_map/out_map index the GLOBAL cell vector; field_dev holds every cell this rank reads) */
int xesn_predict_flow(xesn_reservoir* h, xesn_flow_plan* plan, const double* wout_dev, int32_t n_output, int32_t n_groups,
                      int64_t n_cells, const int32_t* in_map_dev, const double* fill_dev, const int32_t* out_map_dev,
                      double* field_dev, double* r_dev, int32_t n_steps, double* pred_dev, void* scratch_dev,
                      int64_t scratch_bytes, void* stream);
int xesn_memset_async(void* dst_dev, int32_t byte_value, int64_t bytes, void* stream);

/* ---- building blocks exposed for parity tests and benchmarks ---- */
/* C = alpha * op(A) op(B) + beta * C on the FP64 tensor pipe; flags: 1 = A is [M][K] (else [K][M]),
 * 2 = B is [N][K] (else [K][N]), 4 = only lower-triangular tiles (square C). */
int xesn_gemm_f64(const double* a_dev, int64_t lda, const double* b_dev, int64_t ldb, double* c_dev, int64_t ldc,
                  int32_t m, int32_t n, int32_t k, double alpha, double beta, int32_t flags, void* stream);
/* gram_dev (N x N, lower triangle) += S^T S for S = states_dev [n_rows][N], |S| <= 1, computed
 * on the int8 tensor cores (tcgen05 kind::i8) from the fixed-point digits of S: absolute error <= n_rows * 2^-50
 * per entry (FP64-DGEMM level for |S| ~ 1, exact for S on a 2^-42 grid), whatever the magnitude of S.  Allocates
 * its own scratch and synchronises: a test / measurement entry (the training pipeline runs the same
 * device code inside the fused chunk kernel).  Replaces xesn/esn.py:510. */
int xesn_gram_i8(const double* states_dev, int64_t lds, int32_t n_rows, int32_t n_reservoir, double* gram_dev,
                 int64_t ldg, void* stream);
/* Host-side helper (no GPU needed): the placement of the reservoir state in shared memory that the recurrence
 * kernels use.  slot_out[j] (a permutation into [0, n_slots), n_slots a multiple of 16 and >= n_reservoir) is chosen so
 * that the states one half-warp gathers for W r (xesn/esn.py:459, one row of W per thread, rows of a CTA in order,
 * `units_per_cta` rows per CTA) fall into different banks.  Returns the shared-memory wavefronts of those gathers per
 * time step, summed over the CTAs, for the chosen placement and (in *natural_out) for slot[j] = j; negative on error.
 * The arithmetic does not depend on the placement.  Exposed for tests and measurements. */
int64_t xesn_bank_layout(int32_t n_reservoir, const int32_t* indptr, const int32_t* indices, int32_t units_per_cta,
                         int32_t n_slots, int32_t* slot_out, int64_t* natural_out);
/* readout v[g][o] = Wout[g][o][:] . r[g][:]   (xesn/esn.py:268, xesn/lazyesn.py:325-327) */
int xesn_readout(const double* wout_dev, const double* r_dev, double* v_dev, int32_t n_reservoir,
                 int32_t n_output, int32_t n_groups, void* stream);
/* ---- cost of a batch of forecasts: the NRMSE term of xesn/cost.py:227-254 and the
 *      per-candidate average + clamp of xesn/cost.py:201-225.
 *        pred_dev  : [n_windows][n_groups][n_output][n_times] forecasts (xesn_predict with n_windows * n_groups groups)
 *        truth_dev : row (f, o) = truth_dev[(f*n_output + o) * truth_ld + t0 + s], s < n_times
 *        nrmse_fg_dev : [n_windows][n_groups] out;  cost_dev : [n_groups] out = mean over the windows, with NaN, inf and
 *                       values above upper_bound replaced by upper_bound */
int xesn_cost_nrmse(const double* pred_dev, const double* truth_dev, int64_t truth_ld, int64_t t0, int32_t n_windows,
                    int32_t n_groups, int32_t n_output, int32_t n_times, double upper_bound, double* nrmse_fg_dev,
                    double* cost_dev, void* stream);
/* the PSD term of the cost for fields with ONE spatial axis (xesn/cost.py:257-272 over xesn/psd.py:12-65): NRMSE of the
 * 1-D power spectral density, per window and candidate; same array conventions as xesn_cost_nrmse.  The caller combines
 * the terms (cost.py:204-210: factor-weighted sum per window), averages over the windows and clamps. */
int xesn_cost_psd_nrmse(const double* pred_dev, const double* truth_dev, int64_t truth_ld, int64_t t0, int32_t n_windows,
                        int32_t n_groups, int32_t n_output, int32_t n_times, double* psd_nrmse_fg_dev, void* stream);
/* Radially binned power spectrum of fields with two or three spatial axes (xesn/psd.py:42-65): spec_dev holds the
 * 2-D FFT of every image as interleaved (re, im) doubles, [n_images][n_z][n_pix]; out[img][b] = factor[b] * mean over
 * the spectral cells pix_idx[bin_ptr[b] .. bin_ptr[b+1]) of (mean over z of |spec|^2); NaN for an annulus without cells
 * (scipy.stats.binned_statistic).  The cell lists depend only on the grid size and come from the host. */
int xesn_psd_radial(const double* spec_dev, int64_t n_images, int32_t n_z, int32_t n_pix, const int32_t* bin_ptr_dev,
                    const int32_t* pix_idx_dev, const double* factor_dev, int32_t n_bins, double* out_dev, void* stream);
/* NRMSE with xarray's skip-NaN reductions between pred [n_windows][n_groups][n_rows][n_times] and truth
 * [n_windows][n_rows][n_times] (the psd_nrmse of xesn/cost.py:257-272 applied to spectra that may hold NaN bins). */
int xesn_nrmse_skipnan(const double* pred_dev, const double* truth_dev, int32_t n_windows, int32_t n_groups,
                       int32_t n_rows, int32_t n_times, double* nrmse_fg_dev, void* stream);
/* number of kernel launches issued by this library since load (bench.py's gpu_launches) */
int64_t xesn_launch_count(void);
/* per-phase device timers (CUDA events on the launching stream).  enable(1) clears and starts
 * recording, enable(0) stops.  read() synchronises on the recorded events and returns the number
 * of timed regions and their summed duration for class
 * 0 drive GEMM, 1 recurrence, 2 Gram SYRK, 3 Y.R GEMM, 4 ridge solve, 5 closed loop, 6 halo gather,
 * 7 fused chunk kernel (recurrence of chunk c beside the Gram of chunk c-1). */
/* The training pipeline normally runs recurrence(c) and Gram(c-1) side by side in one cooperative
 * SM-specialised kernel; xesn_set_fused(h, 0) falls back to separate launches (A/B measurements). */
int xesn_set_fused(xesn_reservoir* h, int32_t on);
/* Recurrence engine: 0 = team model (a team of CTAs per reservoir, full state copy in shared memory),
 * 1 (default) = pull model (every thread polls exactly the states its rows of W need, spread over all SMs) for
 * forced runs and spin-up, team model inside the fused training kernel, 2 = pull model there too (run by the
 * idle warps of the Gram CTAs).  Same arithmetic, bitwise the same results. */
int xesn_set_recur_mode(xesn_reservoir* h, int32_t mode);
/* Gram accumulation engine of the fused pipeline: 1 (default) = fixed-point int8 digit planes on the
 * tcgen05 tensor cores (needs |r| <= 1), 0 = FP64 DMMA.  xesn_get_gram_mode returns the engine that will
 * actually run: a leak rate outside [0, 1] (accepted by the reference) lets |r| exceed 1, so such
 * reservoirs always use the FP64 engine. */
int xesn_set_gram_mode(xesn_reservoir* h, int32_t mode);
int xesn_get_gram_mode(xesn_reservoir* h, int32_t* mode_out);
int xesn_profile_enable(int32_t on);
int xesn_profile_read(int32_t cls, int64_t* n_regions, double* total_ms);
/* the same, region by region in launch order: fills ms_out[0 .. min(n, capacity)) and returns n (or a negative code) */
int64_t xesn_profile_regions(int32_t cls, double* ms_out, int64_t capacity);

#ifdef __cplusplus
}
#endif
#endif /* XESN_B200_H */
