/*
 * jabble_b200.h -- C ABI of libjabble_b200.so: the B200 (sm_100a) implementation of Jabble's
 * loss / gradient / RV-Fisher hot path.
 *
 * The reference (mdd423/wobble_jax) has no FFI layer: the seam this library replaces is the pair
 * of Python callables handed to scipy in jabble/model.py:226-232
 *     func   = loss.loss_all                  (jabble/loss.py:47-75)
 *     fprime = jax.grad(loss.loss_all)        (jabble/model.py:228)
 * plus EpochSpecificModel.f_info (jabble/model.py:857-901) and Model.__call__
 * (jabble/model.py:138-160).  Everything below those calls -- Shifting/Stretching warp,
 * CardinalSplineMixture evaluation, Additive/Composite combination, ChiSquare, the reverse pass,
 * the per-epoch Fisher sums -- runs inside this library.  INTEGRATION.md shows the ctypes
 * binding a maintainer of the reference would add.
 *
 * Conventions: plain pointers and sizes only; every function returns 0 on success or a negative
 * jb_status; jb_last_error() gives the message for the calling thread.  No exceptions cross the
 * boundary and there is NO CPU fallback: without a CUDA device every compute entry point fails
 * with JB_ERR_CUDA.  Host buffers are owned by the caller and never retained after the call,
 * except that jb_plan_create copies the data block into HBM once (the reference bakes the data
 * into the jitted program per optimize() call, loss.py:47).
 *
 * Threading: calls on one plan are serial (scipy is serial); different plans may be driven from
 * different host threads.  Each plan owns one CUDA stream.
 */
#ifndef JABBLE_B200_H
#define JABBLE_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define JB_ABI_VERSION 1

typedef enum {
    JB_OK = 0,
    JB_ERR_BAD_ARG = -1,      /* null pointer, negative size, inconsistent descriptor            */
    JB_ERR_UNSUPPORTED = -2,  /* model shape outside the compiled grammar (see jb_component_desc) */
    JB_ERR_CUDA = -3,         /* CUDA runtime error, or no device                                 */
    JB_ERR_OUT_OF_RANGE = -4, /* an unmasked warped pixel falls outside the knot grid (the         */
                              /* reference clamps the gather silently, model.py:985)              */
    JB_ERR_WINDOW = -5,       /* knot window of a pixel tile exceeds the on-chip capacity          */
    JB_ERR_NONFINITE = -6     /* loss is NaN/Inf                                                   */
} jb_status;

typedef enum {
    JB_COMP_SPLINE = 0,       /* CardinalSplineMixture: one control-point vector (model.py:991)     */
    JB_COMP_NORM_SPLINE = 1   /* NormalizationModel over a CardinalSplineMixture: one control-point */
                              /* vector PER KEY (model.py:1110-1143)                                */
} jb_comp_kind;

/*
 * One additive component  a[k_a(r)] * Spline(x - d[k_d(r)]; ctrl)   (CompositeModel.call,
 * model.py:670-675, of [ShiftingModel?, CardinalSplineMixture|NormalizationModel, StretchingModel?]).
 * All *_offset fields index the plan's flat parameter vector ("params_all": every parameter of
 * every leaf model, fit or fixed, in DFS order of the tree); -1 = that leaf is absent.
 * Key arrays are (n_rows,) int32: the metablock column named by the leaf's which_key
 * (dataset.py:214-235); NULL = row index.
 */
typedef struct {
    int32_t kind;               /* jb_comp_kind                                                  */
    int32_t p_val;              /* spline degree, 1..3                                           */
    double xp0;                 /* xs[0] of the knot grid (model.py:979)                          */
    double dx;                  /* xs[1] - xs[0]          (model.py:974)                          */
    int64_t n_knots;            /* M (SPLINE) or K = knots per key (NORM_SPLINE)                  */
    int64_t ctrl_offset;        /* control points: M values, or n_ctrl_keys * K values            */
    int64_t n_ctrl_keys;        /* NORM_SPLINE: NormalizationModel.size; SPLINE: 0                */
    const int32_t *ctrl_key;    /* NORM_SPLINE: row -> key                                        */
    int64_t shift_offset;       /* ShiftingModel.p (model.py:927: x - p[key]) or -1               */
    int64_t n_shift;
    const int32_t *shift_key;
    int64_t stretch_offset;     /* StretchingModel.p (model.py:946: p[key] * x) or -1             */
    int64_t n_stretch;
    const int32_t *stretch_key;
} jb_component_desc;

/* The data block of Data.blockify (dataset.py:192-237) plus the compiled model. */
typedef struct {
    int32_t abi_version;        /* JB_ABI_VERSION                                                */
    int32_t device;             /* CUDA device ordinal                                           */
    int64_t n_rows;             /* E                                                             */
    int64_t n_pix;              /* P (padded row length of the block)                            */
    const double *xs;           /* (E,P) row-major log-wavelength                                */
    const double *ys;           /* (E,P) log-flux; may be NULL for forward-only plans            */
    const double *yivar;        /* (E,P) inverse variance; may be NULL for forward-only plans    */
    const uint8_t *mask;        /* (E,P) 1 = ignore pixel; may be NULL (nothing masked)          */
    int32_t n_components;
    const jb_component_desc *components;
    int64_t n_params;           /* length of params_all                                          */
    double coefficient;         /* LossFunc.coefficient (loss.py:31)                             */
    int32_t rows_per_group;     /* 0 = choose; rows accumulated on chip before a partial is written */
    int32_t flags;              /* JB_PLAN_* bits; 0 = defaults                                      */
} jb_plan_desc;

/* jb_plan_desc.flags */
#define JB_PLAN_FLOAT32 2          /* float32 tier: the data block is kept as float (x relative to its
                                      tile, 13 B/pixel) and the fused kernel computes in float; results
                                      within 1e-5 of the float64 ones.  Needs smooth tiles and no
                                      NormalizationModel component; jb_plan_grid_search and _curvature are
                                      float64 paths and not available on such a plan                  */
#define JB_PLAN_NO_STRIP 4         /* never evaluate with the column-walk kernel (k_strip): keep k_fused2 /
                                      k_fused even when the rows are many and regular enough for it
                                      (cross-checks and A/B measurements)                               */
#define JB_PLAN_GENERAL_KERNEL 1   /* always evaluate with the general fused kernel (k_fused), never the
                                      specialised one (k_fused2): cross-checks and A/B measurements   */

typedef struct jb_plan jb_plan;

/* Numbers a caller needs for benchmarking and diagnostics. */
typedef struct {
    int64_t n_rows, n_pix, n_pix_padded;
    int64_t n_params, n_fit;
    int64_t n_tasks;              /* (row group x pixel tile) work items of the fused kernel       */
    int32_t rows_per_group;
    int32_t tile_pixels;
    int64_t device_bytes;         /* HBM held by the plan                                          */
    int64_t scratch_bytes;        /* of which per-evaluation scratch (partials)                     */
    int64_t algorithmic_bytes;    /* E*P*25 + 4*M*8 + E*7*8 for the compiled model     */
    int64_t kernel_launches;      /* kernels launched by this plan since creation                   */
    int64_t evaluations;
    int64_t serial_rows;          /* rows that needed the serial (conflict) path in the last eval   */
    int64_t tiles_smooth;         /* (row, tile) pairs on the branch-free path                       */
    int64_t tiles_general;        /* ... on the general per-pixel path (pixels >= 1 knot apart)      */
    int64_t tiles_serial;         /* ... walked lane by lane (lanes of a class may share knots)      */
    int64_t tiles_empty;          /* ... without any weighted pixel                                  */
    int32_t window_knots;         /* on-chip gradient window per warp and component                  */
    int32_t fused_variant;        /* kernel of the last evaluation: -1 general k_fused, >= 0 flags of k_fused2;
                                     bit 6 (64) float32 tier k_fused32, bit 7 (128) column-walk kernel k_strip */
    int32_t strip_reason;         /* why k_strip does not serve the plan: 0 it does / not decided yet, 1 model shape or
                                     flags, 2 sorted rows more than a knot apart (too few epochs), 3 pixel grid differs
                                     from row to row, 4 more than 8 pixels per knot, 5 a task's knot window too wide,
                                     6 left at run time (the shifts drifted twice)                                   */
    int32_t strip_rows_per_block; /* K of k_strip (rows between two re-basings of a column), 0 when not in use          */
} jb_plan_info_t;

const char *jb_last_error(void);
int jb_abi_version(void);
int jb_device_count(void);   /* 0 when no CUDA device/driver is present */
/* Host helper of the data-format side (jabble/dataset.py:118-160, Data.blockify): out[i] = 64-bit
 * position-sensitive checksum of the nbytes[i] bytes at ptrs[i], for n host buffers (OpenMP over buffers).
 * The host side reuses a resident block only while every frame array is the same object with the same checksum. */
int jb_host_checksum(const void *const *ptrs, const int64_t *nbytes, int64_t n, uint64_t *out);

int jb_plan_create(const jb_plan_desc *desc, jb_plan **out);
void jb_plan_destroy(jb_plan *plan);
int jb_plan_info(const jb_plan *plan, jb_plan_info_t *info);

/* Upload the current value of EVERY parameter (fixed ones are read from here; Model.p). */
int jb_plan_set_params(jb_plan *plan, const double *params_all, int64_t n_params);
/* Which entries of params_all are in fit mode, in get_parameters() order (model.py:478-495). */
int jb_plan_set_fit(jb_plan *plan, const int64_t *fit_index, int64_t n_fit);

/*
 * One fused evaluation at the fit vector p_fit: loss (loss.py:47-75), gradient w.r.t. p_fit
 * (model.py:228) if grad != NULL, both written to host memory after a stream sync.
 */
int jb_plan_eval(jb_plan *plan, const double *p_fit, int64_t n_fit, double *loss, double *grad);

/*
 * Same, device resident: d_p_fit (n_fit doubles, may be NULL to keep the current parameters) and
 * d_out (1 + n_fit doubles: loss, gradient) are DEVICE pointers on the plan's device; work is
 * enqueued on `stream` (a cudaStream_t; NULL = the plan's own stream) and NOT synchronised.
 * Status problems found by the kernels are reported by the next jb_plan_check().
 */
int jb_plan_eval_device(jb_plan *plan, const double *d_p_fit, double *d_out, void *stream);
int jb_plan_check(jb_plan *plan);

/* model(p, xs[r], meta[r]) for every row -> (E,P) host array (Model.__call__, model.py:138-160).
 * Pixels outside the knot grid give NaN. */
int jb_plan_forward(jb_plan *plan, const double *p_fit, int64_t n_fit, double *model_out);

/*
 * Diagonal Fisher information of the ShiftingModel of component `component`
 * (EpochSpecificModel.f_info, model.py:857-901) at the current parameters:
 * F[k] = sum_{rows with key k} sum_pix where(~mask, (d model / d shift_k)^2 * yivar, 0).
 */
int jb_plan_fisher(jb_plan *plan, int32_t component, double *f_out, int64_t n_shift);

/*
 * EpochSpecificModel.grid_search (jabble/model.py:738-786) for the ShiftingModel of additive
 * component `component`: loss_out[k * n_grid + j] = the loss (coefficient * chi-square) of the rows
 * whose shift key is k, evaluated with that shift replaced by grid[k * n_grid + j] and every other
 * parameter as stored (jb_plan_set_params).  The reference evaluates the n_grid columns under
 * jax.vmap and sums rows into keys on the host; here one kernel forms every (row, candidate) sum
 * and the rows of a key are added in row order.  Called by parabola_fit (model.py:788-831).
 */
int jb_plan_grid_search(jb_plan *plan, int32_t component, const double *grid, int64_t n_shift,
                        int64_t n_grid, double *loss_out);

/*
 * Diagonal of the Hessian of the loss with respect to the shifts of `component`
 * (quickplay.get_RV_error_2d, jabble/quickplay.py:455-482, which takes jacrev(grad(loss))[key, key]
 * row by row): h_out[k] = coefficient * sum over the rows of key k and their unmasked pixels of
 * ivar ((a S')^2 - (y - m) a S'').
 */
int jb_plan_curvature(jb_plan *plan, int32_t component, double *h_out, int64_t n_shift);

/* Gradient w.r.t. ALL parameters from the last evaluation (diagnostics / tests).  Entries of
 * parameters that are not in fit mode hold only what the kernel variant of that evaluation formed. */
int jb_plan_grad_all(jb_plan *plan, double *grad_all, int64_t n_params);

/*
 * Batch of independent plans (chunks / echelle orders: separate fits that the reference runs in a
 * multiprocessing.Pool, notebooks/02-51peg.py:66-78) evaluated together: every kernel of the
 * evaluation is launched ONCE for the whole batch, so small per-chunk grids do not leave SMs idle.
 * The plans must live on one device and share the model shape.  jb_batch_bind fixes, per plan, the
 * device pointers of the fit vector (n_fit doubles; the array may be NULL to keep the parameters)
 * and of the output (1 + n_fit doubles: loss, gradient); jb_batch_eval_device enqueues one
 * evaluation of all plans on `stream` (NULL = the batch's own stream) without synchronising.
 * Once two calls in a row enqueue the same launches (same plans, fit state, row order and stream) the
 * sequence is recorded as a CUDA graph and replayed from then on -- one launch per evaluation; the same
 * holds for jb_group_eval_device (its ncclAllReduce included) and for the rounds of jb_lbfgs.  Anything
 * that changes a plan's launch parameters drops the recording; the environment variable JB_GRAPHS=0
 * turns recording off.  A stream that is being captured by the caller is left alone (the launches
 * join the caller's capture).
 */
typedef struct jb_batch jb_batch;
int jb_batch_create(jb_plan *const *plans, int32_t n_plans, jb_batch **out);
void jb_batch_destroy(jb_batch *batch);
int jb_batch_bind(jb_batch *batch, const double *const *d_p_fit, double *const *d_out);
int jb_batch_eval_device(jb_batch *batch, void *stream);
/* The same, then waits and handles knot-window overflows as jb_plan_eval does (re-sort, wider window,
 * shorter row groups, evaluate again); out-of-range pixels are reported.  For device-side drivers
 * before they start iterating. */
int jb_batch_eval_device_checked(jb_batch *batch, void *stream);
/*
 * The same evaluation with HOST buffers (what scipy's func / fprime hand over for every chunk of a
 * Pool of fits): p_fit_host[i] is the fit vector of plan i, out_host[i] receives [loss, gradient]
 * (1 + n_fit doubles).  One host->device copy of all fit vectors from pinned staging, one launch per
 * kernel for all plans, one device->host copy, one synchronisation; errors as jb_plan_eval.
 * Replaces any binding made by jb_batch_bind.
 * A knot-window overflow is handled as in jb_plan_eval (re-sort, wider window, shorter row groups,
 * the batch evaluated again).
 * _begin / _end split the call so that a caller can keep several batches (each on its own stream)
 * in flight: begin copies and enqueues and returns; end waits and hands the results over.
 */
int jb_batch_eval_host(jb_batch *batch, const double *const *p_fit_host, double *const *out_host);
int jb_batch_eval_host_begin(jb_batch *batch, const double *const *p_fit_host);
int jb_batch_eval_host_end(jb_batch *batch, double *const *out_host);
/*
 * The same without the staging copies: jb_batch_host_buffers hands out the batch's own page-locked
 * host buffers -- every plan's fit vector back to back in *p_fit_pinned (plan i at off_p[i] ..
 * off_p[i+1]) and every plan's [loss, gradient] in *out_pinned (plan i at off_out[i] .. off_out[i+1]);
 * off_p and off_out have n_plans + 1 entries.  The caller (one optimiser thread per chunk, the Pool of
 * notebooks/02-51peg.py:66-78) writes its fit vector in place; jb_batch_eval_pinned copies
 * host->device, evaluates every plan, copies device->host and synchronises; the results are read in
 * place.  The buffers stay valid until a plan's fit vector changes size (jb_plan_set_fit), then
 * jb_batch_host_buffers must be called again.  Replaces any binding made by jb_batch_bind.
 */
int jb_batch_host_buffers(jb_batch *batch, double **p_fit_pinned, double **out_pinned,
                          int64_t *off_p, int64_t *off_out);
int jb_batch_eval_pinned(jb_batch *batch);
/* _begin copies host->device and enqueues the evaluation and the copy back on the batch's own stream
 * and returns; _end waits, settles and checks.  Two batches (two halves of a GPU's chunks) driven
 * begin, begin, end, end overlap the copies of one with the kernels of the other. */
int jb_batch_eval_pinned_begin(jb_batch *batch);
int jb_batch_eval_pinned_end(jb_batch *batch);
/* enable != 0: a plan whose trial point puts an unmasked pixel outside a knot grid does not fail the
 * batched host call; its loss comes back NaN (gradient undefined) and the caller decides -- an
 * optimiser's line search treats it as a rejected step (the reference's clamped gather, model.py:985,
 * gives a large finite loss there). */
int jb_batch_soft_range(jb_batch *batch, int enable);
int jb_batch_check(jb_batch *batch);
int64_t jb_batch_launches(const jb_batch *batch);

/*
 * One chunk whose epochs are sharded over the GPUs of a box, ONE PROCESS PER GPU (BASELINE.json
 * configs[4]).  The reference's analogue is the epoch micro-batch loop of LossFunc.loss_all
 * (jabble/loss.py:64-74): per-batch losses add up; here the batches live on different GPUs.
 * Every rank creates the plans of its own rows (same model, the templates replicated; rows are
 * split on which_key boundaries so that every epoch key lives on one rank) and one jb_group over
 * them.  jb_group_eval_device enqueues, on one stream and without host work in between: the rank's
 * plans (one launch per kernel, as jb_batch_eval_device), the sum over the rank's plans of the loss
 * and of the gradient of every fit-mode template control point, one ncclAllReduce of that vector
 * (8 (1 + n_shared) bytes) over NVLink, and the totals written back into every plan's output.
 * Afterwards out[0] of every plan is the loss of the WHOLE chunk and the control-point entries of
 * its gradient are those of the whole chunk; per-epoch entries (shifts, stretches, their Fisher
 * sums) are the rank's own.  world = 1 needs no NCCL and no id.
 * NCCL is loaded at the first use (dlopen: the one already in the process when there is one, else
 * libnccl.so.2, else $JB_NCCL_LIB); rank 0 calls jb_nccl_unique_id and hands the 128 bytes to the
 * other ranks by any means (the host side here uses torch.distributed).
 */
typedef struct jb_group jb_group;
int jb_nccl_unique_id(void *id128);
int jb_nccl_version(void);          /* 0: NCCL not available */
int jb_group_create(jb_plan *const *plans, int32_t n_plans, int32_t rank, int32_t world, const void *nccl_id128,
                    jb_group **out);
void jb_group_destroy(jb_group *group);       /* destroys the communicator, not the plans */
int jb_group_bind(jb_group *group, const double *const *d_p_fit, double *const *d_out);   /* as jb_batch_bind */
int jb_group_eval_device(jb_group *group, void *stream);
/* The same, but the rank's plans are settled first as in jb_batch_eval_device_checked (re-sort, wider
 * windows; a rank off the knot grid contributes a NaN loss, which every rank then sees); waits. */
int jb_group_eval_device_checked(jb_group *group, void *stream);
int jb_group_info(const jb_group *group, int64_t *n_shared, int64_t *collective_bytes, int32_t *world);
int64_t jb_group_launches(const jb_group *group);

/*
 * L-BFGS-B for many independent chunks at once, every vector on the device: the optimiser of
 * Model.optimize (jabble/model.py:226-232: scipy.optimize.fmin_l_bfgs_b(func, x0, fprime, **options),
 * no bounds) and of the per-order process pool of notebooks/02-51peg.py:66-78.  One round = one batched
 * evaluation of every chunk still iterating (as jb_batch_eval_device_checked) + one launch that takes
 * each chunk's new [loss, gradient] through its own state machine (More'-Thuente search, BFGS update,
 * two-loop recursion, next trial point); the host reads back one int per chunk per round.
 * Options and results have scipy's meaning: m, factr, pgtol, maxiter, maxfun, maxls; warnflag 0 =
 * converged, 1 = maxiter / maxfun reached, 2 = other; (task, task_msg) index scipy's status_messages /
 * task_messages tables (4/401 "CONVERGENCE: NORM OF PROJECTED GRADIENT <= PGTOL", 4/402 "... RELATIVE
 * REDUCTION OF F <= FACTR*EPSMCH", 5/504 "STOP: TOTAL NO. OF ITERATIONS REACHED LIMIT", 5/502 "... F,G
 * EVALUATIONS EXCEEDS LIMIT", 8/0 "ABNORMAL: "); task 7 / task_msg 799: the STARTING point puts an
 * unmasked pixel outside a knot grid (a later trial point that does so is a large loss without slope,
 * counted in n_out_of_range).
 * Regularisers of the loss tree (jabble/loss.py:187-222) are parameter-only terms added on the device:
 * kind 0: weight * 0.5 * sum (x[i] - constant)^2, kind 1: weight * sum |x[i] - constant| over `indices`
 * into the fit vector (weight = coefficient * rows of the block: the reference evaluates a regulariser
 * once per row, loss.py:61-74).
 */
typedef struct jb_lbfgs jb_lbfgs;
typedef struct {
    int32_t m;          /* pairs kept (scipy: m = 10) */
    int32_t maxls;      /* trial points per line search (20) */
    int64_t maxiter;    /* 15000 */
    int64_t maxfun;     /* 15000 */
    double factr;       /* 1e7 */
    double pgtol;       /* 1e-5 */
} jb_lbfgs_options;
typedef struct {
    double f;           /* loss at x */
    double pg_norm;     /* max |g_i| at the last accepted iterate */
    int64_t nit, funcalls;
    int32_t warnflag, task, task_msg;
    int32_t n_out_of_range, n_skipped, n_restarts;   /* trial points off a knot grid; BFGS pairs skipped; memory resets */
} jb_lbfgs_result;
int jb_lbfgs_create(jb_plan *const *plans, int32_t n_plans, const jb_lbfgs_options *options, jb_lbfgs **out);
void jb_lbfgs_destroy(jb_lbfgs *opt);         /* not the plans */
int jb_lbfgs_add_reg(jb_lbfgs *opt, int32_t chunk, int32_t kind, double weight, double constant,
                     const int64_t *indices, int64_t n_indices);
/* x0[i], x[i] (and grad[i], when grad and grad[i] are not NULL): host vectors of plan i's n_fit doubles.
 * The plans keep the fit state and the fixed parameters they have (jb_plan_set_params / _set_fit). */
int jb_lbfgs_run(jb_lbfgs *opt, const double *const *x0, double *const *x, double *const *grad,
                 jb_lbfgs_result *results);
int64_t jb_lbfgs_rounds(const jb_lbfgs *opt); /* batched evaluations so far */

/* The device's measured FP64 FMA throughput (TFLOP/s: a register-resident DFMA kernel, best of 3):
 * in float64 the fused kernels sit at the roofline ridge, so benchmarks quote the FP64 pipe next to HBM. */
int jb_double_peak(int32_t device, double *tflops);

/* The plan's own cudaStream_t (so that a caller can order its work, or time it with events). */
int jb_plan_stream(jb_plan *plan, void **stream_out);

/*
 * Per-kernel timing for roofline reporting: while enabled, every evaluation records a CUDA event
 * pair around the fused loss+gradient kernel on the launching stream.  jb_plan_profile_read waits
 * for the stream, returns the summed duration (ms) and number of timed launches since the last
 * read, and resets both.
 */
int jb_plan_profile(jb_plan *plan, int enable);
int jb_plan_profile_read(jb_plan *plan, double *fused_ms_total, int64_t *n_launches);

#ifdef __cplusplus
}
#endif
#endif /* JABBLE_B200_H */
