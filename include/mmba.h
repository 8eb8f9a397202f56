/*
 * mmba.h — C-ABI of libmmba.so, the B200 (sm_100a) bundle-adjustment engine that replaces the
 * `scipy.optimize.least_squares(...)` call inside MeatModeler's `bundleAdjuster.adjustPoints`
 * (reference: bundleAdjuster.py:179-192; caller processor.py:465-470).
 *
 * Conventions
 *   - every function returns 0 on success and a negative MMBA_ERR_* code on failure; the message
 *     is retrievable with mmba_last_error().  No exception crosses this boundary.
 *   - the caller owns every host buffer; the library owns all device memory.
 *   - a handle is not thread-safe; distinct handles are independent.  All device work of a handle
 *     runs on one CUDA stream; every entry point returns only after its work has completed.
 *   - plain pointers and sizes only: no torch / numpy types.
 *   - there is NO CPU fallback: entry points that compute fail with MMBA_ERR_CUDA when no sm_100
 *     device is usable.  Functions in the "host-only" section never touch the GPU.
 *
 * Parameter vector layout (bundleAdjuster.py:175-176, 96-97):
 *   x = [ w0 t0 | w1 t1 | ... (6 doubles per camera) | X0 | X1 | ... (3 doubles per point) ]
 * Residual layout (bundleAdjuster.py:102): f = [du0, dv0, du1, dv1, ...] in the caller's
 * observation order.
 */
#ifndef MMBA_H
#define MMBA_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MMBA_VERSION 100 /* 0.1.0 */

enum {
    MMBA_OK = 0,
    MMBA_ERR_ARG = -1,       /* bad argument (null pointer, index out of range, size <= 0) */
    MMBA_ERR_CUDA = -2,      /* CUDA runtime failure or no usable sm_100 device */
    MMBA_ERR_STATE = -3,     /* call order (solve before set_problem, ...) */
    MMBA_ERR_NONFINITE = -4, /* residuals not finite at the initial point
                                (scipy raises ValueError: least_squares.py:945-946) */
    MMBA_ERR_TRACK = -5,     /* a point has more observations than one tile can hold */
    MMBA_ERR_NCCL = -6,      /* NCCL failure (multi-GPU only) */
    MMBA_ERR_NOMEM = -7
};

typedef struct mmba_handle mmba_handle;
typedef struct mmba_plan mmba_plan;

typedef struct mmba_options {
    int32_t device;       /* CUDA device ordinal */
    int32_t rank;         /* shard rank in [0, nranks) */
    int32_t nranks;       /* 1 = single GPU; >1 = observations sharded by point, NCCL allreduce */
    int32_t verbose;      /* 0 silent, 2 = print scipy's verbose=2 iteration table */
    double ftol;          /* reference value 1e-4 (bundleAdjuster.py:185) */
    double xtol;          /* scipy default 1e-8 */
    double gtol;          /* scipy default 1e-8 */
    int64_t max_nfev;     /* 0 = scipy default 100*n (trf.py:452-453) */
    double pcg_rtol;      /* relative residual stop of the reduced-system PCG */
    int32_t pcg_maxit;
    int32_t profile;      /* bit 0: bracket kernels with CUDA events (mmba_get_profile); bit 1: keep the residual history
                             of every reduced-system PCG solve (mmba_get_pcg_history; read by mmba_set_problem); bit 2:
                             per-phase cycle counters inside the PCG and S-build kernels (mmba_get_phase_cycles) */
    uint8_t nccl_id[128]; /* ncclUniqueId bytes, same on all ranks (nranks > 1 only) */
    int32_t schur_mode;   /* MMBA_SCHUR_*: how the PCG applies the reduced camera system (read by mmba_set_problem) */
    int32_t reserved;
    double pcg_atol;      /* counterpart of LSMR's stopping test 2 (lsmr.py:430-459: ||A^T res|| <= atol ||A|| ||res||,
                             which ends scipy's inner solves) on the reduced system, where A^T res is the PCG residual:
                             stop when ||r|| <= pcg_atol ||f||.  0 = relative rule only.  Default 1e-7 (calibrated on
                             the reference's LSMR iteration counts, see oracle/schur_trf.py). */
    double pcg_ktol;      /* LSMR's test 2 with its growing norm estimate (lsmr.py:430-459: ||A^T res|| <= atol ||A||_est
                             ||res||, ||A||_est ~ sqrt(k) for unit-norm columns) on the reduced system.  LSMR is a
                             minimal-residual method on the normal equations; the minimal-residual norm of the PCG
                             process is nu_k, 1 / nu_k^2 = sum_{j<=k} 1 / ||r_j||^2: stop when
                             nu_k <= pcg_ktol sqrt(k) ||f||.  0 = off.  Default 1.23e-6 = LSMR's atol (1e-6, scipy's default,
                             the one the reference gets) x the measured growth of its estimate, ||A||_est = 1.23 sqrt(k)
                             (tools/lsmr_spy.py): the literal translation, not a calibration. */
} mmba_options;

/* The reduced camera system S = U - W V'^-1 W^T of the damped Gauss-Newton step:
 *   IMPLICIT  S p is evaluated by one streaming pass over J per PCG iteration (152 B/observation);
 *   EXPLICIT  the block-sparse S is formed once per outer iteration (one streaming pass) and the whole PCG
 *             runs in one cooperative kernel on the L2-resident matrix;
 *   AUTO      EXPLICIT when S is small next to the observation stream (video-like visibility), else IMPLICIT. */
enum { MMBA_SCHUR_AUTO = 0, MMBA_SCHUR_IMPLICIT = 1, MMBA_SCHUR_EXPLICIT = 2 };

typedef struct mmba_result {
    double cost;          /* 0.5 * f.f at the returned x */
    double initial_cost;
    double optimality;    /* ||J^T f||_inf, unscaled (trf.py:466) */
    int64_t nfev, njev, nit;
    int32_t status;       /* scipy termination codes: 0 max_nfev, 1 gtol, 2 ftol, 3 xtol, 4 both */
    int32_t reserved;
    int64_t pcg_iterations; /* total reduced-system PCG iterations */
    double solve_ms;      /* device time of the solve (CUDA events on the handle's stream) */
} mmba_result;

/* one row per outer iteration, the columns of scipy's verbose=2 table (common.py:545-563) */
typedef struct mmba_iter_log {
    int64_t iteration, nfev;
    double cost, cost_reduction, step_norm, optimality;
    double reg, delta;
    int64_t pcg_iterations;
} mmba_iter_log;

/* kernel classes of mmba_get_profile */
enum {
    MMBA_K_CAMPREP = 0,
    MMBA_K_BUILD = 1,     /* residual + Jacobian blocks + normal-equation blocks */
    MMBA_K_RESID = 2,     /* trial residual / cost */
    MMBA_K_PTINV = 3,     /* damped 3x3 point-block inversion */
    MMBA_K_RHS = 4,       /* Schur right-hand side + Schur diagonal blocks */
    MMBA_K_MATVEC = 5,    /* implicit Schur-complement product (PCG) */
    MMBA_K_BACKSUB = 6,
    MMBA_K_JV = 7,        /* J*v products (Cauchy step, 2-D subspace Gram) */
    MMBA_K_VEC = 8,       /* small vector kernels */
    MMBA_K_ALLREDUCE = 9,
    MMBA_K_SBUILD = 10,   /* explicit reduced camera matrix + Schur right-hand side (one pass over J) */
    MMBA_K_PCG = 11,      /* the whole PCG solve on the explicit matrix (one cooperative launch) */
    MMBA_K_COUNT = 12
};

int mmba_version(void);
const char* mmba_last_error(const mmba_handle* h); /* h may be NULL: last error of this thread */
void mmba_default_options(mmba_options* opt);

/* replaces: the optimiser object scipy builds inside least_squares (least_squares.py:925-934) */
int mmba_create(mmba_handle** out, const mmba_options* opt);
/* nranks > 1: rank 0 obtains the ncclUniqueId here and ships it to the other ranks (the Python
 * layer broadcasts it with torch.distributed); every rank passes it in mmba_options.nccl_id */
int mmba_nccl_unique_id(uint8_t out[128]);
void mmba_destroy(mmba_handle* h);
/* change the tolerances / limits / verbosity / profiling of an existing handle (device, rank,
 * nranks and nccl_id of `opt` are ignored: a handle keeps its device and communicator), so that
 * one handle — one stream, one NCCL communicator — serves successive problems */
int mmba_set_options(mmba_handle* h, const mmba_options* opt);

/* replaces: pointAdjustmentSparsity (bundleAdjuster.py:55-78, 179) — the block structure is implied
 * by the two index arrays.  Indices are narrowed to int32 and range-checked while they are staged to
 * the device; the tile plan (point order, point-aligned tiles, per-tile camera tables) and the block
 * pattern of the reduced camera matrix are built by device kernels (csrc/devplan.cu).
 * nranks > 1: all ranks pass the full, identical problem; rank r READS only observations
 * [n_obs r / nranks, n_obs (r + 1) / nranks) of the arrays, and one all-to-all over NVLink moves every
 * observation to the rank that owns its point. */
int mmba_set_problem(mmba_handle* h, int64_t n_cams, int64_t n_points, int64_t n_obs,
                     const double K[9], const int64_t* cam_idx, const int64_t* pt_idx,
                     const double* uv /* n_obs x 2 row-major */);

/* replaces: least_squares(pointFun, x0, jac_sparsity=A, x_scale='jac', ftol=1e-4, method='trf')
 * (bundleAdjuster.py:180-192).  x is in/out (n = 6*n_cams + 3*n_points); fun_out (2*n_obs, may be
 * NULL) receives the residuals at the returned x in the caller's observation order. */
int mmba_solve(mmba_handle* h, double* x, mmba_result* result, double* fun_out);

/* The same solve with the two halves of the parameter vector in separate, caller-owned arrays — what adjustPoints holds
 * before it packs them (bundleAdjuster.py:172-176: frameParameters(...) (6 n_cams) and points_3D (3 n_points)) and what it
 * unpacks them into afterwards (:137-157): saves the packed host copies in both directions.  Inputs are not modified. */
int mmba_solve_split(mmba_handle* h, const double* cams_in, const double* points_in, double* cams_out, double* points_out,
                     mmba_result* result, double* fun_out);

/* replaces: least_squares(poseFun, parameters, ftol=1e-4) inside adjustPose (bundleAdjuster.py:232-241):
 * dense-Jacobian defaults, i.e. method='trf', tr_solver='exact', x_scale=1.  Only the 6*n_cams camera
 * parameters of x are variables; the point coordinates in x are constants (the chessboard).  x has
 * the layout of mmba_solve; single-GPU handles only. */
int mmba_solve_pose(mmba_handle* h, double* x, mmba_result* result, double* fun_out);

/* The same solve with the parameters resident in HBM: mmba_set_x uploads a starting point once,
 * mmba_solve_resident restarts from it without host<->device parameter traffic (what bench.py
 * times as the device-resident figure), mmba_get_x downloads the current parameters. */
int mmba_set_x(mmba_handle* h, const double* x);
int mmba_solve_resident(mmba_handle* h, mmba_result* result);
int mmba_get_x(mmba_handle* h, double* x);

int mmba_get_log(const mmba_handle* h, mmba_iter_log* out, int capacity); /* returns row count */
int mmba_get_profile(const mmba_handle* h, int64_t launches[MMBA_K_COUNT], double ms[MMBA_K_COUNT]);
/* diagnostics (options.profile bit 2): cycles one CTA spent in each phase of the instrumented kernels since the last
 * reset; [0..7] PCG iteration phases, [16..31] S-build tile phases */
int mmba_get_phase_cycles(mmba_handle* h, int64_t out[64], int reset);
/* residual history of the reduced-system PCG solves of the last mmba_solve* call (explicit Schur path, handles whose
 * options had profile bit 1 set at mmba_set_problem): outer_iteration < 0 returns the number of inner solves recorded;
 * otherwise the pairs (||r_k||^2, r_k . Pinv r_k), k = 0 .. iterations, are copied to out (capacity doubles) and
 * their count (2 (iterations + 1)) is returned */
int mmba_get_pcg_history(const mmba_handle* h, int outer_iteration, double* out, int capacity);
/* observations / points held by this rank and number of tiles (after set_problem) */
int mmba_get_shard(const mmba_handle* h, int64_t* n_obs_local, int64_t* n_points_local,
                   int64_t* n_tiles);

/* The plan mmba_set_problem built ON THE DEVICE (csrc/devplan.cu), downloaded in the formats of the host builder's
 * exports so that tests can compare the two bit for bit: sizes as mmba_plan_sizes; meta / tile_cams as mmba_plan_raw;
 * obs_perm (n_slots) / point_perm (n_points) as mmba_plan_export.  Any pointer may be NULL. */
int mmba_get_plan_sizes(mmba_handle* h, int64_t sizes[8]);
int mmba_get_plan_raw(mmba_handle* h, void* meta, int32_t* tile_cams, int64_t* obs_perm, int64_t* point_perm);
/* per-point statistics the device plan is built from (n_points entries each, any pointer may be NULL; summed over ranks for
 * a sharded handle): observation count, smallest / largest camera, smallest camera of the upper half of the ids */
int mmba_get_plan_stats(mmba_handle* h, int32_t* count, int32_t* first_cam, int32_t* last_cam, int32_t* first_hi);
/* the device-built block pattern of the reduced camera matrix: sizes[0..6] = upper blocks, full blocks, sum L (L + 1) / 2,
 * PCG CTAs, cameras per CTA, max blocks per CTA, max halo columns per CTA; up_rowptr / up_cols as mmba_host_rcm_pattern */
int mmba_get_rcm_pattern(mmba_handle* h, int64_t sizes[8], int32_t* up_rowptr, int32_t* up_cols, int64_t capacity);

/* ---- evaluation hooks (parity tests; same device code as the solve) ------------------------ */
/* replaces: pointFun(x, ...) (bundleAdjuster.py:81-102) */
int mmba_eval_residual(mmba_handle* h, const double* x, double* f /* 2*n_obs */);
/* replaces: the Jacobian scipy assembles by finite differences (_numdiff.py:770-893); blocks in the
 * caller's observation order: Jc n_obs x 2 x 6 (columns w0 w1 w2 t0 t1 t2), Jp n_obs x 2 x 3 */
int mmba_eval_jacobian(mmba_handle* h, const double* x, double* Jc, double* Jp);
/* J^T J and J^T f block-wise (common.py:590-610): U n_cams x 6 x 6, V n_points x 3 x 3 (full
 * symmetric), gc n_cams x 6, gp n_points x 3, cost = 0.5 f.f */
int mmba_eval_blocks(mmba_handle* h, const double* x, double* U, double* V, double* gc, double* gp,
                     double* cost);
/* replaces: lsmr(J_h, f, damp=sqrt(reg)) (trf.py:494-495) with J_h = J diag(scale):
 * solves (J_h^T J_h + reg I) p = J_h^T f by Schur elimination + block-Jacobi PCG. p has n entries. */
int mmba_eval_gn_step(mmba_handle* h, const double* x, const double* scale, double reg, double* p,
                      int64_t* pcg_iterations, double* pcg_relres);
/* test hook of the explicit reduced camera matrix (schur_mode != IMPLICIT and the matrix was formed):
 * S = D (J_c^T J_c - W V'^-1 W^T) D + reg I as a dense (6 n_cams)^2 row-major matrix and the right-hand side
 * b = D (g_c - W V'^-1 g_p), with D = diag(scale) of the camera parameters; MMBA_ERR_STATE otherwise */
int mmba_eval_reduced_system(mmba_handle* h, const double* x, const double* scale, double reg, double* S_dense,
                             double* rhs);
/* J * s for an n-vector s -> 2*n_obs (caller's observation order); test hook for the J*v kernel
 * (build_quadratic_1d / evaluate_quadratic, common.py:282-288, 348-361): returns ||J s||^2 */
int mmba_eval_jnorm2(mmba_handle* h, const double* x, const double* s, double* jnorm2);
/* time `iters` back-to-back launches of one kernel class on the current linearisation
 * (CUDA events on the handle's stream); used by bench.py for the roofline of each kernel */
int mmba_bench_kernel(mmba_handle* h, const double* x, int kernel_class, int iters, double* avg_ms);

/* ---- batched two-view triangulation ----------------------------------------------
 * replaces: the per-track cv2.triangulatePoints(projection1, projection2, feature, correspondent) loop of
 * processor.triangulatePoints (processor.py:246-261).  projections: n_frames x 3 x 4 row-major; f1 / f2:
 * frame of the first / last observation of each of the n tracks; uv1 / uv2: n x 2; points: n x 3
 * (dehomogenised, as processor.py:259 does); kernel_ms (may be NULL): device time of the kernel. */
int mmba_triangulate(int device, int64_t n_frames, const double* projections, int64_t n, const int64_t* f1,
                     const int64_t* f2, const double* uv1, const double* uv2, double* points, double* kernel_ms);

/* ---- rotate / project ---------------------------------------------------------------------------
 * replaces: rotate(points, rot_vecs) (bundleAdjuster.py:7-28: Rodrigues rotation of row i by rot_vecs[i]; theta = 0 leaves
 * the point unchanged) and project(points, frame_params, camera_matrix) (bundleAdjuster.py:31-52: rotate, translate by
 * frame_params[i, 3:6], multiply by K, divide by depth).  n rows; frame_params has param_stride >= 6 doubles per row. */
int mmba_rotate(int device, int64_t n, const double* points, const double* rot_vecs, double* out /* n x 3 */);
int mmba_project(int device, int64_t n, const double* points, const double* frame_params, int64_t param_stride,
                 const double K[9], double* out /* n x 2 */);

/* ---- host-only functions (no GPU needed; covered by the CPU test-suite) --------------------- */
/* replaces: solve_trust_region_2d (common.py:171-219); B = [b00, b01, b11] */
int mmba_host_tr2d(const double B[3], const double g[2], double delta, double p[2], int* newton);
/* replaces: minimize_quadratic_1d (common.py:298-322) for y = a t^2 + b t on [lb, ub] */
int mmba_host_min_quadratic_1d(double a, double b, double lb, double ub, double* t, double* y);
/* replaces: update_tr_radius (common.py:222-245) */
int mmba_host_update_tr_radius(double delta, double actual, double predicted, double step_norm,
                               int bound_hit, double* delta_new, double* ratio);
/* replaces: check_termination (common.py:705-717); returns 0 for "continue", else 2/3/4 */
int mmba_host_check_termination(double dF, double F, double dx_norm, double x_norm, double ratio,
                                double ftol, double xtol);

/* block pattern of the reduced camera matrix (which camera pairs share a point), upper triangle incl. the
 * diagonal, CSR: sizes[0] = blocks in the upper triangle, sizes[1] = blocks of the full pattern, sizes[2] = sum
 * over points of L (L + 1) / 2.  up_rowptr (n_cams + 1) / up_cols (capacity entries) may be NULL (sizes only);
 * returns MMBA_ERR_NOMEM when capacity is too small. */
int mmba_host_rcm_pattern(int64_t n_cams, int64_t n_points, int64_t n_obs, const int64_t* cam_idx,
                          const int64_t* pt_idx, int64_t sizes[3], int32_t* up_rowptr, int32_t* up_cols,
                          int64_t capacity);

/* tile plan (observation reordering + point sharding) built on the host */
int mmba_plan_create(mmba_plan** out, int64_t n_cams, int64_t n_points, int64_t n_obs,
                     const int64_t* cam_idx, const int64_t* pt_idx, int rank, int nranks);
void mmba_plan_destroy(mmba_plan* p);
/* sizes[0..7] = n_tiles, n_obs_local, n_points_local, point_begin, point_end (internal order),
 * tile_obs, max cameras per tile, padded observation slots */
int mmba_plan_sizes(const mmba_plan* p, int64_t sizes[8]);
/* obs_perm: for every padded slot the caller's observation index or -1; point_perm: internal
 * point -> caller's point index (all n_points); slot_cam_global / slot_point_local: the camera id
 * and the shard-relative internal point index each slot resolves to through the tile tables
 * (-1 for empty slots); the tile of a slot is slot / tile_obs */
int mmba_plan_export(const mmba_plan* p, int64_t* obs_perm, int64_t* point_perm,
                     int32_t* slot_cam_global, int32_t* slot_point_local);

/* the raw tile records (n_tiles x 2592 bytes: header + slot tables, csrc/plan.h TileMeta) and camera lists
 * (n_tiles x 256 global camera ids, -1 padded) of a host-built plan */
int mmba_plan_raw(const mmba_plan* p, void* meta, int32_t* tile_cams);

/* per tile (n_tiles entries each, any pointer may be NULL): distinct cameras, points, the S-build strategy the
 * plan chose (0 = point-pair-major, 1 = camera-pair-major flushed per tile, 2 = one 6x6 block per thread kept
 * across tiles) and the number of observation pairs sum L (L + 1) / 2 */
int mmba_plan_tile_stats(const mmba_plan* p, int32_t* n_cams, int32_t* n_points, int32_t* pair_mode, int32_t* n_pairs);

#ifdef __cplusplus
}
#endif
#endif /* MMBA_H */
