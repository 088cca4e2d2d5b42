/* b200mpc -- C-ABI of the B200-native batched MPC solver (libb200mpc.so).
 *
 * The reference (HybridRobotics/car-racing) is pure Python and has no FFI of its own; the
 * boundary it exposes for this path is the set of module-level solve functions that
 * car_racing/utils/base.py calls by name (base.py:199,256,307,476,558).  Each entry point
 * below is what a ctypes binding for one of those functions binds; car_racing_b200/control.py
 * holds that binding and the drop-in Python functions with the reference signatures.
 *
 *   b200mpc_cbf_solve     <- control.mpc_lti          (car_racing/control/control.py:198-248)  M = 0
 *                         <- control.mpccbf           (control.py:476-607)       M = rivals kept
 *                         <- control.mpc_multi_agents (control.py:251-473)       per-stage targets
 *   b200mpc_ilqr_solve    <- control.ilqr             (control.py:64-195, ilqr_helper.py:4-55)
 *   b200mpc_lmpc_solve    <- control.lmpc             (control.py:610-730)
 *   b200mpc_lmpc_sysid    <- LMPCRacingGame.estimate_ABC (utils/base.py:585-622, control/lmpc_helper.py:26-264)
 *   b200mpc_plant_step    <- DynamicBicycleModel.forward_dynamics (utils/base.py:897-942, system/vehicle_dynamics.py:4-49)
 *   b200mpc_argmin_cost   <- the argmin of OvertakeTrajPlanner.solve_optimization_problem
 *                            (car_racing/planning/overtake_traj_planner.py:244)
 *   b200mpc_comm_*        <- the join of that planner's fan-out (overtake_traj_planner.py:177-204: one forked process per
 *                            candidate, costs gathered through a Manager().dict()) for a batch sharded over several GPUs:
 *                            records exchanged through peer-mapped windows from the solver kernels' epilogue + argmin
 *   b200mpc_planner_*, b200mpc_plan_and_track*, b200mpc_rival_rollout, b200mpc_curv_to_glob: the callers either side of the
 *                            solves (SURVEY 8(f) "next" rows), see the comments at the declarations
 *
 * Conventions: plain pointers + sizes, IEEE double, C-contiguous, caller-owned buffers.
 * `*_solve` take HOST pointers (pageable or pinned) and perform H2D + kernel + D2H on the
 * handle's stream, synchronously.  `*_solve_device` take DEVICE pointers of the same layout
 * and only enqueue the kernel on the handle's stream (no sync) -- used when inputs are
 * already resident in HBM.  Every function returns 0 on success or a negative
 * b200mpc_status; b200mpc_last_error() gives the text.  Non-convergence of an instance is NOT
 * an error: it is reported per instance in rec[].status, as the reference's
 * `except RuntimeError` paths do (control.py:600-603).
 */
#ifndef B200MPC_H
#define B200MPC_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B200MPC_VERSION 100
#define B200MPC_NMAX 64   /* max horizon */
#define B200MPC_MMAX 8    /* max rivals per instance (one warp lane per column of the stage Hessian: 8 + 2M + 1 <= 32) */

typedef struct b200mpc_handle b200mpc_handle;

enum b200mpc_status {
    B200MPC_OK = 0,
    B200MPC_ERR_ARG = -1,
    B200MPC_ERR_CUDA = -2,
    B200MPC_ERR_NOMEM = -3,
    B200MPC_ERR_NODEVICE = -4
};

/* per-instance solver status (rec.status) */
enum b200mpc_solve_status {
    B200MPC_SOLVED = 0,        /* E_0 <= tol (or acceptable_tol for acceptable_iter iterations) */
    B200MPC_MAX_ITER = 1,      /* iterate returned (reference: opti.debug.value, control.py:602) */
    B200MPC_LINESEARCH = 2,    /* filter line search failed after the slack resets */
    B200MPC_INERTIA = 3,       /* inertia correction exhausted */
    B200MPC_INFEASIBLE_X0 = 4  /* x_0 violates its own stage-0 bound rows (control.py:582-586, :228-232 impose v_min <= vx_0 <= v_max,
                                  |ey_0| <= width on the fixed x_0): the reference's NLP is infeasible and IPOPT fails; the
                                  iterate returned is the solution without those rows */
};

/* Problem data shared by every instance of a batch: the *Param objects of the reference
 * (MPCCBFRacingParam base.py:272-291, SystemParam :708-713, CarParam :699-705) plus the
 * constants hard-coded inside control.mpccbf / mpc_multi_agents. */
typedef struct {
    int32_t N;              /* num_horizon */
    int32_t M;              /* rivals per instance in this batch, 0..B200MPC_MMAX (0 = mpc_lti) */
    int32_t xt_per_stage;   /* 0: xtarget is (6,) per instance; 1: (N+1,6) per instance (control.py:373-382) */
    int32_t flags;          /* B200MPC_FLAG_* : optional per-instance blocks appended to the record */
    double A[36], B[12];    /* matrix_A, matrix_B, row-major (control.py:566-570) */
    double Q[36], R[4];     /* matrix_Q, matrix_R (control.py:578-591) */
    double umax[2];         /* delta_max, a_max (control.py:572-576) */
    double vmin, vmax;      /* control.py:582-583 */
    double width;           /* track.width (control.py:585-586) */
    double alpha;           /* mpc_cbf_param.alpha (control.py:558) / 0.6 (control.py:285) */
    double margin;          /* safety_margin 0.2 (control.py:527) / 0.15 (control.py:311) */
    double L, W;            /* l_agent+l_obs, w_agent+w_obs (control.py:532-535) */
    double slack_w;         /* 10000 (control.py:560) */
} b200mpc_cbf_params;

/* flags: planner-candidate QP (planning/overtake_traj_planner.py:248-379) */
#define B200MPC_FLAG_STAGE_BOUNDS 1 /* record carries per-stage bounds on (vx_i, ey_i): (N+1) x {lb_vx, lb_ey, ub_vx, ub_ey};
                                       +-1e300 or +-inf = no bound (:276-324); vmin/vmax/width are then ignored */
#define B200MPC_FLAG_EY_RATE 2      /* record carries wd[0..N-1]: cost += wd[i]*(ey_{i+1}-ey_i)^2 (:325-327) */
#define B200MPC_FLAG_RIVAL_SIZE 4   /* record carries (L_j, W_j), j < M: l_agent+l_obs, w_agent+w_obs of every rival, which the
                                       reference reads per rival (control.py:530-535, :316-319); L, W of the params are then ignored */

/* start point of the interior-point iteration */
#define B200MPC_START_ROLLOUT 0     /* x_i = A^i x_0 (u = 0 roll-out), dynamically feasible; the default */
#define B200MPC_START_ZERO 1        /* x_1..x_N = 0: what CasADi's Opti hands IPOPT (no opti.set_initial in control.py:476-607) */

/* Interior-point options (IPOPT option names where they exist). */
typedef struct {
    double tol;             /* 1e-8 */
    int32_t max_iter;       /* 200 */
    int32_t acceptable_iter;/* 15 */
    double acceptable_tol;  /* 1e-6 */
    double mu_init;         /* 0.1 */
    double rho;             /* l1 weight of the elastic CBF rows (scaled objective units), 1e3 */
    double bound_push;      /* 1e-2 */
    double bound_frac;      /* 1e-2 */
    double max_grad;        /* nlp_scaling_max_gradient, 100 */
    int32_t start;          /* B200MPC_START_ROLLOUT (default) or B200MPC_START_ZERO */
    int32_t max_reset;      /* elastic slack resets per solve when the filter line search fails (our stand-in for IPOPT's
                               restoration phase), 5 */
} b200mpc_ipm_options;

/* 32-byte per-instance record: what the planner's argmin / the multi-GPU all-gather moves. */
typedef struct {
    double cost;            /* optimal objective (the reference's `cost` expression) */
    double u0[2];           /* u_pred[0,:] -- the value control.mpccbf returns (control.py:607) */
    int32_t status;         /* b200mpc_solve_status */
    int32_t iters;
} b200mpc_record;

int b200mpc_version(void);
void b200mpc_default_ipm_options(b200mpc_ipm_options *opt);

/* device < 0: current CUDA device.  max_batch (>= 1) is a capacity HINT for the staging buffers of the host-pointer API:
 * they are allocated on first use and grown on demand, so a larger batch is never an error. */
int b200mpc_create(int device, int max_batch, b200mpc_handle **out);
/* The same with the handle's stream created at the device's highest priority when high_priority != 0: for short kernels
 * that must not queue behind the pending blocks of other handles' batches (the planner's exchange / argmin step). */
int b200mpc_create_ex(int device, int max_batch, int high_priority, b200mpc_handle **out);
void b200mpc_destroy(b200mpc_handle *h);
const char *b200mpc_last_error(const b200mpc_handle *h); /* h may be NULL: last create() error */
/* the handle's CUDA stream (cudaStream_t) as an integer, for event timing by the caller */
uint64_t b200mpc_stream(const b200mpc_handle *h);
/* number of kernel launches issued through this handle so far */
uint64_t b200mpc_launch_count(const b200mpc_handle *h);

/* doubles per instance of the packed input record for (N, M, xt_per_stage) and flags == 0:
 *   [x0 6][lap_off M][pad to even][xtarget 6 or 6(N+1)][obs j=0..M-1: s_0..s_N, ey_0..ey_N][pad to even]
 * with flags the record continues with [bounds 4(N+1)] and/or [wd N][pad to even] and/or [L_0 W_0 .. L_{M-1} W_{M-1}]
 * (b200mpc_cbf_record_doubles_ex gives the total).
 * lap_off[j] = (num_cycle_ego - num_cycle_obs)*lap_length (control.py:538-540); obs rows are rows 4,5
 * of get_trajectory_nsteps' (6,N+1) prediction (control.py:509-511). */
int b200mpc_cbf_record_doubles(int N, int M, int xt_per_stage);
int b200mpc_cbf_record_doubles_ex(int N, int M, int xt_per_stage, int flags);

/* Batched MPC-LTI / MPC-CBF / mpc_multi_agents solve.
 *   in    : B packed records (see above)
 *   rec   : B records (cost, u0, status, iters)
 *   aux   : optional B x 4 doubles {kkt_err, elastic_max, n_refactor, n_backtrack}
 *   xpred : optional B x (N+1) x 6   (x_pred of control.py:598)
 *   upred : optional B x N x 2       (u_pred of control.py:599)
 *   sigma : optional B x M x (N+1)   (cbf_slack of control.py:525)
 */
int b200mpc_cbf_solve(b200mpc_handle *h, const b200mpc_cbf_params *prm, const b200mpc_ipm_options *opt, int B,
                      const double *in, b200mpc_record *rec, double *aux, double *xpred, double *upred, double *sigma);
int b200mpc_cbf_solve_device(b200mpc_handle *h, const b200mpc_cbf_params *prm, const b200mpc_ipm_options *opt, int B,
                             const double *d_in, b200mpc_record *d_rec, double *d_aux, double *d_xpred, double *d_upred,
                             double *d_sigma);
/* The same call without the final stream synchronisation: H2D of the records, the kernel and the D2H of the
 * results are only ENQUEUED on the handle's stream.  The host buffers must stay alive (and should be pinned, or
 * the copies block) until b200mpc_synchronize(h) returns; a second call on the same handle before that reuses the
 * handle's staging buffers in stream order, which is safe.  Several handles (= streams) driven round-robin keep
 * several batches in flight: the reference's planner forks one process per candidate and joins them
 * (planning/overtake_traj_planner.py:182-203); here the join is this call, and the stragglers of one batch
 * overlap the next batch instead of idling the GPU (DESIGN.md "Batches in flight"). */
int b200mpc_cbf_solve_async(b200mpc_handle *h, const b200mpc_cbf_params *prm, const b200mpc_ipm_options *opt, int B,
                            const double *in, b200mpc_record *rec, double *aux, double *xpred, double *upred, double *sigma);
int b200mpc_synchronize(b200mpc_handle *h);
/* page-locked host memory for the asynchronous calls (cudaHostAlloc / cudaFreeHost); NULL on failure */
void *b200mpc_host_alloc(size_t bytes);
void b200mpc_host_free(void *p);

/* iLQR (control.py:64-195).  Shared data: */
typedef struct {
    int32_t N;              /* num_horizon (50) */
    int32_t max_iter;       /* 150 (base.py:176) */
    double A[36], B[12], Q[36], R[4];
    double L, W;            /* l_agent+l_obs, w_agent+w_obs (control.py:107-110) */
} b200mpc_ilqr_params;

/* doubles per instance: [x0 6][xtarget 6][lap_off 1][pad 1][s_0..s_N][ey_0..ey_N][pad to even] */
int b200mpc_ilqr_record_doubles(int N);
int b200mpc_ilqr_solve(b200mpc_handle *h, const b200mpc_ilqr_params *prm, int B, const double *in,
                       b200mpc_record *rec, double *xpred, double *upred);
int b200mpc_ilqr_solve_device(b200mpc_handle *h, const b200mpc_ilqr_params *prm, int B, const double *d_in,
                              b200mpc_record *d_rec, double *d_xpred, double *d_upred);
/* as b200mpc_cbf_solve_async: enqueue only, join with b200mpc_synchronize */
int b200mpc_ilqr_solve_async(b200mpc_handle *h, const b200mpc_ilqr_params *prm, int B, const double *in,
                             b200mpc_record *rec, double *xpred, double *upred);

/* LMPC (control.py:610-730): the per-step QP over the LTV model and the convex hull of the selected safe set.
 * Replaces the CasADi `opti.solve()` at control.py:703.  Shared data: */
#define B200MPC_LMPC_NMAX 16 /* num_horizon (12, base.py:351) */
#define B200MPC_LMPC_KMAX 64 /* selected safe-set points (num_ss_points = 44, base.py:358) */
typedef struct {
    int32_t N, K;
    double Q[36], R[4], dR[4];   /* matrix_Q, matrix_R, matrix_dR (base.py:354-357) */
    double xtrk[6];              /* x_track (control.py:649) */
    double umax[2], vmax, width; /* delta_max, a_max, v_max, lap_width (control.py:658-666) */
} b200mpc_lmpc_params;

/* doubles per instance: [x0 6][u_old 2][A_0..A_{N-1} 36 each, row-major][B_i 12 each][C_i 6 each]
 *                       [SS 6 x K row-major][Qfun K][pad to even]
 * outputs as b200mpc_cbf_solve, plus lambda: optional B x K (lin_comb_lambda of control.py:618). */
int b200mpc_lmpc_record_doubles(int N, int K);
int b200mpc_lmpc_solve(b200mpc_handle *h, const b200mpc_lmpc_params *prm, const b200mpc_ipm_options *opt, int B,
                       const double *in, b200mpc_record *rec, double *aux, double *xpred, double *upred, double *lambda);
int b200mpc_lmpc_solve_device(b200mpc_handle *h, const b200mpc_lmpc_params *prm, const b200mpc_ipm_options *opt, int B,
                              const double *d_in, b200mpc_record *d_rec, double *d_aux, double *d_xpred, double *d_upred,
                              double *d_lambda);
int b200mpc_lmpc_solve_async(b200mpc_handle *h, const b200mpc_lmpc_params *prm, const b200mpc_ipm_options *opt, int B,
                             const double *in, b200mpc_record *rec, double *aux, double *xpred, double *upred, double *lambda);

/* LMPC model identification: LMPCRacingGame.estimate_ABC (utils/base.py:585-622) = per horizon stage
 * lmpc_helper.regression_and_linearization (control/lmpc_helper.py:26-201).  Replaces the N sequential calls (each
 * 2 nearest-neighbour scans over whole stored laps + 3 cvxopt solves) by one launch, one warp per (instance, stage). */
#define B200MPC_SYSID_LMAX 4  /* laps used for the regression (2: iter-2, iter-1; base.py:600-601) */
#define B200MPC_SYSID_PMAX 64 /* max_num_point (40, base.py:602) */
typedef struct {
    int32_t N;                 /* horizon stages (lmpc_param.num_horizon) */
    int32_t num_laps;          /* laps in `laps`, 1..B200MPC_SYSID_LMAX */
    int32_t max_num_point;     /* 1..B200MPC_SYSID_PMAX */
    int32_t num_segments;      /* rows of `segments` (track.point_and_tangent) */
    int32_t lap_rows[B200MPC_SYSID_LMAX]; /* time_ss of each lap: rows 0..time_ss-2 are candidates, row t+1 is the target */
    int32_t lap_stride;        /* allocated rows per (lap, field) */
    int32_t reserved;
    double dt;                 /* timestep */
    double h;                  /* kernel bandwidth, 5 (lmpc_helper.py:45) */
    double lap_length;         /* point_and_tangent[-1,3] + point_and_tangent[-1,4] (lmpc_helper.py:143) */
} b200mpc_sysid_params;

/* lin      : B x N x 8   (lin_points[i, 0:6], lin_input[i, 0:2]) per instance and stage
 * laps     : num_laps x 5 x lap_stride, fields (vx, vy, wz, delta, a) of the stored laps (ss_xcurv[:, 0:3, lap], u_ss[:, :, lap])
 * segments : num_segments x 3  (s_start, length, curvature) = point_and_tangent[:, 3:6]
 * out      : per instance A_0..A_{N-1} (36 each, row-major), B_i (12 each), C_i (6 each) written at
 *            out + b*out_stride + out_offset -- with out_offset = 8 and out_stride = b200mpc_lmpc_record_doubles(N, K)
 *            this is the model block of the LMPC record, so the two kernels chain on the device
 * idx      : optional B x N x num_laps x max_num_point selected rows (-1 padded) = index_used_list (base.py:621)
 * status   : optional B x N, 0 ok, bit0 singular normal equations, bit1 s outside the track table, bit2 < 5 points */
int b200mpc_lmpc_sysid(b200mpc_handle *h, const b200mpc_sysid_params *prm, int B, const double *lin, const double *laps,
                       const double *segments, double *out, int out_stride, int out_offset, int32_t *idx, int32_t *status);
int b200mpc_lmpc_sysid_device(b200mpc_handle *h, const b200mpc_sysid_params *prm, int B, const double *d_lin,
                              const double *d_laps, const double *d_segments, double *d_out, int out_stride, int out_offset,
                              int32_t *d_idx, int32_t *d_status);

/* Plant step: DynamicBicycleModel.forward_dynamics (utils/base.py:897-942) = n_sub Euler sub-steps of
 * system/vehicle_dynamics.py:4-49 + clipped noise + the lap wrap of update_memory (base.py:804-809). */
typedef struct {
    int32_t n_sub;          /* sub-steps: the count of the loop `while (i+1)*0.001 <= timestep` (base.py:909) = 100 */
    int32_t num_segments;   /* rows of `segments` */
    int32_t wrap_lap;       /* 1: s -= lap_length when s > lap_length and laps[b]++ */
    int32_t reserved;
    double delta_t;         /* 0.001 (base.py:901) */
    double lap_length;
    double m, lf, lr, Iz, Df, Cf, Bf, Dr, Cr, Br;   /* BicycleDynamicsParam (base.py:659-696) */
} b200mpc_plant_params;

/* In place: xcurv (vehicle b at xcurv + b*xcurv_stride + xcurv_offset, 6 doubles: with stride = record doubles and
 * offset 0 this is the x0 slot of the MPC records), xglob B x 6; u: delta, a of vehicle b at u + b*u_stride (u_stride = 4
 * reads u0 out of b200mpc_record arrays starting at &rec[0].u0); draws: optional B x 3 standard-normal draws (NULL =
 * zero_noise_flag); segments as in b200mpc_lmpc_sysid; laps: optional B lap counters. */
int b200mpc_plant_step(b200mpc_handle *h, const b200mpc_plant_params *prm, int B, double *xcurv, int xcurv_stride,
                       int xcurv_offset, double *xglob, const double *u, int u_stride, const double *draws,
                       const double *segments, int32_t *laps);
int b200mpc_plant_step_device(b200mpc_handle *h, const b200mpc_plant_params *prm, int B, double *d_xcurv, int xcurv_stride,
                              int xcurv_offset, double *d_xglob, const double *d_u, int u_stride, const double *d_draws,
                              const double *d_segments, int32_t *d_laps);

/* Planner selection + hand-over to the tracking MPC: the part of
 * OvertakeTrajPlanner.solve_optimization_problem after the candidates are gathered
 * (planning/overtake_traj_planner.py:205-246) and the target construction of control.mpc_multi_agents
 * (control/control.py:277, 373-382), on the device, so that candidate solve -> selection -> tracking solve
 * chain on one stream. */
typedef struct {
    int32_t C;                  /* candidates (the reference: num_veh + 1 regions) */
    int32_t N;                  /* num_horizon_planner (10) */
    int32_t num_veh;            /* rivals, in sorted_vehicles order */
    int32_t old_direction_flag; /* previous choice, -1 = None (:238-243) */
    int32_t N_ctrl;             /* num_horizon_ctrl of the tracking MPC (stages of the target block to fill) */
    int32_t M_ctrl;             /* rivals in the tracking record (fixes the offset of its target block) */
    double veh_length, veh_width; /* ego.param.length / width (:209-210) */
    double lap_length;
} b200mpc_planner_select_params;

/* All pointers are device pointers.
 *   rec       : C records of the candidate solve;  xpred : C x (N+1) x 6 its trajectories
 *   heur      : C x (N+1) x 6 heuristic trajectories used where ok0[c] == 0 or the solve failed (:365-374)
 *   ok0       : C, 1 if x_0 satisfies the candidate's stage-0 rows;  region : C, region index of the candidate
 *               (neighbours: rivals region-1 and region; the reference has region[c] = c)
 *   rivals    : num_veh x 2 x (N+1): s and ey predictions (s is lap-wrapped here as :214-215 does)
 * outputs:
 *   sel_cost  : optional C selection costs;  flag : 2 ints {direction_flag = first argmin, its region}
 *   traj      : optional (N+1) x 6 chosen trajectory (traj_xcurv of :245)
 *   track_rec : optional packed record of the tracking MPC (layout of b200mpc_cbf_record_doubles(N_ctrl, M_ctrl, 1));
 *               x0 must be filled in; its per-stage target block is written in place */
int b200mpc_planner_select_device(b200mpc_handle *h, const b200mpc_planner_select_params *prm, const b200mpc_record *d_rec,
                                  const double *d_xpred, const double *d_heur, const int32_t *d_ok0, const int32_t *d_region,
                                  const double *d_rivals, double *d_sel_cost, int32_t *d_flag, double *d_traj,
                                  double *d_track_rec);

/* The whole overtaking step on one stream, host pointers in and out: H2D -> candidate solve (plan_prm: flags
 * STAGE_BOUNDS|EY_RATE, M = 0; C packed records) -> b200mpc_planner_select_device -> tracking solve (track_prm: per-stage
 * targets, alpha 0.6, margin 0.15 as control.py:285,311; track_in = ONE packed record with x0, lap offsets and the
 * rival block filled in, its target block is completed on the device) -> D2H.  Replaces the reference's
 * fork/join of one process per candidate + host selection + a second CasADi/IPOPT solve
 * (overtake_traj_planner.py:162-246 followed by control.py:251-473, called from utils/base.py:540-582).
 * Outputs: cand_rec C (optional), cand_xpred C x (N+1) x 6 (optional), sel_cost C (optional), flag 2 ints, traj (N+1) x 6
 * (optional), track_rec 1 record (u0 = the control to apply), track_xpred (N_ctrl+1) x 6, track_upred N_ctrl x 2 (optional). */
int b200mpc_plan_and_track(b200mpc_handle *h, const b200mpc_cbf_params *plan_prm, const b200mpc_cbf_params *track_prm,
                           const b200mpc_ipm_options *opt, const b200mpc_planner_select_params *sel, const double *cand_in,
                           const double *heur, const int32_t *ok0, const int32_t *region, const double *rivals,
                           const double *track_in, b200mpc_record *cand_rec, double *cand_xpred, double *sel_cost, int32_t *flag,
                           double *traj, b200mpc_record *track_rec, double *track_xpred, double *track_upred);

/* Prediction of rivals that have dynamics: offboard.DynamicBicycleModel.get_trajectory_nsteps
 * (racing/offboard.py:80-94) = n Euler steps of the zero-input Frenet kinematics (get_estimation, :51-77), curvature from
 * the track segments (utils/racing_env.py:225-246), s wrapped into the lap after every step. */
typedef struct {
    int32_t n;              /* predicted steps (N + 1) */
    int32_t num_segments;   /* rows of `segments` */
    double timestep;        /* the model's timestep (0.1) */
    double lap_length;
} b200mpc_rollout_params;

/* xcurv, xglob: B x 6 current states; segments: num_segments x 3 (start s, length, curvature) = point_and_tangent[:, 3:6];
 * xcurv_n, xglob_n: B x 6 x n, the reference's (6, n) blocks per rival (xglob_n optional: every caller in the reference
 * discards it). */
int b200mpc_rival_rollout(b200mpc_handle *h, const b200mpc_rollout_params *prm, int B, const double *xcurv, const double *xglob,
                          const double *segments, double *xcurv_n, double *xglob_n);
int b200mpc_rival_rollout_device(b200mpc_handle *h, const b200mpc_rollout_params *prm, int B, const double *d_xcurv,
                                 const double *d_xglob, const double *d_segments, double *d_xcurv_n, double *d_xglob_n);

/* Curvilinear -> global frame: racing_env.get_global_position / get_orientation
 * (utils/racing_env.py:6-127) for P points at once.  track: num_segments x 6 = ClosedTrack.point_and_tangent
 * (x, y, psi, start s, length, curvature).  s, ey: P values read with the given strides (doubles) -- stride 6 reads
 * columns 4, 5 of (.., 6) state trajectories in place.  out: P x 3 (x, y, psi). */
int b200mpc_curv_to_glob(b200mpc_handle *h, int P, int num_segments, double lap_length, const double *track, const double *s,
                         int s_stride, const double *ey, int ey_stride, double *out);
int b200mpc_curv_to_glob_device(b200mpc_handle *h, int P, int num_segments, double lap_length, const double *d_track,
                                const double *d_s, int s_stride, const double *d_ey, int ey_stride, double *d_out);

/* Candidate preparation on the device: what OvertakeTrajPlanner.get_local_traj computes between the
 * rivals' predictions and the candidate solves (planning/overtake_traj_planner.py:87-117: veh_infos, get_agent_info,
 * get_bezier_control_points, the sampled Bezier curves -- planning/planner_helper.py:46-153, 177-205) and the data part of
 * generate_traj_per_region (:276-334 targets and bounds, :365-374 heuristic trajectory), written as the packed candidate
 * records of the solver (M = 0, per-stage targets, flags STAGE_BOUNDS|EY_RATE). */
typedef struct {
    int32_t N;                  /* num_horizon_planner */
    int32_t num_veh;            /* rivals of interest >= 1; regions C = num_veh + 1 */
    int32_t num_opt;            /* rows of the optimal-trajectory table (>= 2, s ascending) */
    int32_t reserved;
    double prediction_factor;   /* racing_game_param.planning_prediction_factor (utils/base.py:391) */
    double track_width, lap_length;
    double veh_length, veh_width;            /* ego.param.length / width */
    double safety_margin;                    /* 0.15 (overtake_traj_planner.py:262) */
    double vx_max;                           /* 5 (:276) */
    double w_ey_rate, w_progress, w_track;   /* 30, 200, 20 (:327, 328, 333-334) */
} b200mpc_planner_prepare_params;

/* All pointers are device pointers.
 *   ego       : 12 doubles: vehicles["ego"].xcurv, then the xcurv_ego argument of get_local_traj
 *   rivals    : num_veh x 2 x (N+1): s and ey predictions in sorted_vehicles order (s not wrapped)
 *   rival_vx  : num_veh current vx (sorted order);  insertion : num_veh, insertion[i] = position in sorted_vehicles of the
 *               i-th rival of vehicles_interest (the reference fills veh_infos in that order and reads it by region)
 *   opt_traj  : num_opt x 2 (s, ey) of the optimal trajectory
 * outputs:
 *   cand      : C x b200mpc_cbf_record_doubles_ex(N, 0, 1, STAGE_BOUNDS|EY_RATE) packed records
 *   heur      : C x (N+1) x 6;  ok0, region : C ints (as b200mpc_planner_select_device takes them)
 *   offset    : optional C, reference cost = solver cost + offset
 *   ctrl      : optional C x 4 x 2 control points;  bezier : optional C x (N+1) x 2 curve samples
 *   err       : optional 1 int, nonzero if a look-up left the optimal trajectory's range (the reference raises ValueError) */
int b200mpc_planner_prepare_device(b200mpc_handle *h, const b200mpc_planner_prepare_params *prm, const double *d_ego,
                                   const double *d_rivals, const double *d_rival_vx, const int32_t *d_insertion,
                                   const double *d_opt_traj, double *d_cand, double *d_heur, int32_t *d_ok0, int32_t *d_region,
                                   double *d_offset, double *d_ctrl, double *d_bezier, int32_t *d_err);
/* The same with host pointers (H2D, kernel, D2H, synchronised). */
int b200mpc_planner_prepare(b200mpc_handle *h, const b200mpc_planner_prepare_params *prm, const double *ego,
                            const double *rivals, const double *rival_vx, const int32_t *insertion, const double *opt_traj,
                            double *cand, double *heur, int32_t *ok0, int32_t *region, double *offset, double *ctrl,
                            double *bezier, int32_t *err);

/* b200mpc_plan_and_track with the candidates prepared on the device: H2D of the raw planner inputs ->
 * b200mpc_planner_prepare_device -> candidate solve -> selection -> tracking solve -> D2H, one stream.  sel->C must be
 * prep->num_veh + 1 + n_extra; the n_extra additional candidates (BASELINE config 3 evaluates 64) come packed from the
 * host as in b200mpc_plan_and_track (extra_* may be NULL when n_extra = 0).  Outputs as b200mpc_plan_and_track plus
 * heur_out C x (N+1) x 6, ok0_out C, offset C (first num_veh + 1 entries filled), bezier (num_veh+1) x (N+1) x 2 (all optional)
 * and err (1 int, see above).  track_prm = NULL (with track_in = NULL) stops after the selection: the device form of
 * get_local_traj alone (planning/overtake_traj_planner.py:44-161); the tracking outputs are then not written. */
int b200mpc_plan_and_track_prepared(b200mpc_handle *h, const b200mpc_cbf_params *plan_prm, const b200mpc_cbf_params *track_prm,
                                    const b200mpc_ipm_options *opt, const b200mpc_planner_select_params *sel,
                                    const b200mpc_planner_prepare_params *prep, const double *ego, const double *rivals,
                                    const double *rival_vx, const int32_t *insertion, const double *opt_traj, int n_extra,
                                    const double *extra_cand, const double *extra_heur, const int32_t *extra_ok0,
                                    const int32_t *extra_region, const double *track_in, b200mpc_record *cand_rec,
                                    double *cand_xpred, double *sel_cost, int32_t *flag, double *traj, b200mpc_record *track_rec,
                                    double *track_xpred, double *track_upred, double *heur_out, int32_t *ok0_out, double *offset,
                                    double *bezier, int32_t *err);

/* argmin over records (device pointers): index of the smallest cost among status<=max_status,
 * lowest index wins ties (list.index(min(...)), overtake_traj_planner.py:244); *d_out = -1 if none. */
int b200mpc_argmin_cost_device(b200mpc_handle *h, const b200mpc_record *d_rec, int B, int max_status, int32_t *d_out);

/* ---- Exchange step of a batch sharded over G GPUs, one process per GPU.
 * Replaces the join of the reference's fan-out -- one forked process per candidate, results gathered through a
 * Manager().dict(), then `cost_selection.index(min(cost_selection))` (planning/overtake_traj_planner.py:177-204, :244) -- for a
 * candidate / scenario batch whose shards live on different GPUs.  It is not a collective call after the solve: each rank
 * owns a window in its HBM that the peers map over NVLink (CUDA IPC), and the solver kernels' epilogue stores every
 * 32-byte result record straight into all ranks' gathered buffers as the instance finishes (csrc/exchange.cuh).
 *
 *   b200mpc_comm_create   on every rank: allocates the window (slots x world x max_batch records + counters)
 *   b200mpc_comm_export   64-byte handle of the own window; the caller hands all ranks' handles to every rank (any
 *                         transport: torch.distributed, MPI, a file)
 *   b200mpc_comm_connect  maps the peers' windows
 *   b200mpc_comm_publish_next(h, c, slot)   the NEXT b200mpc_{cbf,ilqr,lmpc}_solve* call on handle h also publishes its B
 *                         records (every rank must solve the same B in that step) into slot `slot` of every rank
 *   b200mpc_comm_argmin   on handle h's stream: waits until all world x B records of the slot's oldest unconsumed use
 *                         have arrived, writes the first-min argmin over them (global instance order = rank-major) to d_out
 *                         and, if d_all != NULL, a copy of the gathered records; then releases the slot to the peers.
 * Uses of one slot must be issued in the same order on every rank; `slots` steps can be in flight.  The two waiting
 * kernels (the gate ahead of a solve that reuses a slot, and the argmin) spin on the device: give every stream its own
 * hardware queue (CUDA_DEVICE_MAX_CONNECTIONS >= number of streams in use, set before CUDA initialises; the default of 8
 * aliases streams, and a spinning kernel then blocks an unrelated stream behind it).  A wait gives up after 20 s; the argmin
 * then writes -2. */
typedef struct b200mpc_comm b200mpc_comm;
#define B200MPC_COMM_HANDLE_BYTES 64
#define B200MPC_COMM_MAX_WORLD 16
int b200mpc_comm_create(b200mpc_handle *h, int rank, int world, int max_batch, int slots, b200mpc_comm **out);
int b200mpc_comm_export(b200mpc_comm *c, void *handle_out);
int b200mpc_comm_connect(b200mpc_comm *c, const void *handles /* world x B200MPC_COMM_HANDLE_BYTES, rank-major */);
void b200mpc_comm_destroy(b200mpc_comm *c);
int b200mpc_comm_publish_next(b200mpc_handle *h, b200mpc_comm *c, int slot);
int b200mpc_comm_argmin(b200mpc_handle *h, b200mpc_comm *c, int slot, int max_status, int32_t *d_out, b200mpc_record *d_all);

#ifdef __cplusplus
}
#endif
#endif
