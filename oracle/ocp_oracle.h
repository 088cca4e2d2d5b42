/* TEST INFRASTRUCTURE ONLY -- CPU oracle for the MPC hot path.
 *
 * Restates, in plain C / IEEE double, the optimal-control problems the reference builds with
 * CasADi in car_racing/control/control.py (mpc_lti :198-248, mpccbf :476-607,
 * mpc_multi_agents :251-473) and solves them with a dense-linear-algebra interior point
 * method that follows the published IPOPT algorithm (Waechter & Biegler, Math. Prog. 106(1),
 * 2006) -- the solver inside the CasADi 3.5.5 wheel the reference pins (requirements.txt:6);
 * neither is present under /root/reference nor installable here.
 *
 * PARITY: the SOLVER ALGORITHM is UNPINNED -- the reference holds no golden vector / asserting test for
 * this path and CasADi/IPOPT cannot be run here.  The PROBLEM STATEMENT is pinned: the reference's
 * own functions, run under a recording stand-in for casadi, give the cost / constraint values of
 * tests/golden/nlp_golden.npz, and tests/test_reference_statement.py checks the oracle's problem functions and the
 * packed problem data against them.  The iLQR restatement (ilqr_oracle.c) IS pinned end to end against the
 * reference's own control.ilqr (tests/golden/).
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg
 * may link or call this library.  The product (car_racing_b200/) never does.
 */
#ifndef OCP_ORACLE_H
#define OCP_ORACLE_H

#ifdef __cplusplus
extern "C" {
#endif

#define ORC_NMAX 64 /* max horizon */
#define ORC_MMAX 8  /* max rivals inside the +-2*vx window */

typedef struct {
    int N;              /* horizon (num_horizon) */
    int M;              /* rivals kept by the proximity filter (control.py:499-523) */
    double A[36], B[12]; /* row-major LTI model (control.py:566-570) */
    double Q[36], R[4];  /* stage weights (base.py:231-232) */
    double x0[6];        /* xcurv (control.py:497) */
    double xt[(ORC_NMAX + 1) * 6]; /* per-stage target; mpccbf repeats xtarget, mpc_multi_agents :373-382 */
    double umax[2];      /* delta_max, a_max (control.py:572-576) */
    double vmin, vmax;   /* control.py:582-583 */
    double width;        /* track.width (control.py:585-586) */
    double alpha;        /* CBF decay (0.8 base.py:282 / 0.6 control.py:285) */
    double margin;       /* safety_margin 0.2 (:527) / 0.15 (:311) */
    double L, W;         /* l_agent+l_obs, w_agent+w_obs (:532-535) */
    double slack_w;      /* 10000 (:560) */
    double obs_s[ORC_MMAX][ORC_NMAX + 1];  /* rival s prediction, row 4 of obs_traj */
    double obs_ey[ORC_MMAX][ORC_NMAX + 1]; /* rival ey prediction, row 5 */
    double lap_off[ORC_MMAX]; /* (num_cycle_ego-num_cycle_obs)*lap_length, applied to h only (:539-542) */
    /* planner-candidate extensions (planning/overtake_traj_planner.py:248-379); all off when zero */
    int per_stage_bounds;               /* 1: use xlb/xub instead of vmin/vmax/width */
    double xlb[(ORC_NMAX + 1) * 2];     /* per stage i: lower bound on (vx_i, ey_i), -HUGE_VAL = none (:276-324) */
    double xub[(ORC_NMAX + 1) * 2];     /* per stage i: upper bound on (vx_i, ey_i), +HUGE_VAL = none */
    double wd[ORC_NMAX];                /* cost += wd[i]*(ey_{i+1}-ey_i)^2, i=0..N-1 (:325-327: 30 for i=1..N-2) */
    int per_rival_size;                 /* 1: use Lj/Wj instead of L/W (control.py:530-535 reads the size of every rival) */
    double Lj[ORC_MMAX], Wj[ORC_MMAX];
} orc_problem;

typedef struct {
    double tol;            /* 1e-8  (IPOPT tol) */
    int max_iter;          /* 3000 in IPOPT; we stop earlier by default */
    double mu_init;        /* 0.1 */
    double rho;            /* l1 weight of the elastic CBF rows, scaled-objective units */
    double bound_push;     /* 1e-2 */
    double bound_frac;     /* 1e-2 */
    double acceptable_tol; /* 1e-6 */
    int acceptable_iter;   /* 15 */
    double max_grad;       /* nlp_scaling_max_gradient = 100 */
    int start;             /* 0: u = 0 roll-out of x_0 (DESIGN.md deviation 1); 1: w = 0, what Opti/IPOPT start from
                              (no opti.set_initial in control.py:476-607), pushed into the bounds */
    int max_reset;         /* elastic slack resets when the line search fails (stand-in for IPOPT's restoration phase), 5 */
} orc_options;

typedef struct {
    double x[(ORC_NMAX + 1) * 6];
    double u[ORC_NMAX * 2];
    double sigma[ORC_MMAX * (ORC_NMAX + 1)]; /* [j][i] */
    double cost;        /* unscaled objective, as the reference's `cost` expression */
    double kkt_err;     /* final scaled optimality error E_0 */
    double elastic_max; /* max_r t_r (0 => CBF rows hold exactly) */
    int status;         /* 0 converged, 1 max_iter, 2 line search failed, 3 inertia failed, 4 x_0 violates its own stage-0
                           bound rows (control.py:582-586 impose them on the fixed x_0: IPOPT reports infeasibility) */
    int iters;
    int n_refactor;     /* inertia-correction refactorisations */
    int n_backtrack;    /* line-search halvings */
    /* multipliers for the KKT certificate */
    double lam[ORC_NMAX * 6];
    double y[ORC_MMAX * ORC_NMAX];   /* CBF rows, unscaled problem units */
    double df;                       /* objective scaling used */
} orc_result;

void orc_default_options(orc_options *o);
int orc_solve(const orc_problem *p, const orc_options *o, orc_result *r);
/* batch over B problems with nthreads pthreads (<=0: one thread) */
int orc_solve_batch(const orc_problem *p, int B, const orc_options *o, orc_result *r, int nthreads);

/* ---- iLQR restatement (control.py:64-195, ilqr_helper.py:4-55) ---- */
typedef struct {
    int N;               /* num_horizon (50) */
    int max_iter;        /* 150 (base.py:176) */
    double A[36], B[12], Q[36], R[4];
    double x0[6], xt[6];
    double obs_s[ORC_NMAX + 1], obs_ey[ORC_NMAX + 1]; /* last rival's prediction (control.py:100-105) */
    double lap_off;      /* (num_cycle_ego-num_cycle_obs)*lap_length (ilqr_helper.py:35-37) */
    double L, W;         /* l_agent+l_obs, w_agent+w_obs */
} orc_ilqr_problem;

typedef struct {
    double u0[2];
    double cost;
    int iters;
    int converged;
    double u[ORC_NMAX * 2];
    double x[(ORC_NMAX + 1) * 6];
} orc_ilqr_result;

int orc_ilqr_solve(const orc_ilqr_problem *p, orc_ilqr_result *r);
int orc_ilqr_solve_batch(const orc_ilqr_problem *p, int B, orc_ilqr_result *r, int nthreads);

/* ---- LMPC restatement (control.py:610-730; lmpc_oracle.c) ---- */
#define ORC_KMAX 64 /* max safe-set points (num_ss_points = 44, base.py:358) */
typedef struct {
    int N, K;                          /* num_horizon (12), number of selected safe-set points (44) */
    double Q[36], R[4], dR[4];         /* matrix_Q, matrix_R, matrix_dR (base.py:354-357) */
    double xtrk[6];                    /* x_track = [5,0,0,0,0,0] (control.py:649) */
    double umax[2], vmax, width;       /* delta_max, a_max, v_max, lap_width (control.py:658-666) */
    double x0[6], u_old[2];            /* xcurv (:650), u_old (:675) */
    double A[ORC_NMAX * 36], B[ORC_NMAX * 12], C[ORC_NMAX * 6];   /* matrix_Atv/Btv/Ctv[i] (:653-656) */
    double SS[6 * ORC_KMAX];           /* ss_point_selected_tot, row-major 6 x K (:690-691) */
    double Qfun[ORC_KMAX];             /* Qfun_selected_tot (:695) */
} orc_lmpc_problem;

typedef struct {
    double x[(ORC_NMAX + 1) * 6], u[ORC_NMAX * 2], lambda[ORC_KMAX];
    double cost, kkt_err;
    int status, iters;
} orc_lmpc_result;

int orc_lmpc_solve(const orc_lmpc_problem *p, const orc_options *o, orc_lmpc_result *r);
int orc_lmpc_solve_batch(const orc_lmpc_problem *p, int B, const orc_options *o, orc_lmpc_result *r, int nthreads);

#ifdef __cplusplus
}
#endif
#endif
