/* oracle/usv_oracle.h -- TEST INFRASTRUCTURE ONLY (see usv_oracle.c). */
#ifndef USV_ORACLE_H_
#define USV_ORACLE_H_

#define USVO_MAXN 128  /* max horizon */
#define USVO_NVM 16    /* max nu+nx per stage */
#define USVO_NXM 14    /* max nx */
#define USVO_NBM 16    /* max box rows per stage */
#define USVO_NGM 32    /* max general / nonlinear rows per stage */
#define USVO_NCM (USVO_NBM + USVO_NGM)
#define USVO_NSM 16    /* max soft rows per stage (each has a lower and an upper slack variable) */
#define USVO_NVT (USVO_NVM + 2 * USVO_NSM)  /* [u; x; sl; su] */
#define USVO_NCT (2 * USVO_NCM + 2 * USVO_NSM) /* [lb lg | ub ug | ls | us] */

/* ---- stage-wise OCP QP, HPIPM conventions (SURVEY.md section 3e "data conventions") ---- */
typedef struct
{
    int nx, nu, nb, ng, ns;                  /* this stage; ns soft rows, slack variables sl, su follow [u;x] */
    int idxs_rev[USVO_NCM];                  /* row -> slack index or -1 (HPIPM idxs_rev) */
    double Z[2 * USVO_NSM];                  /* diagonal Hessian of the slacks [Zl; Zu] */
    int idxb[USVO_NBM];                      /* box rows index into [u;x] */
    double BAt[USVO_NVM * USVO_NXM];         /* (nu+nx) x nx_next, column-major, ld = nu+nx : [B';A'] */
    double b[USVO_NXM];
    double RSQ[USVO_NVM * USVO_NVM];         /* (nu+nx)^2 column-major, lower triangle read */
    double rq[USVO_NVT];                     /* gradient of [u;x] then of the slacks (z) */
    double DCt[USVO_NVM * USVO_NGM];         /* (nu+nx) x ng column-major */
    double d[USVO_NCT];                      /* [lb lg | -ub -ug]-style residuals: lower rows then upper rows, then the
                                                lower bounds of the slacks [ls | us] */
} usvo_qp_stage;

typedef struct
{
    double ux[USVO_NVT], pi[USVO_NXM], lam[USVO_NCT], t[USVO_NCT];
} usvo_qp_sol_stage;

typedef struct
{
    double mu0, alpha_min, tol_stat, tol_eq, tol_ineq, tol_comp, reg_prim, lam_min, t_min, tau_min;
    int iter_max, itref_corr_max, pred_corr, cond_pred_corr;
} usvo_ipm_arg;

typedef struct
{
    int iter, status;          /* HPIPM status: 0 ok, 1 max iter, 2 min step, 3 NaN */
    int solve_calls;           /* solve-only Riccati sweeps */
    int lq_wanted;             /* iterations where HPIPM would have switched to its LQ factorisation */
    double res[4], mu;
} usvo_ipm_info;

void usvo_ipm_arg_default(usvo_ipm_arg *arg, int sqp_mode, const double *tol4, int iter_max);
void usvo_qp_solve(int N, const usvo_qp_stage *qp, usvo_qp_sol_stage *sol, const usvo_ipm_arg *arg,
                   usvo_ipm_info *info);

/* flat-array front end used by the tests (layout of tests/refharness.py QP capture) */
void usvo_qp_solve_flat(int N, const int *dims, const double *BAbt, int sBAbt, const double *b, int sb,
                        const double *RSQrq, int sRSQ, const double *rqz, int srq, const double *DCt, int sDCt,
                        const double *d, int sd, const int *idxb, int sidxb, int sqp_mode, const double *tol4,
                        int iter_max, double *ux, int sux, double *pi, int spi, double *lam, double *t, int slam,
                        double *info_out);

/* ---- NMPC problem description (mirror of oracle/ref_harness.c icfg/dcfg) ---- */
typedef struct
{
    int model, N, K, num_steps, num_stages, nlp_type, max_iter, qp_iter_max, nbx, nbu;
    int idxbx[USVO_NBM];
    double dt, tol[4], uh;
    double W[USVO_NVM * USVO_NVM], We[USVO_NXM * USVO_NXM];  /* column-major ny x ny, nx x nx */
    double lbu[USVO_NBM], ubu[USVO_NBM], lbx[USVO_NBM], ubx[USVO_NBM];
    /* soft obstacle rows: the first nsh rows of h (idxsh = 0..nsh-1), one value of lsh, ush, zl, zu, Zl, Zu for all
     * (usv_guidance_ca1/acados_settings.py:105-178) */
    int nsh;
    double lsh, ush, zl, zu, Zl, Zu;
} usvo_problem;

void usvo_problem_init(usvo_problem *P, const int *icfg, const double *dcfg, const double *W, const double *We,
                       const double *lbu, const double *ubu, const int *idxbx, const double *lbx, const double *ubx);

/* discrete dynamics x+ = Phi(x,u) with forward sensitivities: A = dPhi/dx (nx x nx col-major), B = dPhi/du */
void usvo_integrate(const usvo_problem *P, const double *x, const double *u, double *xn, double *A, double *B);

int usvo_solve(const usvo_problem *P, const double *x0, const double *p, int p_per_stage, const double *lh,
               int lh_per_stage, const double *yref, int yref_per_stage, const double *yref_e, const double *xinit,
               const double *uinit, const double *piinit, double *x_out, double *u_out, double *pi_out,
               double *lam_out, double *t_out, double *stats);

/* usvo_solve + the slack values sl, su [N][nsh] of the solution (may be NULL) */
int usvo_solve_ex(const usvo_problem *P, const double *x0, const double *p, int p_per_stage, const double *lh,
                  int lh_per_stage, const double *yref, int yref_per_stage, const double *yref_e, const double *xinit,
                  const double *uinit, const double *piinit, double *x_out, double *u_out, double *pi_out,
                  double *lam_out, double *t_out, double *stats, double *sl_out, double *su_out);

double usvo_solve_batch(const int *icfg, const double *dcfg, const double *W, const double *We, const double *lbu,
                        const double *ubu, const int *idxbx, const double *lbx, const double *ubx, int B,
                        const double *x0, const double *p, int p_per_stage, const double *lh, int lh_per_stage,
                        const double *yref, int yref_per_stage, const double *yref_e, double *x_out, double *u_out,
                        double *stats, int nthreads);

#endif
