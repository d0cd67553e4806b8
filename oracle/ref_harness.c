/* oracle/ref_harness.c -- TEST INFRASTRUCTURE ONLY (never linked into the product).
 *
 * Drives the UNMODIFIED reference C stack (acados + HPIPM + BLASFEO, compiled by
 * oracle/Makefile from /root/reference into oracle/_ref/libacados_ref.so) on the USV
 * collision-avoidance OCP and on the reference's own pendulum test OCP.
 *
 * The reference normally gets this file from its code generator: `acados_create()` is a Tera
 * template (interfaces/acados_template/acados_template/c_templates_tera/acados_solver.in.c:179-1739)
 * and the model callbacks are CasADi-generated C.  Neither generator runs offline, so this file
 * is the hand-rendered equivalent: usvref_create() performs, in the same order, the calls the
 * template would emit for {LINEAR_LS cost, BGH constraints, ERK, GAUSS_NEWTON,
 * PARTIAL_CONDENSING_HPIPM, SQP | SQP_RTI}, and the callbacks follow the CasADi C ABI consumed by
 * external_function_param_casadi (acados/utils/external_function_generic.c:989-1232).
 * All solver arithmetic executed is the reference's own.
 *
 * It also interposes two HPIPM symbols (the harness is earlier in the lookup scope than
 * libacados_ref.so) purely to OBSERVE: d_ocp_qp_ipm_solve (QP capture for QP-level parity
 * tests, IPM iteration totals) and d_ocp_qp_fact_lq_solve_kkt_step (counts LQ fallbacks).
 */
#define _GNU_SOURCE
#include <dlfcn.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#include <pthread.h>

#include "acados/utils/types.h"
#include "acados/utils/external_function_generic.h"
#include "acados_c/external_function_interface.h"
#include "acados_c/ocp_nlp_interface.h"
#include "blasfeo/include/blasfeo_d_aux.h"
#include "hpipm/include/hpipm_d_ocp_qp.h"
#include "hpipm/include/hpipm_d_ocp_qp_sol.h"
#include "hpipm/include/hpipm_d_ocp_qp_ipm.h"

#include "usv_models.h"

/* defined in ocp_nlp_interface.c:909 but not declared in the header at this commit */
void ocp_nlp_eval_residuals(ocp_nlp_solver *solver, ocp_nlp_in *nlp_in, ocp_nlp_out *nlp_out);

/* ------------------------------------------------------------------------------------------
 * process-wide model configuration (CasADi-ABI sparsity callbacks carry no context pointer;
 * the reference's generated library is global state too, acados_solver.in.c:101-107)
 * ---------------------------------------------------------------------------------------- */
static int g_model = 0, g_nx = 6, g_nu = 2, g_K = 0, g_hx = 0, g_hy = 1;

#define MAXNX 8
#define MAXK 64

static int sp_buf[16][3];
static const int *sp_dense(int slot, int nrow, int ncol)
{
    sp_buf[slot][0] = nrow; sp_buf[slot][1] = ncol; sp_buf[slot][2] = 1;
    return sp_buf[slot];
}

/* ---- <model>_expl_ode_fun : (x,u,p) -> f ---- */
static int ode_fun(const double **arg, double **res, int *iw, double *w, void *mem)
{
    usvm_f_jac(g_model, arg[0], arg[1], res[0], NULL, NULL);
    return 0;
}
static int ode_work(int *a, int *r, int *iw, int *w) { *a = 3; *r = 1; *iw = 0; *w = 0; return 0; }
static const int *ode_sp_in(int i)
{
    if (i == 0) return sp_dense(0, g_nx, 1);
    if (i == 1) return sp_dense(1, g_nu, 1);
    return sp_dense(2, 2 * g_K, 1);
}
static const int *ode_sp_out(int i) { return sp_dense(3, g_nx, 1); }
static int ode_n_in(void) { return 3; }
static int ode_n_out(void) { return 1; }

/* ---- <model>_expl_vde_forw : (x,Sx,Sp,u,p) -> (f, Jx*Sx, Jx*Sp + Ju) ---- */
static int vde_fun(const double **arg, double **res, int *iw, double *w, void *mem)
{
    const int nx = g_nx, nu = g_nu;
    double Jx[MAXNX * MAXNX], Ju[MAXNX * MAXNX];
    usvm_f_jac(g_model, arg[0], arg[3], res[0], Jx, Ju);
    const double *Sx = arg[1], *Sp = arg[2];
    for (int j = 0; j < nx; j++)
        for (int i = 0; i < nx; i++)
        {
            double acc = 0.0;
            for (int k = 0; k < nx; k++) acc += Jx[i + nx * k] * Sx[k + nx * j];
            res[1][i + nx * j] = acc;
        }
    for (int j = 0; j < nu; j++)
        for (int i = 0; i < nx; i++)
        {
            double acc = Ju[i + nx * j];
            for (int k = 0; k < nx; k++) acc += Jx[i + nx * k] * Sp[k + nx * j];
            res[2][i + nx * j] = acc;
        }
    return 0;
}
static int vde_work(int *a, int *r, int *iw, int *w) { *a = 5; *r = 3; *iw = 0; *w = 0; return 0; }
static const int *vde_sp_in(int i)
{
    if (i == 0) return sp_dense(4, g_nx, 1);
    if (i == 1) return sp_dense(5, g_nx, g_nx);
    if (i == 2) return sp_dense(6, g_nx, g_nu);
    if (i == 3) return sp_dense(7, g_nu, 1);
    return sp_dense(2, 2 * g_K, 1);
}
static const int *vde_sp_out(int i)
{
    if (i == 0) return sp_dense(3, g_nx, 1);
    if (i == 1) return sp_dense(5, g_nx, g_nx);
    return sp_dense(6, g_nx, g_nu);
}
static int vde_n_in(void) { return 5; }
static int vde_n_out(void) { return 3; }

/* ---- <model>_constr_h_fun : (x,u,z,p) -> h ;  _constr_h_fun_jac_uxt_zt -> (h, (dh/d[u;x])^T, (dh/dz)) ---- */
static int h_fun(const double **arg, double **res, int *iw, double *w, void *mem)
{
    usvm_obstacle_h_at(g_hx, g_hy, g_K, arg[0], arg[3], res[0], NULL, NULL);
    return 0;
}
static int hjac_fun(const double **arg, double **res, int *iw, double *w, void *mem)
{
    double gX[MAXK], gY[MAXK];
    const int nv = g_nu + g_nx;
    usvm_obstacle_h_at(g_hx, g_hy, g_K, arg[0], arg[3], res[0], gX, gY);
    for (int i = 0; i < g_K; i++)
    {
        for (int r = 0; r < nv; r++) res[1][r + nv * i] = 0.0;
        res[1][g_nu + g_hx + nv * i] = gX[i];
        res[1][g_nu + g_hy + nv * i] = gY[i];
    }
    return 0;
}
static int h_work(int *a, int *r, int *iw, int *w) { *a = 4; *r = 1; *iw = 0; *w = 0; return 0; }
static int hjac_work(int *a, int *r, int *iw, int *w) { *a = 4; *r = 3; *iw = 0; *w = 0; return 0; }
static const int *h_sp_in(int i)
{
    if (i == 0) return sp_dense(8, g_nx, 1);
    if (i == 1) return sp_dense(9, g_nu, 1);
    if (i == 2) return sp_dense(10, 0, 1);
    return sp_dense(2, 2 * g_K, 1);
}
static const int *h_sp_out(int i)
{
    if (i == 0) return sp_dense(11, g_K, 1);
    if (i == 1) return sp_dense(12, g_nu + g_nx, g_K);
    return sp_dense(13, 0, g_K);
}
static int h_n_in(void) { return 4; }
static int h_n_out(void) { return 1; }
static int hjac_n_out(void) { return 3; }

/* ------------------------------------------------------------------------------------------
 * observation hooks (thread-local)
 * ---------------------------------------------------------------------------------------- */
typedef struct
{
    int enabled;      /* capture the QP of call number `want` (0-based) since last reset */
    int want;
    int calls;
    long ipm_iters;   /* sum of HPIPM iterations since last reset */
    long lq_calls;    /* LQ fallbacks since last reset */
    long solve_calls; /* solve-only Riccati sweeps (corrector, centering, refinement) since last reset */
    /* captured (flat, per stage, column-major dense) */
    int N, got;
    int nx[128], nu[128], nb[128], ng[128];
    double *BAbt, *b, *RSQrq, *rqz, *DCt, *d;  /* each sized by the caller, stage stride fixed */
    int *idxb;
    double *ux, *pi, *lam, *t;
    int iter, status;
    int sBAbt, sb, sRSQ, srq, sDCt, sd, sidxb, sux, spi, slam;  /* strides per stage */
} qp_tap;

static __thread qp_tap g_tap;

void d_ocp_qp_ipm_solve(struct d_ocp_qp *qp, struct d_ocp_qp_sol *qp_sol, struct d_ocp_qp_ipm_arg *arg,
                        struct d_ocp_qp_ipm_ws *ws)
{
    static void (*real)(struct d_ocp_qp *, struct d_ocp_qp_sol *, struct d_ocp_qp_ipm_arg *,
                        struct d_ocp_qp_ipm_ws *) = NULL;
    if (!real) real = dlsym(RTLD_NEXT, "d_ocp_qp_ipm_solve");
    qp_tap *tp = &g_tap;
    int cap = tp->enabled && tp->calls == tp->want;
    if (cap)
    {
        int N = qp->dim->N;
        tp->N = N;
        for (int k = 0; k <= N; k++)
        {
            int nx = qp->dim->nx[k], nu = qp->dim->nu[k], nb = qp->dim->nb[k], ng = qp->dim->ng[k];
            int nv = nu + nx;
            tp->nx[k] = nx; tp->nu[k] = nu; tp->nb[k] = nb; tp->ng[k] = ng;
            if (k < N)
            {
                int nx1 = qp->dim->nx[k + 1];
                blasfeo_unpack_dmat(nv, nx1, qp->BAbt + k, 0, 0, tp->BAbt + k * tp->sBAbt, nv);
                blasfeo_unpack_dvec(nx1, qp->b + k, 0, tp->b + k * tp->sb, 1);
            }
            blasfeo_unpack_dmat(nv, nv, qp->RSQrq + k, 0, 0, tp->RSQrq + k * tp->sRSQ, nv);
            blasfeo_unpack_dvec(nv, qp->rqz + k, 0, tp->rqz + k * tp->srq, 1);
            if (ng > 0) blasfeo_unpack_dmat(nv, ng, qp->DCt + k, 0, 0, tp->DCt + k * tp->sDCt, nv);
            blasfeo_unpack_dvec(2 * nb + 2 * ng, qp->d + k, 0, tp->d + k * tp->sd, 1);
            for (int j = 0; j < nb; j++) tp->idxb[k * tp->sidxb + j] = qp->idxb[k][j];
        }
    }
    real(qp, qp_sol, arg, ws);
    tp->ipm_iters += ws->iter;
    if (cap)
    {
        int N = qp->dim->N;
        for (int k = 0; k <= N; k++)
        {
            int nv = tp->nu[k] + tp->nx[k], nc = 2 * tp->nb[k] + 2 * tp->ng[k];
            blasfeo_unpack_dvec(nv, qp_sol->ux + k, 0, tp->ux + k * tp->sux, 1);
            if (k < N) blasfeo_unpack_dvec(tp->nx[k + 1], qp_sol->pi + k, 0, tp->pi + k * tp->spi, 1);
            blasfeo_unpack_dvec(nc, qp_sol->lam + k, 0, tp->lam + k * tp->slam, 1);
            blasfeo_unpack_dvec(nc, qp_sol->t + k, 0, tp->t + k * tp->slam, 1);
        }
        tp->iter = ws->iter;
        tp->status = ws->status;
        tp->got = 1;
    }
    tp->calls++;
}

void d_ocp_qp_fact_lq_solve_kkt_step(struct d_ocp_qp *qp, struct d_ocp_qp_sol *qp_sol,
                                     struct d_ocp_qp_ipm_arg *arg, struct d_ocp_qp_ipm_ws *ws)
{
    static void (*real)(struct d_ocp_qp *, struct d_ocp_qp_sol *, struct d_ocp_qp_ipm_arg *,
                        struct d_ocp_qp_ipm_ws *) = NULL;
    if (!real) real = dlsym(RTLD_NEXT, "d_ocp_qp_fact_lq_solve_kkt_step");
    g_tap.lq_calls++;
    real(qp, qp_sol, arg, ws);
}

void d_ocp_qp_solve_kkt_step(struct d_ocp_qp *qp, struct d_ocp_qp_sol *qp_sol, struct d_ocp_qp_ipm_arg *arg,
                             struct d_ocp_qp_ipm_ws *ws)
{
    static void (*real)(struct d_ocp_qp *, struct d_ocp_qp_sol *, struct d_ocp_qp_ipm_arg *,
                        struct d_ocp_qp_ipm_ws *) = NULL;
    if (!real) real = dlsym(RTLD_NEXT, "d_ocp_qp_solve_kkt_step");
    g_tap.solve_calls++;
    real(qp, qp_sol, arg, ws);
}

/* ------------------------------------------------------------------------------------------
 * solver context == what acados_create() builds (one per thread)
 * ---------------------------------------------------------------------------------------- */
/* icfg[ICFG_NSH] soft obstacle rows (the first nsh rows of h: idxsh = 0..nsh-1) with the slack data dcfg[DCFG_LSH..]:
 * lsh, ush, zl, zu, Zl, Zu, one value for all soft rows (usv_guidance_ca1/acados_settings.py:105-178) */
enum { ICFG_MODEL, ICFG_N, ICFG_K, ICFG_NUM_STEPS, ICFG_NUM_STAGES, ICFG_NLP_TYPE, ICFG_MAX_ITER,
       ICFG_QP_ITER_MAX, ICFG_COND_N, ICFG_NBX, ICFG_NBU, ICFG_PRINT, ICFG_NSH, ICFG_LEN };
enum { DCFG_DT, DCFG_TOL_STAT, DCFG_TOL_EQ, DCFG_TOL_INEQ, DCFG_TOL_COMP, DCFG_UH, DCFG_LSH, DCFG_USH, DCFG_ZL, DCFG_ZU,
       DCFG_ZZL, DCFG_ZZU, DCFG_LEN };

typedef struct
{
    int model, N, K, nx, nu, ny, nye, np, nbx, nbu, nlp_type, nsh;
    ocp_nlp_plan *plan;
    ocp_nlp_config *config;
    ocp_nlp_dims *dims;
    ocp_nlp_in *in;
    ocp_nlp_out *out;
    void *opts;
    ocp_nlp_solver *solver;
    external_function_param_casadi *vde, *ode, *hjac, *hfun;
} usvref_ctx;

static void bind(external_function_param_casadi *f, int (*fun)(const double **, double **, int *, double *, void *),
                 int (*work)(int *, int *, int *, int *), const int *(*spi)(int), const int *(*spo)(int),
                 int (*nin)(void), int (*nout)(void), int np)
{
    f->casadi_fun = fun; f->casadi_work = work; f->casadi_sparsity_in = spi; f->casadi_sparsity_out = spo;
    f->casadi_n_in = (int (*)()) nin; f->casadi_n_out = (int (*)()) nout;
    external_function_param_casadi_create(f, np);
}

/* icfg/dcfg: see enums.  W (ny x ny), We (nx x nx) column-major; lbu/ubu (nbu, idxbu=0..nbu-1);
 * idxbx/lbx/ubx (nbx) applied on stages 1..N-1 (template :1202). */
void *usvref_create(const int *icfg, const double *dcfg, const double *W, const double *We, const double *lbu,
                    const double *ubu, const int *idxbx, const double *lbx, const double *ubx)
{
    usvref_ctx *c = calloc(1, sizeof(usvref_ctx));
    const int N = icfg[ICFG_N], K = icfg[ICFG_K];
    c->model = icfg[ICFG_MODEL]; c->N = N; c->K = K; c->nlp_type = icfg[ICFG_NLP_TYPE];
    usvm_dims(c->model, &c->nx, &c->nu);
    const int nxm = c->nx, num = c->nu;
    c->ny = nxm + num; c->nye = nxm; c->np = 2 * K; c->nbx = icfg[ICFG_NBX]; c->nbu = icfg[ICFG_NBU];
    g_model = c->model; g_nx = nxm; g_nu = num; g_K = K;   /* contexts are created serially */
    usvm_pos_states(c->model, &g_hx, &g_hy);
    c->nsh = icfg[ICFG_NSH];
    const int nsh = c->nsh;

    /* plan & config (template :186-239) */
    c->plan = ocp_nlp_plan_create(N);
    c->plan->nlp_solver = c->nlp_type == 0 ? SQP : SQP_RTI;
    c->plan->ocp_qp_solver_plan.qp_solver = PARTIAL_CONDENSING_HPIPM;
    for (int i = 0; i <= N; i++) c->plan->nlp_cost[i] = LINEAR_LS;
    for (int i = 0; i < N; i++)
    {
        c->plan->nlp_dynamics[i] = CONTINUOUS_MODEL;
        c->plan->sim_solver_plan[i].sim_solver = ERK;
    }
    for (int i = 0; i <= N; i++) c->plan->nlp_constraints[i] = BGH;
    c->config = ocp_nlp_config_create(*c->plan);

    /* dims (template :245-357) */
    int nx[N + 1], nu[N + 1], nz[N + 1], ns[N + 1], nbx[N + 1], nbu[N + 1], ng[N + 1], nh[N + 1], ny[N + 1],
        nbxe[N + 1], zero = 0;
    for (int i = 0; i <= N; i++)
    {
        nx[i] = nxm; nu[i] = num; nz[i] = 0; ns[i] = i < N ? nsh : 0; ny[i] = c->ny;
        nbx[i] = c->nbx; nbu[i] = c->nbu; ng[i] = 0; nh[i] = K; nbxe[i] = 0;
    }
    nbx[0] = nxm; nbxe[0] = nxm;
    nu[N] = 0; ny[N] = c->nye; nbx[N] = 0; nbu[N] = 0; nh[N] = 0;
    c->dims = ocp_nlp_dims_create(c->config);
    ocp_nlp_dims_set_opt_vars(c->config, c->dims, "nx", nx);
    ocp_nlp_dims_set_opt_vars(c->config, c->dims, "nu", nu);
    ocp_nlp_dims_set_opt_vars(c->config, c->dims, "nz", nz);
    ocp_nlp_dims_set_opt_vars(c->config, c->dims, "ns", ns);
    for (int i = 0; i <= N; i++)
    {
        ocp_nlp_dims_set_constraints(c->config, c->dims, i, "nbx", &nbx[i]);
        ocp_nlp_dims_set_constraints(c->config, c->dims, i, "nbu", &nbu[i]);
        ocp_nlp_dims_set_constraints(c->config, c->dims, i, "nsbx", &zero);
        ocp_nlp_dims_set_constraints(c->config, c->dims, i, "nsbu", &zero);
        ocp_nlp_dims_set_constraints(c->config, c->dims, i, "ng", &ng[i]);
        ocp_nlp_dims_set_constraints(c->config, c->dims, i, "nsg", &zero);
        ocp_nlp_dims_set_constraints(c->config, c->dims, i, "nbxe", &nbxe[i]);
    }
    for (int i = 0; i < N; i++)
    {
        if (K > 0)
        {
            ocp_nlp_dims_set_constraints(c->config, c->dims, i, "nh", &nh[i]);
            ocp_nlp_dims_set_constraints(c->config, c->dims, i, "nsh", (void *) &nsh);
        }
        ocp_nlp_dims_set_cost(c->config, c->dims, i, "ny", &ny[i]);
    }
    ocp_nlp_dims_set_constraints(c->config, c->dims, N, "nh", &nh[N]);
    ocp_nlp_dims_set_constraints(c->config, c->dims, N, "nsh", &zero);
    ocp_nlp_dims_set_cost(c->config, c->dims, N, "ny", &ny[N]);

    /* external functions, one struct per stage (template :411-510) */
    c->vde = malloc(sizeof(external_function_param_casadi) * N);
    c->ode = malloc(sizeof(external_function_param_casadi) * N);
    for (int i = 0; i < N; i++)
    {
        bind(&c->vde[i], vde_fun, vde_work, vde_sp_in, vde_sp_out, vde_n_in, vde_n_out, c->np);
        bind(&c->ode[i], ode_fun, ode_work, ode_sp_in, ode_sp_out, ode_n_in, ode_n_out, c->np);
    }
    if (K > 0)
    {
        c->hjac = malloc(sizeof(external_function_param_casadi) * N);
        c->hfun = malloc(sizeof(external_function_param_casadi) * N);
        for (int i = 0; i < N; i++)
        {
            bind(&c->hjac[i], hjac_fun, hjac_work, h_sp_in, h_sp_out, h_n_in, hjac_n_out, c->np);
            bind(&c->hfun[i], h_fun, h_work, h_sp_in, h_sp_out, h_n_in, h_n_out, c->np);
        }
    }

    /* nlp_in (template :796-1449) */
    c->in = ocp_nlp_in_create(c->config, c->dims);
    double dt = dcfg[DCFG_DT];
    for (int i = 0; i < N; i++)
    {
        ocp_nlp_in_set(c->config, c->dims, c->in, i, "Ts", &dt);
        ocp_nlp_cost_model_set(c->config, c->dims, c->in, i, "scaling", &dt);
    }
    for (int i = 0; i < N; i++)
    {
        ocp_nlp_dynamics_model_set(c->config, c->dims, c->in, i, "expl_vde_forw", &c->vde[i]);
        ocp_nlp_dynamics_model_set(c->config, c->dims, c->in, i, "expl_ode_fun", &c->ode[i]);
    }
    /* cost: Vx=[I;0], Vu=[0;I] (column-major ny x nx / ny x nu), yref=0 */
    double *Vx = calloc(c->ny * nxm, sizeof(double)), *Vu = calloc(c->ny * num, sizeof(double));
    double *Vxe = calloc(c->nye * nxm, sizeof(double)), *yref0 = calloc(c->ny, sizeof(double));
    for (int j = 0; j < nxm; j++) { Vx[j + c->ny * j] = 1.0; Vxe[j + c->nye * j] = 1.0; }
    for (int j = 0; j < num; j++) Vu[nxm + j + c->ny * j] = 1.0;
    for (int i = 0; i < N; i++)
    {
        ocp_nlp_cost_model_set(c->config, c->dims, c->in, i, "W", (void *) W);
        ocp_nlp_cost_model_set(c->config, c->dims, c->in, i, "yref", yref0);
        ocp_nlp_cost_model_set(c->config, c->dims, c->in, i, "Vx", Vx);
        ocp_nlp_cost_model_set(c->config, c->dims, c->in, i, "Vu", Vu);
    }
    ocp_nlp_cost_model_set(c->config, c->dims, c->in, N, "yref", yref0);
    ocp_nlp_cost_model_set(c->config, c->dims, c->in, N, "W", (void *) We);
    ocp_nlp_cost_model_set(c->config, c->dims, c->in, N, "Vx", Vxe);
    free(Vx); free(Vu); free(Vxe); free(yref0);
    if (nsh > 0)
    {
        /* slack penalties (template :880-930): zl, zu, Zl, Zu per stage */
        double zl[MAXK], zu[MAXK], Zl[MAXK], Zu[MAXK];
        for (int j = 0; j < nsh; j++) { zl[j] = dcfg[DCFG_ZL]; zu[j] = dcfg[DCFG_ZU]; Zl[j] = dcfg[DCFG_ZZL]; Zu[j] = dcfg[DCFG_ZZU]; }
        for (int i = 0; i < N; i++)
        {
            ocp_nlp_cost_model_set(c->config, c->dims, c->in, i, "Zl", Zl);
            ocp_nlp_cost_model_set(c->config, c->dims, c->in, i, "Zu", Zu);
            ocp_nlp_cost_model_set(c->config, c->dims, c->in, i, "zl", zl);
            ocp_nlp_cost_model_set(c->config, c->dims, c->in, i, "zu", zu);
        }
    }

    /* constraints */
    int idxbx0[MAXNX], idxbu[MAXNX];
    double zeros[MAXNX] = {0};
    for (int j = 0; j < nxm; j++) idxbx0[j] = j;
    for (int j = 0; j < c->nbu; j++) idxbu[j] = j;
    ocp_nlp_constraints_model_set(c->config, c->dims, c->in, 0, "idxbx", idxbx0);
    ocp_nlp_constraints_model_set(c->config, c->dims, c->in, 0, "lbx", zeros);
    ocp_nlp_constraints_model_set(c->config, c->dims, c->in, 0, "ubx", zeros);
    ocp_nlp_constraints_model_set(c->config, c->dims, c->in, 0, "idxbxe", idxbx0);
    if (c->nbu > 0)
        for (int i = 0; i < N; i++)
        {
            ocp_nlp_constraints_model_set(c->config, c->dims, c->in, i, "idxbu", idxbu);
            ocp_nlp_constraints_model_set(c->config, c->dims, c->in, i, "lbu", (void *) lbu);
            ocp_nlp_constraints_model_set(c->config, c->dims, c->in, i, "ubu", (void *) ubu);
        }
    if (c->nbx > 0)
        for (int i = 1; i < N; i++)
        {
            ocp_nlp_constraints_model_set(c->config, c->dims, c->in, i, "idxbx", (void *) idxbx);
            ocp_nlp_constraints_model_set(c->config, c->dims, c->in, i, "lbx", (void *) lbx);
            ocp_nlp_constraints_model_set(c->config, c->dims, c->in, i, "ubx", (void *) ubx);
        }
    if (K > 0)
    {
        double lh[MAXK], uh[MAXK];
        for (int j = 0; j < K; j++) { lh[j] = 0.0; uh[j] = dcfg[DCFG_UH]; }
        for (int i = 0; i < N; i++)
        {
            ocp_nlp_constraints_model_set(c->config, c->dims, c->in, i, "nl_constr_h_fun_jac", &c->hjac[i]);
            ocp_nlp_constraints_model_set(c->config, c->dims, c->in, i, "nl_constr_h_fun", &c->hfun[i]);
            ocp_nlp_constraints_model_set(c->config, c->dims, c->in, i, "lh", lh);
            ocp_nlp_constraints_model_set(c->config, c->dims, c->in, i, "uh", uh);
        }
        if (nsh > 0)
        {
            /* soft nonlinear rows (template :1395-1449): idxsh, lsh, ush */
            int idxsh[MAXK];
            double lsh[MAXK], ush[MAXK];
            for (int j = 0; j < nsh; j++) { idxsh[j] = j; lsh[j] = dcfg[DCFG_LSH]; ush[j] = dcfg[DCFG_USH]; }
            for (int i = 0; i < N; i++)
            {
                ocp_nlp_constraints_model_set(c->config, c->dims, c->in, i, "idxsh", idxsh);
                ocp_nlp_constraints_model_set(c->config, c->dims, c->in, i, "lsh", lsh);
                ocp_nlp_constraints_model_set(c->config, c->dims, c->in, i, "ush", ush);
            }
        }
    }

    /* opts (template :1456-1591) */
    c->opts = ocp_nlp_solver_opts_create(c->config, c->dims);
    int num_steps = icfg[ICFG_NUM_STEPS], num_stages = icfg[ICFG_NUM_STAGES], newton_iter = 3;
    bool jac_reuse = false;
    for (int i = 0; i < N; i++) ocp_nlp_solver_opts_set_at_stage(c->config, c->opts, i, "dynamics_num_steps", &num_steps);
    for (int i = 0; i < N; i++) ocp_nlp_solver_opts_set_at_stage(c->config, c->opts, i, "dynamics_num_stages", &num_stages);
    for (int i = 0; i < N; i++) ocp_nlp_solver_opts_set_at_stage(c->config, c->opts, i, "dynamics_newton_iter", &newton_iter);
    for (int i = 0; i < N; i++) ocp_nlp_solver_opts_set_at_stage(c->config, c->opts, i, "dynamics_jac_reuse", &jac_reuse);
    double step_length = 1.0, lm = 0.0;
    ocp_nlp_solver_opts_set(c->config, c->opts, "step_length", &step_length);
    ocp_nlp_solver_opts_set(c->config, c->opts, "levenberg_marquardt", &lm);
    int cond_N = icfg[ICFG_COND_N] > 0 ? icfg[ICFG_COND_N] : N;
    ocp_nlp_solver_opts_set(c->config, c->opts, "qp_cond_N", &cond_N);
    int qp_iter_max = icfg[ICFG_QP_ITER_MAX];
    ocp_nlp_solver_opts_set(c->config, c->opts, "qp_iter_max", &qp_iter_max);
    if (c->nlp_type == 0)
    {
        double ts = dcfg[DCFG_TOL_STAT], te = dcfg[DCFG_TOL_EQ], ti = dcfg[DCFG_TOL_INEQ], tc = dcfg[DCFG_TOL_COMP];
        ocp_nlp_solver_opts_set(c->config, c->opts, "tol_stat", &ts);
        ocp_nlp_solver_opts_set(c->config, c->opts, "tol_eq", &te);
        ocp_nlp_solver_opts_set(c->config, c->opts, "tol_ineq", &ti);
        ocp_nlp_solver_opts_set(c->config, c->opts, "tol_comp", &tc);
        int max_iter = icfg[ICFG_MAX_ITER], its = 0;
        ocp_nlp_solver_opts_set(c->config, c->opts, "max_iter", &max_iter);
        ocp_nlp_solver_opts_set(c->config, c->opts, "initialize_t_slacks", &its);
    }
    int print_level = icfg[ICFG_PRINT];
    ocp_nlp_solver_opts_set(c->config, c->opts, "print_level", &print_level);

    /* out + solver (template :1595-1730) */
    c->out = ocp_nlp_out_create(c->config, c->dims);
    for (int i = 0; i < N; i++)
    {
        ocp_nlp_out_set(c->config, c->dims, c->out, i, "x", zeros);
        ocp_nlp_out_set(c->config, c->dims, c->out, i, "u", zeros);
    }
    ocp_nlp_out_set(c->config, c->dims, c->out, N, "x", zeros);
    c->solver = ocp_nlp_solver_create(c->config, c->dims, c->opts);
    int status = ocp_nlp_precompute(c->solver, c->in, c->out);
    if (status != ACADOS_SUCCESS) { fprintf(stderr, "usvref: ocp_nlp_precompute failed\n"); return NULL; }
    return c;
}

void usvref_free(void *h)
{
    usvref_ctx *c = h;
    if (!c) return;
    ocp_nlp_solver_opts_destroy(c->opts);
    ocp_nlp_in_destroy(c->in);
    ocp_nlp_out_destroy(c->out);
    ocp_nlp_solver_destroy(c->solver);
    ocp_nlp_dims_destroy(c->dims);
    ocp_nlp_config_destroy(c->config);
    ocp_nlp_plan_destroy(c->plan);
    for (int i = 0; i < c->N; i++)
    {
        external_function_param_casadi_free(&c->vde[i]);
        external_function_param_casadi_free(&c->ode[i]);
        if (c->K > 0) { external_function_param_casadi_free(&c->hjac[i]); external_function_param_casadi_free(&c->hfun[i]); }
    }
    free(c->vde); free(c->ode); free(c->hjac); free(c->hfun);
    free(c);
}

/* One solve, driven exactly like NM/scripts/usv_guidance_ca1/main.py:116-175:
 *   set(0,lbx/ubx,x0); per stage set yref, p, lh; solve; get.
 * p: (N+1) x np when p_per_stage else np (broadcast); lh: N x K or K; yref: N x ny or ny.
 * xinit/uinit/piinit may be NULL (-> x_k = x0, u = 0, pi = 0: the cold start of SURVEY 8d).
 * stats[9] = {status, sqp_iter, qp_iter_total, res_stat, res_eq, res_ineq, res_comp, lq_calls, solve_calls}. */
int usvref_solve(void *h, const double *x0, const double *p, int p_per_stage, const double *lh, int lh_per_stage,
                 const double *yref, int yref_per_stage, const double *yref_e, const double *xinit,
                 const double *uinit, const double *piinit, double *x_out, double *u_out, double *pi_out,
                 double *lam_out, double *t_out, double *stats)
{
    usvref_ctx *c = h;
    const int N = c->N, nx = c->nx, nu = c->nu, K = c->K, np = c->np, ny = c->ny;
    double zeros[MAXNX] = {0};
    ocp_nlp_constraints_model_set(c->config, c->dims, c->in, 0, "lbx", (void *) x0);
    ocp_nlp_constraints_model_set(c->config, c->dims, c->in, 0, "ubx", (void *) x0);
    for (int i = 0; i < N; i++)
    {
        ocp_nlp_cost_model_set(c->config, c->dims, c->in, i, "yref", (void *) (yref + (yref_per_stage ? i * ny : 0)));
        if (K > 0)
        {
            double *pi_ = (double *) (p + (p_per_stage ? i * np : 0));
            c->vde[i].set_param(&c->vde[i], pi_);   /* acados_update_params, template :1742-1837 */
            c->ode[i].set_param(&c->ode[i], pi_);
            c->hjac[i].set_param(&c->hjac[i], pi_);
            c->hfun[i].set_param(&c->hfun[i], pi_);
            ocp_nlp_constraints_model_set(c->config, c->dims, c->in, i, "lh", (void *) (lh + (lh_per_stage ? i * K : 0)));
        }
    }
    ocp_nlp_cost_model_set(c->config, c->dims, c->in, N, "yref", (void *) yref_e);
    /* every solve starts from the state of a freshly created solver: x, u, pi as given, multipliers and slacks of
     * the inequalities zero.  (The SQP update blends lam, t as (1-alpha)*old + alpha*qp: a NaN left behind by a
     * diverged instance would otherwise survive 0*NaN and poison every later instance solved with this context.) */
    static const double zeros_ineq[2 * (MAXNX + 8 + MAXK)] = {0};
    for (int i = 0; i <= N; i++)
    {
        ocp_nlp_out_set(c->config, c->dims, c->out, i, "x", (void *) (xinit ? xinit + i * nx : x0));
        if (i < N)
        {
            ocp_nlp_out_set(c->config, c->dims, c->out, i, "u", (void *) (uinit ? uinit + i * nu : zeros));
            ocp_nlp_out_set(c->config, c->dims, c->out, i, "pi", (void *) (piinit ? piinit + i * nx : zeros));
        }
        ocp_nlp_out_set(c->config, c->dims, c->out, i, "lam", (void *) zeros_ineq);
        ocp_nlp_out_set(c->config, c->dims, c->out, i, "t", (void *) zeros_ineq);
        if (c->nsh > 0 && i < N) blasfeo_dvecse(2 * c->nsh, 0.0, c->out->ux + i, nu + nx);  /* slack values sl, su */
    }
    g_tap.ipm_iters = 0; g_tap.lq_calls = 0; g_tap.solve_calls = 0; g_tap.calls = 0; g_tap.got = 0;
    int status = ocp_nlp_solve(c->solver, c->in, c->out);
    int sqp_iter = 0;
    ocp_nlp_get(c->config, c->solver, "sqp_iter", &sqp_iter);
    if (c->nlp_type == 1) ocp_nlp_eval_residuals(c->solver, c->in, c->out);
    double r[4];
    ocp_nlp_get(c->config, c->solver, "res_stat", &r[0]);
    ocp_nlp_get(c->config, c->solver, "res_eq", &r[1]);
    ocp_nlp_get(c->config, c->solver, "res_ineq", &r[2]);
    ocp_nlp_get(c->config, c->solver, "res_comp", &r[3]);
    for (int i = 0; i <= N; i++)
    {
        ocp_nlp_out_get(c->config, c->dims, c->out, i, "x", x_out + i * nx);
        if (i < N)
        {
            ocp_nlp_out_get(c->config, c->dims, c->out, i, "u", u_out + i * nu);
            if (pi_out) ocp_nlp_out_get(c->config, c->dims, c->out, i, "pi", pi_out + i * nx);
        }
    }
    if (lam_out || t_out)
    {
        /* caller lays stages out with stride 2*(nbx_max+nbu+K) + 2*nsh, nbx_max = max(nx, nbx) */
        int nbm = (c->nbx > nx ? c->nbx : nx) + c->nbu + K;
        for (int i = 0; i <= N; i++)
        {
            if (lam_out) ocp_nlp_out_get(c->config, c->dims, c->out, i, "lam", lam_out + i * (2 * nbm + 2 * c->nsh));
            if (t_out) ocp_nlp_out_get(c->config, c->dims, c->out, i, "t", t_out + i * (2 * nbm + 2 * c->nsh));
        }
    }
    if (stats)
    {
        stats[0] = status; stats[1] = sqp_iter; stats[2] = (double) g_tap.ipm_iters;
        stats[3] = r[0]; stats[4] = r[1]; stats[5] = r[2]; stats[6] = r[3]; stats[7] = (double) g_tap.lq_calls; stats[8] = (double) g_tap.solve_calls;
    }
    return status;
}

/* slack values of the last solve: sl, su [N][nsh] (ocp_nlp_get_at_stage "sl"/"su" of the Python wrapper,
 * acados_ocp_solver.py:744-782 reads the same memory) */
void usvref_get_slacks(void *h, double *sl, double *su)
{
    usvref_ctx *c = h;
    for (int i = 0; i < c->N; i++)
    {
        blasfeo_unpack_dvec(c->nsh, c->out->ux + i, c->nu + c->nx, sl + i * c->nsh, 1);
        blasfeo_unpack_dvec(c->nsh, c->out->ux + i, c->nu + c->nx + c->nsh, su + i * c->nsh, 1);
    }
}

/* Arm the QP tap of the calling thread: capture the `want`-th QP (0-based) of the next solve.
 * Buffers are caller-owned; strides are per stage. */
void usvref_tap_arm(int want, double *BAbt, int sBAbt, double *b, int sb, double *RSQrq, int sRSQ, double *rqz, int srq,
                    double *DCt, int sDCt, double *d, int sd, int *idxb, int sidxb, double *ux, int sux, double *pi,
                    int spi, double *lam, double *t, int slam)
{
    qp_tap *tp = &g_tap;
    tp->enabled = 1; tp->want = want;
    tp->BAbt = BAbt; tp->sBAbt = sBAbt; tp->b = b; tp->sb = sb; tp->RSQrq = RSQrq; tp->sRSQ = sRSQ;
    tp->rqz = rqz; tp->srq = srq; tp->DCt = DCt; tp->sDCt = sDCt; tp->d = d; tp->sd = sd;
    tp->idxb = idxb; tp->sidxb = sidxb; tp->ux = ux; tp->sux = sux; tp->pi = pi; tp->spi = spi;
    tp->lam = lam; tp->t = t; tp->slam = slam;
}
/* dims_out: per stage nx,nu,nb,ng (4*(N+1) ints); returns {got, iter, status} through info[3] */
void usvref_tap_read(int *dims_out, int *info)
{
    qp_tap *tp = &g_tap;
    for (int k = 0; k <= tp->N; k++)
    {
        dims_out[4 * k + 0] = tp->nx[k]; dims_out[4 * k + 1] = tp->nu[k];
        dims_out[4 * k + 2] = tp->nb[k]; dims_out[4 * k + 3] = tp->ng[k];
    }
    info[0] = tp->got; info[1] = tp->iter; info[2] = tp->status;
    tp->enabled = 0;
}

static double now_s(void)
{
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return ts.tv_sec + 1e-9 * ts.tv_nsec;
}

/* Batch of B independent instances over `nthreads` host threads, one solver context per thread
 * (the reference is single-threaded per solver; this is "one solver instance per core",
 * BASELINE.md section 3).  Instance-major inputs: x0[B,nx], p[B,(N+1|1),np], lh[B,(N|1),K],
 * yref[B,(N|1),ny], yref_e[B,nye].  Outputs x[B,N+1,nx], u[B,N,nu], stats[B,9].
 * Returns the seconds the busiest thread spent inside the solves (context creation, and re-creation after a failed
 * solve, excluded). */
typedef struct
{
    void *ctx;
    /* creation arguments, kept to re-create the context after a failed solve (see batch_worker) */
    const int *c_icfg, *c_idxbx;
    const double *c_dcfg, *c_W, *c_We, *c_lbu, *c_ubu, *c_lbx, *c_ubx;
    int B, N, K, nx, nu, p_per_stage, lh_per_stage, yref_per_stage;
    long sp, slh, sy;
    const double *x0, *p, *lh, *yref, *yref_e;
    double *x_out, *u_out, *stats;
    int *next;
    double solve_s; /* seconds this thread spent inside usvref_solve (re-creation after a failure not counted) */
} batch_job;

static void *batch_worker(void *arg)
{
    batch_job *j = arg;
    for (;;)
    {
        int i = __atomic_fetch_add(j->next, 1, __ATOMIC_RELAXED);
        if (i >= j->B) break;
        const double ts0 = now_s();
        int st = usvref_solve(j->ctx, j->x0 + (long) i * j->nx, j->p + i * j->sp, j->p_per_stage, j->lh + i * j->slh,
                              j->lh_per_stage, j->yref + i * j->sy, j->yref_per_stage, j->yref_e + (long) i * j->nx, NULL,
                              NULL, NULL, j->x_out + (long) i * (j->N + 1) * j->nx, j->u_out + (long) i * j->N * j->nu,
                              NULL, NULL, NULL, j->stats + (long) i * 9);
        j->solve_s += now_s() - ts0;
        if (st != 0 && st != 2)
        {
            /* a QP failure (NaN / minimum step) leaves non-finite values in the reference's QP-solver memory and every
             * later solve with this context fails in its first QP: what a user of the reference has to do then is
             * to create the solver again, so that is what the harness does (observed: instance 414 of config 2). */
            usvref_free(j->ctx);
            j->ctx = usvref_create(j->c_icfg, j->c_dcfg, j->c_W, j->c_We, j->c_lbu, j->c_ubu, j->c_idxbx, j->c_lbx, j->c_ubx);
        }
    }
    return NULL;
}

double usvref_solve_batch(const int *icfg, const double *dcfg, const double *W, const double *We, const double *lbu,
                          const double *ubu, const int *idxbx, const double *lbx, const double *ubx, int B,
                          const double *x0, const double *p, int p_per_stage, const double *lh, int lh_per_stage,
                          const double *yref, int yref_per_stage, const double *yref_e, double *x_out, double *u_out,
                          double *stats, int nthreads)
{
    const int N = icfg[ICFG_N], K = icfg[ICFG_K];
    int nx, nu;
    usvm_dims(icfg[ICFG_MODEL], &nx, &nu);
    const int np = 2 * K, ny = nx + nu;
    if (nthreads < 1) nthreads = 1;
    if (nthreads > 256) nthreads = 256;
    batch_job jobs[256];
    pthread_t th[256];
    int next = 0;
    for (int t = 0; t < nthreads; t++)
    {
        batch_job *j = &jobs[t];
        j->ctx = usvref_create(icfg, dcfg, W, We, lbu, ubu, idxbx, lbx, ubx);
        j->c_icfg = icfg; j->c_dcfg = dcfg; j->c_W = W; j->c_We = We; j->c_lbu = lbu; j->c_ubu = ubu;
        j->c_idxbx = idxbx; j->c_lbx = lbx; j->c_ubx = ubx;
        j->B = B; j->N = N; j->K = K; j->nx = nx; j->nu = nu;
        j->p_per_stage = p_per_stage; j->lh_per_stage = lh_per_stage; j->yref_per_stage = yref_per_stage;
        j->sp = (long) (p_per_stage ? (N + 1) : 1) * np; j->slh = (long) (lh_per_stage ? N : 1) * K;
        j->sy = (long) (yref_per_stage ? N : 1) * ny;
        j->x0 = x0; j->p = p; j->lh = lh; j->yref = yref; j->yref_e = yref_e;
        j->x_out = x_out; j->u_out = u_out; j->stats = stats; j->next = &next; j->solve_s = 0.0;
    }
    double t0 = now_s();
    for (int t = 1; t < nthreads; t++) pthread_create(&th[t], NULL, batch_worker, &jobs[t]);
    batch_worker(&jobs[0]);
    for (int t = 1; t < nthreads; t++) pthread_join(th[t], NULL);
    double t1 = now_s(), busiest = 0.0;
    for (int t = 0; t < nthreads; t++) { usvref_free(jobs[t].ctx); if (jobs[t].solve_s > busiest) busiest = jobs[t].solve_s; }
    (void) t0; (void) t1;
    return busiest;
}
