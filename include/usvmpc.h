/* usvmpc.h -- C ABI of the B200 batched NMPC solve engine (libusvmpc.so).
 *
 * Drop-in boundary: these entry points are the batched equivalents of what the reference binds over
 * ctypes / links from its ROS nodes for ONE solver instance, i.e. the generated-solver API
 * (interfaces/acados_template/acados_template/c_templates_tera/acados_solver.in.h:44-56) plus the
 * acados_c NLP interface (interfaces/acados_c/ocp_nlp_interface.h:154-390).  One usvmpc_solver
 * holds B independent NMPC instances on one GPU; every per-instance value gains a leading batch
 * dimension, everything else (field names, stage indexing, u-before-x ordering, multiplier
 * ordering [lbu lbx lh | ubu ubx uh], column-major matrices, status codes) is the reference's.
 *
 * Conventions: plain pointers and sizes only.  `value` buffers are caller-owned and are copied
 * before the call returns control of them (host buffers) or ordered on `stream` (device buffers);
 * `on_device` says where `value` lives; `stream` is a cudaStream_t (NULL = default stream).
 * All calls return 0 on success, a negative USVMPC_E_* code otherwise; usvmpc_last_error() gives
 * the message.  Not thread-safe per solver (the reference's generated library is global state,
 * acados_solver.in.c:101-107); different solvers may be driven from different threads.
 *
 * Paths below are relative to catkin_ws/src/nmpc_ca/acados/ of the reference.
 */
#ifndef USVMPC_H_
#define USVMPC_H_

#ifdef __cplusplus
extern "C" {
#endif

#define USVMPC_ALL_STAGES (-1)   /* value is [B][n_stages][dim], one row per stage            */
#define USVMPC_EVERY_STAGE (-2)  /* value is [B][dim], the same row applies to every stage    */

#define USVMPC_MODEL_USV3 0      /* 3-DOF USV, nx=6 nu=2, h = obstacle distances  */
#define USVMPC_MODEL_PENDULUM 1  /* cart-pole of the reference's golden-vector tests          */
#define USVMPC_MODEL_GUIDANCE_CA1 2 /* the deployed CA guidance model "usv_model_guidance_ca1", nx=8 nu=1
                                       (nmpc_ca/scripts/usv_guidance_ca1/usv_model.py:65-128)       */

#define USVMPC_SQP 0
#define USVMPC_SQP_RTI 1

/* solve status per instance = acados `enum return_values` (acados/utils/types.h:59-67) */
#define USVMPC_SUCCESS 0
#define USVMPC_FAILURE 1
#define USVMPC_MAXITER 2
#define USVMPC_MINSTEP 3
#define USVMPC_QP_FAILURE 4

#define USVMPC_E_INVALID (-1)
#define USVMPC_E_CUDA (-2)
#define USVMPC_E_FIELD (-3)
#define USVMPC_E_SIZE (-4)

#define USVMPC_NSTAT 16
/* statistics record per instance (doubles):
 *  0 status  1 sqp_iter  2 qp_iter(total)  3 res_stat  4 res_eq  5 res_ineq  6 res_comp  7 -
 *  8 solve-only Riccati sweeps  9 last QP status (HPIPM)  10 last QP iterations  11 -          */

typedef struct usvmpc_solver usvmpc_solver;

/* The numbers the reference bakes into acados_create() from the AcadosOcp description
 * (acados_solver.in.c:179-1739): dimensions, LINEAR_LS weights, bounds, solver options. */
typedef struct usvmpc_config
{
    int model;                 /* USVMPC_MODEL_*                                                    */
    int N;                     /* dims.N                                                            */
    int K;                     /* number of obstacle rows h (dims.nh), np = 2K                      */
    int num_steps, num_stages; /* sim_method_num_steps / sim_method_num_stages (ERK)                */
    int nlp_type;              /* USVMPC_SQP | USVMPC_SQP_RTI (nlp_solver_type)                     */
    int max_iter;              /* nlp_solver_max_iter                                               */
    int qp_iter_max;           /* qp_solver_iter_max                                                */
    int nbx, nbu;              /* path state boxes (stages 1..N-1) / input boxes on u[0..nbu)       */
    int idxbx[8];
    double dt;                 /* tf / N (shooting interval = LS cost scaling, tpl :806-810)        */
    double tol[4];             /* nlp_solver_tol_stat/eq/ineq/comp (forwarded to the QP in SQP)     */
    double uh;                 /* upper bound of every h row (the scripts use 1e6, not masked)      */
    double lbu[4], ubu[4], lbx[8], ubx[8];
    double W[16 * 16];         /* ny x ny, column-major, y = [x; u] (Vx=[I;0], Vu=[0;I])            */
    double W_e[16 * 16];       /* nx x nx, column-major                                             */
    /* soft obstacle rows (dims.nsh, constraints.idxsh = 0..nsh-1: the first nsh rows of h), each with a lower and an
     * upper slack variable: slack bounds lsh / ush, linear and quadratic penalties zl, zu, Zl, Zu
     * (usv_guidance_ca1/acados_settings.py:105-178; acados_solver.in.c:880-930,1395-1449) */
    int nsh;
    double lsh[32], ush[32], zl[32], zu[32], Zl[32], Zu[32];
} usvmpc_config;

const char* usvmpc_last_error(void);
const char* usvmpc_version(void);

/* fill cfg with the AcadosOcpOptions defaults (acados_template/acados_ocp.py:1747-1780) for a model */
int usvmpc_config_default(usvmpc_config* cfg, int model);

/* replaces acados_create() / acados_free()  (acados_solver.in.h:44,50) */
int usvmpc_create(const usvmpc_config* cfg, int batch, int device, usvmpc_solver** out);
int usvmpc_free(usvmpc_solver* s);

/* replaces acados_solve() -> ocp_nlp_solve() (acados_solver.in.c:1841, ocp_nlp_interface.c:884).
 * Launches the solve of all B instances on `stream` and returns; per-instance status is in the
 * statistics record.  */
int usvmpc_solve(usvmpc_solver* s, void* stream);

/* replaces acados_update_params(stage, p, np) (acados_solver.in.c:1742): p = [ox_1, oy_1, ...], value [B][np] */
int usvmpc_update_params(usvmpc_solver* s, int stage, const double* value, int np, int on_device, void* stream);

/* replaces ocp_nlp_cost_model_set (ocp_nlp_interface.c:402): "yref"/"y_ref" value [B][ny] (stage N: [B][nx]);
 * "W" value [ny*ny] column-major, shared by the batch (stage N: W_e [nx*nx]); "zl", "zu", "Zl", "Zu" value [nsh],
 * shared by the batch and the stages */
int usvmpc_cost_model_set(usvmpc_solver* s, int stage, const char* field, const double* value, int on_device,
                          void* stream);

/* replaces ocp_nlp_constraints_model_set (ocp_nlp_interface.c:413; field map ocp_nlp_constraints_bgh.c:630-822):
 * stage 0 "lbx"/"ubx" = x0, value [B][nx]; "lh" value [B][K]; "lbu","ubu","uh" and "lbx","ubx" of stages >= 1 are
 * shared by the batch (value [nbu] / [K] / [nbx], host memory); so are "lsh", "ush" (value [nsh]) */
int usvmpc_constraints_model_set(usvmpc_solver* s, int stage, const char* field, const double* value,
                                 int on_device, void* stream);

/* replace ocp_nlp_out_set / ocp_nlp_out_get (ocp_nlp_interface.c:452,489): fields x u pi lam t sl su (z: empty).
 * lam / t of a stage: [lbu lbx lh | ubu ubx uh | lsh | ush] with the stage's own counts (acados_ocp_solver.py:732-735) */
int usvmpc_out_set(usvmpc_solver* s, int stage, const char* field, const double* value, int on_device, void* stream);
int usvmpc_out_get(usvmpc_solver* s, int stage, const char* field, double* value, int on_device, void* stream);

/* replaces ocp_nlp_dims_get_from_attr (ocp_nlp_interface.c:536): per-instance length of `field` at `stage`, <0 if unknown */
int usvmpc_dims_get_from_attr(usvmpc_solver* s, int stage, const char* field);

/* replaces ocp_nlp_get(..."sqp_iter"|"res_*"|"statistics"...) (ocp_nlp_interface.c:935): value [B][USVMPC_NSTAT] */
int usvmpc_get_stats(usvmpc_solver* s, double* value, int on_device, void* stream);

/* replaces ocp_nlp_eval_cost + ocp_nlp_get(..."cost_value"...) (ocp_nlp_interface.c:919,935): LINEAR_LS cost of the
 * current iterate, value [B] */
int usvmpc_eval_cost(usvmpc_solver* s, double* value, int on_device, void* stream);

/* replaces ocp_nlp_solver_opts_set (ocp_nlp_interface.c:943).  Fields: "max_iter", "qp_iter_max", "tol_stat",
 * "tol_eq", "tol_ineq", "tol_comp", "nlp_solver_type" (0/1), "cold_start" (extension: 1 = every solve starts from
 * x_k = x0, u = 0, pi = 0 instead of the previous iterate), "lpt_schedule" (extension, default 1: the work queue of a
 * solve is ordered by decreasing iteration count of the previous solve of this solver), "riccati_precision" (extension: 32 = Riccati factorisation in fp32 with an fp64 accuracy test and fallback, BASELINE
 * config 4; 64 = default), "print_level", "rti_phase" (0 preparation + feedback, 1 preparation, 2 feedback:
 * ocp_nlp_sqp_rti.c:459-488), "step_length" (1 only) */
int usvmpc_solver_opts_set(usvmpc_solver* s, const char* field, double value);

/* Obstacle front end of the guidance node (nmpc_ca/src/nmpc_guidance_ca1.cpp:251-363, obstaclesCallback + body2NED):
 * device pointers only.  pose [B][3] = (nedx, nedy, psi); obs_body [B][max_obs][3] = (x, y, radius) in the body frame;
 * len [B] = number of valid obstacles per instance.  Writes p_out [B][2K] = (ox_1, oy_1, ...) in NED -- the argument of
 * usvmpc_update_params -- and r_out [B][K] = radius + boat_radius -- the argument of constraints "lh".  The K obstacles
 * with the smallest clearance are kept; unused slots hold (init_obs_pos, init_obs_pos, 0). */
int usvmpc_obstacle_frontend(const double* pose, const double* obs_body, const int* len, int batch, int max_obs, int K,
                             double boat_radius, double init_obs_pos, double* p_out, double* r_out, void* stream);

/* The QP seam of the reference, qp_solver_config.evaluate(qp_in, qp_out) (acados/ocp_qp/ocp_qp_common.h:62-76 ->
 * ocp_qp_hpipm.c:240 -> d_ocp_qp_ipm_solve): solve ONE given OCP QP per instance -- in the form HPIPM sees it, i.e. after
 * x0 elimination -- with the engine's IPM, tolerances / iteration limit as configured.  Device pointers, engine layout:
 *   G   [B][N][nv*nx]   [B';A'] of every stage, column-major nv x nx (stage 0: x rows zero)
 *   b   [B][N][nx]      rq [B][N+1][nv]      gxy [B][N][2K] = (dh_c/dX, c<K | dh_c/dY, c<K)
 *   d   [B][N][2*ncq]   rows [u boxes | x boxes | h rows] lower side then upper side (upper side negated like HPIPM's d)
 * out: ux [B][N+1][nv], pi [B][N][nx], lam, t [B][N][2*ncq]; the statistics record holds HPIPM status (slot 0), iterations
 * (slot 2) and the four residual norms.  The Hessian is the solver's Gauss-Newton Hessian (LINEAR_LS: constant). */
int usvmpc_qp_solve(usvmpc_solver* s, const double* G, const double* b, const double* rq, const double* gxy, const double* d,
                    double* ux, double* pi, double* lam, double* t, void* stream);

/* Multi-GPU result exchange: give the solver a device buffer [B][width] and the solve kernel's
 * epilogue writes every instance's packed result row into it -- x [N+1][nx] | u [N][nu] | status, sqp_iter, qp_iter,
 * res_stat, res_eq, res_ineq, res_comp -- so that the buffer can be handed to ncclAllGather as it is (it may be this
 * rank's slice of the gathered tensor: in-place all-gather).  NULL switches the epilogue off.  Returns width. */
int usvmpc_set_result_buffer(usvmpc_solver* s, double* device_buffer);

/* engine introspection for benchmarks: kernels launched so far, bytes of HBM held, batch, launch geometry */
int usvmpc_info(usvmpc_solver* s, const char* what, double* value);

#ifdef __cplusplus
}
#endif
#endif
