/* acados_compat.h -- the EXACT symbols of the reference's generated solver library and of the acados_c calls the ROS
 * nodes make, for ONE solver instance (B = 1), exported by libusvmpc.so next to the batched usvmpc_* API.
 *
 * The reference generates `libacados_ocp_solver_<model>.so` from a Tera template
 * (interfaces/acados_template/acados_template/c_templates_tera/acados_solver.in.h:44-56, acados_solver.in.c:179-2037) and the
 * node nmpc_ca/src/nmpc_guidance_ca1.cpp:165,515-586 links it together with libacados:
 *     acados_create(); nlp_in = acados_get_nlp_in(); ...            (:165 and the globals :44-50)
 *     ocp_nlp_constraints_model_set(nlp_config, nlp_dims, nlp_in, 0, "lbx", x0);      (:515-516)
 *     ocp_nlp_cost_model_set(nlp_config, nlp_dims, nlp_in, ii, "yref", yref);         (:570)
 *     acados_update_params(ii, p_obs, 16);                                            (:571)
 *     ocp_nlp_constraints_model_set(nlp_config, nlp_dims, nlp_in, ii, "lh", r_obs);   (:572)
 *     acados_solve();                                                                 (:578)
 *     ocp_nlp_out_get(nlp_config, nlp_dims, nlp_out, 0, "u", u0);                     (:584)
 * Linking the node against libusvmpc.so instead resolves the same names to the B200 engine; the handles are opaque.
 * Like the generated library, this is global state: one solver per process, not re-entrant (acados_solver.in.c:101-107).
 * Which OCP acados_create() builds: the numbers the reference would have baked in at code generation -- by default
 * the deployed collision-avoidance solver (usv_model_guidance_ca1, N = 100, Tf = 5, 8 soft obstacle rows,
 * usv_guidance_ca1/acados_settings.py:70-208); usvmpc_acados_configure() selects another description first. */
#ifndef USVMPC_ACADOS_COMPAT_H_
#define USVMPC_ACADOS_COMPAT_H_

#include "usvmpc.h"

#ifdef __cplusplus
extern "C" {
#endif

/* extension: the description acados_create() will build (copied); NULL restores the default */
int usvmpc_acados_configure(const usvmpc_config* cfg);
/* fills cfg with the deployed CA solver's numbers (usv_guidance_ca1/acados_settings.py:70-208, main.py:54-55) */
int usvmpc_config_guidance_ca1(usvmpc_config* cfg);

/* acados_solver.in.h:44-56 */
int acados_create(void);
int acados_update_params(int stage, double* value, int np);
int acados_solve(void);
int acados_free(void);
void acados_print_stats(void);
void* acados_get_nlp_in(void);
void* acados_get_nlp_out(void);
void* acados_get_nlp_solver(void);
void* acados_get_nlp_config(void);
void* acados_get_nlp_opts(void);
void* acados_get_nlp_dims(void);
void* acados_get_nlp_plan(void);

/* acados_c/ocp_nlp_interface.h:154-390 (the calls the nodes and the Python wrapper make); handles as returned above */
int ocp_nlp_cost_model_set(void* config, void* dims, void* in, int stage, const char* field, void* value);
int ocp_nlp_constraints_model_set(void* config, void* dims, void* in, int stage, const char* field, void* value);
void ocp_nlp_out_set(void* config, void* dims, void* out, int stage, const char* field, void* value);
void ocp_nlp_out_get(void* config, void* dims, void* out, int stage, const char* field, void* value);
int ocp_nlp_dims_get_from_attr(void* config, void* dims, void* out, int stage, const char* field);
/* fields "sl", "su" (the slack values, as the Python wrapper reads them: acados_ocp_solver.py:774-782) */
void ocp_nlp_get_at_stage(void* config, void* dims, void* solver, int stage, const char* field, void* value);
/* fields: "sqp_iter" (int), "time_tot", "time_lin", "time_qp", "res_stat", "res_eq", "res_ineq", "res_comp", "cost_value" (double) */
void ocp_nlp_get(void* config, void* solver, const char* field, void* return_value);
void ocp_nlp_solver_opts_set(void* config, void* opts, const char* field, void* value);
/* residuals of the current iterate are part of every solve's statistics: nothing to do (ocp_nlp_interface.c:909) */
void ocp_nlp_eval_residuals(void* solver, void* in, void* out);
/* cost of the current iterate; read it with ocp_nlp_get(config, solver, "cost_value", &v)  (ocp_nlp_interface.c:919) */
void ocp_nlp_eval_cost(void* solver, void* in, void* out);
/* 2-D size queries of the Python wrapper's cost_set / constraints_set (acados_ocp_solver.py:1022-1030, 1090-1098):
 * vectors report (n, 0), "W" (ny, ny) */
void ocp_nlp_cost_dims_get_from_attr(void* config, void* dims, void* out, int stage, const char* field, int* dims_out);
void ocp_nlp_constraint_dims_get_from_attr(void* config, void* dims, void* out, int stage, const char* field, int* dims_out);

#ifdef __cplusplus
}
#endif
#endif
