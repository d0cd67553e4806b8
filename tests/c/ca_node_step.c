/* ca_node_step.c -- link test of include/acados_compat.h: ONE control step of the deployed collision-avoidance node, written
 * with the generated solver's own symbol names in the order the node calls them
 * (nmpc_ca/src/nmpc_guidance_ca1.cpp:165 acados_create; :515-516 x0; :570-572 yref, p, lh per stage; :578 acados_solve;
 *  :584 ocp_nlp_out_get "u").  Plain C, no CUDA or torch headers: what a maintainer's node sees after relinking against
 * libusvmpc.so.  Input: a binary file of doubles [x0 8 | p 16 | lh 8 | yref 9 | yref_e 8]; output on stdout:
 * status, sqp_iter, then u[0..N-1] and x[0..N][8], one number per line with 17 significant digits. */
#include <stdio.h>
#include <stdlib.h>

#include "acados_compat.h"

#define NX 8
#define NU 1
#define N 100
#define NP 16
#define NH 8

int main(int argc, char** argv)
{
    double in[NX + NP + NH + NX + NU + NX];
    if (argc < 2) return 2;
    FILE* f = fopen(argv[1], "rb");
    if (!f || fread(in, sizeof(double), sizeof(in) / sizeof(double), f) != sizeof(in) / sizeof(double)) return 2;
    fclose(f);
    double *x0 = in, *p = x0 + NX, *lh = p + NP, *yref = lh + NH, *yref_e = yref + NX + NU;

    if (acados_create() != 0) { fprintf(stderr, "acados_create failed\n"); return 1; }
    void* nlp_config = acados_get_nlp_config();
    void* nlp_dims = acados_get_nlp_dims();
    void* nlp_in = acados_get_nlp_in();
    void* nlp_out = acados_get_nlp_out();
    void* nlp_solver = acados_get_nlp_solver();
    if (ocp_nlp_dims_get_from_attr(nlp_config, nlp_dims, nlp_out, 0, "x") != NX) return 3;
    if (ocp_nlp_dims_get_from_attr(nlp_config, nlp_dims, nlp_out, 0, "u") != NU) return 3;

    /* initial guess as the test harness of the reference uses it: the measured state at every node */
    for (int k = 0; k <= N; k++) ocp_nlp_out_set(nlp_config, nlp_dims, nlp_out, k, "x", x0);

    ocp_nlp_constraints_model_set(nlp_config, nlp_dims, nlp_in, 0, "lbx", x0);
    ocp_nlp_constraints_model_set(nlp_config, nlp_dims, nlp_in, 0, "ubx", x0);
    for (int k = 0; k < N; k++)
    {
        ocp_nlp_cost_model_set(nlp_config, nlp_dims, nlp_in, k, "yref", yref);
        acados_update_params(k, p, NP);
        ocp_nlp_constraints_model_set(nlp_config, nlp_dims, nlp_in, k, "lh", lh);
    }
    ocp_nlp_cost_model_set(nlp_config, nlp_dims, nlp_in, N, "yref", yref_e);
    acados_update_params(N, p, NP);

    int status = acados_solve();
    int sqp_iter = -1;
    double time_tot = 0.0;
    ocp_nlp_get(nlp_config, nlp_solver, "sqp_iter", &sqp_iter);
    ocp_nlp_get(nlp_config, nlp_solver, "time_tot", &time_tot);
    printf("%d\n%d\n", status, sqp_iter);
    for (int k = 0; k < N; k++)
    {
        double u;
        ocp_nlp_out_get(nlp_config, nlp_dims, nlp_out, k, "u", &u);
        printf("%.17g\n", u);
    }
    for (int k = 0; k <= N; k++)
    {
        double x[NX];
        ocp_nlp_out_get(nlp_config, nlp_dims, nlp_out, k, "x", x);
        for (int i = 0; i < NX; i++) printf("%.17g\n", x[i]);
    }
    /* the slack of the first soft row at stage 1, read the way the Python wrapper does */
    double sl[NH];
    ocp_nlp_get_at_stage(nlp_config, nlp_dims, nlp_solver, 1, "sl", sl);
    fprintf(stderr, "time_tot %.6f s, sl[0] at stage 1 %.3e\n", time_tot, sl[0]);
    if (!(time_tot > 0.0)) return 4;
    return acados_free();
}
