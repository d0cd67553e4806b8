/* oracle/usv_models.h -- TEST INFRASTRUCTURE ONLY.
 *
 * Hand-derived model functions f(x,u) and Jacobians (df/dx, df/du) for the two models the
 * oracle knows.  They stand in for the CasADi-generated C (`<m>_expl_vde_forw`,
 * `<m>_constr_h_fun_jac_uxt_zt`) that the reference generates at run time
 * (/root/reference/.../acados_template/generate_c_code_explicit_ode.py:73-80,
 *  generate_c_code_constraint.py:98-109); CasADi itself is not available offline.
 *
 *  model 0  "usv3"     3-DOF surface vessel, x=[X,Y,psi,u,v,r], u=[Tport,Tstbd]
 *                      dynamics of NM/scripts/usv_position_control/usv_model.py:61-77,116-128
 *                      (thrust-rate states removed)
 *  model 1  "pendulum" cart-pole, x=[x1,theta,v1,dtheta], u=[F]
 *                      AC/examples/acados_python/getting_started/common/export_pendulum_ode_model.py:37-94
 *  model 2  "usv_model_guidance_ca1"  the guidance model of the deployed collision-avoidance node,
 *                      x=[u,v,ye,chie,psied,xned,yned,psi], u=[Upsieddot]
 *                      NM/scripts/usv_guidance_ca1/usv_model.py:65-128 (T1 = 1, beta = atan2(v, u + 0.001))
 *
 * CasADi differentiation conventions are mirrored: d|a|/da = sign(a) with sign(0)=0,
 * d(if_else(c,a,b))/dc = 0.
 * Shared by oracle/usv_oracle.c (the restatement) and oracle/ref_harness.c (callbacks for
 * the unmodified reference stack).  Matrices are column-major: J[i + n*j] = d f_i / d x_j.
 */
#ifndef USV_MODELS_H_
#define USV_MODELS_H_

#include <math.h>

#define USVM_MODEL_USV3 0
#define USVM_MODEL_PENDULUM 1
#define USVM_MODEL_GUIDANCE_CA1 2

static inline double usvm_sign(double a) { return (a > 0.0) - (a < 0.0); }

static inline void usvm_dims(int model, int *nx, int *nu)
{
    if (model == USVM_MODEL_PENDULUM) { *nx = 4; *nu = 1; }
    else if (model == USVM_MODEL_GUIDANCE_CA1) { *nx = 8; *nu = 1; }
    else { *nx = 6; *nu = 2; }
}

/* indices of the NED position states that enter the obstacle distance */
static inline void usvm_pos_states(int model, int *hx, int *hy)
{
    if (model == USVM_MODEL_GUIDANCE_CA1) { *hx = 5; *hy = 6; }
    else { *hx = 0; *hy = 1; }
}

/* f, Jx (6x6), Ju (6x2); Jx/Ju may be NULL */
static inline void usvm_usv3(const double *x, const double *uc, double *f, double *Jx, double *Ju)
{
    const double X_u_dot = -2.25, Y_v_dot = -23.13, Y_r_dot = -1.31, N_v_dot = -16.41, N_r_dot = -2.79;
    const double Yvv = -99.99, Yvr = -5.49, Nrv = -8.8, Nrr = -3.49;
    const double m = 30, Iz = 4.1, B = 0.41, c = 0.78;
    const double m11 = m - X_u_dot, m22 = m - Y_v_dot, m33 = Iz - N_r_dot;
    const double kY = 0.5 * (-40 * 1000) * (1.1 + 0.0045 * (1.01 / 0.09) - 0.1 * (0.27 / 0.09) + 0.016 * ((0.27 / 0.09) * (0.27 / 0.09)));

    const double psi = x[2], u = x[3], v = x[4], r = x[5];
    const double Tp = uc[0], Ts = uc[1];
    const double Xu = (u > 1.25) ? 64.55 : -25.0;
    const double Xuu = (u > 1.25) ? -70.92 : 0.0;
    const double s = sqrt(u * u + v * v);
    const double Yv = kY * fabs(v);
    const double Nr = -0.52 * s;
    const double Tu = Tp + c * Ts;
    const double Tr = (Tp - c * Ts) * B / 2;
    const double cp = cos(psi), sp = sin(psi);

    f[0] = u * cp - v * sp;
    f[1] = u * sp + v * cp;
    f[2] = r;
    f[3] = (Tu - (-m + 2 * Y_v_dot) * v - (Y_r_dot + N_v_dot) * r * r - (-Xu * u - Xuu * fabs(u) * u)) / m11;
    f[4] = (-(m - X_u_dot) * u * r - (-Yv - Yvv * fabs(v) - Yvr * fabs(r)) * v) / m22;
    f[5] = (Tr - (-2 * Y_v_dot * u * v - (Y_r_dot + N_v_dot) * r * u + X_u_dot * u * r) - (-Nr * r - Nrv * fabs(v) * r - Nrr * fabs(r) * r)) / m33;

    if (Jx)
    {
        for (int i = 0; i < 36; i++) Jx[i] = 0.0;
        Jx[0 + 6 * 2] = -u * sp - v * cp;  Jx[0 + 6 * 3] = cp;  Jx[0 + 6 * 4] = -sp;
        Jx[1 + 6 * 2] = u * cp - v * sp;   Jx[1 + 6 * 3] = sp;  Jx[1 + 6 * 4] = cp;
        Jx[2 + 6 * 5] = 1.0;
        Jx[3 + 6 * 3] = (Xu + 2 * Xuu * fabs(u)) / m11;
        Jx[3 + 6 * 4] = -(-m + 2 * Y_v_dot) / m11;
        Jx[3 + 6 * 5] = -2 * (Y_r_dot + N_v_dot) * r / m11;
        Jx[4 + 6 * 3] = -m11 * r / m22;
        Jx[4 + 6 * 4] = (2 * (kY + Yvv) * fabs(v) + Yvr * fabs(r)) / m22;
        Jx[4 + 6 * 5] = (-m11 * u + Yvr * usvm_sign(r) * v) / m22;
        Jx[5 + 6 * 3] = (2 * Y_v_dot * v + (Y_r_dot + N_v_dot) * r - X_u_dot * r - 0.52 * (u / s) * r) / m33;
        Jx[5 + 6 * 4] = (2 * Y_v_dot * u - 0.52 * (v / s) * r + Nrv * usvm_sign(v) * r) / m33;
        Jx[5 + 6 * 5] = ((Y_r_dot + N_v_dot) * u - X_u_dot * u - 0.52 * s + Nrv * fabs(v) + 2 * Nrr * fabs(r)) / m33;
    }
    if (Ju)
    {
        for (int i = 0; i < 12; i++) Ju[i] = 0.0;
        Ju[3 + 6 * 0] = 1.0 / m11;       Ju[3 + 6 * 1] = c / m11;
        Ju[5 + 6 * 0] = (B / 2) / m33;   Ju[5 + 6 * 1] = -(c * B / 2) / m33;
    }
}

/* f, Jx (4x4), Ju (4x1) */
static inline void usvm_pendulum(const double *x, const double *uc, double *f, double *Jx, double *Ju)
{
    const double M = 1.0, m = 0.1, g = 9.81, l = 0.8;
    const double th = x[1], v1 = x[2], dth = x[3], F = uc[0];
    const double c = cos(th), s = sin(th);
    const double den = M + m - m * c * c;
    const double n3 = -m * l * s * dth * dth + m * g * c * s + F;
    const double n4 = -m * l * c * s * dth * dth + F * c + (M + m) * g * s;
    f[0] = v1;
    f[1] = dth;
    f[2] = n3 / den;
    f[3] = n4 / (l * den);
    if (Jx)
    {
        const double dden = 2 * m * c * s;
        const double dn3_th = -m * l * c * dth * dth + m * g * (c * c - s * s);
        const double dn4_th = -m * l * (c * c - s * s) * dth * dth - F * s + (M + m) * g * c;
        for (int i = 0; i < 16; i++) Jx[i] = 0.0;
        Jx[0 + 4 * 2] = 1.0;
        Jx[1 + 4 * 3] = 1.0;
        Jx[2 + 4 * 1] = (dn3_th * den - n3 * dden) / (den * den);
        Jx[2 + 4 * 3] = -2 * m * l * s * dth / den;
        Jx[3 + 4 * 1] = (dn4_th * den - n4 * dden) / (l * den * den);
        Jx[3 + 4 * 3] = -2 * m * l * c * s * dth / (l * den);
    }
    if (Ju)
    {
        Ju[0] = 0.0; Ju[1] = 0.0; Ju[2] = 1.0 / den; Ju[3] = c / (l * den);
    }
}

/* f, Jx (8x8), Ju (8x1): NM/scripts/usv_guidance_ca1/usv_model.py:117-128 */
static inline void usvm_guidance_ca1(const double *x, const double *uc, double *f, double *Jx, double *Ju)
{
    const double T1 = 1.0;
    const double u = x[0], v = x[1], chie = x[3], psied = x[4], psi = x[7];
    const double ue = u + 0.001;
    const double beta = atan2(v, ue);
    const double psie = chie - beta;
    const double se = sin(psie), ce = cos(psie), sp = sin(psi), cp = cos(psi);
    f[0] = 0.0;
    f[1] = 0.0;
    f[2] = u * se + v * ce;
    f[3] = (psied - psie) / T1;
    f[4] = uc[0];
    f[5] = u * cp - v * sp;
    f[6] = u * sp + v * cp;
    f[7] = (psied - psie) / T1;
    if (Jx)
    {
        const double den = ue * ue + v * v;
        const double db_du = -v / den, db_dv = ue / den;     /* d atan2(v, u + 0.001) */
        const double w = u * ce - v * se;                    /* d f2 / d psie */
        for (int i = 0; i < 64; i++) Jx[i] = 0.0;
        Jx[2 + 8 * 0] = se - w * db_du;
        Jx[2 + 8 * 1] = ce - w * db_dv;
        Jx[2 + 8 * 3] = w;
        Jx[3 + 8 * 0] = db_du / T1;
        Jx[3 + 8 * 1] = db_dv / T1;
        Jx[3 + 8 * 3] = -1.0 / T1;
        Jx[3 + 8 * 4] = 1.0 / T1;
        Jx[5 + 8 * 0] = cp;  Jx[5 + 8 * 1] = -sp;  Jx[5 + 8 * 7] = -u * sp - v * cp;
        Jx[6 + 8 * 0] = sp;  Jx[6 + 8 * 1] = cp;   Jx[6 + 8 * 7] = u * cp - v * sp;
        Jx[7 + 8 * 0] = db_du / T1;
        Jx[7 + 8 * 1] = db_dv / T1;
        Jx[7 + 8 * 3] = -1.0 / T1;
        Jx[7 + 8 * 4] = 1.0 / T1;
    }
    if (Ju)
    {
        for (int i = 0; i < 8; i++) Ju[i] = 0.0;
        Ju[4] = 1.0;
    }
}

static inline void usvm_f_jac(int model, const double *x, const double *uc, double *f, double *Jx, double *Ju)
{
    if (model == USVM_MODEL_PENDULUM) usvm_pendulum(x, uc, f, Jx, Ju);
    else if (model == USVM_MODEL_GUIDANCE_CA1) usvm_guidance_ca1(x, uc, f, Jx, Ju);
    else usvm_usv3(x, uc, f, Jx, Ju);
}

/* obstacle distances h_i = ||(X,Y) - (ox_i, oy_i)||, i<K, p=[ox_1,oy_1,...]
 * (NM/scripts/usv_pf_ca/usv_model.py:165-168).  dh/dX, dh/dY returned in gX,gY (may be NULL). */
static inline void usvm_obstacle_h_at(int hx, int hy, int K, const double *x, const double *p, double *h, double *gX, double *gY)
{
    for (int i = 0; i < K; i++)
    {
        const double dx = x[hx] - p[2 * i], dy = x[hy] - p[2 * i + 1];
        const double d = sqrt(dx * dx + dy * dy);
        h[i] = d;
        if (gX) { gX[i] = dx / d; gY[i] = dy / d; }
    }
}
static inline void usvm_obstacle_h(int K, const double *x, const double *p, double *h, double *gX, double *gY)
{
    for (int i = 0; i < K; i++)
    {
        const double dx = x[0] - p[2 * i], dy = x[1] - p[2 * i + 1];
        const double d = sqrt(dx * dx + dy * dy);
        h[i] = d;
        if (gX) { gX[i] = dx / d; gY[i] = dy / d; }
    }
}

#endif
