// models.cuh -- device model functions: continuous dynamics f(x,u) with Jacobians, replacing the
// CasADi-generated C the reference compiles at run time (`<m>_expl_vde_forw`,
// `<m>_constr_h_fun_jac_uxt_zt`; spec in
// acados_template/generate_c_code_explicit_ode.py:73-80 and generate_c_code_constraint.py:98-109).
// CasADi's differentiation conventions are kept: d|a|/da = sign(a) (0 at 0), d(if_else)/dcond = 0.
//
//  Usv3      3-DOF surface vessel x=[X,Y,psi,u,v,r], u=[Tport,Tstbd]; dynamics of
//            NM/scripts/usv_position_control/usv_model.py:61-77,116-128 with the thrust-rate states
//            removed; obstacle distance h_i = ||(X,Y)-(ox_i,oy_i)|| of
//            NM/scripts/usv_pf_ca/usv_model.py:165-168.
//  Usv8Ca1   guidance model of the deployed collision-avoidance node, x=[u,v,ye,chie,psied,xned,yned,psi],
//            u=[Upsieddot]: NM/scripts/usv_guidance_ca1/usv_model.py:65-128 (T1 = 1, beta = atan2(v, u + 0.001)); the
//            obstacle distances use (xned, yned).
//  Pendulum  cart-pole of the reference's own golden-vector tests
//            (AC/examples/acados_python/getting_started/common/export_pendulum_ode_model.py:37-94).
//
// Jacobians are returned dense, column-major: Jx[i + NX*j] = d f_i / d x_j.
#pragma once
#include "cta_compat.h"

namespace usvmpc {

DEV double dsign(double a) { return (double) ((a > 0.0) - (a < 0.0)); }

struct Usv3 {
    static constexpr int ID = 0;
    static constexpr int NX = 6, NU = 2;
    static constexpr int HX = 0, HY = 1;  // position states entering the obstacle distance
    static constexpr int NKIN = 2;        // leading states (X, Y) the dynamics do not depend on: their sensitivity columns stay unit vectors

    MDEV static void f_jac(const double* x, const double* uc, double* f, double* Jx, double* Ju)
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
        const double s = dsqrt(u * u + v * v);
        const double Yv = kY * dabs(v);
        const double Nr = -0.52 * s;
        const double Tu = Tp + c * Ts;
        const double Tr = (Tp - c * Ts) * B / 2;
        double sp, cp;
        dsincos(psi, &sp, &cp);

        f[0] = u * cp - v * sp;
        f[1] = u * sp + v * cp;
        f[2] = r;
        f[3] = (Tu - (-m + 2 * Y_v_dot) * v - (Y_r_dot + N_v_dot) * r * r - (-Xu * u - Xuu * dabs(u) * u)) / m11;
        f[4] = (-(m - X_u_dot) * u * r - (-Yv - Yvv * dabs(v) - Yvr * dabs(r)) * v) / m22;
        f[5] = (Tr - (-2 * Y_v_dot * u * v - (Y_r_dot + N_v_dot) * r * u + X_u_dot * u * r) - (-Nr * r - Nrv * dabs(v) * r - Nrr * dabs(r) * r)) / m33;

#pragma unroll
        for (int i = 0; i < 36; i++) Jx[i] = 0.0;
        Jx[0 + 6 * 2] = -u * sp - v * cp;  Jx[0 + 6 * 3] = cp;  Jx[0 + 6 * 4] = -sp;
        Jx[1 + 6 * 2] = u * cp - v * sp;   Jx[1 + 6 * 3] = sp;  Jx[1 + 6 * 4] = cp;
        Jx[2 + 6 * 5] = 1.0;
        Jx[3 + 6 * 3] = (Xu + 2 * Xuu * dabs(u)) / m11;
        Jx[3 + 6 * 4] = -(-m + 2 * Y_v_dot) / m11;
        Jx[3 + 6 * 5] = -2 * (Y_r_dot + N_v_dot) * r / m11;
        Jx[4 + 6 * 3] = -m11 * r / m22;
        Jx[4 + 6 * 4] = (2 * (kY + Yvv) * dabs(v) + Yvr * dabs(r)) / m22;
        Jx[4 + 6 * 5] = (-m11 * u + Yvr * dsign(r) * v) / m22;
        Jx[5 + 6 * 3] = (2 * Y_v_dot * v + (Y_r_dot + N_v_dot) * r - X_u_dot * r - 0.52 * (u / s) * r) / m33;
        Jx[5 + 6 * 4] = (2 * Y_v_dot * u - 0.52 * (v / s) * r + Nrv * dsign(v) * r) / m33;
        Jx[5 + 6 * 5] = ((Y_r_dot + N_v_dot) * u - X_u_dot * u - 0.52 * s + Nrv * dabs(v) + 2 * Nrr * dabs(r)) / m33;
#pragma unroll
        for (int i = 0; i < 12; i++) Ju[i] = 0.0;
        Ju[3 + 6 * 0] = 1.0 / m11;       Ju[3 + 6 * 1] = c / m11;
        Ju[5 + 6 * 0] = (B / 2) / m33;   Ju[5 + 6 * 1] = -(c * B / 2) / m33;
    }

    // one column of the variational equation: f(x,u) and ks = (df/dx) s (+ (df/du) e_ucol if ucol >= 0), using the
    // sparsity of the Jacobian (16 of 36 entries).  Same terms in the same order as the dense product, so the result
    // equals the CasADi-generated `<m>_expl_vde_forw` column (generate_c_code_explicit_ode.py:73-80).
    MDEV static void vde_col(const double* x, const double* uc, const double* s, int ucol, double* f, double* ks)
    {
        const double X_u_dot = -2.25, Y_v_dot = -23.13, Y_r_dot = -1.31, N_v_dot = -16.41, N_r_dot = -2.79;
        const double Yvv = -99.99, Yvr = -5.49, Nrv = -8.8, Nrr = -3.49;
        const double m = 30, Iz = 4.1, B = 0.41, c = 0.78;
        const double m11 = m - X_u_dot, m22 = m - Y_v_dot, m33 = Iz - N_r_dot;
        // the inertia terms are compile-time constants: multiply by their reciprocals (folded by the compiler) instead of
        // twelve fp64 divisions per evaluation; differs from the division by at most one ulp per term
        const double i11 = 1.0 / m11, i22 = 1.0 / m22, i33 = 1.0 / m33;
        const double kY = 0.5 * (-40 * 1000) * (1.1 + 0.0045 * (1.01 / 0.09) - 0.1 * (0.27 / 0.09) + 0.016 * ((0.27 / 0.09) * (0.27 / 0.09)));
        const double psi = x[2], u = x[3], v = x[4], r = x[5];
        const double Tp = uc[0], Ts = uc[1];
        const double Xu = (u > 1.25) ? 64.55 : -25.0;
        const double Xuu = (u > 1.25) ? -70.92 : 0.0;
        const double sq = dsqrt(u * u + v * v), isq = 1.0 / sq;
        const double Yv = kY * dabs(v);
        const double Nr = -0.52 * sq;
        const double Tu = Tp + c * Ts;
        const double Tr = (Tp - c * Ts) * B / 2;
        double sp, cp;
        dsincos(psi, &sp, &cp);
        f[0] = u * cp - v * sp;
        f[1] = u * sp + v * cp;
        f[2] = r;
        f[3] = (Tu - (-m + 2 * Y_v_dot) * v - (Y_r_dot + N_v_dot) * r * r - (-Xu * u - Xuu * dabs(u) * u)) * i11;
        f[4] = (-(m - X_u_dot) * u * r - (-Yv - Yvv * dabs(v) - Yvr * dabs(r)) * v) * i22;
        f[5] = (Tr - (-2 * Y_v_dot * u * v - (Y_r_dot + N_v_dot) * r * u + X_u_dot * u * r) - (-Nr * r - Nrv * dabs(v) * r - Nrr * dabs(r) * r)) * i33;
        const double J02 = -u * sp - v * cp, J12 = u * cp - v * sp;
        const double J33 = (Xu + 2 * Xuu * dabs(u)) * i11, J34 = -(-m + 2 * Y_v_dot) / m11, J35 = -2 * (Y_r_dot + N_v_dot) * r * i11;
        const double J43 = -m11 * r * i22, J44 = (2 * (kY + Yvv) * dabs(v) + Yvr * dabs(r)) * i22, J45 = (-m11 * u + Yvr * dsign(r) * v) * i22;
        const double J53 = (2 * Y_v_dot * v + (Y_r_dot + N_v_dot) * r - X_u_dot * r - 0.52 * (u * isq) * r) * i33;
        const double J54 = (2 * Y_v_dot * u - 0.52 * (v * isq) * r + Nrv * dsign(v) * r) * i33;
        const double J55 = ((Y_r_dot + N_v_dot) * u - X_u_dot * u - 0.52 * sq + Nrv * dabs(v) + 2 * Nrr * dabs(r)) * i33;
        const double b3 = ucol == 0 ? 1.0 / m11 : (ucol == 1 ? c / m11 : 0.0);
        const double b5 = ucol == 0 ? (B / 2) / m33 : (ucol == 1 ? -(c * B / 2) / m33 : 0.0);
        ks[0] = ((J02 * s[2]) + cp * s[3]) + (-sp) * s[4];
        ks[1] = ((J12 * s[2]) + sp * s[3]) + cp * s[4];
        ks[2] = s[5];
        ks[3] = ((b3 + J33 * s[3]) + J34 * s[4]) + J35 * s[5];
        ks[4] = ((J43 * s[3]) + J44 * s[4]) + J45 * s[5];
        ks[5] = ((b5 + J53 * s[3]) + J54 * s[4]) + J55 * s[5];
    }
};

DEV double datan2(double y, double x) { return atan2(y, x); }

struct Usv8Ca1 {
    static constexpr int ID = 2;
    static constexpr int NX = 8, NU = 1;
    static constexpr int HX = 5, HY = 6;  // NED position states entering the obstacle distance
    static constexpr int NKIN = 0;        // (states that do not enter the dynamics are not leading here: integrate every column)

    MDEV static void f_jac(const double* x, const double* uc, double* f, double* Jx, double* Ju)
    {
        const double T1 = 1.0;
        const double u = x[0], v = x[1], chie = x[3], psied = x[4], psi = x[7];
        const double ue = u + 0.001;
        const double beta = datan2(v, ue);
        const double psie = chie - beta;
        double se, ce, sp, cp;
        dsincos(psie, &se, &ce);
        dsincos(psi, &sp, &cp);
        f[0] = 0.0;
        f[1] = 0.0;
        f[2] = u * se + v * ce;
        f[3] = (psied - psie) / T1;
        f[4] = uc[0];
        f[5] = u * cp - v * sp;
        f[6] = u * sp + v * cp;
        f[7] = (psied - psie) / T1;
        const double den = ue * ue + v * v;
        const double db_du = -v / den, db_dv = ue / den;  // d atan2(v, u + 0.001)
        const double w = u * ce - v * se;                 // d f2 / d psie
#pragma unroll
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
#pragma unroll
        for (int i = 0; i < 8; i++) Ju[i] = 0.0;
        Ju[4] = 1.0;
    }

    // one column of the variational equation (see Usv3::vde_col), using the sparsity of the Jacobian
    MDEV static void vde_col(const double* x, const double* uc, const double* s, int ucol, double* f, double* ks)
    {
        const double T1 = 1.0;
        const double u = x[0], v = x[1], chie = x[3], psied = x[4], psi = x[7];
        const double ue = u + 0.001;
        const double beta = datan2(v, ue);
        const double psie = chie - beta;
        double se, ce, sp, cp;
        dsincos(psie, &se, &ce);
        dsincos(psi, &sp, &cp);
        f[0] = 0.0;
        f[1] = 0.0;
        f[2] = u * se + v * ce;
        f[3] = (psied - psie) / T1;
        f[4] = uc[0];
        f[5] = u * cp - v * sp;
        f[6] = u * sp + v * cp;
        f[7] = (psied - psie) / T1;
        const double den = ue * ue + v * v;
        const double db_du = -v / den, db_dv = ue / den;
        const double w = u * ce - v * se;
        const double J20 = se - w * db_du, J21 = ce - w * db_dv, J30 = db_du / T1, J31 = db_dv / T1;
        const double J57 = -u * sp - v * cp, J67 = u * cp - v * sp;
        // the same terms in the same (column) order as the dense product Jx s (+ Ju)
        ks[0] = 0.0;
        ks[1] = 0.0;
        ks[2] = (J20 * s[0] + J21 * s[1]) + w * s[3];
        ks[3] = ((J30 * s[0] + J31 * s[1]) + (-1.0 / T1) * s[3]) + (1.0 / T1) * s[4];
        ks[4] = ucol == 0 ? 1.0 : 0.0;
        ks[5] = (cp * s[0] + (-sp) * s[1]) + J57 * s[7];
        ks[6] = (sp * s[0] + cp * s[1]) + J67 * s[7];
        ks[7] = ((J30 * s[0] + J31 * s[1]) + (-1.0 / T1) * s[3]) + (1.0 / T1) * s[4];
    }
};

struct Pendulum {
    static constexpr int ID = 1;
    static constexpr int NX = 4, NU = 1;
    static constexpr int HX = 0, HY = 1;  // unused (the pendulum OCP has no h)
    static constexpr int NKIN = 1;        // the cart position does not enter the dynamics

    MDEV static void f_jac(const double* x, const double* uc, double* f, double* Jx, double* Ju)
    {
        const double M = 1.0, m = 0.1, g = 9.81, l = 0.8;
        const double th = x[1], v1 = x[2], dth = x[3], F = uc[0];
        double s, c;
        dsincos(th, &s, &c);
        const double den = M + m - m * c * c;
        const double n3 = -m * l * s * dth * dth + m * g * c * s + F;
        const double n4 = -m * l * c * s * dth * dth + F * c + (M + m) * g * s;
        f[0] = v1;
        f[1] = dth;
        f[2] = n3 / den;
        f[3] = n4 / (l * den);
        const double dden = 2 * m * c * s;
        const double dn3_th = -m * l * c * dth * dth + m * g * (c * c - s * s);
        const double dn4_th = -m * l * (c * c - s * s) * dth * dth - F * s + (M + m) * g * c;
#pragma unroll
        for (int i = 0; i < 16; i++) Jx[i] = 0.0;
        Jx[0 + 4 * 2] = 1.0;
        Jx[1 + 4 * 3] = 1.0;
        Jx[2 + 4 * 1] = (dn3_th * den - n3 * dden) / (den * den);
        Jx[2 + 4 * 3] = -2 * m * l * s * dth / den;
        Jx[3 + 4 * 1] = (dn4_th * den - n4 * dden) / (l * den * den);
        Jx[3 + 4 * 3] = -2 * m * l * c * s * dth / (l * den);
        Ju[0] = 0.0; Ju[1] = 0.0; Ju[2] = 1.0 / den; Ju[3] = c / (l * den);
    }

    // one column of the variational equation (see Usv3::vde_col): dense product on the 4x4 Jacobian
    MDEV static void vde_col(const double* x, const double* uc, const double* s, int ucol, double* f, double* ks)
    {
        double Jx[NX * NX], Ju[NX * NU];
        f_jac(x, uc, f, Jx, Ju);
#pragma unroll
        for (int i = 0; i < NX; i++)
        {
            double acc = ucol == 0 ? Ju[i] : 0.0;
#pragma unroll
            for (int m = 0; m < NX; m++) acc += Jx[i + NX * m] * s[m];
            ks[i] = acc;
        }
    }
};

}  // namespace usvmpc
