// nmpc_kernel.cuh -- the batched NMPC solve: ONE WARP PER INSTANCE runs the whole
// SQP / SQP_RTI solve (linearise -> x0 elimination -> Mehrotra IPM on a square-root Riccati
// recursion -> variable update) out of its own block of HBM (layout.h).
//
// What it replaces in the reference (AC = catkin_ws/src/nmpc_ca/acados, HP = AC/external/hpipm):
//   ocp_nlp_sqp / ocp_nlp_sqp_rti loop            AC/acados/ocp_nlp/ocp_nlp_sqp.c:466-835, ocp_nlp_sqp_rti.c:445-829
//   linearisation + KKT residuals                 AC/acados/ocp_nlp/ocp_nlp_common.c:1926-2084, 2549-2603
//   ERK with forward sensitivities                AC/acados/sim/sim_erk_integrator.c:668-847
//   LINEAR_LS cost, BGH constraints               AC/acados/ocp_nlp/ocp_nlp_cost_ls.c:713-843, ocp_nlp_constraints_bgh.c:1228-1430
//   x0 elimination / restoration                  HP/ocp_qp/x_ocp_qp_red.c:268-454, 723-871
//   HPIPM IPM (init, delta step, residuals)       HP/ocp_qp/x_ocp_qp_ipm.c:1387-1719, 1888-2350, 2354-2683; x_ocp_qp_res.c:336-633
//   Riccati factorise+solve / solve               HP/ocp_qp/x_ocp_qp_kkt.c:405-766, 1096-1441
//   core vector ops                               HP/ipm_core/x_core_qp_ipm_aux.c:38-357
//   BLASFEO potrf/syrk/trmm/trsv/gemv             BF/blasfeo_hp_pm/d_lapack_lib4.c:1149,1503 etc. -> warp-level code below
//
// Work decomposition inside the warp:
//   * one IPM iteration is FOUR streaming sweeps over the per-stage records (layout.h), each pulling record k
//     into shared memory with cp.async one stage ahead of the arithmetic and writing back only what it changed:
//       B  forward : affine step  (forward substitution, dt/dlam, step length)
//       C  backward: corrector right-hand side + backward substitution
//       D  forward : corrector step + residual of the linear system (iterative-refinement test)
//       A  backward: variable update + QP residuals + Gamma/gamma + Riccati factorisation for the next iteration
//     The Riccati parts are serial in the stage index: lane r owns row r of the (nv+1) x nv factor, inner products
//     go through shared memory / shuffles; everything else of a stage is spread over the lanes.
//   * RK4 sensitivities, cost/constraint evaluation, the IPM start point and the NLP update are independent per
//     stage and are spread over the lanes (one lane per stage or per sensitivity column).
// Inactive variables (x at stage 0 after x0 elimination, u at stage N) and inactive inequality rows
// are kept in the uniform per-stage layout and masked (identity rows in the factor, zero rows in
// B'/A'), so every stage runs the same code.
#pragma once
#include "layout.h"
#include "models.cuh"
#include "warp_compat.h"

namespace usvmpc {

template <class M>
struct WarpSolver {
    static constexpr int NX = M::NX, NU = M::NU, NV = NX + NU, NR = NV + 1, NY = NV;
    static constexpr int HXV = NU + M::HX, HYV = NU + M::HY;
    static constexpr int NE = NV * (NV + 1) / 2 + NV;  // entries of the lower trapezoid of the (NV+1) x NV factor
    // record layout (layout.h): offsets of the fixed-size fields are compile-time constants
    static constexpr int svv = (NV + 1) / 2 * 2, sxx = (NX + 1) / 2 * 2, sLL = (NR * NV + 1) / 2 * 2;
    static constexpr int oBAt = 0, oux = (NV * NX + 1) / 2 * 2, opi = oux + svv, opip = opi + sxx, org = opip + sxx, orb = org + svv;
    static constexpr int oL = orb + sxx, oPb = oL + sLL, odux = oPb + sxx, odpi = odux + svv, odpip = odpi + sxx;
    static constexpr int orq = odpip + sxx, ob = orq + svv, ogxy = ob + sxx;
    // per-warp shared-memory scratch: fixed-size part at compile-time offsets
    static constexpr int qHs = 0, qHes = qHs + NV * NV, qWs = qHes + NV * NV, qWes = qWs + NV * NV, qTp = qWes + NX * NX;
    static constexpr int qAL = qTp + 3 * NE, qz = qAL + NR * NX, qent = qz + sxx, qvrow = qent + (NE + 1) / 2;
    static constexpr int qxrow = qvrow + (NV + 1) / 2, qvar = qxrow + (NX + 1) / 2;

    const Params& P;
    const Layout& Y;
    int lane, N, K, nbu, nbx, ncq, ncz, nbq, nct;
    double* w;
    // shared-memory scratch of this warp
    double *Hs, *Hes, *Ws, *Wes, *Tp, *sAL, *sGs, *sgd, *sdl, *sz;
    int *sent, *srvar, *svrow, *sxrow;
    double *sBA, *sLn, *slx, *sG, *sg, *sL, *sq, *sx1, *sx2, *sgxy;  // rare path, aliased onto the record buffers
    // the two record buffers of the streaming sweeps and the offsets of the record's fields (layout.h)
    double* buf[3];
    unsigned long long* bar;  // one mbarrier per record buffer (TMA completion)
    unsigned phb;             // their phase bits
    int olam, ot, ord, oti, ormc, odlam, odt, od, scq;  // K-dependent record offsets
    // IPM arguments (HP/ocp_qp/x_ocp_qp_ipm.c:133-161 overridden by AC/acados/ocp_qp/ocp_qp_hpipm.c:106-116
    // and, in SQP mode, by AC/acados/ocp_nlp/ocp_nlp_sqp.c:201-227)
    double tol_stat, tol_eq, tol_ineq, tol_comp;
    int iter_max;
    // IPM state (warp-uniform)
    double res_max[4], mu, mu_aff, sigma, alpha;
    double S1, S2;  // sum(lam*dt + t*dlam), sum(dlam*dt) of the last expanded step
    int solve_calls;

    MDEV WarpSolver(const Params& p, int inst, double* sm) : P(p), Y(p.lay)
    {
        lane = lane_id();
        N = P.N; K = P.K; nbu = P.nbu; nbx = P.nbx; ncq = P.ncq; ncz = P.ncz; nbq = nbu + nbx;
        nct = N >= 1 ? 2 * ((nbu + K) + (N - 1) * (nbu + nbx + K)) : 0;
        w = P.ws + (long) inst * P.ws_stride;
        double* s = sm;
        Hs = s + qHs; Hes = s + qHes; Ws = s + qWs; Wes = s + qWes; Tp = s + qTp; sAL = s + qAL; sz = s + qz;
        sent = (int*) (s + qent); svrow = (int*) (s + qvrow); sxrow = (int*) (s + qxrow);
        const int nq = NU + NX + K;
        s += qvar;
        srvar = (int*) s; s += (nq + 1) / 2; sGs = s; s += nq; sgd = s; s += nq; sdl = s; s += nq;
        s = sm + ((s - sm) + 1) / 2 * 2;
        buf[0] = s; s += Y.rec_size; buf[1] = s; s += Y.rec_size; buf[2] = s; s += Y.rec_size;
        bar = (unsigned long long*) s; s += 4;
        phb = 0;
        scq = Y.t.off - Y.lam.off;
        olam = Y.lam.off - Y.rec_off; ot = olam + scq; ord = ot + scq; oti = ord + scq; ormc = oti + scq; odlam = ormc + scq;
        odt = odlam + scq; od = odt + scq;
        // scratch of the rare (iterative refinement) path lives in the record buffers, which are idle then
        double* q = buf[0];
        sBA = q; q += NV * NX; sLn = q; q += NX * NX; slx = q; q += NX; sG = q; q += 2 * (NU + NX + K);
        sg = q; q += 2 * (NU + NX + K); sL = q; q += NR * NV; sq = q; q += NV; sx1 = q; q += NX; sx2 = q; q += NX;
        sgxy = q;
        tol_stat = 1e-6; tol_eq = 1e-8; tol_ineq = 1e-8; tol_comp = 1e-8;
        if (P.nlp_type == 0) { tol_stat = P.tol[0]; tol_eq = P.tol[1]; tol_ineq = P.tol[2]; tol_comp = P.tol[3]; }
        iter_max = P.qp_iter_max > 0 ? P.qp_iter_max : 50;
        solve_calls = 0;
    }

    // the host-side layout (layout.h:make_layout) must agree with the compile-time offsets above
    static bool layout_matches(const Layout& y)
    {
        const int ro = y.rec_off;
        return y.BAt.off - ro == oBAt && y.ux.off - ro == oux && y.pi.off - ro == opi && y.pi_prev.off - ro == opip &&
               y.rg.off - ro == org && y.rb.off - ro == orb && y.L.off - ro == oL && y.Pb.off - ro == oPb &&
               y.dux.off - ro == odux && y.dpi.off - ro == odpi && y.dpi_prev.off - ro == odpip && y.rq.off - ro == orq &&
               y.b.off - ro == ob && y.gxy.off - ro == ogxy && y.lam.off > y.gxy.off && y.rd.off - y.t.off == y.t.off - y.lam.off && y.ti.off - y.rd.off == y.t.off - y.lam.off &&
               y.rmc.off - y.ti.off == y.t.off - y.lam.off &&
               y.d.off - y.dt.off == y.t.off - y.lam.off && y.rec_size == y.d.off - ro + (y.t.off - y.lam.off);
    }
    MDEV double* F(const Field& f, int k) const { return w + f.off + (long) k * f.stride; }
    MDEV bool var_active(int k, int i) const { return k == 0 ? (i < NU) : (k == N ? (i >= NU) : true); }
    MDEV bool row_active(int k, int j) const { return k < N && (j < nbu || j >= nbq || k >= 1); }
    // IPM row of the box on variable c at stage k, or -1
    MDEV int vrow(int k, int c) const
    {
        if (k >= N) return -1;
        if (c < NU) return c < nbu ? c : -1;
        return k >= 1 ? sxrow[c - NU] : -1;
    }

    // ---------------------------------------------------------------- constants into shared memory
    // Gauss-Newton Hessians: ocp_nlp_cost_ls_initialize, AC/acados/ocp_nlp/ocp_nlp_cost_ls.c:713-745.  With
    // Vx=[I;0], Vu=[0;I] the output map y = Cyt'[u;x] is the permutation [x;u].  Lower triangles are read.
    MDEV void load_constants()
    {
        const double* Wg = P.cst;
        const double* Weg = P.cst + NY * NY;
        for (int e = lane; e < NV * NV; e += 32)
        {
            const int i = e % NV, j = e / NV;
            Ws[e] = i >= j ? Wg[i + NY * j] : Wg[j + NY * i];
            const int yi = i < NU ? NX + i : i - NU, yj = j < NU ? NX + j : j - NU;
            Hs[e] = P.dt * (yi >= yj ? Wg[yi + NY * yj] : Wg[yj + NY * yi]);
            Hes[e] = (i >= NU && j >= NU) ? ((i >= j) ? Weg[(i - NU) + NX * (j - NU)] : Weg[(j - NU) + NX * (i - NU)]) : 0.0;
        }
        for (int e = lane; e < NX * NX; e += 32)
        {
            const int i = e % NX, j = e / NX;
            Wes[e] = i >= j ? Weg[i + NX * j] : Weg[j + NX * i];
        }
        if (lane < NX)
        {
            int r = -1;
            for (int j = 0; j < nbx; j++) if (P.idxbx[j] == lane) r = nbu + j;
            sxrow[lane] = r;
        }
        syncwarp();
        // tables of the streaming sweeps: lower-trapezoid entries e -> (r, c, offset); row pair -> boxed variable
        // (or -1: h row); variable -> row pair of its box (or -1); factor templates per stage class
        for (int e = lane; e < NE; e += 32)
        {
            int r = 0;
            while ((r + 1) * (r + 2) / 2 <= e && r < NV) r++;
            const int c = e - r * (r + 1) / 2;
            sent[e] = (r << 24) | (c << 16) | (r * NV + c);
            for (int cls = 0; cls < 3; cls++)
            {
                const double* H = cls == 2 ? Hes : Hs;
                double v = 0.0;
                if (r < NV)
                {
                    const bool ar = cls == 1 || (cls == 0 ? r < NU : r >= NU), ac = cls == 1 || (cls == 0 ? c < NU : c >= NU);
                    if (ar && ac) { v = H[r + NV * c]; if (c == r) v += 1e-15; }  // reg_prim
                    else if (c == r) v = 1.0;
                }
                Tp[cls * NE + e] = v;
            }
        }
        for (int jj = lane; jj < ncq; jj += 32) srvar[jj] = jj < nbu ? jj : (jj < nbq ? NU + P.idxbx[jj - nbu] : -1);
        if (lane < NV) svrow[lane] = lane < NU ? (lane < nbu ? lane : -1) : sxrow[lane - NU];
        syncwarp();
    }
    MDEV const double* Hk(int k) const { return k < N ? Hs : Hes; }

    // ---------------------------------------------------------------- initial guess
    // cold start of the scripts / template (acados_solver.in.c:1595-1623): x_k = x0, u = 0, pi = 0;
    // lam, t start at zero like a freshly created nlp_out.
    MDEV void cold_start(const double* x0)
    {
        for (int k = lane; k <= N; k += 32)
        {
            double* z = F(Y.zux, k);
            for (int i = 0; i < NU; i++) z[i] = 0.0;
            for (int i = 0; i < NX; i++) z[NU + i] = x0[i];
            double* pi = F(Y.zpi, k);
            for (int i = 0; i < NX; i++) pi[i] = 0.0;
            double *l = F(Y.zlam, k), *t = F(Y.zt, k);
            for (int j = 0; j < 2 * ncz; j++) { l[j] = 0.0; t[j] = 0.0; }
        }
        syncwarp();
    }

    // ---------------------------------------------------------------- linearisation
    // ERK with forward sensitivities (AC/acados/sim/sim_erk_integrator.c:762-847, tableaus :253-344; seed S=[I 0],
    // A=Sx(T), B=Su(T): ocp_nlp_dynamics_cont.c:782-804).  One lane integrates [x ; one sensitivity column] of one
    // stage; a stage's NV columns sit on NV consecutive task slots.
    MDEV void integrate_all()
    {
        const int ns = P.num_stages;
        double a21 = 0, a32 = 0, a43 = 0, bv0 = 0, bv1 = 0, bv2 = 0, bv3 = 0;
        if (ns == 1) { bv0 = 1.0; }
        else if (ns == 2) { a21 = 0.5; bv1 = 1.0; }
        else { a21 = 0.5; a32 = 0.5; a43 = 1.0; bv0 = 1.0 / 6.0; bv1 = 1.0 / 3.0; bv2 = 1.0 / 3.0; bv3 = 1.0 / 6.0; }
        const double step = P.dt / P.num_steps;
        for (int task = lane; task < N * NV; task += 32)
        {
            const int k = task / NV, col = task % NV;
            const double* z = F(Y.zux, k);
            double u[NU], x[NX], s[NX];
#pragma unroll
            for (int i = 0; i < NU; i++) u[i] = z[i];
#pragma unroll
            for (int i = 0; i < NX; i++) { x[i] = z[NU + i]; s[i] = (i == col) ? 1.0 : 0.0; }
            for (int istep = 0; istep < P.num_steps; istep++)
            {
                double xr[NX], sr[NX], xa[NX], sa[NX];
#pragma unroll
                for (int i = 0; i < NX; i++) { xr[i] = x[i]; sr[i] = s[i]; xa[i] = x[i]; sa[i] = s[i]; }
#pragma unroll
                for (int st = 0; st < 4; st++)
                {
                    if (st >= ns) break;
                    double f[NX], ks[NX];
                    // VDE right-hand side of this column: Jx*Sx_col, or Jx*Su_col + Ju_col
                    // (acados_template/generate_c_code_explicit_ode.py:73-80)
                    M::vde_col(xr, u, sr, col >= NX ? col - NX : -1, f, ks);
                    const double bb = step * (st == 0 ? bv0 : st == 1 ? bv1 : st == 2 ? bv2 : bv3);
                    const double aa = (st == 0 ? a21 : st == 1 ? a32 : st == 2 ? a43 : 0.0) * step;
#pragma unroll
                    for (int i = 0; i < NX; i++)
                    {
                        xa[i] += bb * f[i]; sa[i] += bb * ks[i];
                        xr[i] = x[i]; sr[i] = s[i];
                        if (aa != 0.0) { xr[i] += aa * f[i]; sr[i] += aa * ks[i]; }
                    }
                }
#pragma unroll
                for (int i = 0; i < NX; i++) { x[i] = xa[i]; s[i] = sa[i]; }
            }
            // BAt = [B'; A'] (nv x nx, column-major): ocp_nlp_dynamics_cont.c:801-804
            double* BAt = F(Y.BAt, k);
            const int row = col < NX ? NU + col : col - NX;
#pragma unroll
            for (int i = 0; i < NX; i++) BAt[row + NV * i] = s[i];
            if (col == 0)
            {
                const double* zn = F(Y.zux, k + 1);
                double* b = F(Y.b, k);
#pragma unroll
                for (int i = 0; i < NX; i++) b[i] = x[i] - zn[NU + i];  // dyn_fun = phi(x,u) - x_next
            }
        }
        syncwarp();
    }

    // cost / constraints / adjoints / NLP residuals / QP vectors, one lane per stage.
    // ocp_nlp_approximate_qp_matrices + _vectors_sqp (ocp_nlp_common.c:1926-2084), ocp_nlp_res_compute (:2549-2603),
    // x0 elimination d_ocp_qp_reduce_eq_dof (HP/ocp_qp/x_ocp_qp_red.c:268-454).  res4 = (stat, eq, ineq, comp).
    MDEV void linearize(const double* x0, const double* pg, const double* lhg, const double* yrg, const double* yre,
                       double* res4)
    {
        integrate_all();
        double r0 = 0, r1 = 0, r2 = 0, r3 = 0;
        for (int k = lane; k <= N; k += 32)
        {
            const double* z = F(Y.zux, k);
            const double *zl = F(Y.zlam, k), *zt = F(Y.zt, k);
            double *zf = F(Y.zfun, k), *rq = F(Y.rq, k), *d = F(Y.d, k);
            double cg[NV], adj[NV];
            // ---- LINEAR_LS cost gradient (ocp_nlp_cost_ls.c:749-843)
            if (k < N)
            {
                const double* yr = yrg + (P.yref_per_stage ? k * NY : 0);
                double r[NY];
#pragma unroll
                for (int i = 0; i < NX; i++) r[i] = z[NU + i] - yr[i];
#pragma unroll
                for (int i = 0; i < NU; i++) r[NX + i] = z[i] - yr[NX + i];
#pragma unroll
                for (int i = 0; i < NY; i++)
                {
                    double acc = 0.0;
#pragma unroll
                    for (int j = 0; j < NY; j++) acc += Ws[i + NY * j] * r[j];
                    if (i < NX) cg[NU + i] = P.dt * acc; else cg[i - NX] = P.dt * acc;
                }
            }
            else
            {
                double r[NX];
#pragma unroll
                for (int i = 0; i < NX; i++) r[i] = z[NU + i] - yre[i];
#pragma unroll
                for (int i = 0; i < NU; i++) cg[i] = 0.0;
#pragma unroll
                for (int i = 0; i < NX; i++)
                {
                    double acc = 0.0;
#pragma unroll
                    for (int j = 0; j < NX; j++) acc += Wes[i + NX * j] * r[j];
                    cg[NU + i] = acc;
                }
            }
#pragma unroll
            for (int i = 0; i < NV; i++) adj[i] = 0.0;
            // ---- BGH constraints (ocp_nlp_constraints_bgh.c:1228-1430): fun = [lb - g ; g - ub], adj = J'(lam_l - lam_u)
            for (int j = 0; j < 2 * ncz; j++) zf[j] = 0.0;
            for (int j = 0; j < 2 * ncq; j++) d[j] = 0.0;
            double dx0[NX];
#pragma unroll
            for (int i = 0; i < NX; i++) dx0[i] = 0.0;
            if (k < N)
            {
#pragma unroll
                for (int j = 0; j < NU; j++)
                    if (j < nbu)
                    {
                        const double g = z[j], fl = P.lbu[j] - g, fu = g - P.ubu[j];
                        zf[j] = fl; zf[ncz + j] = fu; d[j] = fl; d[ncq + j] = fu;
                        adj[j] += zl[j] - zl[ncz + j];
                        const double a = dabs(fl + zt[j]), b = dabs(fu + zt[ncz + j]);
                        r2 = a > r2 ? a : r2; r2 = b > r2 ? b : r2;
                        const double c0 = dabs(zl[j] * zt[j]), c1 = dabs(zl[ncz + j] * zt[ncz + j]);
                        r3 = c0 > r3 ? c0 : r3; r3 = c1 > r3 ? c1 : r3;
                    }
                if (k == 0)
                {
                    // x0 embedding: lbx = ubx = x0 on every state (acados_solver.in.c:1028-1051, bgh.c:1528)
#pragma unroll
                    for (int j = 0; j < NX; j++)
                    {
                        const double g = z[NU + j], fl = x0[j] - g, fu = g - x0[j];
                        const int r = nbu + j;
                        zf[r] = fl; zf[ncz + r] = fu;
                        dx0[j] = fl;
                        adj[NU + j] += zl[r] - zl[ncz + r];
                        const double a = dabs(fl + zt[r]), b = dabs(fu + zt[ncz + r]);
                        r2 = a > r2 ? a : r2; r2 = b > r2 ? b : r2;
                        const double c0 = dabs(zl[r] * zt[r]), c1 = dabs(zl[ncz + r] * zt[ncz + r]);
                        r3 = c0 > r3 ? c0 : r3; r3 = c1 > r3 ? c1 : r3;
                    }
                }
                else
                {
                    for (int j = 0; j < nbx; j++)
                    {
                        const int id = P.idxbx[j], r = nbu + j;
                        const double g = z[NU + id], fl = P.lbx[j] - g, fu = g - P.ubx[j];
                        zf[r] = fl; zf[ncz + r] = fu; d[r] = fl; d[ncq + r] = fu;
                        const double dl = zl[r] - zl[ncz + r];
#pragma unroll
                        for (int i = 0; i < NX; i++) if (i == id) adj[NU + i] += dl;
                        const double a = dabs(fl + zt[r]), b = dabs(fu + zt[ncz + r]);
                        r2 = a > r2 ? a : r2; r2 = b > r2 ? b : r2;
                        const double c0 = dabs(zl[r] * zt[r]), c1 = dabs(zl[ncz + r] * zt[ncz + r]);
                        r3 = c0 > r3 ? c0 : r3; r3 = c1 > r3 ? c1 : r3;
                    }
                }
                // obstacle distances h_c = ||(X,Y) - (ox_c, oy_c)||, dh/d(X,Y) = ((X,Y) - o_c)/h_c
                const double* pk = pg + (P.p_per_stage ? k * 2 * K : 0);
                const double* lhk = lhg + (P.lh_per_stage ? k * K : 0);
                double* gxy = F(Y.gxy, k);
                for (int c = 0; c < K; c++)
                {
                    const double ddx = z[HXV] - pk[2 * c], ddy = z[HYV] - pk[2 * c + 1];
                    const double h = dsqrt(ddx * ddx + ddy * ddy);
                    const double gX = ddx / h, gY = ddy / h;
                    gxy[c] = gX; gxy[K + c] = gY;
                    const double fl = lhk[c] - h, fu = h - P.uh;
                    const int r = nbu + NX + c, rqp = nbq + c;
                    zf[r] = fl; zf[ncz + r] = fu;
                    // stage 0: fold the eliminated x0 step into the bounds (x_ocp_qp_red.c:380-420)
                    const double v = (k == 0) ? gX * dx0[M::HX] + gY * dx0[M::HY] : 0.0;
                    d[rqp] = fl - v; d[ncq + rqp] = fu + v;
                    const double dl = zl[r] - zl[ncz + r];
                    adj[HXV] += gX * dl; adj[HYV] += gY * dl;
                    const double a = dabs(fl + zt[r]), b = dabs(fu + zt[ncz + r]);
                    r2 = a > r2 ? a : r2; r2 = b > r2 ? b : r2;
                    const double c0 = dabs(zl[r] * zt[r]), c1 = dabs(zl[ncz + r] * zt[ncz + r]);
                    r3 = c0 > r3 ? c0 : r3; r3 = c1 > r3 ? c1 : r3;
                }
            }
            // ---- dynamics adjoint -[B';A'] pi_k (+ pi_{k-1} on x): ocp_nlp_common.c:2001-2019 ; stationarity residual
            const double* BAt = F(Y.BAt, k);
            const double* pik = F(Y.zpi, k);
#pragma unroll
            for (int i = 0; i < NV; i++)
            {
                double acc = 0.0;
                if (k < N)
                {
#pragma unroll
                    for (int j = 0; j < NX; j++) acc -= BAt[i + NV * j] * pik[j];
                }
                if (k > 0 && i >= NU) acc += F(Y.zpi, k - 1)[i - NU];
                if (k < N || i >= NU)
                {
                    const double a = dabs(cg[i] - adj[i] - acc);
                    r0 = a > r0 ? a : r0;
                }
            }
            if (k < N)
            {
                const double* b = F(Y.b, k);
#pragma unroll
                for (int i = 0; i < NX; i++) { const double a = dabs(b[i]); r1 = a > r1 ? a : r1; }
            }
            // ---- QP gradient; stage 0: b0 += A0' dx0, r0 += S dx0 (x_ocp_qp_red.c:300-378)
#pragma unroll
            for (int i = 0; i < NV; i++) rq[i] = cg[i];
            if (k == 0 && N > 0)
            {
                double* b = F(Y.b, 0);
#pragma unroll
                for (int j = 0; j < NX; j++)
                {
                    double acc = b[j];
#pragma unroll
                    for (int i = 0; i < NX; i++) acc += BAt[NU + i + NV * j] * dx0[i];
                    b[j] = acc;
                }
#pragma unroll
                for (int i = 0; i < NU; i++)
                {
                    double acc = cg[i];
#pragma unroll
                    for (int j = 0; j < NX; j++) acc += Hs[(NU + j) + NV * i] * dx0[j];
                    rq[i] = acc;
                }
            }
        }
        res4[0] = warp_max(r0); res4[1] = warp_max(r1); res4[2] = warp_max(r2); res4[3] = warp_max(r3);
        syncwarp();
    }

    // ---------------------------------------------------------------- IPM: per-stage (lane-parallel) passes
    // OCP_QP_INIT_VAR, var_init_scheme 1, cold start: HP/ocp_qp/x_ocp_qp_ipm.c:1435-1470,1581-1714 (ns = 0)
    MDEV void ipm_init()
    {
        const double thr0 = 1e-1, mu0 = 1.0;
        for (int k = lane; k <= N; k += 32)
        {
            double *ux = F(Y.ux, k), *pi = F(Y.pi, k), *lam = F(Y.lam, k), *t = F(Y.t, k);
            const double* d = F(Y.d, k);
            for (int i = 0; i < NV; i++) ux[i] = 0.0;
            for (int i = 0; i < NX; i++) pi[i] = 0.0;
            for (int j = 0; j < 2 * ncq; j++) { lam[j] = 0.0; t[j] = 1.0; }
            {
                // the first sweep A applies a zero step: the step and the duplicated neighbour values start at zero
                double *a = F(Y.dux, k), *b = F(Y.dpi, k), *c = F(Y.dlam, k), *e = F(Y.dt, k);
                double *pp = F(Y.pi_prev, k), *dp = F(Y.dpi_prev, k);
                for (int i = 0; i < NV; i++) a[i] = 0.0;
                for (int i = 0; i < NX; i++) { b[i] = 0.0; pp[i] = 0.0; dp[i] = 0.0; }
                for (int j = 0; j < 2 * ncq; j++) { c[j] = 0.0; e[j] = 0.0; }
            }
            if (k >= N) continue;
            for (int j = 0; j < nbq; j++)
            {
                if (!row_active(k, j)) continue;
                const int id = j < nbu ? j : NU + P.idxbx[j - nbu];
                double tl = ux[id] - d[j], tu = -ux[id] - d[ncq + j];
                if (tl < thr0)
                {
                    if (tu < thr0) { ux[id] = 0.5 * (d[j] - d[ncq + j]); tl = thr0; tu = thr0; }
                    else { tl = thr0; ux[id] = d[j] + thr0; }
                }
                else if (tu < thr0) { tu = thr0; ux[id] = -d[ncq + j] - thr0; }
                t[j] = tl; t[ncq + j] = tu;
            }
            const double* gxy = F(Y.gxy, k);
            for (int c = 0; c < K; c++)
            {
                const double v = (k >= 1) ? gxy[c] * ux[HXV] + gxy[K + c] * ux[HYV] : 0.0;
                const double tl = v - d[nbq + c], tu = -v - d[ncq + nbq + c];
                t[nbq + c] = thr0 > tl ? thr0 : tl;
                t[ncq + nbq + c] = thr0 > tu ? thr0 : tu;
            }
            for (int j = 0; j < ncq; j++)
                if (row_active(k, j)) { lam[j] = mu0 / t[j]; lam[ncq + j] = mu0 / t[ncq + j]; }
        }
        syncwarp();
    }

    // ---------------------------------------------------------------- IPM: record streaming
    // The IPM works on one RECORD per stage (layout.h).  Each sweep pulls record k into one of three shared-memory
    // buffers with asynchronous 16-byte copies (cp.async), one stage ahead of the arithmetic, computes in place and
    // writes the range it modified back with coalesced stores; the third buffer keeps the previous stage's record
    // (its Riccati factor / solution is the input of the recursion) so nothing is copied between stages.
    // The sweeps are written for a SMALL instruction footprint (rolled loops over shared-memory operands, work
    // spread over all 32 lanes): a lone warp finishing a hard instance is bound by instruction fetch otherwise.
    MDEV double* rec_g(int k) const { return w + Y.rec_off + (long) k * Y.rec_size; }
    // TMA: one lane issues one bulk copy for the whole record; completion is counted in bytes on the buffer's mbarrier
    MDEV void rec_init()
    {
        if (lane == 0) { mbar_init(bar + 0); mbar_init(bar + 1); mbar_init(bar + 2); fence_mbar_init(); }
        syncwarp();
    }
    MDEV void rec_fetch(int k, int b)
    {
        if (lane == 0) bulk_g2s(buf[b], rec_g(k), Y.rec_size * 8, bar + b);
    }
    MDEV void rec_wait(int b)
    {
        mbar_wait(bar + b, (phb >> b) & 1u);
        phb ^= 1u << b;
    }
    // write two ranges of the record back (bulk store, asynchronous); the generic-proxy writes to the buffer are
    // fenced towards the async proxy first
    MDEV void rec_store2(int k, int b, int f0, int t0, int f1, int t1)
    {
        fence_proxy_async_smem();
        syncwarp();
        if (lane == 0)
        {
            double* g = rec_g(k);
            bulk_s2g(g + f0, buf[b] + f0, (t0 - f0) * 8);
            if (t1 > f1) bulk_s2g(g + f1, buf[b] + f1, (t1 - f1) * 8);
            bulk_commit();
        }
    }
    // before a buffer that was the source of a store becomes the destination of a load again: all but the most recent
    // store group have finished reading shared memory
    MDEV void rec_reuse_guard() { if (lane == 0) bulk_wait_read1(); }
    // sweep boundaries: records written through the generic proxy (start point, rare path) / the async proxy (sweeps)
    MDEV void sweep_begin() { fence_proxy_async(); syncwarp(); }
    MDEV void sweep_end()
    {
        if (lane == 0) bulk_wait_all();
        syncwarp();
        fence_proxy_async();
    }
    // stage 0 after x0 elimination: no x rows in [B';A'], no Jacobian of the h rows (x_ocp_qp_red.c:268-454)
    MDEV void rec_mask_stage0(double* R)
    {
#pragma unroll 1
        for (int e = lane; e < NV * NX; e += 32) if (e % NV >= NU) R[oBAt + e] = 0.0;
#pragma unroll 1
        for (int e = lane; e < 2 * K; e += 32) R[ogxy + e] = 0.0;
        syncwarp();
    }
    // inequality rows are handled as (lower, upper) pairs jj < ncq: pair active at this stage class?
    MDEV bool pair_active(int cls, int jj) const { return cls == 1 || (cls == 0 && (jj < nbu || jj >= nbq)); }
    // IPM row pair of the box on variable i for this stage class, or -1
    MDEV int box_of_var(int cls, int i) const { return cls == 1 ? svrow[i] : (cls == 0 && i < NU ? svrow[i] : -1); }
    // residual-type kernel shared by sweep A (iterate) and sweep D (step): lanes < NV produce
    //   g_i = sum_j H[i][j] v[j] + c_i - pprev_i + (box / h-row multiplier differences) + sum_j [B';A'][i][j] p[j]
    // and lanes NV..NV+NX-1 produce  b_j = cb_j - xnext_j + sum_i [B';A'][i][j] v[i]   (x_ocp_qp_res.c:336-466, 468-592)
    MDEV double res_gb(const double* R, int k, int cls, int ov, int op, int opp, int oc, int ocb, const double* xnext)
    {
        double acc = 0.0;
        if (lane < NV + NX)
        {
            const bool isg = lane < NV;
            if (!isg && k >= N) return 0.0;
            const double* pa = isg ? (cls == 2 ? Hes : Hs) + lane : R + oBAt + NV * (lane - NV);
            const int sa = isg ? NV : 1;
            if (!isg) acc = R[ocb + lane - NV] - xnext[lane - NV];
#pragma unroll
            for (int m = 0; m < NV; m++) acc += pa[m * sa] * R[ov + m];
            if (isg)
            {
                const int i = lane;
                double g = acc + R[oc + i];
                if (i >= NU) g -= R[opp + i - NU];
                const int row = box_of_var(cls, i);
                if (row >= 0) g += sdl[row];
                if (k < N)
                {
                    if (i == HXV || i == HYV)
                    {
                        const double* gq = R + ogxy + (i == HXV ? 0 : K);
#pragma unroll 1
                        for (int c = 0; c < K; c++) g += gq[c] * sdl[nbq + c];
                    }
                    double acc2 = 0.0;
#pragma unroll
                    for (int j = 0; j < NX; j++) acc2 += R[oBAt + i + NV * j] * R[op + j];
                    g += acc2;
                }
                const bool act = cls == 1 || (cls == 0 ? i < NU : i >= NU);
                acc = act ? g : 0.0;
            }
        }
        return acc;
    }

    // Sweep A (backward, k = N..0), three steps of the reference fused per stage:
    //   UPDATE_VAR_QP with step length a (x_core_qp_ipm_aux.c:220-325) -> OCP_QP_RES_COMPUTE at the new iterate
    //   (x_ocp_qp_res.c:336-466; norms into n4, mu) -> backward part of OCP_QP_FACT_SOLVE_KKT_STEP for the affine
    //   right-hand side res_m = lam*t - tau (x_ocp_qp_kkt.c:405-535, COMPUTE_GAMMA_GAMMA_QP x_core_qp_ipm_aux.c:38-86).
    // The factorisation is speculative: if the residuals turn out to be converged it is simply not used.
    // The (NV+1) x NV matrix [H + Gamma terms + AL AL' ; gradient row] is built and factorised in place in the
    // record's L field, one lower-trapezoid entry (or two) per lane.
    MDEV void sweepA(double a, double tau, double reg, double* n4)
    {
        const double lam_min = 1e-16, t_min = 1e-16;
        double n0 = 0, n1 = 0, n2 = 0, n3 = 0, musum = 0;
        int ir = 0, in = 1, ip = 2;
        sweep_begin();
        rec_fetch(N, ir);
#pragma unroll 1
        for (int k = N; k >= 0; k--)
        {
            double *R = buf[ir], *Rp = buf[ip];
            rec_wait(ir);
            if (k > 0) { rec_reuse_guard(); rec_fetch(k - 1, in); }
            const int cls = k == 0 ? 0 : (k < N ? 1 : 2);
            if (k == 0) rec_mask_stage0(R);
            // ---- update: [ux | pi | pi_prev] += a [dux | dpi | dpi_prev]
            if (lane < svv + 2 * sxx) R[oux + lane] += a * R[odux + lane];
            syncwarp();
            // ---- inequality row pairs: lam, t += a (dlam, dt), clipped; rd; mu; Gamma, gamma -> the sums and
            //      differences the factorisation and the stationarity residual need
#pragma unroll 1
            for (int jj = lane; jj < ncq; jj += 32)
            {
                double Gs = 0.0, gd = 0.0, dl = 0.0;
                if (pair_active(cls, jj))
                {
                    double l0 = R[olam + jj] + a * R[odlam + jj], l1 = R[olam + ncq + jj] + a * R[odlam + ncq + jj];
                    double t0 = R[ot + jj] + a * R[odt + jj], t1 = R[ot + ncq + jj] + a * R[odt + ncq + jj];
                    l0 = l0 <= lam_min ? lam_min : l0; l1 = l1 <= lam_min ? lam_min : l1;
                    t0 = t0 <= t_min ? t_min : t0; t1 = t1 <= t_min ? t_min : t1;
                    R[olam + jj] = l0; R[olam + ncq + jj] = l1; R[ot + jj] = t0; R[ot + ncq + jj] = t1;
                    const int var = srvar[jj];
                    const double v = var >= 0 ? R[oux + var]
                                              : R[ogxy + jj - nbq] * R[oux + HXV] + R[ogxy + K + jj - nbq] * R[oux + HYV];
                    const double rd0 = R[od + jj] + t0 - v, rd1 = R[od + ncq + jj] + t1 + v;
                    R[ord + jj] = rd0; R[ord + ncq + jj] = rd1;
                    const double m0 = l0 * t0, m1 = l1 * t1;
                    musum += m0; musum += m1;
                    double q = dabs(m0); n3 = q > n3 ? q : n3; q = dabs(m1); n3 = q > n3 ? q : n3;
                    q = dabs(rd0); n2 = q > n2 ? q : n2; q = dabs(rd1); n2 = q > n2 ? q : n2;
                    const double ti0 = 1.0 / t0, ti1 = 1.0 / t1;
                    R[oti + jj] = ti0; R[oti + ncq + jj] = ti1;  // kept for sweeps B, C, D
                    Gs = ti0 * l0 + ti1 * l1;
                    gd = ti0 * ((m0 - tau) - l0 * rd0) - ti1 * ((m1 - tau) - l1 * rd1);
                    dl = l1 - l0;
                }
                sGs[jj] = Gs; sgd[jj] = gd; sdl[jj] = dl;
            }
            syncwarp();
            // ---- residuals rg (lanes < NV), rb (lanes NV..NV+NX-1)
            {
                const double r = res_gb(R, k, cls, oux, opi, opip, orq, ob, Rp + oux + NU);
                const double q = dabs(r);
                if (lane < NV) { R[org + lane] = r; n0 = q > n0 ? q : n0; }
                else if (lane < NV + NX && k < N) { R[orb + lane - NV] = r; n1 = q > n1 ? q : n1; }
            }
            // ---- matrix to factorise, in place in R.L: template (H + reg, identity rows of inactive variables) ...
            double* Mx = R + oL;
            const double* T = Tp + cls * NE;
#pragma unroll 1
            for (int e = lane; e < NE; e += 32) Mx[sent[e] & 0xffff] = T[e];
            syncwarp();
            // ... + box terms on the diagonal, gradient row = rg + box gamma differences
            if (lane < NV)
            {
                const int row = box_of_var(cls, lane);
                double gr = R[org + lane];
                if (row >= 0) { Mx[lane * NV + lane] += sGs[row]; gr += sgd[row]; }
                Mx[NV * NV + lane] = gr;
            }
            if (k < N)
            {
                // AL = [B'; A'; res_b'] * Lxx_{k+1}  (dtrmm_rlnn), one entry (r, j) per lane and round
                const double* Lx = Rp + oL + NU * NV + NU;  // xx block of the factor of stage k+1, row pitch NV
#pragma unroll 1
                for (int e = lane; e < NR * NX; e += 32)
                {
                    const int r = e / NX, j = e - r * NX;
                    const double* pa = r < NV ? R + oBAt + r : R + orb;
                    const int sa = r < NV ? NV : 1;
                    double acc = 0.0;
#pragma unroll
                    for (int m = 0; m < NX; m++) if (m >= j) acc += pa[m * sa] * Lx[m * NV + j];
                    sAL[e] = acc;
                }
                syncwarp();
                // Pb = Lxx (Lxx' res_b) ; then the last row of AL gets l_{k+1,x} added
                double pb = 0.0;
                if (lane < NX)
                {
#pragma unroll
                    for (int j = 0; j < NX; j++) if (j <= lane) pb += Lx[lane * NV + j] * sAL[NV * NX + j];
                    R[oPb + lane] = pb;
                }
                syncwarp();
                if (lane < NX) sAL[NV * NX + lane] += Rp[oL + NV * NV + NU + lane];
                syncwarp();
                // syrk: M += AL AL' on the lower trapezoid
#pragma unroll 1
                for (int e = lane; e < NE; e += 32)
                {
                    const int rc = sent[e], r = rc >> 24, c = (rc >> 16) & 0xff;
                    double acc = 0.0;
#pragma unroll
                    for (int m = 0; m < NX; m++) acc += sAL[r * NX + m] * sAL[c * NX + m];
                    Mx[rc & 0xffff] += acc;
                }
                // h rows: D diag(Gamma_l + Gamma_u) D' touches only the (X,Y) block; gradient row gets D (gamma_l - gamma_u)
                double oacc = 0.0;
                int ooff = -1;
                if (cls == 1 && lane < 5 && K > 0)
                {
                    const double* gX = R + ogxy;
                    const double* gY = gX + K;
                    const double* pw = (lane < 3 ? sGs : sgd) + nbq;
                    const double* pu = lane == 0 ? gX : gY;
                    const double* pv = (lane == 0 || lane == 1 || lane == 3) ? gX : gY;
                    if (lane < 3)
                    {
#pragma unroll 1
                        for (int c = 0; c < K; c++) oacc += (pu[c] * pw[c]) * pv[c];
                    }
                    else
                    {
#pragma unroll 1
                        for (int c = 0; c < K; c++) oacc += pw[c] * pv[c];
                    }
                    ooff = (lane < 3 ? (lane == 0 ? HXV : HYV) : NV) * NV + ((lane == 0 || lane == 1 || lane == 3) ? HXV : HYV);
                }
                syncwarp();
                if (ooff >= 0) Mx[ooff] += oacc;
            }
            syncwarp();
            // ---- (NV+1) x NV Cholesky, lower; non-positive pivot => zero column
            // (dpotrf_l_mn pivot rule, BF/kernel/generic/kernel_dgemm_4x4_lib4.c:5701-5710).  Lane r <= NV takes row r
            // into registers; pivots and multipliers travel by shuffle.
            {
                const int r = lane <= NV ? lane : NV;
                double Mr[NV], dinv = 0.0;
#pragma unroll
                for (int c = 0; c < NV; c++) Mr[c] = Mx[r * NV + c];
#pragma unroll
                for (int j = 0; j < NV; j++)
                {
                    const double piv = shfl(Mr[j], j);
                    double sq = 0.0, inv = 0.0;
                    if (piv > 0.0) { inv = drsqrt(piv); sq = piv * inv; }
                    if (r == j) { Mr[j] = sq; if (j < NU) dinv = inv; } else Mr[j] *= inv;
#pragma unroll
                    for (int c = j + 1; c < NV; c++)
                    {
                        const double lc = shfl(Mr[j], c);
                        Mr[c] -= Mr[j] * lc;
                    }
                }
                if (lane <= NV)
                {
#pragma unroll
                    for (int c = 0; c < NV; c++) if (c <= lane) Mx[lane * NV + c] = Mr[c];
                    if (lane < NU) Mx[lane * NV + NV - 1] = dinv;  // 1/L[j][j] of the columns solved per stage (unused upper slot)
                }
            }
            syncwarp();
            if (lane < NV) R[odux + lane] = Mx[NV * NV + lane];  // backward vector of the forward substitution
            syncwarp();
            rec_store2(k, ir, oux, odux + svv, olam, ormc);
            { const int t = ip; ip = ir; ir = in; in = t; }
        }
        sweep_end();
        n4[0] = warp_max(n0); n4[1] = warp_max(n1); n4[2] = warp_max(n2); n4[3] = warp_max(n3);
        mu = warp_sum(musum) / nct;
        syncwarp();
    }

    // Sweeps B / D (forward, k = 0..N): forward substitution (x_ocp_qp_kkt.c:537-575 | 1243-1290), then dt, dlam
    // (:748-764 + COMPUTE_LAM_T_QP, x_core_qp_ipm_aux.c:117-142), step length (COMPUTE_ALPHA_QP :146-216) and the
    // sums COMPUTE_MU_AFF_QP (:329-357) needs.  corr = false: affine step after sweep A (dux holds row NV of the
    // factor, res_m = lam*t - tau).  corr = true: corrector / centering step after sweep C (dux holds the backward
    // vector, res_m = rmc) fused with the residual norms of the linear system (OCP_QP_RES_COMPUTE_LIN,
    // x_ocp_qp_res.c:468-633) that decide on iterative refinement.
    MDEV void sweepF(bool corr, double tau, double* nlin)
    {
        double bdn = 1.0, bdd = -1.0, bpn = 1.0, bpd = -1.0, s1 = 0.0, s2 = 0.0;  // best dual / primal ratio = -1
        double l0 = 0, l1 = 0, l2 = 0, l3 = 0;
        int ir = 0, in = 1, ip = 2;
        sweep_begin();
        rec_fetch(0, ir);
        rec_wait(ir);
        if (N >= 1) rec_fetch(1, in);
#pragma unroll 1
        for (int k = 0; k <= N; k++)
        {
            double *R = buf[ir], *Rn = buf[in];
            const int cls = k == 0 ? 0 : (k < N ? 1 : 2);
            if (k == 0)
            {
                rec_mask_stage0(R);
                if (lane < NX) R[odux + NU + lane] = 0.0;  // no x step at stage 0
                syncwarp();
            }
            const double* L = R + oL;
            // ---- columns solved at this stage (dtrsv_ltn): every lane redundantly; x part of dux is already there
            double zu[NU];
#pragma unroll
            for (int i = NU - 1; i >= 0; i--)
            {
                double ax = 0.0, au = -R[odux + i];
#pragma unroll
                for (int m = NU; m < NV; m++) ax -= L[m * NV + i] * R[odux + m];
#pragma unroll
                for (int m = i + 1; m < NU; m++) au -= L[m * NV + i] * zu[m];
                zu[i] = (au + ax) * L[i * NV + NV - 1];
                if (cls == 2) zu[i] = 0.0;
            }
            // ---- dx_{k+1} = res_b + [B A] dux_k
            double x1 = 0.0;
            if (k < N && lane < NX)
            {
                double ax = R[orb + lane], au = 0.0;
#pragma unroll
                for (int i = NU; i < NV; i++) ax += R[oBAt + i + NV * lane] * R[odux + i];
#pragma unroll
                for (int i = 0; i < NU; i++) au += R[oBAt + i + NV * lane] * zu[i];
                x1 = ax + au;
            }
            syncwarp();  // every lane has read the backward vector
            if (lane == 0)
            {
#pragma unroll
                for (int i = 0; i < NU; i++) R[odux + i] = zu[i];
            }
            syncwarp();
            // ---- dt, dlam, step length on the row pairs
#pragma unroll 1
            for (int jj = lane; jj < ncq; jj += 32)
            {
                double dld = 0.0;
                if (pair_active(cls, jj))
                {
                    const int var = srvar[jj];
                    const double dv = var >= 0 ? R[odux + var]
                                               : R[ogxy + jj - nbq] * R[odux + HXV] + R[ogxy + K + jj - nbq] * R[odux + HYV];
#pragma unroll
                    for (int side = 0; side < 2; side++)
                    {
                        const int r = jj + side * ncq;
                        double dtr = side ? -dv : dv;
                        const double lam0 = R[olam + r], t0 = R[ot + r], rd = R[ord + r], tinv = R[oti + r];
                        const double m = corr ? R[ormc + r] : lam0 * t0 - tau;
                        const double dlr = -tinv * (m + (lam0 * dtr) - (lam0 * rd));
                        dtr -= rd;
                        if (corr)
                        {
                            // residual of the linearised rows at the step: rd + dt -+ v, rm + lam*dt + dlam*t
                            const double e2 = dabs(side ? rd + dtr + dv : rd + dtr - dv), e3 = dabs(m + lam0 * dtr + dlr * t0);
                            l2 = e2 > l2 ? e2 : l2; l3 = e3 > l3 ? e3 : l3;
                        }
                        R[odlam + r] = dlr; R[odt + r] = dtr;
                        // COMPUTE_ALPHA_QP keeps the ratio closest to zero among rows with a negative step; the running
                        // best ratio is kept as (numerator, denominator) and compared by cross-multiplication, so the
                        // divisions happen once per sweep instead of once per row
                        if (dlr < 0.0 && bdn * dlr < lam0 * bdd) { bdn = lam0; bdd = dlr; }
                        if (dtr < 0.0 && bpn * dtr < t0 * bpd) { bpn = t0; bpd = dtr; }
                        s1 += lam0 * dtr + t0 * dlr;
                        s2 += dlr * dtr;
                        dld = side ? dlr - dld : dlr;
                    }
                }
                sdl[jj] = dld;  // dlam_upper - dlam_lower, for the residual of the linear system
            }
            // ---- dpi_k = Lxx (Lxx' dx_{k+1} + l_x)  |  p_{k+1} + Lxx (Lxx' dx_{k+1}) : needs the factor of stage k+1
            if (k < N)
            {
                rec_wait(in);  // record k+1 has landed
                const double* Lx = Rn + oL + NU * NV + NU;
                double pj = 0.0;
                if (lane < NX) pj = Rn[odux + NU + lane];  // l_x or p of stage k+1
                syncwarp();
                if (lane < NX) Rn[odux + NU + lane] = x1;  // from now on the x part of dux_{k+1}
                syncwarp();
                if (lane < NX)
                {
                    double acc = 0.0;
#pragma unroll
                    for (int m = 0; m < NX; m++) if (m >= lane) acc += Lx[m * NV + lane] * Rn[odux + NU + m];
                    sz[lane] = corr ? acc : acc + pj;
                }
                syncwarp();
                if (lane < NX)
                {
                    double acc = 0.0;
#pragma unroll
                    for (int j = 0; j < NX; j++) if (j <= lane) acc += Lx[lane * NV + j] * sz[j];
                    const double dp = corr ? pj + acc : acc;
                    R[odpi + lane] = dp;
                    Rn[odpip + lane] = dp;
                }
            }
            syncwarp();
            if (corr)
            {
                const double r = res_gb(R, k, cls, odux, odpi, odpip, org, orb, Rn + odux + NU);
                const double q = dabs(r);
                if (lane < NV) l0 = q > l0 ? q : l0;
                else if (lane < NV + NX && k < N) l1 = q > l1 ? q : l1;
            }
            rec_store2(k, ir, odux, orq, odlam, od);
            { const int t = ip; ip = ir; ir = in; in = t; }
            if (k + 2 <= N) { rec_reuse_guard(); rec_fetch(k + 2, in); }
        }
        sweep_end();
        const double a_prim = warp_max(bpn / bpd), a_dual = warp_max(bdn / bdd);
        alpha = -(a_prim > a_dual ? a_prim : a_dual);
        S1 = warp_sum(s1); S2 = warp_sum(s2);
        if (corr) { nlin[0] = warp_max(l0); nlin[1] = warp_max(l1); nlin[2] = warp_max(l2); nlin[3] = warp_max(l3); }
        syncwarp();
    }

    // Sweep C (backward, k = N..0): backward part of OCP_QP_SOLVE_KKT_STEP (x_ocp_qp_kkt.c:1096-1242) with
    // COMPUTE_GAMMA_QP (x_core_qp_ipm_aux.c:89-113) for the right-hand side
    //   res_m = lam*t + dt_aff*dlam_aff - sigma_mu (corrector)   |   lam*t - sigma_mu (centering only)
    // (x_ocp_qp_ipm.c:2138-2160, 2175-2200), which is stored in rmc for sweep D.
    MDEV void sweepC(bool with_aff, double sigma_mu)
    {
        int ir = 0, in = 1, ip = 2;
        sweep_begin();
        rec_fetch(N, ir);
#pragma unroll 1
        for (int k = N; k >= 0; k--)
        {
            double *R = buf[ir], *Rp = buf[ip];
            rec_wait(ir);
            if (k > 0) { rec_reuse_guard(); rec_fetch(k - 1, in); }
            const int cls = k == 0 ? 0 : (k < N ? 1 : 2);
            if (k == 0) rec_mask_stage0(R);
#pragma unroll 1
            for (int jj = lane; jj < ncq; jj += 32)
            {
                double gd = 0.0, m0 = 0.0, m1 = 0.0;
                if (pair_active(cls, jj))
                {
                    const double la0 = R[olam + jj], la1 = R[olam + ncq + jj], t0 = R[ot + jj], t1 = R[ot + ncq + jj];
                    m0 = la0 * t0; m1 = la1 * t1;
                    if (with_aff) { m0 += R[odt + jj] * R[odlam + jj]; m1 += R[odt + ncq + jj] * R[odlam + ncq + jj]; }
                    m0 -= sigma_mu; m1 -= sigma_mu;
                    gd = R[oti + jj] * (m0 - la0 * R[ord + jj]) - R[oti + ncq + jj] * (m1 - la1 * R[ord + ncq + jj]);
                }
                R[ormc + jj] = m0; R[ormc + ncq + jj] = m1;
                sgd[jj] = gd;
            }
            if (k < N && lane < NX) sz[lane] = Rp[odux + NU + lane] + R[oPb + lane];  // p_{k+1} + Pb
            syncwarp();
            const int i = lane;
            double zi = 0.0;
            if (i < NV && (cls == 1 || (cls == 0 ? i < NU : i >= NU)))
            {
                zi = R[org + i];
                const int row = box_of_var(cls, i);
                if (row >= 0) zi += sgd[row];
                if (k < N)
                {
                    if (i == HXV || i == HYV)
                    {
                        const double* gq = R + ogxy + (i == HXV ? 0 : K);
#pragma unroll 1
                        for (int c = 0; c < K; c++) zi += gq[c] * sgd[nbq + c];
                    }
                    double acc = 0.0;
#pragma unroll
                    for (int j = 0; j < NX; j++) acc += R[oBAt + i + NV * j] * sz[j];
                    zi += acc;
                }
            }
            // dtrsv_lnn_mn(nv, nu): forward-eliminate the columns solved at this stage
            const double* L = R + oL;
#pragma unroll
            for (int m = 0; m < NU; m++)
            {
                if (i == m) zi = zi * L[m * NV + NV - 1];
                const double bm = shfl(zi, m);
                if (i > m && i < NV) zi -= L[i * NV + m] * bm;
            }
            if (i < NV) R[odux + i] = zi;
            syncwarp();
            rec_store2(k, ir, odux, odux + svv, ormc, odlam);
            { const int t = ip; ip = ir; ir = in; in = t; }
        }
        sweep_end();
        solve_calls++;
        syncwarp();
    }

    MDEV bool itref_ok(const double* n) const
    {
        return (n[0] < tol_stat || n[0] < 1e-3 * res_max[0]) && (n[1] < tol_eq || n[1] < 1e-3 * res_max[1]) &&
               (n[2] < tol_ineq || n[2] < 1e-3 * res_max[2]) && (n[3] < tol_comp || n[3] < 1e-3 * res_max[3]);
    }

    // OCP_QP_IPM_SOLVE + OCP_QP_IPM_DELTA_STEP: HP/ocp_qp/x_ocp_qp_ipm.c:2354-2683, 1888-2350 (pred_corr,
    // cond_pred_corr, itref_corr_max = 2); returns HPIPM status 0 ok / 1 max iter / 2 min step / 3 NaN.
    // Per iteration: A (update with the previous step, residuals, factorisation) -> B (affine step) -> C, D (corrector)
    // [-> C, D centering only] [-> refinement, rare].  Each sweep has ONE call site so that its loop exists once in
    // the instruction stream.
    MDEV int ipm_solve(int* iters)
    {
        const double tau_min = 1e-16, alpha_min = 1e-8, reg_prim = 1e-15;
        ipm_init();
        alpha = 1.0;
        double a = 0.0;  // the first sweep A applies no step
        int kk = 0;
        for (;;)
        {
            sweepA(a, tau_min, reg_prim, res_max);
            if (!(kk < iter_max && alpha > alpha_min &&
                  (res_max[0] > tol_stat || res_max[1] > tol_eq || res_max[2] > tol_ineq ||
                   dabs(res_max[3] - tau_min) > tol_comp)))
                break;
            double nlin[4] = {0, 0, 0, 0};
            double sigma_mu = 0.0, mu_aff0 = 0.0;
            for (int pass = 0; pass < 3; pass++)
            {
                // pass 0: affine step; pass 1: corrector; pass 2: centering only (conditional)
                if (pass > 0) sweepC(pass == 1, sigma_mu);
                sweepF(pass > 0, tau_min, nlin);
                const double ma = (mu * nct + alpha * S1 + alpha * alpha * S2) / nct;  // COMPUTE_MU_AFF_QP
                if (pass == 0)
                {
                    mu_aff = ma;
                    const double tmp = mu_aff / mu;
                    sigma = tmp * tmp * tmp;
                    sigma_mu = sigma * mu;
                    sigma_mu = sigma_mu > tau_min ? sigma_mu : tau_min;
                }
                else if (pass == 1)
                {
                    mu_aff0 = mu_aff;
                    mu_aff = ma;
                    if (!(mu_aff > 2.0 * mu_aff0)) break;
                }
            }
            bool refined = false;
            for (int it = 0; it < 2; it++)
            {
                if (itref_ok(nlin)) break;
                // rare path (a fraction of a percent of the iterations): iterative refinement on the stage-major
                // scratch arrays, straight from HBM
                res_pass<true>(nlin);
                solve_sweep(true);
                forward_sweep(Y.rb2, Y.dux2, Y.dpi2, false);
                expand_pass(2, 0.0);
                add_refinement();
                refined = true;
                res_pass<true>(nlin);
            }
            if (refined) alpha_pass();
            a = alpha;
            if (a < 1.0) a = a * ((1.0 - a) * 0.99 + a * 0.9999999);
            kk++;
        }
        *iters = kk;
        if (kk == iter_max) return 1;
        if (alpha <= alpha_min) return 2;
        if (disnan(mu)) return 3;
        return 0;
    }

    // ---------------------------------------------------------------- IPM: rare path (iterative refinement), straight from HBM
    // QP residuals.  LIN = false: OCP_QP_RES_COMPUTE (HP/ocp_qp/x_ocp_qp_res.c:336-466) at the iterate (ux,pi,lam,t)
    // -> (rg,rb,rd), norms into out4, mu.  LIN = true: OCP_QP_RES_COMPUTE_LIN (:468-592): residual of the Newton
    // system with right-hand side (rg,rb,rd,rmc) at the step (dux,dpi,dlam,dt) -> (rg2,rb2,rd2,rm2).
    template <bool LIN>
    MDEV void res_pass(double* out4)
    {
        double n0 = 0, n1 = 0, n2 = 0, n3 = 0, musum = 0;
        const Field &fv = LIN ? Y.dux : Y.ux, &fp = LIN ? Y.dpi : Y.pi, &fl = LIN ? Y.dlam : Y.lam, &ft = LIN ? Y.dt : Y.t;
        const Field &og = LIN ? Y.rg2 : Y.rg, &ob = LIN ? Y.rb2 : Y.rb, &od = LIN ? Y.rd2 : Y.rd;
        for (int k = lane; k <= N; k += 32)
        {
            const double *v = F(fv, k), *pk = F(fp, k), *l = F(fl, k), *tt = F(ft, k);
            const double *H = Hk(k), *cg = LIN ? F(Y.rg, k) : F(Y.rq, k), *cd = LIN ? F(Y.rd, k) : F(Y.d, k);
            double *rg = F(og, k), *rd = F(od, k);
            double g[NV];
#pragma unroll
            for (int i = 0; i < NV; i++)
            {
                double acc = 0.0;
#pragma unroll
                for (int j = 0; j < NV; j++) acc += H[i + NV * j] * v[j];
                g[i] = acc + cg[i];
            }
            if (k > 0)
            {
                const double* pm = F(fp, k - 1);
#pragma unroll
                for (int i = 0; i < NX; i++) g[NU + i] -= pm[i];
            }
            if (k < N)
            {
                for (int j = 0; j < nbq; j++)
                {
                    if (!row_active(k, j)) continue;
                    const int id = j < nbu ? j : NU + P.idxbx[j - nbu];
                    const double dl = l[ncq + j] - l[j], vv = v[id];
#pragma unroll
                    for (int i = 0; i < NV; i++) if (i == id) g[i] += dl;
                    rd[j] = cd[j] + tt[j] - vv;
                    rd[ncq + j] = cd[ncq + j] + tt[ncq + j] + vv;
                }
                const double* gxy = F(Y.gxy, k);
                for (int c = 0; c < K; c++)
                {
                    const int r = nbq + c;
                    const double gX = k >= 1 ? gxy[c] : 0.0, gY = k >= 1 ? gxy[K + c] : 0.0;
                    const double dl = l[ncq + r] - l[r];
                    g[HXV] += gX * dl; g[HYV] += gY * dl;
                    const double vv = gX * v[HXV] + gY * v[HYV];
                    rd[r] = cd[r] + tt[r] - vv;
                    rd[ncq + r] = cd[ncq + r] + tt[ncq + r] + vv;
                }
                const double* BAt = F(Y.BAt, k);
                const double* vn = F(fv, k + 1);
                const double* cb = LIN ? F(Y.rb, k) : F(Y.b, k);
                double* rb = F(ob, k);
#pragma unroll
                for (int j = 0; j < NX; j++)
                {
                    double acc = cb[j] - vn[NU + j];
#pragma unroll
                    for (int i = 0; i < NV; i++) if (var_active(k, i)) acc += BAt[i + NV * j] * v[i];
                    rb[j] = acc;
                    const double a = dabs(acc);
                    n1 = a > n1 ? a : n1;
                }
#pragma unroll
                for (int i = 0; i < NV; i++)
                {
                    double acc = 0.0;
#pragma unroll
                    for (int j = 0; j < NX; j++) acc += BAt[i + NV * j] * pk[j];
                    g[i] += acc;
                }
                if (LIN)
                {
                    const double *lam = F(Y.lam, k), *t = F(Y.t, k), *rm = F(Y.rmc, k);
                    double* rm2 = F(Y.rm2, k);
                    for (int j = 0; j < 2 * ncq; j++)
                    {
                        if (!row_active(k, j % ncq)) { rm2[j] = 0.0; continue; }
                        const double m = rm[j] + lam[j] * tt[j] + l[j] * t[j];
                        rm2[j] = m;
                        const double a = dabs(m), b = dabs(rd[j]);
                        n3 = a > n3 ? a : n3; n2 = b > n2 ? b : n2;
                    }
                }
                else
                {
                    for (int j = 0; j < 2 * ncq; j++)
                    {
                        if (!row_active(k, j % ncq)) continue;
                        const double m = l[j] * tt[j];
                        musum += m;
                        const double a = dabs(m), b = dabs(rd[j]);
                        n3 = a > n3 ? a : n3; n2 = b > n2 ? b : n2;
                    }
                }
            }
#pragma unroll
            for (int i = 0; i < NV; i++)
            {
                const double gi = var_active(k, i) ? g[i] : 0.0;
                rg[i] = gi;
                const double a = dabs(gi);
                n0 = a > n0 ? a : n0;
            }
        }
        out4[0] = warp_max(n0); out4[1] = warp_max(n1); out4[2] = warp_max(n2); out4[3] = warp_max(n3);
        if (!LIN) mu = warp_sum(musum) / nct;
        syncwarp();
    }

    // dt, dlam from dux (tail of HP/ocp_qp/x_ocp_qp_kkt.c:748-764 + COMPUTE_LAM_T_QP, HP/ipm_core/x_core_qp_ipm_aux.c:117-142),
    // fused with COMPUTE_ALPHA_QP (:146-216) and the sums COMPUTE_MU_AFF_QP (:329-357) needs.
    // mode 0: affine rhs (res_m = lam*t - tau_min); 1: rhs m stored in rmc; 2: refinement (rd2, rm2 -> dlam2, dt2)
    MDEV void expand_pass(int mode, double tau)
    {
        const Field &fv = mode == 2 ? Y.dux2 : Y.dux, &fl = mode == 2 ? Y.dlam2 : Y.dlam, &ft = mode == 2 ? Y.dt2 : Y.dt;
        const Field &frd = mode == 2 ? Y.rd2 : Y.rd, &frm = mode == 2 ? Y.rm2 : Y.rmc;
        double a_prim = -1.0, a_dual = -1.0, s1 = 0.0, s2 = 0.0;
        for (int k = lane; k < N; k += 32)
        {
            const double *v = F(fv, k), *lam = F(Y.lam, k), *t = F(Y.t, k), *rd = F(frd, k), *rm = F(frm, k);
            const double* gxy = F(Y.gxy, k);
            double *dl = F(fl, k), *dtt = F(ft, k);
            for (int j = 0; j < ncq; j++)
            {
                if (!row_active(k, j)) continue;
                double dv;
                if (j < nbq) dv = v[j < nbu ? j : NU + P.idxbx[j - nbu]];
                else dv = k >= 1 ? gxy[j - nbq] * v[HXV] + gxy[K + j - nbq] * v[HYV] : 0.0;
#pragma unroll
                for (int side = 0; side < 2; side++)
                {
                    const int r = j + side * ncq;
                    double dtr = side ? -dv : dv;
                    const double lam0 = lam[r], t0 = t[r], tinv = 1.0 / t0;
                    const double m = mode == 0 ? lam0 * t0 - tau : rm[r];
                    const double dlr = -tinv * (m + (lam0 * dtr) - (lam0 * rd[r]));
                    dtr -= rd[r];
                    dl[r] = dlr; dtt[r] = dtr;
                    if (a_dual * dlr > lam0) a_dual = lam0 / dlr;
                    if (a_prim * dtr > t0) a_prim = t0 / dtr;
                    s1 += lam0 * dtr + t0 * dlr;
                    s2 += dlr * dtr;
                }
            }
        }
        if (mode != 2)
        {
            a_prim = warp_max(a_prim); a_dual = warp_max(a_dual);
            alpha = -(a_prim > a_dual ? a_prim : a_dual);
            S1 = warp_sum(s1); S2 = warp_sum(s2);
        }
        syncwarp();
    }

    // COMPUTE_ALPHA_QP on the current step (after iterative refinement changed it)
    MDEV void alpha_pass()
    {
        double a_prim = -1.0, a_dual = -1.0;
        for (int k = lane; k < N; k += 32)
        {
            const double *lam = F(Y.lam, k), *t = F(Y.t, k), *dl = F(Y.dlam, k), *dtt = F(Y.dt, k);
            for (int r = 0; r < 2 * ncq; r++)
            {
                if (!row_active(k, r % ncq)) continue;
                if (a_dual * dl[r] > lam[r]) a_dual = lam[r] / dl[r];
                if (a_prim * dtt[r] > t[r]) a_prim = t[r] / dtt[r];
            }
        }
        a_prim = warp_max(a_prim); a_dual = warp_max(a_dual);
        alpha = -(a_prim > a_dual ? a_prim : a_dual);
    }

    // step += refinement step
    MDEV void add_refinement()
    {
        for (int k = lane; k <= N; k += 32)
        {
            double *a = F(Y.dux, k); const double* b = F(Y.dux2, k);
            for (int i = 0; i < NV; i++) a[i] += b[i];
            if (k < N)
            {
                double* c = F(Y.dpi, k); const double* e = F(Y.dpi2, k);
                double* cn = F(Y.dpi_prev, k + 1);
                for (int i = 0; i < NX; i++) { c[i] += e[i]; cn[i] = c[i]; }
                double *l = F(Y.dlam, k), *t = F(Y.dt, k);
                const double *l2 = F(Y.dlam2, k), *t2 = F(Y.dt2, k);
                for (int r = 0; r < 2 * ncq; r++)
                    if (row_active(k, r % ncq)) { l[r] += l2[r]; t[r] += t2[r]; }
            }
        }
        syncwarp();
    }

    // ---------------------------------------------------------------- IPM: the Riccati sweeps (serial in k)
    MDEV void load_BA(int k)
    {
        if (k < N)
        {
            const double* g = F(Y.BAt, k);
            for (int e = lane; e < NV * NX; e += 32) sBA[e] = var_active(k, e % NV) ? g[e] : 0.0;
        }
        else
            for (int e = lane; e < NV * NX; e += 32) sBA[e] = 0.0;
    }

    MDEV void load_gxy(int k)
    {
        if (k < N)
        {
            const double* g = F(Y.gxy, k);
            for (int e = lane; e < 2 * K; e += 32) sgxy[e] = k >= 1 ? g[e] : 0.0;
        }
    }

    MDEV void load_Lnext(int k1)  // xx block and nothing else of the factor of stage k1
    {
        const double* L = F(Y.L, k1);
        for (int e = lane; e < NX * NX; e += 32)
        {
            const int m = e / NX, j = e % NX;
            sLn[e] = L[(NU + m) * NV + NU + j];
        }
    }

    // forward substitution shared by both solves.  On entry dux[k] holds the backward vector; scaled: its x part is
    // l_{k,x} (row NV of the factor; HP/ocp_qp/x_ocp_qp_kkt.c:537-575), else p_k itself (:1243-1290).
    MDEV void forward_sweep(const Field& frb, const Field& fdux, const Field& fdpi, bool scaled)
    {
        double xc[NX];
#pragma unroll
        for (int i = 0; i < NX; i++) xc[i] = 0.0;
        for (int k = 0; k <= N; k++)
        {
            syncwarp();
            {
                const double* L = F(Y.L, k);
                for (int e = lane; e < NR * NV; e += 32) sL[e] = L[e];
                const double* q = F(fdux, k);
                if (lane < NU) sq[lane] = q[lane];
                if (k < N)
                {
                    load_BA(k);
                    load_Lnext(k + 1);
                    if (lane < NX) { sx1[lane] = F(frb, k)[lane]; sx2[lane] = F(fdux, k + 1)[NU + lane]; }
                }
            }
            syncwarp();
            double z[NV];
#pragma unroll
            for (int i = 0; i < NX; i++) z[NU + i] = xc[i];
#pragma unroll
            for (int i = NU - 1; i >= 0; i--)  // dtrsv_ltn on the columns solved at this stage
            {
                double acc = -sq[i];
#pragma unroll
                for (int m = i + 1; m < NV; m++) acc -= sL[m * NV + i] * z[m];
                z[i] = acc / sL[i * NV + i];
            }
            if (lane == 0)
            {
                double* o = F(fdux, k);
#pragma unroll
                for (int i = 0; i < NV; i++) o[i] = var_active(k, i) ? z[i] : 0.0;
            }
            if (k < N)
            {
                double x1 = 0.0;
                if (lane < NX)
                {
                    double acc = sx1[lane];
#pragma unroll
                    for (int i = 0; i < NV; i++) acc += sBA[i + NV * lane] * z[i];
                    x1 = acc;
                }
#pragma unroll
                for (int m = 0; m < NX; m++) xc[m] = shfl(x1, m);
                // dpi_k = Lxx (Lxx' dx_{k+1} + l_x)  |  p_{k+1} + Lxx (Lxx' dx_{k+1})
                double tmp = 0.0;
                if (lane < NX)
                {
                    double acc = 0.0;
                    for (int m = lane; m < NX; m++) acc += sLn[m * NX + lane] * xc[m];
                    tmp = scaled ? acc + sx2[lane] : acc;
                }
                double tv[NX];
#pragma unroll
                for (int j = 0; j < NX; j++) tv[j] = shfl(tmp, j);
                if (lane < NX)
                {
                    double acc = 0.0;
#pragma unroll
                    for (int j = 0; j < NX; j++) if (j <= lane) acc += sLn[lane * NX + j] * tv[j];
                    F(fdpi, k)[lane] = scaled ? acc : sx2[lane] + acc;
                }
            }
        }
        syncwarp();
    }

    // OCP_QP_SOLVE_KKT_STEP, backward part (HP/ocp_qp/x_ocp_qp_kkt.c:1096-1242) with COMPUTE_GAMMA_QP
    // (x_core_qp_ipm_aux.c:89-113).  refine = false: rhs (rg, rb via Pb, rd, rmc) -> dux ; true: (rg2, rb2, rd2, rm2) -> dux2.
    MDEV void solve_sweep(bool refine)
    {
        const Field &frg = refine ? Y.rg2 : Y.rg, &frb = refine ? Y.rb2 : Y.rb, &frd = refine ? Y.rd2 : Y.rd;
        const Field &frm = refine ? Y.rm2 : Y.rmc, &fo = refine ? Y.dux2 : Y.dux;
        double pn[NX];
#pragma unroll
        for (int i = 0; i < NX; i++) pn[i] = 0.0;
        for (int k = N; k >= 0; k--)
        {
            syncwarp();
            load_BA(k);
            load_gxy(k);
            {
                const double *lam = F(Y.lam, k), *t = F(Y.t, k), *rd = F(frd, k), *rm = F(frm, k);
                for (int r = lane; r < 2 * ncq; r += 32)
                {
                    double g = 0.0;
                    if (row_active(k, r % ncq)) g = (1.0 / t[r]) * (rm[r] - lam[r] * rd[r]);
                    sg[r] = g;
                }
                if (k < N)
                {
                    if (refine) { load_Lnext(k + 1); if (lane < NX) sx1[lane] = F(frb, k)[lane]; }
                    else if (lane < NX) sx1[lane] = F(Y.Pb, k)[lane];
                }
            }
            syncwarp();
            const int i = lane;
            double zi = 0.0;
            if (i < NV && var_active(k, i))
            {
                zi = F(frg, k)[i];
                const int row = vrow(k, i);
                if (row >= 0) zi += sg[row] - sg[ncq + row];
                if (k < N && (i == HXV || i == HYV))
                {
                    for (int c = 0; c < K; c++)
                    {
                        const double gd = sg[nbq + c] - sg[ncq + nbq + c];
                        zi += (i == HXV ? sgxy[c] : sgxy[K + c]) * gd;
                    }
                }
            }
            if (k < N)
            {
                double tmp[NX];
                if (!refine)
                {
#pragma unroll
                    for (int j = 0; j < NX; j++) tmp[j] = pn[j] + sx1[j];
                }
                else
                {
                    double t2[NX];
#pragma unroll
                    for (int j = 0; j < NX; j++)
                    {
                        double acc = 0.0;
#pragma unroll
                        for (int m = j; m < NX; m++) acc += sLn[m * NX + j] * sx1[m];
                        t2[j] = acc;
                    }
#pragma unroll
                    for (int ii = 0; ii < NX; ii++)
                    {
                        double acc = 0.0;
#pragma unroll
                        for (int j = 0; j <= ii; j++) acc += sLn[ii * NX + j] * t2[j];
                        tmp[ii] = acc + pn[ii];
                    }
                }
                if (i < NV)
                {
                    double acc = 0.0;
#pragma unroll
                    for (int j = 0; j < NX; j++) acc += sBA[i + NV * j] * tmp[j];
                    zi += acc;
                }
            }
            // dtrsv_lnn_mn(nv, nu): forward-eliminate the columns solved at this stage
            const double* L = F(Y.L, k);
#pragma unroll
            for (int m = 0; m < NU; m++)
            {
                if (i == m) zi = zi / L[m * NV + m];
                const double bm = shfl(zi, m);
                if (i > m && i < NV) zi -= L[i * NV + m] * bm;
            }
            if (i < NV) F(fo, k)[i] = zi;
#pragma unroll
            for (int j = 0; j < NX; j++) pn[j] = shfl(zi, NU + j);
        }
        solve_calls++;
        syncwarp();
    }


    // ---------------------------------------------------------------- after the QP
    // d_ocp_qp_restore_eq_dof (HP/ocp_qp/x_ocp_qp_red.c:723-871) + ocp_nlp_update_variables_sqp, full step
    // (AC/acados/ocp_nlp/ocp_nlp_common.c:2401-2448): ux += step; pi, lam, t <- QP values.
    MDEV void update_nlp()
    {
        for (int k = lane; k <= N; k += 32)
        {
            double *z = F(Y.zux, k), *zl = F(Y.zlam, k), *zt = F(Y.zt, k);
            const double *ux = F(Y.ux, k), *lam = F(Y.lam, k), *t = F(Y.t, k);
            if (k < N)
            {
                double* zp = F(Y.zpi, k); const double* pi = F(Y.pi, k);
                for (int i = 0; i < NX; i++) zp[i] = pi[i];
                for (int j = 0; j < nbu; j++) { zl[j] = lam[j]; zl[ncz + j] = lam[ncq + j]; zt[j] = t[j]; zt[ncz + j] = t[ncq + j]; }
                for (int c = 0; c < K; c++)
                {
                    const int a = nbu + NX + c, b = nbq + c;
                    zl[a] = lam[b]; zl[ncz + a] = lam[ncq + b]; zt[a] = t[b]; zt[ncz + a] = t[ncq + b];
                }
            }
            if (k == 0)
            {
                // recover the eliminated x0 step and the multipliers of its bounds from stationarity
                const double* zf = F(Y.zfun, 0);
                double s[NV], tmp[NV];
                for (int i = 0; i < NU; i++) s[i] = ux[i];
                for (int i = 0; i < NX; i++) s[NU + i] = zf[nbu + i];  // dx0 = x0 - x_0
                const double* rq = F(Y.rq, 0);
                const double* BAt = F(Y.BAt, 0);
                const double* pi = F(Y.pi, 0);
                const double* gxy = F(Y.gxy, 0);
                for (int i = 0; i < NX; i++)
                {
                    const int iv = NU + i;
                    double acc = rq[iv];
                    double hs = 0.0;
                    for (int j = 0; j < NV; j++) hs += Hs[iv + NV * j] * s[j];
                    acc += hs;
                    if (N > 0) for (int j = 0; j < NX; j++) acc += BAt[iv + NV * j] * pi[j];
                    if (iv == HXV || iv == HYV)
                        for (int c = 0; c < K; c++)
                        {
                            const double dl = lam[ncq + nbq + c] - lam[nbq + c];
                            acc += (iv == HXV ? gxy[c] : gxy[K + c]) * dl;
                        }
                    tmp[iv] = acc;
                }
                for (int j = 0; j < NX; j++)
                {
                    const int r = nbu + j;
                    const double v = tmp[NU + j];
                    zl[r] = 1e-16; zl[ncz + r] = 1e-16; zt[r] = 1e-16; zt[ncz + r] = 1e-16;
                    if (v >= 0) zl[r] = v; else zl[ncz + r] = -v;
                }
                for (int i = 0; i < NV; i++) z[i] += s[i];
            }
            else
            {
                for (int j = 0; j < nbx; j++)
                {
                    const int a = nbu + j;
                    if (k < N) { zl[a] = lam[a]; zl[ncz + a] = lam[ncq + a]; zt[a] = t[a]; zt[ncz + a] = t[ncq + a]; }
                }
                for (int i = 0; i < NV; i++) if (var_active(k, i)) z[i] += ux[i];
            }
        }
        syncwarp();
    }

    // residuals ocp_nlp_eval_residuals reports after an SQP_RTI step: stale linearisation, new lam / t
    // (AC/interfaces/acados_c/ocp_nlp_interface.c:909-916)
    MDEV void rti_residuals(double* res4)
    {
        double r2 = 0, r3 = 0;
        for (int k = lane; k < N; k += 32)
        {
            const double *zl = F(Y.zlam, k), *zt = F(Y.zt, k), *zf = F(Y.zfun, k);
            for (int j = 0; j < 2 * ncz; j++)
            {
                const int jj = j % ncz;
                const bool act = jj < nbu || jj >= nbu + NX || (k == 0 ? true : jj - nbu < nbx);
                if (!act) continue;
                const double a = dabs(zf[j] + zt[j]), c = dabs(zl[j] * zt[j]);
                r2 = a > r2 ? a : r2; r3 = c > r3 ? c : r3;
            }
        }
        res4[2] = warp_max(r2); res4[3] = warp_max(r3);
    }

    // ---------------------------------------------------------------- the solve
    MDEV void run(int inst)
    {
        const double* x0 = P.x0 + (long) inst * NX;
        const double* pg = P.p + (long) inst * (P.p_per_stage ? (N + 1) : 1) * 2 * K;
        const double* lhg = P.lh + (long) inst * (P.lh_per_stage ? N : 1) * K;
        const double* yrg = P.yref + (long) inst * (P.yref_per_stage ? N : 1) * NY;
        const double* yre = P.yref_e + (long) inst * NX;
        load_constants();
        rec_init();
        if (P.cold_start) cold_start(x0);
        int status = 2, sqp_iter = 0, qp_total = 0, qp_status = 0, qp_iter = 0;
        double res[4] = {0, 0, 0, 0};
        const int max_iter = P.nlp_type == 0 ? P.max_iter : 1;
        for (sqp_iter = 0; sqp_iter < max_iter; sqp_iter++)
        {
            linearize(x0, pg, lhg, yrg, yre, res);
            if (P.nlp_type == 0 && res[0] < P.tol[0] && res[1] < P.tol[1] && res[2] < P.tol[2] && res[3] < P.tol[3])
            {
                status = 0;  // ACADOS_SUCCESS, ocp_nlp_sqp.c:641-672
                break;
            }
            qp_status = ipm_solve(&qp_iter);
            qp_total += qp_iter;
            if (qp_status != 0 && qp_status != 1)
            {
                status = 4;  // ACADOS_QP_FAILURE, ocp_nlp_sqp.c:736-773
                break;
            }
            update_nlp();
            if (P.nlp_type == 1)
            {
                status = 0;  // ocp_nlp_sqp_rti.c:810-817
                rti_residuals(res);
                sqp_iter = 1;
                break;
            }
        }
        if (lane == 0)
        {
            double* st = P.stats + (long) inst * NSTAT;
            st[0] = status; st[1] = sqp_iter; st[2] = qp_total;
            st[3] = res[0]; st[4] = res[1]; st[5] = res[2]; st[6] = res[3];
            st[7] = 0; st[8] = solve_calls; st[9] = qp_status; st[10] = qp_iter; st[11] = 0;
        }
    }
};

}  // namespace usvmpc
