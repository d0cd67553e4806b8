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

// pointer to the elements of one stage of a field: element stride 1 (stage-major / record) or nsp (transposed)
struct SP {
    double* p;
    int es;
    MDEV double& operator[](int i) const { return p[(long) i * es]; }
};

// the same for the transposed arrays, whose element stride is the compile-time constant NSPC
struct TP {
    double* p;
    MDEV double& operator[](int i) const { return p[i * NSPC]; }
};

template <class M>
struct WarpSolver {
    static constexpr int NX = M::NX, NU = M::NU, NV = NX + NU, NR = NV + 1, NY = NV;
    static constexpr int HXV = NU + M::HX, HYV = NU + M::HY;
    static constexpr int NE = NV * (NV + 1) / 2 + NV;  // entries of the lower trapezoid of the (NV+1) x NV factor
    // record layout (layout.h): offsets of the fixed-size fields are compile-time constants
    static constexpr int svv = (NV + 1) / 2 * 2, sxx = (NX + 1) / 2 * 2, sLL = (NR * NV + 1) / 2 * 2;
    static constexpr int oBAt = 0, orb = (NV * NX + 1) / 2 * 2, oL = orb + sxx, oPb = oL + sLL, obv = oPb + sxx;
    static constexpr int HEAD = obv + svv;  // what the chain sweeps stream: [B';A'] rb L Pb bv
    // per-warp shared-memory scratch: fixed-size part at compile-time offsets
    static constexpr int qHs = 0, qHes = qHs + NV * NV, qWs = qHes + NV * NV, qWes = qWs + NV * NV, qTp = qWes + NX * NX;
    static constexpr int qAL = qTp + 3 * NE, qent = qAL + (NR * NX + 1) / 2 * 2, qxrow = qent + (NE + 1) / 2;
    static constexpr int qvar = qxrow + (NX + 1) / 2;

    const Params& P;
    const Layout& Y;
    int lane, N, K, nbu, nbx, ncq, ncz, nbq, nct;
    double* w;
    // shared-memory scratch of this warp
    double *Hs, *Hes, *Ws, *Wes, *Tp, *sAL;
    int *sent, *srvar, *sxrow;
    double *sBA, *sLn, *slx, *sG, *sg, *sL, *sq, *sx1, *sx2, *sgxy;  // rare path, aliased onto the record buffers
    // the two record buffers of the streaming sweeps and the offsets of the record's fields (layout.h)
    double* buf[3];
    unsigned long long* bar;  // one mbarrier per record buffer (TMA completion)
    unsigned phb;             // their phase bits
    // IPM arguments (HP/ocp_qp/x_ocp_qp_ipm.c:133-161 overridden by AC/acados/ocp_qp/ocp_qp_hpipm.c:106-116
    // and, in SQP mode, by AC/acados/ocp_nlp/ocp_nlp_sqp.c:201-227)
    double tol_stat, tol_eq, tol_ineq, tol_comp;
    int iter_max;
    // IPM state (warp-uniform)
    double res_max[4], mu, mu_aff, sigma, alpha;
    double S1, S2;  // sum(lam*dt + t*dlam), sum(dlam*dt) of the last expanded step
    double lin_d, lin_m;  // residual norms of the linearised inequality / complementarity rows at the last passF step
    int solve_calls;
#ifdef USVMPC_PROFILE
    // phase clocks of one warp (diagnostic builds only): 0 passA 1 chainA 2 chainF 3 passF 4 passC 5 chainC 6 res_pass
    // 7 refinement 8 linearize 9 update_nlp 10 ipm_init
    long long prof[20], tprev;
#define PROF(i) { const long long t_ = clock64(); prof[i] += t_ - tprev; tprev = t_; }
#else
#define PROF(i)
#endif

    MDEV WarpSolver(const Params& p, int inst, double* sm) : P(p), Y(p.lay)
    {
        lane = lane_id();
        N = P.N; K = P.K; nbu = P.nbu; nbx = P.nbx; ncq = P.ncq; ncz = P.ncz; nbq = nbu + nbx;
        nct = N >= 1 ? 2 * ((nbu + K) + (N - 1) * (nbu + nbx + K)) : 0;
        w = P.ws + (long) inst * P.ws_stride;
        double* s = sm;
        Hs = s + qHs; Hes = s + qHes; Ws = s + qWs; Wes = s + qWes; Tp = s + qTp; sAL = s + qAL;
        sent = (int*) (s + qent); sxrow = (int*) (s + qxrow);
        const int nq = NU + NX + K;
        s += qvar;
        srvar = (int*) s; s += (nq + 1) / 2;
        s = sm + ((s - sm) + 1) / 2 * 2;
        buf[0] = s; s += HEAD; buf[1] = s; s += HEAD; buf[2] = s; s += HEAD;
        bar = (unsigned long long*) s; s += 4;
        phb = 0;
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
        return y.BAt.off - ro == oBAt && y.rb.off - ro == orb && y.L.off - ro == oL && y.Pb.off - ro == oPb &&
               y.bv.off - ro == obv && y.rec_size == HEAD;
    }
    MDEV SP F(const Field& f, int k) const { return SP{w + f.off + (long) k * f.stride, f.es}; }
    MDEV TP FT(const Field& f, int k) const { return TP{w + f.off + k}; }  // transposed fields only
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
        // tables: lower-trapezoid entries e -> (r, c, offset); row pair -> boxed variable (or -1: h row); templates of
        // the matrix to factorise per stage class
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
            const SP z = F(Y.zux, k);
            for (int i = 0; i < NU; i++) z[i] = 0.0;
            for (int i = 0; i < NX; i++) z[NU + i] = x0[i];
            const SP pi = F(Y.zpi, k);
            for (int i = 0; i < NX; i++) pi[i] = 0.0;
            const SP l = F(Y.zlam, k), t = F(Y.zt, k);
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
        // columns of states that do not enter the dynamics stay unit vectors exactly (their VDE right-hand side is
        // Jx e_c = 0): no lane integrates them, the first task of the stage writes them
        constexpr int NCOL = NV - M::NKIN;
        for (int task = lane; task < N * NCOL; task += 32)
        {
            const int k = task / NCOL, col = M::NKIN + task % NCOL;
            const SP z = F(Y.zux, k);
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
            const SP BAt = F(Y.BAt, k);    // record head (chain sweeps)
            const TP BAtT = FT(Y.BAtT, k);  // transposed copy (passes)
            const int row = col < NX ? NU + col : col - NX;
#pragma unroll
            for (int i = 0; i < NX; i++) { BAt[row + NV * i] = s[i]; BAtT[row + NV * i] = s[i]; }
            if (col == M::NKIN)
            {
#pragma unroll
                for (int c = 0; c < M::NKIN; c++)
                {
#pragma unroll
                    for (int i = 0; i < NX; i++) { BAt[NU + c + NV * i] = (i == c) ? 1.0 : 0.0; BAtT[NU + c + NV * i] = (i == c) ? 1.0 : 0.0; }
                }
                const SP zn = F(Y.zux, k + 1);
                const TP b = FT(Y.b, k);
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
            const SP z = F(Y.zux, k);
            const SP zl = F(Y.zlam, k), zt = F(Y.zt, k);
            const SP zf = F(Y.zfun, k), rq = F(Y.rq, k), d = F(Y.d, k);
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
                const SP gxy = F(Y.gxy, k);
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
            const SP BAt = F(Y.BAtT, k);
            const SP pik = F(Y.zpi, k);
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
                const SP b = F(Y.b, k);
#pragma unroll
                for (int i = 0; i < NX; i++) { const double a = dabs(b[i]); r1 = a > r1 ? a : r1; }
            }
            // ---- QP gradient; stage 0: b0 += A0' dx0, r0 += S dx0 (x_ocp_qp_red.c:300-378)
#pragma unroll
            for (int i = 0; i < NV; i++) rq[i] = cg[i];
            if (k == 0 && N > 0)
            {
                const SP b = F(Y.b, 0);
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
            const TP ux = FT(Y.ux, k), pi = FT(Y.pi, k), lam = FT(Y.lam, k), t = FT(Y.t, k);
            const TP d = FT(Y.d, k);
            for (int i = 0; i < NV; i++) ux[i] = 0.0;
            for (int i = 0; i < NX; i++) pi[i] = 0.0;
            for (int j = 0; j < 2 * ncq; j++) { lam[j] = 0.0; t[j] = 1.0; }
            {
                // the first passA applies a zero step: the step starts at zero
                const TP a = FT(Y.dux, k), b = FT(Y.dpi, k), c = FT(Y.dlam, k), e = FT(Y.dt, k);
                for (int i = 0; i < NV; i++) a[i] = 0.0;
                for (int i = 0; i < NX; i++) b[i] = 0.0;
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
            const TP gxy = FT(Y.gxy, k);
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

    // ---------------------------------------------------------------- IPM: one iteration = lean chain sweeps + stage-parallel passes
    // Only the Riccati recursions are serial in the stage index.  They run as CHAIN sweeps: the warp streams the head of
    // each stage record ([B';A'], res_b, the matrix to factorise / its factor, Pb, the backward vector) through shared
    // memory (TMA for chainA, cp.async for the read-only chainF / chainC) and does nothing but the recursion.
    // Everything else of an IPM iteration (variable update, residuals, Gamma/gamma, assembly of the matrix to factorise,
    // dt/dlam, step length, dpi, right-hand sides) is independent per stage and runs as PASSES with one lane per
    // stage, straight on the records in HBM/L2.
    //
    //   passA  : UPDATE_VAR_QP + OCP_QP_RES_COMPUTE + COMPUTE_GAMMA_GAMMA_QP + assembly of H + Gamma terms (+ gradient row)
    //   chainA : backward: AL = [B';A';b'] Lxx, Pb, syrk, Cholesky                       (x_ocp_qp_kkt.c:455-535)
    //   chainF : forward : columns solved per stage, dx_{k+1} = b + [B A] dux             (x_ocp_qp_kkt.c:537-575 | 1243-1290)
    //   passF  : dt, dlam, step length, mu_aff sums, dpi [, residual of the linear system] (x_ocp_qp_kkt.c:748-764, x_core_qp_ipm_aux.c:117-216)
    //   passC  : corrector / centering right-hand side, COMPUTE_GAMMA_QP, constraint part of the backward vector
    //   chainC : backward: z = z0 + [B';A'] (p_{k+1} + Pb), eliminate the columns solved per stage (x_ocp_qp_kkt.c:1096-1242)
    MDEV int stage_class(int k) const { return k == 0 ? 0 : (k < N ? 1 : 2); }

    MDEV void passA(double a, double tau, double* n4)
    {
        const double lam_min = 1e-16, t_min = 1e-16;
        // ---- UPDATE_VAR_QP: HP/ipm_core/x_core_qp_ipm_aux.c:220-325 (split_step = 0).  A load that follows a store to a
        // possibly aliasing address cannot be hoisted, so an element-at-a-time loop pays one L2 latency per element: the
        // loads of a chunk are issued together (CBAR), then the arithmetic, then the stores.
#pragma unroll 1
        for (int k = lane; k <= N; k += 32)
        {
            const TP ux = FT(Y.ux, k), dux = FT(Y.dux, k), pi = FT(Y.pi, k), dpi = FT(Y.dpi, k);
            {
                double z[NV], dz[NV], y[NX], dy[NX];
#pragma unroll
                for (int i = 0; i < NV; i++) { z[i] = ux[i]; dz[i] = dux[i]; }
#pragma unroll
                for (int i = 0; i < NX; i++) { y[i] = pi[i]; dy[i] = dpi[i]; }  // stage N: pads of the arrays, not stored
                CBAR();
#pragma unroll
                for (int i = 0; i < NV; i++) ux[i] = z[i] + a * dz[i];
                if (k < N)
                {
#pragma unroll
                    for (int i = 0; i < NX; i++) pi[i] = y[i] + a * dy[i];
                }
            }
            if (k < N)
            {
                const TP l = FT(Y.lam, k), t = FT(Y.t, k);
                const TP dl = FT(Y.dlam, k), dtt = FT(Y.dt, k);
                constexpr int UC = 10;
                const int ne = 2 * ncq;
#pragma unroll 1
                for (int r0 = 0; r0 < ne; r0 += UC)
                {
                    double lv[UC], dv[UC], tv[UC], ev[UC];
#pragma unroll
                    for (int i = 0; i < UC; i++)
                    {
                        const int r = r0 + i < ne ? r0 + i : ne - 1;
                        lv[i] = l[r]; dv[i] = dl[r]; tv[i] = t[r]; ev[i] = dtt[r];
                    }
                    CBAR();
#pragma unroll
                    for (int i = 0; i < UC; i++)
                    {
                        const int r = r0 + i;
                        if (r >= ne || !row_active(k, r < ncq ? r : r - ncq)) continue;
                        double x = lv[i] + a * dv[i];
                        l[r] = x <= lam_min ? lam_min : x;
                        x = tv[i] + a * ev[i];
                        t[r] = x <= t_min ? t_min : x;
                    }
                }
            }
        }
        syncwarp();
        // ---- OCP_QP_RES_COMPUTE (x_ocp_qp_res.c:336-466) + Gamma, gamma for res_m = lam*t - tau + matrix to factorise
        double n0 = 0, n1 = 0, n2 = 0, n3 = 0, musum = 0;
#pragma unroll 1
        for (int k = lane; k <= N; k += 32)
        {
            const int cls = stage_class(k);
            const TP v = FT(Y.ux, k), rq = FT(Y.rq, k);
            const double* __restrict__ H = Hk(k);
            const SP Lk = F(Y.L, k);
            const double* __restrict__ T = Tp + cls * NE;
            {
                // four entries at a time, loads before stores (a load cannot be hoisted above a store that may alias)
                int e = 0;
#pragma unroll 1
                for (; e + 4 <= NE; e += 4)
                {
                    const int o0 = sent[e] & 0xffff, o1 = sent[e + 1] & 0xffff, o2 = sent[e + 2] & 0xffff, o3 = sent[e + 3] & 0xffff;
                    const double t0 = T[e], t1 = T[e + 1], t2 = T[e + 2], t3 = T[e + 3];
                    Lk[o0] = t0; Lk[o1] = t1; Lk[o2] = t2; Lk[o3] = t3;
                }
#pragma unroll 1
                for (; e < NE; e++) Lk[sent[e] & 0xffff] = T[e];
            }
            double g[NV], dg[NV], gg[NV];  // stationarity residual; additions to the diagonal / to the gradient row
#pragma unroll
            for (int i = 0; i < NV; i++)
            {
                double acc = 0.0;
#pragma unroll
                for (int j = 0; j < NV; j++) acc += H[i + NV * j] * v[j];
                g[i] = acc + rq[i];
                dg[i] = 0.0; gg[i] = 0.0;
            }
            double aYX = 0.0;
            if (k > 0)
            {
                const TP pm = FT(Y.pi, k - 1);
#pragma unroll
                for (int i = 0; i < NX; i++) g[NU + i] -= pm[i];
            }
            if (k < N)
            {
                const TP l = FT(Y.lam, k), tt = FT(Y.t, k), cd = FT(Y.d, k), pk = FT(Y.pi, k);
                const TP rd = FT(Y.rd, k), ti = FT(Y.ti, k);
#pragma unroll 1
                for (int j = 0; j < nbq; j++)
                {
                    if (!row_active(k, j)) continue;
                    const int id = srvar[j];
                    const double l0 = l[j], l1 = l[ncq + j], t0 = tt[j], t1 = tt[ncq + j], vv = v[id];
                    const double dl = l1 - l0;
                    const double rd0 = cd[j] + t0 - vv, rd1 = cd[ncq + j] + t1 + vv;
                    rd[j] = rd0; rd[ncq + j] = rd1;
                    const double m0 = l0 * t0, m1 = l1 * t1;
                    musum += m0; musum += m1;
                    double q = dabs(m0); n3 = q > n3 ? q : n3; q = dabs(m1); n3 = q > n3 ? q : n3;
                    q = dabs(rd0); n2 = q > n2 ? q : n2; q = dabs(rd1); n2 = q > n2 ? q : n2;
                    const double ti0 = 1.0 / t0, ti1 = 1.0 / t1;
                    ti[j] = ti0; ti[ncq + j] = ti1;
                    const double Gs = ti0 * l0 + ti1 * l1;
                    const double gd = ti0 * ((m0 - tau) - l0 * rd0) - ti1 * ((m1 - tau) - l1 * rd1);
#pragma unroll
                    for (int i = 0; i < NV; i++) if (i == id) { g[i] += dl; dg[i] += Gs; gg[i] += gd; }
                }
                const TP gxy = FT(Y.gxy, k);
#pragma unroll 1
                for (int c = 0; c < K; c++)
                {
                    const int r = nbq + c;
                    const double gX = k >= 1 ? gxy[c] : 0.0, gY = k >= 1 ? gxy[K + c] : 0.0;
                    const double l0 = l[r], l1 = l[ncq + r], t0 = tt[r], t1 = tt[ncq + r];
                    const double dl = l1 - l0;
                    g[HXV] += gX * dl; g[HYV] += gY * dl;
                    const double vv = gX * v[HXV] + gY * v[HYV];
                    const double rd0 = cd[r] + t0 - vv, rd1 = cd[ncq + r] + t1 + vv;
                    rd[r] = rd0; rd[ncq + r] = rd1;
                    const double m0 = l0 * t0, m1 = l1 * t1;
                    musum += m0; musum += m1;
                    double q = dabs(m0); n3 = q > n3 ? q : n3; q = dabs(m1); n3 = q > n3 ? q : n3;
                    q = dabs(rd0); n2 = q > n2 ? q : n2; q = dabs(rd1); n2 = q > n2 ? q : n2;
                    const double ti0 = 1.0 / t0, ti1 = 1.0 / t1;
                    ti[r] = ti0; ti[ncq + r] = ti1;
                    const double Gs = ti0 * l0 + ti1 * l1;
                    const double gd = ti0 * ((m0 - tau) - l0 * rd0) - ti1 * ((m1 - tau) - l1 * rd1);
                    dg[HXV] += (gX * Gs) * gX; aYX += (gY * Gs) * gX; dg[HYV] += (gY * Gs) * gY;
                    gg[HXV] += gd * gX; gg[HYV] += gd * gY;
                }
                const TP BAt = FT(Y.BAtT, k);
                const TP vn = FT(Y.ux, k + 1);
                const TP cb = FT(Y.b, k);
                const SP rb = F(Y.rb, k);
                // no store between the loads: the residual goes to the record after both products
                double rbv[NX];
#pragma unroll
                for (int j = 0; j < NX; j++)
                {
                    double acc = cb[j] - vn[NU + j];
#pragma unroll
                    for (int i = 0; i < NV; i++) if (k > 0 || i < NU) acc += BAt[i + NV * j] * v[i];
                    rbv[j] = acc;
                    const double q = dabs(acc);
                    n1 = q > n1 ? q : n1;
                }
#pragma unroll
                for (int i = 0; i < NV; i++)
                {
                    double acc = 0.0;
#pragma unroll
                    for (int j = 0; j < NX; j++) acc += BAt[i + NV * j] * pk[j];
                    g[i] += acc;
                }
#pragma unroll
                for (int j = 0; j < NX; j++) rb[j] = rbv[j];
            }
            const TP rg = FT(Y.rg, k);
            if (K > 0 && k < N) Lk[HYV * NV + HXV] = T[HYV * (HYV + 1) / 2 + HXV] + aYX;
#pragma unroll
            for (int i = 0; i < NV; i++)
            {
                const double gi = var_active(k, i) ? g[i] : 0.0;
                rg[i] = gi;
                // finish the matrix to factorise: diagonal = template + Gamma terms, gradient row = rg + gamma terms
                Lk[i * NV + i] = T[i * (i + 1) / 2 + i] + dg[i];
                Lk[NV * NV + i] = gi + gg[i];
                const double q = dabs(gi);
                n0 = q > n0 ? q : n0;
            }
        }
        n4[0] = warp_max(n0); n4[1] = warp_max(n1); n4[2] = warp_max(n2); n4[3] = warp_max(n3);
        mu = warp_sum(musum) / nct;
        syncwarp();
    }

    // TMA streaming of the record heads for the chain sweeps: two buffers + the previous stage's
    MDEV double* rec_g(int k) const { return w + Y.rec_off + (long) k * Y.rec_size; }
#ifndef USVMPC_STREAM_TMA
#define USVMPC_STREAM_TMA 1
#endif
#if USVMPC_STREAM_TMA
    // TMA: one lane issues one bulk copy for the record head; completion is counted in bytes on the buffer's mbarrier
    MDEV void rec_init()
    {
        if (lane == 0) { mbar_init(bar + 0); mbar_init(bar + 1); mbar_init(bar + 2); fence_mbar_init(); }
        syncwarp();
    }
    MDEV void rec_fetch(const double* g, int b)
    {
        if (lane == 0) bulk_g2s(buf[b], g, HEAD * 8, bar + b);
    }
    MDEV void rec_wait(int b)
    {
        mbar_wait(bar + b, (phb >> b) & 1u);
        phb ^= 1u << b;
    }
    MDEV void rec_store(double* g, int b, int f0, int t0)
    {
        fence_proxy_async_smem();
        syncwarp();
        if (lane == 0) { bulk_s2g(g + f0, buf[b] + f0, (t0 - f0) * 8); bulk_commit(); }
    }
    MDEV void rec_reuse_guard() { if (lane == 0) bulk_wait_read1(); }
    MDEV void sweep_begin() { fence_proxy_async(); syncwarp(); }
    MDEV void sweep_end()
    {
        if (lane == 0) bulk_wait_all();
        syncwarp();
        fence_proxy_async();
    }
    MDEV void sweep_end_nostore() { syncwarp(); fence_proxy_async(); }
    // The sweeps that only READ the records (chainF, chainC) stream them with per-lane 16-byte cp.async instead: no
    // mbarrier round trip per stage (try_wait costs ~90 cycles even when the data is there) and no proxy fences, because
    // cp.async is a generic-proxy access like the passes' stores.  chainA, which also writes the records back, keeps TMA.
#ifndef USVMPC_READ_LDGSTS
#define USVMPC_READ_LDGSTS 1
#endif
#if USVMPC_READ_LDGSTS
    MDEV void rd_fetch(const double* src, int b)
    {
        double* dst = buf[b];
#pragma unroll 1
        for (int c = 2 * lane; c < HEAD; c += 64) cp_async16(dst + c, src + c);
        cp_async_commit();
    }
    MDEV void rd_wait(int b) { (void) b; cp_async_wait_all(); }
    MDEV void rd_begin() { syncwarp(); }
    MDEV void rd_end() { syncwarp(); }
#else
    MDEV void rd_fetch(const double* src, int b) { rec_fetch(src, b); }
    MDEV void rd_wait(int b) { rec_wait(b); }
    MDEV void rd_begin() { sweep_begin(); }
    MDEV void rd_end() { sweep_end_nostore(); }
#endif
#else
    // alternative streaming path: per-lane 16-byte cp.async (LDGSTS) loads and plain coalesced stores, all in the
    // generic proxy (no proxy fences, the L1 keeps what the passes read)
    MDEV void rec_init() {}
    MDEV void rec_fetch(const double* src, int b)
    {
        double* dst = buf[b];
#pragma unroll 1
        for (int c = 2 * lane; c < HEAD; c += 64) cp_async16(dst + c, src + c);
        cp_async_commit();
    }
    MDEV void rec_wait(int b) { (void) b; cp_async_wait_all(); syncwarp(); }
    MDEV void rec_store(double* dst, int b, int f0, int t0)
    {
        syncwarp();
        const double* src = buf[b];
#pragma unroll 1
        for (int c = f0 + 2 * lane; c < t0; c += 64) st2(dst + c, src + c);
    }
    MDEV void rec_reuse_guard() {}
    MDEV void sweep_begin() { syncwarp(); }
    MDEV void sweep_end() { syncwarp(); }
    MDEV void sweep_end_nostore() { syncwarp(); }
    MDEV void rd_fetch(const double* src, int b) { rec_fetch(src, b); }
    MDEV void rd_wait(int b) { rec_wait(b); }
    MDEV void rd_begin() { sweep_begin(); }
    MDEV void rd_end() { sweep_end_nostore(); }
#endif
    // stage 0 after x0 elimination: no x rows in [B';A'] (x_ocp_qp_red.c:268-454)
    MDEV void mask_stage0(double* R)
    {
#pragma unroll 1
        for (int e = lane; e < NV * NX; e += 32) if (e % NV >= NU) R[oBAt + e] = 0.0;
        syncwarp();
    }

    // chainA: Riccati factorisation, backward.  On entry the L field of every record holds H + Gamma terms with the
    // gradient row (passA); on exit the factor (NV+1) x NV, Pb and the backward vector (row NV).
    MDEV void chainA()
    {
        int ir = 0, in = 1, ip = 2;
        sweep_begin();
        double* gk = rec_g(N);  // record head of stage k in HBM
        rec_fetch(gk, ir);
#pragma unroll 1
        for (int k = N; k >= 0; k--, gk -= HEAD)
        {
            double *R = buf[ir], *Rp = buf[ip];
            PROF(1)
            rec_wait(ir);
            PROF(12)
            if (k > 0) { rec_reuse_guard(); rec_fetch(gk - HEAD, in); }
            if (k == 0) mask_stage0(R);
            PROF(13)
            double* Mx = R + oL;
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
                PROF(14)
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
                PROF(15)
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
                syncwarp();
                PROF(16)
            }
            // (NV+1) x NV Cholesky, lower; non-positive pivot => zero column
            // (dpotrf_l_mn pivot rule, BF/kernel/generic/kernel_dgemm_4x4_lib4.c:5701-5710).  Lane r <= NV takes row r
            // into registers; pivots and multipliers travel by shuffle.
            {
                const int r = lane;
                double Mr[NV], dinv = 0.0;
#pragma unroll
                for (int c = 0; c < NV; c++) Mr[c] = r <= NV ? Mx[r * NV + c] : 0.0;
#pragma unroll
                for (int j = 0; j < NV; j++)
                {
                    // the pivot and the unscaled column travel together; scaling the received entries locally gives the
                    // same bits as shuffling the scaled ones and takes the second shuffle off the dependent chain
                    const double piv = shfl(Mr[j], j);
                    double raw[NV];
#pragma unroll
                    for (int c = j + 1; c < NV; c++) raw[c] = shfl(Mr[j], c);
                    double sq = 0.0, inv = 0.0;
                    if (piv > 0.0) { inv = drsqrt(piv); sq = piv * inv; }
                    if (r == j) { Mr[j] = sq; if (j < NU) dinv = inv; } else Mr[j] *= inv;
#pragma unroll
                    for (int c = j + 1; c < NV; c++) Mr[c] -= Mr[j] * (raw[c] * inv);
                }
                if (lane <= NV)
                {
#pragma unroll
                    for (int c = 0; c < NV; c++) if (c <= lane) Mx[lane * NV + c] = Mr[c];
                    if (lane < NU) Mx[lane * NV + NV - 1] = dinv;  // 1/L[j][j] of the columns solved per stage (unused upper slot)
                    if (lane == NV)
                    {
#pragma unroll
                        for (int c = 0; c < NV; c++) R[obv + c] = Mr[c];  // backward vector of the forward substitution
                    }
                }
            }
            PROF(17)
            rec_store(gk, ir, oL, obv + svv);
            PROF(18)
            { const int t = ip; ip = ir; ir = in; in = t; }
        }
        sweep_end();
    }

    // chainF: forward substitution.  bv holds the backward vector (row NV of the factor after chainA, the eliminated
    // right-hand side after chainC); dux receives the primal step.
    MDEV void chainF()
    {
        int ir = 0, in = 1;
        double xc[NX], xme = 0.0;
#pragma unroll
        for (int i = 0; i < NX; i++) xc[i] = 0.0;
        rd_begin();
        const double* gk = rec_g(0);
        rd_fetch(gk, ir);
#pragma unroll 1
        for (int k = 0; k <= N; k++, gk += HEAD)
        {
            double* R = buf[ir];
            rd_wait(ir);
            syncwarp();  // every lane is done with the other buffer before it becomes a copy destination again
            if (k < N) rd_fetch(gk + HEAD, in);
            if (k == 0) mask_stage0(R);
            const double* L = R + oL;
            double zu[NU];
#pragma unroll
            for (int i = NU - 1; i >= 0; i--)  // dtrsv_ltn on the columns solved at this stage, every lane redundantly
            {
                double ax = 0.0, au = -R[obv + i];
#pragma unroll
                for (int m = NU; m < NV; m++) ax -= L[m * NV + i] * xc[m - NU];
#pragma unroll
                for (int m = i + 1; m < NU; m++) au -= L[m * NV + i] * zu[m];
                zu[i] = (au + ax) * L[i * NV + NV - 1];
                if (k == N) zu[i] = 0.0;
            }
            const TP g = FT(Y.dux, k);
            if (lane < NX) g[NU + lane] = xme;
            if (lane == 0)
            {
#pragma unroll
                for (int i = 0; i < NU; i++) g[i] = zu[i];
            }
            if (k < N)
            {
                double x1 = 0.0;
                if (lane < NX)
                {
                    double ax = R[orb + lane], au = 0.0;
#pragma unroll
                    for (int i = NU; i < NV; i++) ax += R[oBAt + i + NV * lane] * xc[i - NU];
#pragma unroll
                    for (int i = 0; i < NU; i++) au += R[oBAt + i + NV * lane] * zu[i];
                    x1 = ax + au;
                }
                xme = x1;
#pragma unroll
                for (int m = 0; m < NX; m++) xc[m] = shfl(x1, m);
            }
            { const int t = ir; ir = in; in = t; }
        }
        rd_end();
    }

    // chainC: backward substitution of OCP_QP_SOLVE_KKT_STEP.  bv holds rhs_g + constraint terms (passC) on entry, the
    // eliminated vector on exit.
    MDEV void chainC()
    {
        int ir = 0, in = 1;
        double pn[NX];
#pragma unroll
        for (int i = 0; i < NX; i++) pn[i] = 0.0;
        rd_begin();
        double* gk = rec_g(N);
        rd_fetch(gk, ir);
#pragma unroll 1
        for (int k = N; k >= 0; k--, gk -= HEAD)
        {
            double* R = buf[ir];
            rd_wait(ir);
            syncwarp();  // every lane is done with the other buffer before it becomes a copy destination again
            if (k > 0) rd_fetch(gk - HEAD, in);
            if (k == 0) mask_stage0(R);
            const int i = lane;
            double zi = 0.0;
            if (i < NV && var_active(k, i))
            {
                zi = R[obv + i];
                if (k < N)
                {
                    double acc = 0.0;
#pragma unroll
                    for (int j = 0; j < NX; j++) acc += R[oBAt + i + NV * j] * (pn[j] + R[oPb + j]);
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
            if (i < NV) gk[obv + i] = zi;
#pragma unroll
            for (int j = 0; j < NX; j++) pn[j] = shfl(zi, NU + j);
            { const int t = ir; ir = in; in = t; }
        }
        solve_calls++;
        rd_end();
    }

    // passC: right-hand side of the corrector / centering solve res_m = lam*t [+ dt_aff*dlam_aff] - sigma_mu
    // (x_ocp_qp_ipm.c:2138-2160, 2175-2200) -> rmc; COMPUTE_GAMMA_QP (x_core_qp_ipm_aux.c:89-113); bv = rhs_g + J'(gamma_l - gamma_u)
    MDEV void passC(bool with_aff, double sigma_mu)
    {
#pragma unroll 1
        for (int k = lane; k <= N; k += 32)
        {
            const TP rg = FT(Y.rg, k);
            const SP bv = F(Y.bv, k);
            double z[NV];
#pragma unroll
            for (int i = 0; i < NV; i++) z[i] = rg[i];
            if (k < N)
            {
                const TP lam = FT(Y.lam, k), t = FT(Y.t, k), ti = FT(Y.ti, k), rd = FT(Y.rd, k);
                const TP dl = FT(Y.dlam, k), dtt = FT(Y.dt, k), gxy = FT(Y.gxy, k);
                const TP rm = FT(Y.rmc, k);
#pragma unroll 1
                for (int j = 0; j < ncq; j++)
                {
                    if (!row_active(k, j)) { rm[j] = 0.0; rm[ncq + j] = 0.0; continue; }
                    // every load of the row before its first store (a load cannot be hoisted above a store that may alias)
                    const double l0 = lam[j], l1 = lam[ncq + j], i0 = ti[j], i1 = ti[ncq + j], e0 = rd[j], e1 = rd[ncq + j];
                    double gX = 0.0, gY = 0.0;
                    if (j >= nbq && k >= 1) { gX = gxy[j - nbq]; gY = gxy[K + j - nbq]; }
                    double m0 = l0 * t[j], m1 = l1 * t[ncq + j];
                    if (with_aff) { m0 += dtt[j] * dl[j]; m1 += dtt[ncq + j] * dl[ncq + j]; }
                    m0 -= sigma_mu; m1 -= sigma_mu;
                    rm[j] = m0; rm[ncq + j] = m1;
                    const double gd = i0 * (m0 - l0 * e0) - i1 * (m1 - l1 * e1);
                    if (j < nbq)
                    {
                        const int id = srvar[j];
#pragma unroll
                        for (int i = 0; i < NV; i++) if (i == id) z[i] += gd;
                    }
                    else if (k >= 1) { z[HXV] += gX * gd; z[HYV] += gY * gd; }
                }
            }
#pragma unroll
            for (int i = 0; i < NV; i++) bv[i] = var_active(k, i) ? z[i] : 0.0;
        }
        syncwarp();
    }

    // passF: dt, dlam from dux (x_ocp_qp_kkt.c:748-764 + COMPUTE_LAM_T_QP), COMPUTE_ALPHA_QP, the sums COMPUTE_MU_AFF_QP
    // needs, and dpi_k = Lxx (Lxx' dx_{k+1} + l_x) | p_{k+1} + Lxx (Lxx' dx_{k+1}) (x_ocp_qp_kkt.c:560-575 | 1270-1290).
    // corr = false: affine step (res_m = lam*t - tau, bv = row NV of the factor); true: res_m = rmc, bv = eliminated rhs.
    MDEV void passF(bool corr, double tau)
    {
        double bdn = 1.0, bdd = -1.0, bpn = 1.0, bpd = -1.0, s1 = 0.0, s2 = 0.0;  // best dual / primal ratio = -1
        double nd = 0.0, nm = 0.0;  // OCP_QP_RES_COMPUTE_LIN rows res_d, res_m of this very step: their inputs are in registers here
#pragma unroll 1
        for (int k = lane; k < N; k += 32)
        {
            const TP v = FT(Y.dux, k), lam = FT(Y.lam, k), t = FT(Y.t, k), ti = FT(Y.ti, k), rd = FT(Y.rd, k);
            const TP rm = FT(Y.rmc, k), gxy = FT(Y.gxy, k);
            const TP dl = FT(Y.dlam, k), dtt = FT(Y.dt, k);
            const double vx = v[HXV], vy = v[HYV];
#pragma unroll 1
            for (int j = 0; j < ncq; j++)
            {
                if (!row_active(k, j)) continue;
                // both sides' loads before the first store of the row (a load cannot be hoisted above a store that may alias)
                double bl[2], bt[2], bi[2], be[2], bm[2];
#pragma unroll
                for (int side = 0; side < 2; side++)
                {
                    const int r = j + side * ncq;
                    bl[side] = lam[r]; bt[side] = t[r]; bi[side] = ti[r]; be[side] = rd[r];
                    bm[side] = corr ? rm[r] : 0.0;
                }
                double dv;
                if (j < nbq) dv = v[srvar[j]];
                else dv = k >= 1 ? gxy[j - nbq] * vx + gxy[K + j - nbq] * vy : 0.0;
#pragma unroll
                for (int side = 0; side < 2; side++)
                {
                    const int r = j + side * ncq;
                    double dtr = side ? -dv : dv;
                    const double lam0 = bl[side], t0 = bt[side];
                    const double m = corr ? bm[side] : lam0 * t0 - tau;
                    const double dlr = -bi[side] * (m + (lam0 * dtr) - (lam0 * be[side]));
                    dtr -= be[side];
                    dl[r] = dlr; dtt[r] = dtr;
                    // COMPUTE_ALPHA_QP keeps the ratio closest to zero among rows with a negative step; the running best
                    // is kept as (numerator, denominator) and compared by cross-multiplication: one division per sweep
                    if (dlr < 0.0 && bdn * dlr < lam0 * bdd) { bdn = lam0; bdd = dlr; }
                    if (dtr < 0.0 && bpn * dtr < t0 * bpd) { bpn = t0; bpd = dtr; }
                    s1 += lam0 * dtr + t0 * dlr;
                    s2 += dlr * dtr;
                    {
                        // rd + dt -/+ J dux  and  rhs_m + lam*dt + dlam*t  (HP/ocp_qp/x_ocp_qp_res.c:560-633), as res_pass has them
                        const double e = side ? be[side] + dtr + dv : be[side] + dtr - dv;
                        const double mm = m + lam0 * dtr + dlr * t0;
                        double q = dabs(e); nd = q > nd ? q : nd;
                        q = dabs(mm); nm = q > nm ? q : nm;
                    }
                }
            }
            // dpi_k from the factor of stage k+1
            const SP Ln = F(Y.L, k + 1);
            const SP bn = F(Y.bv, k + 1);
            const TP xn = FT(Y.dux, k + 1);
            const TP dpi = FT(Y.dpi, k);
            double tmp[NX];
#pragma unroll
            for (int j = 0; j < NX; j++)
            {
                double acc = 0.0;
#pragma unroll
                for (int m = j; m < NX; m++) acc += Ln[(NU + m) * NV + NU + j] * xn[NU + m];
                tmp[j] = corr ? acc : acc + bn[NU + j];
            }
            double dpv[NX];
#pragma unroll
            for (int i = 0; i < NX; i++)
            {
                double acc = 0.0;
#pragma unroll
                for (int j = 0; j <= i; j++) acc += Ln[(NU + i) * NV + NU + j] * tmp[j];
                dpv[i] = corr ? bn[NU + i] + acc : acc;
            }
#pragma unroll
            for (int i = 0; i < NX; i++) dpi[i] = dpv[i];  // stores after the last load
        }
        const double a_prim = warp_max(bpn / bpd), a_dual = warp_max(bdn / bdd);
        alpha = -(a_prim > a_dual ? a_prim : a_dual);
        S1 = warp_sum(s1); S2 = warp_sum(s2);
        lin_d = warp_max(nd); lin_m = warp_max(nm);
        syncwarp();
    }

    MDEV bool itref_ok(const double* n) const
    {
        return (n[0] < tol_stat || n[0] < 1e-3 * res_max[0]) && (n[1] < tol_eq || n[1] < 1e-3 * res_max[1]) &&
               (n[2] < tol_ineq || n[2] < 1e-3 * res_max[2]) && (n[3] < tol_comp || n[3] < 1e-3 * res_max[3]);
    }

    // OCP_QP_IPM_SOLVE + OCP_QP_IPM_DELTA_STEP: HP/ocp_qp/x_ocp_qp_ipm.c:2354-2683, 1888-2350 (pred_corr,
    // cond_pred_corr, itref_corr_max = 2); returns HPIPM status 0 ok / 1 max iter / 2 min step / 3 NaN.
    MDEV int ipm_solve(int* iters)
    {
        const double tau_min = 1e-16, alpha_min = 1e-8;
        ipm_init();
        PROF(10)
        alpha = 1.0;
        double a = 0.0;  // the first passA applies no step
        int kk = 0;
        for (;;)
        {
            passA(a, tau_min, res_max);
            PROF(0)
            if (!(kk < iter_max && alpha > alpha_min &&
                  (res_max[0] > tol_stat || res_max[1] > tol_eq || res_max[2] > tol_ineq ||
                   dabs(res_max[3] - tau_min) > tol_comp)))
                break;
            chainA();
            PROF(1)
            double sigma_mu = 0.0, mu_aff0 = 0.0;
            for (int pass = 0; pass < 3; pass++)
            {
                // pass 0: affine step; pass 1: corrector; pass 2: centering only (conditional)
                if (pass > 0)
                {
                    passC(pass == 1, sigma_mu);
                    PROF(4)
                    chainC();
                    PROF(5)
                }
                chainF();
                PROF(2)
                passF(pass > 0, tau_min);
                PROF(3)
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
            // residual of the linear system at the step (OCP_QP_RES_COMPUTE_LIN); iterative refinement is rare
            double nlin[4];
            res_pass<false>(nlin);
            PROF(6)
            bool refined = false;
            for (int it = 0; it < 2; it++)
            {
                if (itref_ok(nlin)) break;
                res_pass<true>(nlin);
                solve_sweep(true);
                forward_sweep(Y.rb2, Y.dux2, Y.dpi2, false);
                expand_pass(2, 0.0);
                add_refinement();
                refined = true;
                res_pass<true>(nlin);  // the step is no longer passF's: all four norms from scratch (the stores are harmless)
            }
            if (refined) alpha_pass();
            PROF(7)
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
    // OCP_QP_RES_COMPUTE_LIN (HP/ocp_qp/x_ocp_qp_res.c:468-633): residual of the Newton system with right-hand side
    // (rg, rb, rd, rmc) at the step (dux, dpi, dlam, dt); norms into out4.  WRITE: also store it (rg2, rb2, rd2, rm2) as
    // the right-hand side of an iterative-refinement solve.
    template <bool WRITE>
    MDEV void res_pass(double* out4)
    {
        double n0 = 0, n1 = 0, n2 = 0, n3 = 0;
        for (int k = lane; k <= N; k += 32)
        {
            const TP v = FT(Y.dux, k), pk = FT(Y.dpi, k), l = FT(Y.dlam, k), tt = FT(Y.dt, k);
            const double* H = Hk(k);
            const TP cg = FT(Y.rg, k), cd = FT(Y.rd, k);
            const SP rg = F(Y.rg2, k), rd = F(Y.rd2, k);
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
                const TP pm = FT(Y.dpi, k - 1);
#pragma unroll
                for (int i = 0; i < NX; i++) g[NU + i] -= pm[i];
            }
            if (k < N)
            {
                for (int j = 0; j < nbq; j++)
                {
                    if (!row_active(k, j)) continue;
                    const int id = j < nbu ? j : NU + P.idxbx[j - nbu];
                    const double dl = l[ncq + j] - l[j];
#pragma unroll
                    for (int i = 0; i < NV; i++) if (i == id) g[i] += dl;
                    if (WRITE)  // without WRITE the step is passF's and so are these rows' norms (lin_d)
                    {
                        const double vv = v[id];
                        const double e0 = cd[j] + tt[j] - vv, e1 = cd[ncq + j] + tt[ncq + j] + vv;
                        rd[j] = e0; rd[ncq + j] = e1;
                        const double q0 = dabs(e0), q1 = dabs(e1); n2 = q0 > n2 ? q0 : n2; n2 = q1 > n2 ? q1 : n2;
                    }
                }
                const TP gxy = FT(Y.gxy, k);
                for (int c = 0; c < K; c++)
                {
                    const int r = nbq + c;
                    const double gX = k >= 1 ? gxy[c] : 0.0, gY = k >= 1 ? gxy[K + c] : 0.0;
                    const double dl = l[ncq + r] - l[r];
                    g[HXV] += gX * dl; g[HYV] += gY * dl;
                    if (WRITE)
                    {
                        const double vv = gX * v[HXV] + gY * v[HYV];
                        const double e0 = cd[r] + tt[r] - vv, e1 = cd[ncq + r] + tt[ncq + r] + vv;
                        rd[r] = e0; rd[ncq + r] = e1;
                        const double q0 = dabs(e0), q1 = dabs(e1); n2 = q0 > n2 ? q0 : n2; n2 = q1 > n2 ? q1 : n2;
                    }
                }
                const TP BAt = FT(Y.BAtT, k);
                const TP vn = FT(Y.dux, k + 1);
                const SP cb = F(Y.rb, k);
                const SP rb = F(Y.rb2, k);
#pragma unroll
                for (int j = 0; j < NX; j++)
                {
                    double acc = cb[j] - vn[NU + j];
#pragma unroll
                    for (int i = 0; i < NV; i++) if (var_active(k, i)) acc += BAt[i + NV * j] * v[i];
                    if (WRITE) rb[j] = acc;
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
                if (WRITE)
                {
                    const TP lam = FT(Y.lam, k), t = FT(Y.t, k), rm = FT(Y.rmc, k);
                    const SP rm2 = F(Y.rm2, k);
                    for (int j = 0; j < 2 * ncq; j++)
                    {
                        if (!row_active(k, j % ncq)) { if (WRITE) rm2[j] = 0.0; continue; }
                        const double m = rm[j] + lam[j] * tt[j] + l[j] * t[j];
                        if (WRITE) rm2[j] = m;
                        const double a = dabs(m);
                        n3 = a > n3 ? a : n3;
                    }
                }
            }
#pragma unroll
            for (int i = 0; i < NV; i++)
            {
                const double gi = var_active(k, i) ? g[i] : 0.0;
                if (WRITE) rg[i] = gi;
                const double a = dabs(gi);
                n0 = a > n0 ? a : n0;
            }
        }
        out4[0] = warp_max(n0); out4[1] = warp_max(n1); out4[2] = warp_max(n2); out4[3] = warp_max(n3);
        if (!WRITE) { out4[2] = lin_d; out4[3] = lin_m; }
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
            const SP v = F(fv, k), lam = F(Y.lam, k), t = F(Y.t, k), rd = F(frd, k), rm = F(frm, k);
            const SP gxy = F(Y.gxy, k);
            const SP dl = F(fl, k), dtt = F(ft, k);
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
            const SP lam = F(Y.lam, k), t = F(Y.t, k), dl = F(Y.dlam, k), dtt = F(Y.dt, k);
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
            const SP a = F(Y.dux, k); const SP b = F(Y.dux2, k);
            for (int i = 0; i < NV; i++) a[i] += b[i];
            if (k < N)
            {
                const SP c = F(Y.dpi, k); const SP e = F(Y.dpi2, k);
                for (int i = 0; i < NX; i++) c[i] += e[i];
                const SP l = F(Y.dlam, k), t = F(Y.dt, k);
                const SP l2 = F(Y.dlam2, k), t2 = F(Y.dt2, k);
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
            const SP g = F(Y.BAt, k);
            for (int e = lane; e < NV * NX; e += 32) sBA[e] = var_active(k, e % NV) ? g[e] : 0.0;
        }
        else
            for (int e = lane; e < NV * NX; e += 32) sBA[e] = 0.0;
    }

    MDEV void load_gxy(int k)
    {
        if (k < N)
        {
            const SP g = F(Y.gxy, k);
            for (int e = lane; e < 2 * K; e += 32) sgxy[e] = k >= 1 ? g[e] : 0.0;
        }
    }

    MDEV void load_Lnext(int k1)  // xx block and nothing else of the factor of stage k1
    {
        const SP L = F(Y.L, k1);
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
                const SP L = F(Y.L, k);
                for (int e = lane; e < NR * NV; e += 32) sL[e] = L[e];
                const SP q = F(fdux, k);
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
                const SP o = F(fdux, k);
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
                const SP lam = F(Y.lam, k), t = F(Y.t, k), rd = F(frd, k), rm = F(frm, k);
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
            const SP L = F(Y.L, k);
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
            const SP z = F(Y.zux, k), zl = F(Y.zlam, k), zt = F(Y.zt, k);
            const SP ux = F(Y.ux, k), lam = F(Y.lam, k), t = F(Y.t, k);
            if (k < N)
            {
                const SP zp = F(Y.zpi, k); const SP pi = F(Y.pi, k);
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
                const SP zf = F(Y.zfun, 0);
                double s[NV], tmp[NV];
                for (int i = 0; i < NU; i++) s[i] = ux[i];
                for (int i = 0; i < NX; i++) s[NU + i] = zf[nbu + i];  // dx0 = x0 - x_0
                const SP rq = F(Y.rq, 0);
                const SP BAt = F(Y.BAt, 0);
                const SP pi = F(Y.pi, 0);
                const SP gxy = F(Y.gxy, 0);
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
            const SP zl = F(Y.zlam, k), zt = F(Y.zt, k), zf = F(Y.zfun, k);
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
#ifdef USVMPC_PROFILE
        for (int i = 0; i < 20; i++) prof[i] = 0;
        tprev = clock64();
#endif
        int status = 2, sqp_iter = 0, qp_total = 0, qp_status = 0, qp_iter = 0;
        double res[4] = {0, 0, 0, 0};
        const int max_iter = P.nlp_type == 0 ? P.max_iter : 1;
        for (sqp_iter = 0; sqp_iter < max_iter; sqp_iter++)
        {
            linearize(x0, pg, lhg, yrg, yre, res);
            PROF(8)
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
            PROF(9)
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
#ifdef USVMPC_PROFILE
            if (qp_total >= USVMPC_PROFILE)  // the define is the threshold: only long-running instances report
                printf("PROF inst %d sqp %d qp %d | passA %lld chainA %lld chainF %lld passF %lld passC %lld chainC %lld "
                       "res %lld refine %lld lin %lld upd %lld init %lld\n", inst, sqp_iter, qp_total, prof[0], prof[1],
                       prof[2], prof[3], prof[4], prof[5], prof[6], prof[7], prof[8], prof[9], prof[10]);
            if (qp_total >= USVMPC_PROFILE)
                printf("PROFA wait %lld fetch %lld trmm %lld pb %lld syrk %lld chol %lld store %lld\n", prof[12], prof[13],
                       prof[14], prof[15], prof[16], prof[17], prof[18]);
#endif
        }
    }
};

}  // namespace usvmpc
