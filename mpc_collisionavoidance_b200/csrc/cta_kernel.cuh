// cta_kernel.cuh -- the batched NMPC solve: ONE THREAD BLOCK PER INSTANCE, working set resident in shared memory.
//
// A persistent block pulls instances from a device-side queue and runs the whole SQP / SQP_RTI solve of one
// instance at a time: linearise -> x0 elimination -> Mehrotra IPM on a Riccati recursion -> variable update.
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
//   BLASFEO potrf/syrk/trmm/trsv/gemv             BF/blasfeo_hp_pm/d_lapack_lib4.c:1149,1503 etc. -> block/warp-level code below
//
// Work decomposition inside the block (T threads, W = T/32 warps):
//   * everything of an IPM iteration that is independent per stage -- variable update, residuals, Gamma/gamma, assembly
//     of the matrices to factorise, gains, dt/dlam, step lengths, right-hand sides -- is a PASS: the (stage, row) or
//     (stage, variable) items are spread over all T threads, block-wide reductions give norms / step lengths;
//   * only the Riccati recursions are serial in the stage index.  They run as CHAINS on warp 0, straight out of shared
//     memory, on the fp64 tensor cores (mma.sync.m8n8k4.f64), and are cut down to what really is serial:
//       chainA  backward: P_k, p_k from P_{k+1}, p_{k+1}: M = (H + Gamma terms) + G' P_{k+1} G, eliminate the NU input
//               columns (Cholesky with HPIPM's pivot rule); the Schur complement IS P_k, so the x block is never
//               factorised (the reference factorises it only to form G' P G as a product of triangular factors);
//       chainC  backward: p_k = Acl_k' p_{k+1} + e_k          (one NX x NX mat-vec per stage)
//       chainF  forward : dx_{k+1} = Acl_k dx_k + c_k          (one NX x NX mat-vec per stage)
//     with the closed-loop matrices Acl_k = A_k + B_k K_k, the gains K_k and the affine terms c_k, e_k computed by
//     passes between the chains.  Mathematically this is the reference's recursion (x_ocp_qp_kkt.c:455-575,
//     1096-1290); iteration counts, status and converged trajectories are the reference's (tests/).
//     A lone warp issues about one instruction per four cycles whatever the instruction, so the chains are written for
//     instruction count: per-lane shared addresses held in registers, no predicated accesses, DMMA for the products.
// Inactive variables (x at stage 0 after x0 elimination, u at stage N) and inactive inequality rows are kept in
// the uniform per-stage layout and masked (identity rows in the matrix, zero rows in [B';A']), so every stage runs
// the same code.
#pragma once
#include "cta_compat.h"
#include "cta_layout.h"
#include "models.cuh"

namespace usvmpc {

// exact x / d for 0 <= x, x * (d - 1) < 2^32
struct FastDiv {
    int d;
    unsigned magic;
    MDEV void set(int dd) { d = dd > 0 ? dd : 1; magic = (unsigned) ((0x100000000ull + (unsigned) d - 1) / (unsigned) d); }
    MDEV int div(int x) const { return d == 1 ? x : (int) (((unsigned long long) (unsigned) x * magic) >> 32); }
};

// ---------------------------------------------------------------- the serial recursions
#ifndef USVMPC_CHAIN_WARPS
#define USVMPC_CHAIN_WARPS 2
#endif
#ifndef USVMPC_CHAIN_MMA
#define USVMPC_CHAIN_MMA 1   // fp64 factorisation on the tensor cores (chainA_mma); 0: chainA_impl<double> on CHAIN_WARPS warps
#endif
constexpr int CHAIN_WARPS = USVMPC_CHAIN_WARPS;  // warps that share one stage step of the factorisation
template <int CT>
DEV void chain_sync()
{
    if (CT == 32) syncwarp(); else named_barrier_sync(1, CT);
}

// chainA: Riccati factorisation, backward (x_ocp_qp_kkt.c:455-535).  On entry the M field of every stage holds
// H + Gamma terms with the gradient row (passA); on exit, per stage: columns < NU = the Cholesky columns of the input
// block (rows NU.. = Lxu, row NV = l_u), the rest = P_k (lower, packed) and p_k (row NV); dinv = 1 / diag(Luu);
// Pb = P_{k+1} res_b.  Per stage three steps, each spread over the 32 lanes, with one __syncwarp between them:
//   1. sW = P_{k+1} [G' | res_b]  (NX x (NV+1); the last column + p_{k+1} is what the gradient row needs)
//   2. sM = M + G sW              (lower trapezoid, NE entries)
//   3. eliminate the NU input columns: every lane redoes the NU x NU Cholesky of the input block (2 dependent
//      reciprocal square roots for NU = 2) and finishes its own entries; the Schur complement goes to M and, as a full
//      symmetric matrix, to sP for the next stage.  Non-positive pivot => that column becomes zero (dpotrf_l_mn,
//      BF/kernel/generic/kernel_dgemm_4x4_lib4.c:5701-5710).
template <class M, class RT>
MDEVNI void chainA_impl(const double* G, double* Mx, const double* rb, double* Pb, double* dinv, double* sWd, double* sPd,
                        int N)
{
    constexpr int NX = M::NX, NU = M::NU, NV = NX + NU, NR = NV + 1, NE = NV * (NV + 1) / 2 + NV;
    constexpr int CT = 32 * CHAIN_WARPS;  // threads that share the items of a step (warps 0 .. CHAIN_WARPS-1)
    constexpr int NQ = (NE + CT - 1) / CT, NWQ = (NX * NR + CT - 1) / CT;
    ASSUME_SHARED(G); ASSUME_SHARED(Mx); ASSUME_SHARED(rb); ASSUME_SHARED(Pb); ASSUME_SHARED(dinv);
    ASSUME_SHARED(sWd); ASSUME_SHARED(sPd);
    // RT = double: the reference's arithmetic.  RT = float: the factorisation (these three steps) runs in fp32 -- the
    // `s_` twin of the reference (x_ocp_qp_kkt.c:457-461 is its only precision-dependent branch) -- while residuals,
    // the solves with the factor and iterative refinement stay fp64 (BASELINE.json config 4).
    RT* sW = (RT*) sWd; RT* sP = (RT*) sPd;
    RT* sM = sW + NX * NR;  // scratch copy of the matrix between steps 2 and 3
    const int lane = thread_id();  // index among the CT chain threads
    // Uniform control flow: every lane runs every step on clamped item indices (a surplus lane recomputes the last item
    // and stores the same value to the same address), so there is no divergent branch between the __syncwarps.
    // step 1 items: e -> output (m, i) of sW = P [G' | rb]; source vector = row i of G (stride NV) or rb (stride 1)
    int w_e[NWQ], w_m[NWQ], w_src[NWQ], w_str[NWQ];
    bool w_last[NWQ];
#pragma unroll
    for (int q = 0; q < NWQ; q++)
    {
        int e = lane + CT * q;
        e = e < NX * NR ? e : NX * NR - 1;
        const int i = e / NX;
        w_e[q] = e; w_m[q] = e - i * NX; w_last[q] = i == NV;
        w_src[q] = i < NV ? i : 0; w_str[q] = i < NV ? NV : 1;
    }
    // step 2 / 3 items: e -> entry (r, c) of the lower trapezoid
    int m_e[NQ], m_r[NQ], m_c[NQ];
#pragma unroll
    for (int q = 0; q < NQ; q++)
    {
        int e = lane + CT * q;
        e = e < NE ? e : NE - 1;
        int r = 0;
        while ((r + 1) * (r + 2) / 2 <= e && r < NV) r++;
        m_e[q] = e; m_r[q] = r; m_c[q] = e - r * (r + 1) / 2;
    }
#pragma unroll 1
    for (int k = N; k >= 0; k--)
    {
        double* Mk = Mx + k * NE;
        if (k < N)
        {
            const double* Gk = G + k * (NV * NX);
            const double* rbk = rb + k * NX;
            const double* pn = Mk + NE + (NV * (NV + 1)) / 2 + NU;  // p_{k+1}: x part of row NV of stage k+1
#pragma unroll
            for (int q = 0; q < NWQ; q++)
            {
                const double* ga = w_last[q] ? rbk : Gk + w_src[q];
                const RT* pr = sP + w_m[q] * NX;
                RT a0 = 0, a1 = 0;
#pragma unroll
                for (int n = 0; n < NX; n += 2)
                {
                    a0 += pr[n] * (RT) ga[n * w_str[q]];
                    if (n + 1 < NX) a1 += pr[n + 1] * (RT) ga[(n + 1) * w_str[q]];
                }
                const RT acc = a0 + a1;
                if (w_last[q]) Pb[k * NX + w_m[q]] = acc;
                sW[w_e[q]] = w_last[q] ? acc + (RT) pn[w_m[q]] : acc;
            }
            chain_sync<CT>();
#pragma unroll
            for (int q = 0; q < NQ; q++)
            {
                const int r = m_r[q], c = m_c[q];
                const double* gr = Gk + (r < NV ? r : c);
                const RT* wc = sW + (r < NV ? c : NV) * NX;
                RT a0 = (RT) Mk[m_e[q]], a1 = 0;
#pragma unroll
                for (int m = 0; m < NX; m += 2)
                {
                    a0 += (RT) gr[NV * m] * wc[m];
                    if (m + 1 < NX) a1 += (RT) gr[NV * (m + 1)] * wc[m + 1];
                }
                sM[m_e[q]] = a0 + a1;
            }
        }
        else
        {
#pragma unroll
            for (int q = 0; q < NQ; q++) sM[m_e[q]] = (RT) Mk[m_e[q]];
        }
        chain_sync<CT>();
        // Cholesky of the NU x NU input block, redundantly on every lane
        RT Lt[NU][NU], inv[NU];
        if (NU == 2)
        {
            // both reciprocal square roots of the 2 x 2 block at once: piv_1 = M11 - M10^2 / M00 = det / M00, so
            // 1 / sqrt(piv_1) = sqrt(M00) / sqrt(det) and rsqrt(M00), rsqrt(det) do not depend on each other (half the
            // dependent chain of the two-column Cholesky).  Pivot rule unchanged: piv_1 > 0 <=> det > 0 when M00 > 0.
            const RT m00 = sM[0], m10 = sM[1], m11 = sM[2];
            const RT det = m00 * m11 - m10 * m10;
            const RT i0 = m00 > (RT) 0 ? drsqrt(m00) : (RT) 0;
            RT i1;
            if (m00 > (RT) 0) i1 = det > (RT) 0 ? drsqrt(det) * (m00 * i0) : (RT) 0;
            else i1 = m11 > (RT) 0 ? drsqrt(m11) : (RT) 0;  // zero first column: the second pivot is M11 itself
            const RT l10 = m10 * i0;
            inv[0] = i0; inv[NU - 1] = i1;
            Lt[0][0] = m00 * i0; Lt[NU - 1][0] = l10;
            Lt[NU - 1][NU - 1] = (m11 - l10 * l10) * i1;
        }
        else
#pragma unroll
        for (int j = 0; j < NU; j++)
        {
            RT piv = sM[j * (j + 1) / 2 + j];
#pragma unroll
            for (int i = 0; i < j; i++) piv -= Lt[j][i] * Lt[j][i];
            inv[j] = piv > (RT) 0 ? drsqrt(piv) : (RT) 0;
            Lt[j][j] = piv * inv[j];
#pragma unroll
            for (int jj = j + 1; jj < NU; jj++)
            {
                RT v = sM[jj * (jj + 1) / 2 + j];
#pragma unroll
                for (int i = 0; i < j; i++) v -= Lt[jj][i] * Lt[j][i];
                Lt[jj][j] = v * inv[j];
            }
        }
#pragma unroll
        for (int j = 0; j < NU; j++) if (lane == j) dinv[k * NU + j] = inv[j];
#pragma unroll
        for (int q = 0; q < NQ; q++)
        {
            const int r = m_r[q], c = m_c[q];
            // first NU entries of rows r and c of the factor by substitution.  For a row of the input block itself the
            // same formula gives its entries up to the diagonal (what lies beyond is not used); row c is only used for
            // c >= NU.
            const RT* rr = sM + r * (r + 1) / 2;
            const RT* rc = sM + c * (c + 1) / 2;
            RT v[NU], u[NU];
#pragma unroll
            for (int j = 0; j < NU; j++)
            {
                RT x = rr[j], y = rc[j];
#pragma unroll
                for (int i = 0; i < j; i++) { x -= v[i] * Lt[j][i]; y -= u[i] * Lt[j][i]; }
                v[j] = x * inv[j]; u[j] = y * inv[j];
            }
            RT out = sM[m_e[q]];
#pragma unroll
            for (int j = 0; j < NU; j++) out -= v[j] * u[j];
#pragma unroll
            for (int j = 0; j < NU; j++) out = c == j ? v[j] : out;
            Mk[m_e[q]] = out;
            if (c >= NU && r < NV)
            {
                sP[(r - NU) * NX + (c - NU)] = out;
                sP[(c - NU) * NX + (r - NU)] = out;
            }
        }
        chain_sync<CT>();
    }
}

// chainA in fp64 on the tensor cores (the product path; chainA_impl above stays for the fp32 variant and as the
// readable statement of the step).  ONE warp; per stage two small products as DMMA m8n8k4 tiles and the elimination
// of the input columns on the accumulator fragments:
//   1. W (NX x NR) = P_{k+1} [G' | res_b]          A = P from sPp (row stride WS), B = G / res_b straight from the fields
//   2. S (NV x NR) = M + G W                        A = the G fragments of step 1, B = W through sWt ([column][row], stride WS)
//      column NV of S is the gradient row of the packed matrix (S[i][NV] = entry (NV, i))
//   3. Cholesky of the NU x NU input block (redone by every lane) and the Schur complement on the lane's own entries;
//      columns 0 .. NU-1 and NV of S travel through a small exchange array X[row] = {S[row][0], S[row][1], S[row][NV]}.
// With g = lane / 4, t = lane % 4 a lane holds S[8 mt + g][8 nt + 2 t + e], e = 0, 1.
// A lone warp runs at ~4 cycles per instruction whatever the instruction is, so the loop is written for instruction
// count: every shared-memory address is a per-lane 32-bit register that is only decremented by the stage stride
// (lds64 / sts64), lanes whose index falls outside a matrix read a clamped in-bounds entry (finite, multiplied by an
// exact zero or feeding an unused entry) and write to a dump slot, so the loop has no predicated memory access.
// Padding that must be exact (k-padding of P, rows >= NX of W) reads a zero slot.
// Same pivot rule and the same outputs as chainA_impl: Mx (factor columns, P_k, p_k), Pb, dinv.
template <class M>
MDEVNI void chainA_mma(const double* G, double* Mx, const double* rb, double* Pb, double* dinv, double* sWt, double* sPp,
                       double* dumps, int N)
{
    constexpr int NX = M::NX, NU = M::NU, NV = NX + NU, NR = NV + 1, NE = NV * (NV + 1) / 2 + NV;
    constexpr int MT = (NV + 7) / 8, NT = (NR + 7) / 8, KS = (NX + 3) / 4, WS = 12, XS = 4;
    constexpr int NTV = NV / 8, TV = (NV % 8) / 2, EV = NV % 2;  // where column NV sits in the fragment layout
    static_assert(NX <= 8 && NV <= 16 && NU <= 2 && MT <= NT, "fragment layout of chainA_mma");
    const int lane = lane_id(), g = lane >> 2, t = lane & 3;
    double* X = sWt + NR * WS;                 // exchange array [8 MT][XS]
    double* slots = X + 8 * MT * XS;           // [0] = 0.0 (read only)
    if (lane == 0) slots[0] = 0.0;
    // dumps: one write-only slot per lane
    ASSUME_SHARED(dumps);
    const saddr zero = smem_addr(slots), dump = smem_addr(dumps + lane);
    // ---- stage-independent addresses
    saddr aP[KS], aWst[NT][2], aWb[NT][KS], aXst[MT][NU], aXr[MT], aXc[NT][2], aPst[MT][NT][2][2];
    const saddr aXv = (t == TV && g < NU) ? smem_addr(X + g * XS + 2) : dump;   // S[j][NV], j < NU, lives in tile (0, NTV)
#pragma unroll
    for (int ks = 0; ks < KS; ks++)
    {
        const int kk = 4 * ks + t;
        aP[ks] = (g < NX && kk < NX) ? smem_addr(sPp + g * WS + kk) : zero;
#pragma unroll
        for (int nt = 0; nt < NT; nt++)
        {
            const int n = 8 * nt + g;
            aWb[nt][ks] = smem_addr(sWt + (n < NR ? n : NR - 1) * WS + kk);   // columns beyond NR feed only unused columns of S
        }
    }
#pragma unroll
    for (int mt = 0; mt < MT; mt++)
    {
        const int row = 8 * mt + g;
        aXr[mt] = smem_addr(X + row * XS);
#pragma unroll
        for (int j = 0; j < NU; j++) aXst[mt][j] = t == 0 ? smem_addr(X + row * XS + j) : dump;
    }
#pragma unroll
    for (int nt = 0; nt < NT; nt++)
#pragma unroll
        for (int e = 0; e < 2; e++)
        {
            const int col = 8 * nt + 2 * t + e;
            aWst[nt][e] = col < NR ? smem_addr(sWt + col * WS + g) : dump;
            aXc[nt][e] = smem_addr(X + (col < NV ? col : 0) * XS);
#pragma unroll
            for (int mt = 0; mt < MT; mt++)
            {
                const int row = 8 * mt + g;
                const bool ok = col >= NU && col <= row && row < NV;
                aPst[mt][nt][e][0] = ok ? smem_addr(sPp + (row - NU) * WS + (col - NU)) : dump;
                aPst[mt][nt][e][1] = ok ? smem_addr(sPp + (col - NU) * WS + (row - NU)) : dump;
            }
        }
    // ---- addresses that move with the stage (set for stage N, decremented at the end of every stage)
    saddr aM[MT][NT][2], aMst[MT][NT][2], aG[NT][KS];
    int dMst[MT][NT][2], dG[NT];
#pragma unroll
    for (int mt = 0; mt < MT; mt++)
#pragma unroll
        for (int nt = 0; nt < NT; nt++)
#pragma unroll
            for (int e = 0; e < 2; e++)
            {
                const int row = 8 * mt + g, col = 8 * nt + 2 * t + e;
                const int r = row < NV ? row : NV - 1, c = col <= NV ? col : NV;           // clamped: any in-bounds entry
                const int hi = r > c ? r : c, lo = r > c ? c : r;
                const saddr a = smem_addr(Mx + N * NE + hi * (hi + 1) / 2 + lo);             // c == NV: the gradient row
                aM[mt][nt][e] = a;
                const bool st = row < NV && (col == NV || col <= row);
                aMst[mt][nt][e] = st ? a : dump;
                dMst[mt][nt][e] = st ? NE * 8 : 0;
            }
#pragma unroll
    for (int nt = 0; nt < NT; nt++)
    {
        const int n = 8 * nt + g, nn = n <= NV ? n : NV;       // rows beyond NR: read res_b again (their W columns are dumped)
        dG[nt] = nn == NV ? NX * 8 : NV * NX * 8;
#pragma unroll
        for (int ks = 0; ks < KS; ks++)
        {
            const int kk = 4 * ks + t, kc = kk < NX ? kk : NX - 1;   // k-padding: P's fragment is an exact zero there
            aG[nt][ks] = nn == NV ? smem_addr(rb + (N - 1) * NX + kc) : smem_addr(G + (N - 1) * (NV * NX) + nn + NV * kc);
        }
    }
    // column NV of W (lanes t == TV, rows g < NX): P res_b goes out as Pb, p_{k+1} (x part of row NV of stage k+1) comes in
    const bool pb_lane = t == TV && g < NX, dinv_lane = lane < NU;
    saddr aPb = pb_lane ? smem_addr(Pb + (N - 1) * NX + g) : dump;
    saddr aPn = pb_lane ? smem_addr(Mx + N * NE + (NV * (NV + 1)) / 2 + NU + g) : zero;
    saddr aDi = dinv_lane ? smem_addr(dinv + N * NU + lane) : dump;
    const int dPb = pb_lane ? NX * 8 : 0, dPn = pb_lane ? NE * 8 : 0, dDi = dinv_lane ? NU * 8 : 0;
    syncwarp();
#pragma unroll 1
    for (int k = N; k >= 0; k--)
    {
        double s[MT][NT][2];
#pragma unroll
        for (int mt = 0; mt < MT; mt++)
#pragma unroll
            for (int nt = 0; nt < NT; nt++)
#pragma unroll
                for (int e = 0; e < 2; e++)
                    if (8 * nt + e <= NV) s[mt][nt][e] = lds64(aM[mt][nt][e]); else s[mt][nt][e] = 0.0;
        if (k < N)
        {
            double pa[KS], gb[NT][KS], w[NT][2];
#pragma unroll
            for (int nt = 0; nt < NT; nt++)
#pragma unroll
                for (int ks = 0; ks < KS; ks++) gb[nt][ks] = lds64(aG[nt][ks]);
            const double pnv = lds64(aPn);
#pragma unroll
            for (int ks = 0; ks < KS; ks++) pa[ks] = lds64(aP[ks]);
#pragma unroll
            for (int nt = 0; nt < NT; nt++)
            {
                w[nt][0] = 0.0; w[nt][1] = 0.0;
#pragma unroll
                for (int ks = 0; ks < KS; ks++) dmma884(w[nt][0], w[nt][1], pa[ks], gb[nt][ks]);
            }
            // column NV of W = P res_b: goes out as Pb, and with p_{k+1} added into the product of step 2
            sts64(aPb, w[NTV][EV]);
            w[NTV][EV] += pnv;
#pragma unroll
            for (int nt = 0; nt < NT; nt++)
#pragma unroll
                for (int e = 0; e < 2; e++)
                    if (8 * nt + e < NR) sts64(aWst[nt][e], w[nt][e]);
            syncwarp();
            double wb[NT][KS];
#pragma unroll
            for (int nt = 0; nt < NT; nt++)
#pragma unroll
                for (int ks = 0; ks < KS; ks++) wb[nt][ks] = lds64(aWb[nt][ks]);
#pragma unroll
            for (int mt = 0; mt < MT; mt++)
#pragma unroll
                for (int ks = 0; ks < KS; ks++)
#pragma unroll
                    for (int nt = 0; nt < NT; nt++)
                        if (8 * nt <= NV) dmma884(s[mt][nt][0], s[mt][nt][1], gb[mt][ks], wb[nt][ks]);
        }
        // ---- exchange: columns 0 .. NU-1 of every row, column NV of the rows of the input block
#pragma unroll
        for (int mt = 0; mt < MT; mt++)
#pragma unroll
            for (int j = 0; j < NU; j++) sts64(aXst[mt][j], s[mt][0][j]);
        sts64(aXv, s[0][NTV][EV]);
        syncwarp();
        double ar[MT][2], ac[NT][2][2], blk[2][2], sp[2];
#pragma unroll
        for (int mt = 0; mt < MT; mt++) lds128(aXr[mt], ar[mt][0], ar[mt][1]);
#pragma unroll
        for (int nt = 0; nt < NT; nt++)
#pragma unroll
            for (int e = 0; e < 2; e++)
                if (8 * nt + e < NV) lds128(aXc[nt][e], ac[nt][e][0], ac[nt][e][1]); else { ac[nt][e][0] = 0.0; ac[nt][e][1] = 0.0; }
#pragma unroll
        for (int j = 0; j < NU; j++)
        {
            lds128(smem_addr(X + j * XS), blk[j][0], blk[j][1]);
            sp[j] = lds64(smem_addr(X + j * XS + 2));
        }
        // ---- Cholesky of the input block, redundantly on every lane (see chainA_impl)
        double Lt[NU][NU], inv[NU];
        if (NU == 2)
        {
            const double m00 = blk[0][0], m10 = blk[NU - 1][0], m11 = blk[NU - 1][NU - 1];
            const double det = m00 * m11 - m10 * m10;
            const double i0 = m00 > 0.0 ? drsqrt_pos(m00) : 0.0;
            double i1;
            if (m00 > 0.0) i1 = det > 0.0 ? drsqrt_pos(det) * (m00 * i0) : 0.0;
            else i1 = m11 > 0.0 ? drsqrt_pos(m11) : 0.0;
            const double l10 = m10 * i0;
            inv[0] = i0; inv[NU - 1] = i1;
            Lt[0][0] = m00 * i0; Lt[NU - 1][0] = l10;
            Lt[NU - 1][NU - 1] = (m11 - l10 * l10) * i1;
        }
        else
        {
            const double piv = blk[0][0];
            inv[0] = piv > 0.0 ? drsqrt_pos(piv) : 0.0;
            Lt[0][0] = piv * inv[0];
        }
        sts64(aDi, lane == 0 ? inv[0] : inv[NU - 1]);
        // ---- Schur complement on the lane's entries
        double vr[MT][NU];
#pragma unroll
        for (int mt = 0; mt < MT; mt++)
#pragma unroll
            for (int j = 0; j < NU; j++)
            {
                double x = ar[mt][j];
#pragma unroll
                for (int i = 0; i < j; i++) x -= vr[mt][i] * Lt[j][i];
                vr[mt][j] = x * inv[j];
            }
#pragma unroll
        for (int nt = 0; nt < NT; nt++)
#pragma unroll
            for (int e = 0; e < 2; e++)
            {
                if (8 * nt + e > NV) continue;     // no lane holds a column of the matrix there
                const int col = 8 * nt + 2 * t + e;
                double uc[NU];
#pragma unroll
                for (int j = 0; j < NU; j++)
                {
                    double y = col == NV ? sp[j] : ac[nt][e][j];
#pragma unroll
                    for (int i = 0; i < j; i++) y -= uc[i] * Lt[j][i];
                    uc[j] = y * inv[j];
                }
#pragma unroll
                for (int mt = 0; mt < MT; mt++)
                {
                    const int row = 8 * mt + g;
                    double out = s[mt][nt][e];
#pragma unroll
                    for (int j = 0; j < NU; j++) out -= vr[mt][j] * uc[j];
#pragma unroll
                    for (int j = 0; j < NU; j++)
                    {
                        out = col == j ? vr[mt][j] : out;                 // factor column j
                        out = (col == NV && row == j) ? uc[j] : out;      // its entry in the gradient row
                    }
                    sts64(aMst[mt][nt][e], out);
                    if (8 * nt + e < NV + 0 && 8 * nt + 6 + e >= NU)
                    {
                        sts64(aPst[mt][nt][e][0], out);
                        sts64(aPst[mt][nt][e][1], out);
                    }
                }
            }
        // ---- next stage
#pragma unroll
        for (int mt = 0; mt < MT; mt++)
#pragma unroll
            for (int nt = 0; nt < NT; nt++)
#pragma unroll
                for (int e = 0; e < 2; e++)
                {
                    aM[mt][nt][e] -= NE * 8;
                    aMst[mt][nt][e] -= dMst[mt][nt][e];
                }
        if (k < N)
        {
#pragma unroll
            for (int nt = 0; nt < NT; nt++)
#pragma unroll
                for (int ks = 0; ks < KS; ks++) aG[nt][ks] -= dG[nt];
            aPb -= dPb;
            aPn -= dPn;
        }
        aDi -= dDi;
        syncwarp();
    }
}

// The two vector recursions of a solve with the factor, on the fp64 tensor cores:
//   FWD : dx_0 = 0,   dx_{k+1} = Acl_k  dx_k    + c_k  -> out[(k+1) * NV + NU + i]               (x_ocp_qp_kkt.c:537-575)
//   !FWD: p_N = e_N,  p_k      = Acl_k' p_{k+1} + e_k  -> out[k * NV + NU + i]                   (x_ocp_qp_kkt.c:1096-1242)
// One warp.  A step is ONE 8 x 8 x 8 product: A = Acl_k (or its transpose) from the field, B = the vector in column 0 (lanes
// g = 0), C = the affine term in column 0; the result's column 0 sits in the accumulators of the lanes t = 0 and travels to
// the next step's B fragment with two shuffles.  ~20 instructions and ~90 cycles per stage -- a mat-vec with FMAs,
// shuffles for the broadcast and its own index arithmetic took 40 / 180 (and a blocked variant of the recursion, 19
// dependent steps instead of N, three times the instructions).  The next stage's operands are loaded before the current
// product (they do not depend on the recursion); the load one stage past the end reads the neighbouring field and is unused.
// slots: [0] = 0.0, [1 + lane] = the lane's dump slot.
template <class M, bool FWD>
MDEVNI void chain_vec_mma(const double* Acl, const double* aff, double* out, double* slots, int N, int NV, int NU)
{
    constexpr int NX = M::NX, KS = (NX + 3) / 4;
    static_assert(NX <= 8, "fragment layout of chain_vec_mma");
    ASSUME_SHARED(Acl); ASSUME_SHARED(aff); ASSUME_SHARED(out); ASSUME_SHARED(slots);
    const int lane = lane_id(), g = lane >> 2, t = lane & 3;
    if (lane == 0) slots[0] = 0.0;
    syncwarp();
    const saddr zero = smem_addr(slots), dump = smem_addr(slots + 1 + lane);
    const int k0 = FWD ? 0 : N - 1, sgn = FWD ? 1 : -1;
    const bool col0 = t == 0 && g < NX;     // the lanes that hold column 0 of the result
    saddr aA[KS], aC, aO;
    int dA[KS];
#pragma unroll
    for (int ks = 0; ks < KS; ks++)
    {
        const int c = 4 * ks + t;
        const bool ok = g < NX && c < NX;
        aA[ks] = ok ? smem_addr(Acl + k0 * (NX * NX) + (FWD ? g * NX + c : c * NX + g)) : zero;
        dA[ks] = ok ? sgn * NX * NX * 8 : 0;
    }
    aC = col0 ? smem_addr(aff + k0 * NX + g) : zero;
    aO = col0 ? smem_addr(out + (FWD ? 1 : N - 1) * NV + NU + g) : dump;
    const int dC = col0 ? sgn * NX * 8 : 0, dO = col0 ? sgn * NV * 8 : 0;
    // the start vector: B fragments (lanes g = 0 hold component 4 ks + t) and its place in `out`
    double b[KS];
#pragma unroll
    for (int ks = 0; ks < KS; ks++)
    {
        const int c = 4 * ks + t;
        b[ks] = (!FWD && g == 0 && c < NX) ? aff[N * NX + c] : 0.0;
    }
    double an[KS], cn;
#pragma unroll
    for (int ks = 0; ks < KS; ks++) an[ks] = lds64(aA[ks]);
    cn = lds64(aC);
    const double v0 = (!FWD && col0) ? aff[N * NX + g] : 0.0;
    if (col0) out[(FWD ? 0 : N) * NV + NU + g] = v0;   // after the first affine term is in a register: `out` may be `aff`
#pragma unroll 1
    for (int k = 0; k < N; k++)
    {
        double a[KS], d0[KS], d1[KS];
#pragma unroll
        for (int ks = 0; ks < KS; ks++) { a[ks] = an[ks]; aA[ks] += dA[ks]; d0[ks] = ks == 0 ? cn : 0.0; d1[ks] = 0.0; }
        aC += dC;
#pragma unroll
        for (int ks = 0; ks < KS; ks++) an[ks] = lds64(aA[ks]);
        cn = lds64(aC);
        // the two halves of the inner dimension as independent products (they pipeline), summed afterwards
#pragma unroll
        for (int ks = 0; ks < KS; ks++) dmma884(d0[ks], d1[ks], a[ks], b[ks]);
        const double y = KS == 2 ? d0[0] + d0[KS - 1] : d0[0];
        sts64(aO, y);
        aO += dO;
        // every column of the next B fragment becomes a copy of the vector: only column 0 is used, the others stay finite
#pragma unroll
        for (int ks = 0; ks < KS; ks++) b[ks] = shfl(y, 4 * (4 * ks + t));
    }
    syncwarp();
}

template <class M, bool SOFT>
struct CtaSolver {
    static constexpr int NX = M::NX, NU = M::NU, NV = NX + NU, NR = NV + 1, NY = NV;
    static constexpr int HXV = NU + M::HX, HYV = NU + M::HY;
    static constexpr int NE = NV * (NV + 1) / 2 + NV;  // entries of the lower trapezoid of the (NV+1) x NV matrix
    static constexpr int NEP = NX * (NX + 1) / 2;
    MDEV static constexpr int MI(int r, int c) { return r * (r + 1) / 2 + c; }  // packed index, c <= r (row NV = gradient row)

    const Params& P;
    int tid, lane, wid, T, W;
    int N, K, nbu, nbx, ncq, ncz, nbq, nct, s2, ns;  // ns: soft rows (the first ns rows of h); 0 unless SOFT
    FastDiv dq, ds2;
    double *sm, *gs, *w;
    // constants in shared memory
    double *Hs, *Hes, *Tp, *red, *sA0, *sW, *sP, *slots;
    double** ftab;
    int *sxrow, *srvar;
    int redbuf;
    // working-set fields: no pointer is kept in registers; every access forms the address from the block's shared-memory
    // base (or its global scratch) and the offset in the kernel parameters (constant bank), see the accessors below
    // IPM arguments (HP/ocp_qp/x_ocp_qp_ipm.c:133-161 overridden by AC/acados/ocp_qp/ocp_qp_hpipm.c:106-116 and, in SQP
    // mode, by AC/acados/ocp_nlp/ocp_nlp_sqp.c:201-227)
    double tol_stat, tol_eq, tol_ineq, tol_comp;
    int iter_max;
    // IPM state (block-uniform)
    double res_max[4], mu, mu_aff, sigma, alpha;
    double S1, S2;        // sum(lam*dt + t*dlam), sum(dlam*dt) of the last expanded step
    double lin_d, lin_m, lin_g;  // residual norms of the linearised inequality / complementarity / slack-stationarity rows at the last expanded step
    int solve_calls, lq_count, itref_count, fp32_count;
    bool use_fp32;  // this factorisation runs in fp32 (config 4); cleared for the rest of a QP once its accuracy test fails
#ifdef USVMPC_PROFILE
    // phase clocks of thread 0 (diagnostic builds only)
    long long prof[24], tprev;
#define PROF(i) { if (tid == 0) { const long long t_ = clock_now(); prof[i] += t_ - tprev; tprev = t_; } }
#else
#define PROF(i)
#endif

    MDEV double* fld(int id) const
    {
        const SField& f = P.plan.f[id];
        if (id < F_FIRST_FLEX) return sm + f.off;  // the chain fields are always in shared memory
        return ftab[id];                            // resolved once per block (shared memory or the block's scratch)
    }
    MDEV double* G_() const { double* p_ = sm + P.plan.f[F_G].off; ASSUME_SHARED(p_); return p_; }
    MDEV double* Mx_() const { double* p_ = sm + P.plan.f[F_M].off; ASSUME_SHARED(p_); return p_; }
    MDEV double* Acl_() const { double* p_ = sm + P.plan.f[F_ACL].off; ASSUME_SHARED(p_); return p_; }
    MDEV double* Kg_() const { double* p_ = sm + P.plan.f[F_KG].off; ASSUME_SHARED(p_); return p_; }
    MDEV double* cc_() const { double* p_ = sm + P.plan.f[F_CC].off; ASSUME_SHARED(p_); return p_; }
    MDEV double* ee_() const { double* p_ = sm + P.plan.f[F_EE].off; ASSUME_SHARED(p_); return p_; }
    MDEV double* Pb_() const { double* p_ = sm + P.plan.f[F_PB].off; ASSUME_SHARED(p_); return p_; }
    MDEV double* rb_() const { double* p_ = sm + P.plan.f[F_RB].off; ASSUME_SHARED(p_); return p_; }
    MDEV double* zv_() const { double* p_ = sm + P.plan.f[F_ZV].off; ASSUME_SHARED(p_); return p_; }
    MDEV double* dux_() const { double* p_ = sm + P.plan.f[F_DUX].off; ASSUME_SHARED(p_); return p_; }
    MDEV double* kk_() const { double* p_ = sm + P.plan.f[F_KK].off; ASSUME_SHARED(p_); return p_; }
    MDEV double* dinv_() const { double* p_ = sm + P.plan.f[F_DINV].off; ASSUME_SHARED(p_); return p_; }
    MDEV double* ux_() const { return fld(F_UX); }
    MDEV double* pi_() const { return fld(F_PI); }
    MDEV double* lam_() const { return fld(F_LAM); }
    MDEV double* t_() const { return fld(F_T); }
    MDEV double* dlam_() const { return fld(F_DLAM); }
    MDEV double* dt_() const { return fld(F_DT); }
    MDEV double* rd_() const { return fld(F_RD); }
    MDEV double* rmc_() const { return fld(F_RMC); }
    MDEV double* rg_() const { return fld(F_RG); }
    MDEV double* dpi_() const { return fld(F_DPI); }
    MDEV double* gxy_() const { return fld(F_GXY); }
    MDEV double* d_() const { return fld(F_D); }
    MDEV double* rq_() const { return fld(F_RQ); }
    MDEV double* b_() const { return fld(F_B); }
    MDEV double* sv_() const { return fld(F_SV); }
    MDEV double* dsv_() const { return fld(F_DSV); }
    MDEV double* rgs_() const { return fld(F_RGS); }
    MDEV double* zsi_() const { return fld(F_ZSI); }
    MDEV double* rqs_() const { return fld(F_RQS); }
    MDEV double* dsv2_() const { return fld(F_DSV2); }
    MDEV double* rgs2_() const { return fld(F_RGS2); }
    MDEV double* rg2_() const { return fld(F_RG2); }
    MDEV double* rb2_() const { return fld(F_RB2); }
    MDEV double* rd2_() const { return fld(F_RD2); }
    MDEV double* rm2_() const { return fld(F_RM2); }
    MDEV double* dux2_() const { return fld(F_DUX2); }
    MDEV double* dpi2_() const { return fld(F_DPI2); }
    MDEV double* dlam2_() const { return fld(F_DLAM2); }
    MDEV double* dt2_() const { return fld(F_DT2); }

    MDEV CtaSolver(const Params& p, double* smem, double* gscratch) : P(p), sm(smem), gs(gscratch)
    {
        tid = thread_id(); lane = tid & 31; wid = tid >> 5; T = block_threads(); W = T >> 5;
        N = P.N; K = P.K; nbu = P.nbu; nbx = P.nbx; ncq = P.ncq; ncz = P.ncz; nbq = nbu + nbx;
        ns = SOFT ? P.ns : 0;
        s2 = 2 * ncq + 2 * ns;
        nct = N >= 1 ? 2 * ((nbu + K + ns) + (N - 1) * (nbu + nbx + K + ns)) : 0;
        dq.set(ncq); ds2.set(s2);
        double* c = sm + P.plan.const_off;
        Hs = c; c += NV * NV; Hes = c; c += NV * NV; Tp = c;
        red = sm + P.plan.red_off;
        redbuf = 0;
        double* m = sm + P.plan.misc_off;
        sA0 = m; m += NV * NX; sW = m; m += chain_w_doubles(NX, NU); sP = m; m += chain_p_doubles(NX);  // chain scratch: [P G' | matrix], P
        slots = m; m += 34;   // [0] = 0.0 and one dump slot per lane for the recursions' unpredicated stores
        ftab = (double**) m; m += F_COUNT;   // field id -> address; filled here, first read after load_constants' barrier
        for (int id = tid; id < F_COUNT; id += T) ftab[id] = (P.plan.f[id].space ? gs : sm) + P.plan.f[id].off;
        sxrow = (int*) m; srvar = sxrow + NX + (NX & 1);
        // address-space hints: these always point into shared memory
        ASSUME_SHARED(sm); ASSUME_SHARED(Hs); ASSUME_SHARED(Hes); ASSUME_SHARED(Tp);
        ASSUME_SHARED(red); ASSUME_SHARED(sA0); ASSUME_SHARED(sW); ASSUME_SHARED(sP); ASSUME_SHARED(sxrow); ASSUME_SHARED(srvar);
        ASSUME_SHARED(slots); ASSUME_SHARED(ftab);
        tol_stat = 1e-6; tol_eq = 1e-8; tol_ineq = 1e-8; tol_comp = 1e-8;
        if (P.nlp_type == 0) { tol_stat = P.tol[0]; tol_eq = P.tol[1]; tol_ineq = P.tol[2]; tol_comp = P.tol[3]; }
        iter_max = P.qp_iter_max > 0 ? P.qp_iter_max : 50;
        solve_calls = 0; lq_count = 0; itref_count = 0; fp32_count = 0; use_fp32 = false;
    }

    MDEV double* Z(const Field& f, int k) const { return w + f.off + (long) k * f.stride; }
    MDEV bool var_active(int k, int i) const { return k == 0 ? (i < NU) : (k == N ? (i >= NU) : true); }
    MDEV bool row_active(int k, int j) const { return k < N && (j < nbu || j >= nbq || k >= 1); }
    // element e of a row array -> stage k; is the row live?  (lower side | upper side | slack-bound rows)
    MDEV bool elem_active(int e, int& k) const
    {
        k = ds2.div(e);
        int j = e - k * s2;
        if (j >= 2 * ncq) return k < N;
        if (j >= ncq) j -= ncq;
        return row_active(k, j);
    }
    // soft row?  (h row c = j - nbq < ns)  -> slack index or -1
    MDEV int soft_index(int j) const { return (SOFT && j >= nbq && j - nbq < ns) ? j - nbq : -1; }
    // slack penalties of slack `is`, scaled like the stage cost (ocp_nlp_cost_ls.c:733,826-841): side 0 lower, 1 upper
    MDEV double zlin(int is, int side) const { return P.dt * P.zs[side * ns + is]; }
    MDEV double zquad(int is, int side) const { return P.dt * P.zs[(2 + side) * ns + is]; }
    // IPM row of the box on variable c at stage k, or -1
    MDEV int vrow(int k, int c) const
    {
        if (k >= N) return -1;
        if (c < NU) return c < nbu ? c : -1;
        return k >= 1 ? sxrow[c - NU] : -1;
    }
    MDEV int stage_class(int k) const { return k == 0 ? 0 : (k < N ? 1 : 2); }
    MDEV const double* Hk(int k) const { return k < N ? Hs : Hes; }

    // ---------------------------------------------------------------- block-wide reductions (every thread gets the result)
    // The first NSG entries of vmax may have either sign (step-length ratios), the others are norms (>= 0, never -0 or NaN).
    template <int NM, int NS, int NSG = 0>
    MDEV void block_reduce(double* vmax, double* vsum)
    {
#pragma unroll
        for (int i = 0; i < NM; i++) vmax[i] = i < NSG ? warp_max(vmax[i]) : warp_max_nonneg(vmax[i]);
#pragma unroll
        for (int i = 0; i < NS; i++) vsum[i] = warp_sum(vsum[i]);
        double* r = red + redbuf * W * 8;
        redbuf ^= 1;
        if (lane == 0)
        {
#pragma unroll
            for (int i = 0; i < NM; i++) r[wid * 8 + i] = vmax[i];
#pragma unroll
            for (int i = 0; i < NS; i++) r[wid * 8 + NM + i] = vsum[i];
        }
        syncthreads();
        // every warp reduces the W per-warp maxima again (lane l holds warp l mod W's value)
        const int src = (lane < W ? lane : lane % W) * 8;
#pragma unroll
        for (int i = 0; i < NM; i++) vmax[i] = i < NSG ? warp_max(r[src + i]) : warp_max_nonneg(r[src + i]);
        // sums in warp order so that every thread adds the same numbers in the same order
        if (NS > 0)
        {
            double s[NS > 0 ? NS : 1];
#pragma unroll
            for (int i = 0; i < NS; i++) s[i] = 0.0;
            for (int q = 0; q < W; q++)
            {
#pragma unroll
                for (int i = 0; i < NS; i++) s[i] += r[q * 8 + NM + i];
            }
#pragma unroll
            for (int i = 0; i < NS; i++) vsum[i] = s[i];
        }
    }

    // ---------------------------------------------------------------- constants into shared memory
    // Gauss-Newton Hessians: ocp_nlp_cost_ls_initialize, AC/acados/ocp_nlp/ocp_nlp_cost_ls.c:713-745.  With
    // Vx=[I;0], Vu=[0;I] the output map y = Cyt'[u;x] is the permutation [x;u].  Lower triangles are read.
    MDEV void load_constants()
    {
        const double* Wg = P.cst;
        const double* Weg = P.cst + NY * NY;
        for (int e = tid; e < NV * NV; e += T)
        {
            const int i = e % NV, j = e / NV;
            const int yi = i < NU ? NX + i : i - NU, yj = j < NU ? NX + j : j - NU;
            Hs[e] = P.dt * (yi >= yj ? Wg[yi + NY * yj] : Wg[yj + NY * yi]);
            Hes[e] = (i >= NU && j >= NU) ? ((i >= j) ? Weg[(i - NU) + NX * (j - NU)] : Weg[(j - NU) + NX * (i - NU)]) : 0.0;
        }
        if (tid < NX)
        {
            int r = -1;
            for (int j = 0; j < nbx; j++) if (P.idxbx[j] == tid) r = nbu + j;
            sxrow[tid] = r;
        }
        for (int jj = tid; jj < ncq; jj += T) srvar[jj] = jj < nbu ? jj : (jj < nbq ? NU + P.idxbx[jj - nbu] : -1);
        syncthreads();
        // templates of the matrix to factorise per stage class (0: stage 0, 1: path, 2: terminal): Hessian + reg_prim on
        // the active variables, identity on the inactive ones
        for (int e = tid; e < 3 * NE; e += T)
        {
            const int cls = e / NE, q = e - cls * NE;
            int r = 0;
            while ((r + 1) * (r + 2) / 2 <= q && r < NV) r++;
            const int c = q - r * (r + 1) / 2;
            const double* H = cls == 2 ? Hes : Hs;
            double v = 0.0;
            if (r < NV)
            {
                const bool ar = cls == 1 || (cls == 0 ? r < NU : r >= NU), ac = cls == 1 || (cls == 0 ? c < NU : c >= NU);
                if (ar && ac) { v = H[r + NV * c]; if (c == r) v += 1e-15; }  // reg_prim
                else if (c == r) v = 1.0;
            }
            Tp[e] = v;
        }
        syncthreads();
    }

    // ---------------------------------------------------------------- initial guess
    // cold start of the scripts / template (acados_solver.in.c:1595-1623): x_k = x0, u = 0, pi = 0;
    // lam, t start at zero like a freshly created nlp_out.
    MDEV void cold_start(const double* x0)
    {
        for (int k = tid; k <= N; k += T)
        {
            double* z = Z(P.lay.zux, k);
            for (int i = 0; i < NU; i++) z[i] = 0.0;
            for (int i = 0; i < NX; i++) z[NU + i] = x0[i];
            double* p = Z(P.lay.zpi, k);
            for (int i = 0; i < NX; i++) p[i] = 0.0;
            double* l = Z(P.lay.zlam, k); double* tt = Z(P.lay.zt, k);
            for (int j = 0; j < 2 * ncz + 2 * ns; j++) { l[j] = 0.0; tt[j] = 0.0; }
            if (SOFT) { double* zs = Z(P.lay.zsv, k); for (int j = 0; j < 2 * ns; j++) zs[j] = 0.0; }
        }
        syncthreads();
    }

    // ---------------------------------------------------------------- linearisation
    // ERK with forward sensitivities (AC/acados/sim/sim_erk_integrator.c:762-847, tableaus :253-344; seed S=[I 0],
    // A=Sx(T), B=Su(T): ocp_nlp_dynamics_cont.c:782-804).  One thread integrates [x ; one sensitivity column] of one
    // stage.
    MDEV void integrate_all()
    {
        const int ns = P.num_stages;
        double a21 = 0, a32 = 0, a43 = 0, bv0 = 0, bv1 = 0, bv2 = 0, bv3 = 0;
        if (ns == 1) { bv0 = 1.0; }
        else if (ns == 2) { a21 = 0.5; bv1 = 1.0; }
        else { a21 = 0.5; a32 = 0.5; a43 = 1.0; bv0 = 1.0 / 6.0; bv1 = 1.0 / 3.0; bv2 = 1.0 / 3.0; bv3 = 1.0 / 6.0; }
        const double step = P.dt / P.num_steps;
        // columns of states that do not enter the dynamics stay unit vectors exactly (their VDE right-hand side is
        // Jx e_c = 0): no thread integrates them, the first task of the stage writes them
        constexpr int NCOL = NV - M::NKIN;
        for (int task = tid; task < N * NCOL; task += T)
        {
            const int k = task / NCOL, col = M::NKIN + task % NCOL;
            const double* z = Z(P.lay.zux, k);
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
            // G = [B'; A'] (nv x nx, column-major): ocp_nlp_dynamics_cont.c:801-804
            double* Gk = G_() + k * (NV * NX);
            const int row = col < NX ? NU + col : col - NX;
#pragma unroll
            for (int i = 0; i < NX; i++) Gk[row + NV * i] = s[i];
            if (col == M::NKIN)
            {
#pragma unroll
                for (int c = 0; c < M::NKIN; c++)
                {
#pragma unroll
                    for (int i = 0; i < NX; i++) Gk[NU + c + NV * i] = (i == c) ? 1.0 : 0.0;
                }
                const double* zn = Z(P.lay.zux, k + 1);
                double* bk = b_() + k * NX;
#pragma unroll
                for (int i = 0; i < NX; i++) bk[i] = x[i] - zn[NU + i];  // dyn_fun = phi(x,u) - x_next
            }
        }
        syncthreads();
    }

    // cost / constraints / adjoints / NLP residuals / QP vectors, one thread per stage.
    // ocp_nlp_approximate_qp_matrices + _vectors_sqp (ocp_nlp_common.c:1926-2084), ocp_nlp_res_compute (:2549-2603),
    // x0 elimination d_ocp_qp_reduce_eq_dof (HP/ocp_qp/x_ocp_qp_red.c:268-454).  res4 = (stat, eq, ineq, comp).
    // SQP_RTI with rti_phase 1 / 2 (ocp_nlp_sqp_rti.c:459-488): the preparation phase runs the integrator with its
    // sensitivities -- the expensive part of the linearisation -- and parks [B';A'] and phi(x,u) - x_next in the
    // instance's HBM block; the feedback phase picks them up, evaluates cost / constraints / x0 embedding with the x0
    // that has arrived meanwhile, solves the QP and updates.
    MDEV void prep_store(int inst)
    {
        double* g = P.prep + (long) inst * N * (NV * NX + NX);
        for (int e = tid; e < N * NV * NX; e += T) g[e] = G_()[e];
        for (int e = tid; e < N * NX; e += T) g[N * NV * NX + e] = b_()[e];
        syncthreads();
    }
    MDEV void prep_load(int inst)
    {
        const double* g = P.prep + (long) inst * N * (NV * NX + NX);
        for (int e = tid; e < N * NV * NX; e += T) G_()[e] = g[e];
        for (int e = tid; e < N * NX; e += T) b_()[e] = g[N * NV * NX + e];
        syncthreads();
    }

    MDEV void linearize(int inst, const double* x0, const double* pg, const double* lhg, const double* yrg, const double* yre,
                       double* res4)
    {
        if (P.nlp_type == 1 && P.rti_phase == 2) prep_load(inst); else integrate_all();
        const double* Wg = P.cst;                 // cost weights: read from global (L2) once per SQP iteration
        const double* Weg = P.cst + NY * NY;
        double r0 = 0, r1 = 0, r2 = 0, r3 = 0;
        for (int k = tid; k <= N; k += T)
        {
            const double* z = Z(P.lay.zux, k);
            const double* zl = Z(P.lay.zlam, k);
            const double* zt = Z(P.lay.zt, k);
            double* zf = Z(P.lay.zfun, k);
            double* rqk = rq_() + k * NV;
            double* dk = d_() + k * s2;
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
                    for (int j = 0; j < NY; j++) acc += (i >= j ? Wg[i + NY * j] : Wg[j + NY * i]) * r[j];   // lower triangle of W (global, L2)
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
                    for (int j = 0; j < NX; j++) acc += (i >= j ? Weg[i + NX * j] : Weg[j + NX * i]) * r[j];
                    cg[NU + i] = acc;
                }
            }
#pragma unroll
            for (int i = 0; i < NV; i++) adj[i] = 0.0;
            // ---- BGH constraints (ocp_nlp_constraints_bgh.c:1228-1430): fun = [lb - g ; g - ub], adj = J'(lam_l - lam_u)
            for (int j = 0; j < 2 * ncz + 2 * ns; j++) zf[j] = 0.0;
            for (int j = 0; j < s2; j++) dk[j] = 0.0;
            double dx0[NX];
#pragma unroll
            for (int i = 0; i < NX; i++) dx0[i] = 0.0;
            if (k < N)
            {
                for (int j = 0; j < nbu; j++)
                {
                    const double g = z[j], fl = P.lbu[k * nbu + j] - g, fu = g - P.ubu[k * nbu + j];
                    zf[j] = fl; zf[ncz + j] = fu; dk[j] = fl; dk[ncq + j] = fu;
                    const double dl = zl[j] - zl[ncz + j];
#pragma unroll
                    for (int i = 0; i < NU; i++) if (i == j) adj[i] += dl;
                    const double a = dabs(fl + zt[j]), bb = dabs(fu + zt[ncz + j]);
                    r2 = a > r2 ? a : r2; r2 = bb > r2 ? bb : r2;
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
                        const double a = dabs(fl + zt[r]), bb = dabs(fu + zt[ncz + r]);
                        r2 = a > r2 ? a : r2; r2 = bb > r2 ? bb : r2;
                        const double c0 = dabs(zl[r] * zt[r]), c1 = dabs(zl[ncz + r] * zt[ncz + r]);
                        r3 = c0 > r3 ? c0 : r3; r3 = c1 > r3 ? c1 : r3;
                    }
                }
                else
                {
                    for (int j = 0; j < nbx; j++)
                    {
                        const int id = P.idxbx[j], r = nbu + j;
                        const double g = z[NU + id], fl = P.lbx[k * nbx + j] - g, fu = g - P.ubx[k * nbx + j];
                        zf[r] = fl; zf[ncz + r] = fu; dk[r] = fl; dk[ncq + r] = fu;
                        const double dl = zl[r] - zl[ncz + r];
#pragma unroll
                        for (int i = 0; i < NX; i++) if (i == id) adj[NU + i] += dl;
                        const double a = dabs(fl + zt[r]), bb = dabs(fu + zt[ncz + r]);
                        r2 = a > r2 ? a : r2; r2 = bb > r2 ? bb : r2;
                        const double c0 = dabs(zl[r] * zt[r]), c1 = dabs(zl[ncz + r] * zt[ncz + r]);
                        r3 = c0 > r3 ? c0 : r3; r3 = c1 > r3 ? c1 : r3;
                    }
                }
                // obstacle distances h_c = ||(X,Y) - (ox_c, oy_c)||, dh/d(X,Y) = ((X,Y) - o_c)/h_c
                const double* pk = pg + (P.p_per_stage ? k * 2 * K : 0);
                const double* lhk = lhg + (P.lh_per_stage ? k * K : 0);
                double* gk = gxy_() + k * 2 * K;
                for (int c = 0; c < K; c++)
                {
                    const double ddx = z[HXV] - pk[2 * c], ddy = z[HYV] - pk[2 * c + 1];
                    const double h = dsqrt(ddx * ddx + ddy * ddy);
                    const double gX = ddx / h, gY = ddy / h;
                    gk[c] = gX; gk[K + c] = gY;
                    double fl = lhk[c] - h, fu = h - P.uh[k * K + c];
                    const int r = nbu + NX + c, rqp = nbq + c;
                    if (SOFT && c < ns)
                    {
                        // soft row (ocp_nlp_constraints_bgh.c:1404-1427, ocp_nlp_cost_ls.c:826-841): the slacks enter the row, have
                        // their own bound rows fun = bound - slack, cost gradient scaling (z + Z s) and adjoint lam_row + lam_bound
                        const double* zs = Z(P.lay.zsv, k);
#pragma unroll
                        for (int side = 0; side < 2; side++)
                        {
                            const double sj = zs[side * ns + c];
                            if (side) fu -= sj; else fl -= sj;
                            const int rs = 2 * ncz + side * ns + c;
                            const double fs = (side ? P.ush[k * ns + c] : P.lsh[k * ns + c]) - sj;
                            zf[rs] = fs;
                            dk[2 * ncq + side * ns + c] = fs;
                            const double cgs = P.dt * (P.zs[side * ns + c] + P.zs[(2 + side) * ns + c] * sj);
                            rqs_()[k * 2 * ns + side * ns + c] = cgs;
                            const double adjs = zl[side * ncz + r] + zl[rs];
                            double a = dabs(cgs - adjs); r0 = a > r0 ? a : r0;
                            a = dabs(fs + zt[rs]); r2 = a > r2 ? a : r2;
                            a = dabs(zl[rs] * zt[rs]); r3 = a > r3 ? a : r3;
                        }
                    }
                    zf[r] = fl; zf[ncz + r] = fu;
                    // stage 0: fold the eliminated x0 step into the bounds (x_ocp_qp_red.c:380-420)
                    const double v = (k == 0) ? gX * dx0[M::HX] + gY * dx0[M::HY] : 0.0;
                    dk[rqp] = fl - v; dk[ncq + rqp] = fu + v;
                    const double dl = zl[r] - zl[ncz + r];
                    adj[HXV] += gX * dl; adj[HYV] += gY * dl;
                    const double a = dabs(fl + zt[r]), bb = dabs(fu + zt[ncz + r]);
                    r2 = a > r2 ? a : r2; r2 = bb > r2 ? bb : r2;
                    const double c0 = dabs(zl[r] * zt[r]), c1 = dabs(zl[ncz + r] * zt[ncz + r]);
                    r3 = c0 > r3 ? c0 : r3; r3 = c1 > r3 ? c1 : r3;
                }
            }
            // ---- dynamics adjoint -[B';A'] pi_k (+ pi_{k-1} on x): ocp_nlp_common.c:2001-2019 ; stationarity residual
            const double* Gk = G_() + k * (NV * NX);
            const double* pik = Z(P.lay.zpi, k);
#pragma unroll
            for (int i = 0; i < NV; i++)
            {
                double acc = 0.0;
                if (k < N)
                {
#pragma unroll
                    for (int j = 0; j < NX; j++) acc -= Gk[i + NV * j] * pik[j];
                }
                if (k > 0 && i >= NU) acc += Z(P.lay.zpi, k - 1)[i - NU];
                if (k < N || i >= NU)
                {
                    const double a = dabs(cg[i] - adj[i] - acc);
                    r0 = a > r0 ? a : r0;
                }
            }
            if (k < N)
            {
                const double* bk = b_() + k * NX;
#pragma unroll
                for (int i = 0; i < NX; i++) { const double a = dabs(bk[i]); r1 = a > r1 ? a : r1; }
            }
            // ---- QP gradient; stage 0: b0 += A0' dx0, r0 += S dx0 (x_ocp_qp_red.c:300-378)
#pragma unroll
            for (int i = 0; i < NV; i++) rqk[i] = cg[i];
            if (k == 0 && N > 0)
            {
                double* bk = b_();
#pragma unroll
                for (int j = 0; j < NX; j++)
                {
                    double acc = bk[j];
#pragma unroll
                    for (int i = 0; i < NX; i++) acc += Gk[NU + i + NV * j] * dx0[i];
                    bk[j] = acc;
                }
#pragma unroll
                for (int i = 0; i < NU; i++)
                {
                    double acc = cg[i];
#pragma unroll
                    for (int j = 0; j < NX; j++) acc += Hs[(NU + j) + NV * i] * dx0[j];
                    rqk[i] = acc;
                }
                // stage 0 after x0 elimination has no x rows in [B';A'] (x_ocp_qp_red.c:268-454): keep the unmasked block
                // for the restoration of the stage-0 multipliers, mask the working copy
                for (int e = 0; e < NV * NX; e++)
                {
                    sA0[e] = Gk[e];
                    if (e % NV >= NU) G_()[e] = 0.0;
                }
            }
        }
        double vm[4] = {r0, r1, r2, r3};
        block_reduce<4, 0>(vm, nullptr);
        res4[0] = vm[0]; res4[1] = vm[1]; res4[2] = vm[2]; res4[3] = vm[3];
    }

    // ---------------------------------------------------------------- IPM: start point
    // OCP_QP_INIT_VAR, var_init_scheme 1, cold start: HP/ocp_qp/x_ocp_qp_ipm.c:1435-1470,1581-1714 (ns = 0)
    MDEV void ipm_init()
    {
        const double thr0 = 1e-1, mu0 = 1.0;
        for (int k = tid; k <= N; k += T)
        {
            double* v = ux_() + k * NV; double* pk = pi_() + k * NX; double* l = lam_() + k * s2; double* tt = t_() + k * s2;
            const double* dk = d_() + k * s2;
            for (int i = 0; i < NV; i++) v[i] = 0.0;
            for (int i = 0; i < NX; i++) pk[i] = 0.0;
            for (int j = 0; j < s2; j++) { l[j] = 0.0; tt[j] = 1.0; }
            {
                // the first passA applies a zero step: the step starts at zero
                double* a = dux_() + k * NV; double* bq = dpi_() + k * NX; double* c = dlam_() + k * s2; double* e = dt_() + k * s2;
                for (int i = 0; i < NV; i++) a[i] = 0.0;
                for (int i = 0; i < NX; i++) bq[i] = 0.0;
                for (int j = 0; j < s2; j++) { c[j] = 0.0; e[j] = 0.0; }
            }
            if (SOFT)
            {
                // slack variables start at zero, lifted onto their lower bound + thr0 where needed (x_ocp_qp_ipm.c:1610-1626)
                double* sv = sv_() + k * 2 * ns; double* dsv = dsv_() + k * 2 * ns;
                for (int j = 0; j < 2 * ns; j++)
                {
                    double sj = 0.0;
                    if (k < N)
                    {
                        double ts = sj - dk[2 * ncq + j];
                        if (ts < thr0) { ts = thr0; sj = dk[2 * ncq + j] + ts; }
                        tt[2 * ncq + j] = ts;
                    }
                    sv[j] = sj; dsv[j] = 0.0;
                }
            }
            if (k >= N) continue;
            for (int j = 0; j < nbq; j++)
            {
                if (!row_active(k, j)) continue;
                const int id = srvar[j];
                double tl = v[id] - dk[j], tu = -v[id] - dk[ncq + j];
                if (tl < thr0)
                {
                    if (tu < thr0) { v[id] = 0.5 * (dk[j] - dk[ncq + j]); tl = thr0; tu = thr0; }
                    else { tl = thr0; v[id] = dk[j] + thr0; }
                }
                else if (tu < thr0) { tu = thr0; v[id] = -dk[ncq + j] - thr0; }
                tt[j] = tl; tt[ncq + j] = tu;
            }
            const double* gk = gxy_() + k * 2 * K;
            for (int c = 0; c < K; c++)
            {
                const double vv = (k >= 1) ? gk[c] * v[HXV] + gk[K + c] * v[HYV] : 0.0;
                const int is = soft_index(nbq + c);
                const double sl = is >= 0 ? sv_()[k * 2 * ns + is] : 0.0, su = is >= 0 ? sv_()[k * 2 * ns + ns + is] : 0.0;
                const double tl = (vv + sl) - dk[nbq + c], tu = (-vv + su) - dk[ncq + nbq + c];
                tt[nbq + c] = thr0 > tl ? thr0 : tl;
                tt[ncq + nbq + c] = thr0 > tu ? thr0 : tu;
            }
            for (int j = 0; j < ncq; j++)
                if (row_active(k, j)) { l[j] = mu0 / tt[j]; l[ncq + j] = mu0 / tt[ncq + j]; }
            for (int j = 2 * ncq; j < s2; j++) l[j] = mu0 / tt[j];
        }
        syncthreads();
    }

    // ---------------------------------------------------------------- IPM: passes
    // passA: UPDATE_VAR_QP (x_core_qp_ipm_aux.c:220-325, split_step = 0) + OCP_QP_RES_COMPUTE (x_ocp_qp_res.c:336-466)
    // + COMPUTE_GAMMA_GAMMA_QP (x_core_qp_ipm_aux.c:38-86) for res_m = lam*t - tau + assembly of the matrix to factorise
    // (H + Gamma terms, gradient row) in the M field of every stage.
    MDEV void passA(double a, double tau, double* n4)
    {
        const double lam_min = 1e-16, t_min = 1e-16;
        // ---- A1: variable update, element-parallel
        for (int e = tid; e < (N + 1) * NV; e += T) ux_()[e] += a * dux_()[e];
        for (int e = tid; e < N * NX; e += T) pi_()[e] += a * dpi_()[e];
        if (SOFT) for (int e = tid; e < N * 2 * ns; e += T) sv_()[e] += a * dsv_()[e];
        // (rows that do not exist -- the x boxes of stage 0 -- hold lam = 0, t = 1 and a zero step since ipm_init: no test)
        for (int e = tid; e < N * s2; e += T)
        {
            double x = lam_()[e] + a * dlam_()[e];
            lam_()[e] = x <= lam_min ? lam_min : x;
            x = t_()[e] + a * dt_()[e];
            t_()[e] = x <= t_min ? t_min : x;
        }
        syncthreads();
        PROF(18)
        // ---- A2: inequality rows, one (stage, row pair) per thread: res_d, res_m norms, mu, 1/t, and the row's
        // contributions (lam_u - lam_l, Gamma_l + Gamma_u, gamma_l - gamma_u) for A3, parked in the (dead) step arrays
        double n0 = 0, n1 = 0, n2 = 0, n3 = 0, musum = 0;
        for (int it = tid; it < N * ncq; it += T)
        {
            const int k = dq.div(it), j = it - k * ncq;
            const int r0 = k * s2 + j, r1 = r0 + ncq;
            if (!row_active(k, j)) { dlam_()[r0] = 0.0; dlam_()[r1] = 0.0; dt_()[r0] = 0.0; continue; }
            const double* v = ux_() + k * NV;
            double vv;
            if (j < nbq) vv = v[srvar[j]];
            else vv = k >= 1 ? gxy_()[k * 2 * K + j - nbq] * v[HXV] + gxy_()[k * 2 * K + K + j - nbq] * v[HYV] : 0.0;
            const double l0 = lam_()[r0], l1 = lam_()[r1], t0 = t_()[r0], t1 = t_()[r1];
            const int is = soft_index(j);
            double rd0 = d_()[r0] + t0 - vv, rd1 = d_()[r1] + t1 + vv;
            if (is >= 0) { rd0 -= sv_()[k * 2 * ns + is]; rd1 -= sv_()[k * 2 * ns + ns + is]; }  // soft row: the slack enters the row
            rd_()[r0] = rd0; rd_()[r1] = rd1;
            const double m0 = l0 * t0, m1 = l1 * t1;
            musum += m0; musum += m1;
            double q = dabs(m0); n3 = q > n3 ? q : n3; q = dabs(m1); n3 = q > n3 ? q : n3;
            q = dabs(rd0); n2 = q > n2 ? q : n2; q = dabs(rd1); n2 = q > n2 ? q : n2;
            const double ti0 = 1.0 / t0, ti1 = 1.0 / t1;
            double G0 = ti0 * l0, G1 = ti1 * l1;
            double g0 = ti0 * ((m0 - tau) - l0 * rd0), g1 = ti1 * ((m1 - tau) - l1 * rd1);
            if (is >= 0)
            {
                // the row's two slack variables and their bound rows: residuals (x_ocp_qp_res.c:416-438), Gamma / gamma, and
                // the elimination of the slacks COND_SLACKS_FACT_SOLVE (x_ocp_qp_kkt.c:220-291)
#pragma unroll
                for (int side = 0; side < 2; side++)
                {
                    const int rs = k * s2 + 2 * ncq + side * ns + is, si = k * 2 * ns + side * ns + is;
                    const double ls = lam_()[rs], ts = t_()[rs], sj = sv_()[si];
                    const double rds = d_()[rs] + ts - sj, ms = ls * ts;
                    rd_()[rs] = rds;
                    musum += ms;
                    q = dabs(ms); n3 = q > n3 ? q : n3;
                    q = dabs(rds); n2 = q > n2 ? q : n2;
                    const double tis = 1.0 / ts;
                    const double Gsl = tis * ls, gsl = tis * ((ms - tau) - ls * rds);
                    const double Z = zquad(is, side);
                    const double rgs = Z * sj + rqs_()[si] - ls - (side ? l1 : l0);
                    rgs_()[si] = rgs;
                    q = dabs(rgs); n0 = q > n0 ? q : n0;
                    const double Gr = side ? G1 : G0, gr = side ? g1 : g0;
                    const double zi = 1.0 / (Z + 1e-15 + Gr + Gsl);
                    zsi_()[si] = zi;
                    const double rhs = rgs + gr + gsl;
                    dsv_()[si] = rhs;
                    const double tmp = rhs * zi;
                    if (side) { G1 = Gr - Gr * zi * Gr; g1 = gr - Gr * tmp; } else { G0 = Gr - Gr * zi * Gr; g0 = gr - Gr * tmp; }
                }
            }
            dlam_()[r0] = l1 - l0;
            dlam_()[r1] = G0 + G1;
            dt_()[r0] = g0 - g1;
        }
        PROF(19)
        // ---- res_b = b + [B A] ux - x_{k+1}, one (stage, state) per thread
        for (int it = tid; it < N * NX; it += T)
        {
            const int k = (it / NX), j = it - k * NX;
            const double* v = ux_() + k * NV;
            const double* Gk = G_() + k * (NV * NX) + NV * j;
            double acc = b_()[it] - ux_()[(k + 1) * NV + NU + j];
#pragma unroll
            for (int i = 0; i < NV; i++) acc += Gk[i] * v[i];  // stage 0: the x rows of G are zero
            rb_()[it] = acc;
            const double q = dabs(acc);
            n1 = q > n1 ? q : n1;
        }
        // ---- template of the matrix to factorise (off-diagonal entries), four-entry chunks
        if (T < NE)
        {
            for (int it = tid; it < (N + 1) * NE; it += T) Mx_()[it] = Tp[stage_class(it / NE) * NE + it % NE];
        }
        else
        {
            // thread -> (stage k0 + q * SPR, entry e): the path-stage template value stays in a register
            const int SPR = T / NE, k0 = tid / NE, e = tid - k0 * NE;
            if (k0 < SPR)
            {
                const double tv = Tp[NE + e];
                double* dst = Mx_() + e;
                for (int k = k0; k <= N; k += SPR) dst[k * NE] = (k == 0 || k == N) ? Tp[(k == 0 ? 0 : 2) * NE + e] : tv;
            }
        }
        syncthreads();
        PROF(20)
        // ---- A3: stationarity residual, diagonal and gradient row, one (stage, variable) per thread
        for (int it = tid; it < (N + 1) * NV; it += T)
        {
            const int k = (it / NV), i = it - k * NV;
            const double* v = ux_() + k * NV;
            const double* H = Hk(k);
            double g = rq_()[it], dg = 0.0, gg = 0.0;
#pragma unroll
            for (int j = 0; j < NV; j++) g += H[i + NV * j] * v[j];
            if (k > 0 && i >= NU) g -= pi_()[(k - 1) * NX + i - NU];
            if (k < N)
            {
                const double* Gk = G_() + k * (NV * NX) + i;
                const double* pk = pi_() + k * NX;
                double acc = 0.0;
#pragma unroll
                for (int j = 0; j < NX; j++) acc += Gk[NV * j] * pk[j];
                g += acc;
                const int row = vrow(k, i);
                if (row >= 0)
                {
                    const int r0 = k * s2 + row;
                    g += dlam_()[r0]; dg += dlam_()[r0 + ncq]; gg += dt_()[r0];
                }
                if ((i == HXV || i == HYV) && k >= 1)
                {
                    const double* gk = gxy_() + k * 2 * K;
                    const double* gi = i == HXV ? gk : gk + K;
                    double aYX = 0.0;
                    const double* pl = dlam_() + k * s2 + nbq;   // parked by A2: lam_u - lam_l | Gamma_l + Gamma_u | gamma_l - gamma_u
                    const double* pG = pl + ncq;
                    const double* pg = dt_() + k * s2 + nbq;
#pragma unroll 5
                    for (int c = 0; c < K; c++)
                    {
                        const double Gs = pG[c], gc = gi[c];
                        g += gc * pl[c];
                        dg += (gc * Gs) * gc;
                        gg += pg[c] * gc;
                        aYX += (gc * Gs) * gk[c];   // used by the HY lane only (for it gc = gy[c], gk[c] = gx[c]); no branch in the loop
                    }
                    if (i == HYV) Mx_()[k * NE + MI(HYV > HXV ? HYV : HXV, HYV > HXV ? HXV : HYV)] += aYX;
                }
            }
            const double gi = var_active(k, i) ? g : 0.0;
            rg_()[it] = gi;
            Mx_()[k * NE + MI(i, i)] += dg;
            Mx_()[k * NE + MI(NV, i)] = gi + gg;
            const double q = dabs(gi);
            n0 = q > n0 ? q : n0;
        }
        double vm[4] = {n0, n1, n2, n3}, vs[1] = {musum};
        PROF(21)
        block_reduce<4, 1>(vm, vs);
        PROF(22)
        n4[0] = vm[0]; n4[1] = vm[1]; n4[2] = vm[2]; n4[3] = vm[3];
        mu = nct > 0 ? vs[0] / nct : 0.0;
    }

    // passM: assemble the matrices to factorise once more from the current iterate (H + Gamma terms, gradient row =
    // res_g + gamma terms) -- what the tail of passA does -- after a factorisation consumed them (fp32 attempt of
    // config 4 that failed its accuracy test).  Rare path.
    MDEV void passM()
    {
        const double tau = 1e-16;
        if (T < NE)
        {
            for (int it = tid; it < (N + 1) * NE; it += T) Mx_()[it] = Tp[stage_class(it / NE) * NE + it % NE];
        }
        else
        {
            // thread -> (stage k0 + q * SPR, entry e): the path-stage template value stays in a register
            const int SPR = T / NE, k0 = tid / NE, e = tid - k0 * NE;
            if (k0 < SPR)
            {
                const double tv = Tp[NE + e];
                double* dst = Mx_() + e;
                for (int k = k0; k <= N; k += SPR) dst[k * NE] = (k == 0 || k == N) ? Tp[(k == 0 ? 0 : 2) * NE + e] : tv;
            }
        }
        syncthreads();
        for (int it = tid; it < (N + 1) * NV; it += T)
        {
            const int k = (it / NV), i = it - k * NV;
            double dg = 0.0, gg = 0.0;
            if (k < N)
            {
                auto row_terms = [&](int r0, double& Gs, double& gd) {
                    const int r1 = r0 + ncq;
                    const double l0 = lam_()[r0], l1 = lam_()[r1], t0 = t_()[r0], t1 = t_()[r1], i0 = 1.0 / t0, i1 = 1.0 / t1;
                    double G[2] = {i0 * l0, i1 * l1};
                    double g[2] = {i0 * ((l0 * t0 - tau) - l0 * rd_()[r0]), i1 * ((l1 * t1 - tau) - l1 * rd_()[r1])};
                    const int is = soft_index(r0 - k * s2);
                    if (is >= 0)
                    {
#pragma unroll
                        for (int side = 0; side < 2; side++)
                        {
                            const int rs = k * s2 + 2 * ncq + side * ns + is, si = k * 2 * ns + side * ns + is;
                            const double ls = lam_()[rs], ts = t_()[rs], tis = 1.0 / ts;
                            const double Gsl = tis * ls, gsl = tis * ((ls * ts - tau) - ls * rd_()[rs]);
                            const double zi = zsi_()[si], tmp = (rgs_()[si] + g[side] + gsl) * zi, Gr = G[side];
                            (void) Gsl;
                            G[side] = Gr - Gr * zi * Gr;
                            g[side] = g[side] - Gr * tmp;
                        }
                    }
                    Gs = G[0] + G[1];
                    gd = g[0] - g[1];
                };
                const int row = vrow(k, i);
                if (row >= 0) { double Gs, gd; row_terms(k * s2 + row, Gs, gd); dg += Gs; gg += gd; }
                if ((i == HXV || i == HYV) && k >= 1)
                {
                    const double* gk = gxy_() + k * 2 * K;
                    const double* gi = i == HXV ? gk : gk + K;
                    double aYX = 0.0;
                    for (int c = 0; c < K; c++)
                    {
                        double Gs, gd;
                        row_terms(k * s2 + nbq + c, Gs, gd);
                        dg += (gi[c] * Gs) * gi[c];
                        gg += gd * gi[c];
                        if (i == HYV) aYX += (gk[K + c] * Gs) * gk[c];
                    }
                    if (i == HYV) Mx_()[k * NE + MI(HYV > HXV ? HYV : HXV, HYV > HXV ? HXV : HYV)] += aYX;
                }
            }
            Mx_()[k * NE + MI(i, i)] += dg;
            Mx_()[k * NE + MI(NV, i)] = rg_()[it] + gg;
        }
        syncthreads();
    }

    // ---------------------------------------------------------------- IPM: the serial recursions (warp 0)
    // The recursions are free functions (chain_*_impl below, not inlined: their own register allocation and compact
    // code, whatever the state of the surrounding solver) called by warp 0 only.
    MDEV void chainA()
    {
        if (wid >= CHAIN_WARPS) return;
        if (use_fp32) chainA_impl<M, float>(G_(), Mx_(), rb_(), Pb_(), dinv_(), sW, sP, N);
#if USVMPC_CHAIN_MMA
        else if (wid == 0) chainA_mma<M>(G_(), Mx_(), rb_(), Pb_(), dinv_(), sW, sP, slots + 1, N);
#else
        else chainA_impl<M, double>(G_(), Mx_(), rb_(), Pb_(), dinv_(), sW, sP, N);
#endif
    }
    // dx_0 = 0, dx_{k+1} = Acl_k dx_k + c_k -> x part of `out` (stage stride NV)
    // `out` in the block's global scratch (the refinement step on a long horizon): the recursion runs in place on the
    // affine terms (every c_k is in a register before x_k lands on it) and the block copies the result out.
    MDEV void chainF(double* out, bool out_shared)
    {
        if (out_shared)
        {
            if (wid == 0) chain_vec_mma<M, true>(Acl_(), cc_(), out, slots, N, NV, NU);
            return;
        }
        if (wid == 0) chain_vec_mma<M, true>(Acl_(), cc_(), cc_(), slots, N, NX, 0);
        syncthreads();
        for (int it = tid; it < (N + 1) * NX; it += T)
        {
            const int k = it / NX, i = it - k * NX;
            out[k * NV + NU + i] = cc_()[it];
        }
    }
    // p_N = e_N, p_k = Acl_k' p_{k+1} + e_k -> x part of zv
    MDEV void chainC()
    {
        if (wid == 0) chain_vec_mma<M, false>(Acl_(), ee_(), zv_(), slots, N, NV, NU);
    }

    // ---------------------------------------------------------------- IPM: passes around the chains
    // gains after chainA: K_k = -Luu^-T Lxu', Acl_k = A_k + B_k K_k
    MDEV void gains_pass()
    {
        for (int it = tid; it < (N + 1) * NX; it += T)
        {
            const int k = (it / NX), j = it - k * NX;
            const double* Mk = Mx_() + k * NE;
            double kg[NU];
#pragma unroll
            for (int i = NU - 1; i >= 0; i--)
            {
                double acc = -Mk[MI(NU + j, i)];
#pragma unroll
                for (int m = i + 1; m < NU; m++) acc -= Mk[MI(m, i)] * kg[m];
                kg[i] = acc * dinv_()[k * NU + i];
                Kg_()[k * (NU * NX) + i * NX + j] = kg[i];
            }
        }
        syncthreads();
        for (int it = tid; it < N * NX * NX; it += T)
        {
            const int k = it / (NX * NX), e = it - k * (NX * NX), i = e / NX, j = e - i * NX;
            const double* Gk = G_() + k * (NV * NX) + NV * i;
            double acc = Gk[NU + j];
#pragma unroll
            for (int m = 0; m < NU; m++) acc += Gk[m] * Kg_()[k * (NU * NX) + m * NX + j];
            Acl_()[it] = acc;
        }
    }

    // feed-forward terms: kk_k = -Luu^-T l_u,k and c_k = rhs_b,k + B_k kk_k.  from_factor: l_u = row NV of the factor
    // (factorise+solve sweep); else l_u = Luu^-1 (z0_u + B'(p_{k+1} + Pb_k)) with p in the x part of zv.  One
    // (stage, state) item per thread; the NU-vector kk is recomputed by each of the stage's NX items (cheaper than a
    // barrier).
    MDEV void feedforward_pass(bool from_factor, const double* rbp)
    {
        for (int it = tid; it < N * NX; it += T)
        {
            const int k = (it / NX), mm = it - k * NX;
            const double* Mk = Mx_() + k * NE;
            const double* Gk = G_() + k * (NV * NX);
            double lu[NU], kv[NU];
#pragma unroll
            for (int i = 0; i < NU; i++)
            {
                double acc;
                if (from_factor) acc = Mk[MI(NV, i)];
                else
                {
                    acc = zv_()[k * NV + i];
#pragma unroll
                    for (int m = 0; m < NX; m++) acc += Gk[i + NV * m] * (zv_()[(k + 1) * NV + NU + m] + Pb_()[k * NX + m]);
#pragma unroll
                    for (int m = 0; m < i; m++) acc -= Mk[MI(i, m)] * lu[m];
                    acc *= dinv_()[k * NU + i];
                }
                lu[i] = acc;
            }
#pragma unroll
            for (int i = NU - 1; i >= 0; i--)
            {
                double acc = -lu[i];
#pragma unroll
                for (int m = i + 1; m < NU; m++) acc -= Mk[MI(m, i)] * kv[m];
                kv[i] = acc * dinv_()[k * NU + i];
                if (mm == 0) kk_()[k * NU + i] = kv[i];
            }
            double acc = rbp[it];
#pragma unroll
            for (int i = 0; i < NU; i++) acc += Gk[i + NV * mm] * kv[i];
            cc_()[it] = acc;
        }
        if (tid < NU) kk_()[N * NU + tid] = 0.0;  // no inputs at the terminal stage
        syncthreads();
    }

    // right-hand side of a solve with the existing factorisation (OCP_QP_SOLVE_KKT_STEP, x_ocp_qp_kkt.c:1096-1242):
    // gamma rows (COMPUTE_GAMMA_QP, x_core_qp_ipm_aux.c:89-113) -> z0 = rhs_g + J'(gamma_l - gamma_u) -> e_k.
    //   mode 0: corrector, res_m = lam*t + dt_aff*dlam_aff - sigma_mu -> rmc      (x_ocp_qp_ipm.c:2138-2160)
    //   mode 1: centering only, res_m = lam*t - sigma_mu -> rmc                    (:2175-2200)
    //   mode 2: iterative refinement, right-hand side (rg2, rb2, rd2, rm2)          (:2221-2311)
    // gsc: a dead array with the stride of the row arrays that receives gamma_l - gamma_u.
    MDEV void rhs_pass(int mode, double sigma_mu)
    {
        const double* rgp = mode == 2 ? rg2_() : rg_();
        const double* rdp = mode == 2 ? rd2_() : rd_();
        double* gsc = mode == 2 ? dlam2_() : dlam_();
        for (int it = tid; it < N * ncq; it += T)
        {
            const int k = dq.div(it), j = it - k * ncq;
            const int r0 = k * s2 + j, r1 = r0 + ncq;
            if (!row_active(k, j)) { if (mode != 2) { rmc_()[r0] = 0.0; rmc_()[r1] = 0.0; } gsc[r0] = 0.0; continue; }
            const double l0 = lam_()[r0], l1 = lam_()[r1];
            double m0, m1;
            if (mode == 2) { m0 = rm2_()[r0]; m1 = rm2_()[r1]; }
            else
            {
                m0 = l0 * t_()[r0]; m1 = l1 * t_()[r1];
                if (mode == 0) { m0 += dt_()[r0] * dlam_()[r0]; m1 += dt_()[r1] * dlam_()[r1]; }
                m0 -= sigma_mu; m1 -= sigma_mu;
                rmc_()[r0] = m0; rmc_()[r1] = m1;   // kept for a refinement's residual (rare; the field may live in L2)
                dt_()[r0] = m0; dt_()[r1] = m1;     // the affine step is dead from here on: expand_pass picks the values up
            }
            const double ti0 = 1.0 / t_()[r0], ti1 = 1.0 / t_()[r1];   // 1 / t is recomputed where it is needed (same bits as in passA)
            double g0 = ti0 * (m0 - l0 * rdp[r0]), g1 = ti1 * (m1 - l1 * rdp[r1]);
            const int is = soft_index(j);
            if (is >= 0)
            {
                // COND_SLACKS_SOLVE (x_ocp_qp_kkt.c:295-353): gamma of the slack-bound rows, right-hand side of the slack
                // variables, gamma of the row with the slacks eliminated
                const double* rgsp = mode == 2 ? rgs2_() : rgs_();
                double* dso = mode == 2 ? dsv2_() : dsv_();
#pragma unroll
                for (int side = 0; side < 2; side++)
                {
                    const int rs = k * s2 + 2 * ncq + side * ns + is, si = k * 2 * ns + side * ns + is;
                    const double ls = lam_()[rs];
                    double ms;
                    if (mode == 2) ms = rm2_()[rs];
                    else
                    {
                        ms = ls * t_()[rs];
                        if (mode == 0) ms += dt_()[rs] * dlam_()[rs];
                        ms -= sigma_mu;
                        rmc_()[rs] = ms;
                        dt_()[rs] = ms;
                    }
                    const double gsl = (1.0 / t_()[rs]) * (ms - ls * rdp[rs]);
                    const double gr = side ? g1 : g0, Gr = side ? ti1 * l1 : ti0 * l0;
                    const double rhs = rgsp[si] + gr + gsl;
                    dso[si] = rhs;
                    const double tmp = rhs * zsi_()[si];
                    if (side) g1 = gr - Gr * tmp; else g0 = gr - Gr * tmp;
                }
            }
            gsc[r0] = g0 - g1;
        }
        syncthreads();
        for (int it = tid; it < (N + 1) * NV; it += T)
        {
            const int k = (it / NV), i = it - k * NV;
            double z = rgp[it];
            const int row = vrow(k, i);
            if (row >= 0) z += gsc[k * s2 + row];
            if ((i == HXV || i == HYV) && k >= 1 && k < N)
            {
                const double* gi = gxy_() + k * 2 * K + (i == HXV ? 0 : K);
                const double* gs_ = gsc + k * s2 + nbq;
#pragma unroll 5
                for (int c = 0; c < K; c++) z += gi[c] * gs_[c];
            }
            zv_()[it] = var_active(k, i) ? z : 0.0;
        }
        if (mode == 2)
        {
            // Pb = P_{k+1} rhs_b for the new right-hand side (the factor's Pb belongs to res_b)
            for (int it = tid; it < N * NX; it += T)
            {
                const int k = (it / NX), m = it - k * NX;
                const double* Mn = Mx_() + (k + 1) * NE;
                double acc = 0.0;
#pragma unroll
                for (int n = 0; n < NX; n++) acc += (m >= n ? Mn[MI(NU + m, NU + n)] : Mn[MI(NU + n, NU + m)]) * rb2_()[k * NX + n];
                Pb_()[it] = acc;
            }
        }
        syncthreads();
        // e_k = Acl_k' Pb_k + z0_x + K_k' z0_u   (e_N = z0_x)
        for (int it = tid; it < (N + 1) * NX; it += T)
        {
            const int k = (it / NX), j = it - k * NX;
            double acc = zv_()[k * NV + NU + j];
#pragma unroll
            for (int m = 0; m < NU; m++) acc += Kg_()[k * (NU * NX) + m * NX + j] * zv_()[k * NV + m];
            if (k < N)
            {
#pragma unroll
                for (int m = 0; m < NX; m++) acc += Acl_()[k * (NX * NX) + m * NX + j] * Pb_()[k * NX + m];
            }
            ee_()[it] = acc;
        }
        syncthreads();
    }

    // expand_pass: du = K dx + kk; dt, dlam from the primal step (x_ocp_qp_kkt.c:748-764 + COMPUTE_LAM_T_QP,
    // x_core_qp_ipm_aux.c:117-142), COMPUTE_ALPHA_QP (:146-216), the sums COMPUTE_MU_AFF_QP (:329-357) needs, the
    // inequality / complementarity rows of OCP_QP_RES_COMPUTE_LIN (x_ocp_qp_res.c:560-633) of this very step, and
    // dpi_k = P_{k+1} dx_{k+1} + p_{k+1} (x_ocp_qp_kkt.c:560-575 | 1270-1290).
    //   mode 0: affine step (res_m = lam*t - tau, p = row NV of the factor); 1: rhs rmc, p = x part of zv;
    //   mode 2: refinement (rd2, rm2 -> dux2, dpi2, dlam2, dt2), no step length
    MDEV void expand_pass(int mode, double tau)
    {
        double* vo = mode == 2 ? dux2_() : dux_();
        double* dlo = mode == 2 ? dlam2_() : dlam_();
        double* dto = mode == 2 ? dt2_() : dt_();
        double* dpo = mode == 2 ? dpi2_() : dpi_();
        const double* rdp = mode == 2 ? rd2_() : rd_();
        const double* rmp = mode == 2 ? rm2_() : rmc_();
        for (int it = tid; it < (N + 1) * NU; it += T)
        {
            const int k = it / NU, m = it - k * NU;
            double acc = kk_()[it];
            const double* Kr = Kg_() + k * (NU * NX) + m * NX;
            const double* dx = vo + k * NV + NU;
#pragma unroll
            for (int j = 0; j < NX; j++) acc += Kr[j] * dx[j];
            vo[k * NV + m] = k < N ? acc : 0.0;
        }
        syncthreads();
        double bdn = 1.0, bdd = -1.0, bpn = 1.0, bpd = -1.0, s1 = 0.0, s2s = 0.0, nd = 0.0, nm = 0.0, ng = 0.0;
        // one inequality row: dlam, dt from the row's expanded step `dtr` (J dux, + slack step on a soft row), step-length
        // candidates, mu_aff sums, linear-system residual of the row (`lin` = the part of the row's equation that is not dt)
        auto do_row = [&](int r, double dtr, double lin) -> double {
            const double lam0 = lam_()[r], t0 = t_()[r], e0 = rdp[r];
            double m;
            if (mode == 0) m = lam0 * t0 - tau;
            else if (mode == 2) m = rmp[r];
            else m = dto[r];   // the right-hand side rhs_pass formed, parked in the slot this row's new dt is about to overwrite
            const double dlr = -(1.0 / t0) * (m + (lam0 * dtr) - (lam0 * e0));
            dtr -= e0;
            dlo[r] = dlr; dto[r] = dtr;
            // COMPUTE_ALPHA_QP keeps the ratio closest to zero among rows with a negative step; the running best is
            // kept as (numerator, denominator) and compared by cross-multiplication: one division per thread
            if (dlr < 0.0 && bdn * dlr < lam0 * bdd) { bdn = lam0; bdd = dlr; }
            if (dtr < 0.0 && bpn * dtr < t0 * bpd) { bpn = t0; bpd = dtr; }
            s1 += lam0 * dtr + t0 * dlr;
            s2s += dlr * dtr;
            const double e = e0 + dtr - lin;
            const double mm = m + lam0 * dtr + dlr * t0;
            double q = dabs(e); nd = q > nd ? q : nd;
            q = dabs(mm); nm = q > nm ? q : nm;
            return dlr;
        };
        for (int it = tid; it < N * ncq; it += T)
        {
            const int k = dq.div(it), j = it - k * ncq;
            if (!row_active(k, j)) continue;
            const double* v = vo + k * NV;
            double dv;
            if (j < nbq) dv = v[srvar[j]];
            else dv = k >= 1 ? gxy_()[k * 2 * K + j - nbq] * v[HXV] + gxy_()[k * 2 * K + K + j - nbq] * v[HYV] : 0.0;
            const int r0 = k * s2 + j, r1 = r0 + ncq;
            const int is = soft_index(j);
            if (is < 0)
            {
                do_row(r0, dv, dv);
                do_row(r1, -dv, -dv);
            }
            else
            {
                // EXPAND_SLACKS (x_ocp_qp_kkt.c:357-400): the slack steps from the row's J dux, then the four rows
                double* dso = mode == 2 ? dsv2_() : dsv_();
                const double* rgsp = mode == 2 ? rgs2_() : rgs_();
                const int s0 = k * 2 * ns + is, s1i = s0 + ns, rs0 = k * s2 + 2 * ncq + is, rs1 = rs0 + ns;
                const double ds0 = -zsi_()[s0] * (dso[s0] + dv * ((1.0 / t_()[r0]) * lam_()[r0]));
                const double ds1 = -zsi_()[s1i] * (dso[s1i] + (-dv) * ((1.0 / t_()[r1]) * lam_()[r1]));
                dso[s0] = ds0; dso[s1i] = ds1;
                const double dl0 = do_row(r0, dv + ds0, dv + ds0);
                const double dl1 = do_row(r1, -dv + ds1, -dv + ds1);
                const double dls0 = do_row(rs0, ds0, ds0);
                const double dls1 = do_row(rs1, ds1, ds1);
                // stationarity rows of the slacks in the linear system (x_ocp_qp_res.c:540-556)
                double q = dabs(zquad(is, 0) * ds0 + rgsp[s0] - dls0 - dl0); ng = q > ng ? q : ng;
                q = dabs(zquad(is, 1) * ds1 + rgsp[s1i] - dls1 - dl1); ng = q > ng ? q : ng;
            }
        }
        // dpi_k = P_{k+1} dx_{k+1} + p_{k+1}
        for (int it = tid; it < N * NX; it += T)
        {
            const int k = (it / NX), i = it - k * NX;
            const double* Mn = Mx_() + (k + 1) * NE;
            const double* xn = vo + (k + 1) * NV + NU;
            double acc = mode == 0 ? Mn[MI(NV, NU + i)] : zv_()[(k + 1) * NV + NU + i];
#pragma unroll
            for (int n = 0; n < NX; n++) acc += (i >= n ? Mn[MI(NU + i, NU + n)] : Mn[MI(NU + n, NU + i)]) * xn[n];
            dpo[it] = acc;
        }
        if (mode != 2)
        {
            double vm[5] = {bpn / bpd, bdn / bdd, nd, nm, ng}, vs[2] = {s1, s2s};
            block_reduce<5, 2, 2>(vm, vs);
            alpha = -(vm[0] > vm[1] ? vm[0] : vm[1]);
            lin_d = vm[2]; lin_m = vm[3]; lin_g = vm[4];
            S1 = vs[0]; S2 = vs[1];
        }
        else syncthreads();
    }

    // OCP_QP_RES_COMPUTE_LIN (HP/ocp_qp/x_ocp_qp_res.c:468-633): residual of the Newton system with right-hand side
    // (rg, rb, rd, rmc) at the step (dux, dpi, dlam, dt); norms into out4.  WRITE: also store it (rg2, rb2, rd2, rm2) as
    // the right-hand side of an iterative-refinement solve and compute all four norms from scratch; without WRITE the
    // step is expand_pass's and so are the norms of the inequality / complementarity rows (lin_d, lin_m).
    MDEV void res_pass(const bool WRITE, double* out4)
    {
        double n0 = 0, n1 = 0, n2 = 0, n3 = 0;
        for (int it = tid; it < (N + 1) * NV; it += T)
        {
            const int k = (it / NV), i = it - k * NV;
            const double* v = dux_() + k * NV;
            const double* H = Hk(k);
            double g = rg_()[it];
#pragma unroll
            for (int j = 0; j < NV; j++) g += H[i + NV * j] * v[j];
            if (k > 0 && i >= NU) g -= dpi_()[(k - 1) * NX + i - NU];
            if (k < N)
            {
                const double* Gk = G_() + k * (NV * NX) + i;
                const double* pk = dpi_() + k * NX;
                double acc = 0.0;
#pragma unroll
                for (int j = 0; j < NX; j++) acc += Gk[NV * j] * pk[j];
                g += acc;
                const int row = vrow(k, i);
                if (row >= 0) g += dlam_()[k * s2 + ncq + row] - dlam_()[k * s2 + row];
                if ((i == HXV || i == HYV) && k >= 1)
                {
                    const double* gi = gxy_() + k * 2 * K + (i == HXV ? 0 : K);
                    const double* dl_ = dlam_() + k * s2 + nbq;
                    const double* du_ = dl_ + ncq;
#pragma unroll 5
                    for (int c = 0; c < K; c++) g += gi[c] * (du_[c] - dl_[c]);
                }
            }
            const double gi = var_active(k, i) ? g : 0.0;
            if (WRITE) rg2_()[it] = gi;
            const double q = dabs(gi);
            n0 = q > n0 ? q : n0;
        }
        PROF(14)
        for (int it = tid; it < N * NX; it += T)
        {
            const int k = (it / NX), j = it - k * NX;
            const double* v = dux_() + k * NV;
            const double* Gk = G_() + k * (NV * NX) + NV * j;
            double acc = rb_()[it] - dux_()[(k + 1) * NV + NU + j];
#pragma unroll
            for (int i = 0; i < NV; i++) acc += Gk[i] * v[i];
            if (WRITE) rb2_()[it] = acc;
            const double q = dabs(acc);
            n1 = q > n1 ? q : n1;
        }
        PROF(15)
        if (WRITE)
        {
            for (int it = tid; it < N * ncq; it += T)
            {
                const int k = dq.div(it), j = it - k * ncq;
                const int r0 = k * s2 + j, r1 = r0 + ncq;
                if (!row_active(k, j)) { rd2_()[r0] = 0.0; rd2_()[r1] = 0.0; rm2_()[r0] = 0.0; rm2_()[r1] = 0.0; continue; }
                const double* v = dux_() + k * NV;
                double vv;
                if (j < nbq) vv = v[srvar[j]];
                else vv = k >= 1 ? gxy_()[k * 2 * K + j - nbq] * v[HXV] + gxy_()[k * 2 * K + K + j - nbq] * v[HYV] : 0.0;
                double e0 = rd_()[r0] + dt_()[r0] - vv, e1 = rd_()[r1] + dt_()[r1] + vv;
                const int is = soft_index(j);
                if (is >= 0)
                {
                    // soft row: the slack step enters the row; rows of the slack bounds and of the slacks' stationarity
#pragma unroll
                    for (int side = 0; side < 2; side++)
                    {
                        const int rs = k * s2 + 2 * ncq + side * ns + is, si = k * 2 * ns + side * ns + is;
                        const double ds = dsv_()[si];
                        if (side) e1 -= ds; else e0 -= ds;
                        const double es = rd_()[rs] + dt_()[rs] - ds;
                        rd2_()[rs] = es;
                        double q = dabs(es); n2 = q > n2 ? q : n2;
                        const double ms = rmc_()[rs] + lam_()[rs] * dt_()[rs] + dlam_()[rs] * t_()[rs];
                        rm2_()[rs] = ms;
                        q = dabs(ms); n3 = q > n3 ? q : n3;
                        const double gs = zquad(is, side) * ds + rgs_()[si] - dlam_()[rs] - dlam_()[side ? r1 : r0];
                        rgs2_()[si] = gs;
                        q = dabs(gs); n0 = q > n0 ? q : n0;
                    }
                }
                rd2_()[r0] = e0; rd2_()[r1] = e1;
                double q = dabs(e0); n2 = q > n2 ? q : n2; q = dabs(e1); n2 = q > n2 ? q : n2;
                const double m0 = rmc_()[r0] + lam_()[r0] * dt_()[r0] + dlam_()[r0] * t_()[r0];
                const double m1 = rmc_()[r1] + lam_()[r1] * dt_()[r1] + dlam_()[r1] * t_()[r1];
                rm2_()[r0] = m0; rm2_()[r1] = m1;
                q = dabs(m0); n3 = q > n3 ? q : n3; q = dabs(m1); n3 = q > n3 ? q : n3;
            }
        }
        double vm[4] = {n0, n1, n2, n3};
        PROF(16)
        block_reduce<4, 0>(vm, nullptr);
        PROF(17)
        out4[0] = (WRITE || !SOFT) ? vm[0] : (vm[0] > lin_g ? vm[0] : lin_g);
        out4[1] = vm[1]; out4[2] = WRITE ? vm[2] : lin_d; out4[3] = WRITE ? vm[3] : lin_m;
    }

    // COMPUTE_ALPHA_QP on the current step (after iterative refinement changed it)
    MDEV void alpha_pass()
    {
        double a_prim = -1.0, a_dual = -1.0;
        for (int it = tid; it < N * s2; it += T)
        {
            int k;
            if (!elem_active(it, k)) continue;
            if (a_dual * dlam_()[it] > lam_()[it]) a_dual = lam_()[it] / dlam_()[it];
            if (a_prim * dt_()[it] > t_()[it]) a_prim = t_()[it] / dt_()[it];
        }
        double vm[2] = {a_prim, a_dual};
        block_reduce<2, 0, 2>(vm, nullptr);
        alpha = -(vm[0] > vm[1] ? vm[0] : vm[1]);
    }

    // step += refinement step
    MDEV void add_refinement()
    {
        for (int e = tid; e < (N + 1) * NV; e += T) dux_()[e] += dux2_()[e];
        for (int e = tid; e < N * NX; e += T) dpi_()[e] += dpi2_()[e];
        if (SOFT) for (int e = tid; e < N * 2 * ns; e += T) dsv_()[e] += dsv2_()[e];
        for (int e = tid; e < N * s2; e += T)
        {
            int k;
            if (elem_active(e, k)) { dlam_()[e] += dlam2_()[e]; dt_()[e] += dt2_()[e]; }
        }
        syncthreads();
    }

    MDEV bool itref_ok(const double* n) const
    {
        return (n[0] < tol_stat || n[0] < 1e-3 * res_max[0]) && (n[1] < tol_eq || n[1] < 1e-3 * res_max[1]) &&
               (n[2] < tol_ineq || n[2] < 1e-3 * res_max[2]) && (n[3] < tol_comp || n[3] < 1e-3 * res_max[3]);
    }

    // OCP_QP_IPM_SOLVE + OCP_QP_IPM_DELTA_STEP: HP/ocp_qp/x_ocp_qp_ipm.c:2354-2683, 1888-2350 (pred_corr,
    // cond_pred_corr, itref_corr_max = 2); returns HPIPM status 0 ok / 1 max iter / 2 min step / 3 NaN.
    // One IPM iteration is a sequence of linear solves with the same matrix -- AFF (factorise + affine step), COR
    // (corrector), CEN (centering only, conditional), REF (iterative refinement, rare) -- written as ONE loop over
    // solve rounds so that every pass and chain has a single call site (compact code: the whole iteration stays in
    // the instruction cache).
    MDEV int ipm_solve(int* iters)
    {
        enum { AFF, COR, CEN, REF };
        const double tau_min = 1e-16, alpha_min = 1e-8;
        ipm_init();
        PROF(10)
        alpha = 1.0;
        double a = 0.0;  // the first passA applies no step
        int it = 0;
        use_fp32 = P.chain_fp32 != 0;
        for (;;)
        {
            passA(a, nct > 0 ? tau_min : 0.0, res_max);
            PROF(0)
            if (nct == 0 && it == 1) { it = 0; break; }  // no inequality rows: one direct solve (x_ocp_qp_ipm.c:2444-2478)
            if (!(it < iter_max && alpha > alpha_min &&
                  (res_max[0] > tol_stat || res_max[1] > tol_eq || res_max[2] > tol_ineq ||
                   dabs(res_max[3] - tau_min) > tol_comp)))
                break;
            int kind = AFF, nref = 0;
            bool refined = false, refactor = false;
            double sigma_mu = 0.0;
            for (;;)
            {
                if (kind == AFF)
                {
                    if (refactor) passM();  // the fp32 attempt consumed the matrices: assemble them again
                    chainA();
                    syncthreads();
                    if (use_fp32) fp32_count++;
                    PROF(1)
                    gains_pass();
                    syncthreads();
                }
                else
                {
                    rhs_pass(kind == COR ? 0 : (kind == CEN ? 1 : 2), sigma_mu);
                    PROF(6)
                    chainC();
                    syncthreads();
                    solve_calls++;
                    PROF(7)
                }
                feedforward_pass(kind == AFF, kind == REF ? rb2_() : rb_());
                PROF(2)
                chainF(kind == REF ? dux2_() : dux_(), kind != REF || P.plan.f[F_DUX2].space == 0);
                syncthreads();
                PROF(3)
                expand_pass(kind == AFF ? 0 : (kind == REF ? 2 : 1), tau_min);
                PROF(4)
                if (kind == REF) { add_refinement(); refined = true; itref_count++; }
                if (kind == COR)
                {
                    const double mu_aff0 = mu_aff;
                    mu_aff = (mu * nct + alpha * S1 + alpha * alpha * S2) / nct;  // COMPUTE_MU_AFF_QP
                    if (mu_aff > 2.0 * mu_aff0) { kind = CEN; continue; }
                }
                // residual of the linear system at the step (OCP_QP_RES_COMPUTE_LIN): accuracy test of the factorisation
                // after AFF, iterative-refinement test after COR / CEN / REF
                double nlin[4];
                bool wr = kind == REF, more = false;
                for (;;)
                {
                    res_pass(wr, nlin);
                    if (kind == AFF || nct == 0) break;
                    if (itref_ok(nlin) || nref >= 2) break;
                    if (wr) { more = true; break; }
                    wr = true;  // once more, storing the residual as the right-hand side of the refinement solve
                }
                PROF(9)
                if (nct == 0) break;
                if (kind == AFF && use_fp32 && !itref_ok(nlin))
                {
                    // fp32 factorisation (config 4): the fp64 residual of the affine solve is the accuracy test; when it fails,
                    // this and every later factorisation of the QP run in fp64 (Gamma only grows towards the solution)
                    use_fp32 = false; refactor = true;
                    continue;
                }
                if (kind == AFF)
                {
                    // lq_fact = 1 (x_ocp_qp_ipm.c:1941-1974): HPIPM re-factorises with its LQ-based routine when the residual
                    // of the Cholesky-based solve exceeds 1e-5; counted, see DESIGN.md
                    if (!(nlin[0] <= 1e-5 && nlin[1] <= 1e-5 && nlin[2] <= 1e-5 && nlin[3] <= 1e-5)) lq_count++;
                    mu_aff = (mu * nct + alpha * S1 + alpha * alpha * S2) / nct;
                    const double tmp = mu_aff / mu;
                    sigma = tmp * tmp * tmp;
                    sigma_mu = sigma * mu;
                    sigma_mu = sigma_mu > tau_min ? sigma_mu : tau_min;
                    kind = COR;
                    continue;
                }
                if (more) { kind = REF; nref++; continue; }
                break;
            }
            if (refined) alpha_pass();
            PROF(11)
            a = alpha;
            if (a < 1.0) a = a * ((1.0 - a) * 0.99 + a * 0.9999999);
            it++;
        }
        *iters = it;
        if (it == iter_max) return 1;
        if (alpha <= alpha_min) return 2;
        if (disnan(mu)) return 3;
        return 0;
    }

    // ---------------------------------------------------------------- after the QP
    // d_ocp_qp_restore_eq_dof (HP/ocp_qp/x_ocp_qp_red.c:723-871) + ocp_nlp_update_variables_sqp, full step
    // (AC/acados/ocp_nlp/ocp_nlp_common.c:2401-2448): ux += step; pi, lam, t <- QP values.
    MDEV void update_nlp()
    {
        for (int k = tid; k <= N; k += T)
        {
            double* z = Z(P.lay.zux, k); double* zl = Z(P.lay.zlam, k); double* zt = Z(P.lay.zt, k);
            const double* v = ux_() + k * NV; const double* l = lam_() + k * s2; const double* tt = t_() + k * s2;
            if (k < N)
            {
                double* zp = Z(P.lay.zpi, k);
                for (int i = 0; i < NX; i++) zp[i] = pi_()[k * NX + i];
                for (int j = 0; j < nbu; j++) { zl[j] = l[j]; zl[ncz + j] = l[ncq + j]; zt[j] = tt[j]; zt[ncz + j] = tt[ncq + j]; }
                for (int c = 0; c < K; c++)
                {
                    const int a = nbu + NX + c, bq = nbq + c;
                    zl[a] = l[bq]; zl[ncz + a] = l[ncq + bq]; zt[a] = tt[bq]; zt[ncz + a] = tt[ncq + bq];
                }
                if (SOFT)
                {
                    // the QP's slack values are a step like [u; x]; the multipliers / slacks of the slack bounds are the QP's
                    double* zs = Z(P.lay.zsv, k);
                    for (int j = 0; j < 2 * ns; j++)
                    {
                        zs[j] += sv_()[k * 2 * ns + j];
                        zl[2 * ncz + j] = l[2 * ncq + j]; zt[2 * ncz + j] = tt[2 * ncq + j];
                    }
                }
            }
            if (k == 0)
            {
                // recover the eliminated x0 step and the multipliers of its bounds from stationarity
                const double* zf = Z(P.lay.zfun, 0);
                double s[NV], tmp[NV];
                for (int i = 0; i < NU; i++) s[i] = v[i];
                for (int i = 0; i < NX; i++) s[NU + i] = zf[nbu + i];  // dx0 = x0 - x_0
                for (int i = 0; i < NX; i++)
                {
                    const int iv = NU + i;
                    double acc = rq_()[iv];
                    double hs = 0.0;
                    for (int j = 0; j < NV; j++) hs += Hs[iv + NV * j] * s[j];
                    acc += hs;
                    if (N > 0) for (int j = 0; j < NX; j++) acc += sA0[iv + NV * j] * pi_()[j];
                    if (iv == HXV || iv == HYV)
                        for (int c = 0; c < K; c++)
                        {
                            const double dl = l[ncq + nbq + c] - l[nbq + c];
                            acc += (iv == HXV ? gxy_()[c] : gxy_()[K + c]) * dl;
                        }
                    tmp[iv] = acc;
                }
                for (int j = 0; j < NX; j++)
                {
                    const int r = nbu + j;
                    const double vv = tmp[NU + j];
                    zl[r] = 1e-16; zl[ncz + r] = 1e-16; zt[r] = 1e-16; zt[ncz + r] = 1e-16;
                    if (vv >= 0) zl[r] = vv; else zl[ncz + r] = -vv;
                }
                for (int i = 0; i < NV; i++) z[i] += s[i];
            }
            else
            {
                for (int j = 0; j < nbx; j++)
                {
                    const int a = nbu + j;
                    if (k < N) { zl[a] = l[a]; zl[ncz + a] = l[ncq + a]; zt[a] = tt[a]; zt[ncz + a] = tt[ncq + a]; }
                }
                for (int i = 0; i < NV; i++) if (var_active(k, i)) z[i] += v[i];
            }
        }
        syncthreads();
    }

    // residuals ocp_nlp_eval_residuals reports after an SQP_RTI step: stale linearisation, new lam / t
    // (AC/interfaces/acados_c/ocp_nlp_interface.c:909-916)
    MDEV void rti_residuals(double* res4)
    {
        double r2 = 0, r3 = 0;
        for (int k = tid; k < N; k += T)
        {
            const double* zl = Z(P.lay.zlam, k); const double* zt = Z(P.lay.zt, k); const double* zf = Z(P.lay.zfun, k);
            for (int j = 0; j < 2 * ncz + 2 * ns; j++)
            {
                const int jj = j % ncz;
                const bool act = j >= 2 * ncz || jj < nbu || jj >= nbu + NX || (k == 0 ? true : jj - nbu < nbx);
                if (!act) continue;
                const double a = dabs(zf[j] + zt[j]), c = dabs(zl[j] * zt[j]);
                r2 = a > r2 ? a : r2; r3 = c > r3 ? c : r3;
            }
        }
        double vm[2] = {r2, r3};
        block_reduce<2, 0>(vm, nullptr);
        res4[2] = vm[0]; res4[3] = vm[1];
    }

    // ---------------------------------------------------------------- QP-only entry
    // The seam qp_solver_config.evaluate(qp_in, qp_out) of the reference (AC/acados/ocp_qp/ocp_qp_common.h:62-76): solve ONE
    // given OCP QP per instance (after x0 elimination, the form HPIPM sees) with the IPM and return its solution.  The
    // Hessian is the solver's Gauss-Newton Hessian (constant for LINEAR_LS); everything else comes from P.qp.
    MDEV void run_qp(int inst)
    {
        const QpIo& q = P.qp;
        const double* gG = q.G + (long) inst * N * (NV * NX);
        const double* gb = q.b + (long) inst * N * NX;
        const double* grq = q.rq + (long) inst * (N + 1) * NV;
        const double* gg = q.gxy + (long) inst * N * 2 * K;
        const double* gd = q.d + (long) inst * N * s2;
        for (int e = tid; e < N * NV * NX; e += T) G_()[e] = gG[e];
        for (int e = tid; e < N * NX; e += T) b_()[e] = gb[e];
        for (int e = tid; e < (N + 1) * NV; e += T) rq_()[e] = grq[e];
        for (int e = tid; e < N * 2 * K; e += T) gxy_()[e] = gg[e];
        for (int e = tid; e < N * s2; e += T) d_()[e] = gd[e];
        for (int e = tid; e < s2; e += T) d_()[N * s2 + e] = 0.0;
        syncthreads();
        solve_calls = 0; lq_count = 0; itref_count = 0; fp32_count = 0;
        int it = 0;
        const int qp_status = ipm_solve(&it);
        double* oux = q.ux + (long) inst * (N + 1) * NV;
        double* opi = q.pi + (long) inst * N * NX;
        double* ol = q.lam + (long) inst * N * s2;
        double* ot = q.t + (long) inst * N * s2;
        for (int e = tid; e < (N + 1) * NV; e += T) oux[e] = ux_()[e];
        for (int e = tid; e < N * NX; e += T) opi[e] = pi_()[e];
        for (int e = tid; e < N * s2; e += T) { ol[e] = lam_()[e]; ot[e] = t_()[e]; }
        if (tid == 0)
        {
            double* st = P.stats + (long) inst * NSTAT;
            for (int i = 0; i < NSTAT; i++) st[i] = 0.0;
            st[0] = qp_status; st[2] = it; st[3] = res_max[0]; st[4] = res_max[1]; st[5] = res_max[2]; st[6] = res_max[3];
            st[7] = lq_count; st[8] = solve_calls; st[9] = qp_status; st[10] = it; st[11] = itref_count; st[15] = fp32_count;
        }
        syncthreads();
    }

    // ---------------------------------------------------------------- the solve
    MDEV void run(int inst)
    {
        if (P.qp.G) { run_qp(inst); return; }
        w = P.ws + (long) inst * P.ws_stride;
        const double* x0 = P.x0 + (long) inst * NX;
        const double* pg = P.p + (long) inst * (P.p_per_stage ? (N + 1) : 1) * 2 * K;
        const double* lhg = P.lh + (long) inst * (P.lh_per_stage ? N : 1) * K;
        const double* yrg = P.yref + (long) inst * (P.yref_per_stage ? N : 1) * NY;
        const double* yre = P.yref_e + (long) inst * NX;
        if (P.cold_start) cold_start(x0);
        solve_calls = 0; lq_count = 0; itref_count = 0; fp32_count = 0;
#ifdef USVMPC_PROFILE
        for (int i = 0; i < 24; i++) prof[i] = 0;
        tprev = clock_now();
#endif
        const long long t_start = clock_now();
        long long t_lin = 0, t_qp = 0;
        int status = 2, sqp_iter = 0, qp_total = 0, qp_status = 0, qp_iter = 0;
        double res[4] = {0, 0, 0, 0};
        const int max_iter = P.nlp_type == 0 ? P.max_iter : 1;
        if (P.nlp_type == 1 && P.rti_phase == 1)
        {
            // preparation phase only: no QP, the iterate and the statistics of the last feedback step stay
            integrate_all();
            prep_store(inst);
            return;
        }
        for (sqp_iter = 0; sqp_iter < max_iter; sqp_iter++)
        {
            long long t0 = clock_now();
            linearize(inst, x0, pg, lhg, yrg, yre, res);
            t_lin += clock_now() - t0;
            PROF(12)
            if (P.nlp_type == 0 && res[0] < P.tol[0] && res[1] < P.tol[1] && res[2] < P.tol[2] && res[3] < P.tol[3])
            {
                status = 0;  // ACADOS_SUCCESS, ocp_nlp_sqp.c:641-672
                break;
            }
            t0 = clock_now();
            qp_status = ipm_solve(&qp_iter);
            t_qp += clock_now() - t0;
            qp_total += qp_iter;
            if (qp_status != 0 && qp_status != 1)
            {
                status = 4;  // ACADOS_QP_FAILURE, ocp_nlp_sqp.c:736-773
                break;
            }
            update_nlp();
            PROF(13)
            if (P.nlp_type == 1)
            {
                status = 0;  // ocp_nlp_sqp_rti.c:810-817
                rti_residuals(res);
                sqp_iter = 1;
                break;
            }
        }
        if (tid == 0)
        {
            double* st = P.stats + (long) inst * NSTAT;
            st[0] = status; st[1] = sqp_iter; st[2] = qp_total;
            st[3] = res[0]; st[4] = res[1]; st[5] = res[2]; st[6] = res[3];
            st[7] = lq_count; st[8] = solve_calls; st[9] = qp_status; st[10] = qp_iter; st[11] = itref_count;
            st[12] = (double) (clock_now() - t_start); st[13] = (double) t_lin; st[14] = (double) t_qp; st[15] = fp32_count;
#ifdef USVMPC_PROFILE
            if (qp_total >= USVMPC_PROFILE)
                printf("PROF inst %d sqp %d qp %d | passA %lld chainA %lld gains+ff %lld chainF %lld expand %lld - %lld rhs %lld "
                       "chainC %lld - %lld res %lld init %lld alpha %lld lin %lld upd %lld\n", inst, sqp_iter, qp_total,
                       prof[0], prof[1], prof[2], prof[3], prof[4], prof[5], prof[6], prof[7], prof[8], prof[9], prof[10], prof[11],
                       prof[12], prof[13]);
            if (qp_total >= USVMPC_PROFILE)
                printf("PROF2 res: var %lld state %lld write %lld reduce %lld | passA: A1 %lld A2 %lld resb+tmpl %lld A3 %lld reduce %lld\n",
                       prof[14], prof[15], prof[16], prof[17], prof[18], prof[19], prof[20], prof[21], prof[22]);
#endif
        }
        if (P.packed)
        {
            // epilogue: the instance's result row of the send buffer of the all-gather:
            // x [N+1][NX] | u [N][NU] | status, sqp_iter, qp_iter, res_stat, res_eq, res_ineq, res_comp
            double* row = P.packed + (long) inst * P.packed_width;
            const int nxs = (N + 1) * NX, nus = N * NU;
            for (int e = tid; e < nxs; e += T) { const int k = e / NX, i = e - k * NX; row[e] = Z(P.lay.zux, k)[NU + i]; }
            for (int e = tid; e < nus; e += T) { const int k = e / NU, i = e - k * NU; row[nxs + e] = Z(P.lay.zux, k)[i]; }
            if (tid == 0)
            {
                double* r7 = row + nxs + nus;
                r7[0] = status; r7[1] = sqp_iter; r7[2] = qp_total; r7[3] = res[0]; r7[4] = res[1]; r7[5] = res[2]; r7[6] = res[3];
            }
        }
        syncthreads();
    }
};

// The persistent block: pull instances until the queue is drained.
template <class M, bool SOFT>
MDEV void cta_main(const Params& P, double* smem, int block_id)
{
    CtaSolver<M, SOFT> s(P, smem, P.scratch ? P.scratch + (long) block_id * P.plan.scratch_doubles : nullptr);
    s.load_constants();
    int* slot = (int*) (smem + P.plan.red_off);  // first reduction buffer doubles as the broadcast slot between solves
    for (;;)
    {
        if (s.tid == 0)
        {
            const int ticket = atomic_fetch_add(P.queue, 1);
            slot[0] = ticket < P.B ? (P.order ? P.order[ticket] : ticket) : -1;
        }
        syncthreads();
        const int inst = slot[0];
        syncthreads();
        if (inst < 0) break;
        s.run(inst);
    }
}

}  // namespace usvmpc
