// cta_layout.h -- kernel parameter block, the per-instance HBM block (NLP iterate) and the placement plan of the
// CTA-resident working set of one NMPC instance.
// Shared by the CUDA C-ABI (usvmpc_api.cu), the kernel (cta_kernel.cuh) and the CPU emulation used by the tests.
//
// One thread block solves one instance at a time.  Everything the interior-point iterations touch lives in the
// block's shared memory as dense [stage][dim] arrays; fields that do not fit the shared-memory budget (long
// horizons, many obstacle rows) overflow into a per-BLOCK scratch area in global memory, which is small
// (blocks x a few hundred KB) and therefore stays in L2.  Only the NLP iterate (x, u, pi, lam, t: the reference's
// nlp_out, which persists between solves = warm start) and the inputs / outputs live in HBM per instance.
//
// Row layouts of the inequality vectors (one "side" = lower or upper; the upper side follows the lower side at
// +ncq / +ncz):
//   IPM / QP level  (ncq = nbu + nbx + K): [ u boxes | x boxes (idxbx) | h rows ]
//   NLP level       (ncz = nbu + NX  + K): [ u boxes | x boxes: NX slots (stage 0 holds the x0 embedding,
//                                            stages 1..N-1 use the first nbx) | h rows ]
// Soft rows (the first ns rows of h, each with a lower and an upper slack variable): the lower bounds of the slacks
// follow both sides as rows [ ls (ns) | us (ns) ], like HPIPM's [lb lg | ub ug | ls | us].
#pragma once

namespace usvmpc {

constexpr int KMAX = 32;     // max obstacle rows per stage
constexpr int NMAX = 255;    // max horizon
constexpr int NBXMAX = 8;
constexpr int NBUMAX = 4;
constexpr int NSTAT = 16;    // per-instance statistics record (doubles)
// stats: 0 status, 1 sqp_iter, 2 qp_iter (total), 3..6 res_stat/eq/ineq/comp, 7 IPM iterations whose factorisation
//        failed HPIPM's accuracy test (lq_fact), 8 solve-only Riccati sweeps, 9 last QP status, 10 last QP
//        iterations, 11 iterative-refinement solves, 12 time_tot, 13 time_lin, 14 time_qp (SM clock cycles), 15 factorisations
//        that ran in fp32 (riccati_precision = 32)

struct Field { int off, stride, es; };  // offset of (stage 0, element 0), stage stride, element stride (doubles)

// per-instance block in HBM
struct Layout {
    Field zux, zpi, zlam, zt, zfun, zsv;   // zsv: slack values [sl (ns) | su (ns)] of the NLP iterate
    long total;
};

// fields of the block-resident working set; the first group is touched by the serial Riccati recursions and is
// always in shared memory, the second group in order of decreasing priority for the shared-memory budget
enum FieldId {
    F_G, F_M, F_ACL, F_KG, F_CC, F_EE, F_PB, F_RB, F_ZV, F_DUX, F_KK, F_DINV,
    F_UX, F_PI, F_LAM, F_T, F_DLAM, F_DT, F_RD, F_GXY, F_DPI, F_RG, F_D, F_RQ, F_B, F_RMC,
    F_SV, F_DSV, F_RGS, F_ZSI, F_RQS,   // slack variables: values, step (/ condensation right-hand side), res_g, 1/(Z+Gamma..), gradient
    F_RG2, F_RB2, F_RD2, F_RM2, F_DUX2, F_DPI2, F_DLAM2, F_DT2, F_DSV2, F_RGS2,
    F_COUNT
};
constexpr int F_FIRST_FLEX = F_UX;

struct SField { int off, stride, space; };  // offset (doubles) in shared memory (space 0) / block scratch (space 1)

struct Plan {
    SField f[F_COUNT];
    int const_off;       // constants (Hessians, templates): shared memory
    int red_off;         // reduction scratch
    int misc_off;        // small per-block scratch (unmasked A0, integer tables)
    int smem_doubles;    // dynamic shared memory of one block
    int scratch_doubles; // global scratch of one block (0 if everything fits)
};

// device arrays of the QP-only entry (usvmpc_qp_solve); G == nullptr: normal NMPC solve
struct QpIo {
    const double *G, *b, *rq, *gxy, *d;   // [B][N][NV*NX], [B][N][NX], [B][N+1][NV], [B][N][2K], [B][N][2 ncq]
    double *ux, *pi, *lam, *t;            // [B][N+1][NV], [B][N][NX], [B][N][2 ncq], [B][N][2 ncq]
};

struct Params {
    int B, N, K, num_steps, num_stages, nlp_type, max_iter, qp_iter_max, nbx, nbu;
    int idxbx[NBXMAX];
    int p_per_stage, lh_per_stage, yref_per_stage, cold_start;
    int ncq, ncz;
    int ns;                 // soft rows: the first ns rows of h (0: all constraints hard)
    int chain_fp32;         // 1: Riccati factorisation in fp32 (residuals, solves, refinement fp64): BASELINE.json config 4
    int rti_phase;          // 0: prepare + feedback, 1: prepare only, 2: feedback only (ocp_nlp_sqp_rti.c:459-488)
    double dt, tol[4];
    const double* lbu;      // [N][nbu]   (shared by the batch, per stage like the reference's nlp_in)
    const double* ubu;      // [N][nbu]
    const double* lbx;      // [N][nbx]   (row 0 unused: stage 0 holds the x0 embedding)
    const double* ubx;      // [N][nbx]
    const double* uh;       // [N][K]
    const double* lsh;      // [N][ns] lower bounds of the lower / upper slacks (ocp_nlp_constraints_bgh.c:760-790)
    const double* ush;      // [N][ns]
    const double* zs;       // [4][ns]: zl, zu, Zl, Zu -- linear and quadratic slack penalties (ocp_nlp_cost_ls.c:826-841)
    const double* cst;      // [W (NY*NY col-major) | W_e (NX*NX col-major)]
    const double* x0;       // [B][NX]
    const double* p;        // [B][N+1][2K] or [B][2K]
    const double* lh;       // [B][N][K]   or [B][K]
    const double* yref;     // [B][N][NY]  or [B][NY]
    const double* yref_e;   // [B][NX]
    double* ws;             // [B][ws_stride]
    long ws_stride;
    double* stats;          // [B][NSTAT]
    double* scratch;        // [grid][plan.scratch_doubles]
    double* prep;           // [B][N][NV*NX + NX]: linearisation parked between the RTI preparation and feedback phases (or null)
    double* packed;         // optional [B][packed_width]: result rows written by the solve's epilogue (all-gather send buffer)
    int packed_width;
    int* queue;             // work queue: [0] next ticket
    const int* order;       // ticket -> instance (longest-first order from the previous solve's iteration counts) or null
    QpIo qp;
    Layout lay;
    Plan plan;
};

inline int round_up(int a, int m) { return (a + m - 1) / m * m; }
// scratch of the Riccati factorisation: [P G' | copy of the matrix] (chainA_impl) or W with column stride 12 (chainA_mma),
// and P_{k+1} as a full matrix (row stride NX or 12)
#if defined(__CUDACC__)
#define USVMPC_HD __host__ __device__
#else
#define USVMPC_HD
#endif
USVMPC_HD constexpr int chain_w_doubles(int nx, int nu)
{
    const int nv = nx + nu, ne = nv * (nv + 1) / 2 + nv;
    const int a = nx * (nv + 1) + ne + (ne & 1), b = 12 * (nv + 1) + 32 * ((nv + 7) / 8) + 2;   // W + exchange array + zero slot
    return a > b ? a : b;
}
USVMPC_HD constexpr int chain_p_doubles(int nx) { return nx * nx + 2 > 12 * nx ? nx * nx + 2 : 12 * nx; }

inline Layout make_layout(int nx, int nu, int N, int K, int ns = 0)
{
    Layout L;
    const int N1 = N + 1, nv = nx + nu;
    const int ncz = nu + nx + K + ns;  // room for nbu <= nu input boxes; + the slack-bound rows
    long o = 0;
    auto put = [&](Field& f, int stride) { f.off = (int) o; f.stride = stride; f.es = 1; o += (long) stride * N1; };
    put(L.zux, round_up(nv, 2)); put(L.zpi, round_up(nx, 2));
    put(L.zlam, round_up(2 * ncz, 2)); put(L.zt, round_up(2 * ncz, 2)); put(L.zfun, round_up(2 * ncz, 2));
    put(L.zsv, round_up(2 * ns, 2) > 0 ? round_up(2 * ns, 2) : 2);
    L.total = (o + 15) / 16 * 16;  // 128-byte multiple
    return L;
}

// stage strides of the working-set fields
inline void field_dims(int nx, int nu, int K, int nbx, int nbu, int ns, int* dim)
{
    const int nv = nx + nu, ne = nv * (nv + 1) / 2 + nv, r2 = 2 * (nbu + nbx + K) + 2 * ns;
    dim[F_G] = nv * nx; dim[F_M] = ne; dim[F_ACL] = nx * nx; dim[F_KG] = nu * nx; dim[F_CC] = nx; dim[F_EE] = nx;
    dim[F_PB] = nx; dim[F_RB] = nx; dim[F_ZV] = nv; dim[F_DUX] = nv; dim[F_KK] = nu; dim[F_DINV] = nu;
    dim[F_UX] = nv; dim[F_PI] = nx; dim[F_LAM] = r2; dim[F_T] = r2; dim[F_DLAM] = r2; dim[F_DT] = r2; dim[F_RD] = r2;
    dim[F_RMC] = r2; dim[F_RG] = nv; dim[F_DPI] = nx; dim[F_GXY] = 2 * K; dim[F_D] = r2; dim[F_RQ] = nv;
    dim[F_B] = nx;
    dim[F_SV] = 2 * ns; dim[F_DSV] = 2 * ns; dim[F_RGS] = 2 * ns; dim[F_ZSI] = 2 * ns; dim[F_RQS] = 2 * ns;
    dim[F_DSV2] = 2 * ns; dim[F_RGS2] = 2 * ns;
    dim[F_RG2] = nv; dim[F_RB2] = nx; dim[F_RD2] = r2; dim[F_RM2] = r2; dim[F_DUX2] = nv; dim[F_DPI2] = nx;
    dim[F_DLAM2] = r2; dim[F_DT2] = r2;
}

// Place the working set: chain fields and constants in shared memory, then the pass fields in priority order while
// they fit `smem_budget` bytes; the rest goes to the block's global scratch.  Returns false if even the chain fields
// do not fit.
inline bool make_plan(int nx, int nu, int N, int K, int nbx, int nbu, int ns, int warps, long smem_budget, Plan* out)
{
    Plan P;
    const int N1 = N + 1, nv = nx + nu, ne = nv * (nv + 1) / 2 + nv;
    int dim[F_COUNT];
    field_dims(nx, nu, K, nbx, nbu, ns, dim);
    long o = 0;
    P.const_off = (int) o;
    o += 2 * nv * nv + 3 * ne;                          // Hs, Hes, Tp
    o = round_up((int) o, 2);
    P.red_off = (int) o;
    o += 2 * warps * 8;                                  // two reduction buffers of 8 values per warp
    P.misc_off = (int) o;
    o += nv * nx + chain_w_doubles(nx, nu) + chain_p_doubles(nx) + (nx + 2 * (nu + nx + K) + 8) / 2 + 8;  // A0, chain scratch, int tables
    o += 34;                                             // zero slot + one dump slot per lane of the recursions
    o += F_COUNT;                                        // field id -> address table
    o = round_up((int) o, 2);
    // [B';A'] and the closed-loop matrices exist for stages 0 .. N-1 only
    auto stages = [&](int i) { return (i == F_G || i == F_ACL) ? N : N1; };
    for (int i = 0; i < F_FIRST_FLEX; i++)
    {
        P.f[i].off = (int) o; P.f[i].stride = dim[i]; P.f[i].space = 0;
        o += (long) dim[i] * stages(i);
    }
    if (o * 8 > smem_budget) return false;
    long g = 0;
    for (int i = F_FIRST_FLEX; i < F_COUNT; i++)
    {
        const long need = (long) dim[i] * N1;
        P.f[i].stride = dim[i];
        if ((o + need) * 8 <= smem_budget) { P.f[i].off = (int) o; P.f[i].space = 0; o += need; }
        else { P.f[i].off = (int) g; P.f[i].space = 1; g += need; }
    }
    P.smem_doubles = round_up((int) o, 2);
    P.scratch_doubles = round_up((int) g, 16);
    *out = P;
    return true;
}

}  // namespace usvmpc
