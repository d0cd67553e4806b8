// tests/emu/emu_solver.cpp -- TEST INFRASTRUCTURE ONLY.
//
// Builds the product's warp-per-instance kernel source (mpc_collisionavoidance_b200/csrc/nmpc_kernel.cuh)
// for the CPU with -DUSVMPC_EMULATE: the 32 lanes of a warp become 32 cooperative fibers that switch at
// every shuffle / __syncwarp (csrc/warp_compat.h).  This lets the CPU test-suite (-m "not gpu") check the
// kernel's host-visible logic -- indexing, masks, the IPM control flow -- against the oracle, and lets
// AddressSanitizer watch the shared-memory / workspace accesses, on a machine without a GPU.
// Nothing in the product loads this library.
#include <pthread.h>
#include <time.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "nmpc_kernel.cuh"

namespace usvmpc {
namespace emu {

thread_local Warp* g_warp = nullptr;

asm(".text\n"
    ".globl usvmpc_fiber_switch\n"
    ".type usvmpc_fiber_switch,@function\n"
    "usvmpc_fiber_switch:\n"
    "  pushq %rbp\n  pushq %rbx\n  pushq %r12\n  pushq %r13\n  pushq %r14\n  pushq %r15\n"
    "  movq %rsp, (%rdi)\n"
    "  movq %rsi, %rsp\n"
    "  popq %r15\n  popq %r14\n  popq %r13\n  popq %r12\n  popq %rbx\n  popq %rbp\n"
    "  ret\n"
    ".size usvmpc_fiber_switch, .-usvmpc_fiber_switch\n");

static void fiber_entry()
{
    Warp* w = g_warp;
    w->body(w->arg);
    const int me = w->cur;
    if (me + 1 < WARP) { w->cur = me + 1; usvmpc_fiber_switch(&w->sp[me], w->sp[me + 1]); }
    else usvmpc_fiber_switch(&w->sp[me], w->main_sp);
    fprintf(stderr, "usvmpc emu: a finished lane was resumed (non-uniform sync point in the kernel)\n");
    abort();
}

void run_warp(void (*body)(void*), void* arg)
{
    const size_t STACK = 512 * 1024;
    Warp w;
    memset(&w, 0, sizeof(w));
    w.stacks = (char*) aligned_alloc(64, STACK * WARP);
    w.body = body; w.arg = arg;
    for (int i = 0; i < WARP; i++)
    {
        uint64_t* sp = (uint64_t*) (w.stacks + STACK * (i + 1));
        *(--sp) = 0;
        *(--sp) = (uint64_t) (uintptr_t) &fiber_entry;
        for (int r = 0; r < 6; r++) *(--sp) = 0;
        w.sp[i] = sp;
    }
    Warp* saved = g_warp;
    g_warp = &w;
    w.cur = 0;
    usvmpc_fiber_switch(&w.main_sp, w.sp[0]);
    g_warp = saved;
    free(w.stacks);
}

}  // namespace emu
}  // namespace usvmpc

using namespace usvmpc;

namespace {

struct Job {
    const Params* P;
    int model;
    int* next;
    int smem_doubles;
};

struct LaneArg { const Params* P; int inst; double* sm; int model; };

void lane_body(void* a)
{
    LaneArg* la = (LaneArg*) a;
    if (la->model == 1) { WarpSolver<Pendulum> s(*la->P, la->inst, la->sm); s.run(la->inst); }
    else { WarpSolver<Usv3> s(*la->P, la->inst, la->sm); s.run(la->inst); }
}

void* worker(void* arg)
{
    Job* j = (Job*) arg;
    std::vector<double> sm(j->smem_doubles + 16);
    for (;;)
    {
        int i = __atomic_fetch_add(j->next, 1, __ATOMIC_RELAXED);
        if (i >= j->P->B) break;
        for (auto& v : sm) v = 0.0 / 0.0;  // poison: uninitialised shared memory must not be consumed
        LaneArg la{j->P, i, sm.data(), j->model};
        emu::run_warp(lane_body, &la);
    }
    return nullptr;
}

enum { ICFG_MODEL, ICFG_N, ICFG_K, ICFG_NUM_STEPS, ICFG_NUM_STAGES, ICFG_NLP_TYPE, ICFG_MAX_ITER, ICFG_QP_ITER_MAX,
       ICFG_COND_N, ICFG_NBX, ICFG_NBU, ICFG_PRINT };
enum { DCFG_DT, DCFG_TOL_STAT, DCFG_TOL_EQ, DCFG_TOL_INEQ, DCFG_TOL_COMP, DCFG_UH };

}  // namespace

// same calling convention as oracle/usv_oracle.c:usvo_solve_batch so the tests can swap one for the other;
// optional initial guess (xinit [B][N+1][nx], uinit [B][N][nu], piinit [B][N][nx]) and full multiplier output.
extern "C" double usvemu_solve_batch(const int* icfg, const double* dcfg, const double* W, const double* We,
                                     const double* lbu, const double* ubu, const int* idxbx, const double* lbx,
                                     const double* ubx, int B, const double* x0, const double* p, int p_per_stage,
                                     const double* lh, int lh_per_stage, const double* yref, int yref_per_stage,
                                     const double* yref_e, const double* xinit, const double* uinit,
                                     const double* piinit, double* x_out, double* u_out, double* pi_out,
                                     double* lam_out, double* t_out, double* stats, int nthreads)
{
    const int model = icfg[ICFG_MODEL];
    const int nx = model == 1 ? 4 : 6, nu = model == 1 ? 1 : 2, nv = nx + nu;
    Params P;
    memset(&P, 0, sizeof(P));
    P.B = B; P.N = icfg[ICFG_N]; P.K = icfg[ICFG_K]; P.num_steps = icfg[ICFG_NUM_STEPS];
    P.num_stages = icfg[ICFG_NUM_STAGES]; P.nlp_type = icfg[ICFG_NLP_TYPE]; P.max_iter = icfg[ICFG_MAX_ITER];
    P.qp_iter_max = icfg[ICFG_QP_ITER_MAX]; P.nbx = icfg[ICFG_NBX]; P.nbu = icfg[ICFG_NBU];
    for (int i = 0; i < P.nbx; i++) { P.idxbx[i] = idxbx[i]; P.lbx[i] = lbx[i]; P.ubx[i] = ubx[i]; }
    for (int i = 0; i < P.nbu; i++) { P.lbu[i] = lbu[i]; P.ubu[i] = ubu[i]; }
    P.p_per_stage = p_per_stage; P.lh_per_stage = lh_per_stage; P.yref_per_stage = yref_per_stage;
    P.cold_start = xinit ? 0 : 1;
    P.ncq = P.nbu + P.nbx + P.K; P.ncz = P.nbu + nx + P.K;
    P.dt = dcfg[DCFG_DT]; P.uh = dcfg[DCFG_UH];
    for (int i = 0; i < 4; i++) P.tol[i] = dcfg[DCFG_TOL_STAT + i];
    std::vector<double> cst(nv * nv + nx * nx);
    memcpy(cst.data(), W, sizeof(double) * nv * nv);
    memcpy(cst.data() + nv * nv, We, sizeof(double) * nx * nx);
    P.cst = cst.data();
    P.x0 = x0; P.p = p; P.lh = lh; P.yref = yref; P.yref_e = yref_e;
    P.lay = make_layout(nx, nu, P.N, P.K, P.nbx, P.nbu);
    P.ws_stride = P.lay.total;
    if (!(model == 1 ? WarpSolver<Pendulum>::layout_matches(P.lay) : WarpSolver<Usv3>::layout_matches(P.lay)))
    {
        fprintf(stderr, "usvmpc emu: layout.h and the kernel's compile-time offsets disagree\n");
        abort();
    }
    // exact-size heap block (so a sanitizer sees overruns), poisoned with NaN
    double* ws = (double*) malloc(sizeof(double) * P.ws_stride * B);
    for (long i = 0; i < (long) P.ws_stride * B; i++) ws[i] = 0.0 / 0.0;
    P.ws = ws;
    std::vector<double> st((size_t) B * NSTAT, 0.0);
    P.stats = st.data();
    const Layout& Y = P.lay;
    const int N = P.N, ncz = P.ncz;
    for (int b = 0; b < B; b++)
    {
        double* w = ws + (long) b * P.ws_stride;
        // a freshly created solver: multipliers zero (cold_start() does the same when no guess is given)
        for (int k = 0; k <= N; k++)
        {
            for (int j = 0; j < 2 * ncz; j++) { w[Y.zlam.off + k * Y.zlam.stride + j] = 0.0; w[Y.zt.off + k * Y.zt.stride + j] = 0.0; }
            if (xinit)
            {
                for (int i = 0; i < nu; i++) w[Y.zux.off + k * Y.zux.stride + i] = (k < N && uinit) ? uinit[((long) b * N + k) * nu + i] : 0.0;
                for (int i = 0; i < nx; i++) w[Y.zux.off + k * Y.zux.stride + nu + i] = xinit[((long) b * (N + 1) + k) * nx + i];
                for (int i = 0; i < nx; i++) w[Y.zpi.off + k * Y.zpi.stride + i] = (k < N && piinit) ? piinit[((long) b * N + k) * nx + i] : 0.0;
            }
        }
    }
    Job job;
    int next = 0;
    job.P = &P; job.model = model; job.next = &next; job.smem_doubles = warp_smem_doubles(nx, nu, P.N, P.K, P.nbx, P.nbu);
    if (nthreads < 1) nthreads = 1;
    if (nthreads > 64) nthreads = 64;
    struct timespec t0, t1;
    clock_gettime(CLOCK_MONOTONIC, &t0);
    pthread_t th[64];
    for (int t = 1; t < nthreads; t++) pthread_create(&th[t], nullptr, worker, &job);
    worker(&job);
    for (int t = 1; t < nthreads; t++) pthread_join(th[t], nullptr);
    clock_gettime(CLOCK_MONOTONIC, &t1);
    for (int b = 0; b < B; b++)
    {
        const double* w = ws + (long) b * P.ws_stride;
        for (int k = 0; k <= N; k++)
        {
            for (int i = 0; i < nx; i++) x_out[((long) b * (N + 1) + k) * nx + i] = w[Y.zux.off + k * Y.zux.stride + nu + i];
            if (k < N)
            {
                for (int i = 0; i < nu; i++) u_out[((long) b * N + k) * nu + i] = w[Y.zux.off + k * Y.zux.stride + i];
                if (pi_out) for (int i = 0; i < nx; i++) pi_out[((long) b * N + k) * nx + i] = w[Y.zpi.off + k * Y.zpi.stride + i];
            }
            if (lam_out) for (int j = 0; j < 2 * ncz; j++) lam_out[((long) b * (N + 1) + k) * 2 * ncz + j] = w[Y.zlam.off + k * Y.zlam.stride + j];
            if (t_out) for (int j = 0; j < 2 * ncz; j++) t_out[((long) b * (N + 1) + k) * 2 * ncz + j] = w[Y.zt.off + k * Y.zt.stride + j];
        }
        for (int i = 0; i < 9; i++) stats[(long) b * 9 + i] = st[(size_t) b * NSTAT + i];
    }
    free(ws);
    return (t1.tv_sec - t0.tv_sec) + 1e-9 * (t1.tv_nsec - t0.tv_nsec);
}
