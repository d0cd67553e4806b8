// tests/emu/emu_solver.cpp -- TEST INFRASTRUCTURE ONLY.
//
// Builds the product's CTA-per-instance kernel source (mpc_collisionavoidance_b200/csrc/cta_kernel.cuh) for the
// CPU with -DUSVMPC_EMULATE: the threads of a thread block become cooperative fibers with real warp / block
// barriers (csrc/cta_compat.h); every OS worker thread plays one persistent block pulling instances from the same
// work queue the device uses.  This lets the CPU test-suite (-m "not gpu") check the kernel's logic -- indexing,
// masks, barriers, the IPM control flow -- against the oracle, and lets AddressSanitizer watch the shared-memory /
// workspace accesses, on a machine without a GPU.  Nothing in the product loads this library.
#include <pthread.h>
#include <time.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "cta_kernel.cuh"

namespace usvmpc {
namespace emu {

thread_local Block* g_blk = nullptr;

extern "C" void usvmpc_fiber_switch(void** save_sp, void* new_sp);
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

static inline bool runnable(const Fiber& f) { return !f.done && (f.wait == nullptr || f.wait->gen != f.wait_gen); }

// hand the processor to the next runnable fiber (same warp first); if every fiber is finished, back to the caller of
// run_block; if fibers remain but none can run, the kernel has a divergent barrier
static void switch_next()
{
    Block* b = g_blk;
    const int me = b->cur, T = b->T, W = T / WARP, w0 = me / WARP, l0 = me % WARP;
    int next = -1;
    for (int i = 1; i <= WARP && next < 0; i++)
    {
        const int j = w0 * WARP + (l0 + i) % WARP;
        if (j != me && runnable(b->f[j])) next = j;
    }
    for (int dw = 1; dw < W && next < 0; dw++)
    {
        const int wq = (w0 + dw) % W;
        for (int l = 0; l < WARP; l++)
            if (runnable(b->f[wq * WARP + l])) { next = wq * WARP + l; break; }
    }
    if (next < 0 && runnable(b->f[me])) return;
    if (next < 0)
    {
        bool all_done = true;
        for (int i = 0; i < T; i++) all_done = all_done && b->f[i].done;
        if (!all_done)
        {
            fprintf(stderr, "usvmpc emu: deadlock -- some threads wait on a barrier the others never reach\n");
            abort();
        }
        usvmpc_fiber_switch(&b->f[me].sp, b->main_sp);
        abort();
    }
    b->cur = next;
    usvmpc_fiber_switch(&b->f[me].sp, b->f[next].sp);
}

void arrive(Barrier* bar)
{
    Block* b = g_blk;
    if (++bar->count == bar->need) { bar->count = 0; bar->gen++; return; }
    Fiber& me = b->f[b->cur];
    me.wait = bar; me.wait_gen = bar->gen;
    while (bar->gen == me.wait_gen) switch_next();
    me.wait = nullptr;
}

static void fiber_entry()
{
    Block* b = g_blk;
    b->body(b->arg);
    b->f[b->cur].done = true;
    switch_next();
    fprintf(stderr, "usvmpc emu: a finished thread was resumed\n");
    abort();
}

void run_block(int T, void (*body)(void*), void* arg)
{
    const size_t STACK = 256 * 1024;
    if (T % WARP || T > MAXT) abort();
    Block* b = (Block*) calloc(1, sizeof(Block));
    b->T = T; b->body = body; b->arg = arg;
    b->stacks = (char*) aligned_alloc(64, STACK * T);
    b->cta.need = T;
    for (int w = 0; w < T / WARP; w++) b->warp[w].need = WARP;
    for (int i = 0; i < T; i++)
    {
        uint64_t* sp = (uint64_t*) (b->stacks + STACK * (i + 1));
        *(--sp) = 0;
        *(--sp) = (uint64_t) (uintptr_t) &fiber_entry;
        for (int r = 0; r < 6; r++) *(--sp) = 0;
        b->f[i].sp = sp;
    }
    Block* saved = g_blk;
    g_blk = b;
    b->cur = 0;
    usvmpc_fiber_switch(&b->main_sp, b->f[0].sp);
    g_blk = saved;
    free(b->stacks);
    free(b);
}

}  // namespace emu
}  // namespace usvmpc

using namespace usvmpc;

namespace {

struct Job {
    const Params* P;
    int model, threads;
    int next_block;
};

struct BlockArg { const Params* P; int model; double* sm; int block; };

void block_body(void* a)
{
    BlockArg* ba = (BlockArg*) a;
    const bool soft = ba->P->ns > 0;
    if (ba->model == 1) cta_main<Pendulum, false>(*ba->P, ba->sm, ba->block);
    else if (ba->model == 2) { if (soft) cta_main<Usv8Ca1, true>(*ba->P, ba->sm, ba->block); else cta_main<Usv8Ca1, false>(*ba->P, ba->sm, ba->block); }
    else if (soft) cta_main<Usv3, true>(*ba->P, ba->sm, ba->block);
    else cta_main<Usv3, false>(*ba->P, ba->sm, ba->block);
}

void* worker(void* arg)
{
    Job* j = (Job*) arg;
    const int block = __atomic_fetch_add(&j->next_block, 1, __ATOMIC_RELAXED);
    std::vector<double> sm(j->P->plan.smem_doubles + 16);
    for (auto& v : sm) v = 0.0 / 0.0;  // poison: uninitialised shared memory must not be consumed
    BlockArg ba{j->P, j->model, sm.data(), block};
    emu::run_block(j->threads, block_body, &ba);
    return nullptr;
}

enum { ICFG_MODEL, ICFG_N, ICFG_K, ICFG_NUM_STEPS, ICFG_NUM_STAGES, ICFG_NLP_TYPE, ICFG_MAX_ITER, ICFG_QP_ITER_MAX,
       ICFG_COND_N, ICFG_NBX, ICFG_NBU, ICFG_PRINT, ICFG_NSH };
enum { DCFG_DT, DCFG_TOL_STAT, DCFG_TOL_EQ, DCFG_TOL_INEQ, DCFG_TOL_COMP, DCFG_UH, DCFG_LSH, DCFG_USH, DCFG_ZL, DCFG_ZU, DCFG_ZZL,
       DCFG_ZZU };

}  // namespace

// Emulation knobs (tests only): shared-memory budget in bytes (forces fields into the global scratch), threads per block
extern "C" void usvemu_configure(long smem_budget, int block_threads, int chain_fp32);
static long g_smem_budget = 227 * 1024;
static int g_block_threads = 256, g_chain_fp32 = 0;
extern "C" void usvemu_configure(long smem_budget, int block_threads, int chain_fp32)
{
    g_smem_budget = smem_budget; g_block_threads = block_threads; g_chain_fp32 = chain_fp32;
}
// optional output of the slack values of the next solve: [B][N][2 nsh] = (sl | su) per stage
static double* sv_out = nullptr;
extern "C" void usvemu_slack_output(double* buf) { sv_out = buf; }

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
    const int nx = model == 1 ? 4 : (model == 2 ? 8 : 6), nu = model == 0 ? 2 : 1, nv = nx + nu;
    Params P;
    memset(&P, 0, sizeof(P));
    P.B = B; P.N = icfg[ICFG_N]; P.K = icfg[ICFG_K]; P.num_steps = icfg[ICFG_NUM_STEPS];
    P.num_stages = icfg[ICFG_NUM_STAGES]; P.nlp_type = icfg[ICFG_NLP_TYPE]; P.max_iter = icfg[ICFG_MAX_ITER];
    P.qp_iter_max = icfg[ICFG_QP_ITER_MAX]; P.nbx = icfg[ICFG_NBX]; P.nbu = icfg[ICFG_NBU];
    const int N = P.N, K = P.K;
    for (int i = 0; i < P.nbx; i++) P.idxbx[i] = idxbx[i];
    std::vector<double> vlbu((size_t) N * P.nbu + 1), vubu((size_t) N * P.nbu + 1), vlbx((size_t) N * P.nbx + 1),
        vubx((size_t) N * P.nbx + 1), vuh((size_t) N * K + 1);
    for (int k = 0; k < N; k++)
    {
        for (int i = 0; i < P.nbu; i++) { vlbu[k * P.nbu + i] = lbu[i]; vubu[k * P.nbu + i] = ubu[i]; }
        for (int i = 0; i < P.nbx; i++) { vlbx[k * P.nbx + i] = lbx[i]; vubx[k * P.nbx + i] = ubx[i]; }
        for (int i = 0; i < K; i++) vuh[k * K + i] = dcfg[DCFG_UH];
    }
    P.lbu = vlbu.data(); P.ubu = vubu.data(); P.lbx = vlbx.data(); P.ubx = vubx.data(); P.uh = vuh.data();
    const int ns = icfg[ICFG_NSH];
    P.ns = ns;
    std::vector<double> vlsh((size_t) N * ns + 1), vush((size_t) N * ns + 1), vzs((size_t) 4 * ns + 1);
    for (int k = 0; k < N; k++) for (int i = 0; i < ns; i++) { vlsh[k * ns + i] = dcfg[DCFG_LSH]; vush[k * ns + i] = dcfg[DCFG_USH]; }
    for (int i = 0; i < ns; i++) { vzs[i] = dcfg[DCFG_ZL]; vzs[ns + i] = dcfg[DCFG_ZU]; vzs[2 * ns + i] = dcfg[DCFG_ZZL]; vzs[3 * ns + i] = dcfg[DCFG_ZZU]; }
    P.lsh = vlsh.data(); P.ush = vush.data(); P.zs = vzs.data();
    P.p_per_stage = p_per_stage; P.lh_per_stage = lh_per_stage; P.yref_per_stage = yref_per_stage;
    P.cold_start = xinit ? 0 : 1;
    P.chain_fp32 = g_chain_fp32;
    P.ncq = P.nbu + P.nbx + K; P.ncz = P.nbu + nx + K;
    P.dt = dcfg[DCFG_DT];
    for (int i = 0; i < 4; i++) P.tol[i] = dcfg[DCFG_TOL_STAT + i];
    std::vector<double> cst(nv * nv + nx * nx);
    memcpy(cst.data(), W, sizeof(double) * nv * nv);
    memcpy(cst.data() + nv * nv, We, sizeof(double) * nx * nx);
    P.cst = cst.data();
    P.x0 = x0; P.p = p; P.lh = lh; P.yref = yref; P.yref_e = yref_e;
    P.lay = make_layout(nx, nu, N, K, ns);
    P.ws_stride = P.lay.total;
    if (!make_plan(nx, nu, N, K, P.nbx, P.nbu, ns, g_block_threads / 32, g_smem_budget, &P.plan))
    {
        fprintf(stderr, "usvmpc emu: the chain fields do not fit the shared-memory budget\n");
        abort();
    }
    if (nthreads < 1) nthreads = 1;
    if (nthreads > 64) nthreads = 64;
    if (nthreads > B) nthreads = B;
    // exact-size heap blocks (so a sanitizer sees overruns), poisoned with NaN
    double* ws = (double*) malloc(sizeof(double) * P.ws_stride * B);
    for (long i = 0; i < (long) P.ws_stride * B; i++) ws[i] = 0.0 / 0.0;
    P.ws = ws;
    std::vector<double> scratch((size_t) P.plan.scratch_doubles * nthreads + 1, 0.0 / 0.0);
    P.scratch = P.plan.scratch_doubles ? scratch.data() : nullptr;
    std::vector<int> queue(8, 0);
    P.queue = queue.data();
    std::vector<double> st((size_t) B * NSTAT, 0.0);
    P.stats = st.data();
    const Layout& Y = P.lay;
    const int ncz = P.ncz;
    for (int b = 0; b < B; b++)
    {
        double* w = ws + (long) b * P.ws_stride;
        // a freshly created solver: multipliers zero (cold_start() does the same when no guess is given)
        for (int k = 0; k <= N; k++)
        {
            for (int j = 0; j < 2 * ncz + 2 * ns; j++) { w[Y.zlam.off + k * Y.zlam.stride + j] = 0.0; w[Y.zt.off + k * Y.zt.stride + j] = 0.0; }
            for (int j = 0; j < 2 * ns; j++) w[Y.zsv.off + k * Y.zsv.stride + j] = 0.0;
            if (xinit)
            {
                for (int i = 0; i < nu; i++) w[Y.zux.off + k * Y.zux.stride + i] = (k < N && uinit) ? uinit[((long) b * N + k) * nu + i] : 0.0;
                for (int i = 0; i < nx; i++) w[Y.zux.off + k * Y.zux.stride + nu + i] = xinit[((long) b * (N + 1) + k) * nx + i];
                for (int i = 0; i < nx; i++) w[Y.zpi.off + k * Y.zpi.stride + i] = (k < N && piinit) ? piinit[((long) b * N + k) * nx + i] : 0.0;
            }
        }
    }
    Job job;
    job.P = &P; job.model = model; job.threads = g_block_threads; job.next_block = 0;
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
            const int wl = 2 * ncz + 2 * ns;   // [lower | upper | slack bounds]
            if (lam_out) for (int j = 0; j < wl; j++) lam_out[((long) b * (N + 1) + k) * wl + j] = w[Y.zlam.off + k * Y.zlam.stride + j];
            if (t_out) for (int j = 0; j < wl; j++) t_out[((long) b * (N + 1) + k) * wl + j] = w[Y.zt.off + k * Y.zt.stride + j];
            if (sv_out && k < N) for (int j = 0; j < 2 * ns; j++) sv_out[((long) b * N + k) * 2 * ns + j] = w[Y.zsv.off + k * Y.zsv.stride + j];
        }
        for (int i = 0; i < 12; i++) stats[(long) b * 12 + i] = st[(size_t) b * NSTAT + (i == 9 ? 15 : i)];  // slot 9: fp32 factorisations
    }
    free(ws);
    return (t1.tv_sec - t0.tv_sec) + 1e-9 * (t1.tv_nsec - t0.tv_nsec);
}
