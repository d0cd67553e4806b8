// usvmpc_api.cu -- C ABI (include/usvmpc.h) over the CTA-per-instance NMPC kernel.
// Host side: owns the per-instance NLP iterates in HBM and the inputs of B instances, moves caller data in and out
// with the reference's field names, and launches ONE persistent kernel per solve: thread blocks pull instances from a
// device-side queue and solve each of them out of shared memory (csrc/cta_kernel.cuh).
// Built by mpc_collisionavoidance_b200/build.py:  nvcc -gencode arch=compute_100a,code=sm_100a -shared ...
#include <cuda_runtime.h>

#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "../../include/usvmpc.h"
#include "cta_kernel.cuh"

using namespace usvmpc;

namespace {

#ifndef USVMPC_THREADS
#define USVMPC_THREADS 256  // threads of one block (= one instance at a time)
#endif
#ifndef USVMPC_MIN_CTAS
#define USVMPC_MIN_CTAS 2   // resident blocks per SM the register allocation must allow
#endif
constexpr int THREADS = USVMPC_THREADS;
constexpr long SMEM_PER_SM = 227 * 1024, SMEM_RESERVED_PER_CTA = 1024;

thread_local char g_err[512] = "";

int fail(int code, const char* fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}

#define CU(call)                                                                                        \
    do {                                                                                                \
        cudaError_t e_ = (call);                                                                        \
        if (e_ != cudaSuccess) return fail(USVMPC_E_CUDA, "%s: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
    } while (0)

template <class M, bool SOFT>
__global__ void __launch_bounds__(THREADS, USVMPC_MIN_CTAS) nmpc_solve_kernel(const __grid_constant__ Params P)
{
    extern __shared__ double smem[];
    cta_main<M, SOFT>(P, smem, blockIdx.x);
}

// Longest-processing-time-first order of the work queue: instances sorted by decreasing IPM iteration count of the
// PREVIOUS solve of this solver (statistics record, column 2).  In closed-loop / Monte-Carlo use consecutive solves of
// an instance cost about the same, so the expensive instances start first and the launch does not end with a few
// blocks finishing instances they picked up late.  One block, counting sort on min(iterations, NBIN-1).
constexpr int LPT_BINS = 4096;
__global__ void lpt_order_kernel(const double* stats, int B, int* order)
{
    __shared__ int hist[LPT_BINS];
    for (int i = threadIdx.x; i < LPT_BINS; i += blockDim.x) hist[i] = 0;
    __syncthreads();
    for (int b = threadIdx.x; b < B; b += blockDim.x)
    {
        int key = (int) stats[(long) b * NSTAT + 2];
        key = key < 0 ? 0 : (key > LPT_BINS - 1 ? LPT_BINS - 1 : key);
        atomicAdd(&hist[key], 1);
    }
    __syncthreads();
    if (threadIdx.x == 0)
    {
        int run = 0;
        for (int i = LPT_BINS - 1; i >= 0; i--) { const int c = hist[i]; hist[i] = run; run += c; }  // descending keys
    }
    __syncthreads();
    for (int b = threadIdx.x; b < B; b += blockDim.x)
    {
        int key = (int) stats[(long) b * NSTAT + 2];
        key = key < 0 ? 0 : (key > LPT_BINS - 1 ? LPT_BINS - 1 : key);
        order[atomicAdd(&hist[key], 1)] = b;
    }
}

// buf[b][k][i] <-> ws[b*stride + off + (k0+k)*fstride + c0 + i]
__global__ void ws_copy_kernel(double* ws, long stride, int off, int fstride, int c0, int dim, int k0, int nst, int B,
                               double* buf, int to_ws)
{
    const long n = (long) B * nst * dim;
    for (long idx = blockIdx.x * (long) blockDim.x + threadIdx.x; idx < n; idx += (long) gridDim.x * blockDim.x)
    {
        const int i = (int) (idx % dim);
        const int k = (int) ((idx / dim) % nst);
        const long b = idx / ((long) dim * nst);
        double* a = ws + b * stride + off + (long) (k0 + k) * fstride + c0 + i;
        if (to_ws) *a = buf[idx]; else buf[idx] = *a;
    }
}

// multipliers / slacks of one stage between the engine's NLP row layout (cta_layout.h) and the reference's order
// [lbu lbx lh | ubu ubx uh] with the stage's own counts (acados_template/acados_ocp_solver.py:732-735)
__global__ void ws_rows_kernel(double* ws, long stride, int off, int ncz, int nbu, int nxslots, int nbxk, int K, int ns, int B,
                               double* buf, int to_ws)
{
    const int nck = nbu + nbxk + K, w = 2 * nck + 2 * ns;   // [lbu lbx lh | ubu ubx uh | lsh | ush]
    const long n = (long) B * w;
    for (long idx = blockIdx.x * (long) blockDim.x + threadIdx.x; idx < n; idx += (long) gridDim.x * blockDim.x)
    {
        const int j = (int) (idx % w);
        const long b = idx / w;
        int pos;
        if (j >= 2 * nck) pos = 2 * ncz + (j - 2 * nck);
        else
        {
            const int side = j / nck, jj = j % nck;
            pos = side * ncz + (jj < nbu + nbxk ? jj : nbu + nxslots + (jj - nbu - nbxk));
        }
        double* a = ws + b * stride + off + pos;
        if (to_ws) *a = buf[idx]; else buf[idx] = *a;
    }
}

// dst[b][k][i] = src[b][i]
__global__ void broadcast_kernel(double* dst, const double* src, int B, int nst, int dim)
{
    const long n = (long) B * nst * dim;
    for (long idx = blockIdx.x * (long) blockDim.x + threadIdx.x; idx < n; idx += (long) gridDim.x * blockDim.x)
    {
        const int i = (int) (idx % dim);
        const long b = idx / ((long) dim * nst);
        dst[idx] = src[b * dim + i];
    }
}

// LINEAR_LS cost of the current iterate, one thread per instance: sum_k dt/2 ||[x_k;u_k] - yref_k||^2_W + 1/2 ||x_N - yref_e||^2_We
// (ocp_nlp_cost_ls_compute_fun, acados/ocp_nlp/ocp_nlp_cost_ls.c:847; stage scaling dt, acados_solver.in.c:806-810)
__global__ void eval_cost_kernel(Params P, int nx, int nu, double* out)
{
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= P.B) return;
    const int ny = nx + nu, N = P.N;
    const double* W = P.cst;
    const double* We = P.cst + ny * ny;
    const double* w = P.ws + (long) b * P.ws_stride;
    double cost = 0.0;
    for (int k = 0; k <= N; k++)
    {
        const double* z = w + P.lay.zux.off + (long) k * P.lay.zux.stride;
        const double* yr = k < N ? P.yref + ((long) b * (P.yref_per_stage ? N : 1) + (P.yref_per_stage ? k : 0)) * ny : P.yref_e + (long) b * nx;
        const int n = k < N ? ny : nx;
        double acc = 0.0;
        for (int i = 0; i < n; i++)
        {
            const double ri = (i < nx ? z[nu + i] : z[i - nx]) - yr[i];
            double wr = 0.0;
            for (int j = 0; j < n; j++)
            {
                const double rj = (j < nx ? z[nu + j] : z[j - nx]) - yr[j];
                const double wij = k < N ? (i >= j ? W[i + ny * j] : W[j + ny * i]) : (i >= j ? We[i + nx * j] : We[j + nx * i]);
                wr += wij * rj;
            }
            acc += ri * wr;
        }
        cost += (k < N ? P.dt : 1.0) * 0.5 * acc;
    }
    out[b] = cost;
}

// Obstacle front end of the guidance node (nmpc_ca/src/nmpc_guidance_ca1.cpp:251-363): per instance, from up to M
// obstacles given in the BODY frame as (x, y, radius): inflate by the boat radius, keep the K with the smallest
// clearance sqrt(x^2+y^2) - radius when there are more than K, rotate/translate to NED with the vessel pose
// (nedx, nedy, psi), pad with far-away zero-radius obstacles otherwise.  The node does the transform in float
// (Eigen::Vector3f / Matrix3f); so does this kernel.  One thread per instance.
__global__ void obstacle_frontend_kernel(const double* pose, const double* obs, const int* len, int B, int M, int K,
                                         double boat_radius, double init_pos, double* p_out, double* r_out)
{
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    const double* o = obs + (long) b * M * 3;
    const double nedx = pose[b * 3 + 0], nedy = pose[b * 3 + 1], psi = pose[b * 3 + 2];
    const float c = (float) cos(psi), s = (float) sin(psi);
    int n = len[b];
    n = n < 0 ? 0 : (n > M ? M : n);
    unsigned long long taken = 0ull;  // M <= 64
    for (int i = 0; i < K; i++)
    {
        int pick = -1;
        if (n > K)
        {
            // i-th smallest clearance (ties: lower index first)
            double best = 0.0;
            for (int j = 0; j < n; j++)
            {
                if (taken >> j & 1ull) continue;
                const double bx = o[j * 3], by = o[j * 3 + 1], rad = o[j * 3 + 2] + boat_radius;
                const double d = sqrt(bx * bx + by * by) - rad;
                if (pick < 0 || d < best) { pick = j; best = d; }
            }
            taken |= 1ull << pick;
        }
        else if (i < n) pick = i;
        float x, y, r;
        if (pick >= 0)
        {
            const float bx = (float) o[pick * 3], by = (float) o[pick * 3 + 1];
            x = (float) ((double) (c * bx + (-s) * by + 0.0f * 0.0f) + nedx);
            y = (float) ((double) (s * bx + c * by + 0.0f * 0.0f) + nedy);
            r = (float) (o[pick * 3 + 2] + boat_radius);
        }
        else { x = (float) init_pos; y = (float) init_pos; r = 0.0f; }
        p_out[(long) b * 2 * K + 2 * i] = x;
        p_out[(long) b * 2 * K + 2 * i + 1] = y;
        r_out[(long) b * K + i] = r;
    }
}

int grid_for(long n) { long g = (n + 255) / 256; return (int) (g < 1 ? 1 : (g > 148 * 8 ? 148 * 8 : g)); }

}  // namespace

struct usvmpc_solver
{
    usvmpc_config cfg;
    int B, device, nx, nu, nv;
    Params P;
    double *d_ws, *d_stats, *d_cst, *d_x0, *d_yref_e, *d_scratch, *d_prep;
    double *d_bnd;                         // per-stage bounds shared by the batch: lbu | ubu | lbx | ubx | uh
    double *h_bnd;                         // host mirror of d_bnd
    int o_lbu, o_ubu, o_lbx, o_ubx, o_uh, o_lsh, o_ush, o_zs, n_bnd;
    int* d_queue;
    int* d_order;                          // longest-first queue order from the previous solve
    int have_history, lpt;
    double *d_p[2], *d_lh[2], *d_yref[2];  // [0] one row per instance, [1] one row per instance and stage
    double* d_stage;
    size_t stage_bytes;
    long launches;
    int grid, ctas_per_sm, num_sms;
    double sm_clock_hz;
};

namespace {

int model_dims(int model, int* nx, int* nu)
{
    if (model == USVMPC_MODEL_USV3) { *nx = Usv3::NX; *nu = Usv3::NU; return 0; }
    if (model == USVMPC_MODEL_PENDULUM) { *nx = Pendulum::NX; *nu = Pendulum::NU; return 0; }
    if (model == USVMPC_MODEL_GUIDANCE_CA1) { *nx = Usv8Ca1::NX; *nu = Usv8Ca1::NU; return 0; }
    return -1;
}

int upload_constants(usvmpc_solver* s, cudaStream_t st)
{
    const int ny = s->nv, nx = s->nx;
    double tmp[16 * 16 * 2];
    for (int j = 0; j < ny; j++) for (int i = 0; i < ny; i++) tmp[i + ny * j] = s->cfg.W[i + ny * j];
    for (int j = 0; j < nx; j++) for (int i = 0; i < nx; i++) tmp[ny * ny + i + nx * j] = s->cfg.W_e[i + nx * j];
    CU(cudaMemcpyAsync(s->d_cst, tmp, sizeof(double) * (ny * ny + nx * nx), cudaMemcpyHostToDevice, st));
    CU(cudaStreamSynchronize(st));
    return 0;
}

int upload_bounds(usvmpc_solver* s, cudaStream_t st)
{
    CU(cudaMemcpyAsync(s->d_bnd, s->h_bnd, sizeof(double) * s->n_bnd, cudaMemcpyHostToDevice, st));
    CU(cudaStreamSynchronize(st));
    return 0;
}

void refresh_params(usvmpc_solver* s)
{
    Params& P = s->P;
    const usvmpc_config& c = s->cfg;
    P.B = s->B; P.N = c.N; P.K = c.K; P.num_steps = c.num_steps; P.num_stages = c.num_stages; P.nlp_type = c.nlp_type;
    P.max_iter = c.max_iter; P.qp_iter_max = c.qp_iter_max; P.nbx = c.nbx; P.nbu = c.nbu;
    for (int i = 0; i < NBXMAX; i++) P.idxbx[i] = c.idxbx[i];
    P.ncq = c.nbu + c.nbx + c.K; P.ncz = c.nbu + s->nx + c.K; P.ns = c.nsh;
    P.dt = c.dt;
    for (int i = 0; i < 4; i++) P.tol[i] = c.tol[i];
    P.lbu = s->d_bnd + s->o_lbu; P.ubu = s->d_bnd + s->o_ubu; P.lbx = s->d_bnd + s->o_lbx; P.ubx = s->d_bnd + s->o_ubx;
    P.uh = s->d_bnd + s->o_uh; P.lsh = s->d_bnd + s->o_lsh; P.ush = s->d_bnd + s->o_ush; P.zs = s->d_bnd + s->o_zs;
    P.cst = s->d_cst; P.x0 = s->d_x0; P.yref_e = s->d_yref_e;
    P.p = s->d_p[P.p_per_stage]; P.lh = s->d_lh[P.lh_per_stage]; P.yref = s->d_yref[P.yref_per_stage];
    P.ws = s->d_ws; P.stats = s->d_stats; P.scratch = s->d_scratch; P.queue = s->d_queue; P.prep = s->d_prep;
    P.order = (s->lpt && s->have_history) ? s->d_order : nullptr;
}

// `value` as the caller gave it -> host pointer (values shared by the batch are kept on the host)
int to_host(const double* value, int n, int on_device, double* tmp)
{
    if (on_device) CU(cudaMemcpy(tmp, value, sizeof(double) * n, cudaMemcpyDeviceToHost));
    else memcpy(tmp, value, sizeof(double) * n);
    return 0;
}

// make sure the per-stage copy of an input exists and is current before a single stage of it is written
int want_per_stage(usvmpc_solver* s, double** pair, int* flag, int nst, int dim, cudaStream_t st)
{
    if (dim == 0) return 0;
    if (!pair[1]) CU(cudaMalloc(&pair[1], sizeof(double) * (size_t) s->B * nst * dim));
    if (!*flag)
    {
        broadcast_kernel<<<grid_for((long) s->B * nst * dim), 256, 0, st>>>(pair[1], pair[0], s->B, nst, dim);
        CU(cudaGetLastError());
        s->launches++;
        *flag = 1;
    }
    return 0;
}

// write value [B][dim] (stage >= 0), [B][nst][dim] (ALL_STAGES) or [B][dim] for every stage (EVERY_STAGE); the copy kind
// is inferred from the pointers (unified addressing), `on_device` only decides whether the call waits for the copy
int set_input(usvmpc_solver* s, double** pair, int* flag, int nst, int dim, int stage, const double* value, cudaStream_t st)
{
    if (dim == 0) return 0;
    if (stage == USVMPC_EVERY_STAGE)
    {
        CU(cudaMemcpyAsync(pair[0], value, sizeof(double) * (size_t) s->B * dim, cudaMemcpyDefault, st));
        *flag = 0;
    }
    else if (stage == USVMPC_ALL_STAGES)
    {
        if (!pair[1]) CU(cudaMalloc(&pair[1], sizeof(double) * (size_t) s->B * nst * dim));
        CU(cudaMemcpyAsync(pair[1], value, sizeof(double) * (size_t) s->B * nst * dim, cudaMemcpyDefault, st));
        *flag = 1;
    }
    else
    {
        if (stage < 0 || stage >= nst) return fail(USVMPC_E_INVALID, "stage %d out of range [0,%d)", stage, nst);
        int rc = want_per_stage(s, pair, flag, nst, dim, st);
        if (rc) return rc;
        CU(cudaMemcpy2DAsync(pair[1] + (size_t) stage * dim, sizeof(double) * nst * dim, value, sizeof(double) * dim,
                             sizeof(double) * dim, s->B, cudaMemcpyDefault, st));
    }
    refresh_params(s);
    return 0;
}

int need_stage_buf(usvmpc_solver* s, size_t bytes)
{
    if (bytes <= s->stage_bytes) return 0;
    if (s->d_stage) CU(cudaFree(s->d_stage));
    s->d_stage = nullptr; s->stage_bytes = 0;
    CU(cudaMalloc(&s->d_stage, bytes));
    s->stage_bytes = bytes;
    return 0;
}

// field of the NLP iterate: resolve to (layout field, first column, dim, number of stages it exists on)
int out_field(usvmpc_solver* s, const char* field, Field* f, int* c0, int* dim, int* nst)
{
    const Layout& Y = s->P.lay;
    if (!strcmp(field, "x")) { *f = Y.zux; *c0 = s->nu; *dim = s->nx; *nst = s->cfg.N + 1; return 0; }
    if (!strcmp(field, "u")) { *f = Y.zux; *c0 = 0; *dim = s->nu; *nst = s->cfg.N; return 0; }
    if (!strcmp(field, "pi")) { *f = Y.zpi; *c0 = 0; *dim = s->nx; *nst = s->cfg.N; return 0; }
    if (s->cfg.nsh > 0 && !strcmp(field, "sl")) { *f = Y.zsv; *c0 = 0; *dim = s->cfg.nsh; *nst = s->cfg.N; return 0; }
    if (s->cfg.nsh > 0 && !strcmp(field, "su")) { *f = Y.zsv; *c0 = s->cfg.nsh; *dim = s->cfg.nsh; *nst = s->cfg.N; return 0; }
    return -1;
}

int out_copy(usvmpc_solver* s, int stage, const char* field, double* value, int on_device, void* stream, int to_ws)
{
    if (!s || !field || !value) return fail(USVMPC_E_INVALID, "null argument");
    cudaStream_t st = (cudaStream_t) stream;
    CU(cudaSetDevice(s->device));
    const int N = s->cfg.N, B = s->B;
    Field f; int c0, dim, nst;
    if (out_field(s, field, &f, &c0, &dim, &nst) == 0)
    {
        int k0 = stage, ns = 1;
        if (stage == USVMPC_ALL_STAGES) { k0 = 0; ns = nst; }
        else if (stage < 0 || stage >= nst) return fail(USVMPC_E_INVALID, "field %s has no stage %d", field, stage);
        const size_t bytes = sizeof(double) * (size_t) B * ns * dim;
        double* buf = value;
        if (!on_device)
        {
            int rc = need_stage_buf(s, bytes);
            if (rc) return rc;
            buf = s->d_stage;
            if (to_ws) CU(cudaMemcpyAsync(buf, value, bytes, cudaMemcpyHostToDevice, st));
        }
        ws_copy_kernel<<<grid_for((long) B * ns * dim), 256, 0, st>>>(s->d_ws, s->P.ws_stride, f.off, f.stride, c0, dim, k0, ns, B, buf, to_ws);
        CU(cudaGetLastError());
        s->launches++;
        if (!on_device)
        {
            if (!to_ws) CU(cudaMemcpyAsync(value, buf, bytes, cudaMemcpyDeviceToHost, st));
            CU(cudaStreamSynchronize(st));
        }
        return 0;
    }
    if (!strcmp(field, "lam") || !strcmp(field, "t"))
    {
        if (stage < 0 || stage > N) return fail(USVMPC_E_INVALID, "field %s needs a stage in [0,%d]", field, N);
        const int nbxk = stage == 0 ? s->nx : (stage < N ? s->cfg.nbx : 0);
        const int nbuk = stage < N ? s->cfg.nbu : 0, Kk = stage < N ? s->cfg.K : 0;
        const int nck = nbuk + nbxk + Kk, nsk = stage < N ? s->cfg.nsh : 0;
        if (nck == 0) return 0;
        const Field& fl = !strcmp(field, "lam") ? s->P.lay.zlam : s->P.lay.zt;
        const size_t bytes = sizeof(double) * (size_t) B * (2 * nck + 2 * nsk);
        double* buf = value;
        if (!on_device)
        {
            int rc = need_stage_buf(s, bytes);
            if (rc) return rc;
            buf = s->d_stage;
            if (to_ws) CU(cudaMemcpyAsync(buf, value, bytes, cudaMemcpyHostToDevice, st));
        }
        ws_rows_kernel<<<grid_for((long) B * (2 * nck + 2 * nsk)), 256, 0, st>>>(s->d_ws, s->P.ws_stride, fl.off + stage * fl.stride,
                                                                                 s->P.ncz, s->cfg.nbu, s->nx, nbxk, s->cfg.K, nsk, B, buf, to_ws);
        CU(cudaGetLastError());
        s->launches++;
        if (!on_device)
        {
            if (!to_ws) CU(cudaMemcpyAsync(value, buf, bytes, cudaMemcpyDeviceToHost, st));
            CU(cudaStreamSynchronize(st));
        }
        return 0;
    }
    if (!strcmp(field, "sl") || !strcmp(field, "su") || !strcmp(field, "z")) return 0;  // ns = nz = 0
    return fail(USVMPC_E_FIELD, "unknown field '%s' (x, u, pi, lam, t, sl, su, z)", field);
}

static int create_impl(usvmpc_solver* s)
{
    const usvmpc_config* cfg = &s->cfg;
    const int nx = s->nx, nu = s->nu, N = cfg->N, K = cfg->K, B = s->B, ny = nx + nu;
    cudaDeviceProp prop;
    CU(cudaGetDeviceProperties(&prop, s->device));
    s->num_sms = prop.multiProcessorCount;
    int khz = 0;
    CU(cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, s->device));
    s->sm_clock_hz = 1e3 * khz;
    s->P.lay = make_layout(nx, nu, N, K, cfg->nsh);
    s->P.ws_stride = s->P.lay.total;
    // placement of the working set: as many resident blocks per SM as the chain fields allow (at most USVMPC_MIN_CTAS,
    // which bounds the registers), everything else that fits goes to shared memory too, the rest to the L2-resident
    // block scratch.  USVMPC_CTAS_PER_SM overrides (diagnostics).
    int want = USVMPC_MIN_CTAS;
    if (const char* e = getenv("USVMPC_CTAS_PER_SM")) want = atoi(e) > 0 ? atoi(e) : want;
    if (want > USVMPC_MIN_CTAS) want = USVMPC_MIN_CTAS;
    const long max_optin = (long) prop.sharedMemPerBlockOptin;
    bool ok = false;
    for (int c = want; c >= 1 && !ok; c--)
    {
        long budget = SMEM_PER_SM / c - SMEM_RESERVED_PER_CTA;
        if (budget > max_optin) budget = max_optin;
        if (make_plan(nx, nu, N, K, cfg->nbx, cfg->nbu, cfg->nsh, THREADS / 32, budget, &s->P.plan)) { ok = true; s->ctas_per_sm = c; }
    }
    if (!ok) return fail(USVMPC_E_INVALID, "N=%d, K=%d: the Riccati working set does not fit the shared memory of one SM", N, K);
    s->grid = s->ctas_per_sm * s->num_sms;
    if (s->grid > B) s->grid = B;
    const int kk = K > 0 ? K : 1;
    CU(cudaMalloc(&s->d_ws, sizeof(double) * (size_t) s->P.ws_stride * B));
    CU(cudaMemset(s->d_ws, 0, sizeof(double) * (size_t) s->P.ws_stride * B));
    CU(cudaMalloc(&s->d_stats, sizeof(double) * (size_t) B * NSTAT));
    CU(cudaMemset(s->d_stats, 0, sizeof(double) * (size_t) B * NSTAT));
    CU(cudaMalloc(&s->d_cst, sizeof(double) * (ny * ny + nx * nx)));
    CU(cudaMalloc(&s->d_x0, sizeof(double) * (size_t) B * nx));
    CU(cudaMemset(s->d_x0, 0, sizeof(double) * (size_t) B * nx));
    CU(cudaMalloc(&s->d_yref_e, sizeof(double) * (size_t) B * nx));
    CU(cudaMemset(s->d_yref_e, 0, sizeof(double) * (size_t) B * nx));
    CU(cudaMalloc(&s->d_p[0], sizeof(double) * (size_t) B * 2 * kk));
    CU(cudaMemset(s->d_p[0], 0, sizeof(double) * (size_t) B * 2 * kk));
    CU(cudaMalloc(&s->d_lh[0], sizeof(double) * (size_t) B * kk));
    CU(cudaMemset(s->d_lh[0], 0, sizeof(double) * (size_t) B * kk));
    CU(cudaMalloc(&s->d_yref[0], sizeof(double) * (size_t) B * ny));
    CU(cudaMemset(s->d_yref[0], 0, sizeof(double) * (size_t) B * ny));
    if (s->P.plan.scratch_doubles)
    {
        CU(cudaMalloc(&s->d_scratch, sizeof(double) * (size_t) s->P.plan.scratch_doubles * s->grid));
        CU(cudaMemset(s->d_scratch, 0, sizeof(double) * (size_t) s->P.plan.scratch_doubles * s->grid));
    }
    CU(cudaMalloc(&s->d_queue, sizeof(int) * 8));
    CU(cudaMemset(s->d_queue, 0, sizeof(int) * 8));
    CU(cudaMalloc(&s->d_order, sizeof(int) * (size_t) B));
    s->lpt = 1; s->have_history = 0;
    // per-stage bounds shared by the batch, initialised from the description like acados_create() does
    // (acados_solver.in.c:1028-1449: the same lbx/ubx/lbu/ubu/lh/uh on every stage)
    const int nbu = cfg->nbu, nbx = cfg->nbx;
    s->o_lbu = 0; s->o_ubu = s->o_lbu + N * nbu; s->o_lbx = s->o_ubu + N * nbu; s->o_ubx = s->o_lbx + N * nbx;
    const int ns = cfg->nsh;
    s->o_uh = s->o_ubx + N * nbx; s->o_lsh = s->o_uh + N * K; s->o_ush = s->o_lsh + N * ns; s->o_zs = s->o_ush + N * ns;
    s->n_bnd = s->o_zs + 4 * ns + 1;
    s->h_bnd = (double*) calloc(s->n_bnd, sizeof(double));
    if (!s->h_bnd) return fail(USVMPC_E_INVALID, "out of host memory");
    for (int k = 0; k < N; k++)
    {
        for (int i = 0; i < nbu; i++) { s->h_bnd[s->o_lbu + k * nbu + i] = cfg->lbu[i]; s->h_bnd[s->o_ubu + k * nbu + i] = cfg->ubu[i]; }
        for (int i = 0; i < nbx; i++) { s->h_bnd[s->o_lbx + k * nbx + i] = cfg->lbx[i]; s->h_bnd[s->o_ubx + k * nbx + i] = cfg->ubx[i]; }
        for (int i = 0; i < K; i++) s->h_bnd[s->o_uh + k * K + i] = cfg->uh;
        for (int i = 0; i < ns; i++) { s->h_bnd[s->o_lsh + k * ns + i] = cfg->lsh[i]; s->h_bnd[s->o_ush + k * ns + i] = cfg->ush[i]; }
    }
    for (int i = 0; i < ns; i++)
    {
        s->h_bnd[s->o_zs + i] = cfg->zl[i]; s->h_bnd[s->o_zs + ns + i] = cfg->zu[i];
        s->h_bnd[s->o_zs + 2 * ns + i] = cfg->Zl[i]; s->h_bnd[s->o_zs + 3 * ns + i] = cfg->Zu[i];
    }
    CU(cudaMalloc(&s->d_bnd, sizeof(double) * s->n_bnd));
    int rc = upload_bounds(s, 0);
    if (rc) return rc;
    refresh_params(s);
    return upload_constants(s, 0);
}

template <class M, bool SOFT>
static int launch_solve(usvmpc_solver* s, cudaStream_t st)
{
    const size_t smem = sizeof(double) * (size_t) s->P.plan.smem_doubles;
    // the attribute is per function and process: set it for THIS solver's size right before its launch
    CU(cudaFuncSetAttribute(nmpc_solve_kernel<M, SOFT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem));
    CU(cudaMemsetAsync(s->d_queue, 0, sizeof(int) * 8, st));
    if (s->lpt && s->have_history && s->B > s->grid)
    {
        lpt_order_kernel<<<1, 1024, 0, st>>>(s->d_stats, s->B, s->d_order);
        CU(cudaGetLastError());
        s->launches++;
    }
    refresh_params(s);
    if (s->B <= s->grid) s->P.order = nullptr;
    nmpc_solve_kernel<M, SOFT><<<s->grid, THREADS, smem, st>>>(s->P);
    CU(cudaGetLastError());
    s->have_history = 1;
    return 0;
}

// bounds shared by the batch: one row per stage like the reference's nlp_in (ocp_nlp_constraints_bgh.c:630-822);
// stage = USVMPC_EVERY_STAGE writes the row of every stage
static int set_shared_bound(usvmpc_solver* s, int off, int dim, int k_first, int stage, const double* value, int on_device,
                            cudaStream_t st)
{
    const int N = s->cfg.N;
    if (dim == 0) return 0;
    double tmp[KMAX > NBXMAX ? KMAX : NBXMAX];
    int rc = to_host(value, dim, on_device, tmp);
    if (rc) return rc;
    if (stage == USVMPC_EVERY_STAGE) { for (int k = 0; k < N; k++) memcpy(s->h_bnd + off + k * dim, tmp, sizeof(double) * dim); }
    else
    {
        if (stage < k_first || stage >= N) return fail(USVMPC_E_INVALID, "no such bound at stage %d", stage);
        memcpy(s->h_bnd + off + stage * dim, tmp, sizeof(double) * dim);
    }
    return upload_bounds(s, st);
}

// the kernel instantiation of the solver's model / constraint kind
static int launch_model(usvmpc_solver* s, cudaStream_t st)
{
    const bool soft = s->cfg.nsh > 0;
    switch (s->cfg.model)
    {
    case USVMPC_MODEL_PENDULUM: return launch_solve<Pendulum, false>(s, st);
    case USVMPC_MODEL_GUIDANCE_CA1: return soft ? launch_solve<Usv8Ca1, true>(s, st) : launch_solve<Usv8Ca1, false>(s, st);
    default: return soft ? launch_solve<Usv3, true>(s, st) : launch_solve<Usv3, false>(s, st);
    }
}

}  // namespace

extern "C" {

const char* usvmpc_last_error(void) { return g_err; }
const char* usvmpc_version(void) { return "usvmpc 0.2 (sm_100a, block-per-instance fp64, shared-memory resident IPM)"; }

int usvmpc_config_default(usvmpc_config* c, int model)
{
    if (!c) return fail(USVMPC_E_INVALID, "null config");
    int nx, nu;
    if (model_dims(model, &nx, &nu)) return fail(USVMPC_E_INVALID, "unknown model %d", model);
    memset(c, 0, sizeof(*c));
    c->model = model; c->N = 20; c->K = 0; c->num_steps = 1; c->num_stages = 4;
    c->nlp_type = USVMPC_SQP_RTI; c->max_iter = 100; c->qp_iter_max = 50;
    c->dt = 0.05; c->uh = 1e6;
    for (int i = 0; i < 4; i++) c->tol[i] = 1e-6;
    for (int i = 0; i < nx + nu; i++) c->W[i + (nx + nu) * i] = 1.0;
    for (int i = 0; i < nx; i++) c->W_e[i + nx * i] = 1.0;
    return 0;
}

int usvmpc_create(const usvmpc_config* cfg, int batch, int device, usvmpc_solver** out)
{
    if (!cfg || !out) return fail(USVMPC_E_INVALID, "null argument");
    *out = nullptr;
    int nx, nu;
    if (model_dims(cfg->model, &nx, &nu)) return fail(USVMPC_E_INVALID, "unknown model %d", cfg->model);
    if (batch < 1 || cfg->N < 1) return fail(USVMPC_E_INVALID, "batch and N must be >= 1");
    if (cfg->N > NMAX) return fail(USVMPC_E_INVALID, "N=%d: the engine is built for horizons up to %d", cfg->N, NMAX);
    if (cfg->K < 0 || cfg->K > KMAX) return fail(USVMPC_E_INVALID, "K=%d outside [0,%d]", cfg->K, KMAX);
    if (cfg->nsh < 0 || cfg->nsh > cfg->K) return fail(USVMPC_E_INVALID, "nsh=%d outside [0,K=%d]", cfg->nsh, cfg->K);
    if (cfg->nbx < 0 || cfg->nbx > nx || cfg->nbx > NBXMAX || cfg->nbu < 0 || cfg->nbu > nu || cfg->nbu > NBUMAX)
        return fail(USVMPC_E_INVALID, "nbx/nbu out of range");
    if (cfg->num_stages != 1 && cfg->num_stages != 2 && cfg->num_stages != 4) return fail(USVMPC_E_INVALID, "ERK num_stages must be 1, 2 or 4");
    if (cfg->num_steps < 1) return fail(USVMPC_E_INVALID, "num_steps must be >= 1");
    for (int j = 0; j < cfg->nbx; j++)
        if (cfg->idxbx[j] < 0 || cfg->idxbx[j] >= nx) return fail(USVMPC_E_INVALID, "idxbx[%d] out of range", j);
    int ndev = 0;
    CU(cudaGetDeviceCount(&ndev));
    if (device < 0 || device >= ndev) return fail(USVMPC_E_CUDA, "no CUDA device %d (found %d): the engine has no CPU path", device, ndev);
    CU(cudaSetDevice(device));
    usvmpc_solver* s = (usvmpc_solver*) calloc(1, sizeof(usvmpc_solver));
    if (!s) return fail(USVMPC_E_INVALID, "out of host memory");
    s->cfg = *cfg; s->B = batch; s->device = device; s->nx = nx; s->nu = nu; s->nv = nx + nu;
    const int rc = create_impl(s);
    if (rc) { usvmpc_free(s); return rc; }  // one cleanup path: everything allocated so far is released
    *out = s;
    return 0;
}

int usvmpc_free(usvmpc_solver* s)
{
    if (!s) return 0;
    cudaSetDevice(s->device);
    cudaFree(s->d_ws); cudaFree(s->d_stats); cudaFree(s->d_cst); cudaFree(s->d_x0); cudaFree(s->d_yref_e);
    cudaFree(s->d_scratch); cudaFree(s->d_prep); cudaFree(s->d_bnd); cudaFree(s->d_queue); cudaFree(s->d_order);
    for (int i = 0; i < 2; i++) { cudaFree(s->d_p[i]); cudaFree(s->d_lh[i]); cudaFree(s->d_yref[i]); }
    cudaFree(s->d_stage);
    free(s->h_bnd);
    free(s);
    return 0;
}

int usvmpc_solve(usvmpc_solver* s, void* stream)
{
    if (!s) return fail(USVMPC_E_INVALID, "null solver");
    CU(cudaSetDevice(s->device));
    cudaStream_t st = (cudaStream_t) stream;
    const int rc = launch_model(s, st);
    if (rc) return rc;
    s->launches++;
    return 0;
}

int usvmpc_update_params(usvmpc_solver* s, int stage, const double* value, int np, int on_device, void* stream)
{
    if (!s || !value) return fail(USVMPC_E_INVALID, "null argument");
    if (np != 2 * s->cfg.K) return fail(USVMPC_E_SIZE, "np=%d but the model has np=%d", np, 2 * s->cfg.K);  // tpl :1746-1751
    CU(cudaSetDevice(s->device));
    (void) on_device;  // per-instance values: the copy kind is inferred from the pointer (unified addressing)
    return set_input(s, s->d_p, &s->P.p_per_stage, s->cfg.N + 1, 2 * s->cfg.K, stage, value, (cudaStream_t) stream);
}

int usvmpc_cost_model_set(usvmpc_solver* s, int stage, const char* field, const double* value, int on_device, void* stream)
{
    if (!s || !field || !value) return fail(USVMPC_E_INVALID, "null argument");
    CU(cudaSetDevice(s->device));
    cudaStream_t st = (cudaStream_t) stream;
    const int N = s->cfg.N, ny = s->nv, nx = s->nx;
    if (!strcmp(field, "yref") || !strcmp(field, "y_ref"))
    {
        if (stage == N)
        {
            CU(cudaMemcpyAsync(s->d_yref_e, value, sizeof(double) * (size_t) s->B * nx, cudaMemcpyDefault, st));
            return 0;
        }
        return set_input(s, s->d_yref, &s->P.yref_per_stage, N, ny, stage, value, st);
    }
    if (!strcmp(field, "W"))
    {
        // shared by the batch and by all path stages; column-major like the reference (tpl :854)
        double tmp[16 * 16];
        const int n = stage == N ? nx * nx : ny * ny;
        int rc = to_host(value, n, on_device, tmp);
        if (rc) return rc;
        memcpy(stage == N ? s->cfg.W_e : s->cfg.W, tmp, sizeof(double) * n);
        return upload_constants(s, st);
    }
    if (!strcmp(field, "zl") || !strcmp(field, "zu") || !strcmp(field, "Zl") || !strcmp(field, "Zu"))
    {
        // slack penalties: shared by the batch and by the stages (ocp_nlp_cost_ls.c:826-841)
        const int ns = s->cfg.nsh, which = field[0] == 'z' ? (field[1] == 'l' ? 0 : 1) : (field[1] == 'l' ? 2 : 3);
        if (ns == 0) return 0;
        double tmp[KMAX];
        int rc = to_host(value, ns, on_device, tmp);
        if (rc) return rc;
        memcpy(s->h_bnd + s->o_zs + which * ns, tmp, sizeof(double) * ns);
        return upload_bounds(s, st);
    }
    return fail(USVMPC_E_FIELD, "unknown cost field '%s' (yref, y_ref, W, zl, zu, Zl, Zu)", field);
}

int usvmpc_constraints_model_set(usvmpc_solver* s, int stage, const char* field, const double* value, int on_device, void* stream)
{
    if (!s || !field || !value) return fail(USVMPC_E_INVALID, "null argument");
    CU(cudaSetDevice(s->device));
    cudaStream_t st = (cudaStream_t) stream;
    const int N = s->cfg.N;
    if (!strcmp(field, "lbx") || !strcmp(field, "ubx"))
    {
        if (stage == 0)
        {
            CU(cudaMemcpyAsync(s->d_x0, value, sizeof(double) * (size_t) s->B * s->nx, cudaMemcpyDefault, st));
            return 0;
        }
        return set_shared_bound(s, field[0] == 'l' ? s->o_lbx : s->o_ubx, s->cfg.nbx, 1, stage, value, on_device, st);
    }
    if (!strcmp(field, "lbu") || !strcmp(field, "ubu"))
        return set_shared_bound(s, field[0] == 'l' ? s->o_lbu : s->o_ubu, s->cfg.nbu, 0, stage, value, on_device, st);
    if (!strcmp(field, "lh"))
    {
        return set_input(s, s->d_lh, &s->P.lh_per_stage, N, s->cfg.K, stage, value, st);
    }
    if (!strcmp(field, "uh")) return set_shared_bound(s, s->o_uh, s->cfg.K, 0, stage, value, on_device, st);
    if (!strcmp(field, "lsh")) return set_shared_bound(s, s->o_lsh, s->cfg.nsh, 0, stage, value, on_device, st);
    if (!strcmp(field, "ush")) return set_shared_bound(s, s->o_ush, s->cfg.nsh, 0, stage, value, on_device, st);
    return fail(USVMPC_E_FIELD, "unknown constraint field '%s' (lbx, ubx, lbu, ubu, lh, uh)", field);
}

int usvmpc_out_set(usvmpc_solver* s, int stage, const char* field, const double* value, int on_device, void* stream)
{
    return out_copy(s, stage, field, (double*) value, on_device, stream, 1);
}

int usvmpc_out_get(usvmpc_solver* s, int stage, const char* field, double* value, int on_device, void* stream)
{
    return out_copy(s, stage, field, value, on_device, stream, 0);
}

int usvmpc_dims_get_from_attr(usvmpc_solver* s, int stage, const char* field)
{
    if (!s || !field) return fail(USVMPC_E_INVALID, "null argument");
    const int N = s->cfg.N;
    if (stage < 0 || stage > N) return fail(USVMPC_E_INVALID, "stage %d out of range", stage);
    const int nbxk = stage == 0 ? s->nx : (stage < N ? s->cfg.nbx : 0);
    const int nbuk = stage < N ? s->cfg.nbu : 0, Kk = stage < N ? s->cfg.K : 0;
    if (!strcmp(field, "x")) return s->nx;
    if (!strcmp(field, "u")) return stage < N ? s->nu : 0;
    if (!strcmp(field, "pi")) return stage < N ? s->nx : 0;
    const int nsk = stage < N ? s->cfg.nsh : 0;
    if (!strcmp(field, "lam") || !strcmp(field, "t")) return 2 * (nbuk + nbxk + Kk) + 2 * nsk;
    if (!strcmp(field, "sl") || !strcmp(field, "su") || !strcmp(field, "lsh") || !strcmp(field, "ush") || !strcmp(field, "zl") ||
        !strcmp(field, "zu") || !strcmp(field, "Zl") || !strcmp(field, "Zu"))
        return nsk;
    if (!strcmp(field, "p")) return 2 * s->cfg.K;
    if (!strcmp(field, "yref") || !strcmp(field, "y_ref")) return stage < N ? s->nv : s->nx;
    if (!strcmp(field, "lbx") || !strcmp(field, "ubx")) return nbxk;
    if (!strcmp(field, "lbu") || !strcmp(field, "ubu")) return nbuk;
    if (!strcmp(field, "lh") || !strcmp(field, "uh")) return Kk;
    if (!strcmp(field, "z")) return 0;
    return fail(USVMPC_E_FIELD, "unknown field '%s'", field);
}

int usvmpc_get_stats(usvmpc_solver* s, double* value, int on_device, void* stream)
{
    if (!s || !value) return fail(USVMPC_E_INVALID, "null argument");
    CU(cudaSetDevice(s->device));
    cudaStream_t st = (cudaStream_t) stream;
    CU(cudaMemcpyAsync(value, s->d_stats, sizeof(double) * (size_t) s->B * NSTAT, cudaMemcpyDefault, st));
    if (!on_device) CU(cudaStreamSynchronize(st));
    return 0;
}

int usvmpc_eval_cost(usvmpc_solver* s, double* value, int on_device, void* stream)
{
    if (!s || !value) return fail(USVMPC_E_INVALID, "null argument");
    CU(cudaSetDevice(s->device));
    cudaStream_t st = (cudaStream_t) stream;
    double* buf = value;
    if (!on_device)
    {
        int rc = need_stage_buf(s, sizeof(double) * (size_t) s->B);
        if (rc) return rc;
        buf = s->d_stage;
    }
    eval_cost_kernel<<<(s->B + 127) / 128, 128, 0, st>>>(s->P, s->nx, s->nu, buf);
    CU(cudaGetLastError());
    s->launches++;
    if (!on_device)
    {
        CU(cudaMemcpyAsync(value, buf, sizeof(double) * (size_t) s->B, cudaMemcpyDeviceToHost, st));
        CU(cudaStreamSynchronize(st));
    }
    return 0;
}

int usvmpc_solver_opts_set(usvmpc_solver* s, const char* field, double value)
{
    if (!s || !field) return fail(USVMPC_E_INVALID, "null argument");
    usvmpc_config& c = s->cfg;
    if (!strcmp(field, "max_iter") || !strcmp(field, "nlp_solver_max_iter")) c.max_iter = (int) value;
    else if (!strcmp(field, "qp_iter_max") || !strcmp(field, "qp_solver_iter_max")) c.qp_iter_max = (int) value;
    else if (!strcmp(field, "tol_stat")) c.tol[0] = value;
    else if (!strcmp(field, "tol_eq")) c.tol[1] = value;
    else if (!strcmp(field, "tol_ineq")) c.tol[2] = value;
    else if (!strcmp(field, "tol_comp")) c.tol[3] = value;
    else if (!strcmp(field, "nlp_solver_type")) c.nlp_type = (int) value ? USVMPC_SQP_RTI : USVMPC_SQP;
    else if (!strcmp(field, "cold_start")) s->P.cold_start = value != 0.0;
    else if (!strcmp(field, "print_level")) {}
    else if (!strcmp(field, "lpt_schedule")) s->lpt = value != 0.0;
    else if (!strcmp(field, "riccati_precision"))
    {
        if (value != 32.0 && value != 64.0) return fail(USVMPC_E_INVALID, "riccati_precision must be 32 or 64");
        s->P.chain_fp32 = value == 32.0;
    }
    else if (!strcmp(field, "rti_phase"))
    {
        // 0: preparation + feedback, 1: preparation, 2: feedback (ocp_nlp_sqp_rti.c:459-488)
        if (value != 0.0 && value != 1.0 && value != 2.0) return fail(USVMPC_E_INVALID, "rti_phase must be 0, 1 or 2");
        if (value != 0.0 && !s->d_prep)
        {
            const size_t n = (size_t) s->B * c.N * (s->nv * s->nx + s->nx);
            CU(cudaSetDevice(s->device));
            CU(cudaMalloc(&s->d_prep, sizeof(double) * n));
            CU(cudaMemset(s->d_prep, 0, sizeof(double) * n));
        }
        s->P.rti_phase = (int) value;
    }
    else if (!strcmp(field, "step_length")) { if (value != 1.0) return fail(USVMPC_E_INVALID, "step_length %g: the engine takes full SQP steps like the reference scripts", value); }
    else return fail(USVMPC_E_FIELD, "unknown option '%s'", field);
    refresh_params(s);
    return 0;
}

int usvmpc_obstacle_frontend(const double* pose, const double* obs_body, const int* len, int batch, int max_obs, int K,
                             double boat_radius, double init_obs_pos, double* p_out, double* r_out, void* stream)
{
    if (!pose || !obs_body || !len || !p_out || !r_out) return fail(USVMPC_E_INVALID, "null argument");
    if (batch < 1 || K < 1 || max_obs < 1 || max_obs > 64) return fail(USVMPC_E_INVALID, "need batch >= 1, K >= 1, 1 <= max_obs <= 64");
    // launch on the device that owns the buffers (the caller may have another device current)
    cudaPointerAttributes attr;
    CU(cudaPointerGetAttributes(&attr, pose));
    if (attr.type != cudaMemoryTypeDevice && attr.type != cudaMemoryTypeManaged)
        return fail(USVMPC_E_INVALID, "usvmpc_obstacle_frontend: pose must be a device pointer");
    int prev = 0;
    CU(cudaGetDevice(&prev));
    if (prev != attr.device) CU(cudaSetDevice(attr.device));
    obstacle_frontend_kernel<<<(batch + 127) / 128, 128, 0, (cudaStream_t) stream>>>(pose, obs_body, len, batch, max_obs, K, boat_radius,
                                                                                     init_obs_pos, p_out, r_out);
    const cudaError_t err = cudaGetLastError();
    if (prev != attr.device) cudaSetDevice(prev);
    CU(err);
    return 0;
}

int usvmpc_qp_solve(usvmpc_solver* s, const double* G, const double* b, const double* rq, const double* gxy, const double* d,
                    double* ux, double* pi, double* lam, double* t, void* stream)
{
    if (!s || !G || !b || !rq || !d || !ux || !pi || !lam || !t) return fail(USVMPC_E_INVALID, "null argument");
    if (s->cfg.K > 0 && !gxy) return fail(USVMPC_E_INVALID, "null argument");
    CU(cudaSetDevice(s->device));
    const double* g2 = gxy ? gxy : d;
    s->P.qp = QpIo{G, b, rq, g2, d, ux, pi, lam, t};
    const int have = s->have_history;
    s->have_history = 0;   // the iteration counts of an NMPC solve say nothing about these QPs: plain queue order
    if (s->cfg.nsh > 0) return fail(USVMPC_E_INVALID, "usvmpc_qp_solve: soft rows are not part of the QP seam's layout");
    const int rc = launch_model(s, (cudaStream_t) stream);
    s->P.qp = QpIo{};
    (void) have;
    s->have_history = 0;   // the statistics record now holds QP statistics: no history for the next NMPC solve
    if (rc) return rc;
    s->launches++;
    return 0;
}

int usvmpc_set_result_buffer(usvmpc_solver* s, double* device_buffer)
{
    if (!s) return fail(USVMPC_E_INVALID, "null solver");
    s->P.packed = device_buffer;
    s->P.packed_width = (s->cfg.N + 1) * s->nx + s->cfg.N * s->nu + 7;
    return s->P.packed_width;
}

int usvmpc_info(usvmpc_solver* s, const char* what, double* value)
{
    if (!s || !what || !value) return fail(USVMPC_E_INVALID, "null argument");
    if (!strcmp(what, "launches")) *value = (double) s->launches;
    else if (!strcmp(what, "workspace_bytes")) *value = (double) sizeof(double) * s->P.ws_stride * s->B;
    else if (!strcmp(what, "workspace_doubles_per_instance")) *value = (double) s->P.ws_stride;
    else if (!strcmp(what, "batch")) *value = s->B;
    else if (!strcmp(what, "threads_per_cta")) *value = THREADS;
    else if (!strcmp(what, "ctas_per_sm")) *value = s->ctas_per_sm;
    else if (!strcmp(what, "smem_bytes_per_cta")) *value = (double) sizeof(double) * s->P.plan.smem_doubles;
    else if (!strcmp(what, "scratch_bytes_per_cta")) *value = (double) sizeof(double) * s->P.plan.scratch_doubles;
    else if (!strcmp(what, "grid")) *value = s->grid;
    else if (!strcmp(what, "sm_clock_hz")) *value = s->sm_clock_hz;
    else return fail(USVMPC_E_FIELD, "unknown info '%s'", what);
    return 0;
}

}  // extern "C"

// ------------------------------------------------------------------------------------------------------------------
// The exact symbols of the reference's generated solver + acados_c calls for B = 1 (include/acados_compat.h)
// ------------------------------------------------------------------------------------------------------------------
#include "../../include/acados_compat.h"

namespace {
usvmpc_solver* g_acados = nullptr;     // the one solver of the process, like the generated library's globals
usvmpc_config g_acados_cfg;
bool g_acados_cfg_set = false;
double g_acados_stats[NSTAT];
double g_acados_cost = 0.0;
}  // namespace

extern "C" {

int usvmpc_config_guidance_ca1(usvmpc_config* c)
{
    int rc = usvmpc_config_default(c, USVMPC_MODEL_GUIDANCE_CA1);
    if (rc) return rc;
    const int nx = 8, nu = 1, ny = nx + nu, K = 8;
    c->N = 100; c->K = K; c->dt = 5.0 / 100; c->num_steps = 1; c->num_stages = 4; c->nlp_type = USVMPC_SQP_RTI;
    memset(c->W, 0, sizeof(c->W)); memset(c->W_e, 0, sizeof(c->W_e));
    c->W[2 + ny * 2] = 0.05; c->W[3 + ny * 3] = 0.01; c->W[8 + ny * 8] = 0.2;      // Q = diag(0,0,.05,.01,0,0,0,0), R = .2
    c->W_e[2 + nx * 2] = 0.1; c->W_e[3 + nx * 3] = 0.05;
    c->nbu = 1; c->lbu[0] = -0.5; c->ubu[0] = 0.5; c->nbx = 0;
    c->uh = 1000000.0;
    c->nsh = K;
    for (int i = 0; i < K; i++) { c->lsh[i] = -0.2; c->ush[i] = 0.0; c->zl[i] = 1.0; c->zu[i] = 1.0; c->Zl[i] = 0.0; c->Zu[i] = 0.0; }
    return 0;
}

int usvmpc_acados_configure(const usvmpc_config* cfg)
{
    g_acados_cfg_set = cfg != nullptr;
    if (cfg) g_acados_cfg = *cfg;
    return 0;
}

int acados_create(void)
{
    if (g_acados) return 0;
    if (!g_acados_cfg_set) { int rc = usvmpc_config_guidance_ca1(&g_acados_cfg); if (rc) return 1; }
    int dev = 0;
    cudaGetDevice(&dev);
    if (usvmpc_create(&g_acados_cfg, 1, dev, &g_acados) != 0) { fprintf(stderr, "acados_create: %s\n", g_err); return 1; }
    const usvmpc_config& c = g_acados_cfg;
    // initial values of the generated code (acados_solver.in.c:1028-1449): lh from the description would be baked in; the node
    // sets lh, p, yref, x0 before every solve, so zeros are as good
    (void) c;
    return 0;
}

int acados_free(void)
{
    usvmpc_free(g_acados);
    g_acados = nullptr;
    return 0;
}

int acados_update_params(int stage, double* value, int np)
{
    if (!g_acados) return 1;
    if (usvmpc_update_params(g_acados, stage, value, np, 0, nullptr) != 0)
    {
        // the generated code exits on a size mismatch (acados_solver.in.c:1746-1751)
        fprintf(stderr, "acados_update_params: %s\n", g_err);
        exit(1);
    }
    return 0;
}

int acados_solve(void)
{
    if (!g_acados) return 1;
    if (usvmpc_solve(g_acados, nullptr) != 0 || usvmpc_get_stats(g_acados, g_acados_stats, 0, nullptr) != 0)
    {
        fprintf(stderr, "acados_solve: %s\n", g_err);
        return 1;   // ACADOS_FAILURE
    }
    return (int) g_acados_stats[0];
}

void acados_print_stats(void)
{
    printf("iter\tqp_iter\tres_stat\tres_eq\t\tres_ineq\tres_comp\n%d\t%d\t%e\t%e\t%e\t%e\n", (int) g_acados_stats[1],
           (int) g_acados_stats[2], g_acados_stats[3], g_acados_stats[4], g_acados_stats[5], g_acados_stats[6]);
}

void* acados_get_nlp_in(void) { return g_acados; }
void* acados_get_nlp_out(void) { return g_acados; }
void* acados_get_nlp_solver(void) { return g_acados; }
void* acados_get_nlp_config(void) { return g_acados; }
void* acados_get_nlp_opts(void) { return g_acados; }
void* acados_get_nlp_dims(void) { return g_acados; }
void* acados_get_nlp_plan(void) { return g_acados; }

static void acados_die(const char* what)
{
    // the reference's C side exits on an unknown field
    fprintf(stderr, "%s: %s\n", what, g_err);
    exit(1);
}

int ocp_nlp_cost_model_set(void* config, void* dims, void* in, int stage, const char* field, void* value)
{
    (void) config; (void) dims;
    if (usvmpc_cost_model_set((usvmpc_solver*) in, stage, field, (const double*) value, 0, nullptr) != 0) acados_die("ocp_nlp_cost_model_set");
    return 0;   // ACADOS_SUCCESS
}

int ocp_nlp_constraints_model_set(void* config, void* dims, void* in, int stage, const char* field, void* value)
{
    (void) config; (void) dims;
    if (usvmpc_constraints_model_set((usvmpc_solver*) in, stage, field, (const double*) value, 0, nullptr) != 0)
        acados_die("ocp_nlp_constraints_model_set");
    return 0;
}

void ocp_nlp_out_set(void* config, void* dims, void* out, int stage, const char* field, void* value)
{
    (void) config; (void) dims;
    if (usvmpc_out_set((usvmpc_solver*) out, stage, field, (const double*) value, 0, nullptr) != 0) acados_die("ocp_nlp_out_set");
}

void ocp_nlp_out_get(void* config, void* dims, void* out, int stage, const char* field, void* value)
{
    (void) config; (void) dims;
    if (usvmpc_out_get((usvmpc_solver*) out, stage, field, (double*) value, 0, nullptr) != 0) acados_die("ocp_nlp_out_get");
}

void ocp_nlp_get_at_stage(void* config, void* dims, void* solver, int stage, const char* field, void* value)
{
    (void) config; (void) dims;
    // the slack values are the fields the wrapper reads this way (acados_ocp_solver.py:774-782); the dynamics matrices
    // ("A", "B") are not kept per instance after a solve
    if (strcmp(field, "sl") && strcmp(field, "su")) { fail(USVMPC_E_FIELD, "unknown field '%s'", field); acados_die("ocp_nlp_get_at_stage"); }
    if (usvmpc_out_get((usvmpc_solver*) solver, stage, field, (double*) value, 0, nullptr) != 0) acados_die("ocp_nlp_get_at_stage");
}

int ocp_nlp_dims_get_from_attr(void* config, void* dims, void* out, int stage, const char* field)
{
    (void) config; (void) dims;
    const int d = usvmpc_dims_get_from_attr((usvmpc_solver*) out, stage, field);
    if (d < 0) acados_die("ocp_nlp_dims_get_from_attr");
    return d;
}

void ocp_nlp_get(void* config, void* solver, const char* field, void* return_value)
{
    (void) config;
    usvmpc_solver* s = (usvmpc_solver*) solver;
    const double* st = g_acados_stats;
    if (!strcmp(field, "sqp_iter")) *(int*) return_value = (int) st[1];
    else if (!strcmp(field, "qp_iter")) *(int*) return_value = (int) st[2];
    else if (!strcmp(field, "status")) *(int*) return_value = (int) st[0];
    else if (!strcmp(field, "res_stat")) *(double*) return_value = st[3];
    else if (!strcmp(field, "res_eq")) *(double*) return_value = st[4];
    else if (!strcmp(field, "res_ineq")) *(double*) return_value = st[5];
    else if (!strcmp(field, "res_comp")) *(double*) return_value = st[6];
    else if (!strcmp(field, "cost_value")) *(double*) return_value = g_acados_cost;
    else if (!strcmp(field, "time_tot")) *(double*) return_value = st[12] / s->sm_clock_hz;
    else if (!strcmp(field, "time_lin")) *(double*) return_value = st[13] / s->sm_clock_hz;
    else if (!strcmp(field, "time_qp") || !strcmp(field, "time_qp_sol")) *(double*) return_value = st[14] / s->sm_clock_hz;
    else { fail(USVMPC_E_FIELD, "unknown field '%s'", field); acados_die("ocp_nlp_get"); }
}

void ocp_nlp_solver_opts_set(void* config, void* opts, const char* field, void* value)
{
    (void) config;
    // integer-valued options arrive as int*, tolerances as double* (ocp_nlp_sqp_rti.c:168-242)
    const bool is_int = !strcmp(field, "rti_phase") || !strcmp(field, "max_iter") || !strcmp(field, "qp_iter_max") || !strcmp(field, "print_level");
    const double v = is_int ? (double) *(int*) value : *(double*) value;
    if (usvmpc_solver_opts_set((usvmpc_solver*) opts, field, v) != 0) acados_die("ocp_nlp_solver_opts_set");
}

void ocp_nlp_eval_residuals(void* solver, void* in, void* out) { (void) solver; (void) in; (void) out; }

void ocp_nlp_eval_cost(void* solver, void* in, void* out)
{
    (void) in; (void) out;
    if (usvmpc_eval_cost((usvmpc_solver*) solver, &g_acados_cost, 0, nullptr) != 0) acados_die("ocp_nlp_eval_cost");
}

// the 2-D size queries of the Python wrapper's cost_set / constraints_set (acados_ocp_solver.py:1022-1030, 1090-1098):
// vectors report (n, 0), the weight matrix (ny, ny)
static void dims2(void* out, int stage, const char* field, int* dims_out, const char* who)
{
    usvmpc_solver* s = (usvmpc_solver*) out;
    if (!strcmp(field, "W"))
    {
        const int ny = usvmpc_dims_get_from_attr(s, stage, "yref");
        if (ny < 0) acados_die(who);
        dims_out[0] = ny; dims_out[1] = ny;
        return;
    }
    const int d = usvmpc_dims_get_from_attr(s, stage, field);
    if (d < 0) acados_die(who);
    dims_out[0] = d; dims_out[1] = 0;
}

void ocp_nlp_cost_dims_get_from_attr(void* config, void* dims, void* out, int stage, const char* field, int* dims_out)
{
    (void) config; (void) dims;
    dims2(out, stage, field, dims_out, "ocp_nlp_cost_dims_get_from_attr");
}

void ocp_nlp_constraint_dims_get_from_attr(void* config, void* dims, void* out, int stage, const char* field, int* dims_out)
{
    (void) config; (void) dims;
    dims2(out, stage, field, dims_out, "ocp_nlp_constraint_dims_get_from_attr");
}

}  // extern "C"
