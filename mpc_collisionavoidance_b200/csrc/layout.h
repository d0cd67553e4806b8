// layout.h -- kernel parameter block and the HBM layout of one NMPC instance's working set.
// Shared by the CUDA C-ABI (usvmpc_api.cu), the kernel (nmpc_kernel.cuh) and the CPU warp emulator
// used by the tests.
//
// Every instance owns one contiguous block of `ws_stride` doubles.  Inside it each field is an array
// [stage][dim] (stage-major) so that the warp that owns the instance streams whole stage rows with
// unit stride; `Field{off, stride}` gives the offset of stage 0 and the stage pitch, both in doubles
// and both multiples of 2 (16 B).
//
// Row layouts of the inequality vectors (one "side" = lower or upper; the upper side follows the
// lower side at +ncq / +ncz):
//   IPM / QP level  (ncq = nbu + nbx + K): [ u boxes | x boxes (idxbx) | h rows ]
//   NLP level       (ncz = nbu + NX  + K): [ u boxes | x boxes: NX slots (stage 0 holds the x0
//                                            embedding, stages 1..N-1 use the first nbx) | h rows ]
#pragma once

namespace usvmpc {

constexpr int KMAX = 32;     // max obstacles per stage
constexpr int NSPC = 128;    // stage pitch of the transposed arrays (compile-time so that element offsets fold); N <= 127
constexpr int NBXMAX = 8;
constexpr int NBUMAX = 4;
constexpr int NSTAT = 12;    // per-instance statistics record (doubles)
// stats: 0 status, 1 sqp_iter, 2 qp_iter (total), 3..6 res_stat/eq/ineq/comp, 7 reserved,
//        8 solve-only Riccati sweeps, 9 last QP status, 10 last QP iterations, 11 reserved

struct Field { int off, stride, es; };  // offset of (stage 0, element 0), stage stride, element stride (doubles)

struct Layout {
    // NLP iterate (persists between solves = warm start, like nlp_out in the reference); stage-major arrays
    Field zux, zpi, zlam, zt, zfun;
    // ---- the per-stage RECORD HEAD: what the serial Riccati (chain) sweeps stream through shared memory with TMA,
    //      one contiguous block per stage (stride rec_size, compile-time offsets in the kernel):
    //        BAt | rb | L | Pb | bv
    Field BAt, rb, L, Pb, bv;
    int rec_off, rec_size;
    // ---- everything only the stage-parallel passes touch (one lane per stage) is stored TRANSPOSED,
    //      [element][stage] with the stage index fastest, so that the 32 lanes of a pass read consecutive addresses
    Field BAtT, ux, pi, rg, dux, dpi, rq, b, gxy, lam, t, rd, ti, rmc, dlam, dt, d;
    int nsp;  // padded number of stages of the transposed arrays
    // iterative refinement scratch (rare path), stage-major arrays
    Field dux2, dpi2, dlam2, dt2, rg2, rb2, rd2, rm2;
    long total;
};

struct Params {
    int B, N, K, num_steps, num_stages, nlp_type, max_iter, qp_iter_max, nbx, nbu;
    int idxbx[NBXMAX];
    int p_per_stage, lh_per_stage, yref_per_stage, cold_start;
    int ncq, ncz;
    double dt, tol[4], uh;
    double lbu[NBUMAX], ubu[NBUMAX], lbx[NBXMAX], ubx[NBXMAX];
    const double* cst;      // [W (NY*NY col-major) | W_e (NX*NX col-major)]
    const double* x0;       // [B][NX]
    const double* p;        // [B][N+1][2K] or [B][2K]
    const double* lh;       // [B][N][K]   or [B][K]
    const double* yref;     // [B][N][NY]  or [B][NY]
    const double* yref_e;   // [B][NX]
    double* ws;             // [B][ws_stride]
    long ws_stride;
    double* stats;          // [B][NSTAT]
    Layout lay;
};

inline int round_up(int a, int m) { return (a + m - 1) / m * m; }

// Compute the per-instance layout.  nx/nu are the model dimensions.
inline Layout make_layout(int nx, int nu, int N, int K, int nbx, int nbu)
{
    Layout L;
    const int N1 = N + 1, nv = nx + nu;
    const int sv = round_up(nv, 2), sx = round_up(nx, 2);
    const int ncq = nbu + nbx + K, ncz = nbu + nx + K;
    const int scq = round_up(2 * ncq, 2) > 0 ? round_up(2 * ncq, 2) : 2, scz = round_up(2 * ncz, 2);
    const int sBA = round_up(nv * nx, 2), sg = round_up(2 * K, 2) > 0 ? round_up(2 * K, 2) : 2;
    const int sL = round_up((nv + 1) * nv, 2);
    long o = 0;
    auto put = [&](Field& f, int stride) { f.off = (int) o; f.stride = stride; f.es = 1; o += (long) stride * N1; };
    put(L.zux, sv); put(L.zpi, sx); put(L.zlam, scz); put(L.zt, scz); put(L.zfun, scz);
    int r = 0;
    auto rec = [&](Field& f, int n) { f.off = r; r += n; };
    rec(L.BAt, sBA); rec(L.rb, sx); rec(L.L, sL); rec(L.Pb, sx); rec(L.bv, sv);
    L.rec_size = round_up(r, 2);
    L.rec_off = (int) o;
    Field* fs[] = {&L.BAt, &L.rb, &L.L, &L.Pb, &L.bv};
    for (Field* f : fs) { f->off += L.rec_off; f->stride = L.rec_size; f->es = 1; }
    o += (long) L.rec_size * N1;
    L.nsp = NSPC;
    auto tr = [&](Field& f, int dim) { f.off = (int) o; f.stride = 1; f.es = L.nsp; o += (long) dim * L.nsp; };
    tr(L.BAtT, sBA); tr(L.ux, sv); tr(L.pi, sx); tr(L.rg, sv); tr(L.dux, sv); tr(L.dpi, sx); tr(L.rq, sv); tr(L.b, sx);
    tr(L.gxy, sg); tr(L.lam, scq); tr(L.t, scq); tr(L.rd, scq); tr(L.ti, scq); tr(L.rmc, scq); tr(L.dlam, scq);
    tr(L.dt, scq); tr(L.d, scq);
    put(L.dux2, sv); put(L.dpi2, sx); put(L.dlam2, scq); put(L.dt2, scq);
    put(L.rg2, sv); put(L.rb2, sx); put(L.rd2, scq); put(L.rm2, scq);
    L.total = (o + 15) / 16 * 16;  // 128-byte multiple
    return L;
}

// per-warp shared-memory (doubles) used by nmpc_kernel.cuh: two record buffers + scratch
inline int warp_smem_doubles(int nx, int nu, int N, int K, int nbx, int nbu)
{
    const int nv = nx + nu, ncq2 = 2 * (nu + nx + K);
    const Layout L = make_layout(nx, nu, N, K, nbx, nbu);
    const int ne = nv * (nv + 1) / 2 + nv, nq = nu + nx + K;
    int n = 3 * L.rec_size + 4;  // record-head buffers (the rare path's scratch is aliased onto them) + mbarriers
    {
        const int rare = nv * nx + nx * nx + nx + 4 * nq + (nv + 1) * nv + nv + 2 * nx + 2 * (K > 0 ? K : 1);
        if (n < rare + 4) n = rare + 4;
    }
    n += 2 * nv * nv;          // Hs, Hes
    n += nv * nv + nx * nx;    // Ws, Wes
    n += 3 * ne;               // Tp
    n += (nv + 1) * nx;        // sAL
    n += (ne + nq + nv + nx + 1) / 2 + 16;  // int tables + alignment slack
    (void) ncq2;
    return (n + 1) / 2 * 2;
}

}  // namespace usvmpc
