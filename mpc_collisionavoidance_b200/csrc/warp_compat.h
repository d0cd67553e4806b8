// warp_compat.h -- one source for the warp-per-instance NMPC kernel, two ways to build it.
//
//  * nvcc (the product): the macros below map 1:1 onto CUDA warp intrinsics.
//  * g++ -DUSVMPC_EMULATE (tests/emu only): 32 cooperative fibers stand in for the 32 lanes of ONE
//    warp, switching at every shuffle / __syncwarp, so the very same device functions can be run,
//    address-sanitised and compared with the oracle on a machine without a GPU.  The emulation is
//    test infrastructure; nothing in the product path is built with USVMPC_EMULATE.
#pragma once

#ifndef USVMPC_EMULATE
// ------------------------------------------------------------------ CUDA
#include <cuda_runtime.h>
// compiler-level memory barrier: the loads written before it are issued before anything written after it
#define CBAR() asm volatile("" ::: "memory")
#define DEV __device__ __forceinline__
#define MDEV __device__ __forceinline__
#define MDEVNI __device__ __noinline__
#define DEVNI __device__ __noinline__
#define HD __host__ __device__ __forceinline__

namespace usvmpc {
DEV int lane_id() { return threadIdx.x & 31; }
DEV double shfl(double v, int src) { return __shfl_sync(0xffffffffu, v, src); }
DEV int shfl(int v, int src) { return __shfl_sync(0xffffffffu, v, src); }
DEV double shfl_xor(double v, int m) { return __shfl_xor_sync(0xffffffffu, v, m); }
DEV int shfl_xor(int v, int m) { return __shfl_xor_sync(0xffffffffu, v, m); }
DEV void syncwarp() { __syncwarp(); }
DEV double dsqrt(double a) { return sqrt(a); }
DEV double drsqrt(double a) { return rsqrt(a); }
DEV double dabs(double a) { return fabs(a); }
DEV void dsincos(double a, double* s, double* c) { sincos(a, s, c); }
DEV bool disnan(double a) { return isnan(a); }
// asynchronous 16-byte global -> shared copy (LDGSTS, L2-only caching: the records are streamed, not reused in L1)
DEV void cp_async16(double* smem_dst, const double* gsrc)
{
    const unsigned sa = (unsigned) __cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(sa), "l"(gsrc) : "memory");
}
DEV void st2(double* dst, const double* src) { *reinterpret_cast<double2*>(dst) = *reinterpret_cast<const double2*>(src); }
// ---- TMA bulk copies (cp.async.bulk, SASS UBLKCP) with mbarrier completion: one instruction moves a whole record
DEV unsigned smem_u32(const void* p) { return (unsigned) __cvta_generic_to_shared(p); }
DEV void mbar_init(unsigned long long* bar)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;\n" ::"r"(smem_u32(bar)) : "memory");
}
DEV void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory"); }
DEV void fence_proxy_async() { asm volatile("fence.proxy.async;\n" ::: "memory"); }
DEV void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory"); }
// global -> shared, completion counted in bytes on `bar` (issued by ONE lane)
DEV void bulk_g2s(double* dst, const double* src, int bytes, unsigned long long* bar)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
// every lane waits for the phase to complete; a bounded spin turns a protocol bug into a trap instead of a hang
DEV void mbar_wait(unsigned long long* bar, unsigned phase)
{
    const unsigned a = smem_u32(bar);
    unsigned done = 0;
    for (int it = 0; it < (1 << 24) && !done; it++)
        asm volatile("{\n.reg .pred P1;\nmbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2;\nselp.b32 %0, 1, 0, P1;\n}\n"
                     : "=r"(done)
                     : "r"(a), "r"(phase)
                     : "memory");
    if (!done) __trap();
}
// shared -> global (issued by ONE lane), grouped; wait_read: the shared source may be overwritten again
DEV void bulk_s2g(double* dst, const double* src, int bytes)
{
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;\n" ::"l"(dst), "r"(smem_u32(src)), "r"(bytes) : "memory");
}
DEV void bulk_commit() { asm volatile("cp.async.bulk.commit_group;\n" ::: "memory"); }
DEV void bulk_wait_read1() { asm volatile("cp.async.bulk.wait_group.read 1;\n" ::: "memory"); }
DEV void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;\n" ::: "memory"); }
DEV void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
DEV void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;\n" ::: "memory"); }
}  // namespace usvmpc

#else
// ------------------------------------------------------------------ CPU fiber emulation of one warp
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#define CBAR() asm volatile("" ::: "memory")
#define DEV static inline
#define MDEV inline
#define MDEVNI inline
#define DEVNI static
#define HD static inline

namespace usvmpc {
namespace emu {
constexpr int WARP = 32;
extern "C" void usvmpc_fiber_switch(void** save_sp, void* new_sp);
struct Warp {
    void* sp[WARP];
    void* main_sp;
    char* stacks;
    int cur;
    double slot_d[WARP];
    long slot_i[WARP];
    void (*body)(void*);
    void* arg;
    // deferred asynchronous copies, per lane: the data lands only at cp_async_wait_all(), and the destination is
    // poisoned in between, so that reading a prefetch buffer before waiting shows up as NaNs in the CPU tests
    struct Pending { double* dst; const double* src; };
    Pending pend[WARP][512];
    int npend[WARP];
    // bulk (TMA-like) copies: loads land when their barrier is waited on, stores when their group is waited on
    struct Bulk { double* dst; const double* src; int n; const void* bar; int group; };
    Bulk bl[64];
    int nbl;
    Bulk bs[64];
    int nbs, group;
};
extern thread_local Warp* g_warp;
// hand control to the next lane; returns when every other lane has reached its own next yield
inline void yield_lane()
{
    Warp* w = g_warp;
    int me = w->cur, nx = (me + 1) % WARP;
    w->cur = nx;
    usvmpc_fiber_switch(&w->sp[me], w->sp[nx]);
}
void run_warp(void (*body)(void*), void* arg);  // defined in tests/emu/emu_runtime.cpp
}  // namespace emu

DEV int lane_id() { return emu::g_warp->cur; }
DEV double shfl(double v, int src)
{
    emu::Warp* w = emu::g_warp;
    w->slot_d[w->cur] = v;
    emu::yield_lane();
    double r = w->slot_d[src & 31];
    emu::yield_lane();
    return r;
}
DEV int shfl(int v, int src)
{
    emu::Warp* w = emu::g_warp;
    w->slot_i[w->cur] = v;
    emu::yield_lane();
    int r = (int) w->slot_i[src & 31];
    emu::yield_lane();
    return r;
}
DEV double shfl_xor(double v, int m) { return shfl(v, lane_id() ^ m); }
DEV int shfl_xor(int v, int m) { return shfl(v, lane_id() ^ m); }
DEV void syncwarp() { emu::yield_lane(); }
DEV double dsqrt(double a) { return std::sqrt(a); }
DEV double drsqrt(double a) { return 1.0 / std::sqrt(a); }
DEV double dabs(double a) { return std::fabs(a); }
DEV void dsincos(double a, double* s, double* c) { *s = std::sin(a); *c = std::cos(a); }
DEV bool disnan(double a) { return std::isnan(a); }
DEV void cp_async16(double* smem_dst, const double* gsrc)
{
    emu::Warp* w = emu::g_warp;
    const int l = w->cur;
    if (w->npend[l] >= 512) abort();
    w->pend[l][w->npend[l]++] = {smem_dst, gsrc};
    smem_dst[0] = smem_dst[1] = std::nan("");
}
DEV void st2(double* dst, const double* src) { dst[0] = src[0]; dst[1] = src[1]; }
DEV void mbar_init(unsigned long long* bar) { *bar = 0; }
DEV void fence_mbar_init() {}
DEV void fence_proxy_async() {}
DEV void fence_proxy_async_smem() {}
DEV void bulk_g2s(double* dst, const double* src, int bytes, unsigned long long* bar)
{
    emu::Warp* w = emu::g_warp;
    if (w->nbl >= 64 || (bytes & 15) || ((uintptr_t) dst & 15) || ((uintptr_t) src & 15)) abort();
    w->bl[w->nbl++] = {dst, src, bytes / 8, bar, 0};
    for (int i = 0; i < bytes / 8; i++) dst[i] = std::nan("");  // not there yet
}
DEV void mbar_wait(unsigned long long* bar, unsigned phase)
{
    emu::yield_lane();  // lane 0 (which issues the copies) always runs first after a yield
    emu::Warp* w = emu::g_warp;
    int found = 0, j = 0;
    for (int i = 0; i < w->nbl; i++)
        if (w->bl[i].bar == bar) { memcpy(w->bl[i].dst, w->bl[i].src, sizeof(double) * w->bl[i].n); found = 1; }
        else w->bl[j++] = w->bl[i];
    w->nbl = j;
    if (found) { if ((unsigned) (*bar & 1) != phase) abort(); *bar ^= 1; }   // phase bookkeeping must match the kernel's
    else if ((unsigned) (*bar & 1) == phase) abort();                          // nothing in flight: the device would hang
    emu::yield_lane();
}
DEV void bulk_s2g(double* dst, const double* src, int bytes)
{
    emu::Warp* w = emu::g_warp;
    if (w->nbs >= 64 || (bytes & 15) || ((uintptr_t) dst & 15) || ((uintptr_t) src & 15)) abort();
    w->bs[w->nbs++] = {dst, src, bytes / 8, nullptr, w->group};
}
DEV void bulk_commit() { emu::g_warp->group++; }
DEV void bulk_wait_keep(int keep)
{
    emu::Warp* w = emu::g_warp;
    int j = 0;
    for (int i = 0; i < w->nbs; i++)
        if (w->bs[i].group < w->group - keep) memcpy(w->bs[i].dst, w->bs[i].src, sizeof(double) * w->bs[i].n);
        else w->bs[j++] = w->bs[i];
    w->nbs = j;
}
DEV void bulk_wait_read1() { bulk_wait_keep(1); }
DEV void bulk_wait_all() { bulk_wait_keep(0); }
DEV void cp_async_commit() {}
DEV void cp_async_wait_all()
{
    emu::Warp* w = emu::g_warp;
    const int l = w->cur;
    for (int i = 0; i < w->npend[l]; i++) { w->pend[l][i].dst[0] = w->pend[l][i].src[0]; w->pend[l][i].dst[1] = w->pend[l][i].src[1]; }
    w->npend[l] = 0;
}
}  // namespace usvmpc
#endif

namespace usvmpc {
// butterfly reductions; every lane ends up with the result
DEV double warp_max(double v)
{
#pragma unroll
    for (int m = 16; m > 0; m >>= 1) { double o = shfl_xor(v, m); v = o > v ? o : v; }
    return v;
}
DEV double warp_sum(double v)
{
#pragma unroll
    for (int m = 16; m > 0; m >>= 1) v += shfl_xor(v, m);
    return v;
}
DEV int warp_or(int v)
{
#pragma unroll
    for (int m = 16; m > 0; m >>= 1) v |= shfl_xor(v, m);
    return v;
}
}  // namespace usvmpc
