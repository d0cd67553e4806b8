// cta_compat.h -- one source for the CTA-per-instance NMPC kernel, two ways to build it.
//
//  * nvcc (the product): the functions below map 1:1 onto CUDA thread / warp / block intrinsics.
//  * g++ -DUSVMPC_EMULATE (tests/emu only): the threads of ONE thread block become cooperative fibers that switch
//    at every shuffle / __syncwarp / __syncthreads (real barriers: a fiber that arrives early is parked until its
//    warp / block has arrived), so the very same device functions can be run, address-sanitised and compared with
//    the oracle on a machine without a GPU.  The emulation is test infrastructure; nothing in the product path is
//    built with USVMPC_EMULATE.
#pragma once

#ifndef USVMPC_EMULATE
// ------------------------------------------------------------------ CUDA
#include <cuda_runtime.h>
#define DEV __device__ __forceinline__
#define MDEV __device__ __forceinline__
#define MDEVNI __device__ __noinline__
// address-space hint: the pointer is known to point into shared memory (lets the compiler emit LDS / STS)
#define ASSUME_SHARED(p) __builtin_assume(__isShared(p))

namespace usvmpc {
DEV int thread_id() { return threadIdx.x; }
DEV int block_threads() { return blockDim.x; }
DEV int lane_id() { return threadIdx.x & 31; }
DEV double shfl(double v, int src) { return __shfl_sync(0xffffffffu, v, src); }
DEV float shfl(float v, int src) { return __shfl_sync(0xffffffffu, v, src); }
DEV int shfl(int v, int src) { return __shfl_sync(0xffffffffu, v, src); }
DEV double shfl_xor(double v, int m) { return __shfl_xor_sync(0xffffffffu, v, m); }
DEV int shfl_xor(int v, int m) { return __shfl_xor_sync(0xffffffffu, v, m); }
DEV void syncwarp() { __syncwarp(); }
DEV void syncthreads() { __syncthreads(); }
// barrier `id` (1..15) among the first `count` threads of the block (count a multiple of 32)
DEV void named_barrier_sync(int id, int count) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(count) : "memory"); }
DEV double dsqrt(double a) { return sqrt(a); }
DEV double drsqrt(double a) { return rsqrt(a); }
DEV float drsqrt(float a) { return rsqrtf(a); }
DEV double dabs(double a) { return fabs(a); }
DEV void dsincos(double a, double* s, double* c) { sincos(a, s, c); }
DEV bool disnan(double a) { return isnan(a); }
// work queue of the persistent CTAs (device-scope)
DEV int atomic_fetch_add(int* p, int v) { return atomicAdd(p, v); }
DEV int volatile_load(const int* p) { return *(const volatile int*) p; }
DEV void volatile_store(int* p, int v) { *(volatile int*) p = v; }
DEV void threadfence() { __threadfence(); }
DEV void backoff() { __nanosleep(200); }
DEV long long clock_now() { return clock64(); }
}  // namespace usvmpc

#else
// ------------------------------------------------------------------ CPU fiber emulation of one thread block
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#define DEV static inline
#define MDEV inline
#define MDEVNI inline
#define ASSUME_SHARED(p) ((void) 0)

namespace usvmpc {
namespace emu {
constexpr int WARP = 32;
constexpr int MAXT = 1024;
struct Barrier { int count, gen, need; };
struct Fiber {
    void* sp;
    const Barrier* wait;  // barrier this fiber is parked on (nullptr: runnable)
    int wait_gen;
    bool done;
};
struct Block {
    int T, cur;
    Fiber f[MAXT];
    Barrier cta, warp[MAXT / WARP], named[16];
    double slot_d[MAXT];
    long slot_i[MAXT];
    void* main_sp;
    char* stacks;
    void (*body)(void*);
    void* arg;
};
extern thread_local Block* g_blk;
void arrive(Barrier* b);                       // tests/emu/emu_solver.cpp
void run_block(int T, void (*body)(void*), void* arg);
}  // namespace emu

DEV int thread_id() { return emu::g_blk->cur; }
DEV int block_threads() { return emu::g_blk->T; }
DEV int lane_id() { return emu::g_blk->cur & 31; }
DEV void syncwarp() { emu::Block* b = emu::g_blk; emu::arrive(&b->warp[b->cur >> 5]); }
DEV void syncthreads() { emu::arrive(&emu::g_blk->cta); }
DEV void named_barrier_sync(int id, int count)
{
    emu::Barrier* b = &emu::g_blk->named[id];
    if (b->need != count) { if (b->count != 0) abort(); b->need = count; }
    emu::arrive(b);
}
DEV double shfl(double v, int src)
{
    emu::Block* b = emu::g_blk;
    b->slot_d[b->cur] = v;
    syncwarp();
    const double r = b->slot_d[(b->cur & ~31) | (src & 31)];
    syncwarp();
    return r;
}
DEV float shfl(float v, int src) { return (float) shfl((double) v, src); }
DEV int shfl(int v, int src)
{
    emu::Block* b = emu::g_blk;
    b->slot_i[b->cur] = v;
    syncwarp();
    const int r = (int) b->slot_i[(b->cur & ~31) | (src & 31)];
    syncwarp();
    return r;
}
DEV double shfl_xor(double v, int m) { return shfl(v, lane_id() ^ m); }
DEV int shfl_xor(int v, int m) { return shfl(v, lane_id() ^ m); }
DEV double dsqrt(double a) { return std::sqrt(a); }
DEV double drsqrt(double a) { return 1.0 / std::sqrt(a); }
DEV float drsqrt(float a) { return 1.0f / std::sqrt(a); }
DEV double dabs(double a) { return std::fabs(a); }
DEV void dsincos(double a, double* s, double* c) { *s = std::sin(a); *c = std::cos(a); }
DEV bool disnan(double a) { return std::isnan(a); }
DEV int atomic_fetch_add(int* p, int v) { return __atomic_fetch_add(p, v, __ATOMIC_SEQ_CST); }
DEV int volatile_load(const int* p) { return __atomic_load_n(p, __ATOMIC_SEQ_CST); }
DEV void volatile_store(int* p, int v) { __atomic_store_n(p, v, __ATOMIC_SEQ_CST); }
DEV void threadfence() { __atomic_thread_fence(__ATOMIC_SEQ_CST); }
DEV void backoff() {}
DEV long long clock_now() { return 0; }
}  // namespace usvmpc
#endif

namespace usvmpc {
// butterfly reductions over one warp; every lane ends up with the result
DEV double warp_max(double v)
{
#pragma unroll
    for (int m = 16; m > 0; m >>= 1) { const double o = shfl_xor(v, m); v = o > v ? o : v; }
    return v;
}
DEV double warp_sum(double v)
{
#pragma unroll
    for (int m = 16; m > 0; m >>= 1) v += shfl_xor(v, m);
    return v;
}
}  // namespace usvmpc
