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
DEV unsigned warp_max_u32(unsigned v) { return __reduce_max_sync(0xffffffffu, v); }   // one REDUX instruction
DEV unsigned long long double_bits(double v) { return (unsigned long long) __double_as_longlong(v); }
DEV double bits_double(unsigned long long u) { return __longlong_as_double((long long) u); }
// one fp64 tensor-core product of the warp: D (8 x 8) += A (8 x 4, row-major fragments) * B (4 x 8).  With g = lane / 4,
// t = lane % 4 the lane holds A[g][t], B[t][g] and D[g][2t], D[g][2t + 1]   (mma.sync.m8n8k4.f64, SASS DMMA)
DEV void dmma884(double& d0, double& d1, double a, double b)
{
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}
// shared-memory accesses by 32-bit shared-window byte address (no address arithmetic left to the compiler)
typedef unsigned saddr;
DEV saddr smem_addr(const void* p) { return (saddr) __cvta_generic_to_shared(p); }
DEV double lds64(saddr a) { double v; asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(a)); return v; }
DEV void lds128(saddr a, double& x, double& y) { asm volatile("ld.shared.v2.f64 {%0,%1}, [%2];" : "=d"(x), "=d"(y) : "r"(a)); }
DEV void sts64(saddr a, double v) { asm volatile("st.shared.f64 [%0], %1;" ::"r"(a), "d"(v) : "memory"); }
DEV void syncwarp() { __syncwarp(); }
DEV void syncthreads() { __syncthreads(); }
// barrier `id` (1..15) among the first `count` threads of the block (count a multiple of 32)
DEV void named_barrier_sync(int id, int count) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(count) : "memory"); }
DEV double dsqrt(double a) { return sqrt(a); }
DEV double drsqrt(double a) { return rsqrt(a); }
DEV float drsqrt(float a) { return rsqrtf(a); }
// 1 / sqrt(a) for a positive, normal a: the hardware seed (MUFU.RSQ64H, ~2^-21) and two Newton steps, ~10 instructions
// against ~28 of rsqrt(double) with its special-case handling; within an ulp or two of it
DEV double drsqrt_pos(double a)
{
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(a));
#pragma unroll
    for (int it = 0; it < 2; it++)
    {
        const double e = fma(-(a * y), y, 1.0);
        y = fma(0.5 * y, e, y);
    }
    return y;
}
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
    double slot_d[MAXT], slot_e[MAXT];
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

typedef uintptr_t saddr;
DEV saddr smem_addr(const void* p) { return (saddr) p; }
DEV double lds64(saddr a) { return *(const double*) a; }
DEV void lds128(saddr a, double& x, double& y) { x = ((const double*) a)[0]; y = ((const double*) a)[1]; }
DEV void sts64(saddr a, double v) { *(double*) a = v; }
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
DEV void dmma884(double& d0, double& d1, double a, double b)
{
    emu::Block* blk = emu::g_blk;
    const int base = blk->cur & ~31, g = (blk->cur & 31) >> 2, t = blk->cur & 3;
    blk->slot_d[blk->cur] = a; blk->slot_e[blk->cur] = b;
    syncwarp();
    for (int k = 0; k < 4; k++)
    {
        const double ak = blk->slot_d[base + 4 * g + k];
        d0 = std::fma(ak, blk->slot_e[base + 4 * (2 * t) + k], d0);
        d1 = std::fma(ak, blk->slot_e[base + 4 * (2 * t + 1) + k], d1);
    }
    syncwarp();
}
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
DEV unsigned warp_max_u32(unsigned v)
{
    emu::Block* b = emu::g_blk;
    b->slot_i[b->cur] = (long) v;
    syncwarp();
    unsigned r = 0;
    const int base = b->cur & ~31;
    for (int l = 0; l < 32 && base + l < b->T; l++) { const unsigned o = (unsigned) b->slot_i[base + l]; r = o > r ? o : r; }
    syncwarp();
    return r;
}
DEV unsigned long long double_bits(double v) { unsigned long long u; std::memcpy(&u, &v, 8); return u; }
DEV double bits_double(unsigned long long u) { double v; std::memcpy(&v, &u, 8); return v; }
DEV double dsqrt(double a) { return std::sqrt(a); }
DEV double drsqrt(double a) { return 1.0 / std::sqrt(a); }
DEV float drsqrt(float a) { return 1.0f / std::sqrt(a); }
DEV double drsqrt_pos(double a) { return 1.0 / std::sqrt(a); }
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
// reductions over one warp; every lane ends up with the result
// max: the doubles are mapped to unsigned integers of the same order (sign bit flipped for v >= 0, all bits for v < 0),
// then two 32-bit warp reductions (high word, low word among the lanes that hold the maximal high word) replace a
// five-step shuffle butterfly.  The value returned is bit for bit one of the inputs.  Callers never pass NaN (their
// running maxima are built with `q > m ? q : m`, which drops NaN).
DEV double warp_max(double v)
{
    unsigned long long u = double_bits(v);
    u ^= (u >> 63) ? ~0ull : 0x8000000000000000ull;
    const unsigned hi = (unsigned) (u >> 32), lo = (unsigned) u;
    const unsigned mh = warp_max_u32(hi);
    const unsigned ml = warp_max_u32(hi == mh ? lo : 0u);
    u = ((unsigned long long) mh << 32) | ml;
    u ^= (u >> 63) ? 0x8000000000000000ull : ~0ull;
    return bits_double(u);
}
// the same for values that are known to be >= 0 (not -0, not NaN): their bit patterns already order like the values
DEV double warp_max_nonneg(double v)
{
    const unsigned long long u = double_bits(v);
    const unsigned hi = (unsigned) (u >> 32), lo = (unsigned) u;
    const unsigned mh = warp_max_u32(hi);
    const unsigned ml = warp_max_u32(hi == mh ? lo : 0u);
    return bits_double(((unsigned long long) mh << 32) | ml);
}
DEV double warp_sum(double v)
{
#pragma unroll
    for (int m = 16; m > 0; m >>= 1) v += shfl_xor(v, m);
    return v;
}
}  // namespace usvmpc
