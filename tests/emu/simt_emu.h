// tests/emu/simt_emu.h - TEST INFRASTRUCTURE ONLY.
//
// A minimal SIMT emulator so that the *same* kernel sources that nvcc compiles for sm_100a can be
// compiled by g++ and stepped on a CPU-only development box (this container has no GPU).  Every CUDA
// thread of a block is a fiber (hand-rolled x86-64 context switch); warp collectives (__shfl_*_sync,
// __ballot_sync, __syncwarp) and __syncthreads are rendez-vous points between fibers.  Blocks run one
// after another (optionally spread over host threads).  Float arithmetic is the host's IEEE-754
// binary32/binary64 with contraction disabled, which is what the kernels get from nvcc with
// -fmad=false, so emulated results are bit-identical to the device's by construction.
//
// The product library (liblamegpu.so) is NEVER built with this header: it is only used to build
// tests/emu/liblamegpu_emu.so, which the CPU-side tests load to check the kernels' logic against the
// oracle before spending GPU time.  There is no CPU fallback in the product.
#pragma once
#include <cstdint>
#include <cstring>
#include <cstdlib>
#include <cmath>
#include <functional>

namespace emu {

struct dim3_t { unsigned x, y, z; dim3_t(unsigned a = 1, unsigned b = 1, unsigned c = 1) : x(a), y(b), z(c) {} };

struct WarpSync { uint64_t slot[2][32]; int count; int gen; };

struct Fiber { void *sp; char *stack; bool done; };

struct Block {
    int nthreads;
    Fiber *fibers;
    void *sched_sp;
    int cur;
    dim3_t bidx, bdim, gdim;
    unsigned char *smem;
    int bar_count, bar_gen;
    int nbar_count[16], nbar_gen[16];      /* named barriers (bar.sync id, count) */
    WarpSync warps[32];
    const std::function<void()> *body;
};

extern thread_local Block *g_blk;

void launch(dim3_t grid, dim3_t block, size_t smem_bytes, const std::function<void()> &body);
void yield();

inline int tid() { return g_blk->cur; }
inline dim3_t thread_idx() { return dim3_t(g_blk->cur, 0, 0); }
inline WarpSync &warp() { return g_blk->warps[g_blk->cur >> 5]; }

template <class T> inline T collective(unsigned mask, T v, int src)
{
    static_assert(sizeof(T) <= 8, "shuffle payload too large");
    WarpSync &w = warp();
    int const lane = g_blk->cur & 31;
    int const g = w.gen;
    uint64_t raw = 0;
    std::memcpy(&raw, &v, sizeof(T));
    w.slot[g & 1][lane] = raw;
    if (++w.count == __builtin_popcount(mask)) { w.count = 0; w.gen = g + 1; }
    else while (w.gen == g) yield();
    T r;
    std::memcpy(&r, &w.slot[g & 1][src & 31], sizeof(T));
    return r;
}

inline unsigned ballot(unsigned mask, int pred)
{
    WarpSync &w = warp();
    int const lane = g_blk->cur & 31;
    int const g = w.gen;
    w.slot[g & 1][lane] = pred ? 1 : 0;
    if (++w.count == __builtin_popcount(mask)) { w.count = 0; w.gen = g + 1; }
    else while (w.gen == g) yield();
    unsigned r = 0;
    for (int l = 0; l < 32; l++) if ((mask >> l) & 1) r |= (unsigned) (w.slot[g & 1][l] & 1) << l;
    return r;
}

inline void syncthreads()
{
    Block *b = g_blk;
    int const g = b->bar_gen;
    if (++b->bar_count == b->nthreads) { b->bar_count = 0; b->bar_gen = g + 1; }
    else while (b->bar_gen == g) yield();
}

inline void named_barrier(int id, int count)
{
    Block *b = g_blk;
    int const g = b->nbar_gen[id];
    if (++b->nbar_count[id] == count) { b->nbar_count[id] = 0; b->nbar_gen[id] = g + 1; }
    else while (b->nbar_gen[id] == g) yield();
}

} // namespace emu

// ---- CUDA surface used by the kernels -------------------------------------------------------------
#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __noinline__
#define __restrict__
#define __launch_bounds__(...)
#define __constant__
typedef emu::dim3_t dim3;
#define threadIdx (emu::thread_idx())
#define blockIdx (emu::g_blk->bidx)
#define blockDim (emu::g_blk->bdim)
#define gridDim (emu::g_blk->gdim)

static inline void __syncthreads() { emu::syncthreads(); }
static inline void __syncwarp(unsigned mask = 0xffffffffu) { (void) emu::collective<int>(mask, 0, 0); }
template <class T> static inline T __shfl_sync(unsigned mask, T v, int src) { return emu::collective<T>(mask, v, src); }
template <class T> static inline T __shfl_xor_sync(unsigned mask, T v, int x) { return emu::collective<T>(mask, v, (emu::tid() & 31) ^ x); }
template <class T> static inline T __shfl_down_sync(unsigned mask, T v, unsigned d)
{ int l = emu::tid() & 31; return emu::collective<T>(mask, v, (l + (int) d < 32) ? l + (int) d : l); }
template <class T> static inline T __shfl_up_sync(unsigned mask, T v, unsigned d)
{ int l = emu::tid() & 31; return emu::collective<T>(mask, v, (l - (int) d >= 0) ? l - (int) d : l); }
static inline unsigned __ballot_sync(unsigned mask, int pred) { return emu::ballot(mask, pred); }
static inline int __any_sync(unsigned mask, int pred) { return emu::ballot(mask, pred) != 0; }
static inline int __all_sync(unsigned mask, int pred) { return emu::ballot(mask, pred) == mask; }
static inline unsigned __activemask() { return 0xffffffffu; }
static inline int __popc(unsigned v) { return __builtin_popcount(v); }
static inline int __clz(int v) { return v ? __builtin_clz((unsigned) v) : 32; }
static inline int __ffs(int v) { return __builtin_ffs(v); }
static inline int __float_as_int(float f) { int i; std::memcpy(&i, &f, 4); return i; }
static inline unsigned __float_as_uint(float f) { unsigned i; std::memcpy(&i, &f, 4); return i; }
static inline float __int_as_float(int i) { float f; std::memcpy(&f, &i, 4); return f; }
static inline float __uint_as_float(unsigned i) { float f; std::memcpy(&f, &i, 4); return f; }
static inline long long __double_as_longlong(double d) { long long i; std::memcpy(&i, &d, 8); return i; }
static inline double __longlong_as_double(long long i) { double d; std::memcpy(&d, &i, 8); return d; }
template <class T> static inline T __ldg(const T *p) { return *p; }
static inline int atomicAdd(int *p, int v) { return __atomic_fetch_add(p, v, __ATOMIC_RELAXED); }
static inline unsigned atomicAdd(unsigned *p, unsigned v) { return __atomic_fetch_add(p, v, __ATOMIC_RELAXED); }
static inline int atomicMax(int *p, int v) { int o = *p; if (v > o) *p = v; return o; }
static inline float fminf_(float a, float b) { return a < b ? a : b; }
#ifndef __CUDACC__
static inline int min(int a, int b) { return a < b ? a : b; }
static inline int max(int a, int b) { return a > b ? a : b; }
#endif
struct float2 { float x, y; };
struct float4 { float x, y, z, w; };
struct alignas(8) int2 { int x, y; };
