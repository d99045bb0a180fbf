/* cuda_shim.h -- just enough of the CUDA device vocabulary to compile the engine's device RULE code
 * (sugarscape_b200/csrc/ssb_rules.cuh, ssb_exact.cuh, ssb_ext.cuh) with plain g++ as ONE lane
 * (lane group G = 1: shuffles are the identity, warp syncs are no-ops, atomics are plain updates).
 *
 * TEST INFRASTRUCTURE ONLY.  This is not a CPU fallback of the product: nothing under
 * sugarscape_b200/ or sugarscape.py can load the emulation library; the engine raises when
 * libssb.so or a CUDA device is missing.  The purpose is to run the mode E transaction source --
 * the very text nvcc compiles for sm_100a -- against the golden fixtures in the `-m "not gpu"`
 * suite, so that rule changes are checked in the build container before GPU time is spent. */
#pragma once
#define SSB_HOSTEMU 1
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define __device__
#define __host__
#define __global__
#define __shared__ static
#define __forceinline__ inline
#define __noinline__
#define __launch_bounds__(...)
#define __align__(n) alignas(n)

struct ssb_emu_dim3 { unsigned x, y, z; };
struct int2 { int x, y; };
static inline int2 make_int2(int x, int y) { int2 v = {x, y}; return v; }
static const ssb_emu_dim3 threadIdx = {0, 0, 0};

template <class T> static inline T __shfl_sync(unsigned, T v, int, int = 32) { return v; }
template <class T> static inline T __shfl_xor_sync(unsigned, T v, int, int = 32) { return v; }
static inline void __syncwarp(unsigned = 0xffffffffu) {}
static inline int __any_sync(unsigned, int p) { return p; }
static inline unsigned __ballot_sync(unsigned, int p) { return p ? 1u : 0u; }
static inline int __all_sync(unsigned, int p) { return p; }
static inline void __threadfence_block() {}
static inline void __threadfence() {}

static inline int __popc(unsigned v) { return __builtin_popcount(v); }
static inline int __ffs(int v) { return __builtin_ffs(v); }
static inline int __popcll(unsigned long long v) { return __builtin_popcountll(v); }
static inline int __clz(unsigned v) { return v ? __builtin_clz(v) : 32; }
static inline long long __double_as_longlong(double d) { long long u; memcpy(&u, &d, 8); return u; }
static inline double __longlong_as_double(long long u) { double d; memcpy(&d, &u, 8); return d; }

static inline unsigned long long atomicAdd(unsigned long long *p, unsigned long long v) { const unsigned long long o = *p; *p = o + v; return o; }
static inline int atomicAdd(int *p, int v) { const int o = *p; *p = o + v; return o; }
static inline unsigned long long atomicExch(unsigned long long *p, unsigned long long v) { const unsigned long long o = *p; *p = v; return o; }
static inline unsigned long long atomicMax(unsigned long long *p, unsigned long long v) { const unsigned long long o = *p; if (v > o) *p = v; return o; }

static inline int min(int a, int b) { return a < b ? a : b; }
static inline int max(int a, int b) { return a > b ? a : b; }
static inline long long min(long long a, long long b) { return a < b ? a : b; }
static inline long long max(long long a, long long b) { return a > b ? a : b; }
static inline unsigned min(unsigned a, unsigned b) { return a < b ? a : b; }
static inline unsigned max(unsigned a, unsigned b) { return a > b ? a : b; }

#include "../../include/ssb.h"
namespace ssb {
/* the fields of ssb_shard.cuh's ShardCtx that the rule code reads (world <= 1: not sharded) */
struct ShardCtx { int32_t rank, world; long long slot_lo, slot_hi; };
}
