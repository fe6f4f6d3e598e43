// Lane-group primitives used by the env kernels (a group of BP_NL lanes simulates one environment).
//
// All cross-lane communication in the kernels goes through memory + bp_wsync() or through the few
// primitives below, so the same source also compiles as plain host C++ with a "warp" of ONE lane
// (BP_HOST_EMU: a "group" of ONE lane, used only by tests/ to debug kernel logic on machines without a GPU — the shipped
// library never contains or calls that build).
#pragma once
#include <stdint.h>

#if defined(__CUDACC__) && !defined(BP_HOST_EMU)
#define BP_DEV __device__ __forceinline__
#define BP_DEVNI __device__ __noinline__
// One environment is simulated by a GROUP of BP_NL consecutive lanes of a warp (BP_NL = 32, 16, 8 or
// 4; 32 / BP_NL environments share a warp).  Groups of one warp start every substep together and run
// the same instruction stream, so SIMT divergence/reconvergence keeps them in lockstep through the
// phases of a substep; all cross-lane primitives below are scoped to the calling lane's group
// (sub-warp masks, like cooperative_groups::thread_block_tile).
#ifndef BP_GROUP
#define BP_GROUP 32
#endif
#define BP_NL BP_GROUP
#define BP_GROUPS_PER_WARP (32 / BP_NL)
BP_DEV int bp_lane() { return threadIdx.x & (BP_NL - 1); }
BP_DEV int bp_group_in_warp() { return (threadIdx.x & 31) / BP_NL; }
#if BP_GROUP == 32
BP_DEV unsigned bp_gshift() { return 0u; }
BP_DEV unsigned bp_gmask() { return 0xffffffffu; }
#else
BP_DEV unsigned bp_gshift() { return threadIdx.x & 31u & ~(unsigned)(BP_NL - 1); }
BP_DEV unsigned bp_gmask() { return ((1u << BP_NL) - 1u) << bp_gshift(); }
#endif
BP_DEV void bp_wsync() { __syncwarp(bp_gmask()); }
BP_DEV bool bp_wany(bool p) { return __any_sync(bp_gmask(), p); }
BP_DEV int bp_wbcast(int v) { return __shfl_sync(bp_gmask(), v, 0, BP_NL); }
// rank of this lane among the lanes of its group with pred set (ascending lane order) and the total
BP_DEV int bp_wrank(bool pred, int* total) {
    unsigned m = __ballot_sync(bp_gmask(), pred) >> bp_gshift();
    *total = __popc(m);
    return __popc(m & ((1u << bp_lane()) - 1u));
}
// exclusive prefix sum over the lanes of the group (ascending lane order) and the total
BP_DEV int bp_wscan_excl(int v, int* total) {
    int x = v;
    const unsigned gm = bp_gmask();
#pragma unroll
    for (int o = 1; o < BP_NL; o <<= 1) {
        int y = __shfl_up_sync(gm, x, o, BP_NL);
        if (bp_lane() >= o) x += y;
    }
    *total = __shfl_sync(gm, x, BP_NL - 1, BP_NL);
    return x - v;
}
BP_DEV double bp_wmax(double v) {
    const unsigned gm = bp_gmask();
#pragma unroll
    for (int o = BP_NL / 2; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(gm, v, o, BP_NL));
    return v;
}
BP_DEV void bp_sincos(double a, double* s, double* c) { sincos(a, s, c); }
#else
#include <math.h>
#define BP_DEV static inline
#define BP_DEVNI static
#define BP_NL 1
BP_DEV int bp_lane() { return 0; }
BP_DEV void bp_wsync() {}
BP_DEV bool bp_wany(bool p) { return p; }
BP_DEV int bp_wbcast(int v) { return v; }
BP_DEV int bp_wrank(bool pred, int* total) { *total = pred ? 1 : 0; return 0; }
BP_DEV int bp_wscan_excl(int v, int* total) { *total = v; return 0; }
BP_DEV double bp_wmax(double v) { return v; }
BP_DEV void bp_sincos(double a, double* s, double* c) { *s = sin(a); *c = cos(a); }
#ifndef __CUDACC__
static inline int __popc(unsigned v) { return __builtin_popcount(v); }
static inline int __ffsll(long long v) { return __builtin_ffsll(v); }
#endif
#endif

// atomics (device) / plain ops (single-lane host emulation)
#if defined(__CUDACC__) && !defined(BP_HOST_EMU)
BP_DEV void bp_atomic_or(unsigned* p, unsigned v) { atomicOr(p, v); }
BP_DEV void bp_atomic_and(unsigned* p, unsigned v) { atomicAnd(p, v); }
BP_DEV int bp_atomic_add(int* p, int v) { return atomicAdd(p, v); }
#else
BP_DEV void bp_atomic_or(unsigned* p, unsigned v) { *p |= v; }
BP_DEV void bp_atomic_and(unsigned* p, unsigned v) { *p &= v; }
BP_DEV int bp_atomic_add(int* p, int v) { int o = *p; *p += v; return o; }
#endif
