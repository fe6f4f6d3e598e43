// Warp-level primitives used by the env kernels (one warp simulates one environment).
//
// All cross-lane communication in the kernels goes through memory + bp_wsync() or through the few
// primitives below, so the same source also compiles as plain host C++ with a "warp" of ONE lane
// (BP_HOST_EMU, used only by tests/ to debug kernel logic on machines without a GPU — the shipped
// library never contains or calls that build).
#pragma once
#include <stdint.h>

#if defined(__CUDACC__) && !defined(BP_HOST_EMU)
#define BP_DEV __device__ __forceinline__
#define BP_DEVNI __device__ __noinline__
#define BP_NL 32
BP_DEV int bp_lane() { return threadIdx.x & 31; }
BP_DEV void bp_wsync() { __syncwarp(); }
BP_DEV bool bp_wany(bool p) { return __any_sync(0xffffffffu, p); }
BP_DEV int bp_wbcast(int v) { return __shfl_sync(0xffffffffu, v, 0); }
// rank of this lane among the lanes with pred set (ascending lane order) and the total
BP_DEV int bp_wrank(bool pred, int* total) {
    unsigned m = __ballot_sync(0xffffffffu, pred);
    *total = __popc(m);
    return __popc(m & ((1u << bp_lane()) - 1u));
}
// exclusive prefix sum over lanes (ascending lane order) and the total
BP_DEV int bp_wscan_excl(int v, int* total) {
    int x = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        int y = __shfl_up_sync(0xffffffffu, x, o);
        if (bp_lane() >= o) x += y;
    }
    *total = __shfl_sync(0xffffffffu, x, 31);
    return x - v;
}
BP_DEV double bp_wmax(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
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
