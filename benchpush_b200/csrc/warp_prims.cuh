// Lane-group primitives used by the env kernels (a group of BP_NL lanes simulates one environment).
//
// All cross-lane communication in the kernels goes through memory + bp_wsync() or through the few
// primitives below, so the same source also compiles as plain host C++ with a "warp" of ONE lane
// (BP_HOST_EMU: a "group" of ONE lane, used only by tests/ to debug kernel logic on machines without a GPU — the shipped
// library never contains or calls that build).
#pragma once
#include <stdint.h>
#include <float.h>

// Arithmetic type of the body state, contacts and solver ("real").  The library is built twice from the same sources:
// libbenchpush_b200.so with real = double and no FMA contraction (bit-comparable with the fp64 CPU oracle), and
// libbenchpush_b200_f32.so with -DBP_F32 (real = float): half the bytes per env, twice the ALU rate; its
// transitions agree with the oracle within the tolerance stated in tests/test_gpu_fp32.py and its episode scores are
// statistically indistinguishable (profiles/r02b_fp32_experiment.json).  Bookkeeping that has to stay conservative or
// accumulates over an episode (motion bounds D / G, pair gates, accumulated work, outputs) is double in both builds.
#ifdef BP_F32
typedef float real;
#define BP_RMIN FLT_MIN
#else
typedef double real;
#define BP_RMIN DBL_MIN
#endif
#define RC(x) ((real)(x))

#if defined(__CUDACC__) && !defined(BP_HOST_EMU)
#define BP_DEV __device__ __forceinline__
#define BP_DEVNI __device__ __noinline__
// An address-space fact the compiler cannot see through pointers that travel in structs / shared memory / out-of-line calls:
// env blocks live in device global memory.  With the assumption (made where an EnvView is formed, view_at) those accesses
// compile to LDG / STG instead of generic LD / ST: +2 % on the batched step (profiles/r02l_assume_global_ab.txt).  Extending
// it to the layout tables, the handle constants and the polygon scans (one loop per address space) was measured slower
// (profiles/r02m_address_space_ab.txt) and is not done.  -DBP_NO_ASSUME_SPACE builds the generic code.
#if defined(__CUDA_ARCH__) && !defined(BP_NO_ASSUME_SPACE)
#define BP_ASSUME_GLOBAL_PTR(p) __builtin_assume(__isGlobal((const void*)(p)))
#define BP_ASSUME_SHARED_PTR(p) __builtin_assume(__isShared((const void*)(p)))
#else
#define BP_ASSUME_GLOBAL_PTR(p) ((void)0)
#define BP_ASSUME_SHARED_PTR(p) ((void)0)
#endif
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
// group-relative ballot and indexed shuffle
BP_DEV unsigned bp_wballot(bool pred) { return __ballot_sync(bp_gmask(), pred) >> bp_gshift(); }
BP_DEV int bp_wshfl(int v, int src) { return __shfl_sync(bp_gmask(), v, src, BP_NL); }
BP_DEV double bp_wmax(double v) {
    const unsigned gm = bp_gmask();
#pragma unroll
    for (int o = BP_NL / 2; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(gm, v, o, BP_NL));
    return v;
}
// non-blocking hint: start moving the 128-byte line at p (and the following n_lines - 1) towards L1
BP_DEV void bp_prefetch(const void* p, int n_lines) {
    for (int i = 0; i < n_lines; i++) asm volatile("prefetch.global.L1 [%0];" ::"l"((const char*)p + 128 * i));
}
#else
#include <math.h>
#define BP_DEV static inline
#define BP_DEVNI static
#define BP_ASSUME_GLOBAL_PTR(p) ((void)0)
#define BP_ASSUME_SHARED_PTR(p) ((void)0)
#define BP_NL 1
BP_DEV int bp_lane() { return 0; }
BP_DEV void bp_wsync() {}
BP_DEV bool bp_wany(bool p) { return p; }
BP_DEV int bp_wbcast(int v) { return v; }
BP_DEV int bp_wrank(bool pred, int* total) { *total = pred ? 1 : 0; return 0; }
BP_DEV int bp_wscan_excl(int v, int* total) { *total = v; return 0; }
BP_DEV unsigned bp_wballot(bool pred) { return pred ? 1u : 0u; }
BP_DEV int bp_wshfl(int v, int src) { (void)src; return v; }
BP_DEV double bp_wmax(double v) { return v; }
BP_DEV void bp_prefetch(const void* p, int n_lines) { (void)p; (void)n_lines; }
#ifndef __CUDACC__
static inline int __popc(unsigned v) { return __builtin_popcount(v); }
static inline int __ffsll(long long v) { return __builtin_ffsll(v); }
static inline int __ffs(int v) { return __builtin_ffs(v); }
#endif
#endif

#if defined(__CUDACC__) && !defined(BP_HOST_EMU)
BP_DEV unsigned long long bp_d2bits(double x) { return (unsigned long long)__double_as_longlong(x); }
BP_DEV double bp_bits2d(unsigned long long b) { return __longlong_as_double((long long)b); }
#else
BP_DEV unsigned long long bp_d2bits(double x) { unsigned long long b; __builtin_memcpy(&b, &x, 8); return b; }
BP_DEV double bp_bits2d(unsigned long long b) { double x; __builtin_memcpy(&x, &b, 8); return x; }
#endif
// sin / cos of a body angle.  CUDA's sincos() and the host libm differ in the last bit now and then,
// and the contact dynamics amplify one ulp into millimetres over an episode, so the kernels use their
// own routine: argument reduction by pi/2 in three parts (Cody-Waite, constants of fdlibm's
// e_rem_pio2.c, second iteration always taken) followed by fdlibm's k_sin.c / k_cos.c polynomials,
// written with plain * + - only (the library is compiled with -fmad=false), hence bit-reproducible
// on any IEEE-754 machine.  Error < 1 ulp for |a| < 2^20 * pi/2.
BP_DEV void bp_sincos(double x, double* sn, double* cs) {
    const double invpio2 = 6.36619772367581382433e-01;
    const double pio2_1 = 1.57079632673412561417e+00, pio2_2 = 6.07710050630396597660e-11, pio2_2t = 2.02226624879595063154e-21;
    double fn = rint(x * invpio2);
    double t = x - fn * pio2_1;
    double w = fn * pio2_2;
    double r = t - w;
    w = fn * pio2_2t - ((t - r) - w);
    double y0 = r - w;
    double y1 = (r - y0) - w;
    const double S1 = -1.66666666666666324348e-01, S2 = 8.33333333332248946124e-03, S3 = -1.98412698298579493134e-04,
                 S4 = 2.75573137070700676789e-06, S5 = -2.50507602534068634195e-08, S6 = 1.58969099521155010221e-10;
    const double C1 = 4.16666666666666019037e-02, C2 = -1.38888888888741095749e-03, C3 = 2.48015872894767294178e-05,
                 C4 = -2.75573143513906633035e-07, C5 = 2.08757232129817482790e-09, C6 = -1.13596475577881948265e-11;
    double z = y0 * y0;
    double v = z * y0;
    double rs = S2 + z * (S3 + z * (S4 + z * (S5 + z * S6)));
    double ks = y0 - ((z * (0.5 * y1 - v * rs) - y1) - v * S1);
    double rc = z * (C1 + z * (C2 + z * (C3 + z * (C4 + z * (C5 + z * C6)))));
    double ax = fabs(y0), kc;
    if (ax < 0.3) kc = 1.0 - (0.5 * z - (z * rc - y0 * y1));
    else {
        double qx = (ax > 0.78125) ? 0.28125 : bp_bits2d((bp_d2bits(ax) - 0x0020000000000000ull) & 0xFFFFFFFF00000000ull);
        double hz = 0.5 * z - qx;
        double a = 1.0 - qx;
        kc = a - (hz - (z * rc - y0 * y1));
    }
    int n = (int)fn & 3;
    *sn = (n == 0) ? ks : (n == 1) ? kc : (n == 2) ? -ks : -kc;
    *cs = (n == 0) ? kc : (n == 1) ? -ks : (n == 2) ? -kc : ks;
}

#ifdef BP_F32
BP_DEV void bp_sincos(float x, float* sn, float* cs) { sincosf(x, sn, cs); }
#endif

// atomics (device) / plain ops (single-lane host emulation)
#if defined(__CUDACC__) && !defined(BP_HOST_EMU)
BP_DEV void bp_atomic_or(unsigned* p, unsigned v) { atomicOr(p, v); }
BP_DEV void bp_atomic_and(unsigned* p, unsigned v) { atomicAnd(p, v); }
BP_DEV int bp_atomic_add(int* p, int v) { return atomicAdd(p, v); }
BP_DEV void bp_atomic_max(int* p, int v) { atomicMax(p, v); }
BP_DEV void bp_atomic_max_u64(unsigned long long* p, unsigned long long v) { atomicMax(p, v); }
#else
BP_DEV void bp_atomic_or(unsigned* p, unsigned v) { *p |= v; }
BP_DEV void bp_atomic_and(unsigned* p, unsigned v) { *p &= v; }
BP_DEV int bp_atomic_add(int* p, int v) { int o = *p; *p += v; return o; }
BP_DEV void bp_atomic_max(int* p, int v) { if (v > *p) *p = v; }
BP_DEV void bp_atomic_max_u64(unsigned long long* p, unsigned long long v) { if (v > *p) *p = v; }
#endif
