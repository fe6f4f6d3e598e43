// Rigid-body core of the batched 2D pushing simulator: one warp steps one environment.
//
// Replaces, for the ship-ice path, the reference's per-env `pymunk.Space.step` loop
// (reference call sites ship_ice_env.py:117-120, 218-219, 280-290; semantics = Chipmunk2D 7.0.x
// cpSpaceStep with iterations=10, damping=0, gravity=0 — SURVEY.md §8a rows P1-P6, Appendix A).
//
// B200-first design:
//   * arithmetic in fp64 without FMA contraction (B200 has full-rate-half fp64 units) so that a
//     transition is reproducible to ~1e-12 against the fp64 CPU oracle, not merely "float close";
//   * damping = 0 makes the world quasi-static: a body with zero velocity, zero bias velocity and
//     zero cached impulses is EXACTLY a no-op in the sequential-impulse solver.  The kernel keeps an
//     "awake" set = union of complete contact components that contain any non-zero quantity (or
//     touch the kinematic robot) and steps only that set; everything else is provably untouched;
//   * broadphase = Verlet neighbour list (AABB gap < per-body margin) rebuilt once per env step
//     and whenever a body leaves its margin, so it is always a superset of Chipmunk's
//     cpBBIntersects filter; narrowphase = GJK/EPA on world vertices + 2-point manifolds;
//   * solver order = ascending pair id re-ordered by greedy colour, so all arbiters of one colour
//     (disjoint dynamic bodies) are solved by different lanes at once and the result is identical
//     to the sequential Gauss-Seidel sweep in that order (the CPU oracle runs the same order).
#pragma once
#include "warp_prims.cuh"
#include <float.h>
#include <math.h>

namespace bp {

constexpr int VMAX = 20;          // max vertices per convex shape (ice floes have 10..20)
constexpr int MAXCOL = 64;        // max solver colours

enum StatusBits {
    ST_OK = 0,
    ST_PAIR_OVERFLOW = 1,         // neighbour list exceeded PMAX
    ST_LIVE_OVERFLOW = 2,         // touching pairs exceeded LMAX
    ST_COLOUR_OVERFLOW = 4,       // more than MAXCOL colours needed
    ST_GJK_ITER = 8,              // GJK/EPA hit the iteration cap (informational)
};

struct SimParams {
    int iterations;               // solver sweeps (cfg.sim.iterations = 10)
    int persistence;              // collision_persistence = 3
    int substeps;                 // cfg.sim.steps = 400
    int settle_steps;             // 1000 (ship_ice_env.py:218)
    real dt;                    // cfg.dt / cfg.sim.steps
    real damp;                  // pow(cfg.sim.damping, dt)
    real bias_coef;             // 1 - pow(collision_bias, dt)
    real slop;                  // collision_slop = 0.1
    real target_speed;          // cfg.target_speed
    real yaw_step;              // max_yaw_rate_step = (pi/2)/7
    real yaw_lo, yaw_hi;        // (0, pi)
    real map_w, map_h;          // cfg.occ.map_width / map_height
    real goal_y;                // cfg.goal_y
    real beta;                  // 30
    real boundary_limit;        // 0.0
    real boundary_penalty, terminal_reward, yaw_penalty;   // -50, 200, 0
    real margin_body, margin_robot;                        // neighbour-list margins (m)
    // ---- env kind and the maze-NAMO constants (maze_NAMO_env.py:80-82, 406-410, 528-536)
    int kind;                     // 0 = ship-ice-v0, 1 = maze-NAMO-v0
    int n_walls;                  // static wall segments: bodies [0, n_walls) of a maze layout
    real goal_x, goal_gy;       // cfg.env.goal_x / goal_y
    real goal_reach;            // cfg.goal_radius + cfg.robot.min_r
    real k_increment;           // 150
    real pix_scale;             // cfg.occ.m_to_pix_scale (16)
    const double* dist_map;       // [dist_h][dist_w] normalised goal-point distance map (occupancy_map.py:435-485)
    int dist_h, dist_w;
};
enum EnvKind { KIND_SHIPICE = 0, KIND_MAZE = 1 };

struct Dims {
    int NB;                       // bodies per env incl. robot (capacity)
    int PMAX;                     // neighbour-list capacity (pairs)
    int LMAX;                     // live (touching / recently touching) arbiter capacity
    int n_envs;
    int n_layouts;
};

// Immutable per-layout geometry (shared by every env that currently runs that layout).
struct LayoutTab {
    const real* lv;             // [L][NB][VMAX][2] hull vertices, body-local, CCW
    const real* ln;             // [L][NB][VMAX][2] edge normals (edge ending at vertex i)
    const int* nv;                // [L][NB]
    const real* rad;            // [L][NB] shape radius
    const real* rbound;         // [L][NB] max |vertex - cog|: bounds the motion caused by rotation
    const real* minv;           // [L][NB]
    const real* iinv;           // [L][NB]
    const real* cog;            // [L][NB][2]
    const real* mat_e;          // [L][NB]
    const real* mat_u;          // [L][NB]
    const real* init;           // [L][NB][3] body position x, y, angle
    const int* nbody;             // [L] bodies incl. robot
    const int* robot0;            // [L] first body of the kinematic robot: its shapes are bodies [robot0, nbody), all
                                  //     carrying the same pose and velocity (a compound pymunk body, one record per shape)
};

struct EnvScal {
    int stamp, n_awake, n_pairs, cur, n_free, n_order, n_colours, dirty;
    int status, layout, nbody, robot, t, rebuild, n_apairs, n_cached;
    int any_wake, episode, done, first_layout;
    double G;                     // monotone accumulator: bounds the growth of D[a] + D[b] of ANY pair
    int obs_raw, n_fixed, wall_hit, has_prev_dist;   // n_fixed: bodies that are not dynamic (robot shapes + walls)
    double prev_dist;                                 // maze: distance-map value at the previous step (prev_dist_value)
    int slot_hw, pad_hw;          // every arbiter slot >= slot_hw is free: scans over the slots stop there; pad_hw != 0: the state was
                                  // injected (bp_*_set_bodies) and the next env step must not start with a sleep pass
    double total_work;
    long long stat_gjk, stat_pairs_eval, stat_live_sum, stat_awake_sum, stat_recolour, stat_rebuild, stat_substeps;
    long long stat_epa, stat_skipped, t_phase[8];
};

// Per-env mutable state lives in one block of global memory per env (identical offsets for every
// env).  Records are arrays-of-structs sized to cache lines: a lane that works on one body / one
// arbiter touches 1-2 lines instead of one line per field.
struct alignas(16) BodyRec {      // 128 B = one L1/L2 line: everything the solver needs of a body
    real px, py, pa;              // cog position (world), angle
    real vx, vy, vw;              // velocity
    real bx, by, bw;              // bias (pseudo) velocity
    real minv, iinv;              // inverse mass / moment (0 for the kinematic robot)
    real rc, rs;                  // cos / sin of pa
    double D;                       // monotone bound on the path length travelled by any point of the body
    real pad0, pad1;
};
struct BBox { real l, b, r, t; }; // AABB incl. shape radius
struct ContactRec {                 // 96 B
    real r1x, r1y, r2x, r2y, jn, jt, jb, nm, tm, bias, bounce;
    unsigned hash, pad;
};
struct alignas(16) SlotRec {      // 256 B = two lines: one (recently) touching shape pair = cpArbiter
    int a, b, pair, cnt, stamp;
    unsigned char touch, first, keep, pad;
    real nx, ny, e, u;
    ContactRec c[2];
    real pad2;
};
struct PairRec {                    // 24 B: neighbour-list entry
    unsigned key, gjk;              // (i << 16) | j ; GJK warm-start id
    int live;                       // slot index or -1
    unsigned flags;                 // bit0 = evaluated at the current poses, bit1 = in apairs
    double skip;                    // not touching and D[a] + D[b] < skip  =>  provably still not touching
};

struct EnvView {
    EnvScal* sc;
    BodyRec* body;                                                           // [NB]
    BBox* bb;                                                                // [NB] current AABB
    BBox* b0;                                                                // [NB] AABB at neighbour-list build
    real* wv;                                                              // [NB][VMAX][2] world vertices
    double *cenx, *ceny, *area, *tmp;                                       // [NB] |centroid|, area at last env step; scratch
    int* awake;                                                              // [NB]
    unsigned char *is_awake, *moved, *moved_step, *act;                      // [NB]
    unsigned char* mvrb;                                                     // [NB] AABB changed since the last neighbour-list build
    unsigned long long* colmask;                                             // [NB]
    PairRec* pair[2];                                                        // [PMAX] double-buffered over rebuilds
    int *adj_start, *adj, *deg;                                              // [NB+1], [2*PMAX], [NB]
    int *queue, *queue0, *apairs;                                            // [PMAX] eval queues; pairs incident to awake bodies
    double* ap_gate;                                                         // [PMAX] per apairs entry: no check needed while G < gate
    int* queue2; unsigned* qhint;                                            // [LMAX] EPA stage queue + simplex hint
    unsigned* livemask;                                                      // [PMAX/32]
    SlotRec* slot;                                                           // [LMAX]
    int *free_stack, *order, *colour, *col_start;                            // [LMAX],[LMAX],[LMAX],[MAXCOL+1]
};

#define BP_ALIGN16(x) (((x) + 15) & ~(size_t)15)
#define BP_ALIGN128(x) (((x) + 127) & ~(size_t)127)

#if defined(__CUDACC__) && !defined(BP_HOST_EMU)
#define BP_HD __host__ __device__ inline
#else
#define BP_HD static inline
#endif

// Carve the per-env block.  base == nullptr: returns only the size (in the sc field as an offset).
BP_HD EnvView carve_env(char* base, const Dims& d, size_t* total) {
    EnvView v;
    size_t NB = d.NB, P = d.PMAX, Lm = d.LMAX;
    char* p = base;
#define TAKE(ptr, type, bytes) ptr = (type*)p; p += BP_ALIGN128(bytes);
    TAKE(v.sc, EnvScal, sizeof(EnvScal))
    TAKE(v.body, BodyRec, NB * sizeof(BodyRec))
    TAKE(v.slot, SlotRec, Lm * sizeof(SlotRec))
    TAKE(v.bb, BBox, NB * sizeof(BBox)) TAKE(v.b0, BBox, NB * sizeof(BBox))
    TAKE(v.wv, real, NB * VMAX * 2 * sizeof(real))
    TAKE(v.cenx, double, NB * sizeof(double)) TAKE(v.ceny, double, NB * sizeof(double)) TAKE(v.area, double, NB * sizeof(double)) TAKE(v.tmp, double, NB * sizeof(double))
    TAKE(v.awake, int, NB * 4)
    TAKE(v.is_awake, unsigned char, NB) TAKE(v.moved, unsigned char, NB)
    TAKE(v.moved_step, unsigned char, NB) TAKE(v.act, unsigned char, NB)
    TAKE(v.colmask, unsigned long long, NB * 8)
    TAKE(v.pair[0], PairRec, P * sizeof(PairRec)) TAKE(v.pair[1], PairRec, P * sizeof(PairRec))
    TAKE(v.adj_start, int, (NB + 1) * 4) TAKE(v.adj, int, 2 * P * 4) TAKE(v.deg, int, NB * 4)
    TAKE(v.queue, int, P * 4) TAKE(v.queue0, int, P * 4) TAKE(v.apairs, int, P * 4) TAKE(v.ap_gate, double, P * 8)
    TAKE(v.queue2, int, Lm * 4) TAKE(v.qhint, unsigned, Lm * 4)
    TAKE(v.livemask, unsigned, (P / 32 + 1) * 4)
    TAKE(v.free_stack, int, Lm * 4) TAKE(v.order, int, Lm * 4) TAKE(v.colour, int, Lm * 4)
    TAKE(v.col_start, int, (MAXCOL + 1) * 4)
    TAKE(v.mvrb, unsigned char, NB)
#undef TAKE
    if (total) *total = (size_t)(p - base);
    return v;
}
BP_HD EnvView make_view(char* base, const Dims& d) { return carve_env(base, d, nullptr); }
static inline size_t env_mem_size(const Dims& d) { size_t t = 0; carve_env((char*)0, d, &t); return BP_ALIGN128(t); }

// The offsets of the per-env arrays are the same for every env of a handle: they are computed once
// on the host (env_offsets) so that device code can form `base + constant` for just the arrays it
// touches (view_at is inlined and dead fields disappear), instead of re-deriving the whole carve.
#define BP_ENV_FIELD_LIST(X)                                                                         \
    X(0, sc, EnvScal) X(1, body, BodyRec) X(2, slot, SlotRec) X(3, bb, BBox) X(4, b0, BBox) X(5, wv, real)      \
    X(6, cenx, double) X(7, ceny, double) X(8, area, double) X(9, tmp, double) X(10, awake, int)                 \
    X(11, is_awake, unsigned char) X(12, moved, unsigned char) X(13, moved_step, unsigned char)                  \
    X(14, act, unsigned char) X(15, colmask, unsigned long long) X(16, pair[0], PairRec) X(17, pair[1], PairRec) \
    X(18, adj_start, int) X(19, adj, int) X(20, deg, int) X(21, queue, int) X(22, queue0, int) X(23, apairs, int) \
    X(24, ap_gate, double) X(25, queue2, int) X(26, qhint, unsigned) X(27, livemask, unsigned)                   \
    X(28, free_stack, int) X(29, order, int) X(30, colour, int) X(31, col_start, int) X(32, mvrb, unsigned char)
struct EnvOff { unsigned o[33]; };
static inline EnvOff env_offsets(const Dims& d) {
    EnvOff O;
    EnvView v = carve_env((char*)0, d, nullptr);
#define X(i, member, type) O.o[i] = (unsigned)(size_t)v.member;
    BP_ENV_FIELD_LIST(X)
#undef X
    return O;
}
BP_DEV EnvView view_at(char* base, const EnvOff& O) {
    EnvView v;
    BP_ASSUME_GLOBAL_PTR(base);      // accesses through the view compile to LDG / STG instead of generic LD / ST: +2 % (profiles/r02l_assume_global_ab.txt)
#define X(i, member, type) v.member = (type*)(base + O.o[i]);
    BP_ENV_FIELD_LIST(X)
#undef X
    return v;
}

// The offsets again, in shared memory (one copy per thread block, filled at kernel entry by
// load_shared_offsets): device code of the batched hot path forms `base + bp_soff.o[i]` with an LDS at
// a fixed address instead of a global load through the DevConst pointer.
#if defined(__CUDACC__) && !defined(BP_HOST_EMU)
__shared__ EnvOff bp_soff;
BP_DEV void load_shared_offsets(const EnvOff& O) {
    for (int i = threadIdx.x; i < (int)(sizeof(EnvOff) / sizeof(unsigned)); i += blockDim.x) bp_soff.o[i] = O.o[i];
    __syncthreads();
}
#else
static EnvOff bp_soff;
static inline void load_shared_offsets(const EnvOff& O) { bp_soff = O; }
#endif
BP_DEV EnvView view_s(char* base) { return view_at(base, bp_soff); }

struct LayoutView {
    const real *lv, *ln, *rad, *rbound, *minv, *iinv, *cog, *mat_e, *mat_u, *init;
    const int* nv;
    int nbody, robot0;
};
// Everything a kernel needs that is constant for a handle; lives in device memory (one copy per
// handle) so that out-of-line device functions can take it as ONE pointer.
struct DevConst {
    Dims d; SimParams P; EnvOff off; LayoutTab T;
    // reset snapshots: the env block right after reset depends on the layout only, so it is simulated
    // once per layout (init + settle substeps) and a reset is a copy of that block.  nullptr = none.
    const char* snap; size_t env_bytes;
    const double* snap_pose;      // [L][3] ship pose after reset
    const int* snap_status;       // [L]
    // cfg.random_start (ship_ice_env.py:201-203, maze_NAMO_env.py:229-238): per-env start pose of the robot, read at every
    // reset of env e from start_pose[3e..3e+2] (caller-owned device memory); such resets cannot use the snapshots.  nullptr = the layout's.
    const double* start_pose;
};
BP_DEV LayoutView layout_view(const LayoutTab& T, const Dims& d, int l) {
    LayoutView L;
    size_t o = (size_t)l * d.NB;
    L.lv = T.lv + o * VMAX * 2; L.ln = T.ln + o * VMAX * 2; L.nv = T.nv + o; L.rad = T.rad + o; L.rbound = T.rbound + o;
    L.minv = T.minv + o; L.iinv = T.iinv + o; L.cog = T.cog + o * 2; L.mat_e = T.mat_e + o; L.mat_u = T.mat_u + o;
    L.init = T.init + o * 3; L.nbody = T.nbody[l]; L.robot0 = T.robot0[l];
    return L;
}

// ------------------------------------------------------------------ small vector helpers
struct V2 { real x, y; };
struct alignas(2 * sizeof(real)) D2 { real x, y; };      // one vertex in memory: moved with a single 128-bit load / store
BP_DEV V2 mk(real x, real y) { V2 r; r.x = x; r.y = y; return r; }
BP_DEV V2 operator+(V2 a, V2 b) { return mk(a.x + b.x, a.y + b.y); }
BP_DEV V2 operator-(V2 a, V2 b) { return mk(a.x - b.x, a.y - b.y); }
BP_DEV V2 operator*(V2 a, real s) { return mk(a.x * s, a.y * s); }
BP_DEV V2 neg(V2 a) { return mk(-a.x, -a.y); }
BP_DEV real dot(V2 a, V2 b) { return a.x * b.x + a.y * b.y; }
BP_DEV real cross(V2 a, V2 b) { return a.x * b.y - a.y * b.x; }
BP_DEV V2 perp(V2 a) { return mk(-a.y, a.x); }
BP_DEV V2 rperp(V2 a) { return mk(a.y, -a.x); }
BP_DEV real clampd(real f, real lo, real hi) { return fmin(fmax(f, lo), hi); }
BP_DEV real clamp01(real f) { return fmax(RC(0.0), fmin(f, RC(1.0))); }
BP_DEV V2 lerp(V2 a, V2 b, real t) { return a * (RC(1.0) - t) + b * t; }
BP_DEV V2 rotv(V2 a, V2 b) { return mk(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }

// ------------------------------------------------------------------ geometry refresh of one body
BP_DEV void refresh_body_impl(const EnvView& ev, const LayoutView& L, int b) {
    real s, c;
    bp_sincos(ev.body[b].pa, &s, &c);
    ev.body[b].rc = c; ev.body[b].rs = s;
    real cx = L.cog[2 * b], cy = L.cog[2 * b + 1];
    real tx = ev.body[b].px - (cx * c - cy * s);
    real ty = ev.body[b].py - (cx * s + cy * c);
    real ms = -s;
    const real* lv = L.lv + (size_t)b * VMAX * 2;
    real* wv = ev.wv + (size_t)b * VMAX * 2;
    int n = L.nv[b];
    real l = INFINITY, r = -INFINITY, bt = INFINITY, tp = -INFINITY;
    const D2* lv2 = (const D2*)lv;
    D2* wv2 = (D2*)wv;
#pragma unroll 2
    for (int i = 0; i < n; i++) {
        D2 q = lv2[i];
        real wx = c * q.x + ms * q.y + tx;
        real wy = s * q.x + c * q.y + ty;
        D2 w; w.x = wx; w.y = wy;
        wv2[i] = w;
        l = fmin(l, wx); r = fmax(r, wx); bt = fmin(bt, wy); tp = fmax(tp, wy);
    }
    real rad = L.rad[b];
    BBox bx; bx.l = l - rad; bx.b = bt - rad; bx.r = r + rad; bx.t = tp + rad;
    ev.bb[b] = bx;
}
BP_DEVNI void refresh_body(const EnvView& ev, const LayoutView& L, int b) { refresh_body_impl(ev, L, b); }
// same, for callers that hold only the env block (batched hot path: only the touched arrays are addressed)
BP_DEVNI void refresh_body_b(const DevConst* C, char* base, const LayoutView& L, int b) { refresh_body_impl(view_s(base), L, b); }
// body position (pymunk Body.position = local origin in world space)
BP_DEV V2 body_origin(const EnvView& ev, const LayoutView& L, int b) {
    real c = ev.body[b].rc, s = ev.body[b].rs;
    real cx = L.cog[2 * b], cy = L.cog[2 * b + 1];
    return mk(ev.body[b].px - (cx * c - cy * s), ev.body[b].py - (cx * s + cy * c));
}

// ------------------------------------------------------------------ narrowphase: GJK / EPA on world vertices
struct MPoint { V2 ab; unsigned id; };       // Minkowski difference vertex; id = (index_a << 8) | index_b
struct Poly { const real* v; int n; };

// The one copy of the support scan (GJK, EPA and the manifold builder all call it): the hot path is
// instruction-fetch bound (profiles/), so shared out-of-line leaves beat inlined copies.
BP_DEVNI int support_index(const Poly& P, V2 n) {
    // a 2-vertex shape is a cpSegmentShape (walls): SegmentSupportPoint returns ta only when strictly greater
    if (P.n == 2) return (P.v[0] * n.x + P.v[1] * n.y > P.v[2] * n.x + P.v[3] * n.y) ? 0 : 1;
    const D2* v = (const D2*)P.v;
    real mx = -INFINITY; int idx = 0;
#pragma unroll 4
    for (int i = 0; i < P.n; i++) {
        D2 q = v[i];
        real d = q.x * n.x + q.y * n.y;
        if (d > mx) { mx = d; idx = i; }
    }
    return idx;
}
BP_DEV V2 pvert(const Poly& P, int i) { D2 q = ((const D2*)P.v)[i]; return mk(q.x, q.y); }
BP_DEV MPoint mpoint(const Poly& A, const Poly& B, int ia, int ib) {
    MPoint m; m.ab = pvert(B, ib) - pvert(A, ia); m.id = (unsigned)((ia & 0xFF) << 8 | (ib & 0xFF)); return m;
}
BP_DEV MPoint msupport(const Poly& A, const Poly& B, V2 n) {
    return mpoint(A, B, support_index(A, neg(n)), support_index(B, n));
}
BP_DEV real closest_t(V2 a, V2 b) {
    V2 delta = b - a;
    return -clampd(dot(delta, a + b) / (dot(delta, delta) + BP_RMIN), -RC(1.0), RC(1.0));
}
BP_DEV V2 lerp_t(V2 a, V2 b, real t) {
    real ht = RC(0.5) * t;
    return a * (RC(0.5) - ht) + b * (RC(0.5) + ht);
}
BP_DEV real closest_dist(V2 a, V2 b) { V2 p = lerp_t(a, b, closest_t(a, b)); return dot(p, p); }
BP_DEV bool point_greater(V2 a, V2 b, V2 c) {
    return (b.y - a.y) * (a.x + b.x - 2 * c.x) > (b.x - a.x) * (a.y + b.y - 2 * c.y);
}
struct Closest { V2 n; real d; unsigned id; };
BP_DEVNI Closest closest_points(const MPoint& v0, const MPoint& v1) {
    real t = closest_t(v0.ab, v1.ab);
    V2 p = lerp_t(v0.ab, v1.ab, t);
    V2 delta = v1.ab - v0.ab;
    V2 rp = rperp(delta);
    V2 n = rp * (RC(1.0) / (sqrt(dot(rp, rp)) + BP_RMIN));
    real d = dot(n, p);
    Closest c; c.id = (v0.id & 0xFFFF) << 16 | (v1.id & 0xFFFF);
    if (d <= RC(0.0) || (-RC(1.0) < t && t < RC(1.0))) { c.n = n; c.d = d; }
    else { real d2 = sqrt(dot(p, p)); c.n = p * (RC(1.0) / (d2 + BP_RMIN)); c.d = d2; }
    return c;
}

constexpr int GJK_MAX_IT = 30, EPA_MAX_IT = 30;

BP_DEVNI Closest epa(const Poly& A, const Poly& B, MPoint h0, MPoint h1, MPoint h2, int* capped) {
    MPoint hullA[EPA_MAX_IT + 4], hullB[EPA_MAX_IT + 4];
    MPoint* hull = hullA; MPoint* hull2 = hullB;
    hull[0] = h0; hull[1] = h1; hull[2] = h2;
    int count = 3, iteration = 1;
    for (;;) {
        int mini = 0; real minDist = INFINITY;
#pragma unroll 1
        for (int j = 0, i = count - 1; j < count; i = j, j++) {
            real d = closest_dist(hull[i].ab, hull[j].ab);
            if (d < minDist) { minDist = d; mini = i; }
        }
        MPoint v0 = hull[mini], v1 = hull[(mini + 1) % count];
        MPoint p = msupport(A, B, perp(v1.ab - v0.ab));
        bool dup = (p.id == v0.id || p.id == v1.id);
        if (!dup && point_greater(v0.ab, v1.ab, p.ab) && iteration < EPA_MAX_IT) {
            int count2 = 1;
            hull2[0] = p;
#pragma unroll 1
            for (int i = 0; i < count; i++) {
                int index = (mini + 1 + i) % count;
                V2 q0 = hull2[count2 - 1].ab;
                V2 q1 = hull[index].ab;
                V2 q2 = (i + 1 < count ? hull[(index + 1) % count] : p).ab;
                if (point_greater(q0, q2, q1)) { hull2[count2] = hull[index]; count2++; }
            }
            MPoint* tmp = hull; hull = hull2; hull2 = tmp;
            count = count2; iteration++;
        } else {
            if (iteration >= EPA_MAX_IT) *capped = 1;
            return closest_points(v0, v1);
        }
    }
}

// GJK without the EPA tail.  Returns true when the origin lies inside the simplex (v0, p, v1): the
// caller must then run EPA on it; *hint packs (v0.id, v1.id) so that gjk_epa() can rebuild exactly
// that simplex (p is a pure function of the ordered pair v0, v1).
BP_DEVNI bool gjk_only(const Poly& A, const Poly& B, V2 cenA, V2 cenB, unsigned id, Closest* out, unsigned* hint, int* iters, int* capped) {
    MPoint v0, v1;
    if (id) {
        v0 = mpoint(A, B, (id >> 24) & 0xFF, (id >> 16) & 0xFF);
        v1 = mpoint(A, B, (id >> 8) & 0xFF, id & 0xFF);
    } else {
        V2 axis = perp(cenA - cenB);
        v0 = msupport(A, B, axis);
        v1 = msupport(A, B, neg(axis));
    }
    int iteration = 1;
    for (;;) {
        (*iters)++;
        if (iteration > GJK_MAX_IT) { *capped = 1; *out = closest_points(v0, v1); return false; }
        if (point_greater(v1.ab, v0.ab, mk(0, 0))) { MPoint t = v0; v0 = v1; v1 = t; }
        real t = closest_t(v0.ab, v1.ab);
        V2 n = (-RC(1.0) < t && t < RC(1.0)) ? perp(v1.ab - v0.ab) : neg(lerp_t(v0.ab, v1.ab, t));
        MPoint p = msupport(A, B, n);
        if (point_greater(p.ab, v0.ab, mk(0, 0)) && point_greater(v1.ab, p.ab, mk(0, 0))) {
            *hint = (v0.id & 0xFFFF) << 16 | (v1.id & 0xFFFF);
            return true;
        }
        if (dot(p.ab, n) <= fmax(dot(v0.ab, n), dot(v1.ab, n))) { *out = closest_points(v0, v1); return false; }
        if (closest_dist(v0.ab, p.ab) < closest_dist(p.ab, v1.ab)) v1 = p; else v0 = p;
        iteration++;
    }
}
// EPA stage for a pair whose GJK stage ended inside the simplex described by `hint`.
BP_DEVNI Closest gjk_epa(const Poly& A, const Poly& B, unsigned hint, int* capped) {
    MPoint v0 = mpoint(A, B, (hint >> 24) & 0xFF, (hint >> 16) & 0xFF);
    MPoint v1 = mpoint(A, B, (hint >> 8) & 0xFF, hint & 0xFF);
    real t = closest_t(v0.ab, v1.ab);
    V2 n = (-RC(1.0) < t && t < RC(1.0)) ? perp(v1.ab - v0.ab) : neg(lerp_t(v0.ab, v1.ab, t));
    MPoint p = msupport(A, B, n);
    return epa(A, B, v0, p, v1, capped);
}

struct EdgeW { V2 a, b; int ia, ib; };
BP_DEV EdgeW support_edge(const Poly& P, const real* ln, real rc, real rs, V2 n) {
    int count = P.n;
    real ms = -rs;
    if (count == 2) {
        // SupportEdgeForSegment (cpCollision.c): ln[0] is the segment normal rperp(normalize(b - a))
        V2 tn = mk(rc * ln[0] + ms * ln[1], rs * ln[0] + rc * ln[1]);
        EdgeW e;
        if (dot(tn, n) > RC(0.0)) { e.a = pvert(P, 0); e.b = pvert(P, 1); e.ia = 0; e.ib = 1; }
        else { e.a = pvert(P, 1); e.b = pvert(P, 0); e.ia = 1; e.ib = 0; }
        return e;
    }
    int i1 = support_index(P, n);
    int i0 = (i1 - 1 + count) % count, i2 = (i1 + 1) % count;
    V2 n1 = mk(rc * ln[2 * i1] + ms * ln[2 * i1 + 1], rs * ln[2 * i1] + rc * ln[2 * i1 + 1]);
    V2 n2 = mk(rc * ln[2 * i2] + ms * ln[2 * i2 + 1], rs * ln[2 * i2] + rc * ln[2 * i2 + 1]);
    EdgeW e;
    if (dot(n, n1) > dot(n, n2)) { e.a = pvert(P, i0); e.b = pvert(P, i1); e.ia = i0; e.ib = i1; }
    else { e.a = pvert(P, i1); e.b = pvert(P, i2); e.ia = i1; e.ib = i2; }
    return e;
}

struct Manifold { int count; V2 n; V2 p1[2], p2[2]; unsigned hash[2]; };

BP_DEV Poly body_poly(const EnvView& ev, const LayoutView& L, int b) { Poly P; P.v = ev.wv + (size_t)b * VMAX * 2; P.n = L.nv[b]; return P; }
BP_DEV V2 bb_centre(const BBox& r) { return mk((r.l + r.r) * RC(0.5), (r.b + r.t) * RC(0.5)); }

// Contact manifold (<= 2 points, world space) from the closest-feature result.  The polygons, radii, local edge normals and
// rotations come in as plain values / pointers, so the same arithmetic serves the env block in global memory and the
// shared-memory resident copy of the awake bodies (sim_res.cuh).
BP_DEV void make_manifold_core(const Poly& A, const Poly& B, real ra, real rb, const real* lnA, const real* lnB,
                               real rcA, real rsA, real rcB, real rsB, const Closest& pts, Manifold* m) {
    m->count = 0;
    if (pts.d - ra - rb <= RC(0.0)) {
        EdgeW e1 = support_edge(A, lnA, rcA, rsA, pts.n);
        EdgeW e2 = support_edge(B, lnB, rcB, rsB, neg(pts.n));
        real mindist = ra + rb;
        if (pts.d <= mindist) {
            V2 n = pts.n; m->n = n;
            real d_e1_a = cross(e1.a, n), d_e1_b = cross(e1.b, n);
            real d_e2_a = cross(e2.a, n), d_e2_b = cross(e2.b, n);
            real e1_denom = RC(1.0) / (d_e1_b - d_e1_a + BP_RMIN);
            real e2_denom = RC(1.0) / (d_e2_b - d_e2_a + BP_RMIN);
            {
                V2 p1 = n * ra + lerp(e1.a, e1.b, clamp01((d_e2_b - d_e1_a) * e1_denom));
                V2 p2 = n * (-rb) + lerp(e2.a, e2.b, clamp01((d_e1_a - d_e2_a) * e2_denom));
                if (dot(p2 - p1, n) <= RC(0.0)) {
                    int k = m->count++;
                    m->p1[k] = p1; m->p2[k] = p2; m->hash[k] = (unsigned)(e1.ia << 8 | e2.ib);
                }
            }
            {
                V2 p1 = n * ra + lerp(e1.a, e1.b, clamp01((d_e2_a - d_e1_a) * e1_denom));
                V2 p2 = n * (-rb) + lerp(e2.a, e2.b, clamp01((d_e1_b - d_e2_a) * e2_denom));
                if (dot(p2 - p1, n) <= RC(0.0)) {
                    int k = m->count++;
                    m->p1[k] = p1; m->p2[k] = p2; m->hash[k] = (unsigned)(e1.ib << 8 | e2.ia);
                }
            }
        }
    }
}
BP_DEV void make_manifold_impl(const EnvView& ev, const LayoutView& L, int a, int b, const Closest& pts, Manifold* m) {
    make_manifold_core(body_poly(ev, L, a), body_poly(ev, L, b), L.rad[a], L.rad[b], L.ln + (size_t)a * VMAX * 2, L.ln + (size_t)b * VMAX * 2,
                       ev.body[a].rc, ev.body[a].rs, ev.body[b].rc, ev.body[b].rs, pts, m);
}

BP_DEVNI void make_manifold(const EnvView& ev, const LayoutView& L, int a, int b, const Closest& pts, Manifold* m) { make_manifold_impl(ev, L, a, b, pts, m); }
BP_DEVNI void make_manifold_b(const DevConst* C, char* base, const LayoutView& L, int a, int b, const Closest& pts, Manifold* m) {
    make_manifold_impl(view_s(base), L, a, b, pts, m);
}

}  // namespace bp
