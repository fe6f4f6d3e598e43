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
    double dt;                    // cfg.dt / cfg.sim.steps
    double damp;                  // pow(cfg.sim.damping, dt)
    double bias_coef;             // 1 - pow(collision_bias, dt)
    double slop;                  // collision_slop = 0.1
    double target_speed;          // cfg.target_speed
    double yaw_step;              // max_yaw_rate_step = (pi/2)/7
    double yaw_lo, yaw_hi;        // (0, pi)
    double map_w, map_h;          // cfg.occ.map_width / map_height
    double goal_y;                // cfg.goal_y
    double beta;                  // 30
    double boundary_limit;        // 0.0
    double boundary_penalty, terminal_reward, yaw_penalty;   // -50, 200, 0
    double margin_body, margin_robot;                        // neighbour-list margins (m)
};

struct Dims {
    int NB;                       // bodies per env incl. robot (capacity)
    int PMAX;                     // neighbour-list capacity (pairs)
    int LMAX;                     // live (touching / recently touching) arbiter capacity
    int n_envs;
    int n_layouts;
};

// Immutable per-layout geometry (shared by every env that currently runs that layout).
struct LayoutTab {
    const double* lv;             // [L][NB][VMAX][2] hull vertices, body-local, CCW
    const double* ln;             // [L][NB][VMAX][2] edge normals (edge ending at vertex i)
    const int* nv;                // [L][NB]
    const double* rad;            // [L][NB] shape radius
    const double* minv;           // [L][NB]
    const double* iinv;           // [L][NB]
    const double* cog;            // [L][NB][2]
    const double* mat_e;          // [L][NB]
    const double* mat_u;          // [L][NB]
    const double* init;           // [L][NB][3] body position x, y, angle
    const int* nbody;             // [L] bodies incl. robot; robot index = nbody-1
};

struct EnvScal {
    int stamp, n_awake, n_pairs, cur, n_free, n_order, n_colours, dirty;
    int status, layout, nbody, robot, t, rebuild, n_apairs, n_cached;
    int any_wake, episode, done, first_layout;
    int obs_raw, pad0, pad1, pad2;
    double total_work;
    long long stat_gjk, stat_pairs_eval, stat_live_sum, stat_awake_sum, stat_recolour, stat_rebuild, stat_substeps;
};

// Per-env mutable state (global memory; offsets are identical for every env).
struct EnvView {
    EnvScal* sc;
    double *px, *py, *pa, *vx, *vy, *vw, *bx, *by, *bw, *rc, *rs;           // [NB]
    double *bbl, *bbb, *bbr, *bbt, *b0l, *b0b, *b0r, *b0t;                   // [NB] current / at-build AABB
    double *wv;                                                              // [NB][VMAX][2]
    double *cenx, *ceny, *area, *tmp;                                        // [NB] |centroid|, area at last env step; scratch
    int *awake;                                                              // [NB]
    unsigned char *is_awake, *moved, *moved_step, *act;                      // [NB]
    unsigned long long *colmask;                                             // [NB]
    unsigned *pkey[2], *pgjk[2];                                             // [PMAX]
    int *plive[2];                                                           // [PMAX]
    unsigned char *pflag[2];                                                 // [PMAX] bit0 = evaluated at current poses, bit1 = in apairs
    int *adj_start, *adj, *deg;                                              // [NB+1], [2*PMAX], [NB]
    int *queue, *apairs;                                                     // [PMAX] eval queue; pairs incident to awake bodies
    unsigned *livemask;                                                      // [PMAX/32]
    // live arbiter slots
    int *ls_a, *ls_b, *ls_pair, *ls_cnt, *ls_stamp;                          // [LMAX]
    unsigned char *ls_touch, *ls_first, *ls_keep;                            // [LMAX]
    double *ls_nx, *ls_ny, *ls_e, *ls_u;                                     // [LMAX]
    double *ls_r1x, *ls_r1y, *ls_r2x, *ls_r2y, *ls_jn, *ls_jt, *ls_jb, *ls_nm, *ls_tm, *ls_bias, *ls_bounce;  // [2][LMAX]
    unsigned *ls_hash;                                                       // [2][LMAX]
    int *free_stack, *order, *colour, *col_start;                            // [LMAX],[LMAX],[LMAX],[MAXCOL+1]
};

#define BP_ALIGN16(x) (((x) + 15) & ~(size_t)15)

// Host + device: carve the per-env block.  Order must match make_view().
static inline size_t env_mem_size(const Dims& d) {
    size_t NB = d.NB, P = d.PMAX, Lm = d.LMAX, t = 0;
    t += BP_ALIGN16(sizeof(EnvScal));
    t += 19 * BP_ALIGN16(NB * 8);                 // px..rs (11) + bb (8)
    t += BP_ALIGN16(NB * VMAX * 2 * 8);           // wv
    t += 4 * BP_ALIGN16(NB * 8);                  // cenx, ceny, area, tmp
    t += BP_ALIGN16(NB * 4);                      // awake
    t += 4 * BP_ALIGN16(NB);                      // is_awake, moved, moved_step, act
    t += BP_ALIGN16(NB * 8);                      // colmask
    t += 4 * BP_ALIGN16(P * 4);                   // pkey[2], pgjk[2]
    t += 2 * BP_ALIGN16(P * 4);                   // plive[2]
    t += 2 * BP_ALIGN16(P);                       // pflag[2]
    t += BP_ALIGN16((NB + 1) * 4) + BP_ALIGN16(2 * P * 4) + BP_ALIGN16(NB * 4);   // adj_start, adj, deg
    t += 2 * BP_ALIGN16(P * 4);                   // queue, apairs
    t += BP_ALIGN16((P / 32 + 1) * 4);            // livemask
    t += 5 * BP_ALIGN16(Lm * 4);                  // ls ints
    t += 3 * BP_ALIGN16(Lm);                      // ls bytes
    t += 4 * BP_ALIGN16(Lm * 8);                  // nx ny e u
    t += 11 * BP_ALIGN16(2 * Lm * 8);             // per-contact doubles
    t += BP_ALIGN16(2 * Lm * 4);                  // hash
    t += 3 * BP_ALIGN16(Lm * 4) + BP_ALIGN16((MAXCOL + 1) * 4);
    return BP_ALIGN16(t + 256);
}

#if defined(__CUDACC__) && !defined(BP_HOST_EMU)
#define BP_HD __host__ __device__ inline
#else
#define BP_HD static inline
#endif

BP_HD EnvView make_view(char* base, const Dims& d) {
    EnvView v;
    size_t NB = d.NB, P = d.PMAX, Lm = d.LMAX;
    char* p = base;
#define TAKE(ptr, type, bytes) ptr = (type*)p; p += BP_ALIGN16(bytes);
    TAKE(v.sc, EnvScal, sizeof(EnvScal));
    TAKE(v.px, double, NB * 8) TAKE(v.py, double, NB * 8) TAKE(v.pa, double, NB * 8)
    TAKE(v.vx, double, NB * 8) TAKE(v.vy, double, NB * 8) TAKE(v.vw, double, NB * 8)
    TAKE(v.bx, double, NB * 8) TAKE(v.by, double, NB * 8) TAKE(v.bw, double, NB * 8)
    TAKE(v.rc, double, NB * 8) TAKE(v.rs, double, NB * 8)
    TAKE(v.bbl, double, NB * 8) TAKE(v.bbb, double, NB * 8) TAKE(v.bbr, double, NB * 8) TAKE(v.bbt, double, NB * 8)
    TAKE(v.b0l, double, NB * 8) TAKE(v.b0b, double, NB * 8) TAKE(v.b0r, double, NB * 8) TAKE(v.b0t, double, NB * 8)
    TAKE(v.wv, double, NB * VMAX * 2 * 8)
    TAKE(v.cenx, double, NB * 8) TAKE(v.ceny, double, NB * 8) TAKE(v.area, double, NB * 8) TAKE(v.tmp, double, NB * 8)
    TAKE(v.awake, int, NB * 4)
    TAKE(v.is_awake, unsigned char, NB) TAKE(v.moved, unsigned char, NB)
    TAKE(v.moved_step, unsigned char, NB) TAKE(v.act, unsigned char, NB)
    TAKE(v.colmask, unsigned long long, NB * 8)
    TAKE(v.pkey[0], unsigned, P * 4) TAKE(v.pkey[1], unsigned, P * 4)
    TAKE(v.pgjk[0], unsigned, P * 4) TAKE(v.pgjk[1], unsigned, P * 4)
    TAKE(v.plive[0], int, P * 4) TAKE(v.plive[1], int, P * 4)
    TAKE(v.pflag[0], unsigned char, P) TAKE(v.pflag[1], unsigned char, P)
    TAKE(v.adj_start, int, (NB + 1) * 4) TAKE(v.adj, int, 2 * P * 4) TAKE(v.deg, int, NB * 4)
    TAKE(v.queue, int, P * 4) TAKE(v.apairs, int, P * 4)
    TAKE(v.livemask, unsigned, (P / 32 + 1) * 4)
    TAKE(v.ls_a, int, Lm * 4) TAKE(v.ls_b, int, Lm * 4) TAKE(v.ls_pair, int, Lm * 4)
    TAKE(v.ls_cnt, int, Lm * 4) TAKE(v.ls_stamp, int, Lm * 4)
    TAKE(v.ls_touch, unsigned char, Lm) TAKE(v.ls_first, unsigned char, Lm)
    TAKE(v.ls_keep, unsigned char, Lm)
    TAKE(v.ls_nx, double, Lm * 8) TAKE(v.ls_ny, double, Lm * 8) TAKE(v.ls_e, double, Lm * 8) TAKE(v.ls_u, double, Lm * 8)
    TAKE(v.ls_r1x, double, 2 * Lm * 8) TAKE(v.ls_r1y, double, 2 * Lm * 8)
    TAKE(v.ls_r2x, double, 2 * Lm * 8) TAKE(v.ls_r2y, double, 2 * Lm * 8)
    TAKE(v.ls_jn, double, 2 * Lm * 8) TAKE(v.ls_jt, double, 2 * Lm * 8) TAKE(v.ls_jb, double, 2 * Lm * 8)
    TAKE(v.ls_nm, double, 2 * Lm * 8) TAKE(v.ls_tm, double, 2 * Lm * 8)
    TAKE(v.ls_bias, double, 2 * Lm * 8) TAKE(v.ls_bounce, double, 2 * Lm * 8)
    TAKE(v.ls_hash, unsigned, 2 * Lm * 4)
    TAKE(v.free_stack, int, Lm * 4) TAKE(v.order, int, Lm * 4) TAKE(v.colour, int, Lm * 4)
    TAKE(v.col_start, int, (MAXCOL + 1) * 4)
#undef TAKE
    return v;
}

struct LayoutView {
    const double *lv, *ln, *rad, *minv, *iinv, *cog, *mat_e, *mat_u, *init;
    const int* nv;
    int nbody;
};
BP_DEV LayoutView layout_view(const LayoutTab& T, const Dims& d, int l) {
    LayoutView L;
    size_t o = (size_t)l * d.NB;
    L.lv = T.lv + o * VMAX * 2; L.ln = T.ln + o * VMAX * 2; L.nv = T.nv + o; L.rad = T.rad + o;
    L.minv = T.minv + o; L.iinv = T.iinv + o; L.cog = T.cog + o * 2; L.mat_e = T.mat_e + o; L.mat_u = T.mat_u + o;
    L.init = T.init + o * 3; L.nbody = T.nbody[l];
    return L;
}

// ------------------------------------------------------------------ small vector helpers
struct V2 { double x, y; };
BP_DEV V2 mk(double x, double y) { V2 r; r.x = x; r.y = y; return r; }
BP_DEV V2 operator+(V2 a, V2 b) { return mk(a.x + b.x, a.y + b.y); }
BP_DEV V2 operator-(V2 a, V2 b) { return mk(a.x - b.x, a.y - b.y); }
BP_DEV V2 operator*(V2 a, double s) { return mk(a.x * s, a.y * s); }
BP_DEV V2 neg(V2 a) { return mk(-a.x, -a.y); }
BP_DEV double dot(V2 a, V2 b) { return a.x * b.x + a.y * b.y; }
BP_DEV double cross(V2 a, V2 b) { return a.x * b.y - a.y * b.x; }
BP_DEV V2 perp(V2 a) { return mk(-a.y, a.x); }
BP_DEV V2 rperp(V2 a) { return mk(a.y, -a.x); }
BP_DEV double clampd(double f, double lo, double hi) { return fmin(fmax(f, lo), hi); }
BP_DEV double clamp01(double f) { return fmax(0.0, fmin(f, 1.0)); }
BP_DEV V2 lerp(V2 a, V2 b, double t) { return a * (1.0 - t) + b * t; }
BP_DEV V2 rotv(V2 a, V2 b) { return mk(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }

// ------------------------------------------------------------------ geometry refresh of one body
BP_DEV void refresh_body(const EnvView& ev, const LayoutView& L, int b) {
    double s, c;
    bp_sincos(ev.pa[b], &s, &c);
    ev.rc[b] = c; ev.rs[b] = s;
    double cx = L.cog[2 * b], cy = L.cog[2 * b + 1];
    double tx = ev.px[b] - (cx * c - cy * s);
    double ty = ev.py[b] - (cx * s + cy * c);
    double ms = -s;
    const double* lv = L.lv + (size_t)b * VMAX * 2;
    double* wv = ev.wv + (size_t)b * VMAX * 2;
    int n = L.nv[b];
    double l = INFINITY, r = -INFINITY, bt = INFINITY, tp = -INFINITY;
    for (int i = 0; i < n; i++) {
        double x = lv[2 * i], y = lv[2 * i + 1];
        double wx = c * x + ms * y + tx;
        double wy = s * x + c * y + ty;
        wv[2 * i] = wx; wv[2 * i + 1] = wy;
        l = fmin(l, wx); r = fmax(r, wx); bt = fmin(bt, wy); tp = fmax(tp, wy);
    }
    double rad = L.rad[b];
    ev.bbl[b] = l - rad; ev.bbb[b] = bt - rad; ev.bbr[b] = r + rad; ev.bbt[b] = tp + rad;
}
// body position (pymunk Body.position = local origin in world space)
BP_DEV V2 body_origin(const EnvView& ev, const LayoutView& L, int b) {
    double c = ev.rc[b], s = ev.rs[b];
    double cx = L.cog[2 * b], cy = L.cog[2 * b + 1];
    return mk(ev.px[b] - (cx * c - cy * s), ev.py[b] - (cx * s + cy * c));
}

// ------------------------------------------------------------------ narrowphase: GJK / EPA on world vertices
struct MPoint { V2 ab; unsigned id; };       // Minkowski difference vertex; id = (index_a << 8) | index_b
struct Poly { const double* v; int n; };

BP_DEV int support_index(const Poly& P, V2 n) {
    double mx = -INFINITY; int idx = 0;
    for (int i = 0; i < P.n; i++) {
        double d = P.v[2 * i] * n.x + P.v[2 * i + 1] * n.y;
        if (d > mx) { mx = d; idx = i; }
    }
    return idx;
}
BP_DEV V2 pvert(const Poly& P, int i) { return mk(P.v[2 * i], P.v[2 * i + 1]); }
BP_DEV MPoint mpoint(const Poly& A, const Poly& B, int ia, int ib) {
    MPoint m; m.ab = pvert(B, ib) - pvert(A, ia); m.id = (unsigned)((ia & 0xFF) << 8 | (ib & 0xFF)); return m;
}
BP_DEV MPoint msupport(const Poly& A, const Poly& B, V2 n) {
    return mpoint(A, B, support_index(A, neg(n)), support_index(B, n));
}
BP_DEV double closest_t(V2 a, V2 b) {
    V2 delta = b - a;
    return -clampd(dot(delta, a + b) / (dot(delta, delta) + DBL_MIN), -1.0, 1.0);
}
BP_DEV V2 lerp_t(V2 a, V2 b, double t) {
    double ht = 0.5 * t;
    return a * (0.5 - ht) + b * (0.5 + ht);
}
BP_DEV double closest_dist(V2 a, V2 b) { V2 p = lerp_t(a, b, closest_t(a, b)); return dot(p, p); }
BP_DEV bool point_greater(V2 a, V2 b, V2 c) {
    return (b.y - a.y) * (a.x + b.x - 2 * c.x) > (b.x - a.x) * (a.y + b.y - 2 * c.y);
}
struct Closest { V2 n; double d; unsigned id; };
BP_DEV Closest closest_points(const MPoint& v0, const MPoint& v1) {
    double t = closest_t(v0.ab, v1.ab);
    V2 p = lerp_t(v0.ab, v1.ab, t);
    V2 delta = v1.ab - v0.ab;
    V2 rp = rperp(delta);
    V2 n = rp * (1.0 / (sqrt(dot(rp, rp)) + DBL_MIN));
    double d = dot(n, p);
    Closest c; c.id = (v0.id & 0xFFFF) << 16 | (v1.id & 0xFFFF);
    if (d <= 0.0 || (-1.0 < t && t < 1.0)) { c.n = n; c.d = d; }
    else { double d2 = sqrt(dot(p, p)); c.n = p * (1.0 / (d2 + DBL_MIN)); c.d = d2; }
    return c;
}

constexpr int GJK_MAX_IT = 30, EPA_MAX_IT = 30;

BP_DEVNI Closest epa(const Poly& A, const Poly& B, MPoint h0, MPoint h1, MPoint h2, int* capped) {
    MPoint hullA[EPA_MAX_IT + 4], hullB[EPA_MAX_IT + 4];
    MPoint* hull = hullA; MPoint* hull2 = hullB;
    hull[0] = h0; hull[1] = h1; hull[2] = h2;
    int count = 3, iteration = 1;
    for (;;) {
        int mini = 0; double minDist = INFINITY;
        for (int j = 0, i = count - 1; j < count; i = j, j++) {
            double d = closest_dist(hull[i].ab, hull[j].ab);
            if (d < minDist) { minDist = d; mini = i; }
        }
        MPoint v0 = hull[mini], v1 = hull[(mini + 1) % count];
        MPoint p = msupport(A, B, perp(v1.ab - v0.ab));
        bool dup = (p.id == v0.id || p.id == v1.id);
        if (!dup && point_greater(v0.ab, v1.ab, p.ab) && iteration < EPA_MAX_IT) {
            int count2 = 1;
            hull2[0] = p;
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

BP_DEVNI Closest gjk(const Poly& A, const Poly& B, V2 cenA, V2 cenB, unsigned* id, int* iters, int* capped) {
    MPoint v0, v1;
    if (*id) {
        unsigned g = *id;
        v0 = mpoint(A, B, (g >> 24) & 0xFF, (g >> 16) & 0xFF);
        v1 = mpoint(A, B, (g >> 8) & 0xFF, g & 0xFF);
    } else {
        V2 axis = perp(cenA - cenB);
        v0 = msupport(A, B, axis);
        v1 = msupport(A, B, neg(axis));
    }
    Closest out;
    int iteration = 1;
    for (;;) {
        (*iters)++;
        if (iteration > GJK_MAX_IT) { *capped = 1; out = closest_points(v0, v1); break; }
        if (point_greater(v1.ab, v0.ab, mk(0, 0))) { MPoint t = v0; v0 = v1; v1 = t; (*iters)++; }
        double t = closest_t(v0.ab, v1.ab);
        V2 n = (-1.0 < t && t < 1.0) ? perp(v1.ab - v0.ab) : neg(lerp_t(v0.ab, v1.ab, t));
        MPoint p = msupport(A, B, n);
        if (point_greater(p.ab, v0.ab, mk(0, 0)) && point_greater(v1.ab, p.ab, mk(0, 0))) {
            out = epa(A, B, v0, p, v1, capped);
            break;
        }
        if (dot(p.ab, n) <= fmax(dot(v0.ab, n), dot(v1.ab, n))) { out = closest_points(v0, v1); break; }
        if (closest_dist(v0.ab, p.ab) < closest_dist(p.ab, v1.ab)) v1 = p; else v0 = p;
        iteration++;
    }
    *id = out.id;
    return out;
}

struct EdgeW { V2 a, b; int ia, ib; };
BP_DEV EdgeW support_edge(const Poly& P, const double* ln, double rc, double rs, V2 n) {
    int count = P.n;
    int i1 = support_index(P, n);
    int i0 = (i1 - 1 + count) % count, i2 = (i1 + 1) % count;
    double ms = -rs;
    V2 n1 = mk(rc * ln[2 * i1] + ms * ln[2 * i1 + 1], rs * ln[2 * i1] + rc * ln[2 * i1 + 1]);
    V2 n2 = mk(rc * ln[2 * i2] + ms * ln[2 * i2 + 1], rs * ln[2 * i2] + rc * ln[2 * i2 + 1]);
    EdgeW e;
    if (dot(n, n1) > dot(n, n2)) { e.a = pvert(P, i0); e.b = pvert(P, i1); e.ia = i0; e.ib = i1; }
    else { e.a = pvert(P, i1); e.b = pvert(P, i2); e.ia = i1; e.ib = i2; }
    return e;
}

struct Manifold { int count; V2 n; V2 p1[2], p2[2]; unsigned hash[2]; };

// Collide body a (first shape) with body b.  Returns contacts in world space (<= 2).
BP_DEV void collide(const EnvView& ev, const LayoutView& L, int a, int b, unsigned* gid, Manifold* m, int* iters, int* capped) {
    Poly A, B;
    A.v = ev.wv + (size_t)a * VMAX * 2; A.n = L.nv[a];
    B.v = ev.wv + (size_t)b * VMAX * 2; B.n = L.nv[b];
    V2 cenA = mk((ev.bbl[a] + ev.bbr[a]) * 0.5, (ev.bbb[a] + ev.bbt[a]) * 0.5);
    V2 cenB = mk((ev.bbl[b] + ev.bbr[b]) * 0.5, (ev.bbb[b] + ev.bbt[b]) * 0.5);
    Closest pts = gjk(A, B, cenA, cenB, gid, iters, capped);
    m->count = 0;
    double ra = L.rad[a], rb = L.rad[b];
    if (pts.d - ra - rb <= 0.0) {
        EdgeW e1 = support_edge(A, L.ln + (size_t)a * VMAX * 2, ev.rc[a], ev.rs[a], pts.n);
        EdgeW e2 = support_edge(B, L.ln + (size_t)b * VMAX * 2, ev.rc[b], ev.rs[b], neg(pts.n));
        double mindist = ra + rb;
        if (pts.d <= mindist) {
            V2 n = pts.n; m->n = n;
            double d_e1_a = cross(e1.a, n), d_e1_b = cross(e1.b, n);
            double d_e2_a = cross(e2.a, n), d_e2_b = cross(e2.b, n);
            double e1_denom = 1.0 / (d_e1_b - d_e1_a + DBL_MIN);
            double e2_denom = 1.0 / (d_e2_b - d_e2_a + DBL_MIN);
            {
                V2 p1 = n * ra + lerp(e1.a, e1.b, clamp01((d_e2_b - d_e1_a) * e1_denom));
                V2 p2 = n * (-rb) + lerp(e2.a, e2.b, clamp01((d_e1_a - d_e2_a) * e2_denom));
                if (dot(p2 - p1, n) <= 0.0) {
                    int k = m->count++;
                    m->p1[k] = p1; m->p2[k] = p2; m->hash[k] = (unsigned)(e1.ia << 8 | e2.ib);
                }
            }
            {
                V2 p1 = n * ra + lerp(e1.a, e1.b, clamp01((d_e2_a - d_e1_a) * e1_denom));
                V2 p2 = n * (-rb) + lerp(e2.a, e2.b, clamp01((d_e1_b - d_e2_a) * e2_denom));
                if (dot(p2 - p1, n) <= 0.0) {
                    int k = m->count++;
                    m->p1[k] = p1; m->p2[k] = p2; m->hash[k] = (unsigned)(e1.ib << 8 | e2.ia);
                }
            }
        }
    }
}

}  // namespace bp
