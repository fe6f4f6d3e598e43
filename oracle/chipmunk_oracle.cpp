// ORACLE — TEST INFRASTRUCTURE ONLY.  Not part of the shipped product path.
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
// legs may load this library.
//
// PARITY UNPINNED: the physics arithmetic of the reference lives in the third-party
// dependency pymunk==6.5.1 (reference requirements.txt:10), a cffi wrapper around the C
// library Chipmunk2D 7.0.3.  Neither is vendored under /root/reference nor installed,
// so this file RESTATES the published Chipmunk2D algorithm (cpSpaceStep.c, cpArbiter.c,
// cpCollision.c, cpPolyShape.c, cpBody.c, cpPolyline/cpConvexHull in chipmunk.c) in
// plain sequential C++ (all arithmetic double, as cpFloat on x86-64) restricted to the
// surface the reference's 2D environments use (call sites: ship_ice_env.py:117-120,
// 218-219,280-281; sim_utils.py:136-153; ship.py:77-98).  There are no golden vectors in
// the reference (it has no tests), so nothing pins this against real pymunk output.
//
// What is NOT restated: the BBTree broadphase.  Its pair discovery order depends on the
// tree's insertion history and is unknowable offline, so the oracle DEFINES the order:
// all shape pairs whose AABBs intersect (cpBBIntersects, the same filter
// cpSpaceCollideShapes/QueryReject applies), visited in ascending (shape_a, shape_b)
// order with shape_a < shape_b; shape_a plays the role of cpCollide's `a`.
// Optionally (cpo_space_set_order) arbiters are re-ordered by a deterministic greedy
// colouring so that a data-parallel solver can reproduce the same Gauss-Seidel order.
//
// Build: make -C oracle   (g++ only; no external dependencies)

#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cfloat>
#include <vector>
#include <algorithm>
#include <unordered_map>

namespace {

#ifdef CPO_FLOAT32
typedef float cpFloat;      // precision experiment only (tools/fp32_experiment.py): every body, contact and solver quantity in fp32
#else
typedef double cpFloat;
#endif
struct V { cpFloat x, y; };
static inline V v(cpFloat x, cpFloat y) { V r = {x, y}; return r; }
static inline V vadd(V a, V b) { return v(a.x + b.x, a.y + b.y); }
static inline V vsub(V a, V b) { return v(a.x - b.x, a.y - b.y); }
static inline V vneg(V a) { return v(-a.x, -a.y); }
static inline V vmult(V a, cpFloat s) { return v(a.x * s, a.y * s); }
static inline cpFloat vdot(V a, V b) { return a.x * b.x + a.y * b.y; }
static inline cpFloat vcross(V a, V b) { return a.x * b.y - a.y * b.x; }
static inline V vperp(V a) { return v(-a.y, a.x); }
static inline V vrperp(V a) { return v(a.y, -a.x); }
static inline V vrotate(V a, V b) { return v(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }
static inline cpFloat vlengthsq(V a) { return vdot(a, a); }
static inline cpFloat vlength(V a) { return std::sqrt(vdot(a, a)); }
static inline V vnormalize(V a) { return vmult(a, 1.0 / (vlength(a) + DBL_MIN)); }
static inline V vlerp(V a, V b, cpFloat t) { return vadd(vmult(a, 1.0 - t), vmult(b, t)); }
static inline cpFloat vdist(V a, V b) { return vlength(vsub(a, b)); }
static inline cpFloat vdistsq(V a, V b) { return vlengthsq(vsub(a, b)); }
static inline cpFloat fclamp(cpFloat f, cpFloat lo, cpFloat hi) { return std::fmin(std::fmax(f, lo), hi); }
static inline cpFloat fclamp01(cpFloat f) { return std::fmax(0.0, std::fmin(f, 1.0)); }

enum { BODY_DYNAMIC = 0, BODY_KINEMATIC = 1, BODY_STATIC = 2 };
enum { ARB_FIRST = 0, ARB_NORMAL = 1, ARB_IGNORE = 2, ARB_CACHED = 3 };

struct Transform { cpFloat a, b, c, d, tx, ty; };
static inline V tpoint(const Transform& t, V p) { return v(t.a * p.x + t.c * p.y + t.tx, t.b * p.x + t.d * p.y + t.ty); }
static inline V tvect(const Transform& t, V p) { return v(t.a * p.x + t.c * p.y, t.b * p.x + t.d * p.y); }

struct Body {
    int type;
    cpFloat m, i, m_inv, i_inv;
    V cog;       // local centre of gravity
    V p;         // world position of the cog
    V vel; cpFloat a, w;
    V v_bias; cpFloat w_bias;
    V f; cpFloat t;
    Transform tr;
    std::vector<int> shapes;
    bool removed = false;   // space.remove(body, shape): box_delivery_env.py:765-767
};

struct Plane { V v0, n; };

struct Shape {
    int body;
    int kind;    // 0 = cpPolyShape, 1 = cpSegmentShape (world[0].v0 = ta, world[1].v0 = tb, world[0].n = tn)
    int count;
    cpFloat r;
    std::vector<Plane> local;   // hull verts (CCW) + edge normals, body-local
    std::vector<Plane> world;   // cached world-space copy (cpPolyShapeCacheData)
    cpFloat bb_l, bb_b, bb_r, bb_t;
    cpFloat e, u;
    int ctype;
    // cpShapeMassInfo
    cpFloat mi_m, mi_i, mi_area; V mi_cog;
};

struct Contact {
    V r1, r2;
    cpFloat nMass, tMass, bounce;
    cpFloat jnAcc, jtAcc, jBias, bias;
    uint64_t hash;
};

struct Arbiter {
    int a, b;                // shape ids (a < b)
    int body_a, body_b;
    cpFloat e, u;
    V n;
    int count;
    Contact con[2];
    int state;
    uint64_t stamp;
    uint32_t gjk_id;         // cpCollisionID warm-start hint for GJK
    bool live;               // in this step's solve list
};

// ---------------------------------------------------------------- convex hull (chipmunk.c: cpConvexHull / QHullReduce)
static void loop_indexes(const V* verts, int count, int* start, int* end) {
    *start = *end = 0;
    V mn = verts[0], mx = mn;
    for (int i = 1; i < count; i++) {
        V p = verts[i];
        if (p.x < mn.x || (p.x == mn.x && p.y < mn.y)) { mn = p; *start = i; }
        else if (p.x > mx.x || (p.x == mx.x && p.y > mx.y)) { mx = p; *end = i; }
    }
}
static int qhull_partition(V* verts, int count, V a, V b, cpFloat tol) {
    if (count == 0) return 0;
    cpFloat mx = 0; int pivot = 0;
    V delta = vsub(b, a);
    cpFloat valueTol = tol * vlength(delta);
    int head = 0;
    for (int tail = count - 1; head <= tail;) {
        cpFloat value = vcross(vsub(verts[head], a), delta);
        if (value > valueTol) {
            if (value > mx) { mx = value; pivot = head; }
            head++;
        } else {
            std::swap(verts[head], verts[tail]);
            tail--;
        }
    }
    if (pivot != 0) std::swap(verts[0], verts[pivot]);
    return head;
}
static int qhull_reduce(cpFloat tol, V* verts, int count, V a, V pivot, V b, V* result) {
    if (count < 0) return 0;
    if (count == 0) { result[0] = pivot; return 1; }
    int left = qhull_partition(verts, count, a, pivot, tol);
    int index = qhull_reduce(tol, verts + 1, left - 1, a, verts[0], pivot, result);
    result[index++] = pivot;
    verts += left; count -= left;
    int right = qhull_partition(verts, count, pivot, b, tol);
    return index + qhull_reduce(tol, verts + 1, right - 1, pivot, verts[0], b, result + index);
}
static int convex_hull(int count, V* result /* in-place */, cpFloat tol) {
    int start, end;
    loop_indexes(result, count, &start, &end);
    if (start == end) return 1;
    std::swap(result[0], result[start]);
    std::swap(result[1], result[end == 0 ? start : end]);
    V a = result[0], b = result[1];
    return qhull_reduce(tol, result + 2, count - 2, a, b, a, result + 1) + 1;
}

// ---------------------------------------------------------------- mass properties (chipmunk.c)
static cpFloat area_for_poly(int count, const V* verts, cpFloat r) {
    cpFloat area = 0, perimeter = 0;
    for (int i = 0; i < count; i++) {
        V v1 = verts[i], v2 = verts[(i + 1) % count];
        area += vcross(v1, v2);
        perimeter += vdist(v1, v2);
    }
    return r * (M_PI * std::fabs(r) + perimeter) + area / 2.0;
}
static V centroid_for_poly(int count, const V* verts) {
    cpFloat sum = 0; V vsum = v(0, 0);
    for (int i = 0; i < count; i++) {
        V v1 = verts[i], v2 = verts[(i + 1) % count];
        cpFloat cr = vcross(v1, v2);
        sum += cr;
        vsum = vadd(vsum, vmult(vadd(v1, v2), cr));
    }
    return vmult(vsum, 1.0 / (3.0 * sum));
}
static cpFloat moment_for_segment(cpFloat m, V a, V b, cpFloat r) {
    V offset = vlerp(a, b, 0.5);
    cpFloat length = vdist(b, a) + 2.0 * r;
    return m * ((length * length + 4.0 * r * r) / 12.0 + vlengthsq(offset));
}
static cpFloat moment_for_poly(cpFloat m, int count, const V* verts, V offset, cpFloat /*r*/) {
    if (count == 2) return moment_for_segment(m, verts[0], verts[1], 0.0);
    cpFloat sum1 = 0, sum2 = 0;
    for (int i = 0; i < count; i++) {
        V v1 = vadd(verts[i], offset), v2 = vadd(verts[(i + 1) % count], offset);
        cpFloat a = vcross(v2, v1);
        cpFloat b = vdot(v1, v1) + vdot(v1, v2) + vdot(v2, v2);
        sum1 += a * b; sum2 += a;
    }
    return (m * sum1) / (6.0 * sum2);
}

// ---------------------------------------------------------------- space
struct Stats {
    long long steps, pairs_bb, pairs_touch, arb_solved, arb_nonzero, gjk_iters, epa_calls;
    long long max_arb, max_nonzero, moved_bodies, max_moved;
};

struct Space {
    int iterations; cpFloat damping; V gravity;
    cpFloat slop, collision_bias; uint64_t persistence;
    uint64_t stamp; cpFloat curr_dt;
    int order_mode;
    std::vector<Body> bodies;
    std::vector<Shape> shapes;
    std::unordered_map<uint64_t, Arbiter> cache;   // keyed by (a<<32)|b
    std::vector<Arbiter*> arbiters;                // this step's solve list
    Stats st;
    // collision-handler bookkeeping: number of pre_solve calls this far for shape pairs whose
    // collision types are {1, 3} (robot chassis vs maze wall: maze_NAMO_env.py:203-216)
    long long hits_13;
    // handler_mode 1 = box-delivery / area-clearing handler set (box_delivery_env.py:196-229):
    //   (1,3) and (2,3) pre_solve = prevent_boundary_intersection (:294-311) on shapes[0] (the type-1 / type-2
    //   shape), (1,4) begin -> False (ignored), (2,4) pre_solve -> False (never solved).
    // robot_hit = the value the LAST (1,3) pre_solve assigned to env.robot_hit_obstacle (:201-203), -1 = no call yet.
    int handler_mode = 0;
    int robot_hit = -1;
    // sweep scratch
    std::vector<int> order_y;
};

// Trigonometry.  Chipmunk calls libm (cpvforangle = cos/sin), and that is the default here.  libm
// implementations differ in the last bit (glibc vs CUDA vs others) and the contact dynamics amplify
// that over an episode, so for WHOLE-EPISODE comparisons the tests can switch the oracle to a
// portable routine (cpo_set_portable_trig): Cody-Waite reduction by pi/2 + the fdlibm k_sin / k_cos
// polynomials (published algorithm: FreeBSD msun e_rem_pio2.c medium path, k_sin.c, k_cos.c),
// evaluated without fused multiply-add, which every IEEE-754 machine reproduces bit for bit.
static int g_portable_trig = 0;
static void portable_sincos(double x, double* sn, double* cs) {
    static const double invpio2 = 6.36619772367581382433e-01, pio2_1 = 1.57079632673412561417e+00,
                        pio2_2 = 6.07710050630396597660e-11, pio2_2t = 2.02226624879595063154e-21;
    static const double S[7] = {0, -1.66666666666666324348e-01, 8.33333333332248946124e-03, -1.98412698298579493134e-04,
                                2.75573137070700676789e-06, -2.50507602534068634195e-08, 1.58969099521155010221e-10};
    static const double Cc[7] = {0, 4.16666666666666019037e-02, -1.38888888888741095749e-03, 2.48015872894767294178e-05,
                                 -2.75573143513906633035e-07, 2.08757232129817482790e-09, -1.13596475577881948265e-11};
    double fn = std::rint(x * invpio2);
    double r0 = x - fn * pio2_1;
    double w = fn * pio2_2;
    double r = r0 - w;
    w = fn * pio2_2t - ((r0 - r) - w);
    double y = r - w, ytail = (r - y) - w;
    double z = y * y, v = z * y;
    double ps = S[2] + z * (S[3] + z * (S[4] + z * (S[5] + z * S[6])));
    double ksin = y - ((z * (0.5 * ytail - v * ps) - ytail) - v * S[1]);
    double pc = z * (Cc[1] + z * (Cc[2] + z * (Cc[3] + z * (Cc[4] + z * (Cc[5] + z * Cc[6])))));
    double kcos, ay = std::fabs(y);
    if (ay < 0.3) kcos = 1.0 - (0.5 * z - (z * pc - y * ytail));
    else {
        double qx;
        if (ay > 0.78125) qx = 0.28125;
        else { uint64_t bits; memcpy(&bits, &ay, 8); bits = (bits - 0x0020000000000000ull) & 0xFFFFFFFF00000000ull; memcpy(&qx, &bits, 8); }
        double hz = 0.5 * z - qx, a = 1.0 - qx;
        kcos = a - (hz - (z * pc - y * ytail));
    }
    switch ((int)fn & 3) {
        case 0: *sn = ksin; *cs = kcos; break;
        case 1: *sn = kcos; *cs = -ksin; break;
        case 2: *sn = -ksin; *cs = -kcos; break;
        default: *sn = -kcos; *cs = ksin; break;
    }
}
static void trig(double a, double* sn, double* cs) {
    if (g_portable_trig) portable_sincos(a, sn, cs);
    else { *cs = std::cos(a); *sn = std::sin(a); }
}

static void set_transform(Body& b) {
    double cs_d, sn_d;
    trig((double)b.a, &sn_d, &cs_d);      // cpvforangle
    cpFloat cs = (cpFloat)cs_d, sn = (cpFloat)sn_d;
    V c = b.cog;
    b.tr.a = cs; b.tr.b = sn; b.tr.c = -sn; b.tr.d = cs;
    b.tr.tx = b.p.x - (c.x * cs - c.y * sn);
    b.tr.ty = b.p.y - (c.x * sn + c.y * cs);
}
static V body_get_position(const Body& b) { return tpoint(b.tr, v(0, 0)); }
static void body_set_position(Body& b, V pos) {
    b.p = vadd(tvect(b.tr, b.cog), pos);
    set_transform(b);
}
static void body_accumulate_mass(Space* sp, Body& b) {   // cpBodyAccumulateMassFromShapes
    if (b.type != BODY_DYNAMIC) return;
    b.m = b.i = 0; b.cog = v(0, 0);
    V pos = body_get_position(b);
    for (int sid : b.shapes) {
        Shape& s = sp->shapes[sid];
        cpFloat m = s.mi_m;
        if (m > 0) {
            cpFloat msum = b.m + m;
            b.i += m * s.mi_i + vdistsq(b.cog, s.mi_cog) * (m * b.m) / msum;
            b.cog = vlerp(b.cog, s.mi_cog, m / msum);
            b.m = msum;
        }
    }
    b.m_inv = 1.0 / b.m; b.i_inv = 1.0 / b.i;
    body_set_position(b, pos);
}
static void shape_cache_data(Space* sp, Shape& s) {      // cpPolyShapeCacheData
    const Transform& t = sp->bodies[s.body].tr;
    cpFloat l = INFINITY, r = -INFINITY, b = INFINITY, tp = -INFINITY;
    for (int i = 0; i < s.count; i++) {
        V p = tpoint(t, s.local[i].v0);
        V n = tvect(t, s.local[i].n);
        s.world[i].v0 = p; s.world[i].n = n;
        l = std::fmin(l, p.x); r = std::fmax(r, p.x);
        b = std::fmin(b, p.y); tp = std::fmax(tp, p.y);
    }
    s.bb_l = l - s.r; s.bb_b = b - s.r; s.bb_r = r + s.r; s.bb_t = tp + s.r;
}

// ---------------------------------------------------------------- narrowphase (cpCollision.c)
struct SupportPoint { V p; int index; };
struct MinkowskiPoint { V a, b, ab; uint32_t id; };
struct ClosestPoints { V a, b, n; cpFloat d; uint32_t id; };
struct Ctx { const Shape* s1; const Shape* s2; Space* sp; };

static inline int poly_support_index(const Shape* s, V n) {
    cpFloat mx = -INFINITY; int index = 0;
    for (int i = 0; i < s->count; i++) {
        cpFloat d = vdot(s->world[i].v0, n);
        if (d > mx) { mx = d; index = i; }
    }
    return index;
}
static inline SupportPoint poly_support(const Shape* s, V n) {
    int i = poly_support_index(s, n);
    SupportPoint sp = {s->world[i].v0, i};
    return sp;
}
static inline MinkowskiPoint mk_new(SupportPoint a, SupportPoint b) {
    MinkowskiPoint m = {a.p, b.p, vsub(b.p, a.p), (uint32_t)((a.index & 0xFF) << 8 | (b.index & 0xFF))};
    return m;
}
// cpCollision.c SegmentSupportPoint: ta (index 0) only when strictly greater, else tb (index 1)
static inline SupportPoint segment_support(const Shape* s, V n) {
    if (vdot(s->world[0].v0, n) > vdot(s->world[1].v0, n)) { SupportPoint sp = {s->world[0].v0, 0}; return sp; }
    SupportPoint sp = {s->world[1].v0, 1}; return sp;
}
static inline SupportPoint shape_support(const Shape* s, V n) { return s->kind == 1 ? segment_support(s, n) : poly_support(s, n); }
static inline MinkowskiPoint support(const Ctx* c, V n) {
    return mk_new(shape_support(c->s1, vneg(n)), shape_support(c->s2, n));
}
static inline cpFloat closest_t(V a, V b) {
    V delta = vsub(b, a);
    return -fclamp(vdot(delta, vadd(a, b)) / (vlengthsq(delta) + DBL_MIN), -1.0, 1.0);
}
static inline V lerp_t(V a, V b, cpFloat t) {
    cpFloat ht = 0.5 * t;
    return vadd(vmult(a, 0.5 - ht), vmult(b, 0.5 + ht));
}
static inline cpFloat closest_dist(V v0, V v1) { return vlengthsq(lerp_t(v0, v1, closest_t(v0, v1))); }
static inline bool check_point_greater(V a, V b, V c) {
    return (b.y - a.y) * (a.x + b.x - 2 * c.x) > (b.x - a.x) * (a.y + b.y - 2 * c.y);
}
static inline bool check_axis(V v0, V v1, V p, V n) { return vdot(p, n) <= std::fmax(vdot(v0, n), vdot(v1, n)); }

static ClosestPoints closest_points_new(MinkowskiPoint v0, MinkowskiPoint v1) {
    cpFloat t = closest_t(v0.ab, v1.ab);
    V p = lerp_t(v0.ab, v1.ab, t);
    V pa = lerp_t(v0.a, v1.a, t), pb = lerp_t(v0.b, v1.b, t);
    uint32_t id = (v0.id & 0xFFFF) << 16 | (v1.id & 0xFFFF);
    V delta = vsub(v1.ab, v0.ab);
    V n = vnormalize(vrperp(delta));
    cpFloat d = vdot(n, p);
    if (d <= 0.0 || (-1.0 < t && t < 1.0)) {
        ClosestPoints cp = {pa, pb, n, d, id};
        return cp;
    } else {
        cpFloat d2 = vlength(p);
        V n2 = vmult(p, 1.0 / (d2 + DBL_MIN));
        ClosestPoints cp = {pa, pb, n2, d2, id};
        return cp;
    }
}

static const int MAX_GJK_ITERATIONS = 30, MAX_EPA_ITERATIONS = 30;

static ClosestPoints epa_recurse(const Ctx* ctx, int count, const MinkowskiPoint* hull, int iteration) {
    int mini = 0; cpFloat minDist = INFINITY;
    for (int j = 0, i = count - 1; j < count; i = j, j++) {
        cpFloat d = closest_dist(hull[i].ab, hull[j].ab);
        if (d < minDist) { minDist = d; mini = i; }
    }
    MinkowskiPoint v0 = hull[mini], v1 = hull[(mini + 1) % count];
    MinkowskiPoint p = support(ctx, vperp(vsub(v1.ab, v0.ab)));
    bool duplicate = (p.id == v0.id || p.id == v1.id);
    if (!duplicate && check_point_greater(v0.ab, v1.ab, p.ab) && iteration < MAX_EPA_ITERATIONS) {
        std::vector<MinkowskiPoint> hull2(count + 1);
        int count2 = 1;
        hull2[0] = p;
        for (int i = 0; i < count; i++) {
            int index = (mini + 1 + i) % count;
            V h0 = hull2[count2 - 1].ab;
            V h1 = hull[index].ab;
            V h2 = (i + 1 < count ? hull[(index + 1) % count] : p).ab;
            if (check_point_greater(h0, h2, h1)) { hull2[count2] = hull[index]; count2++; }
        }
        return epa_recurse(ctx, count2, hull2.data(), iteration + 1);
    }
    return closest_points_new(v0, v1);
}
static ClosestPoints epa(const Ctx* ctx, MinkowskiPoint v0, MinkowskiPoint v1, MinkowskiPoint v2) {
    MinkowskiPoint hull[3] = {v0, v1, v2};
    ctx->sp->st.epa_calls++;
    return epa_recurse(ctx, 3, hull, 1);
}
static ClosestPoints gjk_recurse(const Ctx* ctx, MinkowskiPoint v0, MinkowskiPoint v1, int iteration) {
    ctx->sp->st.gjk_iters++;
    if (iteration > MAX_GJK_ITERATIONS) return closest_points_new(v0, v1);
    if (check_point_greater(v1.ab, v0.ab, v(0, 0))) {
        return gjk_recurse(ctx, v1, v0, iteration);
    } else {
        cpFloat t = closest_t(v0.ab, v1.ab);
        V n = (-1.0 < t && t < 1.0 ? vperp(vsub(v1.ab, v0.ab)) : vneg(lerp_t(v0.ab, v1.ab, t)));
        MinkowskiPoint p = support(ctx, n);
        if (check_point_greater(p.ab, v0.ab, v(0, 0)) && check_point_greater(v1.ab, p.ab, v(0, 0))) {
            return epa(ctx, v0, p, v1);
        } else {
            if (check_axis(v0.ab, v1.ab, p.ab, n)) {
                return closest_points_new(v0, v1);
            } else {
                if (closest_dist(v0.ab, p.ab) < closest_dist(p.ab, v1.ab)) return gjk_recurse(ctx, v0, p, iteration + 1);
                else return gjk_recurse(ctx, p, v1, iteration + 1);
            }
        }
    }
}
static ClosestPoints gjk(const Ctx* ctx, uint32_t* id) {
    MinkowskiPoint v0, v1;
    if (*id) {
        const Shape* s1 = ctx->s1; const Shape* s2 = ctx->s2;
        int i0 = (*id >> 24) & 0xFF, j0 = (*id >> 16) & 0xFF, i1 = (*id >> 8) & 0xFF, j1 = (*id) & 0xFF;
        SupportPoint a0 = {s1->world[i0].v0, i0}, b0 = {s2->world[j0].v0, j0};
        SupportPoint a1 = {s1->world[i1].v0, i1}, b1 = {s2->world[j1].v0, j1};
        v0 = mk_new(a0, b0); v1 = mk_new(a1, b1);
    } else {
        V c1 = v((ctx->s1->bb_l + ctx->s1->bb_r) * 0.5, (ctx->s1->bb_b + ctx->s1->bb_t) * 0.5);
        V c2 = v((ctx->s2->bb_l + ctx->s2->bb_r) * 0.5, (ctx->s2->bb_b + ctx->s2->bb_t) * 0.5);
        V axis = vperp(vsub(c1, c2));
        v0 = support(ctx, axis);
        v1 = support(ctx, vneg(axis));
    }
    ClosestPoints pts = gjk_recurse(ctx, v0, v1, 1);
    *id = pts.id;
    return pts;
}

struct EdgePoint { V p; uint64_t hash; };
struct Edge { EdgePoint a, b; cpFloat r; V n; };
static inline uint64_t feature_id(int shape, int idx) { return ((uint64_t)(uint32_t)shape << 8) | (uint32_t)idx; }

static Edge support_edge_for_poly(const Shape* s, int sid, V n) {
    int count = s->count;
    int i1 = poly_support_index(s, n);
    int i0 = (i1 - 1 + count) % count, i2 = (i1 + 1) % count;
    const std::vector<Plane>& pl = s->world;
    if (vdot(n, pl[i1].n) > vdot(n, pl[i2].n)) {
        Edge e = {{pl[i0].v0, feature_id(sid, i0)}, {pl[i1].v0, feature_id(sid, i1)}, s->r, pl[i1].n};
        return e;
    } else {
        Edge e = {{pl[i1].v0, feature_id(sid, i1)}, {pl[i2].v0, feature_id(sid, i2)}, s->r, pl[i2].n};
        return e;
    }
}

// cpCollision.c SupportEdgeForSegment
static Edge support_edge_for_segment(const Shape* s, int sid, V n) {
    V ta = s->world[0].v0, tb = s->world[1].v0, tn = s->world[0].n;
    if (vdot(tn, n) > 0.0) {
        Edge e = {{ta, feature_id(sid, 0)}, {tb, feature_id(sid, 1)}, s->r, tn};
        return e;
    } else {
        Edge e = {{tb, feature_id(sid, 1)}, {ta, feature_id(sid, 0)}, s->r, vneg(tn)};
        return e;
    }
}

struct CollisionInfo { V n; int count; V p1[2], p2[2]; uint64_t hash[2]; };

// Chipmunk hashes the two vertex hashes with the symmetric CP_HASH_PAIR; here the pair of
// exact feature ids is packed instead (no false matches; a < b is fixed so no symmetry needed).
static inline uint64_t pair_hash(uint64_t ha, uint64_t hb) { return (ha << 32) | hb; }

static void contact_points(const Edge& e1, const Edge& e2, const ClosestPoints& pts, CollisionInfo* info) {
    cpFloat mindist = e1.r + e2.r;
    if (pts.d <= mindist) {
        V n = info->n = pts.n;
        cpFloat d_e1_a = vcross(e1.a.p, n), d_e1_b = vcross(e1.b.p, n);
        cpFloat d_e2_a = vcross(e2.a.p, n), d_e2_b = vcross(e2.b.p, n);
        cpFloat e1_denom = 1.0 / (d_e1_b - d_e1_a + DBL_MIN);
        cpFloat e2_denom = 1.0 / (d_e2_b - d_e2_a + DBL_MIN);
        {
            V p1 = vadd(vmult(n, e1.r), vlerp(e1.a.p, e1.b.p, fclamp01((d_e2_b - d_e1_a) * e1_denom)));
            V p2 = vadd(vmult(n, -e2.r), vlerp(e2.a.p, e2.b.p, fclamp01((d_e1_a - d_e2_a) * e2_denom)));
            cpFloat dist = vdot(vsub(p2, p1), n);
            if (dist <= 0.0) {
                int k = info->count++;
                info->p1[k] = p1; info->p2[k] = p2; info->hash[k] = pair_hash(e1.a.hash, e2.b.hash);
            }
        }
        {
            V p1 = vadd(vmult(n, e1.r), vlerp(e1.a.p, e1.b.p, fclamp01((d_e2_a - d_e1_a) * e1_denom)));
            V p2 = vadd(vmult(n, -e2.r), vlerp(e2.a.p, e2.b.p, fclamp01((d_e1_b - d_e2_a) * e2_denom)));
            cpFloat dist = vdot(vsub(p2, p1), n);
            if (dist <= 0.0) {
                int k = info->count++;
                info->p1[k] = p1; info->p2[k] = p2; info->hash[k] = pair_hash(e1.b.hash, e2.a.hash);
            }
        }
    }
}

static void poly_to_poly(Space* sp, int ia, int ib, uint32_t* id, CollisionInfo* info) {
    const Shape* s1 = &sp->shapes[ia]; const Shape* s2 = &sp->shapes[ib];
    Ctx ctx = {s1, s2, sp};
    ClosestPoints pts = gjk(&ctx, id);
    if (pts.d - s1->r - s2->r <= 0.0) {
        contact_points(support_edge_for_poly(s1, ia, pts.n), support_edge_for_poly(s2, ib, vneg(pts.n)), pts, info);
    }
}

// cpCollision.c SegmentToPoly (the segment is always shape `a`: cpCollide requires type order
// circle < segment < poly).  Neighbour tangents are zero for every segment the envs create
// (sim_utils.py:174-181 never calls set_neighbors), so the end-cap rejection never fires.
static void segment_to_poly(Space* sp, int ia, int ib, uint32_t* id, CollisionInfo* info) {
    const Shape* seg = &sp->shapes[ia]; const Shape* poly = &sp->shapes[ib];
    Ctx ctx = {seg, poly, sp};
    ClosestPoints pts = gjk(&ctx, id);
    if (pts.d - seg->r - poly->r <= 0.0) {
        contact_points(support_edge_for_segment(seg, ia, pts.n), support_edge_for_poly(poly, ib, vneg(pts.n)), pts, info);
    }
}
static void collide_shapes(Space* sp, int ia, int ib, uint32_t* id, CollisionInfo* info) {
    int ka = sp->shapes[ia].kind, kb = sp->shapes[ib].kind;
    if (ka == 0 && kb == 0) poly_to_poly(sp, ia, ib, id, info);
    else if (ka == 1 && kb == 0) segment_to_poly(sp, ia, ib, id, info);
    else if (ka == 1 && kb == 1) { /* cpCollide: segment-segment is only reached for non-static pairs; none exist here */ }
    else { fprintf(stderr, "oracle: create segment shapes before poly shapes (pair %d,%d)\n", ia, ib); abort(); }
}

// ---------------------------------------------------------------- arbiter (cpArbiter.c)
static inline V relative_velocity(const Body& a, const Body& b, V r1, V r2) {
    V v1 = vadd(b.vel, vmult(vperp(r2), b.w));
    V v2 = vadd(a.vel, vmult(vperp(r1), a.w));
    return vsub(v1, v2);
}
static inline cpFloat k_scalar_body(const Body& b, V r, V n) {
    cpFloat rcn = vcross(r, n);
    return b.m_inv + b.i_inv * rcn * rcn;
}
static inline cpFloat k_scalar(const Body& a, const Body& b, V r1, V r2, V n) {
    return k_scalar_body(a, r1, n) + k_scalar_body(b, r2, n);
}
static inline void apply_impulse(Body& b, V j, V r) {
    b.vel = vadd(b.vel, vmult(j, b.m_inv));
    b.w += b.i_inv * vcross(r, j);
}
static inline void apply_impulses(Body& a, Body& b, V r1, V r2, V j) {
    apply_impulse(a, vneg(j), r1); apply_impulse(b, j, r2);
}
static inline void apply_bias_impulse(Body& b, V j, V r) {
    b.v_bias = vadd(b.v_bias, vmult(j, b.m_inv));
    b.w_bias += b.i_inv * vcross(r, j);
}
static inline void apply_bias_impulses(Body& a, Body& b, V r1, V r2, V j) {
    apply_bias_impulse(a, vneg(j), r1); apply_bias_impulse(b, j, r2);
}

static void arbiter_update(Space* sp, Arbiter* arb, const CollisionInfo& info) {   // cpArbiterUpdate
    const Shape& a = sp->shapes[arb->a]; const Shape& b = sp->shapes[arb->b];
    Contact fresh[2];
    for (int i = 0; i < info.count; i++) {
        Contact& c = fresh[i];
        c.r1 = vsub(info.p1[i], sp->bodies[a.body].p);
        c.r2 = vsub(info.p2[i], sp->bodies[b.body].p);
        c.jnAcc = c.jtAcc = 0.0; c.jBias = 0.0; c.bias = 0.0; c.nMass = c.tMass = c.bounce = 0.0;
        c.hash = info.hash[i];
        for (int j = 0; j < arb->count; j++) {
            if (c.hash == arb->con[j].hash) { c.jnAcc = arb->con[j].jnAcc; c.jtAcc = arb->con[j].jtAcc; }
        }
    }
    for (int i = 0; i < info.count; i++) arb->con[i] = fresh[i];
    arb->count = info.count;
    arb->n = info.n;
    arb->e = a.e * b.e;
    arb->u = a.u * b.u;
    if (arb->state == ARB_CACHED) arb->state = ARB_FIRST;
}
static void arbiter_prestep(Space* sp, Arbiter* arb, cpFloat dt, cpFloat slop, cpFloat bias) {
    const Body& a = sp->bodies[arb->body_a]; const Body& b = sp->bodies[arb->body_b];
    V n = arb->n;
    V body_delta = vsub(b.p, a.p);
    for (int i = 0; i < arb->count; i++) {
        Contact& c = arb->con[i];
        c.nMass = 1.0 / k_scalar(a, b, c.r1, c.r2, n);
        c.tMass = 1.0 / k_scalar(a, b, c.r1, c.r2, vperp(n));
        cpFloat dist = vdot(vadd(vsub(c.r2, c.r1), body_delta), n);
        c.bias = -bias * std::fmin(0.0, dist + slop) / dt;
        c.jBias = 0.0;
        c.bounce = vdot(relative_velocity(a, b, c.r1, c.r2), n) * arb->e;
    }
}
static void arbiter_apply_cached(Space* sp, Arbiter* arb, cpFloat dt_coef) {
    if (arb->state == ARB_FIRST) return;
    Body& a = sp->bodies[arb->body_a]; Body& b = sp->bodies[arb->body_b];
    V n = arb->n;
    for (int i = 0; i < arb->count; i++) {
        Contact& c = arb->con[i];
        V j = vrotate(n, v(c.jnAcc, c.jtAcc));
        apply_impulses(a, b, c.r1, c.r2, vmult(j, dt_coef));
    }
}
static void arbiter_apply_impulse(Space* sp, Arbiter* arb) {
    Body& a = sp->bodies[arb->body_a]; Body& b = sp->bodies[arb->body_b];
    V n = arb->n;
    cpFloat friction = arb->u;
    for (int i = 0; i < arb->count; i++) {
        Contact& c = arb->con[i];
        cpFloat nMass = c.nMass;
        V r1 = c.r1, r2 = c.r2;
        V vb1 = vadd(a.v_bias, vmult(vperp(r1), a.w_bias));
        V vb2 = vadd(b.v_bias, vmult(vperp(r2), b.w_bias));
        V vr = relative_velocity(a, b, r1, r2);          // surface_vr == 0 for every shape the envs create
        cpFloat vbn = vdot(vsub(vb2, vb1), n);
        cpFloat vrn = vdot(vr, n);
        cpFloat vrt = vdot(vr, vperp(n));
        cpFloat jbn = (c.bias - vbn) * nMass;
        cpFloat jbnOld = c.jBias;
        c.jBias = std::fmax(jbnOld + jbn, 0.0);
        cpFloat jn = -(c.bounce + vrn) * nMass;
        cpFloat jnOld = c.jnAcc;
        c.jnAcc = std::fmax(jnOld + jn, 0.0);
        cpFloat jtMax = friction * c.jnAcc;
        cpFloat jt = -vrt * c.tMass;
        cpFloat jtOld = c.jtAcc;
        c.jtAcc = fclamp(jtOld + jt, -jtMax, jtMax);
        apply_bias_impulses(a, b, r1, r2, vmult(n, c.jBias - jbnOld));
        apply_impulses(a, b, r1, r2, vrotate(n, v(c.jnAcc - jnOld, c.jtAcc - jtOld)));
    }
}

// ---------------------------------------------------------------- optional solver order (greedy colouring)
// order_mode 0: canonical ascending (a,b).  order_mode 1: stable re-order of the canonical list
// by greedy colour, colour(arb) = smallest c >= 0 not yet used by an earlier (canonical order)
// arbiter that shares a DYNAMIC body with it.  Same-colour arbiters touch disjoint dynamic
// bodies, so a parallel solver that sweeps colour by colour performs exactly this sequence.
static void reorder_by_colour(Space* sp) {
    size_t n = sp->arbiters.size();
    std::vector<std::vector<uint64_t>> used(sp->bodies.size());   // bitmask words per body
    std::vector<int> colour(n);
    int maxc = 0;
    for (size_t k = 0; k < n; k++) {
        Arbiter* arb = sp->arbiters[k];
        int ba = arb->body_a, bb = arb->body_b;
        bool da = sp->bodies[ba].type == BODY_DYNAMIC, db = sp->bodies[bb].type == BODY_DYNAMIC;
        int c = 0;
        for (;; c++) {
            size_t wd = c >> 6; uint64_t bit = 1ull << (c & 63);
            bool ua = da && wd < used[ba].size() && (used[ba][wd] & bit);
            bool ub = db && wd < used[bb].size() && (used[bb][wd] & bit);
            if (!ua && !ub) break;
        }
        size_t wd = c >> 6; uint64_t bit = 1ull << (c & 63);
        if (da) { if (used[ba].size() <= wd) used[ba].resize(wd + 1, 0); used[ba][wd] |= bit; }
        if (db) { if (used[bb].size() <= wd) used[bb].resize(wd + 1, 0); used[bb][wd] |= bit; }
        colour[k] = c; if (c > maxc) maxc = c;
    }
    std::vector<Arbiter*> out; out.reserve(n);
    for (int c = 0; c <= maxc; c++)
        for (size_t k = 0; k < n; k++) if (colour[k] == c) out.push_back(sp->arbiters[k]);
    sp->arbiters.swap(out);
}

// ---------------------------------------------------------------- cpSpaceStep
static void space_step(Space* sp, cpFloat dt) {
    if (dt == 0.0) return;
    sp->stamp++;
    cpFloat prev_dt = sp->curr_dt;
    sp->curr_dt = dt;

    for (Arbiter* arb : sp->arbiters) { arb->state = ARB_NORMAL; arb->live = false; }
    sp->arbiters.clear();

    // integrate positions (dynamic + kinematic)
    long long moved = 0;
    for (Body& b : sp->bodies) {
        if (b.type == BODY_STATIC || b.removed) continue;
        if (b.vel.x != 0 || b.vel.y != 0 || b.w != 0 || b.v_bias.x != 0 || b.v_bias.y != 0 || b.w_bias != 0) moved++;
        b.p = vadd(b.p, vmult(vadd(b.vel, b.v_bias), dt));
        b.a = b.a + (b.w + b.w_bias) * dt;
        set_transform(b);
        b.v_bias = v(0, 0); b.w_bias = 0.0;
    }
    sp->st.moved_bodies += moved; if (moved > sp->st.max_moved) sp->st.max_moved = moved;
    for (Shape& s : sp->shapes) if (sp->bodies[s.body].type != BODY_STATIC && !sp->bodies[s.body].removed) shape_cache_data(sp, s);

    // broadphase: sweep along y, then canonical order
    int ns = (int)sp->shapes.size();
    sp->order_y.resize(ns);
    for (int i = 0; i < ns; i++) sp->order_y[i] = i;
    std::sort(sp->order_y.begin(), sp->order_y.end(), [&](int p, int q) {
        const Shape& a = sp->shapes[p]; const Shape& b = sp->shapes[q];
        return a.bb_b < b.bb_b || (a.bb_b == b.bb_b && p < q); });
    std::vector<uint64_t> pairs;
    for (int ii = 0; ii < ns; ii++) {
        const Shape& a = sp->shapes[sp->order_y[ii]];
        for (int jj = ii + 1; jj < ns; jj++) {
            const Shape& b = sp->shapes[sp->order_y[jj]];
            if (b.bb_b > a.bb_t) break;
            if (!(a.bb_l <= b.bb_r && b.bb_l <= a.bb_r && a.bb_b <= b.bb_t && b.bb_b <= a.bb_t)) continue;
            if (a.body == b.body) continue;
            if (sp->bodies[a.body].removed || sp->bodies[b.body].removed) continue;
            int p = sp->order_y[ii], q = sp->order_y[jj];
            if (sp->bodies[a.body].type == BODY_STATIC && sp->bodies[b.body].type == BODY_STATIC) continue;
            if (p > q) std::swap(p, q);
            pairs.push_back(((uint64_t)p << 32) | (uint32_t)q);
        }
    }
    std::sort(pairs.begin(), pairs.end());
    sp->st.pairs_bb += (long long)pairs.size();

    for (uint64_t key : pairs) {
        int ia = (int)(key >> 32), ib = (int)(key & 0xFFFFFFFFu);
        auto it = sp->cache.find(key);
        uint32_t gid = (it != sp->cache.end()) ? it->second.gjk_id : 0;
        CollisionInfo info; info.count = 0;
        collide_shapes(sp, ia, ib, &gid, &info);
        if (info.count == 0) { if (it != sp->cache.end()) it->second.gjk_id = gid; continue; }
        Arbiter* arb;
        if (it == sp->cache.end()) {
            Arbiter fresh; memset(&fresh, 0, sizeof(fresh));
            fresh.a = ia; fresh.b = ib; fresh.count = 0; fresh.state = ARB_FIRST; fresh.stamp = 0;
            arb = &sp->cache.emplace(key, fresh).first->second;
        } else arb = &it->second;
        arb->gjk_id = gid;
        arb->body_a = sp->shapes[ia].body; arb->body_b = sp->shapes[ib].body;
        arbiter_update(sp, arb, info);
        int ta = sp->shapes[ia].ctype, tb = sp->shapes[ib].ctype;
        bool solve_ok = true;
        if (sp->handler_mode == 0) {
            // pre_solve of handler (1, 3) — called for every touching pair, solved or not
            if ((ta == 1 && tb == 3) || (ta == 3 && tb == 1)) sp->hits_13++;
        } else {
            int lo = ta < tb ? ta : tb, hi = ta < tb ? tb : ta;
            // begin of handler (1,4) returns False: cpArbiterIgnore until the shapes separate
            if (arb->state == ARB_FIRST && lo == 1 && hi == 4) arb->state = ARB_IGNORE;
            if (arb->state != ARB_IGNORE) {
                if ((lo == 1 || lo == 2) && hi == 3) {
                    // prevent_boundary_intersection (box_delivery_env.py:294-311).  arbiter.shapes[0] is the shape
                    // whose type equals the handler's first type; contact_point_set (cpArbiterGetContactPointSet):
                    // normal = swapped ? -n : n, points[0].distance = dot((b.p + r2) - (a.p + r1), n).
                    bool swapped = (ta != lo);
                    Body& A = sp->bodies[arb->body_a]; Body& B = sp->bodies[arb->body_b];
                    V nrm = swapped ? vneg(arb->n) : arb->n;
                    V p1 = vadd(A.p, arb->con[0].r1), p2 = vadd(B.p, arb->con[0].r2);
                    cpFloat depth = vdot(vsub(p2, p1), arb->n);
                    Body& S0 = swapped ? B : A;
                    V cv = S0.vel;
                    cpFloat k = 2 * vdot(cv, nrm);                  // 2 * current_velocity.dot(normal)
                    V refl = vsub(cv, vmult(nrm, k));               // current_velocity - k * normal
                    V nv = vmult(refl, 0.5);                        // * elasticity
                    bool collision = depth < 0;
                    V corr = vmult(nrm, depth);
                    body_set_position(S0, vadd(body_get_position(S0), corr));
                    S0.vel = nv;
                    if (lo == 1) { sp->robot_hit = collision ? 1 : 0; sp->hits_13++; }
                }
                if (lo == 2 && hi == 4) solve_ok = false;       // (2,4) pre_solve returns False
            }
        }
        // handlers used by ship-ice only do bookkeeping (ship_ice_env.py:150-180): begin/pre_solve return True
        const Body& ba = sp->bodies[arb->body_a]; const Body& bb = sp->bodies[arb->body_b];
        if (arb->state != ARB_IGNORE && solve_ok && !(ba.m == INFINITY && bb.m == INFINITY)) {
            arb->live = true;
            sp->arbiters.push_back(arb);
        } else {
            arb->count = 0;
            if (arb->state != ARB_IGNORE) arb->state = ARB_NORMAL;
        }
        arb->stamp = sp->stamp;
        sp->st.pairs_touch++;
    }
    // cpSpaceArbiterSetFilter
    for (auto it = sp->cache.begin(); it != sp->cache.end();) {
        Arbiter& arb = it->second;
        uint64_t ticks = sp->stamp - arb.stamp;
        const Body& ba = sp->bodies[sp->shapes[arb.a].body]; const Body& bb = sp->bodies[sp->shapes[arb.b].body];
        if (ba.type == BODY_STATIC && bb.type == BODY_STATIC) { ++it; continue; }
        if (ticks >= 1 && arb.state != ARB_CACHED) arb.state = ARB_CACHED;
        if (ticks >= sp->persistence) { it = sp->cache.erase(it); continue; }
        ++it;
    }
    if (sp->order_mode == 1) reorder_by_colour(sp);

    cpFloat slop = sp->slop;
    cpFloat biasCoef = 1.0 - std::pow(sp->collision_bias, dt);
    for (Arbiter* arb : sp->arbiters) arbiter_prestep(sp, arb, dt, slop, biasCoef);

    cpFloat damping = std::pow(sp->damping, dt);
    for (Body& b : sp->bodies) {
        if (b.type != BODY_DYNAMIC || b.removed) continue;   // cpBodyUpdateVelocity skips kinematic; static not in list
        b.vel = vadd(vmult(b.vel, damping), vmult(vadd(sp->gravity, vmult(b.f, b.m_inv)), dt));
        b.w = b.w * damping + b.t * b.i_inv * dt;
        b.f = v(0, 0); b.t = 0.0;
    }
    cpFloat dt_coef = (prev_dt == 0.0 ? 0.0 : dt / prev_dt);
    for (Arbiter* arb : sp->arbiters) arbiter_apply_cached(sp, arb, dt_coef);
    for (int i = 0; i < sp->iterations; i++)
        for (Arbiter* arb : sp->arbiters) arbiter_apply_impulse(sp, arb);

    // statistics only
    long long nz = 0;
    for (Arbiter* arb : sp->arbiters) {
        bool any = false;
        for (int i = 0; i < arb->count; i++) if (arb->con[i].jnAcc != 0 || arb->con[i].jtAcc != 0 || arb->con[i].jBias != 0) any = true;
        if (any) nz++;
    }
    sp->st.steps++; sp->st.arb_solved += (long long)sp->arbiters.size(); sp->st.arb_nonzero += nz;
    if ((long long)sp->arbiters.size() > sp->st.max_arb) sp->st.max_arb = (long long)sp->arbiters.size();
    if (nz > sp->st.max_nonzero) sp->st.max_nonzero = nz;
}

}  // namespace

// ---------------------------------------------------------------- C API (ctypes)
extern "C" {

void cpo_set_portable_trig(int on) { g_portable_trig = on ? 1 : 0; }
int cpo_get_portable_trig(void) { return g_portable_trig; }
void cpo_sincos(double a, double* sn, double* cs) { trig(a, sn, cs); }

void* cpo_space_new(void) {
    Space* sp = new Space();
    sp->iterations = 10; sp->damping = 1.0; sp->gravity = v(0, 0);
    sp->slop = 0.1; sp->collision_bias = std::pow(1.0 - 0.1, 60.0); sp->persistence = 3;
    sp->stamp = 0; sp->curr_dt = 0.0; sp->order_mode = 0; sp->hits_13 = 0;
    memset(&sp->st, 0, sizeof(sp->st));
    return sp;
}
void cpo_space_free(void* h) { delete (Space*)h; }
// Deep copy of a space (bodies, shapes, arbiter cache with its accumulated impulses, step counters): lets a test advance
// the SAME state with two solver orders, or compare one transition from an identical mid-episode state.
void* cpo_space_clone(void* h) {
    Space* src = (Space*)h;
    Space* sp = new Space(*src);
    sp->arbiters.clear();
    for (Arbiter* arb : src->arbiters) {
        auto it = sp->cache.find(((uint64_t)arb->a << 32) | (uint64_t)(uint32_t)arb->b);
        if (it != sp->cache.end()) sp->arbiters.push_back(&it->second);
    }
    return sp;
}
void cpo_space_set(void* h, int iterations, double damping, double gx, double gy) {
    Space* sp = (Space*)h; sp->iterations = iterations; sp->damping = damping; sp->gravity = v(gx, gy);
}
void cpo_space_set_order(void* h, int mode) { ((Space*)h)->order_mode = mode; }

int cpo_body_new(void* h, int type, double x, double y, double angle) {
    Space* sp = (Space*)h;
    Body b;
    b.type = type;
    if (type == BODY_DYNAMIC) { b.m = 0; b.i = 0; b.m_inv = INFINITY; b.i_inv = INFINITY; }
    else { b.m = b.i = INFINITY; b.m_inv = b.i_inv = 0.0; }
    b.cog = v(0, 0); b.p = v(0, 0); b.vel = v(0, 0); b.a = angle; b.w = 0;
    b.v_bias = v(0, 0); b.w_bias = 0; b.f = v(0, 0); b.t = 0;
    set_transform(b);
    body_set_position(b, v(x, y));
    sp->bodies.push_back(b);
    return (int)sp->bodies.size() - 1;
}
// pymunk.Poly(body, verts, radius): convex hull (tolerance 0), planes, mass info with m = 0.
int cpo_poly_new(void* h, int body, int n, const double* xy, double radius) {
    Space* sp = (Space*)h;
    std::vector<V> hull(n);
    for (int i = 0; i < n; i++) hull[i] = v(xy[2 * i], xy[2 * i + 1]);
    int hc = convex_hull(n, hull.data(), 0.0);
    Shape s;
    s.body = body; s.kind = 0; s.count = hc; s.r = radius;
    s.local.resize(hc); s.world.resize(hc);
    for (int i = 0; i < hc; i++) {
        V a = hull[(i - 1 + hc) % hc], b = hull[i];
        s.local[i].v0 = b;
        s.local[i].n = vnormalize(vrperp(vsub(b, a)));
    }
    s.e = 0; s.u = 0; s.ctype = 0;
    V cen = centroid_for_poly(hc, hull.data());
    s.mi_m = 0; s.mi_i = moment_for_poly(1.0, hc, hull.data(), vneg(cen), radius);
    s.mi_cog = cen; s.mi_area = area_for_poly(hc, hull.data(), radius);
    sp->shapes.push_back(s);
    int sid = (int)sp->shapes.size() - 1;
    if (body >= 0) {
        sp->bodies[body].shapes.push_back(sid);
        shape_cache_data(sp, sp->shapes[sid]);
    }
    return sid;
}
// pymunk.Segment(body, a, b, radius): cpSegmentShapeInit — n = rperp(normalize(b - a))
int cpo_segment_new(void* h, int body, double ax, double ay, double bx, double by, double radius) {
    Space* sp = (Space*)h;
    Shape s;
    s.body = body; s.kind = 1; s.count = 2; s.r = radius;
    s.local.resize(2); s.world.resize(2);
    V a = v(ax, ay), b = v(bx, by);
    V n = vrperp(vnormalize(vsub(b, a)));
    s.local[0].v0 = a; s.local[0].n = n;
    s.local[1].v0 = b; s.local[1].n = vneg(n);
    s.e = 0; s.u = 0; s.ctype = 0;
    s.mi_m = 0; s.mi_i = moment_for_segment(1.0, a, b, radius); s.mi_cog = vlerp(a, b, 0.5);
    s.mi_area = radius * (M_PI * radius + 2.0 * vdist(a, b));
    sp->shapes.push_back(s);
    int sid = (int)sp->shapes.size() - 1;
    sp->bodies[body].shapes.push_back(sid);
    shape_cache_data(sp, sp->shapes[sid]);
    return sid;
}
void cpo_shape_set_mass(void* h, int sid, double m) {
    Space* sp = (Space*)h; Shape& s = sp->shapes[sid];
    s.mi_m = m; if (s.body >= 0) { body_accumulate_mass(sp, sp->bodies[s.body]); shape_cache_data(sp, s); }
}
void cpo_shape_set_density(void* h, int sid, double rho) {
    Space* sp = (Space*)h; cpo_shape_set_mass(h, sid, rho * sp->shapes[sid].mi_area);
}
void cpo_shape_set_material(void* h, int sid, double e, double u, int ctype) {
    Shape& s = ((Space*)h)->shapes[sid]; s.e = e; s.u = u; s.ctype = ctype;
}
// pymunk.Poly(None, verts).center_of_gravity  (sim_utils.py:140-141)
void cpo_poly_centroid(int n, const double* xy, double* out2) {
    std::vector<V> hull(n);
    for (int i = 0; i < n; i++) hull[i] = v(xy[2 * i], xy[2 * i + 1]);
    int hc = convex_hull(n, hull.data(), 0.0);
    V c = centroid_for_poly(hc, hull.data());
    out2[0] = c.x; out2[1] = c.y;
}
int cpo_shape_count(void* h, int sid) { return ((Space*)h)->shapes[sid].count; }
// Poly.get_vertices(): hull-ordered local vertices
int cpo_shape_get_verts(void* h, int sid, double* out) {
    Shape& s = ((Space*)h)->shapes[sid];
    for (int i = 0; i < s.count; i++) { out[2 * i] = s.local[i].v0.x; out[2 * i + 1] = s.local[i].v0.y; }
    return s.count;
}
void cpo_shape_get_mass(void* h, int sid, double* out4) {
    Space* sp = (Space*)h; Shape& s = sp->shapes[sid];
    out4[0] = s.mi_m; out4[1] = s.mi_i; out4[2] = s.mi_area;
    out4[3] = (s.body >= 0) ? sp->bodies[s.body].i : 0.0;
}
void cpo_body_set_velocity(void* h, int b, double vx, double vy) { ((Space*)h)->bodies[b].vel = v(vx, vy); }
void cpo_body_set_angular_velocity(void* h, int b, double w) { ((Space*)h)->bodies[b].w = w; }
// out: position.x, position.y, angle, vx, vy, w, m, i   (position = body origin, as pymunk Body.position)
void cpo_body_get(void* h, int b, double* out8) {
    Body& bd = ((Space*)h)->bodies[b];
    V pos = body_get_position(bd);
    out8[0] = pos.x; out8[1] = pos.y; out8[2] = bd.a; out8[3] = bd.vel.x; out8[4] = bd.vel.y; out8[5] = bd.w;
    out8[6] = bd.m; out8[7] = bd.i;
}
int cpo_num_bodies(void* h) { return (int)((Space*)h)->bodies.size(); }
// bulk: [nb][6] = x,y,a,vx,vy,w
void cpo_get_state(void* h, double* out) {
    Space* sp = (Space*)h;
    for (size_t i = 0; i < sp->bodies.size(); i++) {
        Body& bd = sp->bodies[i]; V pos = body_get_position(bd);
        double* o = out + 6 * i;
        o[0] = pos.x; o[1] = pos.y; o[2] = bd.a; o[3] = bd.vel.x; o[4] = bd.vel.y; o[5] = bd.w;
    }
}
// Put every body into an externally given state ([nb][6] = position x, y, angle, vx, vy, w as in cpo_get_state):
// cpBodySetAngle, cpBodySetPosition, cpBodySetVelocity, cpBodySetAngularVelocity; bias velocities cleared.
void cpo_set_state(void* h, const double* in) {
    Space* sp = (Space*)h;
    for (size_t i = 0; i < sp->bodies.size(); i++) {
        Body& bd = sp->bodies[i];
        const double* o = in + 6 * i;
        bd.a = o[2];
        set_transform(bd);
        body_set_position(bd, v(o[0], o[1]));
        bd.vel = v(o[3], o[4]); bd.w = o[5];
        bd.v_bias = v(0, 0); bd.w_bias = 0;
    }
}
// Drop every cached arbiter (contacts and accumulated impulses), as if the shapes had been removed and re-added.
void cpo_clear_cache(void* h) {
    Space* sp = (Space*)h;
    sp->arbiters.clear();
    sp->cache.clear();
}
void cpo_step(void* h, double dt) { space_step((Space*)h, dt); }

// ship_ice_env.py:280-290 — the substep loop of one env step (kept in C so the CPU baseline is
// not dominated by ctypes crossings).  flags bit0 = yaw constraint violated, bit1 = boundary.
int cpo_shipice_substeps(void* h, int ship_body, int steps, double dt, double yaw_lo, double yaw_hi, double map_width) {
    Space* sp = (Space*)h; int flags = 0;
    for (int k = 0; k < steps; k++) {
        space_step(sp, dt);
        Body& sb = sp->bodies[ship_body];
        if (sb.a <= yaw_lo || sb.a >= yaw_hi) { sb.w = 0.0; flags |= 1; }
        V pos = body_get_position(sb);
        if (pos.x < 0 || pos.x > map_width) flags |= 2;
    }
    return flags;
}
// maze_NAMO_env.py:414-419 — the substep loop of one env step.  flags bit1 = boundary constraint
// violated (robot x outside [0, width]); the wall flag is read with cpo_hits_13.
int cpo_maze_substeps(void* h, int robot_body, int steps, double dt, double width) {
    Space* sp = (Space*)h; int flags = 0;
    for (int k = 0; k < steps; k++) {
        space_step(sp, dt);
        V pos = body_get_position(sp->bodies[robot_body]);
        if (pos.x < 0 || pos.x > width) flags |= 2;
    }
    return flags;
}
long long cpo_hits_13(void* h) { return ((Space*)h)->hits_13; }
void cpo_space_set_handler_mode(void* h, int mode) { ((Space*)h)->handler_mode = mode; }
int cpo_robot_hit(void* h) { return ((Space*)h)->robot_hit; }
void cpo_set_robot_hit(void* h, int v) { ((Space*)h)->robot_hit = v; }
// pymunk `body.position = (x, y)` (cpBodySetPosition)
void cpo_body_set_position(void* h, int b, double x, double y) { body_set_position(((Space*)h)->bodies[b], v(x, y)); }
// pymunk `space.remove(body, shape)`: the body stops integrating, its shapes leave the broadphase and
// every cached arbiter that references them is dropped (cpSpaceRemoveShape -> cpSpaceFilterArbiters).
void cpo_body_remove(void* h, int b) {
    Space* sp = (Space*)h; sp->bodies[b].removed = true;
    std::vector<Arbiter*> keep;
    for (Arbiter* arb : sp->arbiters) if (arb->body_a != b && arb->body_b != b) keep.push_back(arb);
    sp->arbiters.swap(keep);
    for (auto it = sp->cache.begin(); it != sp->cache.end();) {
        if (sp->shapes[it->second.a].body == b || sp->shapes[it->second.b].body == b) it = sp->cache.erase(it); else ++it;
    }
}
// body.local_to_world(v) = cpTransformPoint(body->transform, v)
void cpo_body_local_to_world(void* h, int b, double x, double y, double* out2) {
    V w = tpoint(((Space*)h)->bodies[b].tr, v(x, y)); out2[0] = w.x; out2[1] = w.y;
}
// pymunk Space.point_query(p, max_dist, ShapeFilter()) != []  — cpSpacePointQuery: a shape is reported
// when its signed surface distance (rounded radius included) is strictly below max_dist
// (NearestPointQuery: `if(info.distance < context->maxDistance)`); the candidate filter is the
// shape's AABB against the box p +- max(max_dist, 0).
static int point_query_mask(void* h, double px, double py, double max_dist, int stop_first) {
    Space* sp = (Space*)h; V p = v(px, py);
    cpFloat md = std::fmax(max_dist, 0.0);
    int mask = 0;
    for (const Shape& s : sp->shapes) {
        if (sp->bodies[s.body].removed) continue;
        if (!(p.x - md <= s.bb_r && s.bb_l <= p.x + md && p.y - md <= s.bb_t && s.bb_b <= p.y + md)) continue;
        cpFloat dist;
        if (s.kind == 1) {          // cpSegmentShapePointQuery
            V a = s.world[0].v0, b = s.world[1].v0;
            V delta = vsub(a, b);
            cpFloat t = fclamp01(vdot(delta, vsub(p, b)) / vlengthsq(delta));
            V closest = vadd(b, vmult(delta, t));
            dist = vdist(p, closest) - s.r;
        } else {                    // cpPolyShapePointQuery
            int count = s.count; cpFloat r = s.r;
            V v0 = s.world[count - 1].v0; cpFloat minDist = INFINITY; V closestPoint = v(0, 0); bool outside = false;
            for (int i = 0; i < count; i++) {
                V v1 = s.world[i].v0;
                if (vdot(s.world[i].n, vsub(p, v1)) > 0.0) outside = true;
                V d2 = vsub(v0, v1);
                cpFloat t = fclamp01(vdot(d2, vsub(p, v1)) / vlengthsq(d2));
                V closest = vadd(v1, vmult(d2, t));
                cpFloat dd = vdist(p, closest);
                if (dd < minDist) { minDist = dd; closestPoint = closest; }
                v0 = v1;
            }
            (void)closestPoint;
            dist = (outside ? minDist : -minDist) - r;
        }
        if (dist < max_dist) { if (stop_first) return 1; mask |= 1 << s.ctype; }
    }
    return mask;
}
int cpo_point_query_any(void* h, double px, double py, double max_dist) { return point_query_mask(h, px, py, max_dist, 1); }
// bit c set = some shape with collision_type c is reported by Space.point_query(p, max_dist, ShapeFilter())
int cpo_point_query_types(void* h, double px, double py, double max_dist) { return point_query_mask(h, px, py, max_dist, 0); }
void cpo_get_stats(void* h, long long* out11) { memcpy(out11, &((Space*)h)->st, sizeof(Stats)); }
void cpo_reset_stats(void* h) { memset(&((Space*)h)->st, 0, sizeof(Stats)); }
// number of live arbiters in the last step and their (a, b, count, jn0, jt0, jn1, jt1) rows
int cpo_get_arbiters(void* h, double* out, int max_rows) {
    Space* sp = (Space*)h; int k = 0;
    for (Arbiter* arb : sp->arbiters) {
        if (k >= max_rows) break;
        double* o = out + 7 * k++;
        o[0] = arb->a; o[1] = arb->b; o[2] = arb->count;
        o[3] = arb->con[0].jnAcc; o[4] = arb->con[0].jtAcc;
        o[5] = arb->count > 1 ? arb->con[1].jnAcc : 0; o[6] = arb->count > 1 ? arb->con[1].jtAcc : 0;
    }
    return (int)sp->arbiters.size();
}

}  // extern "C"
