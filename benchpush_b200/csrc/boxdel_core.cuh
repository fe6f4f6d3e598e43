// Small-world rigid-body core for box-delivery-v0 (and the other room-sized envs): one warp steps one
// environment whose whole world is ~30 convex shapes (compound kinematic robot, <= 20 dynamic boxes,
// static walls / corners / receptacle).
//
// Replaces the per-env `pymunk.Space.step` calls of BoxDeliveryEnv (reference call sites
// box_delivery_env.py:278-279 settle, :945 controller loop, :1019 until-still loop) together with the
// collision handlers the env installs (:196-229, prevent_boundary_intersection :294-311) and
// `space.remove` (:765-767).  Semantics = Chipmunk2D cpSpaceStep as restated in SURVEY.md Appendix A
// and oracle/chipmunk_oracle.cpp (handler_mode 1): same pair order (ascending shape pair), same
// arbiter cache life cycle, same greedy-colour solver order, fp64 without FMA contraction.
//
// Why not the ship-ice core (sim_step.cuh): that one is built around awake sets and exact motion bounds
// for ~650 bodies; here the handlers teleport bodies inside a step and the env removes bodies, and the
// world is small enough that ALL candidate pairs (a dense, precomputed list of the <= 600 shape pairs
// that can have an effect) are AABB-tested by the 32 lanes every substep.  The arbiter cache is a dense
// table indexed by that pair id: no allocation, no hashing.  Narrowphase leaves (GJK / EPA / support
// edges) are shared with sim_core.cuh.
//
// Pairs that cannot have any effect are not in the list: robot shape vs receptacle ((1,4) begin -> False,
// both bodies of infinite mass), box vs receptacle ((2,4) pre_solve -> False: never solved), wheels /
// bumper vs walls (default handler, both bodies of infinite mass).
#pragma once
#include "sim_core.cuh"

namespace bp {
namespace bd {

constexpr int MAXBOX = 20;        // cfg.boxes.num_boxes_large
constexpr int NROBOT = 6;         // chassis, 4 wheels, front bumper (sim_utils.py:20-76)
constexpr int MAXSTAT = 24;       // receptacle + 4 walls + 3 corners x 3 triangles = 14 in an empty room; + <= 7 columns, or a divider + 2 x 3 triangles = 21
constexpr int SV = 4;             // vertices per shape (rectangles, the bumper trapezoid, corner triangles)
constexpr int NMOV = NROBOT + MAXBOX;
constexpr int MAXCAND = 128;      // AABB-overlapping pairs per substep
constexpr int MAXWP = 24;         // waypoints of one robot path
constexpr int MAXCOLOUR = 32;

enum BdStatus { BD_OK = 0, BD_CAND_OVERFLOW = 1, BD_COLOUR_OVERFLOW = 2, BD_GJK_ITER = 4, BD_PATH_OVERFLOW = 8, BD_STEP_LIMIT = 16 };
enum SlotState { SL_NONE = 0, SL_FIRST = 1, SL_NORMAL = 2, SL_CACHED = 3 };

struct BdParams {
    int NB;                       // boxes per env (capacity = cfg num_boxes)
    int n_static, n_s3;           // static shapes; those with collision_type 3 (walls, corners)
    int s3[MAXSTAT];              // static index of the k-th type-3 shape, ascending
    int recept;                   // static index of the receptacle (collision_type 4)
    int n_slots;                  // dense pair table size
    int iterations, persistence, settle_steps;
    double dt;                    // cfg.controller.dt / cfg.sim.steps (the physics substep)
    double damp, bias_coef, slop;
    double ctrl_dt, target_speed; // cfg.controller.dt, cfg.controller.target_speed
    // maps
    int H, W, LW, SMALL, CROP;    // padded room (430 x 542), local map (224), small map (244), crop (318)
    int small_i0, small_j0, px_off;
    double ppm;                   // local_map_pixels_per_meter
    double local_map_width, step_size, room_length, room_width, robot_radius, robot_half_width;
    double recept_x, recept_y;
    // rewards / episode (box_delivery/config.yaml:135-147, :120-124)
    double partial_scale, goal_reward, collision_penalty, nonmove_penalty, correct_scale, ministep;
    int use_correct, inactivity_cutoff, num_boxes, action_type;   // action_type 0 = position (pixel index), 1 = heading
    double sp_div, sp_scale;      // (sqrt(2) * LW) / ppm ; cfg.env.shortest_path_channel_scale
    // 1 when every free configuration-space cell is visible (skimage line, drawn from the receptacle cell) from the
    // receptacle's cell: then shortest_path(box, receptacle) always prunes to [box, receptacle] whatever the spfa parent
    // chain is (box_delivery_env.py:1243-1250 tests lines from coords[0] = the receptacle cell only while all are free),
    // so shortest_path_distance is the Euclidean distance and needs no spfa solve.  Checked once per handle.
    int recept_visible;           // (AND over the rooms of the handle)
    int exact_parents;            // 1: paths follow the reference solver's own parent tie-breaks (spfa_serial_run)
    int border_blocked;           // 1: the first / last rows and columns of every room's configuration space are blocked
    // ---- env kind: 0 = box-delivery-v0, 1 = area-clearing-v0 (same control chain and handlers, different constants)
    int kind;
    int step_limit;               // STEP_LIMIT of execute_robot_path: 10000 (box_delivery_env.py:62) / 5000 (area_clearing.py:76)
    int post_steps;               // area-clearing: cfg.sim.steps fixed substeps after the path (area_clearing.py:692-693)
    int t_max;                    // area-clearing: cfg.sim.t_max
    double lfc;                   // cfg.controller.Lfc (TargetCourse look-ahead, dp.py:74-84); 0 when absent
    double omega_mul, v_mul;      // apply_controller: omega * 3, v * 2 (box_delivery_env.py:887-889) / omega / 2, v * 5 (area_clearing.py:903-906)
    double bound_x, bound_y;      // PositionController room clamp: map_height (x) and map_width (y) as passed by the env (D17)
    double robot_code, box_code, box_done_code;       // overhead-map segment indices (k / 8 is the pixel value)
    int draw_recept, n_bnd, n_goal, pad1;
    double bnd[8][2];             // area-clearing: clearance polygon (convex), cfg env `boundary`
    double goal[64][2];           // area-clearing: goal points on the clearance boundary (area_clearing.py:225-263)
    double r_boundary, r_cleared, r_putback, r_push_mult, r_terminal, r_truncation, r_nonmove;   // area_clearing.py:41-52
};

struct BdBody {                   // 176 B
    double px, py, pa;            // cog position (world), angle
    double vx, vy, vw, bx, by, bw;
    double minv, iinv, rc, rs, cgx, cgy;
    double cpx, cpy, cpa;         // pose the cached world vertices / AABB of the body's shapes were computed from
    int alive, idx;
    int moved, pad;               // pose differs from (cpx, cpy, cpa) in this substep: shapes re-cached, pairs re-evaluated
};

struct alignas(16) BdSlot {       // one shape pair of the dense table = cpArbiter while state != SL_NONE
    int stamp; unsigned gjk;
    unsigned char state, cnt;
    // Geometry cache of the pair (independent of the arbiter life cycle): a world with damping 0 is quasi-static, most
    // shapes do not move in most substeps, and a pair of two unmoved shapes has the same AABB test and the same
    // closest-feature result as in the previous substep.  ov = AABBs overlapped at the last evaluation; mvalid = the
    // manifold below is the narrowphase result for the current poses.
    unsigned char ov, mvalid;
    unsigned mgjk;
    double nx, ny, e, u;
    ContactRec c[2];
    double mnx, mny, mp1[2][2], mp2[2][2];
    unsigned mhash[2]; int mcnt, pad2;
};

struct BdLayout {                 // immutable per layout
    double start[3];
    int n_boxes, room;            // room: index of the layout's static shapes / maps (BdConst::rooms)
    double box_init[MAXBOX][3];
    double box_lv[MAXBOX][SV][2], box_ln[MAXBOX][SV][2];
    double box_cog[MAXBOX][2], box_minv[MAXBOX], box_iinv[MAXBOX];
};

struct BdStatics {                // immutable per ROOM: static shapes (world) + the receptacle's cell in the room's configuration space
    double st_lv[MAXSTAT][SV][2], st_ln[MAXSTAT][SV][2], st_pose[MAXSTAT][3];   // as created (body-local + body pose)
    alignas(16) double st_v[MAXSTAT][SV][2];                                       // world (filled by bd_init_static); 128-bit vertex loads
    double st_n[MAXSTAT][SV][2], st_p[MAXSTAT][2];
    BBox st_bb[MAXSTAT];
    double st_rad[MAXSTAT], st_e[MAXSTAT], st_u[MAXSTAT];
    int st_nv[MAXSTAT], st_ctype[MAXSTAT];           // st_nv == 0: slot not used in this room (fewer columns than the handle's maximum)
    int recept_ci, recept_cj;                        // closest valid cell of the receptacle position (host-built maps)
    int pad[2];
    // collision_type-3 shapes as the reference rasterises them into the small obstacle map (update_configuration_space,
    // box_delivery_env.py:1151-1162): pixel vertices of the polygon as passed to cv2.fillPoly; st_npx 0 = not drawn.  Used by
    // the on-device room build (boxdel_room.cuh).
    int st_npx[MAXSTAT], st_px[MAXSTAT][SV], st_py[MAXSTAT][SV];
};

struct BdShapes {                 // immutable per handle: robot shapes (body-local), box material
    double rb_lv[NROBOT][SV][2], rb_ln[NROBOT][SV][2], rb_rad[NROBOT], rb_e[NROBOT], rb_u[NROBOT];
    int rb_nv[NROBOT];
    double rb_foot[SV][2]; int n_foot, pad_foot;     // polygon drawn for the robot in the overhead map (body-local)
    double box_rad, box_e, box_u;
};

struct BdSlot;
struct BdEnvCold {                // per env, global memory only: touched once per substep or once per env step
    double path[MAXWP][3];
    double init_dist[MAXBOX], final_dist[MAXBOX];
    double dist_key[MAXBOX][2], dist_val[MAXBOX];   // last solved shortest_path_distance(box, receptacle) and the box origin it was solved for
    unsigned char dist_ok[MAXBOX];
    double prev_boxes[MAXBOX][SV][2];   // CostMap.get_obs_from_poly at the previous env step (work accounting)
    short order[MAXCAND];         // this step's solve list (slot ids), colour-sorted
    short cand[MAXCAND];
    unsigned char ccnt[MAXCAND];
    Manifold cman[MAXCAND];
};
struct BdEnv {                    // mutable per env: the HOT header (6.6 KB) — k_bd_move keeps a copy in shared memory for the whole launch
    int layout, episode, t, status, stamp, n_order, robot_hit, inactivity, cum_boxes, done, truncated, n_prev;
    int cleared_count, room;      // room = layouts[layout].room (static shapes)
    int map_slot, pad_ms;         // slot of this env's maps: room, or the env's own slot when the maps are built on the device at reset
    int nc, fast;                 // this substep: candidate pairs; 1 = nothing overlaps and nothing is cached (phases B-D skipped)
    int n_path, move_sign, need_spfa, src_i, src_j, tgt_i, tgt_j, n_live;      // n_live: slots with state != SL_NONE
    long long stat_substeps, stat_cand, stat_arb;
    double prev_dt, cum_distance, cum_reward, total_work, last_work, reward, ministeps;
    double init_pos[2], init_heading, robot_distance;
    double tgt[2];
    BdEnvCold* cold;              // the env block's cold part and dense pair table (pointers stay valid in any copy of the header)
    BdSlot* slots;
    BdBody body[1 + MAXBOX];      // [0] robot (all six shapes), [1 + i] box i
    alignas(16) double wv[NMOV][SV][2];   // world vertices of the moving shapes: [0, 6) robot, [6 + i] box i (128-bit vertex loads)
    BBox bb[NMOV];
    unsigned char completed[MAXBOX];
};

// env block = [BdEnv header][BdEnvCold][BdSlot x n_slots]
BP_HD size_t bd_env_bytes(int n_slots) { return BP_ALIGN128(sizeof(BdEnv)) + BP_ALIGN128(sizeof(BdEnvCold)) + BP_ALIGN128((size_t)n_slots * sizeof(BdSlot)); }

struct BdConst {
    BdParams P; BdShapes S;
    const BdLayout* layouts; int n_layouts, n_envs;
    const short* pair_a; const short* pair_b;      // [n_slots] shape ids of the dense pair table (a < b)
    // rooms: static shapes and maps (box_delivery_env.py:1140-1175).  One room per handle in the obstacle-free configurations;
    // one per layout when the reference redraws columns / a divider at every reset (box_delivery_env.py:358-471).
    const BdStatics* rooms; int n_rooms;
    int maps_per_env;                                       // 0: map slot = room (host-built maps); 1: map slot = env, built on the device at reset
    const uint8_t* cspace; const uint8_t* cspace_thin;      // [n_rooms][H][W] 1 = free
    const int32_t* closest;                                 // [n_rooms][2][H][W] distance_transform_edt indices
    const uint8_t* small_free;                              // [n_rooms][SMALL][SMALL] 1 = free floor (small_obstacle_map)
    const uint8_t* robot_chan;                              // [LW][LW] robot_state_channel * 255
    const float* ch3_map;                                   // [n_rooms][H][W] fourth observation channel before the window shift: scaled spfa field of the
                                                            //        receptacle (box-delivery) / goal-point map (area-clearing); static per room
    char* envs; size_t env_bytes;
};
BP_DEV const BdStatics& bd_room(const BdConst& C, const BdEnv* E) { return C.rooms[E->room]; }
BP_DEV const uint8_t* bd_cspace(const BdConst& C, int room) { return C.cspace + (size_t)room * C.P.H * C.P.W; }
BP_DEV const uint8_t* bd_cspace_thin(const BdConst& C, int room) { return C.cspace_thin + (size_t)room * C.P.H * C.P.W; }
BP_DEV const int32_t* bd_closest(const BdConst& C, int room) { return C.closest + (size_t)room * 2 * C.P.H * C.P.W; }
BP_DEV const uint8_t* bd_small_free(const BdConst& C, int room) { return C.small_free + (size_t)room * C.P.SMALL * C.P.SMALL; }
BP_DEV const float* bd_ch3_map(const BdConst& C, int room) { return C.ch3_map + (size_t)room * C.P.H * C.P.W; }
BP_DEV BdEnv* bd_env(const BdConst& C, int e) { return (BdEnv*)(C.envs + (size_t)e * C.env_bytes); }
BP_DEV BdSlot* bd_slots(BdEnv* e) { return e->slots; }
// pointers of the header into its env block (set at reset, from the header's GLOBAL address)
BP_DEV void bd_link_env(BdEnv* g) {
    g->cold = (BdEnvCold*)((char*)g + BP_ALIGN128(sizeof(BdEnv)));
    g->slots = (BdSlot*)((char*)g + BP_ALIGN128(sizeof(BdEnv)) + BP_ALIGN128(sizeof(BdEnvCold)));
}

// dense pair table (ascending (shape a, shape b) with a < b):  shape ids: [0,6) robot, [6, 6+NB) boxes, then statics
static inline int bd_build_pairs(const BdParams& P, short* pa, short* pb) {
    int n = 0, NB = P.NB;
    auto put = [&](int a, int b) { if (pa) { pa[n] = (short)a; pb[n] = (short)b; } n++; };
    for (int j = 0; j < NB; j++) put(0, NROBOT + j);
    for (int k = 0; k < P.n_s3; k++) put(0, NROBOT + NB + P.s3[k]);
    for (int r = 1; r < NROBOT; r++) for (int j = 0; j < NB; j++) put(r, NROBOT + j);
    for (int i = 0; i < NB; i++) {
        for (int j = i + 1; j < NB; j++) put(NROBOT + i, NROBOT + j);
        for (int k = 0; k < P.n_s3; k++) put(NROBOT + i, NROBOT + NB + P.s3[k]);
    }
    return n;
}

// The leaves below are out of line on purpose (BP_DEVNI): a substep is one long pass over ~8 k dynamic instructions with
// little reuse, 16 resident warps sit at 16 different places of it, and the first build (everything inlined, 125 KB of
// SASS for the substep) was bound by instruction fetch (ncu: stall_no_instruction 6.6 of 14.7 cycles per issue, i-cache
// hit rate 64 %).  One shared copy per leaf keeps the hot footprint inside the instruction cache.
// ------------------------------------------------------------------ shape access
struct ShapeRef { const double* v; const double* ln; int n; double rad, e, u; double rc, rs; int body; BBox bb; };   // body: -1 = static
BP_DEVNI ShapeRef shape_ref(const BdConst& C, const BdEnv* E, int s) {
    ShapeRef r; const int NB = C.P.NB;
    if (s < NROBOT) {
        r.v = &E->wv[s][0][0]; r.ln = &C.S.rb_ln[s][0][0]; r.n = C.S.rb_nv[s]; r.rad = C.S.rb_rad[s]; r.e = C.S.rb_e[s]; r.u = C.S.rb_u[s];
        r.rc = E->body[0].rc; r.rs = E->body[0].rs; r.body = 0; r.bb = E->bb[s];
    } else if (s < NROBOT + NB) {
        int i = s - NROBOT; const BdLayout& L = C.layouts[E->layout];
        r.v = &E->wv[s][0][0]; r.ln = &L.box_ln[i][0][0]; r.n = SV; r.rad = C.S.box_rad; r.e = C.S.box_e; r.u = C.S.box_u;
        r.rc = E->body[1 + i].rc; r.rs = E->body[1 + i].rs; r.body = 1 + i; r.bb = E->bb[s];
    } else {
        int k = s - NROBOT - NB; const BdStatics& T = bd_room(C, E);
        r.v = &T.st_v[k][0][0]; r.ln = &T.st_n[k][0][0]; r.n = T.st_nv[k]; r.rad = T.st_rad[k]; r.e = T.st_e[k]; r.u = T.st_u[k];
        r.rc = 1.0; r.rs = 0.0; r.body = -1; r.bb = T.st_bb[k];
    }
    return r;
}

// world vertices + AABB of one shape from its body pose (cpPolyShapeCacheData)
BP_DEVNI void cache_shape(const double* lv, int n, double rad, double c, double s, double px, double py, double cgx, double cgy,
                        double* wv, BBox* bb) {
    double tx = px - (cgx * c - cgy * s);
    double ty = py - (cgx * s + cgy * c);
    double ms = -s;
    double l = INFINITY, r = -INFINITY, bt = INFINITY, tp = -INFINITY;
    #pragma unroll 1
    for (int i = 0; i < n; i++) {
        double qx = lv[2 * i], qy = lv[2 * i + 1];
        double wx = c * qx + ms * qy + tx;
        double wy = s * qx + c * qy + ty;
        wv[2 * i] = wx; wv[2 * i + 1] = wy;
        l = fmin(l, wx); r = fmax(r, wx); bt = fmin(bt, wy); tp = fmax(tp, wy);
    }
    bb->l = l - rad; bb->b = bt - rad; bb->r = r + rad; bb->t = tp + rad;
}
// pymunk Body.position (local origin in world space) and its setter (cpBodySetPosition)
BP_DEV V2 bd_origin(const BdBody& b) { return mk(b.px - (b.cgx * b.rc - b.cgy * b.rs), b.py - (b.cgx * b.rs + b.cgy * b.rc)); }
BP_DEV void bd_set_origin(BdBody& b, V2 pos) {
    // p = transform_vect(cog) + position
    b.px = (b.rc * b.cgx + (-b.rs) * b.cgy) + pos.x;
    b.py = (b.rs * b.cgx + b.rc * b.cgy) + pos.y;
}
// body.local_to_world(v)
BP_DEV V2 bd_local_to_world(const BdBody& b, double x, double y) {
    double tx = b.px - (b.cgx * b.rc - b.cgy * b.rs), ty = b.py - (b.cgx * b.rs + b.cgy * b.rc);
    return mk(b.rc * x + (-b.rs) * y + tx, b.rs * x + b.rc * y + ty);
}

// statics: world vertices / normals / AABB from the creation pose (run once per handle)
BP_DEV void bd_init_static(BdStatics& S, int k) {
    if (S.st_nv[k] == 0) {        // unused slot: an AABB nothing overlaps
        S.st_bb[k].l = INFINITY; S.st_bb[k].b = INFINITY; S.st_bb[k].r = -INFINITY; S.st_bb[k].t = -INFINITY;
        S.st_p[k][0] = 0.0; S.st_p[k][1] = 0.0;
        return;
    }
    double s, c; bp_sincos(S.st_pose[k][2], &s, &c);
    // static bodies carry no mass: cog = 0, p = position
    S.st_p[k][0] = S.st_pose[k][0]; S.st_p[k][1] = S.st_pose[k][1];
    cache_shape(&S.st_lv[k][0][0], S.st_nv[k], S.st_rad[k], c, s, S.st_pose[k][0], S.st_pose[k][1], 0.0, 0.0, &S.st_v[k][0][0], &S.st_bb[k]);
    for (int i = 0; i < S.st_nv[k]; i++) {
        double nx = S.st_ln[k][i][0], ny = S.st_ln[k][i][1];
        S.st_n[k][i][0] = c * nx + (-s) * ny; S.st_n[k][i][1] = s * nx + c * ny;
    }
}

// ------------------------------------------------------------------ narrowphase of one shape pair
BP_DEVNI void bd_manifold(const ShapeRef& A, const ShapeRef& B, const Closest& pts, Manifold* m) {
    m->count = 0;
    Poly PA; PA.v = A.v; PA.n = A.n; Poly PB; PB.v = B.v; PB.n = B.n;
    double ra = A.rad, rb = B.rad;
    if (pts.d - ra - rb <= 0.0) {
        EdgeW e1 = support_edge(PA, A.ln, A.rc, A.rs, pts.n);
        EdgeW e2 = support_edge(PB, B.ln, B.rc, B.rs, neg(pts.n));
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
                if (dot(p2 - p1, n) <= 0.0) { int k = m->count++; m->p1[k] = p1; m->p2[k] = p2; m->hash[k] = (unsigned)(e1.ia << 8 | e2.ib); }
            }
            {
                V2 p1 = n * ra + lerp(e1.a, e1.b, clamp01((d_e2_a - d_e1_a) * e1_denom));
                V2 p2 = n * (-rb) + lerp(e2.a, e2.b, clamp01((d_e1_b - d_e2_a) * e2_denom));
                if (dot(p2 - p1, n) <= 0.0) { int k = m->count++; m->p1[k] = p1; m->p2[k] = p2; m->hash[k] = (unsigned)(e1.ib << 8 | e2.ia); }
            }
        }
    }
}

// body record of shape-ref body index (-1 = static: zero velocity, infinite mass, position st_p)
struct BodyV { double px, py, vx, vy, vw, minv, iinv; };
BP_DEVNI BodyV body_view(const BdConst& C, const BdEnv* E, int s, int body) {
    BodyV b;
    if (body >= 0) { const BdBody& q = E->body[body]; b.px = q.px; b.py = q.py; b.vx = q.vx; b.vy = q.vy; b.vw = q.vw; b.minv = q.minv; b.iinv = q.iinv; }
    else { int k = s - NROBOT - C.P.NB; b.px = bd_room(C, E).st_p[k][0]; b.py = bd_room(C, E).st_p[k][1]; b.vx = b.vy = b.vw = 0.0; b.minv = b.iinv = 0.0; }
    return b;
}

// cpArbiterUpdate for candidate q (manifold already computed); returns nothing, marks the slot touched this step
BP_DEVNI void bd_arb_update(const BdConst& C, BdEnv* E, int q) {
    int p = E->cold->cand[q];
    const Manifold& m = E->cold->cman[q];
    BdSlot& S = bd_slots(E)[p];
    int sa = C.pair_a[p], sb = C.pair_b[p];
    ShapeRef A = shape_ref(C, E, sa), B = shape_ref(C, E, sb);
    BodyV ba = body_view(C, E, sa, A.body), bb = body_view(C, E, sb, B.body);
    bool fresh = (S.state == SL_NONE);
    int oc = fresh ? 0 : S.cnt;
    double jn_new[2] = {0, 0}, jt_new[2] = {0, 0};
    #pragma unroll 1
    for (int i = 0; i < m.count; i++)
        #pragma unroll 1
        for (int j = 0; j < oc; j++)
            if (m.hash[i] == S.c[j].hash) { jn_new[i] = S.c[j].jn; jt_new[i] = S.c[j].jt; }
    #pragma unroll 1
    for (int i = 0; i < m.count; i++) {
        ContactRec& c = S.c[i];
        c.r1x = m.p1[i].x - ba.px; c.r1y = m.p1[i].y - ba.py;
        c.r2x = m.p2[i].x - bb.px; c.r2y = m.p2[i].y - bb.py;
        c.hash = m.hash[i]; c.jn = jn_new[i]; c.jt = jt_new[i];
    }
    S.cnt = (unsigned char)m.count;
    S.nx = m.n.x; S.ny = m.n.y;
    S.e = A.e * B.e; S.u = A.u * B.u;
    if (fresh || S.state == SL_CACHED) S.state = SL_FIRST;
    S.stamp = E->stamp;
}

// prevent_boundary_intersection (box_delivery_env.py:294-311) for candidate q = (moving shape a, type-3 static b).
// shapes[0] is always `a` here (robot chassis or box precede the statics in shape order: not swapped).
BP_DEVNI bool bd_boundary_presolve(const BdConst& C, BdEnv* E, int q) {
    int p = E->cold->cand[q];
    BdSlot& S = bd_slots(E)[p];
    int sa = C.pair_a[p], sb = C.pair_b[p];
    int body = (sa < NROBOT) ? 0 : 1 + (sa - NROBOT);
    BdBody& A = E->body[body];
    int k = sb - NROBOT - C.P.NB;
    V2 n = mk(S.nx, S.ny);
    V2 p1 = mk(A.px + S.c[0].r1x, A.py + S.c[0].r1y);
    V2 p2 = mk(bd_room(C, E).st_p[k][0] + S.c[0].r2x, bd_room(C, E).st_p[k][1] + S.c[0].r2y);
    double depth = dot(p2 - p1, n);
    V2 cv = mk(A.vx, A.vy);
    double kk = 2 * dot(cv, n);
    V2 refl = cv - n * kk;
    V2 nv = refl * 0.5;
    V2 corr = n * depth;
    bd_set_origin(A, bd_origin(A) + corr);
    A.vx = nv.x; A.vy = nv.y;
    return depth < 0;
}

BP_DEVNI void bd_prestep(const BdConst& C, BdEnv* E, int p) {
    BdSlot& S = bd_slots(E)[p];
    int sa = C.pair_a[p], sb = C.pair_b[p];
    int NB = C.P.NB;
    BodyV A = body_view(C, E, sa, sa < NROBOT ? 0 : (sa < NROBOT + NB ? 1 + sa - NROBOT : -1));
    BodyV B = body_view(C, E, sb, sb < NROBOT ? 0 : (sb < NROBOT + NB ? 1 + sb - NROBOT : -1));
    V2 n = mk(S.nx, S.ny), tn = perp(n);
    V2 delta = mk(B.px - A.px, B.py - A.py);
    V2 va = mk(A.vx, A.vy), vb = mk(B.vx, B.vy);
    #pragma unroll 1
    for (int i = 0; i < S.cnt; i++) {
        ContactRec& c = S.c[i];
        V2 r1 = mk(c.r1x, c.r1y), r2 = mk(c.r2x, c.r2y);
        double r1n = cross(r1, n), r2n = cross(r2, n);
        double kn = (A.minv + A.iinv * r1n * r1n) + (B.minv + B.iinv * r2n * r2n);
        double r1t = cross(r1, tn), r2t = cross(r2, tn);
        double kt = (A.minv + A.iinv * r1t * r1t) + (B.minv + B.iinv * r2t * r2t);
        c.nm = 1.0 / kn; c.tm = 1.0 / kt;
        double dist = dot((r2 - r1) + delta, n);
        c.bias = -C.P.bias_coef * fmin(0.0, dist + C.P.slop) / C.P.dt;
        c.jb = 0.0;
        V2 vr = (vb + perp(r2) * B.vw) - (va + perp(r1) * A.vw);
        c.bounce = dot(vr, n) * S.e;
    }
}

// One solver visit of slot p: warm == true applies the cached impulses (cpArbiterApplyCachedImpulse), otherwise
// one cpArbiterApplyImpulse.  Bodies of infinite mass are read-only.
BP_DEVNI void bd_solve(const BdConst& C, BdEnv* E, int p, bool warm, double dt_coef) {
    BdSlot& S = bd_slots(E)[p];
    if (warm && S.state == SL_FIRST) return;
    int sa = C.pair_a[p], sb = C.pair_b[p];
    int NB = C.P.NB;
    int ia_ = sa < NROBOT ? 0 : (sa < NROBOT + NB ? 1 + sa - NROBOT : -1);
    int ib_ = sb < NROBOT ? 0 : (sb < NROBOT + NB ? 1 + sb - NROBOT : -1);
    BdBody zero; zero.vx = zero.vy = zero.vw = zero.bx = zero.by = zero.bw = 0.0; zero.minv = zero.iinv = 0.0;
    BdBody* PA = ia_ >= 0 ? &E->body[ia_] : &zero;
    BdBody* PB = ib_ >= 0 ? &E->body[ib_] : &zero;
    double ma = PA->minv, ia = PA->iinv, mb = PB->minv, ib = PB->iinv;
    V2 n = mk(S.nx, S.ny), tn = perp(n);
    V2 va = mk(PA->vx, PA->vy), vb = mk(PB->vx, PB->vy);
    double wa = PA->vw, wb = PB->vw;
    V2 ba = mk(PA->bx, PA->by), bb = mk(PB->bx, PB->by);
    double bwa = PA->bw, bwb = PB->bw;
    #pragma unroll 1
    for (int i = 0; i < S.cnt; i++) {
        ContactRec& c = S.c[i];
        V2 r1 = mk(c.r1x, c.r1y), r2 = mk(c.r2x, c.r2y);
        if (warm) {
            V2 j = rotv(n, mk(c.jn, c.jt)) * dt_coef, nj = neg(j);
            va = va + nj * ma; wa += ia * cross(r1, nj);
            vb = vb + j * mb; wb += ib * cross(r2, j);
            continue;
        }
        V2 vb1 = ba + perp(r1) * bwa, vb2 = bb + perp(r2) * bwb;
        V2 vr = (vb + perp(r2) * wb) - (va + perp(r1) * wa);
        double vbn = dot(vb2 - vb1, n), vrn = dot(vr, n), vrt = dot(vr, tn);
        double jbn = (c.bias - vbn) * c.nm;
        double jbnOld = c.jb;
        double jb = fmax(jbnOld + jbn, 0.0);
        double jn = -(c.bounce + vrn) * c.nm;
        double jnOld = c.jn;
        double jnAcc = fmax(jnOld + jn, 0.0);
        double jtMax = S.u * jnAcc;
        double jt = -vrt * c.tm;
        double jtOld = c.jt;
        double jtAcc = clampd(jtOld + jt, -jtMax, jtMax);
        c.jb = jb; c.jn = jnAcc; c.jt = jtAcc;
        V2 jbv = n * (jb - jbnOld), njbv = neg(jbv);
        ba = ba + njbv * ma; bwa += ia * cross(r1, njbv);
        bb = bb + jbv * mb; bwb += ib * cross(r2, jbv);
        V2 j = rotv(n, mk(jnAcc - jnOld, jtAcc - jtOld)), nj = neg(j);
        va = va + nj * ma; wa += ia * cross(r1, nj);
        vb = vb + j * mb; wb += ib * cross(r2, j);
    }
    if (ma != 0.0 || ia != 0.0) { PA->vx = va.x; PA->vy = va.y; PA->vw = wa; PA->bx = ba.x; PA->by = ba.y; PA->bw = bwa; }
    if (mb != 0.0 || ib != 0.0) { PB->vx = vb.x; PB->vy = vb.y; PB->vw = wb; PB->bx = bb.x; PB->by = bb.y; PB->bw = bwb; }
}

// ------------------------------------------------------------------ one substep (cpSpaceStep) for one env, one warp
// A substep is split into four phases so that a warp that steps several environments can run each phase for all of them
// before moving on (boxdel_env.cuh: bd_move): the code of one phase is then fetched once and executed K times.
// Phase A: integrate, re-cache moved shapes, broadphase; sets E->nc and E->fast.
BP_DEVNI void bd_phase_a(const BdConst& C, BdEnv* E) {
    const int lane = bp_lane();
    const BdParams& P = C.P;
    const int NB = P.NB;
    const BdLayout& L = C.layouts[E->layout];
    BdSlot* slots = bd_slots(E);
    const double dt = P.dt;
    // arbiters solved in the previous step leave FIRST_COLLISION
    #pragma unroll 1
    for (int k = lane; k < E->n_order; k += BP_NL) { BdSlot& S = slots[E->cold->order[k]]; if (S.state != SL_NONE) S.state = SL_NORMAL; }
    bp_wsync();
    if (lane == 0) { E->stamp++; E->stat_substeps++; }
    // 1. integrate positions (kinematic robot + live boxes), refresh the shapes
    #pragma unroll 1
    for (int b = lane; b <= NB; b += BP_NL) {
        BdBody& B = E->body[b];
        if (b > 0 && !B.alive) continue;
        B.px = B.px + (B.vx + B.bx) * dt; B.py = B.py + (B.vy + B.by) * dt;
        B.pa = B.pa + (B.vw + B.bw) * dt;
        B.bx = 0.0; B.by = 0.0; B.bw = 0.0;
        // handlers / teleports edit (px, py) after the shapes were cached, so the comparison also catches those
        int mv = !(B.px == B.cpx && B.py == B.cpy && B.pa == B.cpa);
        if (mv) { if (B.pa != B.cpa) bp_sincos(B.pa, &B.rs, &B.rc); B.cpx = B.px; B.cpy = B.py; B.cpa = B.pa; }
        B.moved = mv;
    }
    bp_wsync();
    #pragma unroll 1
    for (int s = lane; s < NROBOT + NB; s += BP_NL) {
        if (s < NROBOT) {
            const BdBody& B = E->body[0];
            if (B.moved) cache_shape(&C.S.rb_lv[s][0][0], C.S.rb_nv[s], C.S.rb_rad[s], B.rc, B.rs, B.px, B.py, 0.0, 0.0, &E->wv[s][0][0], &E->bb[s]);
        } else {
            int i = s - NROBOT; const BdBody& B = E->body[1 + i];
            if (!B.alive || !B.moved) continue;
            cache_shape(&L.box_lv[i][0][0], SV, C.S.box_rad, B.rc, B.rs, B.px, B.py, B.cgx, B.cgy, &E->wv[s][0][0], &E->bb[s]);
        }
    }
    bp_wsync();
    const int stamp = E->stamp;
    // 2. broadphase over the dense pair table (ballot compaction keeps the canonical order)
    int nc = 0;
    #pragma unroll 1
    for (int base = 0; base < P.n_slots; base += BP_NL) {
        int p = base + lane;
        bool take = false;
        if (p < P.n_slots) {
            int sa = C.pair_a[p], sb = C.pair_b[p];
            bool alive = true;
            if (sa >= NROBOT) alive = E->body[1 + sa - NROBOT].alive;
            if (sb >= NROBOT && sb < NROBOT + NB) alive = alive && E->body[1 + sb - NROBOT].alive;
            if (alive) {
                bool dirty = E->body[sa < NROBOT ? 0 : 1 + sa - NROBOT].moved || (sb < NROBOT + NB && E->body[1 + sb - NROBOT].moved);
                BdSlot& S = slots[p];
                if (dirty) {
                    BBox A = E->bb[sa];
                    BBox B = (sb < NROBOT + NB) ? E->bb[sb] : bd_room(C, E).st_bb[sb - NROBOT - NB];
                    take = (A.l <= B.r && B.l <= A.r && A.b <= B.t && B.b <= A.t);
                    S.ov = take ? 1 : 0; S.mvalid = 0;
                } else take = S.ov != 0;
            }
        }
        int tot; int rk = bp_wrank(take, &tot);
        if (take && nc + rk < MAXCAND) E->cold->cand[nc + rk] = (short)p;
        nc += tot;
    }
    if (nc > MAXCAND) { nc = MAXCAND; if (lane == 0) E->status |= BD_CAND_OVERFLOW; }
    const bool fast = (nc == 0 && E->n_live == 0);
    if (lane == 0) { E->nc = nc; E->fast = fast ? 1 : 0; }
    bp_wsync();
    if (fast) {
        // nothing overlaps and no arbiter is cached: the rest of cpSpaceStep only applies the velocity update
        #pragma unroll 1
        for (int b = 1 + lane; b <= NB; b += BP_NL) {
            BdBody& B = E->body[b];
            if (!B.alive) continue;
            B.vx = B.vx * P.damp + (0.0 + 0.0 * B.minv) * dt;
            B.vy = B.vy * P.damp + (0.0 + 0.0 * B.minv) * dt;
            B.vw = B.vw * P.damp + 0.0 * B.iinv * dt;
        }
        if (lane == 0) { E->n_order = 0; E->prev_dt = dt; }
        bp_wsync();
    }
    (void)stamp;
}

// Phase B: narrowphase.
BP_DEVNI void bd_phase_b(const BdConst& C, BdEnv* E) {
    const int lane = bp_lane();
    BdSlot* slots = bd_slots(E);
    const int nc = E->nc;
    // 3. narrowphase, lanes over candidates (pure geometry on the cached world vertices); a pair of two unmoved shapes
    //    reuses the result of the substep in which it was last evaluated
    #pragma unroll 1
    for (int q = lane; q < nc; q += BP_NL) {
        int p = E->cold->cand[q];
        BdSlot& S = slots[p];
        Manifold m;
#ifdef BD_NO_PAIR_CACHE
        S.mvalid = 0;       // testing aid: always run the narrowphase
#endif
        if (S.mvalid) {
            m.count = S.mcnt; m.n = mk(S.mnx, S.mny);
            #pragma unroll 1
            for (int i = 0; i < 2; i++) { m.p1[i] = mk(S.mp1[i][0], S.mp1[i][1]); m.p2[i] = mk(S.mp2[i][0], S.mp2[i][1]); m.hash[i] = S.mhash[i]; }
            if (S.state != SL_NONE || m.count > 0) S.gjk = S.mgjk;
        } else {
            int sa = C.pair_a[p], sb = C.pair_b[p];
            ShapeRef A = shape_ref(C, E, sa), B = shape_ref(C, E, sb);
            Poly PA; PA.v = A.v; PA.n = A.n; Poly PB; PB.v = B.v; PB.n = B.n;
            unsigned gid = (S.state != SL_NONE) ? S.gjk : 0u;
            Closest pts; unsigned hint = 0; int iters = 0, capped = 0;
            bool inside = gjk_only(PA, PB, bb_centre(A.bb), bb_centre(B.bb), gid, &pts, &hint, &iters, &capped);
            if (inside) pts = gjk_epa(PA, PB, hint, &capped);
            bd_manifold(A, B, pts, &m);
            if (S.state != SL_NONE || m.count > 0) S.gjk = pts.id;
            S.mcnt = m.count; S.mnx = m.n.x; S.mny = m.n.y; S.mgjk = pts.id; S.mvalid = 1;
            #pragma unroll 1
            for (int i = 0; i < m.count; i++) { S.mp1[i][0] = m.p1[i].x; S.mp1[i][1] = m.p1[i].y; S.mp2[i][0] = m.p2[i].x; S.mp2[i][1] = m.p2[i].y; S.mhash[i] = m.hash[i]; }
        }
        E->cold->cman[q] = m; E->cold->ccnt[q] = (unsigned char)m.count;
    }
    bp_wsync();
}

// Phase C: arbiter updates + handlers, cache expiry, solve order.
BP_DEVNI void bd_phase_c(const BdConst& C, BdEnv* E) {
    const int lane = bp_lane();
    const BdParams& P = C.P;
    const int NB = P.NB;
    BdSlot* slots = bd_slots(E);
    const int nc = E->nc, stamp = E->stamp;
    // 4. arbiter updates + handlers, in an order equivalent to the sequential canonical sweep:
    //    (chassis, box) -> (chassis, wall)* sequential, edits the robot -> everything else except (box, wall)
    //    -> (box i, wall)* sequential per box, boxes in parallel.
    const int seg0 = NB, seg1 = NB + P.n_s3;
    #pragma unroll 1
    for (int q = lane; q < nc; q += BP_NL) if (E->cold->ccnt[q] && E->cold->cand[q] < seg0) bd_arb_update(C, E, q);
    bp_wsync();
    if (lane == 0) {
        #pragma unroll 1
        for (int q = 0; q < nc; q++) {
            int p = E->cold->cand[q];
            if (p < seg0 || p >= seg1 || !E->cold->ccnt[q]) continue;
            bd_arb_update(C, E, q);
            E->robot_hit = bd_boundary_presolve(C, E, q) ? 1 : 0;
            // kinematic vs static: never solved — contacts dropped, state back to NORMAL
            BdSlot& S = slots[p]; S.cnt = 0; S.state = SL_NORMAL;
        }
    }
    bp_wsync();
    #pragma unroll 1
    for (int q = lane; q < nc; q += BP_NL) {
        int p = E->cold->cand[q];
        if (p < seg1 || !E->cold->ccnt[q]) continue;
        if (C.pair_b[p] >= NROBOT + NB) continue;       // (box, wall): next pass
        bd_arb_update(C, E, q);
    }
    bp_wsync();
    #pragma unroll 1
    for (int i = lane; i < NB; i += BP_NL) {
        if (!E->body[1 + i].alive) continue;
        #pragma unroll 1
        for (int q = 0; q < nc; q++) {
            int p = E->cold->cand[q];
            if (!E->cold->ccnt[q] || C.pair_a[p] != NROBOT + i || C.pair_b[p] < NROBOT + NB) continue;
            bd_arb_update(C, E, q);
            bd_boundary_presolve(C, E, q);
        }
    }
    bp_wsync();
    // 5. cache expiry (cpSpaceArbiterSetFilter)
    int n_live = 0;
    #pragma unroll 1
    for (int base = 0; base < P.n_slots; base += BP_NL) {
        int p = base + lane;
        bool live = false;
        if (p < P.n_slots) {
            BdSlot& S = slots[p];
            if (S.state != SL_NONE && S.stamp != stamp) {
                int ticks = stamp - S.stamp;
                if (ticks >= 1 && S.state != SL_CACHED) S.state = SL_CACHED;
                if (ticks >= P.persistence) S.state = SL_NONE;
            }
            live = S.state != SL_NONE;
        }
        int tot; bp_wrank(live, &tot);
        n_live += tot;
    }
    if (lane == 0) E->n_live = n_live;
    // 6. solve list in canonical order, then greedy colours over the dynamic bodies (lane 0; <= a few dozen entries)
    if (lane == 0) {
        int n = 0;
        short list[MAXCAND]; unsigned char col[MAXCAND];
        unsigned used[1 + MAXBOX];
        #pragma unroll 1
        for (int b = 0; b <= NB; b++) used[b] = 0u;
        int maxc = 0;
        #pragma unroll 1
        for (int q = 0; q < nc; q++) {
            int p = E->cold->cand[q];
            if (!E->cold->ccnt[q] || (p >= seg0 && p < seg1)) continue;
            int sa = C.pair_a[p], sb = C.pair_b[p];
            int ba = sa < NROBOT ? 0 : 1 + sa - NROBOT;
            int bb = sb < NROBOT + NB ? 1 + sb - NROBOT : -1;
            unsigned m = (ba > 0 ? used[ba] : 0u) | (bb > 0 ? used[bb] : 0u);
            int c = __ffs(~m) - 1;
            if (c < 0 || c >= MAXCOLOUR) { c = MAXCOLOUR - 1; E->status |= BD_COLOUR_OVERFLOW; }
            if (ba > 0) used[ba] |= 1u << c;
            if (bb > 0) used[bb] |= 1u << c;
            list[n] = (short)p; col[n] = (unsigned char)c; n++;
            if (c > maxc) maxc = c;
        }
        int k = 0;
        #pragma unroll 1
        for (int c = 0; c <= maxc && n > 0; c++)
            #pragma unroll 1
            for (int i = 0; i < n; i++) if (col[i] == c) { E->cold->order[k] = list[i]; E->cold->cman[k].count = c; k++; }   // cman[k].count reused: colour of order[k]
        E->n_order = n;
        E->stat_cand += nc; E->stat_arb += n;
    }
    bp_wsync();
}

// Phase D: pre-step, velocity update, warm start, solver iterations.
BP_DEVNI void bd_phase_d(const BdConst& C, BdEnv* E) {
    const int lane = bp_lane();
    const BdParams& P = C.P;
    const int NB = P.NB;
    const double dt = P.dt;
    const int n_order = E->n_order;
    // 7. pre-step, velocity update (dynamic bodies only), warm start, iterations
    #pragma unroll 1
    for (int k = lane; k < n_order; k += BP_NL) bd_prestep(C, E, E->cold->order[k]);
    bp_wsync();
    #pragma unroll 1
    for (int b = 1 + lane; b <= NB; b += BP_NL) {
        BdBody& B = E->body[b];
        if (!B.alive) continue;
        B.vx = B.vx * P.damp + (0.0 + 0.0 * B.minv) * dt;
        B.vy = B.vy * P.damp + (0.0 + 0.0 * B.minv) * dt;
        B.vw = B.vw * P.damp + 0.0 * B.iinv * dt;
    }
    bp_wsync();
    const double dt_coef = (E->prev_dt == 0.0) ? 0.0 : dt / E->prev_dt;
    #pragma unroll 1
    for (int it = -1; it < P.iterations; it++) {
        int k0 = 0;
        #pragma unroll 1
        while (k0 < n_order) {
            int c = E->cold->cman[k0].count, k1 = k0;
            #pragma unroll 1
            while (k1 < n_order && E->cold->cman[k1].count == c) k1++;
            #pragma unroll 1
            for (int k = k0 + lane; k < k1; k += BP_NL) bd_solve(C, E, E->cold->order[k], it < 0, dt_coef);
            bp_wsync();
            k0 = k1;
        }
    }
    if (lane == 0) E->prev_dt = dt;
    bp_wsync();
}

// ------------------------------------------------------------------ one substep (cpSpaceStep) for one env, one warp
BP_DEVNI void bd_substep(const BdConst& C, BdEnv* E) {
    bd_phase_a(C, E);
    if (E->fast) return;
    bd_phase_b(C, E);
    bd_phase_c(C, E);
    bd_phase_d(C, E);
}

// space.remove(box.body, box): the body leaves the world, its arbiters are dropped
BP_DEV void bd_remove_box(const BdConst& C, BdEnv* E, int i) {
    E->body[1 + i].alive = 0;
    BdSlot* slots = bd_slots(E);
    for (int p = 0; p < C.P.n_slots; p++)
        if ((C.pair_a[p] == NROBOT + i || C.pair_b[p] == NROBOT + i) && slots[p].state != SL_NONE) { slots[p].state = SL_NONE; E->n_live--; }
}

// Space.point_query(p, 0, ShapeFilter()) restricted to the static shapes: bit (1 << ctype) per shape that strictly
// contains p (cpPolyShapePointQuery: distance = (outside ? d : -d) - r, reported when < 0)
BP_DEV int bd_point_query_static(const BdConst& C, const BdEnv* E, V2 p) {
    int mask = 0;
    const BdStatics& T = bd_room(C, E);
    #pragma unroll 1
    for (int k = 0; k < C.P.n_static; k++) {
        const BBox& bb = T.st_bb[k];
        if (!(p.x <= bb.r && bb.l <= p.x && p.y <= bb.t && bb.b <= p.y)) continue;
        int n = T.st_nv[k];
        if (n == 0) continue;
        V2 v0 = mk(T.st_v[k][n - 1][0], T.st_v[k][n - 1][1]);
        double minDist = INFINITY; bool outside = false;
        #pragma unroll 1
        for (int i = 0; i < n; i++) {
            V2 v1 = mk(T.st_v[k][i][0], T.st_v[k][i][1]);
            V2 nn = mk(T.st_n[k][i][0], T.st_n[k][i][1]);
            if (dot(nn, p - v1) > 0.0) outside = true;
            V2 d2 = v0 - v1;
            double t = clamp01(dot(d2, p - v1) / dot(d2, d2));
            V2 closest = v1 + d2 * t;
            V2 dd = p - closest;
            double dist = sqrt(dot(dd, dd));
            if (dist < minDist) minDist = dist;
            v0 = v1;
        }
        double dist = (outside ? minDist : -minDist) - T.st_rad[k];
        if (dist < 0.0) mask |= 1 << T.st_ctype[k];
    }
    return mask;
}

}  // namespace bd
}  // namespace bp
