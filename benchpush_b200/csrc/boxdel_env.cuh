// box-delivery-v0 environment logic on top of boxdel_core.cuh: reset, the position-control chain, the
// controller / until-still substep loops, rewards, work, termination.
//
// Follows BoxDeliveryEnv (reference benchpush/environments/box_delivery/box_delivery_env.py):
//   reset :571-632 (init_box_delivery_env :231-292, 1000 settle substeps :278-279),
//   step :634-830, controller :867-889 (DP.ideal_control dp.py:217-248, TargetCourse dp.py:63-117 with
//   metrics.py:176-181 path_length), execute_robot_path :891-988, step_simulation_until_still :990-1023,
//   shortest_path :1209-1264 (== position_controller.py:126-181), shortest_path_distance :1266-1268,
//   PositionController.get_waypoints_to_spatial_action position_controller.py:56-124.
// Third-party primitives restated: skimage.draw.line (Bresenham of skimage/draw/_draw.pyx::_line),
// skimage.measure.approximate_polygon (iterative Douglas-Peucker, skimage/measure/_polygon.py).
#pragma once
#include "boxdel_core.cuh"
#include "spfa.cuh"

namespace bp {
namespace bd {

constexpr double BD_PI = 3.141592653589793;
constexpr double MOVE_STEP_SIZE = 0.05;
constexpr double TURN_STEP_SIZE = 0.2617993877991494;          // np.radians(15)
constexpr double WAYPOINT_MOVING_THRESHOLD = 0.6;
constexpr double WAYPOINT_TURNING_THRESHOLD = 0.17453292519943295;   // np.radians(10)
constexpr double NOT_MOVING_THRESHOLD = 0.005;
constexpr double NONMOVEMENT_DIST_THRESHOLD = 0.05;
constexpr double NONMOVEMENT_TURN_THRESHOLD = 0.0008726646259971648; // np.radians(0.05)
constexpr int STEP_LIMIT = 10000;

struct BdOut {                    // device pointers owned by the caller (all optional)
    double* reward;               // [E]
    unsigned char* terminated;    // [E]
    unsigned char* truncated;     // [E]
    double* info;                 // [E][8]: state x, y, angle (2 dp), cumulative_distance, cumulative_reward, total_work, ministeps, work
    int* info_i;                  // [E][4]: cumulative_boxes, inactivity, boxes left, status
    unsigned char* completed;     // [E][NB] box_completed_statuses
    double* box_verts;            // [E][NB][4][2] info['obs'] (first `boxes left` rows valid)
};

// np.mod(h + pi, 2 pi) - pi
BP_DEV double np_mod(double a, double b) {
    double m = fmod(a, b);
    if (m != 0.0) { if ((b < 0) != (m < 0)) m += b; } else m = copysign(0.0, b);
    return m;
}
BP_DEV double restrict_heading(double h) { return np_mod(h + BD_PI, 2 * BD_PI) - BD_PI; }
BP_DEV double heading_diff(double a, double b) { return restrict_heading(a - b); }
BP_DEV double dist2(double ax, double ay, double bx, double by) { double dx = ax - bx, dy = ay - by; return sqrt(dx * dx + dy * dy); }
BP_DEV double sgn(double x) { return (x > 0) - (x < 0); }
BP_DEV double round2dp(double x) { return rint(x * 100.0) / 100.0; }

BP_DEV void pos_to_pixel(const BdParams& P, double x, double y, int* i, int* j) {
    int pi = (int)floor(P.H / 2.0 - y * P.ppm), pj = (int)floor(P.W / 2.0 + x * P.ppm);
    *i = pi < 0 ? 0 : (pi > P.H - 1 ? P.H - 1 : pi);
    *j = pj < 0 ? 0 : (pj > P.W - 1 ? P.W - 1 : pj);
}
BP_DEV void pixel_to_pos(const BdParams& P, int i, int j, double* x, double* y) {
    *x = (j - P.W / 2.0) / P.ppm; *y = (P.H / 2.0 - i) / P.ppm;
}
BP_DEV void closest_valid(const BdConst& C, int room, int i, int j, int* ci, int* cj) {
    size_t nv = (size_t)C.P.H * C.P.W;
    const int32_t* cl = bd_closest(C, room);
    *ci = cl[(size_t)i * C.P.W + j]; *cj = cl[nv + (size_t)i * C.P.W + j];
}

// skimage.draw.line iterator: calls f(r, c) for every pixel, both ends included
template <class F>
BP_DEV void sk_line(int r0, int c0, int r1, int c1, F f) {
    int steep = 0, r = r0, c = c0;
    int dr = abs(r1 - r0), dc = abs(c1 - c0);
    int sc = (c1 - c) > 0 ? 1 : -1, sr = (r1 - r) > 0 ? 1 : -1;
    if (dr > dc) { steep = 1; int t = c; c = r; r = t; t = dc; dc = dr; dr = t; t = sc; sc = sr; sr = t; }
    int d = 2 * dr - dc;
    for (int i = 0; i < dc; i++) {
        if (steep) f(c, r); else f(r, c);
        while (d >= 0) { r += sr; d -= 2 * dc; }
        c += sc; d += 2 * dr;
    }
    f(r1, c1);
}
BP_DEV bool line_free(const uint8_t* map, int W, int r0, int c0, int r1, int c1) {
    int blocked = 0;
    sk_line(r0, c0, r1, c1, [&](int r, int c) { if (!map[(size_t)r * W + c]) blocked++; });
    return blocked == 0;
}

// shortest_path_distance(box, receptacle) for every live box when the receptacle cell sees the whole free space
// (BdParams::recept_visible): path = [box position, receptacle position] => one Euclidean distance.
BP_DEV void bd_box_dists_visible(const BdConst& C, BdEnv* E, double* out) {
    for (int i = 0; i < C.P.NB; i++) {
        const BdBody& B = E->body[1 + i];
        if (B.alive != 1) continue;
        V2 o = bd_origin(B);
        out[i] = 0.0 + dist2(o.x, o.y, C.P.recept_x, C.P.recept_y);
    }
}

// ------------------------------------------------------------------ reset
BP_DEVNI void bd_reset_env(const BdConst& C, BdEnv* E, int layout, int episode) {
    const int lane = bp_lane();
    const BdParams& P = C.P;
    const BdLayout& L = C.layouts[layout];
    if (lane == 0) bd_link_env(E);
    bp_wsync();
    BdSlot* slots = bd_slots(E);
    #pragma unroll 1
    for (int p = lane; p < P.n_slots; p += BP_NL) { slots[p].state = SL_NONE; slots[p].cnt = 0; slots[p].stamp = 0; slots[p].gjk = 0; slots[p].ov = 0; slots[p].mvalid = 0; }
    if (lane == 0) {
        E->layout = layout; E->room = L.room; E->map_slot = C.maps_per_env ? (int)(((char*)E - C.envs) / C.env_bytes) : L.room; E->episode = episode; E->t = 0; E->status = 0; E->stamp = 0; E->n_order = 0; E->robot_hit = 0;
        E->inactivity = 0; E->cum_boxes = 0; E->done = 0; E->truncated = 0; E->n_path = 0; E->need_spfa = 0; E->n_live = 0; E->cleared_count = 0;
        E->prev_dt = 0.0; E->cum_distance = 0.0; E->cum_reward = 0.0; E->total_work = 0.0; E->last_work = 0.0; E->reward = 0.0; E->ministeps = 0.0;
        E->stat_substeps = 0; E->stat_cand = 0; E->stat_arb = 0;
        BdBody& R = E->body[0];
        R.px = L.start[0]; R.py = L.start[1]; R.pa = L.start[2];
        R.vx = R.vy = R.vw = R.bx = R.by = R.bw = 0.0; R.minv = R.iinv = 0.0; R.cgx = R.cgy = 0.0; R.alive = 1; R.idx = -1; R.cpx = R.cpy = R.cpa = NAN; R.moved = 1;
        bp_sincos(R.pa, &R.rs, &R.rc);
    }
    #pragma unroll 1
    for (int i = lane; i < P.NB; i += BP_NL) {
        BdBody& B = E->body[1 + i];
        B.alive = i < L.n_boxes; B.idx = i;
        E->cold->dist_ok[i] = 0;
        B.pa = L.box_init[i][2];
        bp_sincos(B.pa, &B.rs, &B.rc);
        B.cgx = L.box_cog[i][0]; B.cgy = L.box_cog[i][1];
        B.minv = L.box_minv[i]; B.iinv = L.box_iinv[i];
        B.vx = B.vy = B.vw = B.bx = B.by = B.bw = 0.0;
        B.cpx = B.cpy = B.cpa = NAN; B.moved = 1;
        bd_set_origin(B, mk(L.box_init[i][0], L.box_init[i][1]));
        E->completed[i] = 0;
    }
    bp_wsync();
    #pragma unroll 1
    for (int k = 0; k < P.settle_steps; k++) bd_substep(C, E);
    // prev_boxes = CostMap.get_obs_from_poly(self.boxes)
    #pragma unroll 1
    for (int i = lane; i < P.NB; i += BP_NL) {
        const BdBody& B = E->body[1 + i];
        #pragma unroll 1
        for (int v = 0; v < SV; v++) {
            V2 w = bd_local_to_world(B, L.box_lv[i][v][0], L.box_lv[i][v][1]);
            E->cold->prev_boxes[i][v][0] = w.x; E->cold->prev_boxes[i][v][1] = w.y;
        }
    }
    if (lane == 0) E->n_prev = L.n_boxes;
    bp_wsync();
}

// ------------------------------------------------------------------ step, part 1: action -> target, straight-line test
// Scalar; run by one lane.  action: pixel index (position control) or heading in [-1, 1].
BP_DEV void bd_begin_step(const BdConst& C, BdEnv* E, double action, double action2 = 0.0) {
    const BdParams& P = C.P;
    E->t++;
    E->robot_hit = 0;
    const BdBody& R = E->body[0];
    double x0 = R.px, y0 = R.py, h0 = restrict_heading(R.pa);
    E->init_pos[0] = x0; E->init_pos[1] = y0; E->init_heading = h0;
    if (P.action_type == 2 && P.kind == 1) {
        // area-clearing velocity control (area_clearing.py:656-663): yaw rate = max_yaw_rate_step * a[1] / 2 and surge =
        // target_speed * a[0] are float32 products (float32 action array, NumPy >= 2 promotion); the velocity VECTOR is set
        // once from the heading at step start and not refreshed during the cfg.sim.steps substeps that follow (D1)
        float w32 = ((float)((BD_PI / 2) / 15) * (float)action2) / 2.0f;
        float sv = (float)P.target_speed * (float)action;
        BdBody& Rb = E->body[0];
        double ch = cos(Rb.pa), sh = sin(Rb.pa);
        Rb.vw = (double)w32;
        Rb.vx = ch * (double)sv + (-sh) * 0.0; Rb.vy = sh * (double)sv + ch * 0.0;
        E->tgt[0] = (double)sv; E->tgt[1] = (double)w32;
        E->n_path = 0; E->need_spfa = 0; E->move_sign = 1; E->robot_distance = 0.0;
        return;
    }
    if (P.action_type == 2) {
        // velocity control (box_delivery_env.py:667-700): action = float32 (linear, angular); |linear| is clamped to target_speed
        // in float32 arithmetic (python float * np.float32 scalar -> float32 under NumPy >= 2 promotion)
        float lin = (float)action;
        if (fabsf(lin) >= (float)P.target_speed) lin = (float)P.target_speed * (float)sgn((double)lin);
        E->tgt[0] = (double)lin; E->tgt[1] = action2;
        E->n_path = 0; E->need_spfa = 0; E->move_sign = 1;
        if (P.kind == 0 && P.recept_visible) bd_box_dists_visible(C, E, E->cold->init_dist);
        return;
    }
    long long spatial;
    const int LW = P.LW;
    if (P.action_type == 1) {
        // box_delivery_env.py:706-723
        double angle = (action + 1) * BD_PI + BD_PI / 2;
        double xm = P.step_size * cos(angle), ym = P.step_size * sin(angle);
        int xp = (int)(LW / 2.0 + xm * P.ppm), yp = (int)(LW / 2.0 - ym * P.ppm);
        spatial = (long long)yp * LW + xp;
    } else spatial = (long long)action;
    int row = (int)(spatial / LW), col = (int)(spatial % LW);
    // position_controller.py:56-88
    double x_movement = -P.local_map_width / 2 + (double)col / P.ppm;
    double y_movement = P.local_map_width / 2 - (double)row / P.ppm;
    double sld = sqrt(x_movement * x_movement + y_movement * y_movement);
    double turn = atan2(-x_movement, y_movement);
    double slh = restrict_heading(h0 + turn);
    double tx = x0 + sld * cos(slh), ty = y0 + sld * sin(slh);
    double dxx = tx - x0, dyy = ty - y0;
    double ratio_x = 1, ratio_y = 1;
    double bound_x = sgn(tx) * P.bound_x / 2, bound_y = sgn(ty) * P.bound_y / 2;
    if (fabs(tx) > fabs(bound_x)) ratio_x = (bound_x - x0) / (tx - x0);
    if (fabs(ty) > fabs(bound_y)) ratio_y = (bound_y - y0) / (ty - y0);
    double ratio = fmin(ratio_x, ratio_y);
    tx = x0 + ratio * dxx; ty = y0 + ratio * dyy;
    E->tgt[0] = tx; E->tgt[1] = ty;
    // shortest_path(check_straight=True), first half
    int si, sj, ti, tj;
    pos_to_pixel(P, x0, y0, &si, &sj); pos_to_pixel(P, tx, ty, &ti, &tj);
    if (line_free(bd_cspace_thin(C, E->map_slot), P.W, si, sj, ti, tj)) {
        E->need_spfa = 0; E->n_path = 2;
        E->cold->path[0][0] = x0; E->cold->path[0][1] = y0; E->cold->path[1][0] = tx; E->cold->path[1][1] = ty;
    } else {
        E->need_spfa = 1;
        closest_valid(C, E->map_slot, si, sj, &E->src_i, &E->src_j);
        closest_valid(C, E->map_slot, ti, tj, &E->tgt_i, &E->tgt_j);
    }
    if (P.kind == 0 && P.recept_visible) bd_box_dists_visible(C, E, E->cold->init_dist);
}

// ------------------------------------------------------------------ path from a distance field (second half of shortest_path)
// dist = raw spfa_relax output for source (si, sj).  coords / chain / stk: scratch of H*W ints each.  Serial.
// Writes the waypoints (source first) to out[][2]; returns their count (0 = overflow).
BP_DEV int bd_extract_path(const BdConst& C, const uint8_t* cspace, const int* parents, const int* dist, int si, int sj, int ti, int tj,
                           double sx, double sy, double gx, double gy, int* coords, int* chain, int* stk, double (*out)[3], int max_out) {
    const BdParams& P = C.P;
    const int W = P.W, H = P.H;
    const int s = si * W + sj;
    // parents walk from the target (coords[0]) towards the source
    int n = 0, v = ti * W + tj;
    coords[n++] = v;
    while (v != s) {
        int u = parents ? parents[v] : spfa_parent_of(cspace, H, W, dist, s, v);
        if (u < 0) break;
        v = u; coords[n++] = v;
        if (n >= H * W) break;
    }
    // approximate_polygon(coords, tolerance=1)
    for (int i = 0; i < n; i++) chain[i] = 0;
    chain[0] = 1; chain[n - 1] = 1;
    int sp = 0;
    stk[sp++] = 0; stk[sp++] = n - 1;
    while (sp > 0) {
        int end = stk[--sp], start = stk[--sp];
        double r0 = coords[start] / W, c0 = coords[start] % W, r1 = coords[end] / W, c1 = coords[end] % W;
        double dr = r1 - r0, dc = c1 - c0;
        double ang = -atan2(dr, dc);
        double sa = sin(ang), ca = cos(ang);
        double segment_dist = c0 * sa + r0 * ca;
        double best = -INFINITY; int besti = -1; bool any = false;
        for (int k = start + 1; k < end; k++) {
            double r = coords[k] / W, c = coords[k] % W;
            double dr0 = r - r0, dc0 = c - c0, dr1 = r - r1, dc1 = c - c1;
            double pl0 = dr0 * dr + dc0 * dc, pl1 = -dr1 * dr - dc1 * dc;
            double d;
            if (pl0 > 0 && pl1 > 0) d = fabs(r * ca + c * sa - segment_dist);
            else d = fmin(sqrt(dc0 * dc0 + dr0 * dr0), sqrt(dc1 * dc1 + dr1 * dr1));
            if (d > 1.0) any = true;
            if (d > best) { best = d; besti = k; }      // np.argmax: first maximum
        }
        if (any) {
            // push (new_end, end) then (start, new_end): the latter is popped first
            stk[sp++] = besti; stk[sp++] = end;
            stk[sp++] = start; stk[sp++] = besti;
            chain[besti] = 1;
        }
    }
    int m = 0;
    for (int i = 0; i < n; i++) if (chain[i]) stk[m++] = coords[i];      // stk reused: the simplified chain
    // line-of-sight pruning on the configuration space
    int nn = 0;
    chain[nn++] = stk[0];
    for (int i = 1; i < m - 1; i++) {
        int a = chain[nn - 1], b = stk[i + 1];
        if (!line_free(cspace, W, a / W, a % W, b / W, b % W)) chain[nn++] = stk[i];
    }
    if (m > 1) chain[nn++] = stk[m - 1];
    // pixel -> position, reversed (source first)
    if (nn > max_out) return 0;
    if (nn < 2) {
        out[0][0] = sx; out[0][1] = sy; out[1][0] = gx; out[1][1] = gy;
        return 2;
    }
    for (int k = 0; k < nn; k++) {
        int c = chain[nn - 1 - k];
        pixel_to_pos(P, c / W, c % W, &out[k][0], &out[k][1]);
    }
    out[0][0] = sx; out[0][1] = sy; out[nn - 1][0] = gx; out[nn - 1][1] = gy;
    return nn;
}

// ------------------------------------------------------------------ waypoint headings, final-waypoint back-off
BP_DEV void bd_finalize_path(const BdConst& C, BdEnv* E) {
    int n = E->n_path;
    if (n < 2) return;          // velocity control: no path
    double (*wp)[3] = E->cold->path;
    wp[0][2] = 0.0;   // None
    for (int i = 1; i < n; i++) wp[i][2] = restrict_heading(atan2(wp[i][1] - wp[i - 1][1], wp[i][0] - wp[i - 1][0]));
    double d_end = dist2(wp[n - 2][0], wp[n - 2][1], wp[n - 1][0], wp[n - 1][1]);
    double signed_dist = d_end - C.P.robot_radius;
    int move_sign = (int)sgn(signed_dist);
    if (n > 2 && signed_dist < 0) {
        wp[n - 2][0] = wp[n - 1][0]; wp[n - 2][1] = wp[n - 1][1];
        wp[n - 2][2] = restrict_heading(atan2(wp[n - 2][1] - wp[n - 3][1], wp[n - 2][0] - wp[n - 3][0]));
        move_sign = 1;
    }
    E->move_sign = move_sign;
}

// ------------------------------------------------------------------ step, part 2: execute_robot_path + step_simulation_until_still
struct BdLoop { int flag; int dp_none; double sp_x, sp_y, al, pl; double cx[2], cy[2]; };

// controller() + apply_controller() of one loop iteration (box_delivery_env.py:867-889; dp.py:100-128, 217-248):
// DP over the course (wp[wi-1], wp[wi]); scalar, run by one lane; out of line (libm atan2 / sin / cos are large).
BP_DEVNI void bd_controller(const BdConst& C, BdEnv* E, BdLoop& LS, int wi, double pos_x, double pos_y, double heading, bool done_turning) {
    const BdParams& P = C.P;
    double (*wp)[3] = E->cold->path;
    if (LS.dp_none) {
        LS.cx[0] = wp[wi - 1][0]; LS.cy[0] = wp[wi - 1][1]; LS.cx[1] = wp[wi][0]; LS.cy[1] = wp[wi][1];
        double d0 = hypot(pos_x - LS.cx[0], pos_y - LS.cy[0]), d1 = hypot(pos_x - LS.cx[1], pos_y - LS.cy[1]);
        int ind = (d1 < d0) ? 1 : 0;
        // TargetCourse.init_setpoint: look ahead while Lfc exceeds the distance to the course point (dp.py:74-84)
        while (P.lfc > dist2(pos_x, pos_y, LS.cx[ind], LS.cy[ind])) { if (ind + 1 >= 2) break; ind++; }
        double ddx = LS.cx[1] - LS.cx[0], ddy = LS.cy[1] - LS.cy[0];
        LS.pl = sqrt(ddx * ddx + ddy * ddy);
        LS.al = LS.pl;
        LS.sp_x = LS.cx[ind]; LS.sp_y = LS.cy[ind];
        LS.dp_none = 0;
    }
    double theta_d = atan2(LS.sp_y - pos_y, LS.sp_x - pos_x);
    double theta_e = theta_d - heading;
    theta_e = atan2(sin(theta_e), cos(theta_e));
    double omega = 1.0 * theta_e / P.ctrl_dt;
    double ch = cos(heading), sh = sin(heading);
    double vx = ch * P.target_speed + (-sh) * 0.0, vy = sh * P.target_speed + ch * 0.0;
    LS.al = LS.al + P.target_speed * P.ctrl_dt;
    int ind = (LS.pl < LS.al) ? 1 : 0;
    LS.sp_x = LS.cx[ind]; LS.sp_y = LS.cy[ind];
    BdBody& R = E->body[0];
    R.vw = omega * P.omega_mul;
    if (!done_turning) { vx = vx * 0; vy = vy * 0; }
    R.vx = vx * P.v_mul; R.vy = vy * P.v_mul;
}

// A warp advances BD_MOVE_K environments, phase-major: every phase of the loop body (controller, the four substep phases,
// the exit tests) runs for all of the warp's live envs before the next phase starts.  The three loops of the reference
// (execute_robot_path, step_simulation_until_still / the fixed area-clearing substeps) are one state machine per env; the
// warps of a block meet at one barrier per iteration.  Loop state lives in shared memory (MoveVars), written by lane 0.
// Measured on B200 (profiles/r01e_room_kernels_envs_per_warp_sweep.txt, 16 384 area-clearing envs): K = 1 / 2 / 4 / 8 ->
// 28.4 k / 24.8 k / 22.5 k / 21.1 k env-steps/s.  The kernel is bound by instruction supply (~0.7 IPC per SM whatever the
// occupancy), but re-running a phase K times inside a warp does not help: the other 15 resident warps are in
// other phases and the ~110 KB of kernel code still does not stay in the instruction cache, while K > 1 leaves fewer warps
// to overlap.  Default K = 1; the knob stays for the phase-per-kernel design of DESIGN.md §8.
#if defined(__CUDACC__) && !defined(BP_HOST_EMU)
#define BP_BLOCK_ANY(p) __syncthreads_or(p)
#else
#define BP_BLOCK_ANY(p) (p)
#endif
#ifndef BD_MOVE_K
#define BD_MOVE_K 1
#endif
enum { MV_EXEC = 0, MV_STILL = 1, MV_POST = 2, MV_DONE = 3, MV_VEL = 4 };
struct MoveVars {
    int state, wi, sim_steps, post_k, n, done_turning, have_prev, still_done;
    double pos_x, pos_y, heading, robot_distance, prev_heading_diff, hd;
    double pvx[1 + MAXBOX], pvy[1 + MAXBOX];     // until-still: positions at the previous iteration ([NB] = robot)
    BdLoop LS;
};

BP_DEVNI void mv_pre(const BdConst& C, BdEnv* E, MoveVars& V) {
    const int lane = bp_lane();
    const BdParams& P = C.P;
    if (V.state == MV_EXEC) {
        if (lane == 0) {
            double (*wp)[3] = E->cold->path;
            double hd = heading_diff(V.heading, wp[V.wi][2]);
            V.hd = hd;
            if (!(fabs(hd) > TURN_STEP_SIZE && fabs(hd - V.prev_heading_diff) > 0.001)) V.done_turning = 1;
            bd_controller(C, E, V.LS, V.wi, V.pos_x, V.pos_y, V.heading, V.done_turning != 0);
        }
        bp_wsync();
    } else if (V.state == MV_VEL) {
        if (lane == 0) {
            // the robot keeps moving in the direction it is facing (box_delivery_env.py:687-690)
            BdBody& R = E->body[0];
            double ch = cos(R.pa), sh = sin(R.pa);
            R.vw = E->tgt[1];
            R.vx = ch * E->tgt[0] + (-sh) * 0.0; R.vy = sh * E->tgt[0] + ch * 0.0;
        }
        bp_wsync();
    } else if (V.state == MV_STILL) {
        const BdLayout& L = C.layouts[E->layout];
        // un-stick boxes from the walls
#pragma unroll 1
        for (int i = lane; i < P.NB; i += BP_NL) {
            BdBody& B = E->body[1 + i];
            if (!B.alive) continue;
            bool stuck = false;
#pragma unroll 1
            for (int v = 0; v < SV && !stuck; v++) {
                V2 w = bd_local_to_world(B, L.box_lv[i][v][0], L.box_lv[i][v][1]);
                if (bd_point_query_static(C, E, w) & (1 << 3)) stuck = true;
            }
            if (stuck) {
                V2 o = bd_origin(B);
                int pi, pj, ni, nj; pos_to_pixel(P, o.x, o.y, &pi, &pj);
                closest_valid(C, E->map_slot, pi, pj, &ni, &nj);
                double nx, ny; pixel_to_pos(P, ni, nj, &nx, &ny);
                bd_set_origin(B, mk(nx, ny));
                B.vx = 0.0; B.vy = 0.0;
            }
        }
        bp_wsync();
        bool moved = false;
        const bool have_prev = V.have_prev != 0;
#pragma unroll 1
        for (int b = lane; b <= P.NB; b += BP_NL) {
            const BdBody& B = E->body[b == P.NB ? 0 : 1 + b];
            if (b < P.NB && !B.alive) continue;
            V2 o = bd_origin(B);
            if (have_prev && dist2(V.pvx[b], V.pvy[b], o.x, o.y) > NOT_MOVING_THRESHOLD) moved = true;
            V.pvx[b] = o.x; V.pvy[b] = o.y;
        }
        bool any = bp_wany(moved);
        if (lane == 0) { V.still_done = (have_prev && !any) ? 1 : 0; V.have_prev = 1; }
        bp_wsync();
    }
}

BP_DEVNI void mv_post(const BdConst& C, BdEnv* E, MoveVars& V) {
    const BdParams& P = C.P;
    if (bp_lane() == 0) {
        if (V.state == MV_EXEC) {
            double (*wp)[3] = E->cold->path;
            const int wi = V.wi;
            double pos_x = E->body[0].px, pos_y = E->body[0].py, heading = restrict_heading(E->body[0].pa);
            V.pos_x = pos_x; V.pos_y = pos_y; V.heading = heading;
            V.prev_heading_diff = V.hd;
            bool stop = false;
            if (dist2(wp[wi - 1][0], wp[wi - 1][1], pos_x, pos_y) > MOVE_STEP_SIZE && E->robot_hit) stop = true;
            if (!stop && dist2(pos_x, pos_y, wp[wi][0], wp[wi][1]) < WAYPOINT_MOVING_THRESHOLD && fabs(heading - wp[wi][2]) < WAYPOINT_TURNING_THRESHOLD) {
                V.robot_distance += dist2(wp[wi - 1][0], wp[wi - 1][1], pos_x, pos_y);
                if (wi == V.n - 1) stop = true;
                else { V.wi = wi + 1; V.done_turning = 0; V.LS.dp_none = 1; }
            }
            if (!stop) {
                V.sim_steps++;
                if (V.sim_steps > P.step_limit) { E->status |= BD_STEP_LIMIT; stop = true; }
            }
            if (stop) {
                E->robot_distance = V.robot_distance;
                V.sim_steps = 0;
                // area-clearing: a fixed number of substeps follows the path (area_clearing.py:692-693); box-delivery: until still
                V.state = (P.kind == 1) ? (P.post_steps > 0 ? MV_POST : MV_DONE) : MV_STILL;
            }
        } else if (V.state == MV_VEL) {
            V.sim_steps++;
            if (E->robot_hit || V.sim_steps >= P.post_steps) {
                E->robot_distance = dist2(E->init_pos[0], E->init_pos[1], E->body[0].px, E->body[0].py);
                V.sim_steps = 0;
                V.state = MV_STILL;
            }
        } else if (V.state == MV_STILL) {
            // the reference steps once more after deciding that everything is still (:1007-1019)
            if (V.still_done) V.state = MV_DONE;
            else {
                V.sim_steps++;
                if (V.sim_steps > STEP_LIMIT) { E->status |= BD_STEP_LIMIT; V.state = MV_DONE; }
            }
        } else if (V.state == MV_POST) {
            if (++V.post_k >= P.post_steps) V.state = MV_DONE;
        }
    }
    bp_wsync();
}

// envs [env0, env0 + nk) of this warp; V[k] in shared memory.  Every warp of the block calls this (nk may be 0).
// hot != nullptr: the env headers are copied to hot[k * hot_stride] (shared memory) for the launch and written back at the end.
BP_DEVNI void bd_move(const BdConst& C, int env0, int nk, const uint8_t* mask, MoveVars* V, char* hot, int hot_stride) {
    const int lane = bp_lane();
    BdEnv* Es[BD_MOVE_K];
#pragma unroll 1
    for (int k = 0; k < nk; k++) {
        BdEnv* G = bd_env(C, env0 + k);
        bool part = !(mask && !mask[env0 + k]);
        Es[k] = G;
        if (part && hot) {
            BdEnv* Hh = (BdEnv*)(hot + (size_t)k * hot_stride);
            const D2* src = (const D2*)G; D2* dst = (D2*)Hh;          // 16-byte moves
#pragma unroll 1
            for (int i = lane; i < (int)(sizeof(BdEnv) / 16); i += BP_NL) dst[i] = src[i];
            Es[k] = Hh;
        }
        bp_wsync();
        if (lane == 0) {
            MoveVars& v = V[k];
            v.state = part ? (C.P.action_type == 2 ? (C.P.kind == 1 ? (C.P.post_steps > 0 ? MV_POST : MV_DONE) : MV_VEL) : MV_EXEC) : MV_DONE;
            if (part) {
                BdEnv* E = Es[k];
                bd_finalize_path(C, E);
                v.wi = 1; v.sim_steps = 0; v.post_k = 0; v.n = E->n_path; v.done_turning = 0; v.have_prev = 0; v.still_done = 0;
                v.pos_x = E->init_pos[0]; v.pos_y = E->init_pos[1]; v.heading = E->init_heading;
                v.robot_distance = 0.0; v.prev_heading_diff = 0.0; v.hd = 0.0; v.LS.dp_none = 1;
            }
        }
    }
    bp_wsync();
#pragma unroll 1
    for (;;) {
        unsigned act = 0;
#pragma unroll 1
        for (int k = 0; k < nk; k++) if (V[k].state != MV_DONE) act |= 1u << k;
        if (!BP_BLOCK_ANY(act != 0)) break;
        if (!act) continue;
#pragma unroll 1
        for (int k = 0; k < nk; k++) if (act >> k & 1) mv_pre(C, Es[k], V[k]);
#pragma unroll 1
        for (int k = 0; k < nk; k++) if (act >> k & 1) bd_phase_a(C, Es[k]);
#pragma unroll 1
        for (int k = 0; k < nk; k++) if ((act >> k & 1) && !Es[k]->fast) bd_phase_b(C, Es[k]);
#pragma unroll 1
        for (int k = 0; k < nk; k++) if ((act >> k & 1) && !Es[k]->fast) bd_phase_c(C, Es[k]);
#pragma unroll 1
        for (int k = 0; k < nk; k++) if ((act >> k & 1) && !Es[k]->fast) bd_phase_d(C, Es[k]);
#pragma unroll 1
        for (int k = 0; k < nk; k++) if (act >> k & 1) mv_post(C, Es[k], V[k]);
    }
    bp_wsync();
    if (hot) {
#pragma unroll 1
        for (int k = 0; k < nk; k++) {
            BdEnv* G = bd_env(C, env0 + k);
            if (Es[k] == G) continue;
            const D2* src = (const D2*)Es[k]; D2* dst = (D2*)G;
#pragma unroll 1
            for (int i = lane; i < (int)(sizeof(BdEnv) / 16); i += BP_NL) dst[i] = src[i];
        }
        bp_wsync();
    }
}

// ------------------------------------------------------------------ step, part 3: rewards, removal, work, termination
// Scalar; run by one lane.
BP_DEV void ac_finish_step(const BdConst& C, BdEnv* E, int env, const BdOut& out);
BP_DEV void bd_finish_step(const BdConst& C, BdEnv* E, int env, const BdOut& out) {
    const BdParams& P = C.P;
    if (P.kind == 1) { ac_finish_step(C, E, env, out); return; }
    const BdLayout& L = C.layouts[E->layout];
    double reward = 0.0;
    int robot_boxes = 0;
    if (P.recept_visible) bd_box_dists_visible(C, E, E->cold->final_dist);
    for (int i = 0; i < P.NB; i++) {
        BdBody& B = E->body[1 + i];
        if (!B.alive) continue;
        double dist_moved = E->cold->init_dist[i] - E->cold->final_dist[i];
        if (P.use_correct && dist_moved > 0) dist_moved *= P.correct_scale;
        reward += P.partial_scale * dist_moved;
        bool inside = true;
        for (int v = 0; v < SV && inside; v++) {
            V2 w = bd_local_to_world(B, L.box_lv[i][v][0], L.box_lv[i][v][1]);
            if (!(bd_point_query_static(C, E, w) & (1 << 4))) inside = false;
        }
        if (inside) {
            E->completed[B.idx] = 1; E->inactivity = 0; robot_boxes++; reward += P.goal_reward;
            B.alive = 2;          // marked; removed below (after the loop, as the reference does)
        }
    }
    for (int i = 0; i < P.NB; i++) if (E->body[1 + i].alive == 2) bd_remove_box(C, E, i);
    if (E->robot_hit) reward -= P.collision_penalty;
    const BdBody& R = E->body[0];
    double heading = restrict_heading(R.pa);
    double turn = heading_diff(E->init_heading, heading);
    if (E->robot_distance < NONMOVEMENT_DIST_THRESHOLD && fabs(turn) < NONMOVEMENT_TURN_THRESHOLD) reward -= P.nonmove_penalty;
    E->cum_distance += E->robot_distance; E->cum_boxes += robot_boxes; E->cum_reward += reward;
    // work = total_work_done(prev_boxes, updated_boxes): zip pairs the k-th previous polygon with the k-th live box (D10)
    double work = 0.0;
    int k = 0;
    double upd[MAXBOX][SV][2];
    for (int i = 0; i < P.NB; i++) {
        const BdBody& B = E->body[1 + i];
        if (B.alive != 1) continue;
        for (int v = 0; v < SV; v++) { V2 w = bd_local_to_world(B, L.box_lv[i][v][0], L.box_lv[i][v][1]); upd[k][v][0] = w.x; upd[k][v][1] = w.y; }
        k++;
    }
    int nz = k < E->n_prev ? k : E->n_prev;
    for (int q = 0; q < nz; q++) {
        double A, ax, ay, A2, bx, by;
        area_centroid_abs(&E->cold->prev_boxes[q][0][0], SV, &A, &ax, &ay);
        area_centroid_abs(&upd[q][0][0], SV, &A2, &bx, &by);
        double dx = ax - bx, dy = ay - by;
        work += sqrt(dx * dx + dy * dy) * A;
    }
    for (int q = 0; q < k; q++) for (int v = 0; v < SV; v++) { E->cold->prev_boxes[q][v][0] = upd[q][v][0]; E->cold->prev_boxes[q][v][1] = upd[q][v][1]; }
    E->n_prev = k;
    E->total_work += work; E->last_work = work;
    if (robot_boxes == 0) E->inactivity++;
    int terminated = (E->cum_boxes == P.num_boxes), truncated = 0;
    if (E->inactivity >= P.inactivity_cutoff) { terminated = 1; truncated = 1; }
    E->done = terminated; E->truncated = truncated; E->reward = reward;
    E->ministeps = E->robot_distance / P.ministep;
    if (out.reward) out.reward[env] = reward;
    if (out.terminated) out.terminated[env] = (unsigned char)terminated;
    if (out.truncated) out.truncated[env] = (unsigned char)truncated;
    if (out.info) {
        double* o = out.info + (size_t)env * 8;
        o[0] = round2dp(R.px); o[1] = round2dp(R.py); o[2] = round2dp(R.pa);
        o[3] = E->cum_distance; o[4] = E->cum_reward; o[5] = E->total_work; o[6] = E->ministeps; o[7] = work;
    }
    if (out.info_i) { int* o = out.info_i + (size_t)env * 4; o[0] = E->cum_boxes; o[1] = E->inactivity; o[2] = k; o[3] = E->status; }
    if (out.completed) for (int i = 0; i < P.NB; i++) out.completed[(size_t)env * P.NB + i] = E->completed[i];
    if (out.box_verts) {
        double* o = out.box_verts + (size_t)env * P.NB * SV * 2;
        for (int q = 0; q < k; q++) for (int v = 0; v < SV; v++) { o[(q * SV + v) * 2] = upd[q][v][0]; o[(q * SV + v) * 2 + 1] = upd[q][v][1]; }
    }
}
// ------------------------------------------------------------------ area-clearing-v0: completion, rewards, truncation
// shapely Polygon.intersects for two convex polygons: closed sets, touching counts (separating-axis test)
BP_DEV bool convex_intersects(const double (*A)[2], int na, const double (*B)[2], int nb) {
    for (int pass = 0; pass < 2; pass++) {
        const double (*P)[2] = pass ? B : A; int np = pass ? nb : na;
        const double (*Q)[2] = pass ? A : B; int nq = pass ? na : nb;
        for (int i = 0; i < np; i++) {
            int j = (i + 1) % np;
            double ex = P[j][0] - P[i][0], ey = P[j][1] - P[i][1];
            double nx = ey, ny = -ex;
            double amax = -INFINITY, amin = INFINITY, bmax = -INFINITY, bmin = INFINITY;
            for (int k = 0; k < np; k++) { double d = nx * P[k][0] + ny * P[k][1]; amax = fmax(amax, d); amin = fmin(amin, d); }
            for (int k = 0; k < nq; k++) { double d = nx * Q[k][0] + ny * Q[k][1]; bmax = fmax(bmax, d); bmin = fmin(bmin, d); }
            if (amax < bmin || bmax < amin) return false;
        }
    }
    return true;
}
BP_DEV double min_goal_dist(const BdParams& P, double cx, double cy) {
    double m = INFINITY;
    for (int g = 0; g < P.n_goal; g++) m = fmin(m, dist2(cx, cy, P.goal[g][0], P.goal[g][1]));
    return m;
}
// AreaClearingEnv.step after the substeps (area_clearing.py:695-778): boxes_completed (:1122-1140), obs_to_goal_difference
// (evaluation/metrics.py:73-94, centroids abs-folded by poly_centroid: D4), total_work_done, reward, truncation clock (D13).
// Scalar; run by one lane.
BP_DEV void ac_finish_step(const BdConst& C, BdEnv* E, int env, const BdOut& out) {
    const BdParams& P = C.P;
    const BdLayout& L = C.layouts[E->layout];
    const int NB = L.n_boxes;
    double upd[MAXBOX][SV][2];
    for (int i = 0; i < NB; i++)
        for (int v = 0; v < SV; v++) { V2 w = bd_local_to_world(E->body[1 + i], L.box_lv[i][v][0], L.box_lv[i][v][1]); upd[i][v][0] = w.x; upd[i][v][1] = w.y; }
    double collision_penalty = E->robot_hit ? P.r_boundary : 0.0;
    int num_completed = 0;
    for (int i = 0; i < NB; i++) {
        bool outside = !convex_intersects(P.bnd, P.n_bnd, upd[i], SV);
        num_completed += outside ? 1 : 0;
        E->completed[i] = outside ? 1 : 0;
    }
    bool all_completed = (num_completed == P.num_boxes);
    double diff_reward = 0.0, work = 0.0;
    for (int i = 0; i < NB; i++) {
        double A, ax, ay, A2, bx, by;
        area_centroid_abs(&E->cold->prev_boxes[i][0][0], SV, &A, &ax, &ay);
        area_centroid_abs(&upd[i][0][0], SV, &A2, &bx, &by);
        if (convex_intersects(P.bnd, P.n_bnd, E->cold->prev_boxes[i], SV)) diff_reward += min_goal_dist(P, ax, ay) - min_goal_dist(P, bx, by);
        double dx = ax - bx, dy = ay - by;
        work += sqrt(dx * dx + dy * dy) * A;
    }
    double pushing_reward = diff_reward * P.r_push_mult;
    double box_completion_reward;
    int delta = num_completed - E->cleared_count; if (delta < 0) delta = -delta;
    if (num_completed > E->cleared_count) { box_completion_reward = delta * P.r_cleared; E->t = 0; }
    else box_completion_reward = delta * P.r_putback;
    E->cleared_count = num_completed;
    const BdBody& R = E->body[0];
    double heading = restrict_heading(R.pa);
    double turn = heading_diff(E->init_heading, heading);
    double nonmove = (E->robot_distance < NONMOVEMENT_DIST_THRESHOLD && fabs(turn) < NONMOVEMENT_TURN_THRESHOLD) ? P.r_nonmove : 0.0;
    E->total_work += work; E->last_work = work;
    for (int i = 0; i < NB; i++) for (int v = 0; v < SV; v++) { E->cold->prev_boxes[i][v][0] = upd[i][v][0]; E->cold->prev_boxes[i][v][1] = upd[i][v][1]; }
    int terminated = all_completed ? 1 : 0;
    double reward = box_completion_reward + collision_penalty + pushing_reward + nonmove;
    int truncated = E->t >= P.t_max;
    if (truncated) reward += P.r_truncation; else if (terminated) reward += P.r_terminal;
    E->done = terminated; E->truncated = truncated; E->reward = reward; E->cum_boxes = num_completed; E->cum_reward += reward;
    E->ministeps = E->robot_distance / 2.5;
    E->robot_hit = 0;
    if (out.reward) out.reward[env] = reward;
    if (out.terminated) out.terminated[env] = (unsigned char)terminated;
    if (out.truncated) out.truncated[env] = (unsigned char)truncated;
    if (out.info) {
        double* o = out.info + (size_t)env * 8;
        o[0] = round2dp(R.px); o[1] = round2dp(R.py); o[2] = round2dp(R.pa);
        o[3] = diff_reward; o[4] = box_completion_reward; o[5] = E->total_work; o[6] = E->ministeps; o[7] = work;
    }
    if (out.info_i) { int* o = out.info_i + (size_t)env * 4; o[0] = num_completed; o[1] = E->t; o[2] = NB; o[3] = E->status; }
    if (out.completed) for (int i = 0; i < P.NB; i++) out.completed[(size_t)env * P.NB + i] = E->completed[i];
    if (out.box_verts) {
        double* o = out.box_verts + (size_t)env * P.NB * SV * 2;
        for (int q = 0; q < NB; q++) for (int v = 0; v < SV; v++) { o[(q * SV + v) * 2] = upd[q][v][0]; o[(q * SV + v) * 2 + 1] = upd[q][v][1]; }
    }
}

// info after reset (box_delivery_env.py:618-630)
BP_DEV void bd_reset_out(const BdConst& C, BdEnv* E, int env, const BdOut& out) {
    const BdParams& P = C.P;
    const BdBody& R = E->body[0];
    if (out.reward) out.reward[env] = 0.0;
    if (out.terminated) out.terminated[env] = 0;
    if (out.truncated) out.truncated[env] = 0;
    if (out.info) { double* o = out.info + (size_t)env * 8; o[0] = round2dp(R.px); o[1] = round2dp(R.py); o[2] = round2dp(R.pa); o[3] = o[4] = o[5] = o[6] = o[7] = 0.0; }
    if (out.info_i) { int* o = out.info_i + (size_t)env * 4; o[0] = 0; o[1] = 0; o[2] = E->n_prev; o[3] = E->status; }
    if (out.completed) for (int i = 0; i < P.NB; i++) out.completed[(size_t)env * P.NB + i] = 0;
    if (out.box_verts) {
        double* o = out.box_verts + (size_t)env * P.NB * SV * 2;
        for (int q = 0; q < E->n_prev; q++) for (int v = 0; v < SV; v++) { o[(q * SV + v) * 2] = E->cold->prev_boxes[q][v][0]; o[(q * SV + v) * 2 + 1] = E->cold->prev_boxes[q][v][1]; }
    }
}

}  // namespace bd
}  // namespace bp
