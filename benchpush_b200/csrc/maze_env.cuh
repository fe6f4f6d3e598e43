// maze-NAMO-v0 environment logic on top of the rigid-body core.
// Follows MazeNAMO.reset / step of the reference (maze_NAMO_env.py:221-273, 325-359, 402-485) and the
// (1, 3) robot-wall collision handler (maze_NAMO_env.py:203-216): the maze walls are static pymunk
// Segments of radius 0.5 (sim_utils.py:174-181) = bodies [0, n_walls) of a layout with two vertices;
// the robot is ONE kinematic pymunk body with five shapes (chassis + 4 wheels, robot.py:77-118) =
// bodies [robot0, nbody), one record per shape carrying the same pose and velocity.
#pragma once
#include "shipice_env.cuh"

namespace bp {

// pre_solve of handler (1, 3): fires whenever the chassis shape touches a wall segment, although the
// pair is never solved (both bodies have infinite mass).  Chipmunk evaluates the pair when the AABBs
// intersect; before the first hit there is no cached arbiter, so GJK starts without a warm-start id.
// Executed by ONE lane after a substep (positions are those of the substep's collision phase).
BP_DEVNI bool maze_wall_touch(const EnvView& ev, const LayoutView& L, const SimParams& P, int r) {
    BBox B = ev.bb[r];
    for (int w = 0; w < P.n_walls; w++) {
        BBox A = ev.bb[w];
        if (!(A.l <= B.r && B.l <= A.r && A.b <= B.t && B.b <= A.t)) continue;
        Closest pts; pts.d = 0; pts.id = 0; pts.n = mk(0, 0);
        unsigned hint = 0; int iters = 0, capped = 0;
        Poly PA = body_poly(ev, L, w), PB = body_poly(ev, L, r);
        if (gjk_only(PA, PB, bb_centre(A), bb_centre(B), 0u, &pts, &hint, &iters, &capped)) pts = gjk_epa(PA, PB, hint, &capped);
        Manifold m; m.count = 0;
        make_manifold(ev, L, w, r, pts, &m);
        if (m.count > 0) return true;
    }
    return false;
}

// maze_NAMO_env.py:414-419 after every substep (one lane): bit1 = boundary constraint violated;
// the wall flag is sticky in the env block until the next reset.
BP_DEV int maze_substep_checks(const EnvView& ev, const LayoutView& L, const SimParams& P, int r) {
    int f = 0;
    V2 o = body_origin(ev, L, r);
    if (o.x < 0 || o.x > P.map_w) f |= 2;
    if (!ev.sc->wall_hit && maze_wall_touch(ev, L, P, r)) ev.sc->wall_hit = 1;
    return f;
}

// global_distance_map[int(y * s), int(x * s)] (maze_NAMO_env.py:395-400).  Negative indices wrap like
// numpy's; an index past the end (the reference raises IndexError) is clamped to the last cell.
BP_DEV double maze_distance_value(const SimParams& P, double x, double y) {
    int px = (int)(x * P.pix_scale), py = (int)(y * P.pix_scale);
    if (px < 0) px += P.dist_w;
    if (py < 0) py += P.dist_h;
    px = px < 0 ? 0 : (px >= P.dist_w ? P.dist_w - 1 : px);
    py = py < 0 ? 0 : (py >= P.dist_h ? P.dist_h - 1 : py);
    return P.dist_map[(size_t)py * P.dist_w + px];
}

// maze_NAMO_env.py:421-474: work, termination, reward, info.
BP_DEVNI void maze_step_finish(const EnvView& ev, const LayoutView& L, const SimParams& P, const Dims& d, int env, const StepOut& out,
                               bool boundary_violated) {
    EnvScal* sc = ev.sc;
    const int lane = bp_lane();
    const int r = sc->robot;
    V2 o = body_origin(ev, L, r);
    double work = step_work(ev, L);
    if (lane == 0) {
        sc->total_work += work;
        bool wall = sc->wall_hit != 0;
        double gx = o.x - P.goal_x, gy = o.y - P.goal_gy;
        bool reached = sqrt(gx * gx + gy * gy) <= P.goal_reach;
        bool terminated = reached || wall;
        double incr = RC(0.0);
        if (o.x != P.goal_x || o.y != P.goal_gy) {
            double dv = maze_distance_value(P, o.x, o.y);
            if (sc->has_prev_dist) incr = (sc->prev_dist - dv) * P.k_increment;
            sc->prev_dist = dv; sc->has_prev_dist = 1;
        }
        double collision_reward = -work;
        double reward = P.beta * collision_reward + incr;
        if (boundary_violated || wall) reward += P.boundary_penalty;
        bool success = false;
        if (terminated && !wall) { reward += P.terminal_reward; success = true; }
        out.reward[env] = reward; out.terminated[env] = terminated ? 1 : 0; out.truncated[env] = 0;
        out.pose[3 * env] = o.x; out.pose[3 * env + 1] = o.y; out.pose[3 * env + 2] = ev.body[r].pa;
        out.state[3 * env] = round2(o.x); out.state[3 * env + 1] = round2(o.y); out.state[3 * env + 2] = round2(ev.body[r].pa);
        out.total_work[env] = sc->total_work; out.collision_reward[env] = collision_reward;
        out.scaled_collision_reward[env] = collision_reward * P.beta; out.dist_reward[env] = incr;
        out.trial_success[env] = success ? 1 : 0; out.status[env] = sc->status;
    }
    bp_wsync();
}

// ---- env-kind dispatch used by the step loops (the kind is uniform over a launch)
BP_DEV int env_substep_checks(const EnvView& ev, const LayoutView& L, const SimParams& P, int r) {
    return P.kind == KIND_MAZE ? maze_substep_checks(ev, L, P, r) : shipice_substep_checks(ev, L, P, r);
}
BP_DEV void env_step_finish(const EnvView& ev, const LayoutView& L, const SimParams& P, const Dims& d, int env, const StepOut& out, int flags) {
    if (P.kind == KIND_MAZE) maze_step_finish(ev, L, P, d, env, out, (flags & 2) != 0);
    else shipice_step_finish(ev, L, P, d, env, out, (flags & 1) != 0, (flags & 2) != 0);
}

// One env, one warp: <Env>.step (the first design; the batched path of sim_batch.cuh is the default).
BP_DEVNI void shipice_step_env(const EnvView& ev, const LayoutTab& T, const SimParams& P, const Dims& d, double action, int env, const StepOut& out) {
    EnvScal* sc = ev.sc;
    LayoutView L = layout_view(T, d, sc->layout);
    const int r = sc->robot;
    shipice_step_begin(ev, L, P, d, action);
    int flags = 0;
    for (int k = 0; k < P.substeps; k++) {
        substep(ev, L, P, d, P.dt, P.dt);
        int f = 0;
        if (bp_lane() == 0) f = env_substep_checks(ev, L, P, r);
        flags |= bp_wbcast(f);
        bp_wsync();
    }
    env_step_finish(ev, L, P, d, env, out, flags);
}

}  // namespace bp
