// ship-ice-v0 environment logic on top of the rigid-body core (one warp = one environment).
// Follows ShipIceEnv.reset / step of the reference (ship_ice_env.py:186-220, 223-249, 261-355),
// CostMap.get_obs_from_poly (cost_map.py:274-281), total_work_done (evaluation/metrics.py:96-126),
// poly_area / poly_centroid (geometry/polygon.py:25-41).
#pragma once
#include "sim_step.cuh"

namespace bp {

// Per-env outputs of one env step (SoA over envs; device pointers owned by the caller).
struct StepOut {
    double* reward;             // [E]
    unsigned char* terminated;  // [E]
    unsigned char* truncated;   // [E] always 0 (the reference never truncates: ship_ice_env.py:355)
    double* state;              // [E][3] ship (x, y, theta) rounded to 2 decimals (info['state'])
    double* pose;               // [E][3] unrounded ship pose
    double* total_work;         // [E]
    double* collision_reward;   // [E]  (= -work)
    double* scaled_collision_reward;   // [E]
    double* dist_reward;        // [E]
    unsigned char* trial_success;      // [E]
    int* status;                // [E] StatusBits
};

// geometry/polygon.py:25-41 on the world-space hull vertices of body b (np.roll(y, 1)[i] = y[i-1]).
// Always accumulated in double (also in the fp32-state build): the shoelace sums cancel from ~1e3 m^2 down to a floe area
// below 1 m^2, and the work term differences centroids that move by millimetres per step.
BP_DEVNI void area_centroid_abs(const real* wv, int n, double* area, double* cx, double* cy) {
    double s1 = 0.0, s2 = 0.0;
    for (int i = 0; i < n; i++) {
        int j = (i == 0) ? n - 1 : i - 1;
        s1 += (double)wv[2 * i] * (double)wv[2 * j + 1];
        s2 += (double)wv[2 * i + 1] * (double)wv[2 * j];
    }
    double A = 0.5 * fabs(s1 - s2);
    double sx = 0.0, sy = 0.0;
    for (int i = 0; i < n; i++) {
        int j = (i == 0) ? n - 1 : i - 1;
        double xi = wv[2 * i], yi = wv[2 * i + 1], xj = wv[2 * j], yj = wv[2 * j + 1];
        double u = xi * yj - xj * yi;
        sx += (xi + xj) * u;
        sy += (yi + yj) * u;
    }
    double k = 1.0 / (6.0 * A);
    *area = A; *cx = fabs(k * sx); *cy = fabs(k * sy);
}

BP_DEV double round2(double x) { return rint(x * RC(100.0)) / RC(100.0); }

// True when a substep can change nothing: the ship is the only awake body and does not move, and
// no arbiter is touching or cached (so no stamp-based expiry is pending either).
BP_DEV bool settle_is_idle(const EnvView& ev, const EnvScal* sc) {
    // bodies that are not dynamic never leave the awake list, so "only they are awake" is a count
    if (sc->n_awake != sc->n_fixed || sc->n_order != 0 || sc->n_cached != 0 || sc->dirty || sc->rebuild) return false;
    const BodyRec& R = ev.body[sc->robot];
    return R.vx == RC(0.0) && R.vy == RC(0.0) && R.vw == RC(0.0) && R.bx == RC(0.0) && R.by == RC(0.0) && R.bw == RC(0.0);
}

BP_DEVNI void shipice_reset_env(const EnvView& ev, const LayoutTab& T, const SimParams& P, const Dims& d, int layout_id, int env, const StepOut& out,
                                const double* start = nullptr) {
    LayoutView L = layout_view(T, d, layout_id);
    EnvScal* sc = ev.sc;
    const int lane = bp_lane();
    init_env(ev, L, P, d, layout_id, start);
    // 1000 settle substeps with the ship at rest (ship_ice_env.py:218-219)
    real prev_dt = RC(0.0);
    for (int k = 0; k < P.settle_steps; k++) {
        substep(ev, L, P, d, P.dt, prev_dt);
        prev_dt = P.dt;
        if (k == 0 || (k & 127) == 127) sleep_pass(ev, L, d);
        // Everything asleep, the ship at rest, no arbiter alive or cached: every remaining settle
        // substep is exactly a no-op (nothing integrates, nothing is evaluated or solved), only the
        // step counter advances — skip them.
        if (settle_is_idle(ev, sc)) {
            if (lane == 0) sc->stamp += P.settle_steps - 1 - k;
            bp_wsync();
            break;
        }
    }
    sleep_pass(ev, L, d);
    // prev_obs = get_obs_from_poly(polygons) (ship_ice_env.py:220): cache |centroid| and area per floe
    int nb = sc->nbody;
    for (int b = lane; b < nb; b += BP_NL) {
        if (body_dynamic(L, b)) area_centroid_abs(ev.wv + (size_t)b * VMAX * 2, L.nv[b], &ev.area[b], &ev.cenx[b], &ev.ceny[b]);
        ev.moved_step[b] = 0;
    }
    if (lane == 0) {
        int r = sc->robot;
        V2 o = body_origin(ev, L, r);
        out.pose[3 * env] = o.x; out.pose[3 * env + 1] = o.y; out.pose[3 * env + 2] = ev.body[r].pa;
        out.state[3 * env] = round2(o.x); out.state[3 * env + 1] = round2(o.y); out.state[3 * env + 2] = round2(ev.body[r].pa);
        out.total_work[env] = RC(0.0); out.reward[env] = RC(0.0); out.terminated[env] = 0; out.truncated[env] = 0;
        out.collision_reward[env] = RC(0.0); out.scaled_collision_reward[env] = RC(0.0); out.dist_reward[env] = RC(0.0);
        out.trial_success[env] = 0; out.status[env] = sc->status;
        sc->t = 0; sc->total_work = RC(0.0);
        // maze: wall_collision and prev_dist_value are cleared after the settle steps (maze_NAMO_env.py:337-341)
        sc->wall_hit = 0; sc->has_prev_dist = 0; sc->prev_dist = RC(0.0);
    }
    bp_wsync();
}

// Put env into an externally given body state (bp_*_set_bodies): poses and velocities from bodies6[nbody][6], empty arbiter
// cache (like a pymunk Space whose cached arbiters were dropped), every body awake, neighbour list rebuilt, the work
// accounting (prev_obs) restarted from the new poses.  Episode counters, layout and accumulated work are kept.
BP_DEVNI void shipice_set_bodies_env(const EnvView& ev, const LayoutTab& T, const SimParams& P, const Dims& d, const double* bodies6) {
    EnvScal* sc = ev.sc;
    const int lane = bp_lane();
    const int layout = sc->layout, t = sc->t, episode = sc->episode, first = sc->first_layout, done = sc->done, obs_raw = sc->obs_raw;
    const int wall_hit = sc->wall_hit, has_prev = sc->has_prev_dist;
    const double total_work = sc->total_work, prev_dist = sc->prev_dist;
    bp_wsync();
    LayoutView L = layout_view(T, d, layout);
    init_env(ev, L, P, d, layout, nullptr, bodies6);
    int nb = sc->nbody;
    for (int b = lane; b < nb; b += BP_NL) {
        if (body_dynamic(L, b)) area_centroid_abs(ev.wv + (size_t)b * VMAX * 2, L.nv[b], &ev.area[b], &ev.cenx[b], &ev.ceny[b]);
        ev.moved_step[b] = 0;
    }
    if (lane == 0) {
        sc->t = t; sc->episode = episode; sc->first_layout = first; sc->done = done; sc->obs_raw = 0; (void)obs_raw;
        sc->total_work = total_work; sc->wall_hit = wall_hit; sc->has_prev_dist = has_prev; sc->prev_dist = prev_dist;
        // Nothing may be put to sleep before one substep has run on the injected state: a resting pair that overlaps deeper
        // than the collision slop has no arbiter (hence no bias impulse) yet, so the sleep criterion cannot see that it is
        // about to move.  The next env step keeps every body awake (like the first settle substep after a reset).
        sc->pad_hw = 1;
    }
    bp_wsync();
}

// ShipIceEnv.reset through the per-layout snapshot (DevConst::snap): identical to shipice_reset_env
// because the post-reset state is a pure function of the layout.
BP_DEVNI void shipice_reset_from_snapshot(const DevConst* C, char* base, int layout_id, int env, const StepOut& out) {
    const int lane = bp_lane();
    struct alignas(16) V16 { unsigned long long a, b; };      // 128-bit loads / stores
    const V16* src = (const V16*)(C->snap + C->env_bytes * (size_t)layout_id);
    V16* dst = (V16*)base;
    const size_t n = C->env_bytes / sizeof(V16);
#pragma unroll 4
    for (size_t i = lane; i < n; i += BP_NL) dst[i] = src[i];
    if (lane == 0) {
        const double* p = C->snap_pose + 3 * layout_id;
        out.pose[3 * env] = p[0]; out.pose[3 * env + 1] = p[1]; out.pose[3 * env + 2] = p[2];
        out.state[3 * env] = round2(p[0]); out.state[3 * env + 1] = round2(p[1]); out.state[3 * env + 2] = round2(p[2]);
        out.total_work[env] = RC(0.0); out.reward[env] = RC(0.0); out.terminated[env] = 0; out.truncated[env] = 0;
        out.collision_reward[env] = RC(0.0); out.scaled_collision_reward[env] = RC(0.0); out.dist_reward[env] = RC(0.0);
        out.trial_success[env] = 0; out.status[env] = C->snap_status[layout_id];
    }
    bp_wsync();
}

// ship_ice_env.py:263-274: yaw rate from the action, constant surge along the heading at step start
BP_DEVNI void shipice_step_begin(const EnvView& ev, const LayoutView& L, const SimParams& P, const Dims& d, double action) {
    EnvScal* sc = ev.sc;
    const int r = sc->robot;
    if (bp_lane() == 0) {
        sc->t++;
        real a = ev.body[r].pa;
        real s, c; bp_sincos(a, &s, &c);
        real vw = action * P.yaw_step;
        real vx = c * P.target_speed + (-s) * RC(0.0);
        real vy = s * P.target_speed + c * RC(0.0);
        // every shape of the kinematic robot is one body record with the same pose and velocity
        for (int q = r; q < sc->nbody; q++) { ev.body[q].vw = vw; ev.body[q].vx = vx; ev.body[q].vy = vy; }
    }
    bp_wsync();
    if (sc->pad_hw) {                 // state injected by bp_*_set_bodies: this step runs with every body awake
        bp_wsync();
        if (bp_lane() == 0) sc->pad_hw = 0;
        bp_wsync();
    } else {
        sleep_pass(ev, L, d);
    }
}

// ship_ice_env.py:284-290, after every substep: yaw limit (zero the yaw rate, never clamp the angle) and
// boundary flag.  Executed by ONE lane for its env; returns bit0 = yaw violated, bit1 = boundary violated.
BP_DEV int shipice_substep_checks(const EnvView& ev, const LayoutView& L, const SimParams& P, int r) {
    int f = 0;
    real a = ev.body[r].pa;
    if (a <= P.yaw_lo || a >= P.yaw_hi) { ev.body[r].vw = RC(0.0); f |= 1; }
    V2 o = body_origin(ev, L, r);
    if (o.x < 0 || o.x > P.map_w) f |= 2;
    return f;
}

// total_work_done(prev_obs, updated) (evaluation/metrics.py:96-126):
// work = sum_i area(a_i) * || |c(a_i)| - |c(b_i)| || over the movable polygons in creation order; only
// bodies that moved this step can contribute.  Warp-cooperative; the sum is returned on lane 0.
BP_DEVNI double step_work(const EnvView& ev, const LayoutView& L) {
    EnvScal* sc = ev.sc;
    const int lane = bp_lane();
    int na = sc->n_awake;
    for (int k = lane; k < na; k += BP_NL) {
        int b = ev.awake[k];
        double w = 0.0;
        if (body_dynamic(L, b) && ev.moved_step[b]) {
            double A, cx, cy;
            area_centroid_abs(ev.wv + (size_t)b * VMAX * 2, L.nv[b], &A, &cx, &cy);
            double dx = ev.cenx[b] - cx, dy = ev.ceny[b] - cy;
            w = sqrt(dx * dx + dy * dy) * ev.area[b];
            ev.cenx[b] = cx; ev.ceny[b] = cy; ev.area[b] = A;
            ev.moved_step[b] = 0;
        }
        ev.tmp[b] = w;
    }
    bp_wsync();
    double work = RC(0.0);
    if (lane == 0) {
        // ascending body index, like zip(prev_obs, updated) in the reference
        const int r = sc->robot;
        for (int b = 0; b < r; b++) if (ev.is_awake[b] && body_dynamic(L, b)) work += ev.tmp[b];
    }
    return work;
}

// ship_ice_env.py:291-345: terminal tests, work, reward, info.
BP_DEVNI void shipice_step_finish(const EnvView& ev, const LayoutView& L, const SimParams& P, const Dims& d, int env, const StepOut& out,
                                  bool yaw_violated, bool boundary_violated) {
    EnvScal* sc = ev.sc;
    const int lane = bp_lane();
    const int r = sc->robot;
    V2 o = body_origin(ev, L, r);
    bool boundary_terminal = false;
    if (o.x < 0 && fabs(o.x - 0) >= P.boundary_limit) boundary_terminal = true;
    if (o.x > P.map_w && fabs(o.x - P.map_w) >= P.boundary_limit) boundary_terminal = true;
    double work = step_work(ev, L);
    if (lane == 0) {
        sc->total_work += work;
        double y = o.y;
        bool terminated = (y >= P.goal_y) || boundary_terminal;
        double dist_reward = RC(0.0);
        if (y < P.goal_y) {
            real a = ev.body[r].pa;
            real s, c; bp_sincos(a, &s, &c);
            dist_reward = RC(1.0) * (c * RC(0.0) + s * RC(1.0));
        }
        double collision_reward = -work;
        double reward = P.beta * collision_reward + dist_reward;
        if (yaw_violated) reward += P.yaw_penalty;
        if (boundary_violated) reward += P.boundary_penalty;
        bool success = false;
        if (terminated && !boundary_terminal) { reward += P.terminal_reward; success = true; }
        out.reward[env] = reward; out.terminated[env] = terminated ? 1 : 0; out.truncated[env] = 0;
        out.pose[3 * env] = o.x; out.pose[3 * env + 1] = o.y; out.pose[3 * env + 2] = ev.body[r].pa;
        out.state[3 * env] = round2(o.x); out.state[3 * env + 1] = round2(o.y); out.state[3 * env + 2] = round2(ev.body[r].pa);
        out.total_work[env] = sc->total_work; out.collision_reward[env] = collision_reward;
        out.scaled_collision_reward[env] = collision_reward * P.beta; out.dist_reward[env] = dist_reward;
        out.trial_success[env] = success ? 1 : 0; out.status[env] = sc->status;
    }
    bp_wsync();
}

}  // namespace bp
