// TEST-ONLY host emulation of the env kernels.
// Compiles benchpush_b200/csrc/*.cuh as plain C++ with a one-lane "warp" (BP_HOST_EMU) so kernel
// logic can be debugged and checked against the oracle on machines without a GPU.  It is never
// loaded by the benchpush_b200 package: the product path is the CUDA library only.
#define BP_HOST_EMU 1
#include <cstdio>
#include <cstdlib>
#include <string>
#include <vector>
#include "../../benchpush_b200/csrc/host_tables.h"
#include "../../benchpush_b200/csrc/sim_batch.cuh"
#define BP_RES_NS bk_default
#include "../../benchpush_b200/csrc/sim_res.cuh"
#include "../../benchpush_b200/csrc/raster.cuh"
#include "../../benchpush_b200/csrc/maze_raster.cuh"
#include "../../benchpush_b200/csrc/spfa.cuh"

using namespace bp;
using namespace bp::bk_default;      // sim_batch.cuh, single instantiation

struct Emu {
    Dims d; SimParams P; HostLayouts H; LayoutTab T;
    std::vector<char> mem; size_t env_bytes;
    std::vector<double> reward, state, pose, total_work, coll, scoll, distr;
    std::vector<unsigned char> term, trunc, succ; std::vector<int> status;
    StepOut out;
    ObsParams O; std::vector<unsigned char> raw_mode;
    // maze-NAMO
    MazeConfig mz; MazeObsParams MO; std::vector<uint8_t> wall_map; std::vector<double> dist_norm, dist_raw;
};

extern "C" {

void* emu_create(int n_envs, int max_floes, int pmax, int lmax, double goal_y, int settle_steps, double margin_body, double margin_robot) {
    Emu* e = new Emu();
    e->d.NB = max_floes + 1; e->d.PMAX = pmax; e->d.LMAX = lmax; e->d.n_envs = n_envs; e->d.n_layouts = 0;
    ShipIceConfig c; c.goal_y = goal_y; c.settle_steps = settle_steps; c.margin_body = margin_body; c.margin_robot = margin_robot;
    e->P = make_params(c);
    e->env_bytes = env_mem_size(e->d);
    e->mem.assign(e->env_bytes * n_envs, 0);
    e->reward.resize(n_envs); e->state.resize(3 * n_envs); e->pose.resize(3 * n_envs); e->total_work.resize(n_envs);
    e->coll.resize(n_envs); e->scoll.resize(n_envs); e->distr.resize(n_envs);
    e->term.resize(n_envs); e->trunc.resize(n_envs); e->succ.resize(n_envs); e->status.resize(n_envs);
    e->out.reward = e->reward.data(); e->out.terminated = e->term.data(); e->out.truncated = e->trunc.data();
    e->out.state = e->state.data(); e->out.pose = e->pose.data(); e->out.total_work = e->total_work.data();
    e->out.collision_reward = e->coll.data(); e->out.scaled_collision_reward = e->scoll.data(); e->out.dist_reward = e->distr.data();
    e->out.trial_success = e->succ.data(); e->out.status = e->status.data();
    ObsParams& O = e->O;
    O.grid_w = 300; O.grid_h = 1000; O.scale = 25.0; O.vshift = 2.0; O.goal_y = goal_y; O.g2m = 40.0 / 1000; O.range_x = 12.0; O.range_y = 12.0;
    O.head[0] = 1.0; O.head[1] = 0.0; O.tail[0] = -1.0; O.tail[1] = 0.0; O.n_robot_raw = 0;
    e->raw_mode.assign(n_envs, 1);
    return e;
}
void emu_destroy(void* h) { delete (Emu*)h; }
int emu_load_layouts(void* h, int n_layouts, const int32_t* n_floes, const int32_t* n_verts, const double* verts,
                     const double* centres, const double* ship_state, const double* ship_verts, int n_ship_verts) {
    Emu* e = (Emu*)h;
    ShipIceMaterial mat;
    std::string err = build_shipice_layouts(e->H, n_layouts, e->d.NB - 1, n_floes, n_verts, verts, centres, ship_state, ship_verts, n_ship_verts, mat);
    if (!err.empty()) { fprintf(stderr, "emu_load_layouts: %s\n", err.c_str()); return 1; }
    e->d.n_layouts = n_layouts;
    HostLayouts& H = e->H;
    e->T.lv = H.lv.data(); e->T.ln = H.ln.data(); e->T.nv = H.nv.data(); e->T.rad = H.rad.data(); e->T.rbound = H.rbound.data();
    e->T.minv = H.minv.data(); e->T.iinv = H.iinv.data(); e->T.cog = H.cog.data(); e->T.mat_e = H.mat_e.data();
    e->T.mat_u = H.mat_u.data(); e->T.init = H.init.data(); e->T.nbody = H.nbody.data(); e->T.robot0 = H.robot0.data();
    e->O.n_robot_raw = n_ship_verts;
    for (int i = 0; i < 2 * n_ship_verts; i++) e->O.robot_raw[i] = ship_verts[i];
    return 0;
}
void emu_reset(void* h, int env, int layout) {
    Emu* e = (Emu*)h;
    EnvView ev = make_view(e->mem.data() + e->env_bytes * env, e->d);
    shipice_reset_env(ev, e->T, e->P, e->d, layout, env, e->out);
}
void emu_step(void* h, int env, double action) {
    Emu* e = (Emu*)h;
    EnvView ev = make_view(e->mem.data() + e->env_bytes * env, e->d);
    shipice_step_env(ev, e->T, e->P, e->d, action, env, e->out);
}
// the batched path (sim_batch.cuh): envs [env0, env0 + BK) advance together; actions[n_envs]
void emu_step_batch(void* h, int env0, const float* actions, int auto_reset) {
    Emu* e = (Emu*)h;
    static DevConst C;
    C.d = e->d; C.P = e->P; C.off = env_offsets(e->d); C.T = e->T; C.snap = nullptr; C.env_bytes = e->env_bytes; C.snap_pose = nullptr; C.snap_status = nullptr;
    static BatchShared S;
    load_shared_offsets(C.off);
    shipice_step_batch(&C, S, e->mem.data(), e->env_bytes, env0, actions, nullptr, e->out, auto_reset);
}
// the resident path (sim_res.cuh): same contract, the awake world staged in a "shared memory" pool
void emu_step_resident(void* h, int env0, const float* actions, int auto_reset) {
    Emu* e = (Emu*)h;
    static DevConst C;
    C.d = e->d; C.P = e->P; C.off = env_offsets(e->d); C.T = e->T; C.snap = nullptr; C.env_bytes = e->env_bytes; C.snap_pose = nullptr; C.snap_status = nullptr;
    C.start_pose = nullptr;
    static res::ResShared S;
    load_shared_offsets(C.off);
    res::shipice_step_resident(&C, S, e->mem.data(), e->env_bytes, env0, actions, nullptr, e->out, auto_reset);
}
int emu_res_pool(int what) { return what == 0 ? res::R_PA : what == 1 ? res::R_PV : res::R_PS; }
void emu_set_bodies(void* h, int env, const double* bodies6) {
    Emu* e = (Emu*)h;
    EnvView ev = make_view(e->mem.data() + e->env_bytes * env, e->d);
    shipice_set_bodies_env(ev, e->T, e->P, e->d, bodies6);
}
int emu_batch_size() { return BK; }
// one raw physics substep (for per-substep parity); robot velocity set explicitly
void emu_set_robot_velocity(void* h, int env, double vx, double vy, double w) {
    Emu* e = (Emu*)h;
    EnvView ev = make_view(e->mem.data() + e->env_bytes * env, e->d);
    int r = ev.sc->robot; ev.body[r].vx = vx; ev.body[r].vy = vy; ev.body[r].vw = w;
}
void emu_substeps(void* h, int env, int n, int first) {
    Emu* e = (Emu*)h;
    EnvView ev = make_view(e->mem.data() + e->env_bytes * env, e->d);
    LayoutView L = layout_view(e->T, e->d, ev.sc->layout);
    for (int k = 0; k < n; k++) substep(ev, L, e->P, e->d, e->P.dt, (first && k == 0) ? 0.0 : e->P.dt);
}
void emu_init_only(void* h, int env, int layout) {
    Emu* e = (Emu*)h;
    EnvView ev = make_view(e->mem.data() + e->env_bytes * env, e->d);
    LayoutView L = layout_view(e->T, e->d, layout);
    init_env(ev, L, e->P, e->d, layout);
}
void emu_sleep_pass(void* h, int env) {
    Emu* e = (Emu*)h;
    EnvView ev = make_view(e->mem.data() + e->env_bytes * env, e->d);
    LayoutView L = layout_view(e->T, e->d, ev.sc->layout);
    sleep_pass(ev, L, e->d);
}
// [nbody][6]: origin x, y, angle, vx, vy, w
int emu_get_bodies(void* h, int env, double* out) {
    Emu* e = (Emu*)h;
    EnvView ev = make_view(e->mem.data() + e->env_bytes * env, e->d);
    LayoutView L = layout_view(e->T, e->d, ev.sc->layout);
    int nb = ev.sc->nbody;
    for (int b = 0; b < nb; b++) {
        V2 o = body_origin(ev, L, b);
        out[6 * b] = o.x; out[6 * b + 1] = o.y; out[6 * b + 2] = ev.body[b].pa; out[6 * b + 3] = ev.body[b].vx; out[6 * b + 4] = ev.body[b].vy; out[6 * b + 5] = ev.body[b].vw;
    }
    return nb;
}
void emu_get_out(void* h, int env, double* o13) {
    Emu* e = (Emu*)h;
    o13[0] = e->reward[env]; o13[1] = e->term[env]; o13[2] = e->trunc[env];
    o13[3] = e->state[3 * env]; o13[4] = e->state[3 * env + 1]; o13[5] = e->state[3 * env + 2];
    o13[6] = e->total_work[env]; o13[7] = e->coll[env]; o13[8] = e->scoll[env]; o13[9] = e->distr[env];
    o13[10] = e->succ[env]; o13[11] = e->status[env]; o13[12] = e->pose[3 * env + 2];
}
void emu_get_scal(void* h, int env, long long* o) {
    Emu* e = (Emu*)h;
    EnvView ev = make_view(e->mem.data() + e->env_bytes * env, e->d);
    EnvScal* s = ev.sc;
    o[0] = s->stamp; o[1] = s->n_awake; o[2] = s->n_pairs; o[3] = s->n_free; o[4] = s->n_order; o[5] = s->n_colours;
    o[6] = s->status; o[7] = s->n_apairs; o[8] = s->n_cached; o[9] = s->stat_gjk; o[10] = s->stat_pairs_eval;
    o[11] = s->stat_live_sum; o[12] = s->stat_awake_sum; o[13] = s->stat_recolour; o[14] = s->stat_rebuild; o[15] = s->stat_substeps;
    o[16] = s->stat_epa; o[17] = s->stat_skipped;
}
// hull vertices of body b of layout l (for the hull parity test)
int emu_layout_verts(void* h, int l, int b, double* out, double* minv_iinv) {
    Emu* e = (Emu*)h;
    size_t idx = (size_t)l * e->d.NB + b;
    int n = e->H.nv[idx];
    for (int i = 0; i < 2 * n; i++) out[i] = e->H.lv[idx * VMAX * 2 + i];
    minv_iinv[0] = e->H.minv[idx]; minv_iinv[1] = e->H.iinv[idx]; minv_iinv[2] = e->H.cog[2 * idx]; minv_iinv[3] = e->H.cog[2 * idx + 1];
    return n;
}
// observation of env: raw != 0 rasterises the raw trial vertices (reset), else the hull world vertices
void emu_render_obs(void* h, int env, int raw, unsigned char* out) {
    Emu* e = (Emu*)h;
    EnvView ev = make_view(e->mem.data() + e->env_bytes * env, e->d);
    LayoutView L = layout_view(e->T, e->d, ev.sc->layout);
    static ObsSmem S;
    int r = ev.sc->robot;
    V2 o = body_origin(ev, L, r);
    if (raw) {
        size_t base = (size_t)ev.sc->layout * e->d.NB;
        render_obs(S, e->O, r, e->H.rawv.data() + base * VMAX * 2, VMAX * 2, e->H.rawn.data() + base, e->H.rawcen.data() + base * 2,
                   e->H.rawcen.data() + base * 2 + 1, 2, o.x, o.y, ev.body[r].pa, out);
    } else {
        render_obs(S, e->O, r, ev.wv, VMAX * 2, L.nv, ev.cenx, ev.ceny, 1, o.x, o.y, ev.body[r].pa, out);
    }
}
// cfg.egocentric_obs = false: render_obs_global on the same sources as k_shipice_obs_global; out [2][200][60]
void emu_render_obs_global(void* h, int env, int raw, unsigned char* out) {
    Emu* e = (Emu*)h;
    EnvView ev = make_view(e->mem.data() + e->env_bytes * env, e->d);
    LayoutView L = layout_view(e->T, e->d, ev.sc->layout);
    static ObsGlobalSmem S;
    GlobalObsParams G;       // the default map: 40 x 12 m at 25 px/m, 0.2 m cells (bp_api.cu: global_obs_params)
    G.raw_h = e->O.grid_h; G.raw_w = e->O.grid_w; G.blk_y = G.blk_x = 5; G.gh = 200; G.gw = 60; G.cell_scale = 5.0;
    int r = ev.sc->robot;
    V2 o = body_origin(ev, L, r);
    if (raw) {
        size_t base = (size_t)ev.sc->layout * e->d.NB;
        render_obs_global(S, e->O, G, r, e->H.rawv.data() + base * VMAX * 2, VMAX * 2, e->H.rawn.data() + base, o.x, o.y, ev.body[r].pa, out);
    } else {
        render_obs_global(S, e->O, G, r, ev.wv, VMAX * 2, L.nv, o.x, o.y, ev.body[r].pa, out);
    }
}
// the |centroid| cache k_low_dim compacts (cenx / ceny of the bodies below the robot); returns the count
int emu_centroid_cache(void* h, int env, double* out) {
    Emu* e = (Emu*)h;
    EnvView ev = make_view(e->mem.data() + e->env_bytes * env, e->d);
    LayoutView L = layout_view(e->T, e->d, ev.sc->layout);
    int n = 0;
    for (int b = 0; b < ev.sc->robot; b++) if (body_dynamic(L, b)) { out[2 * n] = ev.cenx[b]; out[2 * n + 1] = ev.ceny[b]; n++; }
    return n;
}
// spfa_serial_parents (the reference solver's queue restated), reference output convention
void emu_spfa_serial(const uint8_t* map, int H, int W, int si, int sj, float* dist, int32_t* parents) {
    size_t nv = (size_t)H * W;
    std::vector<int> ring(nv + 2); std::vector<unsigned char> inq(nv);
    spfa_serial_parents(map, H, W, si, sj, dist, parents, ring.data(), inq.data());
    const float inf = 2.0f * (float)nv;
    for (size_t v = 0; v < nv; v++) if (!(dist[v] < inf)) dist[v] = 0.0f;
}
// world-space hull vertices of body b as the kernels hold them (ev.wv), widened to double; returns the vertex count
int emu_world_verts(void* h, int env, int b, double* out) {
    Emu* e = (Emu*)h;
    EnvView ev = make_view(e->mem.data() + e->env_bytes * env, e->d);
    LayoutView L = layout_view(e->T, e->d, ev.sc->layout);
    int n = L.nv[b];
    for (int i = 0; i < 2 * n; i++) out[i] = ev.wv[(size_t)b * VMAX * 2 + i];
    return n;
}
void emu_spfa(const uint8_t* map, int H, int W, int si, int sj, float* dist, int32_t* parents) {
    std::vector<int> scratch((size_t)3 * H * W);
    SpfaShared S;
    spfa_block(S, map, H, W, si, sj, dist, parents, scratch.data(), scratch.data() + (size_t)H * W, scratch.data() + (size_t)2 * H * W);
}
size_t emu_env_bytes(void* h) { return ((Emu*)h)->env_bytes; }

// ---------------------------------------------------------------- maze-NAMO (default config of maze_version v)
void* emu_maze_create(int n_envs, int max_boxes, int maze_version, int settle_steps) {
    Emu* e = new Emu();
    MazeConfig& m = e->mz;
    double W = maze_version == 2 ? 20.0 : 15.0, Lg = W;
    m.width = W; m.length = Lg; m.settle_steps = settle_steps;
    if (maze_version == 2) {
        m.goal_x = m.goal_y = 3.33; m.start[0] = m.start[1] = 16.66; m.start[2] = 3 * M_PI / 2;
        const double w[6][4] = {{0, 0, W, 0}, {0, 0, 0, Lg}, {W, 0, W, Lg}, {0, Lg, W, Lg}, {W / 3, 0, W / 3, 2 * Lg / 3}, {2 * W / 3, Lg, 2 * W / 3, Lg / 3}};
        memcpy(m.walls, w, sizeof(w));
    } else {
        m.goal_x = m.goal_y = 3.75; m.start[0] = 11.25; m.start[1] = 3.75; m.start[2] = M_PI / 2;
        const double w[6][4] = {{0, 0, W, 0}, {0, 0, 0, Lg}, {W, 0, W, Lg}, {0, Lg, W, Lg}, {2 * W / 2, Lg, 2 * W / 2, 5}, {W / 2, 0, W / 2, Lg - Lg / 3}};
        memcpy(m.walls, w, sizeof(w));
    }
    m.n_walls = 6;
    static const double rv[] = {0.7, -0.5, 0.55, -0.6, -0.55, -0.6, -0.7, -0.5, -0.7, 0.5, -0.55, 0.6, 0.55, 0.6, 0.7, 0.5};
    m.n_robot_verts = 8; memcpy(m.robot_verts, rv, sizeof(rv));
    static const double wv[4][8] = {{0.2, 0.6, 0.2, 0.7, 0.5, 0.7, 0.55, 0.6}, {0.2, -0.7, 0.2, -0.6, 0.55, -0.6, 0.55, -0.7},
                                    {-0.2, 0.6, -0.2, 0.7, -0.55, 0.7, -0.55, 0.6}, {-0.2, -0.6, -0.2, -0.7, -0.55, -0.7, -0.55, -0.6}};
    m.n_wheels = 4;
    for (int w = 0; w < 4; w++) { m.wheel_nv[w] = 4; memcpy(m.wheel_verts + 16 * w, wv[w], sizeof(wv[w])); }
    e->d.NB = m.n_walls + max_boxes + 1 + m.n_wheels;
    e->d.PMAX = e->d.NB * (e->d.NB - 1) / 2 + 8; e->d.LMAX = e->d.PMAX < 256 ? e->d.PMAX : 256; if (e->d.LMAX < 64) e->d.LMAX = 64;
    e->d.n_envs = n_envs; e->d.n_layouts = 0;
    e->P = make_maze_params(m);
    int H = 0, Wd = 0;
    build_maze_maps(m, &H, &Wd, e->wall_map, e->dist_norm, e->dist_raw);
    e->P.dist_map = e->dist_norm.data(); e->P.dist_h = H; e->P.dist_w = Wd;
    e->MO.grid_h = H; e->MO.grid_w = Wd; e->MO.scale = (double)m.m_to_pix_scale; e->MO.n_robot_raw = m.n_robot_verts;
    memcpy(e->MO.robot_raw, m.robot_verts, sizeof(double) * 2 * m.n_robot_verts);
    e->MO.wall_map = e->wall_map.data(); e->MO.dist_map = e->dist_norm.data();
    e->env_bytes = env_mem_size(e->d);
    e->mem.assign(e->env_bytes * n_envs, 0);
    e->reward.resize(n_envs); e->state.resize(3 * n_envs); e->pose.resize(3 * n_envs); e->total_work.resize(n_envs);
    e->coll.resize(n_envs); e->scoll.resize(n_envs); e->distr.resize(n_envs);
    e->term.resize(n_envs); e->trunc.resize(n_envs); e->succ.resize(n_envs); e->status.resize(n_envs);
    e->out.reward = e->reward.data(); e->out.terminated = e->term.data(); e->out.truncated = e->trunc.data();
    e->out.state = e->state.data(); e->out.pose = e->pose.data(); e->out.total_work = e->total_work.data();
    e->out.collision_reward = e->coll.data(); e->out.scaled_collision_reward = e->scoll.data(); e->out.dist_reward = e->distr.data();
    e->out.trial_success = e->succ.data(); e->out.status = e->status.data();
    return e;
}
int emu_maze_load_layouts(void* h, int n_layouts, int max_boxes, const int32_t* n_boxes, const double* centres) {
    Emu* e = (Emu*)h;
    std::string err = build_maze_layouts(e->H, n_layouts, max_boxes, n_boxes, centres, e->mz);
    if (!err.empty()) { fprintf(stderr, "emu_maze_load_layouts: %s\n", err.c_str()); return 1; }
    e->d.n_layouts = n_layouts;
    HostLayouts& H = e->H;
    e->T.lv = H.lv.data(); e->T.ln = H.ln.data(); e->T.nv = H.nv.data(); e->T.rad = H.rad.data(); e->T.rbound = H.rbound.data();
    e->T.minv = H.minv.data(); e->T.iinv = H.iinv.data(); e->T.cog = H.cog.data(); e->T.mat_e = H.mat_e.data();
    e->T.mat_u = H.mat_u.data(); e->T.init = H.init.data(); e->T.nbody = H.nbody.data(); e->T.robot0 = H.robot0.data();
    return 0;
}
void emu_maze_get_maps(void* h, uint8_t* wall, double* dist_norm, double* dist_raw) {
    Emu* e = (Emu*)h;
    memcpy(wall, e->wall_map.data(), e->wall_map.size());
    memcpy(dist_norm, e->dist_norm.data(), e->dist_norm.size() * 8);
    memcpy(dist_raw, e->dist_raw.data(), e->dist_raw.size() * 8);
}
void emu_maze_render_obs(void* h, int env, int raw, unsigned char* out) {
    Emu* e = (Emu*)h;
    EnvView ev = make_view(e->mem.data() + e->env_bytes * env, e->d);
    LayoutView L = layout_view(e->T, e->d, ev.sc->layout);
    static MazeObsSmem S;
    int r = ev.sc->robot;
    V2 o = body_origin(ev, L, r);
    int first = 0;
    while (first < L.robot0 && L.nv[first] == 2) first++;
    int n_obst = L.robot0 - first;
    if (raw) {
        size_t base = (size_t)ev.sc->layout * e->d.NB + first;
        render_maze_obs(S, e->MO, n_obst, e->H.rawv.data() + base * VMAX * 2, VMAX * 2, e->H.rawn.data() + base, o.x, o.y, ev.body[r].pa, out);
    } else {
        render_maze_obs(S, e->MO, n_obst, ev.wv + (size_t)first * VMAX * 2, VMAX * 2, L.nv + first, o.x, o.y, ev.body[r].pa, out);
    }
}
int emu_get_scal2(void* h, int env, int* o) {
    Emu* e = (Emu*)h;
    EnvView ev = make_view(e->mem.data() + e->env_bytes * env, e->d);
    o[0] = ev.sc->robot; o[1] = ev.sc->nbody; o[2] = ev.sc->n_fixed; o[3] = ev.sc->wall_hit; o[4] = ev.sc->has_prev_dist;
    return 0;
}

}  // extern "C"
