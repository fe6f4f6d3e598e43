// TEST-ONLY host emulation of the box-delivery kernels: the same .cuh sources compiled as plain C++ with a
// "warp" / "block" of ONE lane (BP_HOST_EMU), behind the same entry-point names and signatures as the C ABI
// (include/benchpush_b200.h: bp_boxdel_*), with host memory where the library takes device pointers.
// Used by tests/ to check the kernel logic against the oracle on machines without a GPU.  The shipped
// library never contains or calls this build.
#define BP_HOST_EMU 1
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>
#include "../../include/benchpush_b200.h"
#include "../../benchpush_b200/csrc/host_tables.h"
#include "../../benchpush_b200/csrc/raster.cuh"
#include "../../benchpush_b200/csrc/maze_raster.cuh"
#include "../../benchpush_b200/csrc/shipice_env.cuh"
#include "../../benchpush_b200/csrc/spfa.cuh"
#include "../../benchpush_b200/csrc/boxdel_obs.cuh"
#include "../../benchpush_b200/csrc/boxdel_room.cuh"
#include "../../benchpush_b200/csrc/boxdel_host.h"

using namespace bp;
using namespace bp::bd;

struct bp_boxdel {
    bp_boxdel_config cfg;
    BdHost HB;
    std::vector<uint8_t> cspace, thin, small_free, rchan; std::vector<int32_t> closest; std::vector<float> recept;
    std::vector<char> envs; std::vector<int> scratch;
    int64_t launches = 0;
    bool loaded = false;
    bool device_maps = false;
};

static BdOut bd_out(const bp_boxdel_outputs* o) {
    BdOut r; memset(&r, 0, sizeof(r));
    if (o) { r.reward = o->reward; r.terminated = o->terminated; r.truncated = o->truncated; r.info = o->info; r.info_i = o->info_i; r.completed = o->completed; r.box_verts = o->box_verts; }
    return r;
}

extern "C" {

const char* bp_last_error(void) { return "emulation"; }

void bp_boxdel_default_config(bp_boxdel_config* c) {
    memset(c, 0, sizeof(*c));
    c->n_envs = 1; c->num_boxes = 10; c->device = 0; c->action_type = 1;
    c->steps = 100; c->iterations = 10; c->settle_steps = 1000; c->inactivity_cutoff = 200;
    c->ctrl_dt = 0.2; c->target_speed = 0.3; c->damping = 0.0;
    c->box_size = 0.44; c->box_density = 0.001;
    c->room_length = 10; c->room_width = 5; c->receptacle_width = 1.5;
    c->local_map_width = 10; c->local_map_pixel_width = 224; c->step_size = 1.75;
    c->robot_length = 0.6; c->robot_width = 0.92;
    c->partial_rewards_scale = 0.2; c->goal_reward = 10.0; c->collision_penalty = 0.25; c->non_movement_penalty = 0.0;
    c->correct_direction_reward_scale = 3.0; c->use_correct_direction_reward = 0; c->ministep_size = 2.5;
    c->shortest_path_channel_scale = 0.25; c->force_box_spfa = 0; c->phase_kernels = 0;
    c->env_kind = 0; c->step_limit = 10000; c->post_steps = 0; c->t_max = 0; c->n_boundary = 0; c->n_goals = 0;
    c->lfc = 0.0; c->omega_mul = 3.0; c->v_mul = 2.0; c->bound_x = c->room_length; c->bound_y = c->room_width;
    c->static_e = 0.01; c->static_u = 1.0; c->robot_code = 6; c->box_code = 4; c->box_done_code = 4; c->n_footprint = 0;
    c->exact_spfa_parents = 1;       // paths follow the reference solver's own parent tie-breaks; 0 = parallel relaxation + deterministic parents (faster)
    const double chassis[8] = {0.26, -0.35, -0.31, -0.35, -0.31, 0.35, 0.26, 0.35};
    const double bumper[8] = {0.31, -0.53, 0.20, -0.39, 0.20, 0.39, 0.31, 0.53};
    const double wheels[32] = {0.09, 0.35, 0.09, 0.39, 0.24, 0.39, 0.24, 0.35,   0.09, -0.39, 0.09, -0.35, 0.24, -0.35, 0.24, -0.39,
                               -0.09, 0.35, -0.09, 0.39, -0.24, 0.39, -0.24, 0.35,   -0.09, -0.35, -0.09, -0.39, -0.24, -0.39, -0.24, -0.35};
    memcpy(c->chassis_verts, chassis, sizeof(chassis)); memcpy(c->bumper_verts, bumper, sizeof(bumper)); memcpy(c->wheel_verts, wheels, sizeof(wheels));
}

int bp_boxdel_create(const bp_boxdel_config* cfg, bp_boxdel_t** out) {
    bp_boxdel* h = new bp_boxdel();
    h->cfg = *cfg;
    *out = h;
    return 0;
}
int bp_boxdel_destroy(bp_boxdel_t* h) { delete h; return 0; }

int bp_boxdel_load_rooms(bp_boxdel_t* h, int32_t n_layouts, const double* starts, const int32_t* n_boxes, const double* boxes,
                         int32_t n_rooms, const int32_t* layout_room,
                         int32_t n_static, const int32_t* st_nv, const double* st_verts, const double* st_pose, const double* st_rad, const int32_t* st_ctype,
                         int32_t H, int32_t W, const uint8_t* cspace, const uint8_t* cspace_thin, const int32_t* closest,
                         const uint8_t* small_free, const uint8_t* robot_chan, const float* ch3_map) {
    std::string err = bd_host_build(h->HB, h->cfg, n_layouts, starts, n_boxes, boxes, n_rooms, layout_room, n_static, st_nv, st_verts, st_pose, st_rad, st_ctype, H, W);
    if (!err.empty()) { fprintf(stderr, "emu_bd: %s\n", err.c_str()); return -1; }
    BdConst& K = h->HB.K;
    const BdParams& P = K.P;
    const size_t nv = (size_t)H * W, R = (size_t)n_rooms;
    h->cspace.assign(cspace, cspace + R * nv); h->thin.assign(cspace_thin, cspace_thin + R * nv); h->closest.assign(closest, closest + 2 * R * nv);
    h->small_free.assign(small_free, small_free + R * P.SMALL * P.SMALL); h->rchan.assign(robot_chan, robot_chan + (size_t)P.LW * P.LW);
    h->recept.assign(R * nv, 0.0f);
    K.layouts = h->HB.LY.data(); K.pair_a = h->HB.pa.data(); K.pair_b = h->HB.pb.data();
    K.rooms = h->HB.RM.data(); K.n_rooms = n_rooms;
    K.cspace = h->cspace.data(); K.cspace_thin = h->thin.data(); K.closest = h->closest.data(); K.small_free = h->small_free.data();
    K.robot_chan = h->rchan.data(); K.ch3_map = h->recept.data();
    h->envs.assign(K.env_bytes * h->cfg.n_envs, 0);
    K.envs = h->envs.data();
    h->scratch.assign(bd_scratch_ints(P), 0);
    K.P.border_blocked = getenv("BP_EMU_CHECKED_BORDER") ? 0 : bd_border_blocked(cspace, R, H, W);
    int pi = (int)std::floor(H / 2.0 - P.recept_y * P.ppm), pj = (int)std::floor(W / 2.0 + P.recept_x * P.ppm);
    pi = pi < 0 ? 0 : (pi > H - 1 ? H - 1 : pi); pj = pj < 0 ? 0 : (pj > W - 1 ? W - 1 : pj);
    int vis = 1;
    for (size_t r = 0; r < R; r++) {
        BdStatics& T = h->HB.RM[r];
        for (int k = 0; k < P.n_static; k++) bd_init_static(T, k);
        const int32_t* cl = closest + r * 2 * nv;
        T.recept_ci = cl[(size_t)pi * W + pj]; T.recept_cj = cl[nv + (size_t)pi * W + pj];
        if (ch3_map) std::copy(ch3_map + r * nv, ch3_map + (r + 1) * nv, h->recept.begin() + r * nv);
        else {
            std::vector<int32_t> par(nv);
            SpfaShared S;
            spfa_block(S, h->cspace.data() + r * nv, H, W, T.recept_ci, T.recept_cj, h->recept.data() + r * nv, par.data(), h->scratch.data(), h->scratch.data() + nv, h->scratch.data() + 2 * nv);
            bd_scale_receptacle_map(P, h->recept.data() + r * nv, nv);
        }
        if (h->cfg.env_kind == 0) for (size_t v = 0; v < nv && vis; v++) if (!bd_cell_visible(K, (int)r, (int)v)) vis = 0;
    }
    K.P.recept_visible = (vis && !h->cfg.force_box_spfa) ? 1 : 0;
    if (h->cfg.env_kind == 1) K.P.recept_visible = 1;      // no receptacle distances in area-clearing
    h->loaded = true;
    return 0;
}

int bp_boxdel_load_rooms_device(bp_boxdel_t* h, int32_t n_layouts, const double* starts, const int32_t* n_boxes, const double* boxes,
                                int32_t n_static, const int32_t* st_nv, const double* st_verts, const double* st_pose, const double* st_rad, const int32_t* st_ctype,
                                const int32_t* st_npx, const int32_t* st_pxy, int32_t H, int32_t W, const uint8_t* robot_chan) {
    std::vector<int32_t> ident(n_layouts);
    for (int i = 0; i < n_layouts; i++) ident[i] = i;
    std::string err = bd_host_build(h->HB, h->cfg, n_layouts, starts, n_boxes, boxes, n_layouts, ident.data(), n_static, st_nv, st_verts, st_pose, st_rad, st_ctype, H, W, st_npx, st_pxy);
    if (!err.empty()) { fprintf(stderr, "emu_bd: %s\n", err.c_str()); return -1; }
    BdConst& K = h->HB.K;
    const BdParams& P = K.P;
    const size_t nv = (size_t)H * W, R = (size_t)h->cfg.n_envs;
    h->cspace.assign(R * nv, 0); h->thin.assign(R * nv, 0); h->closest.assign(2 * R * nv, 0);
    h->small_free.assign(R * P.SMALL * P.SMALL, 0); h->rchan.assign(robot_chan, robot_chan + (size_t)P.LW * P.LW);
    h->recept.assign(R * nv, 0.0f);
    K.layouts = h->HB.LY.data(); K.pair_a = h->HB.pa.data(); K.pair_b = h->HB.pb.data();
    K.rooms = h->HB.RM.data(); K.n_rooms = n_layouts; K.maps_per_env = 1;
    K.cspace = h->cspace.data(); K.cspace_thin = h->thin.data(); K.closest = h->closest.data(); K.small_free = h->small_free.data();
    K.robot_chan = h->rchan.data(); K.ch3_map = h->recept.data();
    h->envs.assign(K.env_bytes * h->cfg.n_envs, 0);
    K.envs = h->envs.data();
    h->scratch.assign(bd_scratch_ints(P), 0);
    for (int r = 0; r < n_layouts; r++) for (int k = 0; k < P.n_static; k++) bd_init_static(h->HB.RM[r], k);
    K.P.recept_visible = 0; K.P.border_blocked = 1;
    h->device_maps = true;
    h->loaded = true;
    return 0;
}

int bp_boxdel_get_room_maps(bp_boxdel_t* h, int32_t env, uint8_t* cspace, uint8_t* cspace_thin, int32_t* closest, uint8_t* small_free, float* ch3) {
    const BdConst& K = h->HB.K;
    const BdParams& P = K.P;
    const size_t nv = (size_t)P.H * P.W, s = (size_t)bd_env(K, env)->map_slot, nn = (size_t)P.SMALL * P.SMALL;
    if (cspace) memcpy(cspace, K.cspace + s * nv, nv);
    if (cspace_thin) memcpy(cspace_thin, K.cspace_thin + s * nv, nv);
    if (closest) memcpy(closest, K.closest + s * 2 * nv, nv * 8);
    if (small_free) memcpy(small_free, K.small_free + s * nn, nn);
    if (ch3) memcpy(ch3, K.ch3_map + s * nv, nv * 4);
    return 0;
}

int bp_boxdel_load(bp_boxdel_t* h, int32_t n_layouts, const double* starts, const int32_t* n_boxes, const double* boxes,
                   int32_t n_static, const int32_t* st_nv, const double* st_verts, const double* st_pose, const double* st_rad, const int32_t* st_ctype,
                   int32_t H, int32_t W, const uint8_t* cspace, const uint8_t* cspace_thin, const int32_t* closest,
                   const uint8_t* small_free, const uint8_t* robot_chan, const float* ch3_map) {
    return bp_boxdel_load_rooms(h, n_layouts, starts, n_boxes, boxes, 1, nullptr, n_static, st_nv, st_verts, st_pose, st_rad, st_ctype,
                                H, W, cspace, cspace_thin, closest, small_free, robot_chan, ch3_map);
}

static void emu_obs(bp_boxdel* h, const uint8_t* mask, uint8_t* obs) {
    if (!obs) return;
    const BdConst& K = h->HB.K;
    BdObsShared S;
    size_t ob = (size_t)K.P.LW * K.P.LW * 4;
    for (int e = 0; e < h->cfg.n_envs; e++) if (!(mask && !mask[e])) bd_obs_item(S, K, bd_env(K, e), h->scratch.data(), obs + ob * e);
}

int bp_boxdel_reset(bp_boxdel_t* h, const int32_t* layout_ids, const uint8_t* mask, const bp_boxdel_outputs* out, void* stream) {
    const BdConst& K = h->HB.K;
    BdOut o = bd_out(out);
    for (int e = 0; e < h->cfg.n_envs; e++) {
        if (mask && !mask[e]) continue;
        BdEnv* E = bd_env(K, e);
        if (h->device_maps) {
            SpfaShared S; int mm[2];
            int layout = (layout_ids ? layout_ids[e] : e) % K.n_layouts;
            bd_build_room_maps(S, K, K.rooms[K.layouts[layout].room], e, h->scratch.data(), mm);
        }
        bd_reset_env(K, E, (layout_ids ? layout_ids[e] : e) % K.n_layouts, 0);
        bd_reset_out(K, E, e, o);
    }
    emu_obs(h, mask, out ? out->obs : nullptr);
    return 0;
}

int bp_boxdel_step(bp_boxdel_t* h, const float* actions, const uint8_t* mask, const bp_boxdel_outputs* out, void* stream) {
    const BdConst& K = h->HB.K;
    BdOut o = bd_out(out);
    SpfaShared S;
    static MoveVars V[BD_MOVE_K];
    const int n = h->cfg.n_envs;
    for (int e = 0; e < n; e++) {
        if (mask && !mask[e]) continue;
        BdEnv* E = bd_env(K, e);
        if (K.P.action_type == 2) bd_begin_step(K, E, (double)actions[2 * e], (double)actions[2 * e + 1]);
        else bd_begin_step(K, E, (double)actions[e]);
        for (int k = 0; k <= K.P.NB; k++) bd_dist_item(S, K, E, k, 0, h->scratch.data());
    }
    // as the kernel: one "warp" per BD_MOVE_K consecutive envs, phase-major
    for (int e0 = 0; e0 < n; e0 += BD_MOVE_K) bd_move(K, e0, (n - e0 < BD_MOVE_K) ? n - e0 : BD_MOVE_K, mask, V, nullptr, 0);
    for (int e = 0; e < n; e++) {
        if (mask && !mask[e]) continue;
        BdEnv* E = bd_env(K, e);
        for (int k = 0; k < K.P.NB; k++) bd_dist_item(S, K, E, k, 1, h->scratch.data());
        bd_finish_step(K, E, e, o);
    }
    emu_obs(h, mask, out ? out->obs : nullptr);
    return 0;
}

int bp_boxdel_step_host(bp_boxdel_t* h, const float* actions, const bp_boxdel_outputs* out_host) { return bp_boxdel_step(h, actions, nullptr, out_host, nullptr); }
int bp_boxdel_step_host_obs_device(bp_boxdel_t* h, const float* actions, const bp_boxdel_outputs* out_host, uint8_t* obs_dev) {
    bp_boxdel_outputs o = *out_host; o.obs = obs_dev;
    return bp_boxdel_step(h, actions, nullptr, &o, nullptr);
}

int bp_boxdel_get_bodies(bp_boxdel_t* h, int32_t env, double* out) {
    const BdEnv* E = bd_env(h->HB.K, env);
    for (int b = 0; b <= h->cfg.num_boxes; b++) {
        const BdBody& B = E->body[b];
        double ox = B.px - (B.cgx * B.rc - B.cgy * B.rs), oy = B.py - (B.cgx * B.rs + B.cgy * B.rc);
        double* o = out + 7 * b;
        o[0] = ox; o[1] = oy; o[2] = B.pa; o[3] = B.vx; o[4] = B.vy; o[5] = B.vw; o[6] = B.alive;
    }
    return 0;
}

int bp_boxdel_get_step_debug(bp_boxdel_t* h, int32_t env, double* out16, double* dists, double* path) {
    const BdEnv* E = bd_env(h->HB.K, env);
    if (out16) {
        out16[0] = E->n_path; out16[1] = E->move_sign; out16[2] = E->need_spfa; out16[3] = E->status; out16[4] = E->robot_hit;
        out16[5] = (double)E->stat_substeps; out16[6] = (double)E->stat_cand; out16[7] = (double)E->stat_arb; out16[8] = E->robot_distance;
        for (int i = 9; i < 16; i++) out16[i] = 0.0;
    }
    if (dists) for (int i = 0; i < h->cfg.num_boxes; i++) { dists[i] = E->cold->init_dist[i]; dists[h->cfg.num_boxes + i] = E->cold->final_dist[i]; }
    if (path) memcpy(path, E->cold->path, sizeof(E->cold->path));
    return 0;
}
// bd_feature_transform (csrc/boxdel_room.cuh) on an arbitrary map: free u8[H][W] -> indices i32[2][H][W]
void emu_feature_transform(const uint8_t* free_map, int H, int W, int32_t* out) {
    std::vector<int> g((size_t)H * W), r((size_t)H * W);
    int mm[2];
    bd_feature_transform(free_map, H, W, out, g.data(), r.data(), mm);
}
// host twin of bp_spfa_batch_host (same kernel source, one lane)
int bp_spfa_batch_host(const uint8_t* maps, const int32_t* sources, int32_t batch, int32_t H, int32_t W, float* dists, int32_t* parents) {
    size_t nv = (size_t)H * W;
    std::vector<int> scratch(3 * nv);
    SpfaShared S;
    for (int m = 0; m < batch; m++)
        spfa_block(S, maps + m * nv, H, W, sources[2 * m], sources[2 * m + 1], dists + m * nv, parents + m * nv, scratch.data(), scratch.data() + nv, scratch.data() + 2 * nv);
    return 0;
}
int64_t bp_boxdel_launch_count(bp_boxdel_t* h) { return h->HB.K.P.recept_visible; }   // emulation only: reports the visibility flag
size_t bp_boxdel_env_bytes(bp_boxdel_t* h) { return h->HB.K.env_bytes; }
int bp_boxdel_set_profiling(bp_boxdel_t*, int32_t) { return 0; }
int bp_boxdel_get_kernel_ms(bp_boxdel_t*, double* o) { o[0] = o[1] = o[2] = 0.0; return 0; }

}  // extern "C"
