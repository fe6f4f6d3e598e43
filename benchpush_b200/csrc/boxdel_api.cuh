// box-delivery-v0: kernels and C ABI (bp_boxdel_*, declared in include/benchpush_b200.h).
// Included at the end of bp_api.cu (one translation unit, one shared library).
#pragma once
#include "boxdel_obs.cuh"
#include "boxdel_room.cuh"
#include "boxdel_host.h"

using namespace bp::bd;

#ifndef BD_MOVE_MINB
#define BD_MOVE_MINB 4          // resident 128-thread blocks per SM the warp-per-env kernels are compiled for (<= 128 registers)
#endif

// ------------------------------------------------------------------------------------------------ kernels
__global__ void k_bd_init_statics(BdStatics* rooms, int n_static) {
    int k = threadIdx.x;
    if (k < n_static) bd_init_static(rooms[blockIdx.x], k);
}

// thread per cell: clears *flag when a free cell is not visible from the receptacle cell
__global__ void k_bd_visibility(const BdConst* C, int* flag) {
    int v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v < C->P.H * C->P.W && !bd_cell_visible(*C, blockIdx.y, v)) atomicAnd(flag, 0);
}

// block per env pulled from a counter: the env's room maps from its layout's static shapes (boxdel_room.cuh)
__global__ void __launch_bounds__(256) k_bd_build_rooms(const BdConst* C, const int32_t* layout_ids, const uint8_t* mask, int* scratch, int* counter) {
    __shared__ SpfaShared S;
    __shared__ int s_env;
    __shared__ int s_mm[2];
    int* my = scratch + (size_t)blockIdx.x * bd_scratch_ints(C->P);
    for (;;) {
        if (threadIdx.x == 0) {
            int e;
            for (;;) { e = atomicAdd(counter, 1); if (e >= C->n_envs || !(mask && !mask[e])) break; }
            s_env = e;
        }
        __syncthreads();
        int env = s_env;
        __syncthreads();
        if (env >= C->n_envs) break;
        int layout = (layout_ids ? layout_ids[env] : env) % C->n_layouts;
        bd_build_room_maps(S, *C, C->rooms[C->layouts[layout].room], env, my, s_mm);
    }
}

// warp per env
__global__ void __launch_bounds__(128, BD_MOVE_MINB) k_bd_reset(const BdConst* C, const int32_t* layout_ids, const uint8_t* mask, BdOut out) {
    int env = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (env >= C->n_envs || (mask && !mask[env])) return;
    BdEnv* E = bd_env(*C, env);
    int layout = (layout_ids ? layout_ids[env] : env) % C->n_layouts;
    int episode = 0;
    bd_reset_env(*C, E, layout, episode);
    if ((threadIdx.x & 31) == 0) bd_reset_out(*C, E, env, out);
}

// thread per env: action decode, target, straight-line test
__global__ void k_bd_begin(const BdConst* C, const float* actions, const uint8_t* mask) {
    int env = blockIdx.x * blockDim.x + threadIdx.x;
    if (env >= C->n_envs || (mask && !mask[env])) return;
    if (C->P.action_type == 2) bd_begin_step(*C, bd_env(*C, env), (double)actions[2 * env], (double)actions[2 * env + 1]);
    else bd_begin_step(*C, bd_env(*C, env), (double)actions[env]);
}

// block per (env, item) pulled from a counter; item k < NB: box distance, k == NB: robot path (phase 0)
__global__ void __launch_bounds__(256) k_bd_dist(const BdConst* C, const uint8_t* mask, int phase, int* scratch, int* counter) {
    __shared__ SpfaShared S;
    __shared__ int s_item;
    const int per_env = C->P.NB + 1;
    const int total = C->n_envs * per_env;
    int* my = scratch + (size_t)blockIdx.x * bd_scratch_ints(C->P);
    for (;;) {
        if (threadIdx.x == 0) {
            // skip items that need no work without a block-wide round trip
            int it;
            for (;;) {
                it = atomicAdd(counter, 1);
                if (it >= total) break;
                int env = it / per_env, k = it - env * per_env;
                if (mask && !mask[env]) continue;
                BdEnv* E = bd_env(*C, env);
                if (k < C->P.NB) { if (E->body[1 + k].alive == 1 && !C->P.recept_visible && !bd_box_dist_cached(E, k, phase, true)) break; }
                else if (phase == 0 && E->need_spfa) break;
            }
            s_item = it;
        }
        __syncthreads();
        int it = s_item;
        __syncthreads();
        if (it >= total) break;
        int env = it / per_env, k = it - env * per_env;
        bd_dist_item(S, *C, bd_env(*C, env), k, phase, my);
    }
}

// one warp per BD_MOVE_K envs: waypoint headings, then the phase-major substep state machine (block-synchronous: every warp
// of the block takes part)
#ifndef BD_MOVE_BLOCK
#define BD_MOVE_BLOCK 128
#endif
__global__ void __launch_bounds__(BD_MOVE_BLOCK, BD_MOVE_MINB * 128 / BD_MOVE_BLOCK) k_bd_move(const BdConst* C, const uint8_t* mask) {
    __shared__ MoveVars V[BD_MOVE_BLOCK / 32][BD_MOVE_K];
    // hot env headers of this block's envs stay in shared memory for the whole launch (6.6 KB each): the loop touches them
    // ~100 times per substep, and with every resident warp streaming its own header L1 did not hold them (occupancy sweep)
    __shared__ __align__(16) char hot[BD_MOVE_BLOCK / 32][BD_MOVE_K][(sizeof(BdEnv) + 15) / 16 * 16];
    int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    int env0 = warp * BD_MOVE_K;
    int nk = C->n_envs - env0;
    nk = nk < 0 ? 0 : (nk > BD_MOVE_K ? BD_MOVE_K : nk);
    bd_move(*C, env0, nk, mask, V[threadIdx.x >> 5], &hot[threadIdx.x >> 5][0][0], (int)sizeof(hot[0][0]));
}

// ---- phase-per-kernel variant of the substep loop: every warp of the GPU runs the same small piece of code at the same time
// (one kernel per phase over the whole batch, warp per env, loop state in global memory), so the instruction working set of a
// launch fits the instruction cache.  The host enqueues chunks of iterations and reads one flag per chunk.
__global__ void __launch_bounds__(128) k_mv_init(const BdConst* C, const uint8_t* mask, MoveVars* mv) {
    int env = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (env >= C->n_envs || (threadIdx.x & 31)) return;
    MoveVars& v = mv[env];
    bool part = !(mask && !mask[env]);
    v.state = part ? (C->P.action_type == 2 ? (C->P.kind == 1 ? (C->P.post_steps > 0 ? MV_POST : MV_DONE) : MV_VEL) : MV_EXEC) : MV_DONE;
    if (!part) return;
    BdEnv* E = bd_env(*C, env);
    bd_finalize_path(*C, E);
    v.wi = 1; v.sim_steps = 0; v.post_k = 0; v.n = E->n_path; v.done_turning = 0; v.have_prev = 0; v.still_done = 0;
    v.pos_x = E->init_pos[0]; v.pos_y = E->init_pos[1]; v.heading = E->init_heading;
    v.robot_distance = 0.0; v.prev_heading_diff = 0.0; v.hd = 0.0; v.LS.dp_none = 1;
}
__global__ void __launch_bounds__(128) k_mv_pre_a(const BdConst* C, MoveVars* mv) {
    int env = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (env >= C->n_envs || mv[env].state == MV_DONE) return;
    BdEnv* E = bd_env(*C, env);
    mv_pre(*C, E, mv[env]);
    bd_phase_a(*C, E);
}
__global__ void __launch_bounds__(128) k_mv_b(const BdConst* C, MoveVars* mv) {
    int env = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (env >= C->n_envs || mv[env].state == MV_DONE) return;
    BdEnv* E = bd_env(*C, env);
    if (!E->fast) bd_phase_b(*C, E);
}
__global__ void __launch_bounds__(128) k_mv_c(const BdConst* C, MoveVars* mv) {
    int env = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (env >= C->n_envs || mv[env].state == MV_DONE) return;
    BdEnv* E = bd_env(*C, env);
    if (!E->fast) bd_phase_c(*C, E);
}
__global__ void __launch_bounds__(128) k_mv_d_post(const BdConst* C, MoveVars* mv) {
    int env = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (env >= C->n_envs || mv[env].state == MV_DONE) return;
    BdEnv* E = bd_env(*C, env);
    if (!E->fast) bd_phase_d(*C, E);
    mv_post(*C, E, mv[env]);
}
__global__ void k_mv_alive(const BdConst* C, const MoveVars* mv, int* flag) {
    int env = blockIdx.x * blockDim.x + threadIdx.x;
    if (env < C->n_envs && mv[env].state != MV_DONE) *flag = 1;
}

// thread per env
__global__ void k_bd_finish(const BdConst* C, const uint8_t* mask, BdOut out) {
    int env = blockIdx.x * blockDim.x + threadIdx.x;
    if (env >= C->n_envs || (mask && !mask[env])) return;
    bd_finish_step(*C, bd_env(*C, env), env, out);
}

// block per env pulled from a counter
__global__ void __launch_bounds__(256) k_bd_obs(const BdConst* C, const uint8_t* mask, uint8_t* obs, int* scratch, int* counter) {
    __shared__ BdObsShared S;
    __shared__ int s_env;
    int* my = scratch + (size_t)blockIdx.x * bd_scratch_ints(C->P);
    const size_t obs_bytes = (size_t)C->P.LW * C->P.LW * 4;
    for (;;) {
        if (threadIdx.x == 0) {
            int e;
            for (;;) { e = atomicAdd(counter, 1); if (e >= C->n_envs || !(mask && !mask[e])) break; }
            s_env = e;
        }
        __syncthreads();
        int env = s_env;
        __syncthreads();
        if (env >= C->n_envs) break;
        bd_obs_item(S, *C, bd_env(*C, env), my, obs + obs_bytes * env);
    }
}

// ------------------------------------------------------------------------------------------------ handle
struct bp_boxdel {
    bp_boxdel_config cfg;
    BdConst hC;                      // host copy (device pointers inside)
    BdConst* dC = nullptr;
    BdLayout* d_layouts = nullptr; BdStatics* d_rooms = nullptr; short* d_pa = nullptr; short* d_pb = nullptr;
    uint8_t *d_cspace = nullptr, *d_thin = nullptr, *d_small = nullptr, *d_rchan = nullptr; int32_t* d_closest = nullptr; float* d_recept = nullptr;
    char* d_envs = nullptr; int* d_scratch = nullptr; int* d_counter = nullptr; int n_blocks = 0;
    int dist_blocks = 0, dist_threads = 256;      // k_bd_dist launch shape (reference-order parents: many small blocks, one serial solve each)
    bool loaded = false;
    bool device_maps = false;         // rooms built on the device at reset (bp_boxdel_load_rooms_device)
    int64_t launches = 0;
    // optional kernel timing (CUDA events on the launching stream around k_bd_move / k_bd_obs of every step)
    bool profiling = false;
    std::vector<cudaEvent_t> ev;      // 4 per timed step: move start / stop, obs start / stop
    MoveVars* d_mv = nullptr; int* h_flag = nullptr; int phase_chunk = 64;
};


extern "C" {

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
    if (!cfg || !out) return fail(BP_ERR_INVALID, "null argument");
    if (cfg->n_envs <= 0 || cfg->num_boxes <= 0 || cfg->num_boxes > MAXBOX) return fail(BP_ERR_INVALID, "n_envs / num_boxes out of range");
    if (cfg->action_type < 0 || cfg->action_type > 2) return fail(BP_ERR_INVALID, "action_type must be 0 (position), 1 (heading) or 2 (velocity: two floats per env)");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0) return fail(BP_ERR_CUDA, "no CUDA device (there is no CPU fallback)");
    CUDA_TRY(cudaSetDevice(cfg->device));
    bp_boxdel* h = new bp_boxdel();
    h->cfg = *cfg;
    memset(&h->hC, 0, sizeof(h->hC));
    *out = h;
    return BP_OK;
}

int bp_boxdel_destroy(bp_boxdel_t* h) {
    if (!h) return BP_OK;
    cudaSetDevice(h->cfg.device);
    cudaFree(h->dC); cudaFree(h->d_layouts); cudaFree(h->d_rooms); cudaFree(h->d_pa); cudaFree(h->d_pb); cudaFree(h->d_cspace); cudaFree(h->d_thin); cudaFree(h->d_small);
    cudaFree(h->d_rchan); cudaFree(h->d_closest); cudaFree(h->d_recept); cudaFree(h->d_envs); cudaFree(h->d_scratch); cudaFree(h->d_counter); cudaFree(h->d_mv); if (h->h_flag) cudaFreeHost(h->h_flag);
    delete h;
    return BP_OK;
}

// All pointers are HOST pointers.  starts[L][3]; n_boxes[L]; boxes[L][num_boxes][3] (x, y, heading);
// statics: st_nv[ns], st_verts[ns][4][2] (body-local, as passed to pymunk.Poly), st_pose[ns][3], st_rad[ns], st_ctype[ns]
// in shape creation order (receptacle, walls, corners: box_delivery_env.py:256-263, sim_utils.py:96-131);
// maps of the room (box_delivery_env.py:1140-1175): cspace / cspace_thin u8[H][W] (1 = free), closest i32[2][H][W],
// small_free u8[S][S] (S = local_map_pixel_width + 20), robot_chan u8[LW][LW] (0 / 255).
int bp_boxdel_load(bp_boxdel_t* h, int32_t n_layouts, const double* starts, const int32_t* n_boxes, const double* boxes,
                   int32_t n_static, const int32_t* st_nv, const double* st_verts, const double* st_pose, const double* st_rad, const int32_t* st_ctype,
                   int32_t H, int32_t W, const uint8_t* cspace, const uint8_t* cspace_thin, const int32_t* closest,
                   const uint8_t* small_free, const uint8_t* robot_chan, const float* ch3_map) {
    return bp_boxdel_load_rooms(h, n_layouts, starts, n_boxes, boxes, 1, nullptr, n_static, st_nv, st_verts, st_pose, st_rad, st_ctype,
                                H, W, cspace, cspace_thin, closest, small_free, robot_chan, ch3_map);
}

static int bd_load_impl(bp_boxdel_t* h, int32_t n_layouts, const double* starts, const int32_t* n_boxes, const double* boxes,
                        int32_t n_rooms, const int32_t* layout_room,
                        int32_t n_static, const int32_t* st_nv, const double* st_verts, const double* st_pose, const double* st_rad, const int32_t* st_ctype,
                        int32_t H, int32_t W, const uint8_t* cspace, const uint8_t* cspace_thin, const int32_t* closest,
                        const uint8_t* small_free, const uint8_t* robot_chan, const float* ch3_map,
                        const int32_t* st_npx, const int32_t* st_pxy, bool device_maps);

// As bp_boxdel_load with a leading room axis on every room-dependent array (include/benchpush_b200.h).
int bp_boxdel_load_rooms(bp_boxdel_t* h, int32_t n_layouts, const double* starts, const int32_t* n_boxes, const double* boxes,
                         int32_t n_rooms, const int32_t* layout_room,
                         int32_t n_static, const int32_t* st_nv, const double* st_verts, const double* st_pose, const double* st_rad, const int32_t* st_ctype,
                         int32_t H, int32_t W, const uint8_t* cspace, const uint8_t* cspace_thin, const int32_t* closest,
                         const uint8_t* small_free, const uint8_t* robot_chan, const float* ch3_map) {
    if (!cspace || !cspace_thin || !closest || !small_free) return fail(BP_ERR_INVALID, "null argument");
    return bd_load_impl(h, n_layouts, starts, n_boxes, boxes, n_rooms, layout_room, n_static, st_nv, st_verts, st_pose, st_rad, st_ctype,
                        H, W, cspace, cspace_thin, closest, small_free, robot_chan, ch3_map, nullptr, nullptr, false);
}

// One room per layout, static shapes only; the maps of an env are built on the device whenever the env is reset
// (k_bd_build_rooms, boxdel_room.cuh) into the env's own map slot.  st_npx[n_layouts][n_static], st_pxy[n_layouts][n_static][4][2]:
// the pixel polygons of the collision_type-3 shapes as the reference passes them to cv2.fillPoly.
int bp_boxdel_load_rooms_device(bp_boxdel_t* h, int32_t n_layouts, const double* starts, const int32_t* n_boxes, const double* boxes,
                                int32_t n_static, const int32_t* st_nv, const double* st_verts, const double* st_pose, const double* st_rad, const int32_t* st_ctype,
                                const int32_t* st_npx, const int32_t* st_pxy, int32_t H, int32_t W, const uint8_t* robot_chan) {
    if (!st_npx || !st_pxy) return fail(BP_ERR_INVALID, "null argument");
    if (h && h->cfg.env_kind != 0) return fail(BP_ERR_INVALID, "device-built rooms are a box-delivery feature");
    std::vector<int32_t> ident(n_layouts > 0 ? n_layouts : 0);
    for (int i = 0; i < n_layouts; i++) ident[i] = i;
    return bd_load_impl(h, n_layouts, starts, n_boxes, boxes, n_layouts, ident.data(), n_static, st_nv, st_verts, st_pose, st_rad, st_ctype,
                        H, W, nullptr, nullptr, nullptr, nullptr, robot_chan, nullptr, st_npx, st_pxy, true);
}

static int bd_load_impl(bp_boxdel_t* h, int32_t n_layouts, const double* starts, const int32_t* n_boxes, const double* boxes,
                        int32_t n_rooms, const int32_t* layout_room,
                        int32_t n_static, const int32_t* st_nv, const double* st_verts, const double* st_pose, const double* st_rad, const int32_t* st_ctype,
                        int32_t H, int32_t W, const uint8_t* cspace, const uint8_t* cspace_thin, const int32_t* closest,
                        const uint8_t* small_free, const uint8_t* robot_chan, const float* ch3_map,
                        const int32_t* st_npx, const int32_t* st_pxy, bool device_maps) {
    if (!h || !starts || !n_boxes || !boxes || !st_nv || !st_verts || !st_pose || !st_rad || !st_ctype || !robot_chan)
        return fail(BP_ERR_INVALID, "null argument");
    if (h->loaded) return fail(BP_ERR_STATE, "already loaded");
    if (n_layouts <= 0 || n_static <= 0 || n_static > MAXSTAT || n_rooms <= 0) return fail(BP_ERR_INVALID, "n_layouts / n_static / n_rooms out of range");
    const bp_boxdel_config& c = h->cfg;
    CUDA_TRY(cudaSetDevice(c.device));
    BdHost HB;
    {
        std::string err = bd_host_build(HB, c, n_layouts, starts, n_boxes, boxes, n_rooms, layout_room, n_static, st_nv, st_verts, st_pose, st_rad, st_ctype, H, W, st_npx, st_pxy);
        if (!err.empty()) return fail(BP_ERR_INVALID, err);
    }
    h->hC = HB.K;
    BdConst& K = h->hC;
    BdParams& P = K.P;
    std::vector<BdLayout>& LY = HB.LY;
    // ---- upload
    // map slots: one per room (host-built maps), or one per env (built on the device at reset)
    const size_t nv = (size_t)H * W, R = device_maps ? (size_t)c.n_envs : (size_t)n_rooms;
    std::vector<short>& pa = HB.pa; std::vector<short>& pb = HB.pb;
    CUDA_TRY(cudaMalloc((void**)&h->d_layouts, sizeof(BdLayout) * n_layouts));
    CUDA_TRY(cudaMemcpy(h->d_layouts, LY.data(), sizeof(BdLayout) * n_layouts, cudaMemcpyHostToDevice));
    CUDA_TRY(cudaMalloc((void**)&h->d_pa, 2 * P.n_slots)); CUDA_TRY(cudaMalloc((void**)&h->d_pb, 2 * P.n_slots));
    CUDA_TRY(cudaMemcpy(h->d_pa, pa.data(), 2 * P.n_slots, cudaMemcpyHostToDevice));
    CUDA_TRY(cudaMemcpy(h->d_pb, pb.data(), 2 * P.n_slots, cudaMemcpyHostToDevice));
    CUDA_TRY(cudaMalloc((void**)&h->d_cspace, R * nv)); CUDA_TRY(cudaMalloc((void**)&h->d_thin, R * nv));
    CUDA_TRY(cudaMalloc((void**)&h->d_closest, R * nv * 8)); CUDA_TRY(cudaMalloc((void**)&h->d_small, R * P.SMALL * P.SMALL));
    if (!device_maps) {
        CUDA_TRY(cudaMemcpy(h->d_cspace, cspace, R * nv, cudaMemcpyHostToDevice)); CUDA_TRY(cudaMemcpy(h->d_thin, cspace_thin, R * nv, cudaMemcpyHostToDevice));
        CUDA_TRY(cudaMemcpy(h->d_closest, closest, R * nv * 8, cudaMemcpyHostToDevice));
        CUDA_TRY(cudaMemcpy(h->d_small, small_free, R * P.SMALL * P.SMALL, cudaMemcpyHostToDevice));
    } else {
        CUDA_TRY(cudaMemset(h->d_cspace, 0, R * nv)); CUDA_TRY(cudaMemset(h->d_thin, 0, R * nv));
        CUDA_TRY(cudaMemset(h->d_closest, 0, R * nv * 8)); CUDA_TRY(cudaMemset(h->d_small, 0, R * P.SMALL * P.SMALL));
    }
    CUDA_TRY(cudaMalloc((void**)&h->d_rchan, (size_t)P.LW * P.LW)); CUDA_TRY(cudaMemcpy(h->d_rchan, robot_chan, (size_t)P.LW * P.LW, cudaMemcpyHostToDevice));
    CUDA_TRY(cudaMalloc((void**)&h->d_recept, R * nv * 4));
    P.border_blocked = device_maps ? 1 : bd_border_blocked(cspace, R, H, W);
    // receptacle cell of every room (closest valid cell of the receptacle position)
    if (!device_maps) {
        int pi = (int)std::floor(H / 2.0 - P.recept_y * P.ppm), pj = (int)std::floor(W / 2.0 + P.recept_x * P.ppm);
        pi = pi < 0 ? 0 : (pi > H - 1 ? H - 1 : pi); pj = pj < 0 ? 0 : (pj > W - 1 ? W - 1 : pj);
        for (size_t r = 0; r < R; r++) {
            const int32_t* cl = closest + r * 2 * nv;
            HB.RM[r].recept_ci = cl[(size_t)pi * W + pj]; HB.RM[r].recept_cj = cl[nv + (size_t)pi * W + pj];
        }
    }
    CUDA_TRY(cudaMalloc((void**)&h->d_rooms, sizeof(BdStatics) * (size_t)n_rooms));
    CUDA_TRY(cudaMemcpy(h->d_rooms, HB.RM.data(), sizeof(BdStatics) * (size_t)n_rooms, cudaMemcpyHostToDevice));
    K.layouts = h->d_layouts; K.n_layouts = n_layouts; K.n_envs = c.n_envs; K.pair_a = h->d_pa; K.pair_b = h->d_pb;
    K.rooms = h->d_rooms; K.n_rooms = n_rooms; K.maps_per_env = device_maps ? 1 : 0;
    K.cspace = h->d_cspace; K.cspace_thin = h->d_thin; K.closest = h->d_closest; K.small_free = h->d_small; K.robot_chan = h->d_rchan; K.ch3_map = h->d_recept;
    K.env_bytes = bd_env_bytes(P.n_slots);
    CUDA_TRY(cudaMalloc((void**)&h->d_envs, K.env_bytes * c.n_envs));
    CUDA_TRY(cudaMemset(h->d_envs, 0, K.env_bytes * c.n_envs));
    K.envs = h->d_envs;
    int sms = 0; CUDA_TRY(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, c.device));
    h->n_blocks = sms * 4;
    // Reference-order parents: a solve is one thread's sequential queue (the rest of the block only initialises), so k_bd_dist
    // runs as many 64-thread blocks as fit an SM (80 registers: 12) instead of 3 blocks of 256 threads — 4x the concurrent solves.
    h->dist_threads = P.exact_parents ? 64 : 256;
    h->dist_blocks = P.exact_parents ? sms * 12 : h->n_blocks;
    if (const char* e = getenv("BP_BD_DIST_WIDE")) { if (atoi(e) != 0) { h->dist_threads = 256; h->dist_blocks = h->n_blocks; } }      // A/B measurements
    const int max_blocks = h->dist_blocks > h->n_blocks ? h->dist_blocks : h->n_blocks;
    CUDA_TRY(cudaMalloc((void**)&h->d_scratch, (size_t)max_blocks * bd_scratch_ints(P) * sizeof(int)));
    CUDA_TRY(cudaMalloc((void**)&h->d_counter, 64));
    CUDA_TRY(cudaMalloc((void**)&h->d_mv, sizeof(MoveVars) * (size_t)c.n_envs));
    CUDA_TRY(cudaMallocHost((void**)&h->h_flag, 64));
    CUDA_TRY(cudaMalloc((void**)&h->dC, sizeof(BdConst)));
    CUDA_TRY(cudaMemcpy(h->dC, &K, sizeof(BdConst), cudaMemcpyHostToDevice));
    k_bd_init_statics<<<n_rooms, 32>>>(h->d_rooms, P.n_static);
    CUDA_TRY(cudaGetLastError());
    if (device_maps) {
        // maps, receptacle field: built per env at reset; the receptacle does not see every free cell of a room with obstacles
        P.recept_visible = 0; P.border_blocked = 1;      // the padded room's margin is wider than the dilation disk: borders are blocked
        CUDA_TRY(cudaMemcpy(h->dC, &K, sizeof(BdParams), cudaMemcpyHostToDevice));
        h->device_maps = true;
        h->loaded = true;
        return BP_OK;
    }
    // fourth observation channel, static per room: the caller's map (area-clearing goal-point map), or the scaled
    // receptacle-source distance field (the reference recomputes it every step, box_delivery_env.py:1115-1129)
    if (ch3_map) {
        CUDA_TRY(cudaMemcpy(h->d_recept, ch3_map, R * nv * 4, cudaMemcpyHostToDevice));
    } else {
        if (c.env_kind != 0) return fail(BP_ERR_INVALID, "area-clearing needs the goal-point map (ch3_map)");
        std::vector<int32_t> src(2 * R);
        for (size_t r = 0; r < R; r++) { src[2 * r] = HB.RM[r].recept_ci; src[2 * r + 1] = HB.RM[r].recept_cj; }
        int32_t* d_src = nullptr; int32_t* d_par = nullptr;
        // one spfa solve per room, in chunks (the parents are not needed but the solver writes them)
        const int chunk = (int)(R < 64 ? R : 64);
        CUDA_TRY(cudaMalloc((void**)&d_src, 8 * R)); CUDA_TRY(cudaMalloc((void**)&d_par, (size_t)chunk * nv * 4));
        CUDA_TRY(cudaMemcpy(d_src, src.data(), 8 * R, cudaMemcpyHostToDevice));
        int rc = BP_OK; cudaError_t e = cudaSuccess;
        for (size_t r0 = 0; r0 < R && rc == BP_OK && e == cudaSuccess; r0 += chunk) {
            int nb = (int)((R - r0) < (size_t)chunk ? (R - r0) : chunk);
            rc = bp_spfa_batch(h->d_cspace + r0 * nv, d_src + 2 * r0, nb, H, W, h->d_recept + r0 * nv, d_par, nullptr);
            e = cudaDeviceSynchronize();
        }
        cudaFree(d_src); cudaFree(d_par);
        if (rc != BP_OK) return rc;
        if (e != cudaSuccess) return fail(BP_ERR_CUDA, std::string("receptacle spfa: ") + cudaGetErrorString(e));
        std::vector<float> tmp(R * nv);
        CUDA_TRY(cudaMemcpy(tmp.data(), h->d_recept, R * nv * 4, cudaMemcpyDeviceToHost));
        bd_scale_receptacle_map(P, tmp.data(), R * nv);
        CUDA_TRY(cudaMemcpy(h->d_recept, tmp.data(), R * nv * 4, cudaMemcpyHostToDevice));
    }
    // can the receptacle cell see the whole free space of every room?  (BdParams::recept_visible)
    {
        P.recept_visible = 0;
        CUDA_TRY(cudaMemcpy(h->dC, &K, sizeof(BdParams), cudaMemcpyHostToDevice));      // P is the first member of BdConst
        int one = 1;
        CUDA_TRY(cudaMemcpy(h->d_counter, &one, 4, cudaMemcpyHostToDevice));
        k_bd_visibility<<<dim3((unsigned)((nv + 255) / 256), (unsigned)n_rooms), 256>>>(h->dC, h->d_counter);
        CUDA_TRY(cudaGetLastError());
        CUDA_TRY(cudaMemcpy(&one, h->d_counter, 4, cudaMemcpyDeviceToHost));
        P.recept_visible = (one != 0 && !h->cfg.force_box_spfa) ? 1 : 0;
        if (c.env_kind == 1) P.recept_visible = 1;          // no receptacle distances in area-clearing
        CUDA_TRY(cudaMemcpy(h->dC, &K, sizeof(BdParams), cudaMemcpyHostToDevice));
    }
    h->loaded = true;
    return BP_OK;
}

static BdOut bd_out(const bp_boxdel_outputs* o) {
    BdOut r; memset(&r, 0, sizeof(r));
    if (o) { r.reward = o->reward; r.terminated = o->terminated; r.truncated = o->truncated; r.info = o->info; r.info_i = o->info_i; r.completed = o->completed; r.box_verts = o->box_verts; }
    return r;
}

static int bd_launch_obs(bp_boxdel* h, const uint8_t* mask, uint8_t* obs, cudaStream_t st) {
    if (!obs) return BP_OK;
    CUDA_TRY(cudaMemsetAsync(h->d_counter, 0, 4, st));
    int blocks = h->n_blocks < h->cfg.n_envs ? h->n_blocks : h->cfg.n_envs;
    k_bd_obs<<<blocks, 256, 0, st>>>(h->dC, mask, obs, h->d_scratch, h->d_counter);
    h->launches++;
    CUDA_TRY(cudaGetLastError());
    return BP_OK;
}

int bp_boxdel_reset(bp_boxdel_t* h, const int32_t* layout_ids, const uint8_t* mask, const bp_boxdel_outputs* out, void* stream) {
    if (!h || !h->loaded) return fail(BP_ERR_STATE, "bp_boxdel_load first");
    CUDA_TRY(cudaSetDevice(h->cfg.device));
    cudaStream_t st = (cudaStream_t)stream;
    int E = h->cfg.n_envs;
    if (h->device_maps) {
        CUDA_TRY(cudaMemsetAsync(h->d_counter, 0, 4, st));
        int blocks = h->n_blocks < E ? h->n_blocks : E;
        k_bd_build_rooms<<<blocks, 256, 0, st>>>(h->dC, layout_ids, mask, h->d_scratch, h->d_counter);
        h->launches++;
        CUDA_TRY(cudaGetLastError());
    }
    k_bd_reset<<<(E * 32 + 127) / 128, 128, 0, st>>>(h->dC, layout_ids, mask, bd_out(out));
    h->launches++;
    CUDA_TRY(cudaGetLastError());
    return bd_launch_obs(h, mask, out ? out->obs : nullptr, st);
}

int bp_boxdel_step(bp_boxdel_t* h, const float* actions, const uint8_t* mask, const bp_boxdel_outputs* out, void* stream) {
    if (!h || !h->loaded) return fail(BP_ERR_STATE, "bp_boxdel_load first");
    if (!actions) return fail(BP_ERR_INVALID, "null actions");
    CUDA_TRY(cudaSetDevice(h->cfg.device));
    cudaStream_t st = (cudaStream_t)stream;
    int E = h->cfg.n_envs;
    int items = E * (h->cfg.num_boxes + 1);
    int blocks = h->dist_blocks < items ? h->dist_blocks : items;
    k_bd_begin<<<(E + 127) / 128, 128, 0, st>>>(h->dC, actions, mask);
    CUDA_TRY(cudaMemsetAsync(h->d_counter, 0, 4, st));
    k_bd_dist<<<blocks, h->dist_threads, 0, st>>>(h->dC, mask, 0, h->d_scratch, h->d_counter);
    cudaEvent_t e4[4] = {nullptr, nullptr, nullptr, nullptr};
    if (h->profiling) { for (auto& e : e4) { CUDA_TRY(cudaEventCreate(&e)); h->ev.push_back(e); } CUDA_TRY(cudaEventRecord(e4[0], st)); }
    if (h->cfg.phase_kernels) {
        // one kernel per phase over the whole batch; the number of iterations is data dependent (until every env's loop has
        // ended), so the host enqueues chunks and reads one flag per chunk (this path synchronises the stream once per chunk)
        const int wb = (E * 32 + 127) / 128;
        k_mv_init<<<wb, 128, 0, st>>>(h->dC, mask, h->d_mv);
        int* d_flag = h->d_counter + 8;
        for (long it = 0;; it += h->phase_chunk) {
            for (int i = 0; i < h->phase_chunk; i++) {
                k_mv_pre_a<<<wb, 128, 0, st>>>(h->dC, h->d_mv);
                k_mv_b<<<wb, 128, 0, st>>>(h->dC, h->d_mv);
                k_mv_c<<<wb, 128, 0, st>>>(h->dC, h->d_mv);
                k_mv_d_post<<<wb, 128, 0, st>>>(h->dC, h->d_mv);
            }
            h->launches += 4L * h->phase_chunk + 1;
            CUDA_TRY(cudaMemsetAsync(d_flag, 0, 4, st));
            k_mv_alive<<<(E + 255) / 256, 256, 0, st>>>(h->dC, h->d_mv, d_flag);
            CUDA_TRY(cudaMemcpyAsync(h->h_flag, d_flag, 4, cudaMemcpyDeviceToHost, st));
            CUDA_TRY(cudaStreamSynchronize(st));
            if (!*h->h_flag || it > 40000) break;
        }
    } else {
        int warps = (E + BD_MOVE_K - 1) / BD_MOVE_K;
        k_bd_move<<<(warps * 32 + BD_MOVE_BLOCK - 1) / BD_MOVE_BLOCK, BD_MOVE_BLOCK, 0, st>>>(h->dC, mask);
    }
    if (h->profiling) CUDA_TRY(cudaEventRecord(e4[1], st));
    if (!h->hC.P.recept_visible) {
        CUDA_TRY(cudaMemsetAsync(h->d_counter, 0, 4, st));
        k_bd_dist<<<blocks, h->dist_threads, 0, st>>>(h->dC, mask, 1, h->d_scratch, h->d_counter);
        h->launches++;
    }
    k_bd_finish<<<(E + 127) / 128, 128, 0, st>>>(h->dC, mask, bd_out(out));
    h->launches += 4;
    CUDA_TRY(cudaGetLastError());
    if (h->profiling) CUDA_TRY(cudaEventRecord(e4[2], st));
    int rc = bd_launch_obs(h, mask, out ? out->obs : nullptr, st);
    if (h->profiling) CUDA_TRY(cudaEventRecord(e4[3], st));
    return rc;
}

int bp_boxdel_set_profiling(bp_boxdel_t* h, int32_t on) { if (!h) return fail(BP_ERR_INVALID, "null handle"); h->profiling = on != 0; return BP_OK; }
// HOST out[3]: summed k_bd_move ms, summed k_bd_obs ms, timed steps since the last call.  Synchronises; clears the events.
int bp_boxdel_get_kernel_ms(bp_boxdel_t* h, double* out3) {
    if (!h || !out3) return fail(BP_ERR_INVALID, "null argument");
    CUDA_TRY(cudaSetDevice(h->cfg.device));
    CUDA_TRY(cudaDeviceSynchronize());
    out3[0] = out3[1] = 0.0; out3[2] = (double)(h->ev.size() / 4);
    for (size_t i = 0; i + 3 < h->ev.size(); i += 4) {
        float a = 0, b = 0;
        CUDA_TRY(cudaEventElapsedTime(&a, h->ev[i], h->ev[i + 1]));
        CUDA_TRY(cudaEventElapsedTime(&b, h->ev[i + 2], h->ev[i + 3]));
        out3[0] += a; out3[1] += b;
    }
    for (auto e : h->ev) cudaEventDestroy(e);
    h->ev.clear();
    return BP_OK;
}

// host-buffer variant: actions and every non-null output are HOST pointers; copies in, steps, copies out, synchronises
static int bd_step_host_impl(bp_boxdel_t* h, const float* actions, const bp_boxdel_outputs* out_host, uint8_t* obs_dev);
int bp_boxdel_step_host(bp_boxdel_t* h, const float* actions, const bp_boxdel_outputs* out_host) { return bd_step_host_impl(h, actions, out_host, nullptr); }
// As bp_boxdel_step_host, with the observation rendered into the DEVICE buffer obs_dev (u8[E][LW][LW][4]) instead of being copied
// to the host: the call a GPU-resident policy makes (host actions in, host rewards / flags / info out, images stay on the device).
int bp_boxdel_step_host_obs_device(bp_boxdel_t* h, const float* actions, const bp_boxdel_outputs* out_host, uint8_t* obs_dev) {
    if (!obs_dev) return fail(BP_ERR_INVALID, "null obs_dev");
    return bd_step_host_impl(h, actions, out_host, obs_dev);
}
static int bd_step_host_impl(bp_boxdel_t* h, const float* actions, const bp_boxdel_outputs* out_host, uint8_t* obs_dev) {
    if (!h || !h->loaded) return fail(BP_ERR_STATE, "bp_boxdel_load first");
    if (!actions || !out_host) return fail(BP_ERR_INVALID, "null argument");
    CUDA_TRY(cudaSetDevice(h->cfg.device));
    const int E = h->cfg.n_envs, NB = h->cfg.num_boxes;
    const size_t obs_b = (size_t)E * h->hC.P.LW * h->hC.P.LW * 4;
    const size_t act_b = 4 * (size_t)E * (h->cfg.action_type == 2 ? 2 : 1);
    float* d_act = nullptr; CUDA_TRY(cudaMalloc((void**)&d_act, act_b));
    CUDA_TRY(cudaMemcpy(d_act, actions, act_b, cudaMemcpyHostToDevice));
    bp_boxdel_outputs d; memset(&d, 0, sizeof(d));
    struct Item { void** dev; void* host; size_t bytes; };
    Item items[8] = {{(void**)&d.obs, out_host->obs, obs_b}, {(void**)&d.reward, out_host->reward, 8 * (size_t)E},
                     {(void**)&d.terminated, out_host->terminated, (size_t)E}, {(void**)&d.truncated, out_host->truncated, (size_t)E},
                     {(void**)&d.info, out_host->info, 64 * (size_t)E}, {(void**)&d.info_i, out_host->info_i, 16 * (size_t)E},
                     {(void**)&d.completed, out_host->completed, (size_t)E * NB}, {(void**)&d.box_verts, out_host->box_verts, (size_t)E * NB * SV * 16}};
    int rc = BP_OK;
    if (obs_dev) items[0].host = nullptr;          // rendered straight into the caller's device buffer
    for (auto& it : items) if (it.host && cudaMalloc(it.dev, it.bytes) != cudaSuccess) rc = fail(BP_ERR_CUDA, "cudaMalloc (step_host)");
    if (obs_dev) d.obs = obs_dev;
    if (rc == BP_OK) rc = bp_boxdel_step(h, d_act, nullptr, &d, nullptr);
    if (rc == BP_OK) for (auto& it : items) if (it.host && cudaMemcpy(it.host, *it.dev, it.bytes, cudaMemcpyDeviceToHost) != cudaSuccess) rc = fail(BP_ERR_CUDA, "copy out (step_host)");
    if (rc == BP_OK && obs_dev && cudaDeviceSynchronize() != cudaSuccess) rc = fail(BP_ERR_CUDA, "step_host_obs_device");
    if (obs_dev) d.obs = nullptr;
    for (auto& it : items) if (*it.dev) cudaFree(*it.dev);
    cudaFree(d_act);
    return rc;
}

// HOST buffers (any may be NULL): the maps env `env` currently uses — cspace / cspace_thin u8[H][W], closest i32[2][H][W],
// small_free u8[S][S], ch3 f32[H][W].  Parity / debugging aid (the device-built rooms are compared with the host build).  Synchronises.
int bp_boxdel_get_room_maps(bp_boxdel_t* h, int32_t env, uint8_t* cspace, uint8_t* cspace_thin, int32_t* closest, uint8_t* small_free, float* ch3) {
    if (!h || !h->loaded || env < 0 || env >= h->cfg.n_envs) return fail(BP_ERR_INVALID, "bad argument");
    CUDA_TRY(cudaSetDevice(h->cfg.device));
    CUDA_TRY(cudaDeviceSynchronize());
    BdEnv hdr;
    CUDA_TRY(cudaMemcpy(&hdr, h->d_envs + h->hC.env_bytes * env, sizeof(BdEnv), cudaMemcpyDeviceToHost));
    const BdParams& P = h->hC.P;
    const size_t nv = (size_t)P.H * P.W, s = (size_t)hdr.map_slot, nn = (size_t)P.SMALL * P.SMALL;
    if (cspace) CUDA_TRY(cudaMemcpy(cspace, h->d_cspace + s * nv, nv, cudaMemcpyDeviceToHost));
    if (cspace_thin) CUDA_TRY(cudaMemcpy(cspace_thin, h->d_thin + s * nv, nv, cudaMemcpyDeviceToHost));
    if (closest) CUDA_TRY(cudaMemcpy(closest, h->d_closest + s * 2 * nv, nv * 8, cudaMemcpyDeviceToHost));
    if (small_free) CUDA_TRY(cudaMemcpy(small_free, h->d_small + s * nn, nn, cudaMemcpyDeviceToHost));
    if (ch3) CUDA_TRY(cudaMemcpy(ch3, h->d_recept + s * nv, nv * 4, cudaMemcpyDeviceToHost));
    return BP_OK;
}

// HOST out[(1 + num_boxes)][7]: body origin x, y, angle, vx, vy, w, alive — row 0 = robot.  Synchronises.
int bp_boxdel_get_bodies(bp_boxdel_t* h, int32_t env, double* out) {
    if (!h || !h->loaded || env < 0 || env >= h->cfg.n_envs || !out) return fail(BP_ERR_INVALID, "bad argument");
    CUDA_TRY(cudaSetDevice(h->cfg.device));
    CUDA_TRY(cudaDeviceSynchronize());
    std::vector<char> buf(sizeof(BdEnv));
    CUDA_TRY(cudaMemcpy(buf.data(), h->d_envs + h->hC.env_bytes * env, sizeof(BdEnv), cudaMemcpyDeviceToHost));
    const BdEnv* E = (const BdEnv*)buf.data();
    for (int b = 0; b <= h->cfg.num_boxes; b++) {
        const BdBody& B = E->body[b];
        double ox = B.px - (B.cgx * B.rc - B.cgy * B.rs), oy = B.py - (B.cgx * B.rs + B.cgy * B.rc);
        double* o = out + 7 * b;
        o[0] = ox; o[1] = oy; o[2] = B.pa; o[3] = B.vx; o[4] = B.vy; o[5] = B.vw; o[6] = B.alive;
    }
    return BP_OK;
}

// HOST out[16]: n_path, move_sign, need_spfa, status, robot_hit, substeps so far, candidate pairs so far, arbiters so far,
// robot_distance, then init_dist / final_dist are returned through dists[2][num_boxes]; path[24][3].  Debug / parity aid.
int bp_boxdel_get_step_debug(bp_boxdel_t* h, int32_t env, double* out16, double* dists, double* path) {
    if (!h || !h->loaded || env < 0 || env >= h->cfg.n_envs) return fail(BP_ERR_INVALID, "bad argument");
    CUDA_TRY(cudaSetDevice(h->cfg.device));
    CUDA_TRY(cudaDeviceSynchronize());
    const size_t hdr = BP_ALIGN128(sizeof(BdEnv));
    std::vector<char> buf(hdr + sizeof(BdEnvCold));
    CUDA_TRY(cudaMemcpy(buf.data(), h->d_envs + h->hC.env_bytes * env, buf.size(), cudaMemcpyDeviceToHost));
    const BdEnv* E = (const BdEnv*)buf.data();
    const BdEnvCold* K = (const BdEnvCold*)(buf.data() + hdr);
    if (out16) {
        out16[0] = E->n_path; out16[1] = E->move_sign; out16[2] = E->need_spfa; out16[3] = E->status; out16[4] = E->robot_hit;
        out16[5] = (double)E->stat_substeps; out16[6] = (double)E->stat_cand; out16[7] = (double)E->stat_arb; out16[8] = E->robot_distance;
        for (int i = 9; i < 16; i++) out16[i] = 0.0;
    }
    if (dists) for (int i = 0; i < h->cfg.num_boxes; i++) { dists[i] = K->init_dist[i]; dists[h->cfg.num_boxes + i] = K->final_dist[i]; }
    if (path) memcpy(path, K->path, sizeof(K->path));
    return BP_OK;
}

int64_t bp_boxdel_launch_count(bp_boxdel_t* h) { return h ? h->launches : 0; }
size_t bp_boxdel_env_bytes(bp_boxdel_t* h) { return h ? h->hC.env_bytes : 0; }

}  // extern "C"
