// C ABI of libbenchpush_b200.so (declared in include/benchpush_b200.h): handle management, layout
// upload, kernel launches.  No CPU fallback: every entry point fails with BP_ERR_CUDA when the
// CUDA runtime cannot run the kernels.
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/benchpush_b200.h"
#include "host_tables.h"
#include "raster.cuh"
#include "maze_raster.cuh"
// the batched ship-ice / maze step, instantiated for 8 and for 4 envs per warp (sim_batch.cuh)
#ifndef BP_BK_MAIN          // make BK=<n>: envs per warp of the main instantiation (tuning experiments)
#define BP_BK_MAIN 8
#endif
#define BP_BK BP_BK_MAIN
#define BP_BATCH_NS bk_main
#include "sim_batch.cuh"
// the resident step (awake world in shared memory) builds on the main instantiation
#define BP_RES_NS bk_main
#include "sim_res.cuh"
#undef BP_BK
#undef BP_BATCH_NS
#define BP_BK 4
#define BP_BATCH_NS bk_small
#include "sim_batch.cuh"
#undef BP_BK
#undef BP_BATCH_NS
#include "spfa.cuh"

using namespace bp;

static thread_local std::string g_err;
static int fail(int code, const std::string& msg) { g_err = msg; return code; }
#define CUDA_TRY(expr)                                                                            \
    do {                                                                                          \
        cudaError_t _e = (expr);                                                                  \
        if (_e != cudaSuccess) return fail(BP_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(_e)); \
    } while (0)

// ------------------------------------------------------------------------------------------------ kernels
// Persistent warps: each warp pulls batches of 32 / BP_NL environment indices from a global counter
// and simulates them side by side, one per lane group (envs differ widely in cost: contact-free
// water vs. an ice jam, so the batches are claimed dynamically).
#ifndef BP_MIN_BLOCKS
#define BP_MIN_BLOCKS 4          // resident 128-thread blocks per SM the sim kernels are compiled for
#endif

struct StepArgs {
    char* mem; size_t env_bytes; Dims d; SimParams P; LayoutTab T; StepOut out;
    const float* actions; const int32_t* layout_ids; const uint8_t* mask; int* counter; int auto_reset;
    const DevConst* C; int use_snap;
    const int* perm; int n_sched;      // step schedule (k_schedule): perm[n_sched], entries = env id or -1; nullptr = identity
};

// Claim the next batch of environments for this warp: one per lane group.  Returns false when the
// queue is empty.  Full-warp operation: every lane of the warp must call it (the groups of a warp
// reconverge here after each batch, which keeps them in lockstep inside the substep loops).
__device__ __forceinline__ bool claim_envs(int* counter, int n_envs, int* env) {
    int base = 0;
    if ((threadIdx.x & 31) == 0) base = atomicAdd(counter, BP_GROUPS_PER_WARP);
    base = __shfl_sync(0xffffffffu, base, 0);
    *env = base + bp_group_in_warp();
    return base < n_envs;
}

__global__ void __launch_bounds__(128, BP_MIN_BLOCKS) k_shipice_reset(StepArgs a) {
    int env;
    while (claim_envs(a.counter, a.d.n_envs, &env)) {
        if (env < a.d.n_envs && !(a.mask && !a.mask[env])) {
            EnvView ev = make_view(a.mem + a.env_bytes * (size_t)env, a.d);
            int layout = a.layout_ids[env] % a.d.n_layouts;
            const double* sp = a.C->start_pose ? a.C->start_pose + 3 * (size_t)env : nullptr;
            if (a.use_snap && !sp) shipice_reset_from_snapshot(a.C, a.mem + a.env_bytes * (size_t)env, layout, env, a.out);
            else shipice_reset_env(ev, a.T, a.P, a.d, layout, env, a.out, sp);
            if (bp_lane() == 0) { ev.sc->first_layout = layout; ev.sc->episode = 0; ev.sc->done = 0; ev.sc->obs_raw = 1; }
            bp_wsync();
        }
        __syncwarp();
    }
}

// bp_*_set_bodies: one warp puts one env into an externally given body state.
__global__ void __launch_bounds__(32) k_shipice_set_bodies(StepArgs a, int env, const double* bodies6) {
    EnvView ev = make_view(a.mem + a.env_bytes * (size_t)env, a.d);
    shipice_set_bodies_env(ev, a.T, a.P, a.d, bodies6);
}

// One env per warp (the first design; kept for A/B measurements: BP_STEP_IMPL=warp).
__global__ void __launch_bounds__(128, BP_MIN_BLOCKS) k_shipice_step_warp(StepArgs a) {
    int env;
    while (claim_envs(a.counter, a.d.n_envs, &env)) {
        if (env < a.d.n_envs && !(a.mask && !a.mask[env])) {
            EnvView ev = make_view(a.mem + a.env_bytes * (size_t)env, a.d);
            if (a.auto_reset && ev.sc->done) {
                // next-step auto reset: the env moves on to its next trial (episode_idx % n_trials,
                // ship_ice_env.py:188); this call returns the reset observation with reward 0.
                int ep = ev.sc->episode + 1, first = ev.sc->first_layout;
                int layout = (first + ep) % a.d.n_layouts;
                const double* sp = a.C->start_pose ? a.C->start_pose + 3 * (size_t)env : nullptr;
                if (a.use_snap && !sp) shipice_reset_from_snapshot(a.C, a.mem + a.env_bytes * (size_t)env, layout, env, a.out);
                else shipice_reset_env(ev, a.T, a.P, a.d, layout, env, a.out, sp);
                if (bp_lane() == 0) { ev.sc->first_layout = first; ev.sc->episode = ep; ev.sc->done = 0; ev.sc->obs_raw = 1; }
            } else {
                shipice_step_env(ev, a.T, a.P, a.d, (double)a.actions[env], env, a.out);
                if (bp_lane() == 0) { ev.sc->done = a.out.terminated[env]; ev.sc->obs_raw = 0; }
            }
            bp_wsync();
        }
        __syncwarp();
    }
}

// Step schedule.  The cost of an env step varies by more than an order of magnitude with the size of its awake set and the
// number of solver colours, and a warp batch runs as long as its slowest member (10 sweeps x the LARGEST colour count of the
// batch).  One block sorts the envs by a cost key read from their env blocks (colours, awake bodies; descending) and writes
// the order in which the step kernel's warps claim them:
//   * many batches per resident warp: batches of similar envs (few wasted solver rounds), heaviest first (the light ones
//     fill the tail: longest-processing-time-first);
//   * about one batch per resident warp (the per-GPU share of the 8-GPU strong-scaling point): the sorted envs are dealt
//     round-robin over the batches, so that every batch gets the same mix and the launch does not wait for one heavy batch.
// Results do not depend on the schedule (envs never interact; tests run the same envs under different schedules).
__global__ void __launch_bounds__(1024) k_schedule(const char* mem, size_t env_bytes, int n_envs, const uint8_t* mask, int auto_reset,
                                                    int* perm, int bk, int n_batches, int balanced, long long* prev, int* keybuf, int hist_all) {
    __shared__ int hist[256];
    __shared__ int start[256];
    const int n_sched = n_batches * bk;
    for (int i = threadIdx.x; i < 256; i += blockDim.x) hist[i] = 0;
    for (int i = threadIdx.x; i < n_sched; i += blockDim.x) perm[i] = -1;
    __syncthreads();
    auto key_of = [&](int e) -> int {
        if (mask && !mask[e]) return 255;                                   // not stepping: last
        const EnvScal* sc = (const EnvScal*)(mem + env_bytes * (size_t)e);
        if (auto_reset && sc->done) return 254;                             // reset = copy of a snapshot: light
        if (prev && (balanced || hist_all)) {
            // balanced dealing wants the env's COST: arbiters and awake bodies averaged over its previous step (the counters the
            // substep loop keeps), not the state it happens to be in at the end of it
            long long ds = sc->stat_substeps - prev[3 * (size_t)e + 2];
            long long dl = sc->stat_live_sum - prev[3 * (size_t)e], da = sc->stat_awake_sum - prev[3 * (size_t)e + 1];
            prev[3 * (size_t)e] = sc->stat_live_sum; prev[3 * (size_t)e + 1] = sc->stat_awake_sum; prev[3 * (size_t)e + 2] = sc->stat_substeps;
            if (ds > 0 && dl >= 0 && da >= 0) {
                long long cost = (2 * dl + da) / ds;
                return 253 - (int)(cost > 252 ? 252 : cost);
            }
        }
        int nc = sc->n_colours < 15 ? sc->n_colours : 15;
        int na = sc->n_awake / 4; if (na > 14) na = 14;
        return 253 - (nc * 15 + na);                                        // descending cost
    };
    for (int e = threadIdx.x; e < n_envs; e += blockDim.x) { int k = key_of(e); keybuf[e] = k; atomicAdd(&hist[k], 1); }
    __syncthreads();
    if (threadIdx.x == 0) { int acc = 0; for (int i = 0; i < 256; i++) { start[i] = acc; acc += hist[i]; } }
    __syncthreads();
    for (int e = threadIdx.x; e < n_envs; e += blockDim.x) {
        int p = atomicAdd(&start[keybuf[e]], 1);
        // balanced: deal the sorted envs over the batches like cards; balanced == 2 reverses every other round (snake order), so that
        // the batch that got the heaviest env of one round gets the lightest of the next
        int round = p / n_batches, q = p % n_batches;
        if (balanced == 2 && (round & 1)) q = n_batches - 1 - q;
        int idx = balanced ? q * bk + round : p;
        perm[idx] = e;
    }
}

// Batched step: each warp claims BK consecutive envs and advances them together (sim_batch.cuh).
#define BP_DEFINE_STEP_KERNEL(NAME, NS)                                                                         \
    __global__ void __launch_bounds__(128, BP_MIN_BLOCKS) NAME(StepArgs a) {                                    \
        __shared__ NS::BatchShared SH[4];                                                                        \
        NS::BatchShared& S = SH[threadIdx.x >> 5];                                                               \
        load_shared_offsets(a.C->off);                                                                           \
        for (;;) {                                                                                               \
            int env0 = 0;                                                                                        \
            if ((threadIdx.x & 31) == 0) env0 = atomicAdd(a.counter, NS::BK);                                    \
            env0 = __shfl_sync(0xffffffffu, env0, 0);                                                            \
            if (env0 >= (a.perm ? a.n_sched : a.d.n_envs)) break;                                                \
            NS::shipice_step_batch(a.C, S, a.mem, a.env_bytes, env0, a.actions, a.mask, a.out, a.auto_reset, a.perm); \
            __syncwarp();                                                                                        \
        }                                                                                                        \
    }
BP_DEFINE_STEP_KERNEL(k_shipice_step, bk_main)
BP_DEFINE_STEP_KERNEL(k_shipice_step_small, bk_small)

// Resident step: one warp per block, the awake world of its RK envs in dynamic shared memory (sim_res.cuh).
__global__ void __launch_bounds__(32) k_shipice_step_res(StepArgs a) {
    extern __shared__ __align__(128) unsigned char bp_res_smem[];
    res::ResShared& S = *reinterpret_cast<res::ResShared*>(bp_res_smem);
    load_shared_offsets(a.C->off);
    if (threadIdx.x == 0) { bp_mbar_init(&S.mbar, 32); S.mbar_phase = 0; }
    __syncwarp();
    for (;;) {
        int env0 = 0;
        if (threadIdx.x == 0) env0 = atomicAdd(a.counter, res::RK);
        env0 = __shfl_sync(0xffffffffu, env0, 0);
        if (env0 >= (a.perm ? a.n_sched : a.d.n_envs)) break;
        res::shipice_step_resident(a.C, S, a.mem, a.env_bytes, env0, a.actions, a.mask, a.out, a.auto_reset, a.perm);
        __syncwarp();
    }
}

struct ObsArgs {
    char* mem; size_t env_bytes; Dims d; LayoutTab T; ObsParams O;
    const double* rawv; const int* rawn; const double* rawcen; const uint8_t* mask; uint8_t* obs;
};

__global__ void __launch_bounds__(256) k_shipice_obs(ObsArgs a) {
    __shared__ ObsSmem S;
    int env = blockIdx.x;
    if (a.mask && !a.mask[env]) return;
    EnvView ev = make_view(a.mem + a.env_bytes * (size_t)env, a.d);
    const EnvScal* sc = ev.sc;
    LayoutView L = layout_view(a.T, a.d, sc->layout);
    int r = sc->robot;
    V2 o = body_origin(ev, L, r);
    uint8_t* out = a.obs + (size_t)env * BP_OBS_C * OBS_PIX;
    if (sc->obs_raw) {
        size_t base = (size_t)sc->layout * a.d.NB;
        render_obs(S, a.O, r, a.rawv + base * VMAX * 2, VMAX * 2, a.rawn + base, a.rawcen + base * 2, a.rawcen + base * 2 + 1, 2,
                   o.x, o.y, ev.body[r].pa, out);
    } else {
        render_obs(S, a.O, r, ev.wv, VMAX * 2, L.nv, ev.cenx, ev.ceny, 1, o.x, o.y, ev.body[r].pa, out);
    }
}

// cfg.egocentric_obs = false: the 2 x 200 x 60 cell-grid observation (raster.cuh: render_obs_global), one block per env
__global__ void __launch_bounds__(256) k_shipice_obs_global(ObsArgs a, GlobalObsParams G) {
    __shared__ ObsGlobalSmem S;
    int env = blockIdx.x;
    if (a.mask && !a.mask[env]) return;
    EnvView ev = make_view(a.mem + a.env_bytes * (size_t)env, a.d);
    const EnvScal* sc = ev.sc;
    LayoutView L = layout_view(a.T, a.d, sc->layout);
    int r = sc->robot;
    V2 o = body_origin(ev, L, r);
    uint8_t* out = a.obs + (size_t)env * 2 * G.gh * G.gw;
    if (sc->obs_raw) {
        size_t base = (size_t)sc->layout * a.d.NB;
        render_obs_global(S, a.O, G, r, a.rawv + base * VMAX * 2, VMAX * 2, a.rawn + base, o.x, o.y, ev.body[r].pa, out);
    } else {
        render_obs_global(S, a.O, G, r, ev.wv, VMAX * 2, L.nv, o.x, o.y, ev.body[r].pa, out);
    }
}

struct MazeObsArgs {
    char* mem; size_t env_bytes; Dims d; LayoutTab T; MazeObsParams O;
    const double* rawv; const int* rawn; const uint8_t* mask; uint8_t* obs;
};

// maze-NAMO observation: one block per env, the 288x288 crop lives in (dynamic) shared memory
__global__ void __launch_bounds__(256) k_maze_obs(MazeObsArgs a) {
    extern __shared__ __align__(16) unsigned char mz_smem_raw[];
    MazeObsSmem& S = *reinterpret_cast<MazeObsSmem*>(mz_smem_raw);
    int env = blockIdx.x;
    if (a.mask && !a.mask[env]) return;
    EnvView ev = make_view(a.mem + a.env_bytes * (size_t)env, a.d);
    const EnvScal* sc = ev.sc;
    LayoutView L = layout_view(a.T, a.d, sc->layout);
    int r = sc->robot;
    V2 o = body_origin(ev, L, r);
    uint8_t* out = a.obs + (size_t)env * MZ_OBS_C * MZ_PIX;
    int first = 0;                                                   // bodies [n_walls, robot0) are the movable obstacles
    while (first < L.robot0 && L.nv[first] == 2) first++;            // skip the wall segments
    int n_obst = L.robot0 - first;
    if (sc->obs_raw) {
        size_t base = (size_t)sc->layout * a.d.NB + first;
        render_maze_obs(S, a.O, n_obst, a.rawv + base * VMAX * 2, VMAX * 2, a.rawn + base, o.x, o.y, ev.body[r].pa, out);
    } else {
        render_maze_obs(S, a.O, n_obst, ev.wv + (size_t)first * VMAX * 2, VMAX * 2, L.nv + first, o.x, o.y, ev.body[r].pa, out);
    }
}

__global__ void __launch_bounds__(512) k_spfa(const uint8_t* maps, const int32_t* sources, int batch, int H, int W,
                                               float* dists, int32_t* parents, int* scratch, int* counter) {
    __shared__ SpfaShared S;
    __shared__ int s_map;
    const size_t nv = (size_t)H * W;
    int* q0 = scratch + (size_t)blockIdx.x * 3 * nv;
    for (;;) {
        if (threadIdx.x == 0) s_map = atomicAdd(counter, 1);
        __syncthreads();
        int m = s_map;
        __syncthreads();
        if (m >= batch) break;
        spfa_block(S, maps + m * nv, H, W, sources[2 * m], sources[2 * m + 1], dists + m * nv, parents + m * nv, q0, q0 + nv, q0 + 2 * nv);
    }
}

// one warp per map: the lanes initialise, lane 0 restates the reference solver's queue (spfa_serial_run), the lanes apply the
// reference's output convention (unreachable -> 0)
__global__ void __launch_bounds__(32) k_spfa_serial(const uint8_t* maps, const int32_t* sources, int batch, int H, int W, float* dists, int32_t* parents,
                                                     int* ring, unsigned char* inq) {
    const int m = blockIdx.x;
    if (m >= batch) return;
    const size_t nv = (size_t)H * W;
    float* d = dists + m * nv;
    spfa_serial_init(maps + m * nv, (int)nv, d, parents + m * nv, inq + m * nv, threadIdx.x, 32);
    __syncwarp();
    if (threadIdx.x == 0) spfa_serial_run(maps + m * nv, H, W, sources[2 * m], sources[2 * m + 1], d, parents + m * nv, ring + m * (nv + 2), inq + m * nv);
    __syncwarp();
    const float inf = 2.0f * (float)nv;
    for (size_t v = threadIdx.x; v < nv; v += 32) if (!(d[v] < inf) || d[v] < 0.0f) d[v] = 0.0f;      // unreachable and blocked cells report 0
}

// generate_observation_low_dim (ship_ice_env.py:358-370; maze_NAMO_env.py:488-504): |poly_centroid| of every movable
// polygon in creation order.  One warp per env: the |centroid| cache of the env block (cenx / ceny: refreshed by
// step_work for every body that moved, set at reset) is compacted over the dynamic bodies below the robot.
// maze = 1: <robot_x, robot_y> first, and — as the reference's loop starts at index 1 — obstacle 0 is never written
// and the last pair stays 0.
struct LowDimArgs { char* mem; size_t env_bytes; Dims d; LayoutTab T; double* out; int32_t* counts; int stride; int maze; };

__global__ void __launch_bounds__(128) k_low_dim(LowDimArgs a) {
    const int env = blockIdx.x * 4 + (threadIdx.x >> 5);
    if (env >= a.d.n_envs) return;
    const int lane = threadIdx.x & 31;
    EnvView ev = make_view(a.mem + a.env_bytes * (size_t)env, a.d);
    const EnvScal* sc = ev.sc;
    LayoutView L = layout_view(a.T, a.d, sc->layout);
    const int r = sc->robot;
    double* row = a.out + (size_t)env * a.stride;
    int n = 0;                                       // movable polygons seen so far (warp-uniform)
    for (int b0 = 0; b0 < r; b0 += 32) {
        int b = b0 + lane;
        bool dyn = b < r && body_dynamic(L, b);
        unsigned m = __ballot_sync(0xffffffffu, dyn);
        if (dyn) {
            int i = n + __popc(m & ((1u << lane) - 1u));
            if (!(a.maze && i == 0)) { row[2 * i] = ev.cenx[b]; row[2 * i + 1] = ev.ceny[b]; }
        }
        n += __popc(m);
    }
    const int used = a.maze ? 2 * (n + 1) : 2 * n;
    if (a.maze) {
        if (lane == 0) { V2 o = body_origin(ev, L, r); row[0] = o.x; row[1] = o.y; }
        if (lane == 1 && n > 0) { row[2 * n] = 0.0; row[2 * n + 1] = 0.0; }
    }
    for (int k = used + lane; k < a.stride; k += 32) row[k] = 0.0;
    if (lane == 0 && a.counts) a.counts[env] = used;
}

// ------------------------------------------------------------------------------------------------ handle
struct bp_shipice {
    bp_shipice_config cfg;
    Dims d; SimParams P; ObsParams O;
    int kind = KIND_SHIPICE;       // KIND_MAZE: created by bp_maze_create
    bp_maze_config mcfg; MazeConfig mz; MazeObsParams MO;
    std::vector<uint8_t> wall_map; std::vector<double> dist_norm, dist_raw;   // maze: host copies of the constant maps
    int max_boxes = 0;
    HostLayouts H;
    LayoutTab T;               // device pointers
    std::vector<void*> dev_allocs;
    double *d_rawv = nullptr, *d_rawcen = nullptr; int* d_rawn = nullptr;
    char* mem = nullptr; size_t env_bytes = 0;
    int* counter = nullptr;
    DevConst* d_const = nullptr;   // Dims, SimParams, array offsets, layout tables (device copy)
    DevConst hc_host;              // its host copy (bp_shipice_set_start_poses patches one pointer)
    char* snap = nullptr;          // post-reset env block of every layout (reset = copy)
    bool warp_impl = false;        // BP_STEP_IMPL=warp: one env per warp (A/B measurements)
    int* perm = nullptr; int perm_cap = 0;   // step schedule (k_schedule); BP_NO_SCHEDULE=1 keeps the identity order
    bool use_schedule = true;
    bool sched_history = true;         // balanced dealing by the previous step's average arbiters / awake bodies (BP_SCHED_HISTORY=0: by the current colours / awake count)
    long long* sched_prev = nullptr; int* sched_keys = nullptr;
    bool sched_hist_all = false;       // BP_SCHED_HISTORY=2: the history key also orders the grouped (many batches per warp) mode
    bool snake = true;                 // BP_NO_SNAKE=1: plain round-robin dealing (A/B measurements)
    bool pad_batches = true;           // BP_NO_PAD=1 turns the partly-filled-batch schedule off (A/B measurements)
    bool res_impl = false;         // resident step (sim_res.cuh): the default for ship-ice; BP_STEP_IMPL=batch selects sim_batch.cuh
    int res_blocks_per_sm = 1;
    StepOut out;               // internal device outputs
    int32_t* d_layout_ids = nullptr;
    float* d_actions = nullptr; uint8_t* d_obs = nullptr;   // only for the host-buffer entry point
    float* h_actions = nullptr; void* h_stage = nullptr; size_t h_stage_bytes = 0;
    bool layouts_loaded = false, was_reset = false;
    int64_t launches = 0;
    int sm_count = 148;
    // optional per-kernel device timing (bench): events around the step and obs kernels
    bool profiling = false;
    std::vector<cudaEvent_t> ev_pool; size_t ev_used = 0;
    std::vector<int> ev_kind;            // 0 = step/reset kernel, 1 = obs kernel (one entry per event pair)
    double ms_acc[2] = {0, 0}; int64_t n_acc[2] = {0, 0};
};

template <typename T>
static int dev_upload(bp_shipice* h, const std::vector<T>& v, const T** out) {
    void* p = nullptr;
    CUDA_TRY(cudaMalloc(&p, v.size() * sizeof(T) + 16));
    CUDA_TRY(cudaMemcpy(p, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice));
    h->dev_allocs.push_back(p);
    *out = (const T*)p;
    return 0;
}

extern "C" {

int bp_version(void) { return 100; }
const char* bp_last_error(void) { return g_err.c_str(); }
int bp_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

void bp_shipice_default_config(bp_shipice_config* c) {
    memset(c, 0, sizeof(*c));
    c->n_envs = 1; c->max_floes = 653; c->device = 0; c->pair_capacity = 0; c->live_capacity = 0; c->auto_reset = 0;
    c->dt = 0.8; c->steps = 400; c->iterations = 10; c->damping = 0.0; c->obstacle_density = 0.001;
    c->target_speed = 0.3; c->goal_y = 9.0; c->map_width = 12.0; c->map_height = 40.0; c->m_to_pix_scale = 25;
    c->settle_steps = 1000;
    static const double sv[] = {1., -0., 0.9, 0.10, 0.5, 0.25, 0.25, 0.25, 0, 0.25, -0.25, 0.25, -0.5, 0.25, -0.75, 0.25, -1., 0.25,
                                -1., -0.25, -0.75, -0.25, -0.5, -0.25, -0.25, -0.25, 0, -0.25, 0.25, -0.25, 0.5, -0.25, 0.9, -0.10};
    c->n_ship_verts = 17;
    memcpy(c->ship_verts, sv, sizeof(sv));
    c->ship_head[0] = 1.0; c->ship_head[1] = 0.0; c->ship_tail[0] = -1.0; c->ship_tail[1] = 0.0;
}

int bp_shipice_create(const bp_shipice_config* cfg, bp_shipice_t** out_h) {
    if (!cfg || !out_h) return fail(BP_ERR_INVALID, "null argument");
    if (cfg->n_envs <= 0 || cfg->max_floes <= 0 || cfg->max_floes > 60000) return fail(BP_ERR_INVALID, "n_envs / max_floes out of range");
    if (cfg->n_ship_verts < 3 || cfg->n_ship_verts > 32) return fail(BP_ERR_INVALID, "n_ship_verts out of range");
    if (cfg->steps <= 0 || cfg->iterations <= 0 || cfg->dt <= 0) return fail(BP_ERR_INVALID, "steps / iterations / dt must be positive");
    int ndev = 0;
    CUDA_TRY(cudaGetDeviceCount(&ndev));
    if (cfg->device < 0 || cfg->device >= ndev) return fail(BP_ERR_CUDA, "CUDA device not available");
    CUDA_TRY(cudaSetDevice(cfg->device));
    bp_shipice* h = new bp_shipice();
    h->cfg = *cfg;
    cudaDeviceProp prop;
    CUDA_TRY(cudaGetDeviceProperties(&prop, cfg->device));
    h->sm_count = prop.multiProcessorCount;
    Dims& d = h->d;
    d.NB = cfg->max_floes + 1;
    d.PMAX = cfg->pair_capacity > 0 ? cfg->pair_capacity : 8 * d.NB;
    d.LMAX = cfg->live_capacity > 0 ? cfg->live_capacity : (2 * d.NB > 256 ? 2 * d.NB : 256);
    d.n_envs = cfg->n_envs; d.n_layouts = 0;
    ShipIceConfig sc;
    sc.dt = cfg->dt; sc.steps = cfg->steps; sc.iterations = cfg->iterations; sc.damping = cfg->damping;
    sc.target_speed = cfg->target_speed; sc.goal_y = cfg->goal_y; sc.map_width = cfg->map_width; sc.map_height = cfg->map_height;
    sc.m_to_pix_scale = cfg->m_to_pix_scale; sc.settle_steps = cfg->settle_steps;
    h->P = make_params(sc);
    ObsParams& O = h->O;
    double grid = 1.0 / cfg->m_to_pix_scale;                       // ship_ice_env.py:92
    O.grid_w = (int)(cfg->map_width / grid); O.grid_h = (int)(cfg->map_height / grid);   // occupancy_map.py:20-21
    O.scale = (double)cfg->m_to_pix_scale; O.vshift = 2.0; O.goal_y = cfg->goal_y;
    O.g2m = cfg->map_height / O.grid_h; O.range_x = 12.0; O.range_y = 12.0;
    O.n_robot_raw = cfg->n_ship_verts;
    memcpy(O.robot_raw, cfg->ship_verts, sizeof(double) * 2 * cfg->n_ship_verts);
    O.head[0] = cfg->ship_head[0]; O.head[1] = cfg->ship_head[1]; O.tail[0] = cfg->ship_tail[0]; O.tail[1] = cfg->ship_tail[1];
    h->env_bytes = env_mem_size(d);
    size_t total = h->env_bytes * (size_t)d.n_envs;
    cudaError_t e = cudaMalloc((void**)&h->mem, total);
    if (e != cudaSuccess) { delete h; return fail(BP_ERR_CUDA, std::string("cudaMalloc of env state (") + std::to_string(total >> 20) + " MiB): " + cudaGetErrorString(e)); }
    CUDA_TRY(cudaMemset(h->mem, 0, total));
    CUDA_TRY(cudaMalloc((void**)&h->counter, 64));
    size_t E = d.n_envs;
    auto A = [&](void** p, size_t bytes) -> cudaError_t { cudaError_t r = cudaMalloc(p, bytes); if (r == cudaSuccess) { h->dev_allocs.push_back(*p); cudaMemset(*p, 0, bytes); } return r; };
    CUDA_TRY(A((void**)&h->out.reward, E * 8)); CUDA_TRY(A((void**)&h->out.terminated, E)); CUDA_TRY(A((void**)&h->out.truncated, E));
    CUDA_TRY(A((void**)&h->out.state, E * 24)); CUDA_TRY(A((void**)&h->out.pose, E * 24)); CUDA_TRY(A((void**)&h->out.total_work, E * 8));
    CUDA_TRY(A((void**)&h->out.collision_reward, E * 8)); CUDA_TRY(A((void**)&h->out.scaled_collision_reward, E * 8));
    CUDA_TRY(A((void**)&h->out.dist_reward, E * 8)); CUDA_TRY(A((void**)&h->out.trial_success, E)); CUDA_TRY(A((void**)&h->out.status, E * 4));
    CUDA_TRY(A((void**)&h->d_layout_ids, E * 4));
    *out_h = h;
    return BP_OK;
}

int bp_shipice_destroy(bp_shipice_t* h) {
    if (!h) return BP_OK;
    cudaSetDevice(h->cfg.device);
    for (void* p : h->dev_allocs) cudaFree(p);
    for (cudaEvent_t e : h->ev_pool) cudaEventDestroy(e);
    if (h->mem) cudaFree(h->mem);
    if (h->counter) cudaFree(h->counter);
    if (h->perm) cudaFree(h->perm);
    if (h->sched_keys) cudaFree(h->sched_keys);
    if (h->sched_prev) cudaFree(h->sched_prev);
    if (h->d_actions) cudaFree(h->d_actions);
    if (h->d_obs) cudaFree(h->d_obs);
    if (h->h_actions) cudaFreeHost(h->h_actions);
    if (h->h_stage) cudaFreeHost(h->h_stage);
    delete h;
    return BP_OK;
}

static int finish_load_layouts(bp_shipice* h, int32_t n_layouts) {
    HostLayouts& H = h->H;
    int rc;
    if ((rc = dev_upload(h, H.lv, &h->T.lv))) return rc;
    if ((rc = dev_upload(h, H.ln, &h->T.ln))) return rc;
    if ((rc = dev_upload(h, H.nv, &h->T.nv))) return rc;
    if ((rc = dev_upload(h, H.rad, &h->T.rad))) return rc;
    if ((rc = dev_upload(h, H.rbound, &h->T.rbound))) return rc;
    if ((rc = dev_upload(h, H.minv, &h->T.minv))) return rc;
    if ((rc = dev_upload(h, H.iinv, &h->T.iinv))) return rc;
    if ((rc = dev_upload(h, H.cog, &h->T.cog))) return rc;
    if ((rc = dev_upload(h, H.mat_e, &h->T.mat_e))) return rc;
    if ((rc = dev_upload(h, H.mat_u, &h->T.mat_u))) return rc;
    if ((rc = dev_upload(h, H.init, &h->T.init))) return rc;
    if ((rc = dev_upload(h, H.nbody, &h->T.nbody))) return rc;
    if ((rc = dev_upload(h, H.robot0, &h->T.robot0))) return rc;
    const double* p; const int* q;
    if ((rc = dev_upload(h, H.rawv, &p))) return rc; h->d_rawv = (double*)p;
    if ((rc = dev_upload(h, H.rawcen, &p))) return rc; h->d_rawcen = (double*)p;
    if ((rc = dev_upload(h, H.rawn, &q))) return rc; h->d_rawn = (int*)q;
    h->d.n_layouts = n_layouts;
    DevConst hc;
    hc.d = h->d; hc.P = h->P; hc.off = env_offsets(h->d); hc.T = h->T;
    hc.snap = nullptr; hc.env_bytes = h->env_bytes; hc.snap_pose = nullptr; hc.snap_status = nullptr; hc.start_pose = nullptr;
    CUDA_TRY(cudaMalloc((void**)&h->d_const, sizeof(DevConst)));
    h->dev_allocs.push_back(h->d_const);
    CUDA_TRY(cudaMemcpy(h->d_const, &hc, sizeof(DevConst), cudaMemcpyHostToDevice));
    // reset snapshots: simulate the reset of every layout once (init + settle), keep the env blocks
    if (!getenv("BP_NO_SNAPSHOT")) {
        size_t Ls = (size_t)n_layouts;
        CUDA_TRY(cudaMalloc((void**)&h->snap, h->env_bytes * Ls));
        h->dev_allocs.push_back(h->snap);
        CUDA_TRY(cudaMemset(h->snap, 0, h->env_bytes * Ls));
        StepOut so; void* p = nullptr;
        auto T = [&](size_t bytes) -> void* { p = nullptr; if (cudaMalloc(&p, bytes) == cudaSuccess) { cudaMemset(p, 0, bytes); h->dev_allocs.push_back(p); } return p; };
        so.reward = (double*)T(Ls * 8); so.terminated = (unsigned char*)T(Ls); so.truncated = (unsigned char*)T(Ls);
        so.state = (double*)T(Ls * 24); so.pose = (double*)T(Ls * 24); so.total_work = (double*)T(Ls * 8);
        so.collision_reward = (double*)T(Ls * 8); so.scaled_collision_reward = (double*)T(Ls * 8); so.dist_reward = (double*)T(Ls * 8);
        so.trial_success = (unsigned char*)T(Ls); so.status = (int*)T(Ls * 4);
        int32_t* ids = (int32_t*)T(Ls * 4);
        if (!so.reward || !so.terminated || !so.truncated || !so.state || !so.pose || !so.total_work || !so.collision_reward ||
            !so.scaled_collision_reward || !so.dist_reward || !so.trial_success || !so.status || !ids)
            return fail(BP_ERR_CUDA, "cudaMalloc of the reset snapshots failed");
        std::vector<int32_t> hid(Ls);
        for (size_t i = 0; i < Ls; i++) hid[i] = (int32_t)i;
        CUDA_TRY(cudaMemcpy(ids, hid.data(), Ls * 4, cudaMemcpyHostToDevice));
        CUDA_TRY(cudaMemset(h->counter, 0, 4));
        StepArgs a;
        a.mem = h->snap; a.env_bytes = h->env_bytes; a.d = h->d; a.d.n_envs = n_layouts; a.P = h->P; a.T = h->T; a.out = so;
        a.actions = nullptr; a.layout_ids = ids; a.mask = nullptr; a.counter = h->counter; a.auto_reset = 0; a.C = h->d_const; a.use_snap = 0; a.perm = nullptr; a.n_sched = 0;
        int blocks = (n_layouts + 3) / 4; if (blocks > h->sm_count * BP_MIN_BLOCKS) blocks = h->sm_count * BP_MIN_BLOCKS;
        k_shipice_reset<<<blocks, 128>>>(a);
        CUDA_TRY(cudaGetLastError());
        CUDA_TRY(cudaDeviceSynchronize());
        hc.snap = h->snap; hc.snap_pose = so.pose; hc.snap_status = so.status;
        CUDA_TRY(cudaMemcpy(h->d_const, &hc, sizeof(DevConst), cudaMemcpyHostToDevice));
    }
    if (const char* e = getenv("BP_STEP_IMPL")) h->warp_impl = (strcmp(e, "warp") == 0);
    if (const char* e = getenv("BP_NO_SCHEDULE")) h->use_schedule = !(atoi(e) != 0);
    if (const char* e = getenv("BP_NO_PAD")) h->pad_batches = !(atoi(e) != 0);
    if (const char* e = getenv("BP_NO_SNAKE")) h->snake = !(atoi(e) != 0);
    if (const char* e = getenv("BP_SCHED_HISTORY")) { h->sched_history = atoi(e) != 0; h->sched_hist_all = atoi(e) == 2; }
    // tuning experiments only: DRAM -> L2 fetch granularity hint (32 / 64 / 128 bytes); the env state is accessed sector by sector
    if (const char* e = getenv("BP_L2_FETCH")) { int g = atoi(e); if (g == 32 || g == 64 || g == 128) cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, (size_t)g); }
    // resident step (sim_res.cuh), ship-ice only (the per-substep checks of the maze run on the env block).  Selected with
    // BP_STEP_IMPL=res: in fp64 only three pools fit an SM, and three warps per SM issue too slowly to beat the batched
    // kernel (profiles/r02a_*), so the batched kernel stays the default.
    h->res_impl = false;
    if (const char* e = getenv("BP_STEP_IMPL")) if (strcmp(e, "res") == 0 && h->kind == KIND_SHIPICE) h->res_impl = true;
    if (h->res_impl) {
        cudaDeviceProp prop;
        CUDA_TRY(cudaGetDeviceProperties(&prop, h->cfg.device));
        size_t need = sizeof(res::ResShared);
        if (need > prop.sharedMemPerBlockOptin) return fail(BP_ERR_CUDA, "resident step needs more shared memory per block than the device offers");
        CUDA_TRY(cudaFuncSetAttribute(k_shipice_step_res, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)need));
        int per_sm = (int)(prop.sharedMemPerMultiprocessor / (need + prop.reservedSharedMemPerBlock + sizeof(EnvOff) + 64));
        if (per_sm < 1) per_sm = 1;
        if (const char* e = getenv("BP_RES_BLOCKS_PER_SM")) { int b = atoi(e); if (b > 0 && b < per_sm) per_sm = b; }   // tuning experiments only
        h->res_blocks_per_sm = per_sm;
    }
    h->hc_host = hc;
    h->layouts_loaded = true;
    return BP_OK;
}

int bp_shipice_load_layouts(bp_shipice_t* h, int32_t n_layouts, const int32_t* n_floes, const int32_t* n_verts,
                            const double* verts, const double* centres, const double* ship_state) {
    if (!h || n_layouts <= 0 || !n_floes || !n_verts || !verts || !centres || !ship_state) return fail(BP_ERR_INVALID, "null / empty layout arguments");
    if (h->layouts_loaded) return fail(BP_ERR_STATE, "layouts already loaded for this handle");
    CUDA_TRY(cudaSetDevice(h->cfg.device));
    ShipIceMaterial mat;
    mat.density = h->cfg.obstacle_density;
    std::string err = build_shipice_layouts(h->H, n_layouts, h->cfg.max_floes, n_floes, n_verts, verts, centres, ship_state,
                                            h->cfg.ship_verts, h->cfg.n_ship_verts, mat);
    if (!err.empty()) return fail(BP_ERR_INVALID, err);
    return finish_load_layouts(h, n_layouts);
}

static int copy_outputs(bp_shipice* h, const bp_shipice_outputs* o, cudaStream_t st, cudaMemcpyKind kind) {
    size_t E = h->d.n_envs;
    if (!o) return BP_OK;
#define CP(dst, src, bytes) if (dst) CUDA_TRY(cudaMemcpyAsync(dst, src, bytes, kind, st));
    CP(o->reward, h->out.reward, E * 8) CP(o->terminated, h->out.terminated, E) CP(o->truncated, h->out.truncated, E)
    CP(o->state, h->out.state, E * 24) CP(o->pose, h->out.pose, E * 24) CP(o->total_work, h->out.total_work, E * 8)
    CP(o->collision_reward, h->out.collision_reward, E * 8) CP(o->scaled_collision_reward, h->out.scaled_collision_reward, E * 8)
    CP(o->dist_reward, h->out.dist_reward, E * 8) CP(o->trial_success, h->out.trial_success, E) CP(o->status, h->out.status, E * 4)
#undef CP
    return BP_OK;
}

static void prof_begin(bp_shipice* h, int kind, cudaStream_t st) {
    if (!h->profiling) return;
    if (h->ev_used + 2 > h->ev_pool.size()) {
        for (int i = 0; i < 2; i++) { cudaEvent_t e; cudaEventCreate(&e); h->ev_pool.push_back(e); }
    }
    cudaEventRecord(h->ev_pool[h->ev_used], st);
    h->ev_kind.push_back(kind);
}
static void prof_end(bp_shipice* h, cudaStream_t st) {
    if (!h->profiling) return;
    cudaEventRecord(h->ev_pool[h->ev_used + 1], st);
    h->ev_used += 2;
}

static int launch_obs(bp_shipice* h, uint8_t* obs, const uint8_t* mask, cudaStream_t st) {
    if (!obs) return BP_OK;
    if (h->kind == KIND_MAZE) {
        MazeObsArgs ma;
        ma.mem = h->mem; ma.env_bytes = h->env_bytes; ma.d = h->d; ma.T = h->T; ma.O = h->MO;
        ma.rawv = h->d_rawv; ma.rawn = h->d_rawn; ma.mask = mask; ma.obs = obs;
        prof_begin(h, 1, st);
        k_maze_obs<<<h->d.n_envs, 256, sizeof(MazeObsSmem), st>>>(ma);
        prof_end(h, st);
        h->launches++;
        CUDA_TRY(cudaGetLastError());
        return BP_OK;
    }
    ObsArgs oa;
    oa.mem = h->mem; oa.env_bytes = h->env_bytes; oa.d = h->d; oa.T = h->T; oa.O = h->O;
    oa.rawv = h->d_rawv; oa.rawn = h->d_rawn; oa.rawcen = h->d_rawcen; oa.mask = mask; oa.obs = obs;
    prof_begin(h, 1, st);
    k_shipice_obs<<<h->d.n_envs, 256, 0, st>>>(oa);
    prof_end(h, st);
    h->launches++;
    CUDA_TRY(cudaGetLastError());
    return BP_OK;
}


static int launch_low_dim(bp_shipice* h, double* out, int32_t* counts, int maze, cudaStream_t st) {
    LowDimArgs la;
    la.mem = h->mem; la.env_bytes = h->env_bytes; la.d = h->d; la.T = h->T; la.out = out; la.counts = counts; la.maze = maze;
    la.stride = maze ? 2 * (h->max_boxes + 1) : 2 * h->cfg.max_floes;
    k_low_dim<<<(h->d.n_envs + 3) / 4, 128, 0, st>>>(la);
    h->launches++;
    CUDA_TRY(cudaGetLastError());
    return BP_OK;
}

static int sim_grid(bp_shipice* h) {
    // 4 warps per block, 32 / BP_NL envs per warp; enough resident blocks to fill the SMs, never more groups than envs
    int blocks = h->sm_count * BP_MIN_BLOCKS;
    if (const char* e = getenv("BP_BLOCKS_PER_SM")) { int b = atoi(e); if (b > 0 && b < BP_MIN_BLOCKS) blocks = h->sm_count * b; }   // tuning experiments only
    int per_block = 4 * BP_GROUPS_PER_WARP;
    int need = (h->d.n_envs + per_block - 1) / per_block;
    return blocks < need ? blocks : need;
}

// envs per warp for this handle: the small instantiation when the batch cannot fill the resident warps of the GPU with 8 per warp
static bool use_small_batches(bp_shipice* h) {
    if (const char* e = getenv("BP_FORCE_BK")) return atoi(e) == bk_small::BK;       // tuning experiments only
    return (long)h->d.n_envs * 2 <= (long)h->sm_count * BP_MIN_BLOCKS * 4 * bk_main::BK;
}
static int batch_grid(bp_shipice* h, int sched_batches = 0) {
    // 4 warps per block, BK envs per warp batch; never more warps than batches
    int blocks = h->sm_count * BP_MIN_BLOCKS;
    if (const char* e = getenv("BP_BLOCKS_PER_SM")) { int b = atoi(e); if (b > 0 && b < BP_MIN_BLOCKS) blocks = h->sm_count * b; }   // tuning experiments only
    int per_block = 4 * (use_small_batches(h) ? bk_small::BK : bk_main::BK);
    int need = sched_batches ? (sched_batches + 3) / 4 : (h->d.n_envs + per_block - 1) / per_block;
    return blocks < need ? blocks : need;
}

int bp_shipice_reset(bp_shipice_t* h, const int32_t* layout_ids, const uint8_t* mask, const bp_shipice_outputs* out, void* stream) {
    if (!h || !layout_ids) return fail(BP_ERR_INVALID, "null argument");
    if (!h->layouts_loaded) return fail(BP_ERR_STATE, "load layouts before reset");
    CUDA_TRY(cudaSetDevice(h->cfg.device));
    cudaStream_t st = (cudaStream_t)stream;
    CUDA_TRY(cudaMemsetAsync(h->counter, 0, 4, st));
    StepArgs a;
    a.mem = h->mem; a.env_bytes = h->env_bytes; a.d = h->d; a.P = h->P; a.T = h->T; a.out = h->out;
    a.actions = nullptr; a.layout_ids = layout_ids; a.mask = mask; a.counter = h->counter; a.auto_reset = 0; a.C = h->d_const; a.use_snap = h->snap ? 1 : 0;
    a.perm = nullptr; a.n_sched = 0;
    k_shipice_reset<<<sim_grid(h), 128, 0, st>>>(a);
    h->launches++;
    CUDA_TRY(cudaGetLastError());
    int rc = launch_obs(h, out ? out->obs : nullptr, mask, st);
    if (rc) return rc;
    h->was_reset = true;
    return copy_outputs(h, out, st, cudaMemcpyDeviceToDevice);
}

int bp_shipice_step_masked(bp_shipice_t* h, const float* actions, const uint8_t* mask, const bp_shipice_outputs* out, void* stream) {
    if (!h || !actions) return fail(BP_ERR_INVALID, "null argument");
    if (!h->was_reset) return fail(BP_ERR_STATE, "reset before step");
    CUDA_TRY(cudaSetDevice(h->cfg.device));
    cudaStream_t st = (cudaStream_t)stream;
    CUDA_TRY(cudaMemsetAsync(h->counter, 0, 4, st));
    StepArgs a;
    a.mem = h->mem; a.env_bytes = h->env_bytes; a.d = h->d; a.P = h->P; a.T = h->T; a.out = h->out;
    a.actions = actions; a.layout_ids = nullptr; a.mask = mask; a.counter = h->counter; a.auto_reset = h->cfg.auto_reset; a.C = h->d_const; a.use_snap = h->snap ? 1 : 0;
    a.perm = nullptr; a.n_sched = 0;
    prof_begin(h, 0, st);
    int sched_batches = 0;
    if (h->use_schedule && !h->warp_impl) {
        const int bk = h->res_impl ? res::RK : (use_small_batches(h) ? bk_small::BK : bk_main::BK);
        int n_batches = (h->d.n_envs + bk - 1) / bk;
        const int warp_slots = h->res_impl ? h->sm_count * h->res_blocks_per_sm : h->sm_count * BP_MIN_BLOCKS * 4;
        // Fewer batches than resident warps (the per-GPU share of a strong-scaling run): spread the envs over ALL warp slots with
        // partly filled batches (perm entries -1) instead of leaving slots idle — the step is latency bound, extra warps are free.
        if (h->pad_batches && n_batches < warp_slots) n_batches = warp_slots < h->d.n_envs ? warp_slots : h->d.n_envs;
        sched_batches = n_batches;
        if (!h->sched_keys) {
            CUDA_TRY(cudaMalloc((void**)&h->sched_keys, (size_t)h->d.n_envs * 4));
            CUDA_TRY(cudaMalloc((void**)&h->sched_prev, (size_t)h->d.n_envs * 24));
            CUDA_TRY(cudaMemsetAsync(h->sched_prev, 0, (size_t)h->d.n_envs * 24, st));
        }
        if (!h->perm || h->perm_cap < n_batches * bk) {
            if (h->perm) cudaFree(h->perm);
            CUDA_TRY(cudaMalloc((void**)&h->perm, (size_t)n_batches * bk * 4));
            h->perm_cap = n_batches * bk;
        }
        k_schedule<<<1, 1024, 0, st>>>(h->mem, h->env_bytes, h->d.n_envs, mask, h->cfg.auto_reset, h->perm, bk, n_batches, n_batches < 2 * warp_slots ? (h->snake ? 2 : 1) : 0, h->sched_history ? h->sched_prev : nullptr, h->sched_keys, h->sched_hist_all ? 1 : 0);
        a.perm = h->perm; a.n_sched = n_batches * bk;
    }
    if (h->res_impl) {
        int blocks = h->sm_count * h->res_blocks_per_sm;
        int need = sched_batches ? sched_batches : (h->d.n_envs + res::RK - 1) / res::RK;
        if (blocks > need) blocks = need;
        k_shipice_step_res<<<blocks, 32, sizeof(res::ResShared), st>>>(a);
    }
    else if (h->warp_impl) k_shipice_step_warp<<<sim_grid(h), 128, 0, st>>>(a);
    else if (use_small_batches(h)) k_shipice_step_small<<<batch_grid(h, sched_batches), 128, 0, st>>>(a);
    else k_shipice_step<<<batch_grid(h, sched_batches), 128, 0, st>>>(a);
    prof_end(h, st);
    h->launches++;
    CUDA_TRY(cudaGetLastError());
    int rc = launch_obs(h, out ? out->obs : nullptr, mask, st);
    if (rc) return rc;
    return copy_outputs(h, out, st, cudaMemcpyDeviceToDevice);
}

int bp_shipice_step(bp_shipice_t* h, const float* actions, const bp_shipice_outputs* out, void* stream) {
    return bp_shipice_step_masked(h, actions, nullptr, out, stream);
}

static int step_host_impl(bp_shipice_t* h, const float* actions, const bp_shipice_outputs* oh, uint8_t* obs_dev) {
    if (!h || !actions || !oh) return fail(BP_ERR_INVALID, "null argument");
    CUDA_TRY(cudaSetDevice(h->cfg.device));
    size_t E = h->d.n_envs;
    if (!h->d_actions) CUDA_TRY(cudaMalloc((void**)&h->d_actions, E * 4));
    const size_t obs_bytes = h->kind == KIND_MAZE ? (size_t)MZ_OBS_C * MZ_PIX : (size_t)BP_OBS_C * OBS_PIX;
    if (oh->obs && !obs_dev && !h->d_obs) CUDA_TRY(cudaMalloc((void**)&h->d_obs, E * obs_bytes));
    CUDA_TRY(cudaMemcpyAsync(h->d_actions, actions, E * 4, cudaMemcpyHostToDevice, 0));
    bp_shipice_outputs dev;
    memset(&dev, 0, sizeof(dev));
    dev.obs = obs_dev ? obs_dev : (oh->obs ? h->d_obs : nullptr);
    int rc = bp_shipice_step(h, h->d_actions, &dev, nullptr);
    if (rc) return rc;
    if (oh->obs) CUDA_TRY(cudaMemcpyAsync(oh->obs, dev.obs, E * obs_bytes, cudaMemcpyDeviceToHost, 0));
    rc = copy_outputs(h, oh, 0, cudaMemcpyDeviceToHost);
    if (rc) return rc;
    CUDA_TRY(cudaStreamSynchronize(0));
    return BP_OK;
}
int bp_shipice_step_host(bp_shipice_t* h, const float* actions, const bp_shipice_outputs* oh) { return step_host_impl(h, actions, oh, nullptr); }
int bp_shipice_step_host_obs_device(bp_shipice_t* h, const float* actions, const bp_shipice_outputs* oh, uint8_t* obs_dev) {
    if (!obs_dev) return fail(BP_ERR_INVALID, "null observation buffer");
    return step_host_impl(h, actions, oh, obs_dev);
}

size_t bp_shipice_state_bytes(bp_shipice_t* h) { return h ? h->env_bytes : 0; }
int bp_shipice_get_state(bp_shipice_t* h, int32_t env, void* out_host) {
    if (!h || !out_host || env < 0 || env >= h->d.n_envs) return fail(BP_ERR_INVALID, "bad argument");
    CUDA_TRY(cudaSetDevice(h->cfg.device));
    CUDA_TRY(cudaDeviceSynchronize());
    CUDA_TRY(cudaMemcpy(out_host, h->mem + h->env_bytes * (size_t)env, h->env_bytes, cudaMemcpyDeviceToHost));
    return BP_OK;
}
int bp_shipice_set_state(bp_shipice_t* h, int32_t env, const void* in_host) {
    if (!h || !in_host || env < 0 || env >= h->d.n_envs) return fail(BP_ERR_INVALID, "bad argument");
    if (!h->was_reset) return fail(BP_ERR_STATE, "reset before set_state");
    CUDA_TRY(cudaSetDevice(h->cfg.device));
    CUDA_TRY(cudaDeviceSynchronize());
    CUDA_TRY(cudaMemcpy(h->mem + h->env_bytes * (size_t)env, in_host, h->env_bytes, cudaMemcpyHostToDevice));
    return BP_OK;
}
int bp_shipice_set_bodies(bp_shipice_t* h, int32_t env, const double* bodies6, int32_t n_bodies) {
    if (!h || !bodies6 || env < 0 || env >= h->d.n_envs) return fail(BP_ERR_INVALID, "bad argument");
    if (!h->was_reset) return fail(BP_ERR_STATE, "reset before set_bodies");
    if (n_bodies <= 0 || n_bodies > h->d.NB) return fail(BP_ERR_INVALID, "body count out of range");
    CUDA_TRY(cudaSetDevice(h->cfg.device));
    double* dbuf = nullptr;
    CUDA_TRY(cudaMalloc((void**)&dbuf, (size_t)h->d.NB * 6 * 8));
    CUDA_TRY(cudaMemset(dbuf, 0, (size_t)h->d.NB * 6 * 8));
    CUDA_TRY(cudaMemcpy(dbuf, bodies6, (size_t)n_bodies * 6 * 8, cudaMemcpyHostToDevice));
    StepArgs a;
    a.mem = h->mem; a.env_bytes = h->env_bytes; a.d = h->d; a.P = h->P; a.T = h->T; a.out = h->out;
    a.actions = nullptr; a.layout_ids = nullptr; a.mask = nullptr; a.counter = h->counter; a.auto_reset = 0; a.C = h->d_const; a.use_snap = 0; a.perm = nullptr; a.n_sched = 0;
    k_shipice_set_bodies<<<1, 32>>>(a, env, dbuf);
    h->launches++;
    cudaError_t e = cudaDeviceSynchronize();
    cudaFree(dbuf);
    if (e != cudaSuccess) return fail(BP_ERR_CUDA, std::string("k_shipice_set_bodies: ") + cudaGetErrorString(e));
    return BP_OK;
}

static int fetch_env(bp_shipice* h, int env, std::vector<char>& buf) {
    if (env < 0 || env >= h->d.n_envs) return fail(BP_ERR_INVALID, "env index out of range");
    CUDA_TRY(cudaSetDevice(h->cfg.device));
    buf.resize(h->env_bytes);
    CUDA_TRY(cudaDeviceSynchronize());
    CUDA_TRY(cudaMemcpy(buf.data(), h->mem + h->env_bytes * (size_t)env, h->env_bytes, cudaMemcpyDeviceToHost));
    return BP_OK;
}
static LayoutView host_layout_view(bp_shipice* h, int l) {
    LayoutTab T;
    HostLayouts& H = h->H;
    T.lv = H.lv.data(); T.ln = H.ln.data(); T.nv = H.nv.data(); T.rad = H.rad.data(); T.rbound = H.rbound.data(); T.minv = H.minv.data(); T.iinv = H.iinv.data();
    T.cog = H.cog.data(); T.mat_e = H.mat_e.data(); T.mat_u = H.mat_u.data(); T.init = H.init.data(); T.nbody = H.nbody.data();
    LayoutView L;
    size_t o = (size_t)l * h->d.NB;
    L.lv = T.lv + o * VMAX * 2; L.ln = T.ln + o * VMAX * 2; L.nv = T.nv + o; L.rad = T.rad + o; L.rbound = T.rbound + o; L.minv = T.minv + o; L.iinv = T.iinv + o;
    L.cog = T.cog + o * 2; L.mat_e = T.mat_e + o; L.mat_u = T.mat_u + o; L.init = T.init + o * 3; L.nbody = T.nbody[l]; L.robot0 = H.robot0[l];
    return L;
}

int bp_shipice_set_start_poses(bp_shipice_t* h, const double* poses) {
    if (!h) return fail(BP_ERR_INVALID, "null argument");
    if (!h->layouts_loaded) return fail(BP_ERR_STATE, "load layouts first");
    CUDA_TRY(cudaSetDevice(h->cfg.device));
    h->hc_host.start_pose = poses;
    CUDA_TRY(cudaMemcpy(h->d_const, &h->hc_host, sizeof(DevConst), cudaMemcpyHostToDevice));
    return BP_OK;
}
int bp_maze_set_start_poses(bp_maze_t* h, const double* poses) { return bp_shipice_set_start_poses(h, poses); }

// ship_ice_env.py:96 (OccupancyGrid with 0.2 m cells when egocentric_obs is false) and occupancy_map.py:20-21, 105
static int global_obs_params(bp_shipice* h, GlobalObsParams* G) {
    const double cell = 0.2, scale = (double)h->cfg.m_to_pix_scale;
    G->raw_h = (int)(h->cfg.map_height * scale); G->raw_w = (int)(h->cfg.map_width * scale);
    G->blk_y = (int)(cell * scale); G->blk_x = (int)(cell * scale);
    G->gh = (int)(h->cfg.map_height / cell); G->gw = (int)(h->cfg.map_width / cell);
    G->cell_scale = G->gw / h->cfg.map_width;
    if (G->blk_y < 1 || G->gh < 1 || G->gw < 1) return fail(BP_ERR_INVALID, "global observation: m_to_pix_scale below 5 gives an empty block");
    if ((long long)G->raw_h * G->raw_w > (long long)OBS_RAW_WORDS * 32) return fail(BP_ERR_INVALID, "global observation: map_height * map_width * scale^2 exceeds the 307200-pixel mask");
    if (G->cell_scale != G->gh / h->cfg.map_height) return fail(BP_ERR_INVALID, "global observation: anisotropic cell scale is not supported");
    // block_reduce pads the image up to a multiple of the block; its output must be the cell grid the footprint is drawn on
    if ((G->raw_h + G->blk_y - 1) / G->blk_y != G->gh || (G->raw_w + G->blk_x - 1) / G->blk_x != G->gw)
        return fail(BP_ERR_INVALID, "global observation: concentration grid and footprint grid differ (the reference's np.concatenate would fail)");
    return BP_OK;
}

int bp_shipice_obs_global_shape(bp_shipice_t* h, int32_t* hw) {
    if (!h || !hw) return fail(BP_ERR_INVALID, "null argument");
    if (h->kind != KIND_SHIPICE) return fail(BP_ERR_INVALID, "not a ship-ice handle");
    GlobalObsParams G;
    int rc = global_obs_params(h, &G);
    if (rc) return rc;
    hw[0] = G.gh; hw[1] = G.gw;
    return BP_OK;
}

int bp_shipice_obs_global(bp_shipice_t* h, uint8_t* obs, const uint8_t* mask, void* stream) {
    if (!h || !obs) return fail(BP_ERR_INVALID, "null argument");
    if (h->kind != KIND_SHIPICE) return fail(BP_ERR_INVALID, "not a ship-ice handle");
    if (!h->was_reset) return fail(BP_ERR_STATE, "reset first");
    GlobalObsParams G;
    int rc = global_obs_params(h, &G);
    if (rc) return rc;
    CUDA_TRY(cudaSetDevice(h->cfg.device));
    cudaStream_t st = (cudaStream_t)stream;
    ObsArgs oa;
    oa.mem = h->mem; oa.env_bytes = h->env_bytes; oa.d = h->d; oa.T = h->T; oa.O = h->O;
    oa.rawv = h->d_rawv; oa.rawn = h->d_rawn; oa.rawcen = h->d_rawcen; oa.mask = mask; oa.obs = obs;
    prof_begin(h, 1, st);
    k_shipice_obs_global<<<h->d.n_envs, 256, 0, st>>>(oa, G);
    prof_end(h, st);
    h->launches++;
    CUDA_TRY(cudaGetLastError());
    return BP_OK;
}

int bp_shipice_low_dim(bp_shipice_t* h, double* out, int32_t* counts, void* stream) {
    if (!h || !out) return fail(BP_ERR_INVALID, "null argument");
    if (h->kind != KIND_SHIPICE) return fail(BP_ERR_INVALID, "not a ship-ice handle");
    if (!h->was_reset) return fail(BP_ERR_STATE, "reset first");
    CUDA_TRY(cudaSetDevice(h->cfg.device));
    return launch_low_dim(h, out, counts, 0, (cudaStream_t)stream);
}

int bp_shipice_get_obstacles(bp_shipice_t* h, int32_t env, double* verts, int32_t* n_verts) {
    if (!h || !verts || !n_verts) return fail(BP_ERR_INVALID, "null argument");
    if (!h->was_reset) return fail(BP_ERR_STATE, "reset first");
    std::vector<char> buf;
    int rc = fetch_env(h, env, buf);
    if (rc) return rc;
    EnvView ev = make_view(buf.data(), h->d);
    LayoutView L = host_layout_view(h, ev.sc->layout);
    int nf = 0;
    for (int b = 0; b < ev.sc->robot; b++) {
        if (L.minv[b] == 0.0) continue;              // static walls
        n_verts[nf] = L.nv[b];
        for (int i = 0; i < 2 * L.nv[b]; i++) verts[(size_t)nf * VMAX * 2 + i] = (double)ev.wv[(size_t)b * VMAX * 2 + i];
        nf++;
    }
    return nf;
}

int bp_shipice_get_bodies(bp_shipice_t* h, int32_t env, double* out) {
    if (!h || !out) return fail(BP_ERR_INVALID, "null argument");
    if (!h->was_reset) return fail(BP_ERR_STATE, "reset first");
    std::vector<char> buf;
    int rc = fetch_env(h, env, buf);
    if (rc) return rc;
    EnvView ev = make_view(buf.data(), h->d);
    LayoutView L = host_layout_view(h, ev.sc->layout);
    int nb = ev.sc->nbody;
    for (int b = 0; b < nb; b++) {
        double c = ev.body[b].rc, s = ev.body[b].rs, cx = L.cog[2 * b], cy = L.cog[2 * b + 1];
        out[6 * b] = ev.body[b].px - (cx * c - cy * s); out[6 * b + 1] = ev.body[b].py - (cx * s + cy * c); out[6 * b + 2] = ev.body[b].pa;
        out[6 * b + 3] = ev.body[b].vx; out[6 * b + 4] = ev.body[b].vy; out[6 * b + 5] = ev.body[b].vw;
    }
    return nb;
}

int bp_shipice_get_counters(bp_shipice_t* h, int32_t env, int64_t* o) {
    if (!h || !o) return fail(BP_ERR_INVALID, "null argument");
    std::vector<char> buf;
    int rc = fetch_env(h, env, buf);
    if (rc) return rc;
    EnvView ev = make_view(buf.data(), h->d);
    EnvScal* s = ev.sc;
    o[0] = s->stamp; o[1] = s->n_awake; o[2] = s->n_pairs; o[3] = s->n_free; o[4] = s->n_order; o[5] = s->n_colours;
    o[6] = s->status; o[7] = s->n_apairs; o[8] = s->n_cached; o[9] = s->stat_gjk; o[10] = s->stat_pairs_eval;
    o[11] = s->stat_live_sum; o[12] = s->stat_awake_sum; o[13] = s->stat_recolour; o[14] = s->stat_rebuild; o[15] = s->stat_substeps;
    o[16] = s->stat_epa; o[17] = s->stat_skipped;
    for (int k = 0; k < 8; k++) o[18 + k] = s->t_phase[k];
    return BP_OK;
}

int bp_shipice_set_profiling(bp_shipice_t* h, int32_t on) {
    if (!h) return fail(BP_ERR_INVALID, "null handle");
    h->profiling = on != 0;
    return BP_OK;
}

int bp_shipice_get_kernel_ms(bp_shipice_t* h, double* out4) {
    if (!h || !out4) return fail(BP_ERR_INVALID, "null argument");
    CUDA_TRY(cudaSetDevice(h->cfg.device));
    CUDA_TRY(cudaDeviceSynchronize());
    for (size_t i = 0; i * 2 < h->ev_used; i++) {
        float ms = 0.f;
        CUDA_TRY(cudaEventElapsedTime(&ms, h->ev_pool[2 * i], h->ev_pool[2 * i + 1]));
        int k = h->ev_kind[i];
        h->ms_acc[k] += ms; h->n_acc[k]++;
    }
    h->ev_used = 0; h->ev_kind.clear();
    out4[0] = h->ms_acc[0]; out4[1] = (double)h->n_acc[0]; out4[2] = h->ms_acc[1]; out4[3] = (double)h->n_acc[1];
    h->ms_acc[0] = h->ms_acc[1] = 0; h->n_acc[0] = h->n_acc[1] = 0;
    return BP_OK;
}

size_t bp_shipice_env_bytes(bp_shipice_t* h) { return h ? h->env_bytes : 0; }
int64_t bp_shipice_launch_count(bp_shipice_t* h) { return h ? h->launches : 0; }

// ------------------------------------------------------------------------------------------------ maze-NAMO
void bp_maze_default_config(bp_maze_config* c, int32_t maze_version) {
    memset(c, 0, sizeof(*c));
    c->n_envs = 1; c->max_boxes = 5; c->device = 0; c->auto_reset = 0;
    c->dt = 0.8; c->steps = 400; c->iterations = 10; c->damping = 0.0; c->obstacle_density = 0.001; c->obstacle_size = 0.5;
    c->target_speed = 0.15; c->goal_radius = 1.0; c->min_r = 0.86; c->m_to_pix_scale = 16; c->settle_steps = 1000;
    double W, Lg;
    if (maze_version == 2) {
        W = Lg = 20.0; c->goal_x = 3.33; c->goal_y = 3.33;
        c->start[0] = 16.66; c->start[1] = 16.66; c->start[2] = 3 * M_PI / 2;
        const double w[6][4] = {{0, 0, W, 0}, {0, 0, 0, Lg}, {W, 0, W, Lg}, {0, Lg, W, Lg}, {W / 3, 0, W / 3, 2 * Lg / 3}, {2 * W / 3, Lg, 2 * W / 3, Lg / 3}};
        memcpy(c->walls, w, sizeof(w));
    } else {
        W = Lg = 15.0; c->goal_x = 3.75; c->goal_y = 3.75;
        c->start[0] = 11.25; c->start[1] = 3.75; c->start[2] = M_PI / 2;
        const double w[6][4] = {{0, 0, W, 0}, {0, 0, 0, Lg}, {W, 0, W, Lg}, {0, Lg, W, Lg}, {2 * W / 2, Lg, 2 * W / 2, 5}, {W / 2, 0, W / 2, Lg - Lg / 3}};
        memcpy(c->walls, w, sizeof(w));
    }
    c->width = W; c->length = Lg; c->n_walls = 6;
    static const double rv[] = {0.7, -0.5, 0.55, -0.6, -0.55, -0.6, -0.7, -0.5, -0.7, 0.5, -0.55, 0.6, 0.55, 0.6, 0.7, 0.5};
    c->n_robot_verts = 8; memcpy(c->robot_verts, rv, sizeof(rv));
    static const double wv[4][8] = {{0.2, 0.6, 0.2, 0.7, 0.5, 0.7, 0.55, 0.6}, {0.2, -0.7, 0.2, -0.6, 0.55, -0.6, 0.55, -0.7},
                                    {-0.2, 0.6, -0.2, 0.7, -0.55, 0.7, -0.55, 0.6}, {-0.2, -0.6, -0.2, -0.7, -0.55, -0.7, -0.55, -0.6}};
    c->n_wheels = 4;
    for (int w = 0; w < 4; w++) { c->wheel_nv[w] = 4; memcpy(c->wheel_verts + 16 * w, wv[w], sizeof(wv[w])); }
}

int bp_maze_create(const bp_maze_config* cfg, bp_maze_t** out_h) {
    if (!cfg || !out_h) return fail(BP_ERR_INVALID, "null argument");
    if (cfg->n_envs <= 0 || cfg->max_boxes < 0 || cfg->max_boxes > 4096) return fail(BP_ERR_INVALID, "n_envs / max_boxes out of range");
    if (cfg->n_walls < 0 || cfg->n_walls > BP_MAZE_MAX_WALLS) return fail(BP_ERR_INVALID, "n_walls out of range");
    if (cfg->n_robot_verts < 3 || cfg->n_robot_verts > 32 || cfg->n_wheels < 0 || cfg->n_wheels > BP_MAZE_MAX_WHEELS) return fail(BP_ERR_INVALID, "robot shape out of range");
    for (int w = 0; w < cfg->n_wheels; w++) if (cfg->wheel_nv[w] < 3 || cfg->wheel_nv[w] > 8) return fail(BP_ERR_INVALID, "wheel vertex count out of range");
    if (cfg->steps <= 0 || cfg->iterations <= 0 || cfg->dt <= 0 || cfg->m_to_pix_scale <= 0) return fail(BP_ERR_INVALID, "steps / iterations / dt / scale must be positive");
    int ndev = 0;
    CUDA_TRY(cudaGetDeviceCount(&ndev));
    if (cfg->device < 0 || cfg->device >= ndev) return fail(BP_ERR_CUDA, "CUDA device not available");
    CUDA_TRY(cudaSetDevice(cfg->device));
    bp_shipice* h = new bp_shipice();
    h->kind = KIND_MAZE; h->mcfg = *cfg; h->max_boxes = cfg->max_boxes;
    memset(&h->cfg, 0, sizeof(h->cfg));
    h->cfg.n_envs = cfg->n_envs; h->cfg.device = cfg->device; h->cfg.auto_reset = cfg->auto_reset;
    cudaDeviceProp prop;
    CUDA_TRY(cudaGetDeviceProperties(&prop, cfg->device));
    h->sm_count = prop.multiProcessorCount;
    MazeConfig& m = h->mz;
    m.dt = cfg->dt; m.steps = cfg->steps; m.iterations = cfg->iterations; m.damping = cfg->damping; m.target_speed = cfg->target_speed;
    m.obstacle_density = cfg->obstacle_density; m.obstacle_size = cfg->obstacle_size; m.width = cfg->width; m.length = cfg->length;
    m.goal_x = cfg->goal_x; m.goal_y = cfg->goal_y; m.goal_radius = cfg->goal_radius; m.min_r = cfg->min_r;
    m.m_to_pix_scale = cfg->m_to_pix_scale; m.settle_steps = cfg->settle_steps;
    memcpy(m.start, cfg->start, sizeof(m.start));
    m.n_walls = cfg->n_walls; memcpy(m.walls, cfg->walls, sizeof(double) * 4 * cfg->n_walls);
    m.n_robot_verts = cfg->n_robot_verts; memcpy(m.robot_verts, cfg->robot_verts, sizeof(double) * 2 * cfg->n_robot_verts);
    m.n_wheels = cfg->n_wheels;
    for (int w = 0; w < cfg->n_wheels; w++) { m.wheel_nv[w] = cfg->wheel_nv[w]; memcpy(m.wheel_verts + 16 * w, cfg->wheel_verts + 16 * w, sizeof(double) * 16); }
    Dims& d = h->d;
    d.NB = cfg->n_walls + cfg->max_boxes + 1 + cfg->n_wheels;
    d.PMAX = d.NB * (d.NB - 1) / 2 + 8;
    d.LMAX = d.PMAX < 256 ? d.PMAX : 256; if (d.LMAX < 64) d.LMAX = 64;
    d.n_envs = cfg->n_envs; d.n_layouts = 0;
    h->P = make_maze_params(m);
    // the two state-independent maps
    int H = 0, W = 0;
    build_maze_maps(m, &H, &W, h->wall_map, h->dist_norm, h->dist_raw);
    if (H <= 0 || W <= 0) { delete h; return fail(BP_ERR_INVALID, "empty maze grid"); }
    auto A = [&](void** p, size_t bytes) -> cudaError_t { cudaError_t r = cudaMalloc(p, bytes); if (r == cudaSuccess) { h->dev_allocs.push_back(*p); cudaMemset(*p, 0, bytes); } return r; };
    uint8_t* d_wall = nullptr; double* d_dist = nullptr;
    CUDA_TRY(A((void**)&d_wall, (size_t)H * W)); CUDA_TRY(A((void**)&d_dist, (size_t)H * W * 8));
    CUDA_TRY(cudaMemcpy(d_wall, h->wall_map.data(), (size_t)H * W, cudaMemcpyHostToDevice));
    CUDA_TRY(cudaMemcpy(d_dist, h->dist_norm.data(), (size_t)H * W * 8, cudaMemcpyHostToDevice));
    h->P.dist_map = d_dist; h->P.dist_h = H; h->P.dist_w = W;
    MazeObsParams& O = h->MO;
    O.grid_h = H; O.grid_w = W; O.scale = (double)cfg->m_to_pix_scale; O.n_robot_raw = cfg->n_robot_verts;
    memcpy(O.robot_raw, cfg->robot_verts, sizeof(double) * 2 * cfg->n_robot_verts);
    O.wall_map = d_wall; O.dist_map = d_dist;
    CUDA_TRY(cudaFuncSetAttribute(k_maze_obs, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(MazeObsSmem)));
    h->env_bytes = env_mem_size(d);
    size_t total = h->env_bytes * (size_t)d.n_envs;
    cudaError_t e = cudaMalloc((void**)&h->mem, total);
    if (e != cudaSuccess) { delete h; return fail(BP_ERR_CUDA, std::string("cudaMalloc of env state (") + std::to_string(total >> 20) + " MiB): " + cudaGetErrorString(e)); }
    CUDA_TRY(cudaMemset(h->mem, 0, total));
    CUDA_TRY(cudaMalloc((void**)&h->counter, 64));
    size_t E = d.n_envs;
    CUDA_TRY(A((void**)&h->out.reward, E * 8)); CUDA_TRY(A((void**)&h->out.terminated, E)); CUDA_TRY(A((void**)&h->out.truncated, E));
    CUDA_TRY(A((void**)&h->out.state, E * 24)); CUDA_TRY(A((void**)&h->out.pose, E * 24)); CUDA_TRY(A((void**)&h->out.total_work, E * 8));
    CUDA_TRY(A((void**)&h->out.collision_reward, E * 8)); CUDA_TRY(A((void**)&h->out.scaled_collision_reward, E * 8));
    CUDA_TRY(A((void**)&h->out.dist_reward, E * 8)); CUDA_TRY(A((void**)&h->out.trial_success, E)); CUDA_TRY(A((void**)&h->out.status, E * 4));
    CUDA_TRY(A((void**)&h->d_layout_ids, E * 4));
    *out_h = h;
    return BP_OK;
}

int bp_maze_destroy(bp_maze_t* h) { return bp_shipice_destroy(h); }

int bp_maze_load_layouts(bp_maze_t* h, int32_t n_layouts, const int32_t* n_boxes, const double* centres) {
    if (!h || n_layouts <= 0 || !n_boxes || !centres) return fail(BP_ERR_INVALID, "null / empty layout arguments");
    if (h->kind != KIND_MAZE) return fail(BP_ERR_INVALID, "not a maze handle");
    if (h->layouts_loaded) return fail(BP_ERR_STATE, "layouts already loaded for this handle");
    CUDA_TRY(cudaSetDevice(h->cfg.device));
    std::string err = build_maze_layouts(h->H, n_layouts, h->max_boxes, n_boxes, centres, h->mz);
    if (!err.empty()) return fail(BP_ERR_INVALID, err);
    return finish_load_layouts(h, n_layouts);
}

int bp_maze_reset(bp_maze_t* h, const int32_t* layout_ids, const uint8_t* mask, const bp_maze_outputs* out, void* stream) {
    return bp_shipice_reset(h, layout_ids, mask, out, stream);
}
int bp_maze_step(bp_maze_t* h, const float* actions, const bp_maze_outputs* out, void* stream) {
    return bp_shipice_step_masked(h, actions, nullptr, out, stream);
}
int bp_maze_step_masked(bp_maze_t* h, const float* actions, const uint8_t* mask, const bp_maze_outputs* out, void* stream) {
    return bp_shipice_step_masked(h, actions, mask, out, stream);
}
int bp_maze_step_host(bp_maze_t* h, const float* actions, const bp_maze_outputs* out_host) { return bp_shipice_step_host(h, actions, out_host); }
int bp_maze_step_host_obs_device(bp_maze_t* h, const float* actions, const bp_maze_outputs* out_host, uint8_t* obs_dev) {
    return bp_shipice_step_host_obs_device(h, actions, out_host, obs_dev);
}
int bp_maze_low_dim(bp_maze_t* h, double* out, int32_t* counts, void* stream) {
    if (!h || !out) return fail(BP_ERR_INVALID, "null argument");
    if (h->kind != KIND_MAZE) return fail(BP_ERR_INVALID, "not a maze handle");
    if (!h->was_reset) return fail(BP_ERR_STATE, "reset first");
    CUDA_TRY(cudaSetDevice(h->cfg.device));
    return launch_low_dim(h, out, counts, 1, (cudaStream_t)stream);
}
int bp_maze_get_obstacles(bp_maze_t* h, int32_t env, double* verts, int32_t* n_verts) { return bp_shipice_get_obstacles(h, env, verts, n_verts); }
int bp_maze_get_bodies(bp_maze_t* h, int32_t env, double* out) { return bp_shipice_get_bodies(h, env, out); }
size_t bp_maze_state_bytes(bp_maze_t* h) { return bp_shipice_state_bytes(h); }
int bp_maze_get_state(bp_maze_t* h, int32_t env, void* out_host) { return bp_shipice_get_state(h, env, out_host); }
int bp_maze_set_state(bp_maze_t* h, int32_t env, const void* in_host) { return bp_shipice_set_state(h, env, in_host); }
int bp_maze_set_bodies(bp_maze_t* h, int32_t env, const double* bodies6, int32_t n_bodies) { return bp_shipice_set_bodies(h, env, bodies6, n_bodies); }
int bp_maze_get_maps(bp_maze_t* h, uint8_t* wall, double* dist_norm, double* dist_raw, int32_t* hw) {
    if (!h || h->kind != KIND_MAZE) return fail(BP_ERR_INVALID, "not a maze handle");
    size_t n = h->wall_map.size();
    if (wall) memcpy(wall, h->wall_map.data(), n);
    if (dist_norm) memcpy(dist_norm, h->dist_norm.data(), n * 8);
    if (dist_raw) memcpy(dist_raw, h->dist_raw.data(), n * 8);
    if (hw) { hw[0] = h->P.dist_h; hw[1] = h->P.dist_w; }
    return BP_OK;
}


// ------------------------------------------------------------------------------------------------ spfa
static int* g_spfa_scratch = nullptr; static size_t g_spfa_scratch_bytes = 0; static int* g_spfa_counter = nullptr;

int bp_spfa_batch(const uint8_t* maps, const int32_t* sources, int32_t batch, int32_t H, int32_t W, float* dists, int32_t* parents, void* stream) {
    if (!maps || !sources || !dists || !parents) return fail(BP_ERR_INVALID, "null argument");
    if (batch <= 0 || H <= 0 || W <= 0 || (long long)H * W > (1ll << 26)) return fail(BP_ERR_INVALID, "bad batch / grid size");
    cudaStream_t st = (cudaStream_t)stream;
    int dev = 0; CUDA_TRY(cudaGetDevice(&dev));
    int sms = 0; CUDA_TRY(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    int blocks = sms * 4; if (blocks > batch) blocks = batch;
    size_t need = (size_t)blocks * 3 * H * W * sizeof(int);
    if (need > g_spfa_scratch_bytes) {
        if (g_spfa_scratch) CUDA_TRY(cudaFree(g_spfa_scratch));
        g_spfa_scratch = nullptr; g_spfa_scratch_bytes = 0;
        CUDA_TRY(cudaMalloc((void**)&g_spfa_scratch, need));
        g_spfa_scratch_bytes = need;
    }
    if (!g_spfa_counter) CUDA_TRY(cudaMalloc((void**)&g_spfa_counter, 64));
    CUDA_TRY(cudaMemsetAsync(g_spfa_counter, 0, 4, st));
    k_spfa<<<blocks, 512, 0, st>>>(maps, sources, batch, H, W, dists, parents, g_spfa_scratch, g_spfa_counter);
    CUDA_TRY(cudaGetLastError());
    return BP_OK;
}

int bp_spfa_batch_host(const uint8_t* maps, const int32_t* sources, int32_t batch, int32_t H, int32_t W, float* dists, int32_t* parents) {
    if (!maps || !sources || !dists || !parents) return fail(BP_ERR_INVALID, "null argument");
    size_t nv = (size_t)H * W * batch;
    uint8_t* dm = nullptr; int32_t* ds = nullptr; float* dd = nullptr; int32_t* dp = nullptr;
    CUDA_TRY(cudaMalloc((void**)&dm, nv)); CUDA_TRY(cudaMalloc((void**)&ds, (size_t)batch * 8));
    CUDA_TRY(cudaMalloc((void**)&dd, nv * 4)); CUDA_TRY(cudaMalloc((void**)&dp, nv * 4));
    CUDA_TRY(cudaMemcpy(dm, maps, nv, cudaMemcpyHostToDevice));
    CUDA_TRY(cudaMemcpy(ds, sources, (size_t)batch * 8, cudaMemcpyHostToDevice));
    int rc = bp_spfa_batch(dm, ds, batch, H, W, dd, dp, nullptr);
    if (rc == BP_OK) {
        cudaError_t e1 = cudaMemcpy(dists, dd, nv * 4, cudaMemcpyDeviceToHost);
        cudaError_t e2 = cudaMemcpy(parents, dp, nv * 4, cudaMemcpyDeviceToHost);
        if (e1 != cudaSuccess || e2 != cudaSuccess) rc = fail(BP_ERR_CUDA, "spfa result copy failed");
    }
    cudaFree(dm); cudaFree(ds); cudaFree(dd); cudaFree(dp);
    return rc;
}

// Reference-order parents (spfa.cuh spfa_serial_run): one warp per map, lane 0 restates the reference solver's queue, so that
// parents[] equals spfa.spfa's own output also where several predecessors tie.  Distances are the same as bp_spfa_batch's.
// Slow by construction (a sequential algorithm per map); meant for parity runs and for callers that walk the parents.
int bp_spfa_batch_exact(const uint8_t* maps, const int32_t* sources, int32_t batch, int32_t H, int32_t W, float* dists, int32_t* parents, void* stream) {
    if (!maps || !sources || !dists || !parents) return fail(BP_ERR_INVALID, "null argument");
    if (batch <= 0 || H <= 0 || W <= 0 || (long long)H * W > (1ll << 26)) return fail(BP_ERR_INVALID, "bad batch / grid size");
    cudaStream_t st = (cudaStream_t)stream;
    const size_t nv = (size_t)H * W;
    int* ring = nullptr; unsigned char* inq = nullptr;
    CUDA_TRY(cudaMalloc((void**)&ring, (size_t)batch * (nv + 2) * sizeof(int)));
    if (cudaMalloc((void**)&inq, (size_t)batch * nv) != cudaSuccess) { cudaFree(ring); return fail(BP_ERR_CUDA, "spfa scratch allocation failed"); }
    k_spfa_serial<<<batch, 32, 0, st>>>(maps, sources, batch, H, W, dists, parents, ring, inq);
    cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);          // the scratch is released below
    cudaFree(ring); cudaFree(inq);
    if (e != cudaSuccess) return fail(BP_ERR_CUDA, std::string("bp_spfa_batch_exact: ") + cudaGetErrorString(e));
    return BP_OK;
}

int bp_spfa_batch_exact_host(const uint8_t* maps, const int32_t* sources, int32_t batch, int32_t H, int32_t W, float* dists, int32_t* parents) {
    if (!maps || !sources || !dists || !parents) return fail(BP_ERR_INVALID, "null argument");
    size_t nv = (size_t)H * W * batch;
    uint8_t* dm = nullptr; int32_t* ds = nullptr; float* dd = nullptr; int32_t* dp = nullptr;
    CUDA_TRY(cudaMalloc((void**)&dm, nv)); CUDA_TRY(cudaMalloc((void**)&ds, (size_t)batch * 8));
    CUDA_TRY(cudaMalloc((void**)&dd, nv * 4)); CUDA_TRY(cudaMalloc((void**)&dp, nv * 4));
    CUDA_TRY(cudaMemcpy(dm, maps, nv, cudaMemcpyHostToDevice));
    CUDA_TRY(cudaMemcpy(ds, sources, (size_t)batch * 8, cudaMemcpyHostToDevice));
    int rc = bp_spfa_batch_exact(dm, ds, batch, H, W, dd, dp, nullptr);
    if (rc == BP_OK) {
        cudaError_t e1 = cudaMemcpy(dists, dd, nv * 4, cudaMemcpyDeviceToHost);
        cudaError_t e2 = cudaMemcpy(parents, dp, nv * 4, cudaMemcpyDeviceToHost);
        if (e1 != cudaSuccess || e2 != cudaSuccess) rc = fail(BP_ERR_CUDA, "spfa result copy failed");
    }
    cudaFree(dm); cudaFree(ds); cudaFree(dd); cudaFree(dp);
    return rc;
}

}  // extern "C"

// box-delivery-v0 / area-clearing-v0 keep fp64 state (their own small-world core, boxdel_*.cuh): not part of the fp32 build
#ifndef BP_F32
#include "boxdel_api.cuh"
#endif
