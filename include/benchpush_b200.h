/* benchpush_b200 — C ABI of the B200-native batched simulator for BenchPush's 2D pushing envs.
 *
 * This is the drop-in boundary for the hot path.  The reference has one native FFI
 * on this path (spfa, pybind11) and otherwise reaches native code through pymunk's cffi binding of
 * Chipmunk2D; each entry point below names the reference interface it replaces (paths relative to
 * the reference tree).  Plain pointers and sizes only; no torch types.  Unless a function says
 * "host", pointers are DEVICE pointers owned by the caller (e.g. torch tensors) and the work is
 * enqueued on `stream` (a cudaStream_t passed as void*; NULL = default stream) without
 * synchronising the host.  Every function returns 0 on success or a negative bp_status code;
 * bp_last_error() returns a message for the calling thread's last failure.
 */
#ifndef BENCHPUSH_B200_H
#define BENCHPUSH_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define BP_VMAX 20            /* max vertices per convex obstacle */
#define BP_OBS_C 4            /* ship-ice egocentric observation: 4 x 150 x 150 u8 */
#define BP_OBS_H 150
#define BP_OBS_W 150

enum bp_status {
    BP_OK = 0,
    BP_ERR_INVALID = -1,      /* bad argument */
    BP_ERR_CUDA = -2,         /* CUDA runtime error (no GPU, OOM, launch failure) */
    BP_ERR_CAPACITY = -3,     /* a per-env capacity (pairs / live arbiters / colours) overflowed */
    BP_ERR_STATE = -4         /* call order (e.g. step before load_layouts / reset) */
};

/* env status bits reported per env in bp_shipice_outputs.status */
#define BP_ENV_PAIR_OVERFLOW 1
#define BP_ENV_LIVE_OVERFLOW 2
#define BP_ENV_COLOUR_OVERFLOW 4
#define BP_ENV_GJK_ITER_CAP 8

int bp_version(void);
const char* bp_last_error(void);
int bp_device_count(void);

/* ------------------------------------------------------------------------------------------------
 * ship-ice-v0.  Replaces ShipIceEnv.__init__/reset/step/generate_observation
 * (benchpush/environments/ship_ice_nav/ship_ice_env.py:37-106, 223-249, 261-355, 378-409), the
 * pymunk.Space it drives (ship_ice_env.py:117-120, 218-219, 280-281) and the OccupancyGrid channel
 * methods it calls (benchpush/common/occupancy_grid/occupancy_map.py:37-65, 112-140, 300-337,
 * 379-433, 492-587) for a batch of independent environments.
 * ---------------------------------------------------------------------------------------------- */
typedef struct bp_shipice bp_shipice_t;

typedef struct bp_shipice_config {
    int32_t n_envs;
    int32_t max_floes;        /* capacity of obstacles per layout */
    int32_t device;           /* CUDA device ordinal */
    int32_t pair_capacity;    /* neighbour-list capacity per env (0 = default 8 * (max_floes+1)) */
    int32_t live_capacity;    /* touching-pair capacity per env (0 = default max(256, 2*(max_floes+1))) */
    int32_t auto_reset;       /* 1: an env that terminated is reset by its NEXT step call (next trial) */
    /* reference cfg keys that reach the hot path (ship_ice_nav/config.yaml) */
    double dt;                /* cfg.dt = 0.8 */
    int32_t steps;            /* cfg.sim.steps = 400 */
    int32_t iterations;       /* cfg.sim.iterations = 10 */
    double damping;           /* cfg.sim.damping = 0 */
    double obstacle_density;  /* cfg.sim.obstacle_density = 0.001 */
    double target_speed;      /* cfg.target_speed = 0.3 */
    double goal_y;            /* cfg.goal_y = 9 */
    double map_width;         /* cfg.occ.map_width = 12 */
    double map_height;        /* cfg.occ.map_height = 40 */
    int32_t m_to_pix_scale;   /* cfg.occ.m_to_pix_scale = 25 */
    int32_t settle_steps;     /* 1000 (ship_ice_env.py:218) */
    int32_t n_ship_verts;     /* cfg.ship.vertices */
    double ship_verts[64];    /* x0,y0,x1,y1,... (<= 32 vertices) */
    double ship_head[2];      /* cfg.ship.head */
    double ship_tail[2];      /* cfg.ship.tail */
} bp_shipice_config;

/* Fill cfg with the reference defaults (ship_ice_nav/config.yaml). */
void bp_shipice_default_config(bp_shipice_config* cfg);

int bp_shipice_create(const bp_shipice_config* cfg, bp_shipice_t** out);
int bp_shipice_destroy(bp_shipice_t* h);

/* Ice-field layouts = the trials of the reference's experiment pickles
 * (ship_ice_env.py:75-80, 186-216).  HOST pointers; the tables are built (hull, mass, normals —
 * sim_utils.py:136-153, ship.py:77-98) and uploaded once.
 *   n_floes[L], n_verts[L*max_floes], verts[L*max_floes*BP_VMAX*2] (world coordinates),
 *   centres[L*max_floes*2], ship_state[L*3] (x, y, theta). */
int bp_shipice_load_layouts(bp_shipice_t* h, int32_t n_layouts, const int32_t* n_floes, const int32_t* n_verts,
                            const double* verts, const double* centres, const double* ship_state);

typedef struct bp_shipice_outputs {   /* all DEVICE pointers, each may be NULL except obs/reward/terminated */
    uint8_t* obs;                     /* [E][4][150][150]  footprint, goal-edt, orientation, occupancy */
    double* reward;                   /* [E] */
    uint8_t* terminated;              /* [E] */
    uint8_t* truncated;               /* [E] (always 0: ship_ice_env.py:355) */
    double* state;                    /* [E][3] info['state']: pose rounded to 2 decimals */
    double* pose;                     /* [E][3] unrounded pose */
    double* total_work;               /* [E] info['total_work'] */
    double* collision_reward;         /* [E] */
    double* scaled_collision_reward;  /* [E] */
    double* dist_reward;              /* [E] */
    uint8_t* trial_success;           /* [E] */
    int32_t* status;                  /* [E] BP_ENV_* bits */
} bp_shipice_outputs;

/* ShipIceEnv.reset (ship_ice_env.py:223-249) for the envs with mask[e] != 0 (mask NULL = all):
 * build the world from layout_ids[e], run the settle substeps, write the reset observation and info. */
int bp_shipice_reset(bp_shipice_t* h, const int32_t* layout_ids, const uint8_t* mask,
                     const bp_shipice_outputs* out, void* stream);

/* ShipIceEnv.step (ship_ice_env.py:261-355) for every env: actions[E] in [-1, 1]. */
int bp_shipice_step(bp_shipice_t* h, const float* actions, const bp_shipice_outputs* out, void* stream);

/* Same, for the envs with mask[e] != 0 only (mask NULL = all): the others keep their state and their
 * previous outputs.  The reference has no counterpart (one env per process); batched callers use it
 * to hold finished envs, or to de-synchronise episode phases. */
int bp_shipice_step_masked(bp_shipice_t* h, const float* actions, const uint8_t* mask, const bp_shipice_outputs* out, void* stream);

/* Same call with HOST buffers (pageable or pinned): copies actions to the device, steps, copies
 * obs / reward / terminated / truncated (and the non-NULL info arrays) back, and synchronises. */
int bp_shipice_step_host(bp_shipice_t* h, const float* actions, const bp_shipice_outputs* out_host);
/* The same with the observation left on the device for a policy that runs there: actions and every scalar output are HOST
 * buffers as above, the observation kernel writes obs_dev[E][4][150][150] (DEVICE, caller-owned); when out_host->obs is set
 * the image is additionally copied to that host buffer. */
int bp_shipice_step_host_obs_device(bp_shipice_t* h, const float* actions, const bp_shipice_outputs* out_host, uint8_t* obs_dev);

/* cfg.random_start (ship_ice_env.py:201-203; maze_NAMO_env.py:229-238): from now on every reset of env e — explicit or
 * the next-step auto reset inside bp_*_step — starts the ship / robot at poses[3e..3e+2] = (x, y, theta) instead of the
 * layout's start pose, through the full reset (init + settle substeps; the per-layout reset snapshots do not apply).
 * poses: DEVICE, [E][3], owned by the caller and read by later kernels (it may be rewritten between calls to draw a
 * new start for the next episode); NULL restores the layouts' poses.  Synchronous w.r.t. the host. */
int bp_shipice_set_start_poses(bp_shipice_t* h, const double* poses);

/* ShipIceEnv.generate_observation with cfg.egocentric_obs = false (ship_ice_env.py:96, 397-409): DEVICE
 * obs[E][2][gh][gw] u8 on the 0.2 m cell grid the reference builds in that mode (gh = int(map_height / 0.2),
 * gw = int(map_width / 0.2): 200 x 60 by default; bp_shipice_obs_global_shape returns them in hw[2]):
 * channel 0 the ice concentration (5 x 5 block mean of the 1000 x 300 occupancy image of ALL floes, compute_con_gridmap),
 * channel 1 the ship footprint (compute_ship_footprint_planner), both (x * 255) truncated to u8.
 * Envs with mask[e] == 0 keep their previous contents (mask NULL = all).  Enqueued on `stream`. */
int bp_shipice_obs_global_shape(bp_shipice_t* h, int32_t* hw);
int bp_shipice_obs_global(bp_shipice_t* h, uint8_t* obs, const uint8_t* mask, void* stream);

/* ShipIceEnv.generate_observation_low_dim (ship_ice_env.py:358-370), the observation when cfg.low_dim_state:
 * DEVICE out[E][2*max_floes] = |poly_centroid| (x, y) of every floe in creation order, zero padded;
 * counts[E] (DEVICE, may be NULL) = valid entries per env (2 * n_floes).  Enqueued on `stream`. */
int bp_shipice_low_dim(bp_shipice_t* h, double* out, int32_t* counts, void* stream);

/* CostMap.get_obs_from_poly (cost_map.py:274-281) = info['obs']: world-space hull vertices of every
 * floe of env `env`.  HOST outputs: verts[max_floes*BP_VMAX*2], n_verts[max_floes]; returns n_floes. */
int bp_shipice_get_obstacles(bp_shipice_t* h, int32_t env, double* verts, int32_t* n_verts);

/* Body state of env `env` (parity fixtures / checkpoints), HOST: out[(n_floes+1)*6] =
 * position x, y, angle, vx, vy, w per body; the ship is the last row.  Returns the body count. */
int bp_shipice_get_bodies(bp_shipice_t* h, int32_t env, double* out);

/* State of one env for checkpoint / resume and parity fixtures (SURVEY §8b `bp_get_state / bp_set_state`; the reference never
 * serialises env state).  The state is the env's whole block: body poses / velocities, arbiter cache with its
 * accumulated impulses, neighbour list, awake set, step counters, accumulated work.  HOST buffers of bp_shipice_state_bytes(h)
 * bytes; a block is only meaningful for a handle with the same configuration and layouts.  Both calls synchronise. */
size_t bp_shipice_state_bytes(bp_shipice_t* h);
int bp_shipice_get_state(bp_shipice_t* h, int32_t env, void* out_host);
int bp_shipice_set_state(bp_shipice_t* h, int32_t env, const void* in_host);
/* Put env `env` into an EXTERNALLY given body state (e.g. one taken from the reference / the oracle in the middle of an
 * episode): HOST bodies6[n_bodies][6] = position x, y, angle, vx, vy, w in the order of bp_shipice_get_bodies.  The arbiter
 * cache starts empty (as in a pymunk Space whose cached arbiters were dropped), every body is awake, the work accounting
 * restarts from the new poses; layout, episode counters and accumulated work are kept.  Synchronises. */
int bp_shipice_set_bodies(bp_shipice_t* h, int32_t env, const double* bodies6, int32_t n_bodies);

/* Counters of env `env`, HOST out[32]: stamp, n_awake, n_pairs, n_free, n_order, n_colours, status,
 * n_apairs, n_cached, gjk_iters, pairs_evaluated, live_sum, awake_sum, recolours, rebuilds, substeps,
 * epa_pairs, pairs_skipped_by_motion_bound, then 8 per-phase clock sums (only with BP_PHASE_TIMERS). */
int bp_shipice_get_counters(bp_shipice_t* h, int32_t env, int64_t* out);

/* Device timing of the kernels this handle launches (bench.py's roofline): when enabled, CUDA
 * events bracket every step/reset kernel and every observation kernel on their launch stream.
 * get_kernel_ms synchronises and returns, since the last call, out[4] = {step-kernel ms total,
 * step-kernel launches, obs-kernel ms total, obs-kernel launches}. */
int bp_shipice_set_profiling(bp_shipice_t* h, int32_t on);
int bp_shipice_get_kernel_ms(bp_shipice_t* h, double* out4);

/* Bytes of device memory held per env and in total. */
size_t bp_shipice_env_bytes(bp_shipice_t* h);

/* Number of kernel launches issued by this handle so far (for bench accounting). */
int64_t bp_shipice_launch_count(bp_shipice_t* h);

/* ------------------------------------------------------------------------------------------------
 * maze-NAMO-v0.  Replaces MazeNAMO.__init__/reset/step/generate_observation
 * (benchpush/environments/maze_NAMO/maze_NAMO_env.py:42-132, 325-359, 402-485, 514-525), the
 * pymunk.Space it drives (maze_NAMO_env.py:145-216, 268, 415: static Segment walls of
 * common/utils/sim_utils.py:174-181, the compound kinematic robot of common/robot.py:77-118, the
 * (1,3) robot-wall pre_solve handler) and OccupancyGrid.ego_view_map_maze with the maps it needs
 * (common/occupancy_grid/occupancy_map.py:67-94, 142-202, 338-376, 435-485) for a batch of envs.
 * Same handle life cycle, output struct and call semantics as the ship-ice functions above; the
 * observation is [E][4][192][192] u8 (footprint, movable obstacles, walls, goal-distance map) and
 * `dist_reward` carries info['dist increment reward'].
 * ---------------------------------------------------------------------------------------------- */
#define BP_MAZE_OBS_C 4
#define BP_MAZE_OBS_H 192
#define BP_MAZE_OBS_W 192
#define BP_MAZE_MAX_WALLS 16
#define BP_MAZE_MAX_WHEELS 8

typedef struct bp_shipice bp_maze_t;           /* same opaque handle type */
typedef bp_shipice_outputs bp_maze_outputs;

typedef struct bp_maze_config {
    int32_t n_envs;
    int32_t max_boxes;        /* capacity of movable obstacles per layout (cfg.num_obstacles = 5) */
    int32_t device;
    int32_t auto_reset;       /* 1: an env that terminated is reset by its NEXT step call (next layout) */
    /* reference cfg keys that reach the hot path (maze_NAMO/config.yaml) */
    double dt;                /* cfg.dt = 0.8 */
    int32_t steps;            /* cfg.sim.steps = 400 */
    int32_t iterations;       /* cfg.sim.iterations = 10 */
    double damping;           /* cfg.sim.damping = 0 */
    double obstacle_density;  /* cfg.sim.obstacle_density = 0.001 */
    double obstacle_size;     /* cfg.obstacle_size = 0.5 (half side of the square obstacles) */
    double target_speed;      /* cfg.target_speed = 0.15 */
    double width, length;     /* cfg.env.width / length (15 x 15 for maze_version 1) */
    double goal_x, goal_y;    /* cfg.env.goal_x / goal_y */
    double goal_radius;       /* cfg.goal_radius = 1 */
    double min_r;             /* cfg.robot.min_r = 0.86 */
    int32_t m_to_pix_scale;   /* cfg.occ.m_to_pix_scale = 16 (local window 12 m = 192 cells is fixed) */
    int32_t settle_steps;     /* 1000 (maze_NAMO_env.py:268) */
    double start[3];          /* robot start pose x, y, theta (maze_NAMO_env.py:247-250) */
    int32_t n_walls;          /* construct_maze_walls (maze_NAMO_env.py:361-379) */
    double walls[4 * BP_MAZE_MAX_WALLS];            /* x0, y0, x1, y1 per wall */
    int32_t n_robot_verts;    /* cfg.robot.vertices */
    double robot_verts[64];
    int32_t n_wheels;         /* cfg.robot.wheel_vertices */
    int32_t wheel_nv[BP_MAZE_MAX_WHEELS];
    double wheel_verts[BP_MAZE_MAX_WHEELS * 16];    /* up to 8 vertices per wheel, 16 doubles per wheel slot */
} bp_maze_config;

/* Fill cfg with the reference defaults for maze_version 1 or 2 (maze_NAMO/config.yaml, construct_maze_walls). */
void bp_maze_default_config(bp_maze_config* cfg, int32_t maze_version);

/* Builds the two state-independent global maps on the host (wall occupancy: compute_occ_img_walls,
 * occupancy_map.py:67-94; goal-point wavefront distance map: global_goal_point_dist_transform, :435-485). */
int bp_maze_create(const bp_maze_config* cfg, bp_maze_t** out);
int bp_maze_destroy(bp_maze_t* h);

/* Obstacle layouts (the reference samples them with the unseeded global `random`, maze_NAMO_env.py:275-301;
 * here they are inputs).  HOST pointers: n_boxes[L], centres[L*max_boxes*2]. */
int bp_maze_load_layouts(bp_maze_t* h, int32_t n_layouts, const int32_t* n_boxes, const double* centres);

int bp_maze_reset(bp_maze_t* h, const int32_t* layout_ids, const uint8_t* mask, const bp_maze_outputs* out, void* stream);
int bp_maze_step(bp_maze_t* h, const float* actions, const bp_maze_outputs* out, void* stream);
int bp_maze_step_masked(bp_maze_t* h, const float* actions, const uint8_t* mask, const bp_maze_outputs* out, void* stream);
int bp_maze_step_host(bp_maze_t* h, const float* actions, const bp_maze_outputs* out_host);
/* observation (u8[E][4][192][192]) rendered into the DEVICE buffer obs_dev, everything else through host buffers */
int bp_maze_step_host_obs_device(bp_maze_t* h, const float* actions, const bp_maze_outputs* out_host, uint8_t* obs_dev);

/* Same as bp_shipice_set_start_poses, for the maze robot (chassis and wheels share the pose). */
int bp_maze_set_start_poses(bp_maze_t* h, const double* poses);

/* MazeNAMO.generate_observation_low_dim (maze_NAMO_env.py:488-504): DEVICE out[E][2*(max_boxes+1)] =
 * <robot_x, robot_y, |centroid| of obstacle 1, ..., obstacle n-1, 0, 0> — the reference's loop starts at index 1, so
 * obstacle 0 never appears and the last pair stays zero (preserved).  counts[E] (may be NULL) = 2 * (n_boxes + 1). */
int bp_maze_low_dim(bp_maze_t* h, double* out, int32_t* counts, void* stream);

/* info['obs']: world-space hull vertices of every movable obstacle of env `env`.  HOST outputs:
 * verts[max_boxes*BP_VMAX*2], n_verts[max_boxes]; returns the obstacle count. */
int bp_maze_get_obstacles(bp_maze_t* h, int32_t env, double* verts, int32_t* n_verts);
/* Body state of env `env`, HOST out[n*6] (x, y, angle, vx, vy, w) in shape creation order: walls, obstacles,
 * robot chassis, wheels.  Returns the body count. */
int bp_maze_get_bodies(bp_maze_t* h, int32_t env, double* out);
/* Same as bp_shipice_state_bytes / get_state / set_state / set_bodies, for a maze handle. */
size_t bp_maze_state_bytes(bp_maze_t* h);
int bp_maze_get_state(bp_maze_t* h, int32_t env, void* out_host);
int bp_maze_set_state(bp_maze_t* h, int32_t env, const void* in_host);
int bp_maze_set_bodies(bp_maze_t* h, int32_t env, const double* bodies6, int32_t n_bodies);
/* The constant maps, HOST outputs (each may be NULL): wall[H*W] u8, dist_norm[H*W] (global_distance_map),
 * dist_raw[H*W] (info['goal_dt'], the un-normalised hop counts); hw[2] = {H, W}. */
int bp_maze_get_maps(bp_maze_t* h, uint8_t* wall, double* dist_norm, double* dist_raw, int32_t* hw);

/* ------------------------------------------------------------------------------------------------
 * Grid shortest paths.  Replaces spfa.spfa(input_map, source) -> (dists, parents)
 * (deps/spfa/src/main.cpp:10-116; call sites box_delivery_env.py:1121,1134,1226,
 * area_clearing.py:1049,1061, position_controller.py:143) for a batch of maps:
 *   maps[B][H][W] u8 (non-zero = free), sources[B][2] (row, col)
 *   dists[B][H][W] f32 (unreachable -> 0), parents[B][H][W] i32 (raveled index or -1).
 * Distances are bit-identical to the reference (same fp32 relaxation fixed point); parents are a
 * valid shortest-path tree (ties may be broken differently from the reference's queue order).
 * ---------------------------------------------------------------------------------------------- */
int bp_spfa_batch(const uint8_t* maps, const int32_t* sources, int32_t batch, int32_t height, int32_t width,
                  float* dists, int32_t* parents, void* stream);
/* scratch-free host-buffer variant (copies in, runs, copies out, synchronises) */
int bp_spfa_batch_host(const uint8_t* maps, const int32_t* sources, int32_t batch, int32_t height, int32_t width,
                       float* dists, int32_t* parents);
/* Same signatures, REFERENCE-ORDER parents: the reference solver's queue discipline (main.cpp:62-85: FIFO with the
 * small-label-first swap, parent recorded at every strict improvement) restated by one thread per map, so parents[] equals
 * spfa.spfa's own output also where equal-cost predecessors tie.  Sequential per map by construction — for parity runs and for
 * callers that walk the parent chain; bp_spfa_batch_exact synchronises `stream` before returning (it frees its scratch). */
int bp_spfa_batch_exact(const uint8_t* maps, const int32_t* sources, int32_t batch, int32_t height, int32_t width,
                        float* dists, int32_t* parents, void* stream);
int bp_spfa_batch_exact_host(const uint8_t* maps, const int32_t* sources, int32_t batch, int32_t height, int32_t width,
                             float* dists, int32_t* parents);

/* ------------------------------------------------------------------------------------------------
 * box-delivery-v0.  Replaces BoxDeliveryEnv.reset / step / generate_observation
 * (benchpush/environments/box_delivery/box_delivery_env.py:571-632, 634-830, 1045-1059), the
 * pymunk.Space it drives with its collision handlers (:185-229, 294-311, 765-767), the position-control
 * chain (benchpush/common/controller/position_controller.py:56-181, dp.py:217-248) and the 2 N + 2..3
 * spfa.spfa calls per step (:651-655, 736-740, 1115-1138, 1226) for a batch of independent environments.
 * The room maps (configuration spaces, closest-free-cell indices, small obstacle map: :1140-1175) are
 * per-handle inputs built by the host mirror with the libraries the reference itself calls.
 * Observation: u8 [n_envs][224][224][4] (HWC, as the reference stacks the channels on axis 2).
 * ---------------------------------------------------------------------------------------------- */
typedef struct bp_boxdel bp_boxdel_t;

typedef struct bp_boxdel_config {
    int32_t n_envs, num_boxes, device;
    int32_t action_type;          /* 0 = position (pixel index), 1 = heading in [-1, 1], 2 = velocity (linear, angular: two floats per env;
                                     box-delivery: refreshed every substep, box_delivery_env.py:667-700; area-clearing: scaled by
                                     target_speed / max_yaw_rate_step / 2 and set once, area_clearing.py:656-663) — cfg.agent.action_type */
    int32_t steps, iterations, settle_steps, inactivity_cutoff;   /* cfg.sim.steps / iterations, 1000, cfg.misc.inactivity_cutoff */
    int32_t local_map_pixel_width, use_correct_direction_reward;
    int32_t force_box_spfa;       /* 1: always solve spfa per box even when the exact visibility shortcut applies (testing aid) */
    int32_t phase_kernels;        /* 1: substep loop as one kernel per phase over the whole batch (synchronises the stream once per 64 iterations) */
    double ctrl_dt, target_speed, damping;                        /* cfg.controller.dt / target_speed, cfg.sim.damping */
    double box_size, box_density;
    double room_length, room_width, receptacle_width;
    double local_map_width, step_size, robot_length, robot_width;
    double partial_rewards_scale, goal_reward, collision_penalty, non_movement_penalty, correct_direction_reward_scale;
    double ministep_size, shortest_path_channel_scale;
    double chassis_verts[8], bumper_verts[8], wheel_verts[32];   /* cfg.agent.vertices / front_bumper_vertices / wheel_vertices */
    /* ---- area-clearing-v0 (env_kind 1) runs on the same handle type: AreaClearingEnv.reset / step / generate_observation
     * (benchpush/environments/area_clearing/area_clearing.py:563-778, 927-950), same control chain and wall handlers. */
    int32_t env_kind;             /* 0 = box-delivery-v0, 1 = area-clearing-v0 */
    int32_t step_limit;           /* STEP_LIMIT: 10000 (box_delivery_env.py:62) / 5000 (area_clearing.py:76) */
    int32_t post_steps, t_max;    /* area-clearing: cfg.sim.steps substeps after the path; cfg.sim.t_max */
    int32_t n_boundary, n_goals;  /* area-clearing: clearance polygon vertices (<= 8), goal points (<= 64) */
    double lfc, omega_mul, v_mul; /* cfg.controller.Lfc; apply_controller gains (3, 2 box-delivery; 0.5, 5 area-clearing) */
    double bound_x, bound_y;      /* PositionController room clamp in x / y */
    double static_e, static_u;    /* elasticity / friction of the static shapes (0.01, 1.0 box-delivery; 0, 0.99 area-clearing) */
    double robot_code, box_code, box_done_code;   /* overhead-map segment indices of the robot, boxes, completed boxes */
    double footprint_verts[8];    /* polygon drawn for the robot (area-clearing: cfg.agent.footprint_vertices; 0 = chassis hull) */
    int32_t n_footprint;
    int32_t exact_spfa_parents;   /* 1 (default): robot paths / box distances follow the reference solver's own parent tie-breaks (its queue
                                   * restated by one thread per solve).  0: parallel relaxation + deterministic parents: same distances, faster */
    double boundary[16], goals[128];
    double r_boundary, r_cleared, r_putback, r_push_mult, r_terminal, r_truncation, r_nonmove;   /* area_clearing.py:41-52 */
} bp_boxdel_config;

typedef struct bp_boxdel_outputs {  /* every pointer optional */
    uint8_t* obs;                 /* [E][224][224][4] */
    double* reward;               /* [E] */
    uint8_t* terminated;          /* [E] */
    uint8_t* truncated;           /* [E] */
    double* info;                 /* [E][8]: state x, y, angle (2 dp), cumulative_distance, cumulative_reward, total_work, ministeps, work of this step */
    int32_t* info_i;              /* [E][4]: cumulative_boxes, inactivity, boxes left, env status bits */
    uint8_t* completed;           /* [E][num_boxes] box_completed_statuses */
    double* box_verts;            /* [E][num_boxes][4][2] info['obs'] (first `boxes left` polygons valid) */
} bp_boxdel_outputs;

void bp_boxdel_default_config(bp_boxdel_config* cfg);
int bp_boxdel_create(const bp_boxdel_config* cfg, bp_boxdel_t** out);
int bp_boxdel_destroy(bp_boxdel_t* h);
/* HOST pointers; see csrc/boxdel_api.cuh for the array shapes */
int bp_boxdel_load(bp_boxdel_t* h, int32_t n_layouts, const double* starts, const int32_t* n_boxes, const double* boxes,
                   int32_t n_static, const int32_t* st_nv, const double* st_verts, const double* st_pose, const double* st_rad, const int32_t* st_ctype,
                   int32_t H, int32_t W, const uint8_t* cspace, const uint8_t* cspace_thin, const int32_t* closest,
                   const uint8_t* small_free, const uint8_t* robot_chan, const float* ch3_map);
/* Rooms that differ between layouts (obstacle_config small_columns / large_columns / large_divider: the reference redraws the
 * columns / the divider at every reset, box_delivery_env.py:358-471, and rebuilds its configuration space, :1140-1175).
 * Every array of bp_boxdel_load that describes the room gains a leading axis of n_rooms: st_nv / st_rad / st_ctype
 * [n_rooms][n_static], st_verts [n_rooms][n_static][4][2], st_pose [n_rooms][n_static][3], cspace / cspace_thin
 * [n_rooms][H][W], closest [n_rooms][2][H][W], small_free [n_rooms][S][S], ch3_map [n_rooms][H][W] or NULL (box-delivery:
 * computed on the device, one spfa solve per room).  The static slots have the same types in every room; st_nv == 0 marks a
 * collision_type-3 slot that room does not use (fewer columns than the maximum).  layout_room[n_layouts] names the room of
 * each layout (NULL: room 0).  bp_boxdel_load is this call with n_rooms = 1. */
int bp_boxdel_load_rooms(bp_boxdel_t* h, int32_t n_layouts, const double* starts, const int32_t* n_boxes, const double* boxes,
                         int32_t n_rooms, const int32_t* layout_room,
                         int32_t n_static, const int32_t* st_nv, const double* st_verts, const double* st_pose, const double* st_rad, const int32_t* st_ctype,
                         int32_t H, int32_t W, const uint8_t* cspace, const uint8_t* cspace_thin, const int32_t* closest,
                         const uint8_t* small_free, const uint8_t* robot_chan, const float* ch3_map);
/* Rooms built ON THE DEVICE: one set of static shapes per layout (arrays [n_layouts][n_static]...), no maps.  At
 * every reset of an env a thread block rebuilds that env's small obstacle map, both configuration spaces, the nearest-free-cell
 * map (scipy's distance_transform_edt indices incl. its tie-breaks) and the receptacle distance field from the layout's shapes
 * (csrc/boxdel_room.cuh), as the reference does in update_configuration_space at every reset (box_delivery_env.py:1140-1175).
 * st_npx[n_layouts][n_static] / st_pxy[n_layouts][n_static][4][2]: the pixel polygon (x, y) of every collision_type-3 shape as the
 * reference passes it to cv2.fillPoly ((world * ppm).astype(int32) + offset, y flipped); 0 vertices = not drawn. */
int bp_boxdel_load_rooms_device(bp_boxdel_t* h, int32_t n_layouts, const double* starts, const int32_t* n_boxes, const double* boxes,
                                int32_t n_static, const int32_t* st_nv, const double* st_verts, const double* st_pose, const double* st_rad, const int32_t* st_ctype,
                                const int32_t* st_npx, const int32_t* st_pxy, int32_t H, int32_t W, const uint8_t* robot_chan);
/* HOST buffers (any may be NULL): the maps env `env` uses at the moment (parity aid; synchronises) */
int bp_boxdel_get_room_maps(bp_boxdel_t* h, int32_t env, uint8_t* cspace, uint8_t* cspace_thin, int32_t* closest, uint8_t* small_free, float* ch3);
/* DEVICE pointers, enqueued on `stream`, no host synchronisation */
int bp_boxdel_reset(bp_boxdel_t* h, const int32_t* layout_ids, const uint8_t* mask, const bp_boxdel_outputs* out, void* stream);
int bp_boxdel_step(bp_boxdel_t* h, const float* actions, const uint8_t* mask, const bp_boxdel_outputs* out, void* stream);
/* HOST buffers: copies in, steps, copies out, synchronises */
int bp_boxdel_step_host(bp_boxdel_t* h, const float* actions, const bp_boxdel_outputs* out_host);
/* same, with the observation rendered into the DEVICE buffer obs_dev (u8[E][LW][LW][4]) for a GPU-resident policy instead of being
 * copied to the host (out_host->obs is ignored) */
int bp_boxdel_step_host_obs_device(bp_boxdel_t* h, const float* actions, const bp_boxdel_outputs* out_host, uint8_t* obs_dev);
/* HOST out[(1 + num_boxes)][7]: origin x, y, angle, vx, vy, w, alive; row 0 = robot */
int bp_boxdel_get_bodies(bp_boxdel_t* h, int32_t env, double* out);
int bp_boxdel_get_step_debug(bp_boxdel_t* h, int32_t env, double* out16, double* dists, double* path);
/* kernel timing with CUDA events on the launching stream: out3 = summed k_bd_move ms, summed k_bd_obs ms, timed steps */
int bp_boxdel_set_profiling(bp_boxdel_t* h, int32_t on);
int bp_boxdel_get_kernel_ms(bp_boxdel_t* h, double* out3);
int64_t bp_boxdel_launch_count(bp_boxdel_t* h);
size_t bp_boxdel_env_bytes(bp_boxdel_t* h);

#ifdef __cplusplus
}
#endif
#endif /* BENCHPUSH_B200_H */
