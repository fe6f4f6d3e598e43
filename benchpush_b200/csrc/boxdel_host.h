// Host-side builder of the box-delivery tables (parameters, shapes, layouts, dense pair table).
// Shared by the CUDA library (which uploads them) and by the host-emulation test harness.
#pragma once
#include <cmath>
#include <cstring>
#include <string>
#include <vector>
#include "../../include/benchpush_b200.h"
#include "layout_prep.h"
#include "boxdel_core.cuh"

namespace bp {
namespace bd {

struct BdHost { BdConst K; std::vector<BdLayout> LY; std::vector<BdStatics> RM; std::vector<short> pa, pb; };

inline void bd_fill_prep(const bp_prep::BodyPrep& b, double (*lv)[2], double (*ln)[2], int* nv) {
    *nv = (int)b.verts.size();
    for (int i = 0; i < *nv && i < SV; i++) { lv[i][0] = b.verts[i].x; lv[i][1] = b.verts[i].y; ln[i][0] = b.normals[i].x; ln[i][1] = b.normals[i].y; }
}

// Returns "" or an error message.  Static-shape arrays carry a leading room axis: st_nv[n_rooms][n_static], st_verts[n_rooms][n_static][4][2],
// st_pose[n_rooms][n_static][3], st_rad[n_rooms][n_static], st_ctype[n_rooms][n_static] (the same slot types in every room; st_nv 0 = slot
// unused in that room).  layout_room[n_layouts] (nullptr: every layout uses room 0).
inline std::string bd_host_build(BdHost& HB, const bp_boxdel_config& c, int n_layouts, const double* starts, const int32_t* n_boxes, const double* boxes,
                                 int n_rooms, const int32_t* layout_room,
                                 int n_static, const int32_t* st_nv, const double* st_verts, const double* st_pose, const double* st_rad,
                                 const int32_t* st_ctype, int H, int W, const int32_t* st_npx = nullptr, const int32_t* st_pxy = nullptr) {
    memset(&HB.K, 0, sizeof(HB.K));
    BdConst& K = HB.K;
    BdParams& P = K.P;
    P.exact_parents = c.exact_spfa_parents ? 1 : 0;
    P.NB = c.num_boxes; P.n_static = n_static; P.n_s3 = 0; P.recept = -1;
    for (int k = 0; k < n_static; k++) { if (st_ctype[k] == 3) P.s3[P.n_s3++] = k; else if (st_ctype[k] == 4) P.recept = k; }
    if (P.recept < 0 && c.env_kind == 0) return std::string("no receptacle (collision_type 4) among the static shapes");
    P.n_slots = bd_build_pairs(P, nullptr, nullptr);
    P.iterations = c.iterations; P.persistence = 3; P.settle_steps = c.settle_steps;
    P.dt = c.ctrl_dt / c.steps; P.damp = std::pow(c.damping, P.dt);
    P.bias_coef = 1.0 - std::pow(std::pow(1.0 - 0.1, 60.0), P.dt); P.slop = 0.1;
    P.ctrl_dt = c.ctrl_dt; P.target_speed = c.target_speed;
    P.H = H; P.W = W; P.LW = c.local_map_pixel_width; P.SMALL = P.LW + 20;
    P.CROP = (int)(std::ceil(P.LW * std::sqrt(2.0) / 2) * 2);
    P.ppm = (double)P.LW / c.local_map_width;
    P.small_i0 = (int)(H / 2.0 - P.SMALL / 2.0); P.small_j0 = (int)(W / 2.0 - P.SMALL / 2.0);
    P.px_off = (int)(c.local_map_width * P.ppm / 2) + 10;
    P.local_map_width = c.local_map_width; P.step_size = c.step_size; P.room_length = c.room_length; P.room_width = c.room_width;
    P.robot_radius = (std::pow(c.robot_length * c.robot_length + c.robot_width * c.robot_width, 0.5) / 2) * 1.2;
    P.robot_half_width = (c.robot_length > c.robot_width ? c.robot_length : c.robot_width) / 2;
    P.recept_x = c.room_length / 2 - c.receptacle_width / 2; P.recept_y = c.room_width / 2 - c.receptacle_width / 2;
    P.partial_scale = c.partial_rewards_scale; P.goal_reward = c.goal_reward; P.collision_penalty = c.collision_penalty;
    P.nonmove_penalty = c.non_movement_penalty; P.correct_scale = c.correct_direction_reward_scale; P.ministep = c.ministep_size;
    P.use_correct = c.use_correct_direction_reward; P.inactivity_cutoff = c.inactivity_cutoff; P.num_boxes = c.num_boxes; P.action_type = c.action_type;
    P.sp_div = (std::sqrt(2.0) * P.LW) / P.ppm; P.sp_scale = c.shortest_path_channel_scale;
    P.kind = c.env_kind; P.step_limit = c.step_limit; P.post_steps = (c.env_kind == 0) ? c.steps : c.post_steps;   // box-delivery velocity control runs cfg.sim.steps substeps
    P.t_max = c.t_max;
    P.lfc = c.lfc; P.omega_mul = c.omega_mul; P.v_mul = c.v_mul; P.bound_x = c.bound_x; P.bound_y = c.bound_y;
    P.robot_code = c.robot_code; P.box_code = c.box_code; P.box_done_code = c.box_done_code; P.draw_recept = (c.env_kind == 0);
    if (c.n_boundary < 0 || c.n_boundary > 8 || c.n_goals < 0 || c.n_goals > 64) return std::string("n_boundary / n_goals out of range");
    P.n_bnd = c.n_boundary; P.n_goal = c.n_goals;
    for (int i = 0; i < c.n_boundary; i++) { P.bnd[i][0] = c.boundary[2 * i]; P.bnd[i][1] = c.boundary[2 * i + 1]; }
    for (int i = 0; i < c.n_goals; i++) { P.goal[i][0] = c.goals[2 * i]; P.goal[i][1] = c.goals[2 * i + 1]; }
    P.r_boundary = c.r_boundary; P.r_cleared = c.r_cleared; P.r_putback = c.r_putback; P.r_push_mult = c.r_push_mult;
    P.r_terminal = c.r_terminal; P.r_truncation = c.r_truncation; P.r_nonmove = c.r_nonmove;
    if (c.env_kind == 1) { P.recept_x = 0.0; P.recept_y = 0.0; }
    // ---- shapes
    BdShapes& S = K.S;
    {
        std::vector<bp_prep::P2> cv(4);
        for (int i = 0; i < 4; i++) cv[i] = bp_prep::P2{c.chassis_verts[2 * i], c.chassis_verts[2 * i + 1]};
        bp_prep::BodyPrep ch = bp_prep::prep_body(bp_prep::recentre_on_hull_centroid(cv), 0.0, 0.0, false);   // D8: only the chassis is recentred
        if (ch.verts.size() < 3 || ch.verts.size() > SV) return std::string("chassis hull vertex count out of range");
        bd_fill_prep(ch, S.rb_lv[0], S.rb_ln[0], &S.rb_nv[0]); S.rb_rad[0] = 0.0; S.rb_e[0] = 0.01; S.rb_u[0] = 1.0;
        for (int w = 0; w < 5; w++) {
            const double* src = (w < 4) ? &c.wheel_verts[8 * w] : c.bumper_verts;
            for (int i = 0; i < 4; i++) cv[i] = bp_prep::P2{src[2 * i], src[2 * i + 1]};
            bp_prep::BodyPrep b = bp_prep::prep_body(cv, 0.02, 0.0, false);
            if (b.verts.size() < 3 || b.verts.size() > SV) return std::string("wheel / bumper hull vertex count out of range");
            bd_fill_prep(b, S.rb_lv[1 + w], S.rb_ln[1 + w], &S.rb_nv[1 + w]); S.rb_rad[1 + w] = 0.02; S.rb_e[1 + w] = 0.01; S.rb_u[1 + w] = 0.0;
        }
        HB.RM.assign(n_rooms, BdStatics());
        for (int r = 0; r < n_rooms; r++) {
            BdStatics& T = HB.RM[r];
            memset(&T, 0, sizeof(T));
            for (int k = 0; k < n_static; k++) {
                const int q = r * n_static + k;
                if (st_ctype[q] != st_ctype[k]) return std::string("static slot types differ between rooms");
                T.st_ctype[k] = st_ctype[q];
                if (st_nv[q] == 0 && st_ctype[q] == 3) { T.st_nv[k] = 0; continue; }     // slot unused in this room
                if (st_nv[q] < 3 || st_nv[q] > SV) return std::string("static shape vertex count out of range");
                std::vector<bp_prep::P2> sv(st_nv[q]);
                for (int i = 0; i < st_nv[q]; i++) sv[i] = bp_prep::P2{st_verts[((size_t)q * SV + i) * 2], st_verts[((size_t)q * SV + i) * 2 + 1]};
                bp_prep::BodyPrep b = bp_prep::prep_body(sv, st_rad[q], 0.0, false);
                if (b.verts.size() < 3) return std::string("degenerate static shape");
                bd_fill_prep(b, T.st_lv[k], T.st_ln[k], &T.st_nv[k]);
                T.st_pose[k][0] = st_pose[3 * q]; T.st_pose[k][1] = st_pose[3 * q + 1]; T.st_pose[k][2] = st_pose[3 * q + 2];
                T.st_rad[k] = st_rad[q]; T.st_e[k] = c.static_e; T.st_u[k] = c.static_u;
            }
            // pixel polygons for the on-device room build: st_npx[n_rooms][n_static], st_pxy[n_rooms][n_static][4][2] (x, y)
            for (int k = 0; k < n_static && st_npx && st_pxy; k++) {
                const int q = r * n_static + k;
                if (st_npx[q] < 0 || st_npx[q] > SV) return std::string("st_npx out of range");
                T.st_npx[k] = st_npx[q];
                for (int i = 0; i < st_npx[q]; i++) { T.st_px[k][i] = st_pxy[((size_t)q * SV + i) * 2]; T.st_py[k][i] = st_pxy[((size_t)q * SV + i) * 2 + 1]; }
            }
        }
        S.box_rad = 0.02; S.box_e = 0.01; S.box_u = 1.0;     // create_polygon ignores its radius argument (D7)
        // polygon drawn for the robot in the overhead map: cfg.agent.footprint_vertices (area-clearing) or the chassis hull
        if (c.n_footprint > 0) {
            if (c.n_footprint > SV) return std::string("n_footprint out of range");
            S.n_foot = c.n_footprint;
            for (int i = 0; i < c.n_footprint; i++) { S.rb_foot[i][0] = c.footprint_verts[2 * i]; S.rb_foot[i][1] = c.footprint_verts[2 * i + 1]; }
        } else {
            S.n_foot = S.rb_nv[0];
            for (int i = 0; i < S.n_foot; i++) { S.rb_foot[i][0] = S.rb_lv[0][i][0]; S.rb_foot[i][1] = S.rb_lv[0][i][1]; }
        }
    }
    // ---- layouts
    std::vector<BdLayout>& LY = HB.LY; LY.assign(n_layouts, BdLayout());
    const double bs = c.box_size / 2;
    for (int l = 0; l < n_layouts; l++) {
        BdLayout& L = LY[l];
        memset(&L, 0, sizeof(L));
        L.start[0] = starts[3 * l]; L.start[1] = starts[3 * l + 1]; L.start[2] = starts[3 * l + 2];
        if (n_boxes[l] < 0 || n_boxes[l] > c.num_boxes) return std::string("n_boxes out of range");
        L.n_boxes = n_boxes[l];
        L.room = layout_room ? layout_room[l] : 0;
        if (L.room < 0 || L.room >= n_rooms) return std::string("layout_room out of range");
        for (int i = 0; i < L.n_boxes; i++) {
            const double* b = boxes + ((size_t)l * c.num_boxes + i) * 3;
            double x = b[0], y = b[1];
            // generate_boxes :548-556 world vertices, then generate_sim_boxes: vertices - position (sim_utils.py:160-165)
            double wv[8] = {x + bs, y + bs, x - bs, y + bs, x - bs, y - bs, x + bs, y - bs};
            std::vector<bp_prep::P2> rel(4);
            for (int q = 0; q < 4; q++) rel[q] = bp_prep::P2{wv[2 * q] - x, wv[2 * q + 1] - y};
            bp_prep::BodyPrep bp_ = bp_prep::prep_body(bp_prep::recentre_on_hull_centroid(rel), 0.02, c.box_density, true);
            if (bp_.verts.size() != SV) return std::string("degenerate box");
            int nvv; bd_fill_prep(bp_, L.box_lv[i], L.box_ln[i], &nvv);
            L.box_cog[i][0] = bp_.cog.x; L.box_cog[i][1] = bp_.cog.y; L.box_minv[i] = bp_.m_inv; L.box_iinv[i] = bp_.i_inv;
            L.box_init[i][0] = x; L.box_init[i][1] = y; L.box_init[i][2] = b[2];
        }
    }
    HB.pa.assign(P.n_slots, 0); HB.pb.assign(P.n_slots, 0);
    bd_build_pairs(P, HB.pa.data(), HB.pb.data());
    K.n_layouts = n_layouts; K.n_rooms = n_rooms; K.n_envs = c.n_envs; K.env_bytes = bd_env_bytes(P.n_slots);
    return "";
}

// 1 when the first / last rows and columns of every room's configuration space are blocked (BdParams::border_blocked)
inline int bd_border_blocked(const uint8_t* cspace, size_t n_rooms, int H, int W) {
    const size_t nv = (size_t)H * W;
    for (size_t r = 0; r < n_rooms; r++) {
        const uint8_t* m = cspace + r * nv;
        for (int j = 0; j < W; j++) if (m[j] || m[(size_t)(H - 1) * W + j]) return 0;
        for (int i = 0; i < H; i++) if (m[(size_t)i * W] || m[(size_t)i * W + W - 1]) return 0;
    }
    return 1;
}

// fourth observation channel of box-delivery before the window shift: create_global_shortest_path_to_receptacle_map
// (box_delivery_env.py:1115-1129): spfa field / ppm (f32), / ((sqrt 2 LW) / ppm) (f64, rounded to f32), * scale (f32)
inline void bd_scale_receptacle_map(const BdParams& P, float* m, size_t n) {
    for (size_t i = 0; i < n; i++) {
        float a = m[i] / (float)P.ppm;
        float b = (float)((double)a / P.sp_div);
        m[i] = b * (float)P.sp_scale;
    }
}

}  // namespace bd
}  // namespace bp
