// Host-side builders for the layout tables and simulation parameters of ship-ice-v0.
// Shared by the CUDA library (which uploads the tables) and by the host-emulation test harness.
#pragma once
#include <cmath>
#include <cstdint>
#include <cstring>
#include <string>
#include <vector>
#include "layout_prep.h"
#include "sim_core.cuh"

namespace bp {

struct HostLayouts {
    int NB = 0, L = 0;
    std::vector<real> lv, ln, rad, rbound, minv, iinv, cog, mat_e, mat_u, init;      // body-state precision of the build
    std::vector<int> nv, nbody, robot0;
    // raw (un-hulled) world-space trial vertices: the reset observation rasterises these
    // (ship_ice_env.py:207,248), plus |centroid| of each raw polygon for the local-range test
    std::vector<double> rawv, rawcen;
    std::vector<int> rawn;
    // robot footprint vertices as configured (all of cfg.ship.vertices, not the hull: occupancy_map.py:320-336)
    std::vector<double> robot_raw;
    double head[2] = {1.0, 0.0}, tail[2] = {-1.0, 0.0};
};

struct ShipIceMaterial {
    double density = 0.001, floe_radius = 0.02, floe_e = 0.01, floe_u = 1.0;
    double ship_radius = 0.02, ship_e = 0.01, ship_u = 1.0;
};

// np.abs(poly_centroid(v)) of geometry/polygon.py:32-41 on a raw vertex list
inline void raw_centroid_abs(const double* v, int n, double* cx, double* cy) {
    double s1 = 0.0, s2 = 0.0;
    for (int i = 0; i < n; i++) { int j = (i == 0) ? n - 1 : i - 1; s1 += v[2 * i] * v[2 * j + 1]; s2 += v[2 * i + 1] * v[2 * j]; }
    double A = 0.5 * std::fabs(s1 - s2);
    double sx = 0.0, sy = 0.0;
    for (int i = 0; i < n; i++) {
        int j = (i == 0) ? n - 1 : i - 1;
        double u = v[2 * i] * v[2 * j + 1] - v[2 * j] * v[2 * i + 1];
        sx += (v[2 * i] + v[2 * j]) * u; sy += (v[2 * i + 1] + v[2 * j + 1]) * u;
    }
    double k = 1.0 / (6.0 * A);
    *cx = std::fabs(k * sx); *cy = std::fabs(k * sy);
}

// n_floes[L]; n_verts[L*max_floes]; verts[L*max_floes*VMAX*2] world coords; centres[L*max_floes*2];
// ship_state[L*3]; ship_verts[n_ship_verts*2].  Returns "" or an error message.
inline std::string build_shipice_layouts(HostLayouts& H, int n_layouts, int max_floes, const int32_t* n_floes,
                                         const int32_t* n_verts, const double* verts, const double* centres,
                                         const double* ship_state, const double* ship_verts, int n_ship_verts,
                                         const ShipIceMaterial& mat) {
    const int NB = max_floes + 1;
    H.NB = NB; H.L = n_layouts;
    size_t LB = (size_t)n_layouts * NB;
    H.lv.assign(LB * VMAX * 2, 0.0); H.ln.assign(LB * VMAX * 2, 0.0); H.nv.assign(LB, 0);
    H.rad.assign(LB, 0.0); H.rbound.assign(LB, 0.0); H.minv.assign(LB, 0.0); H.iinv.assign(LB, 0.0); H.cog.assign(LB * 2, 0.0);
    H.mat_e.assign(LB, 0.0); H.mat_u.assign(LB, 0.0); H.init.assign(LB * 3, 0.0); H.nbody.assign(n_layouts, 0); H.robot0.assign(n_layouts, 0);
    H.rawv.assign(LB * VMAX * 2, 0.0); H.rawn.assign(LB, 0); H.rawcen.assign(LB * 2, 0.0);
    if (n_ship_verts > 32) return "too many ship vertices";
    H.robot_raw.assign(ship_verts, ship_verts + 2 * n_ship_verts);
    std::vector<bp_prep::P2> sv(n_ship_verts);
    for (int i = 0; i < n_ship_verts; i++) sv[i] = bp_prep::P2{ship_verts[2 * i], ship_verts[2 * i + 1]};
    bp_prep::BodyPrep ship = bp_prep::prep_body(sv, mat.ship_radius, 0.0, false);
    if ((int)ship.verts.size() > VMAX) return "ship hull has too many vertices";
    auto store = [&](size_t idx, const bp_prep::BodyPrep& b, double radius, double e, double u, double x, double y, double a) {
        H.nv[idx] = (int)b.verts.size();
        for (size_t i = 0; i < b.verts.size(); i++) {
            H.lv[(idx * VMAX + i) * 2] = b.verts[i].x; H.lv[(idx * VMAX + i) * 2 + 1] = b.verts[i].y;
            H.ln[(idx * VMAX + i) * 2] = b.normals[i].x; H.ln[(idx * VMAX + i) * 2 + 1] = b.normals[i].y;
        }
        double rb = 0.0;
        for (size_t i = 0; i < b.verts.size(); i++) rb = std::fmax(rb, std::hypot(b.verts[i].x - b.cog.x, b.verts[i].y - b.cog.y));
        H.rbound[idx] = rb * 1.000001 + radius;
        H.rad[idx] = radius; H.minv[idx] = b.m_inv; H.iinv[idx] = b.i_inv;
        H.cog[idx * 2] = b.cog.x; H.cog[idx * 2 + 1] = b.cog.y;
        H.mat_e[idx] = e; H.mat_u[idx] = u;
        H.init[idx * 3] = x; H.init[idx * 3 + 1] = y; H.init[idx * 3 + 2] = a;
    };
    for (int l = 0; l < n_layouts; l++) {
        int nf = n_floes[l];
        if (nf > max_floes || nf < 0) return "n_floes out of range";
        if (nf + 1 > 65535) return "too many bodies";
        H.nbody[l] = nf + 1; H.robot0[l] = nf;
        for (int f = 0; f < nf; f++) {
            size_t src = (size_t)l * max_floes + f;
            size_t idx = (size_t)l * NB + f;
            int n = n_verts[src];
            if (n < 3 || n > VMAX) return "floe vertex count out of range";
            const double* v = verts + src * VMAX * 2;
            double cx = centres[src * 2], cy = centres[src * 2 + 1];
            std::vector<bp_prep::P2> rel(n);
            for (int i = 0; i < n; i++) rel[i] = bp_prep::P2{v[2 * i] - cx, v[2 * i + 1] - cy};
            bp_prep::BodyPrep b = bp_prep::prep_body(bp_prep::recentre_on_hull_centroid(rel), mat.floe_radius, mat.density, true);
            if (b.verts.size() < 3) return "degenerate floe";
            store(idx, b, mat.floe_radius, mat.floe_e, mat.floe_u, cx, cy, 0.0);
            H.rawn[idx] = n;
            memcpy(&H.rawv[idx * VMAX * 2], v, sizeof(double) * 2 * n);
            raw_centroid_abs(v, n, &H.rawcen[idx * 2], &H.rawcen[idx * 2 + 1]);
        }
        size_t ridx = (size_t)l * NB + nf;
        store(ridx, ship, mat.ship_radius, mat.ship_e, mat.ship_u, ship_state[3 * l], ship_state[3 * l + 1], ship_state[3 * l + 2]);
    }
    return "";
}

struct ShipIceConfig {           // mirrors the reference cfg keys that reach the hot path
    double dt = 0.8; int steps = 400; int iterations = 10; double damping = 0.0;
    double target_speed = 0.3; double goal_y = 9.0; double map_width = 12.0, map_height = 40.0;
    int m_to_pix_scale = 25; int settle_steps = 1000;
    double margin_body = 0.3, margin_robot = 0.6;
};

inline SimParams make_params(const ShipIceConfig& c) {
    SimParams P;
    P.iterations = c.iterations; P.persistence = 3; P.substeps = c.steps; P.settle_steps = c.settle_steps;
    P.dt = c.dt / c.steps;
    P.damp = std::pow(c.damping, P.dt);
    double collision_bias = std::pow(1.0 - 0.1, 60.0);
    P.bias_coef = 1.0 - std::pow(collision_bias, P.dt);
    P.slop = 0.1;
    P.target_speed = c.target_speed;
    P.yaw_step = (M_PI / 2) / 7;
    P.yaw_lo = 0.0; P.yaw_hi = M_PI;
    P.map_w = c.map_width; P.map_h = c.map_height; P.goal_y = c.goal_y;
    P.beta = 30.0; P.boundary_limit = 0.0; P.boundary_penalty = -50.0; P.terminal_reward = 200.0; P.yaw_penalty = 0.0;
    P.margin_body = c.margin_body; P.margin_robot = c.margin_robot;
    P.kind = KIND_SHIPICE; P.n_walls = 0; P.goal_x = 0.0; P.goal_gy = 0.0; P.goal_reach = 0.0; P.k_increment = 0.0;
    P.pix_scale = (double)c.m_to_pix_scale; P.dist_map = nullptr; P.dist_h = 0; P.dist_w = 0;
    return P;
}

// ------------------------------------------------------------------------------------------------ maze-NAMO
struct MazeConfig {              // mirrors the reference cfg keys that reach the hot path (maze_NAMO/config.yaml)
    double dt = 0.8; int steps = 400; int iterations = 10; double damping = 0.0;
    double target_speed = 0.15; double obstacle_density = 0.001; double obstacle_size = 0.5;
    double width = 15.0, length = 15.0, goal_x = 3.75, goal_y = 3.75, goal_radius = 1.0, min_r = 0.86;
    int m_to_pix_scale = 16; int settle_steps = 1000;
    double start[3] = {11.25, 3.75, M_PI / 2};
    int n_walls = 0; double walls[4 * 16] = {0};          // x0, y0, x1, y1 per wall segment
    int n_robot_verts = 0; double robot_verts[64] = {0};  // chassis (cfg.robot.vertices)
    int n_wheels = 0; int wheel_nv[8] = {0}; double wheel_verts[8 * 16] = {0};   // up to 8 verts per wheel
    double margin_body = 0.3, margin_robot = 0.6;
};

// Bodies of a maze layout, in the reference's shape creation order (maze_NAMO_env.py:224-262: walls, then
// obstacles, then the robot chassis and wheels): [0, n_walls) static segments of radius 0.5 (e .5, mu .5),
// [n_walls, n_walls + n_boxes) dynamic squares (density, r .02, e .01, mu 1; sim_utils.py:136-163),
// then the kinematic robot shapes (r .02, e .01, friction left at pymunk's default 0; robot.py:77-118).
// centres[L*max_boxes*2]; n_boxes[L].
inline std::string build_maze_layouts(HostLayouts& H, int n_layouts, int max_boxes, const int32_t* n_boxes, const double* centres,
                                      const MazeConfig& c) {
    const int n_robot = 1 + c.n_wheels;
    const int NB = c.n_walls + max_boxes + n_robot;
    H.NB = NB; H.L = n_layouts;
    size_t LB = (size_t)n_layouts * NB;
    H.lv.assign(LB * VMAX * 2, 0.0); H.ln.assign(LB * VMAX * 2, 0.0); H.nv.assign(LB, 0);
    H.rad.assign(LB, 0.0); H.rbound.assign(LB, 0.0); H.minv.assign(LB, 0.0); H.iinv.assign(LB, 0.0); H.cog.assign(LB * 2, 0.0);
    H.mat_e.assign(LB, 0.0); H.mat_u.assign(LB, 0.0); H.init.assign(LB * 3, 0.0); H.nbody.assign(n_layouts, 0); H.robot0.assign(n_layouts, 0);
    H.rawv.assign(LB * VMAX * 2, 0.0); H.rawn.assign(LB, 0); H.rawcen.assign(LB * 2, 0.0);
    H.robot_raw.assign(c.robot_verts, c.robot_verts + 2 * c.n_robot_verts);
    auto store = [&](size_t idx, const bp_prep::BodyPrep& b, double radius, double e, double u, double x, double y, double a) {
        H.nv[idx] = (int)b.verts.size();
        for (size_t i = 0; i < b.verts.size(); i++) {
            H.lv[(idx * VMAX + i) * 2] = b.verts[i].x; H.lv[(idx * VMAX + i) * 2 + 1] = b.verts[i].y;
            H.ln[(idx * VMAX + i) * 2] = b.normals[i].x; H.ln[(idx * VMAX + i) * 2 + 1] = b.normals[i].y;
        }
        double rb = 0.0;
        for (size_t i = 0; i < b.verts.size(); i++) rb = std::fmax(rb, std::hypot(b.verts[i].x - b.cog.x, b.verts[i].y - b.cog.y));
        H.rbound[idx] = rb * 1.000001 + radius;
        H.rad[idx] = radius; H.minv[idx] = b.m_inv; H.iinv[idx] = b.i_inv;
        H.cog[idx * 2] = b.cog.x; H.cog[idx * 2 + 1] = b.cog.y;
        H.mat_e[idx] = e; H.mat_u[idx] = u;
        H.init[idx * 3] = x; H.init[idx * 3 + 1] = y; H.init[idx * 3 + 2] = a;
    };
    std::vector<bp_prep::P2> cv(c.n_robot_verts);
    for (int i = 0; i < c.n_robot_verts; i++) cv[i] = bp_prep::P2{c.robot_verts[2 * i], c.robot_verts[2 * i + 1]};
    bp_prep::BodyPrep chassis = bp_prep::prep_body(cv, 0.02, 0.0, false);
    if ((int)chassis.verts.size() > VMAX || chassis.verts.size() < 3) return "robot hull vertex count out of range";
    std::vector<bp_prep::BodyPrep> wheels;
    for (int w = 0; w < c.n_wheels; w++) {
        std::vector<bp_prep::P2> wv(c.wheel_nv[w]);
        for (int i = 0; i < c.wheel_nv[w]; i++) wv[i] = bp_prep::P2{c.wheel_verts[(w * 8 + i) * 2], c.wheel_verts[(w * 8 + i) * 2 + 1]};
        wheels.push_back(bp_prep::prep_body(wv, 0.02, 0.0, false));
        if (wheels.back().verts.size() < 3) return "degenerate wheel";
    }
    for (int l = 0; l < n_layouts; l++) {
        int nbx = n_boxes[l];
        if (nbx < 0 || nbx > max_boxes) return "n_boxes out of range";
        size_t base = (size_t)l * NB;
        for (int w = 0; w < c.n_walls; w++) {
            // pymunk.Segment(static_body, a, b, 0.5): cpSegmentShapeInit, n = rperp(normalize(b - a))
            size_t idx = base + w;
            double ax = c.walls[4 * w], ay = c.walls[4 * w + 1], bx = c.walls[4 * w + 2], by = c.walls[4 * w + 3];
            double dx = bx - ax, dy = by - ay;
            double inv = 1.0 / (std::sqrt(dx * dx + dy * dy) + 2.2250738585072014e-308);
            double ux = dx * inv, uy = dy * inv;
            H.nv[idx] = 2;
            H.lv[(idx * VMAX) * 2] = ax; H.lv[(idx * VMAX) * 2 + 1] = ay; H.lv[(idx * VMAX + 1) * 2] = bx; H.lv[(idx * VMAX + 1) * 2 + 1] = by;
            H.ln[(idx * VMAX) * 2] = uy; H.ln[(idx * VMAX) * 2 + 1] = -ux; H.ln[(idx * VMAX + 1) * 2] = -uy; H.ln[(idx * VMAX + 1) * 2 + 1] = ux;
            H.rad[idx] = 0.5; H.rbound[idx] = 0.0; H.minv[idx] = 0.0; H.iinv[idx] = 0.0;
            H.mat_e[idx] = 0.5; H.mat_u[idx] = 0.5;
        }
        for (int k = 0; k < nbx; k++) {
            size_t idx = base + c.n_walls + k;
            double cx = centres[((size_t)l * max_boxes + k) * 2], cy = centres[((size_t)l * max_boxes + k) * 2 + 1];
            double sz = c.obstacle_size;
            // maze_NAMO_env.py:312-319: world vertices of the square, in this order
            double wv[8] = {cx + sz, cy + sz, cx - sz, cy + sz, cx - sz, cy - sz, cx + sz, cy - sz};
            std::vector<bp_prep::P2> rel(4);
            for (int i = 0; i < 4; i++) rel[i] = bp_prep::P2{wv[2 * i] - cx, wv[2 * i + 1] - cy};
            bp_prep::BodyPrep b = bp_prep::prep_body(bp_prep::recentre_on_hull_centroid(rel), 0.02, c.obstacle_density, true);
            if (b.verts.size() < 3) return "degenerate obstacle";
            store(idx, b, 0.02, 0.01, 1.0, cx, cy, 0.0);
            H.rawn[idx] = 4;
            memcpy(&H.rawv[idx * VMAX * 2], wv, sizeof(wv));
            raw_centroid_abs(wv, 4, &H.rawcen[idx * 2], &H.rawcen[idx * 2 + 1]);
        }
        int r0 = c.n_walls + nbx;
        H.robot0[l] = r0; H.nbody[l] = r0 + n_robot;
        store(base + r0, chassis, 0.02, 0.01, 0.0, c.start[0], c.start[1], c.start[2]);
        for (int w = 0; w < c.n_wheels; w++) store(base + r0 + 1 + w, wheels[w], 0.02, 0.01, 0.0, c.start[0], c.start[1], c.start[2]);
    }
    return "";
}

// skimage point_in_polygon (crossing test; vertices and edges count as inside) on pixel-space vertices —
// host twin of pip_inclusive (raster.cuh) for the maps that are built once per handle.
inline bool host_pip(const double* px, const double* py, int n, double x, double y) {
    const double eps = 1e-12;
    unsigned rc = 0, lc = 0;
    double x1 = px[n - 1] - x, y1 = py[n - 1] - y;
    for (int i = 0; i < n; i++) {
        double x0 = px[i] - x, y0 = py[i] - y;
        if ((-eps < x0 && x0 < eps) && (-eps < y0 && y0 < eps)) return true;
        if ((y0 > 0) != (y1 > 0)) { if (((x0 * y1 - x1 * y0) / (y1 - y0)) > 0) rc++; }
        if ((y0 < 0) != (y1 < 0)) { if (((x0 * y1 - x1 * y0) / (y1 - y0)) < 0) lc++; }
        x1 = x0; y1 = y0;
    }
    if ((rc & 1) != (lc & 1)) return true;
    return (rc & 1) != 0;
}

// OccupancyGrid.compute_occ_img_walls (occupancy_map.py:67-94) and global_goal_point_dist_transform
// (:435-485) for the maze of `c`: wall[H*W] (1 = wall), dist_norm (walls = 1.0), dist_raw (hop count + 1,
// 0 = unreachable or wall).  H = int(length / (1/scale)), W = int(width / (1/scale)).
inline void build_maze_maps(const MazeConfig& c, int* H_out, int* W_out, std::vector<uint8_t>& wall, std::vector<double>& dist_norm,
                            std::vector<double>& dist_raw) {
    const double grid = 1.0 / c.m_to_pix_scale;
    const int W = (int)(c.width / grid), H = (int)(c.length / grid);
    *H_out = H; *W_out = W;
    wall.assign((size_t)H * W, 0); dist_raw.assign((size_t)H * W, 0.0); dist_norm.assign((size_t)H * W, 0.0);
    const double scale = (double)H / c.length;         // meter_to_pixel_scale = height / self.map_height
    const double wr = 0.5;
    for (int w = 0; w < c.n_walls; w++) {
        double ax = c.walls[4 * w], ay = c.walls[4 * w + 1], bx = c.walls[4 * w + 2], by = c.walls[4 * w + 3];
        double dx = bx - ax, dy = by - ay;
        double len = std::sqrt(dx * dx + dy * dy);
        double ux = dx / len, uy = dy / len;
        double qx = -uy, qy = ux;
        double vx[4] = {(ax + wr * qx) - wr * ux, (ax - wr * qx) - wr * ux, (bx - wr * qx) + wr * ux, (bx + wr * qx) + wr * ux};
        double vy[4] = {(ay + wr * qy) - wr * uy, (ay - wr * qy) - wr * uy, (by - wr * qy) + wr * uy, (by + wr * qy) + wr * uy};
        double rmin = INFINITY, rmax = -INFINITY, cmin = INFINITY, cmax = -INFINITY;
        for (int i = 0; i < 4; i++) {
            vx[i] *= scale; vy[i] *= scale;
            cmin = std::fmin(cmin, vx[i]); cmax = std::fmax(cmax, vx[i]); rmin = std::fmin(rmin, vy[i]); rmax = std::fmax(rmax, vy[i]);
        }
        long long minr = (long long)std::fmax(0.0, rmin), maxr = (long long)std::ceil(rmax);
        long long minc = (long long)std::fmax(0.0, cmin), maxc = (long long)std::ceil(cmax);
        if (maxr > H - 1) maxr = H - 1;
        if (maxc > W - 1) maxc = W - 1;
        for (long long r = minr; r <= maxr; r++)
            for (long long q = minc; q <= maxc; q++)
                if (host_pip(vx, vy, 4, (double)q, (double)r)) wall[(size_t)r * W + q] = 1;
    }
    const double g2m = c.width / W;
    int gx = (int)(c.goal_x / g2m), gy = (int)(c.goal_y / g2m);
    if (gx < 0 || gx >= W || gy < 0 || gy >= H) return;
    std::vector<int> frontier, next;
    std::vector<uint8_t> visited((size_t)H * W, 0);
    dist_raw[(size_t)gy * W + gx] = 1.0; visited[(size_t)gy * W + gx] = 1;
    frontier.push_back(gy * W + gx);
    const int dyx[8][2] = {{0, 1}, {0, -1}, {1, 0}, {-1, 0}, {1, 1}, {1, -1}, {-1, 1}, {-1, -1}};
    double mx = 1.0;
    while (!frontier.empty()) {
        next.clear();
        for (int cell : frontier) {
            int y = cell / W, x = cell % W;
            for (int k = 0; k < 8; k++) {
                int ny = y + dyx[k][0], nx = x + dyx[k][1];
                if (ny < 0 || ny >= H || nx < 0 || nx >= W) continue;
                size_t ni = (size_t)ny * W + nx;
                if (visited[ni] || wall[ni]) continue;
                visited[ni] = 1;
                dist_raw[ni] = dist_raw[cell] + 1;
                mx = std::fmax(mx, dist_raw[ni]);
                next.push_back((int)ni);
            }
        }
        frontier.swap(next);
    }
    for (size_t i = 0; i < dist_raw.size(); i++) dist_norm[i] = wall[i] ? 1.0 : dist_raw[i] / mx;
}

inline SimParams make_maze_params(const MazeConfig& c) {
    ShipIceConfig sc;
    sc.dt = c.dt; sc.steps = c.steps; sc.iterations = c.iterations; sc.damping = c.damping; sc.target_speed = c.target_speed;
    sc.map_width = c.width; sc.map_height = c.length; sc.m_to_pix_scale = c.m_to_pix_scale; sc.settle_steps = c.settle_steps;
    sc.margin_body = c.margin_body; sc.margin_robot = c.margin_robot;
    SimParams P = make_params(sc);
    P.kind = KIND_MAZE; P.n_walls = c.n_walls;
    P.yaw_step = (M_PI / 2) / 15;                 // maze_NAMO_env.py:93
    P.beta = 1.5; P.k_increment = 150.0;          // :80-82
    P.goal_x = c.goal_x; P.goal_gy = c.goal_y; P.goal_reach = c.goal_radius + c.min_r;
    P.goal_y = 0.0;
    return P;
}

}  // namespace bp
