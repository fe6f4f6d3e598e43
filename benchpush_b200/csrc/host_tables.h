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
    std::vector<double> lv, ln, rad, rbound, minv, iinv, cog, mat_e, mat_u, init;
    std::vector<int> nv, nbody;
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
    H.mat_e.assign(LB, 0.0); H.mat_u.assign(LB, 0.0); H.init.assign(LB * 3, 0.0); H.nbody.assign(n_layouts, 0);
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
        H.nbody[l] = nf + 1;
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
    return P;
}

}  // namespace bp
