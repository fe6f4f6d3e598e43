// On-device room build for box-delivery's column / divider configurations.
//
// Replaces, per env and per reset, BoxDeliveryEnv.update_configuration_space (box_delivery_env.py:1140-1175) and
// create_global_shortest_path_to_receptacle_map (:1115-1138): the reference redraws the random columns / the divider at every
// reset (:358-471) and rebuilds
//     small_obstacle_map      cv2.fillPoly of every wall / column / divider / corner polygon into a (LW + 20)^2 image
//     obstacle_map            a padded room of ones with the small map pasted into its centre
//     configuration_space     1 - binary_dilation(obstacle_map, disk(floor(robot_radius * ppm)))      (skimage disk, scipy dilation)
//     configuration_space_thin  the same with disk(floor(robot_half_width * ppm))
//     closest_cspace_indices  scipy.ndimage.distance_transform_edt(1 - configuration_space, return_indices=True)
// with library calls.  Here one thread block builds the five maps of an env into that env's own map slot:
//   * fillPoly: bd_fill_poly (boxdel_obs.cuh, the OpenCV scan-line rule) on the pixel polygons the host computed per layout
//     exactly as the reference does (world vertices * ppm -> int32, offset, y flip): BdStatics::st_px / st_py;
//   * dilation by a disk of radius r == "some obstacle pixel within squared distance r^2" (skimage's disk is x^2 + y^2 <= r^2,
//     scipy's border value is 0): a column pass (distance to the nearest obstacle pixel above / below, capped) and a row pass
//     over |dj| <= r;
//   * the feature transform: column pass (nearest free cell above / below, ties -> the upper one) and a row pass that minimises
//     g(i, j')^2 + (j - j')^2 over the columns j' that hold free cells, taking the FIRST minimum.  That is scipy's result
//     including its tie-breaks — smallest column, then smallest row — checked against scipy itself on random and on real
//     room maps (tests/test_kernel_logic_boxdel_cpu.py, tests/test_gpu_boxdel.py compare every map byte for byte with the
//     host build that calls cv2 / scipy).
//   * the receptacle-source distance field (fourth observation channel): spfa_relax on the new configuration space, scaled in
//     the reference's mixed fp32 / fp64 arithmetic.
#pragma once
#include "boxdel_obs.cuh"

namespace bp {
namespace bd {

// squared-distance test of the dilation: blocked when an obstacle pixel lies within r2 of (i, j).  gcol[(i, j)] = vertical distance
// of (i, j) to the nearest obstacle pixel of its column (0 on obstacle pixels and everywhere outside the small window).
BP_DEV bool bd_dilated(const int* gcol, int H, int W, int i, int j, int r, int r2) {
    for (int dj = -r; dj <= r; dj++) {
        int jj = j + dj;
        if (jj < 0 || jj >= W) continue;              // outside the image: not an obstacle (border value 0)
        int g = gcol[(size_t)i * W + jj];
        if (g * g + dj * dj <= r2) return true;
    }
    return false;
}

// scipy.ndimage.distance_transform_edt(1 - free, return_distances=False, return_indices=True) for a 2-D map: for every cell the
// (row, column) of the nearest free cell, a free cell mapping to itself; among equally near free cells the one in the smallest
// column, then the smallest row (scipy's own outcome; tests/test_kernel_logic_boxdel_cpu.py holds this against scipy on random
// maps).  Block-cooperative; gcol / grow: H * W ints of scratch each; sh_minmax: 2 ints shared by the block.
// out[0][H][W] = rows, out[1][H][W] = columns.
BP_DEV void bd_feature_transform(const uint8_t* cs, int H, int W, int32_t* closest, int* gcol, int* grow, int* sh_minmax) {
    const int tid = BP_TID, nt = BP_NT;
    const size_t nv = (size_t)H * W;
    const int BIG = 1 << 20;
    if (tid == 0) { sh_minmax[0] = W; sh_minmax[1] = -1; }
    BP_BSYNC();
    for (int j = tid; j < W; j += nt) {
        int last = -BIG; bool any = false;
        for (int i = 0; i < H; i++) {
            if (cs[(size_t)i * W + j]) { last = i; any = true; }
            gcol[(size_t)i * W + j] = i - last; grow[(size_t)i * W + j] = last;          // nearest free cell above (or far away)
        }
        last = BIG;
        for (int i = H - 1; i >= 0; i--) {
            if (cs[(size_t)i * W + j]) last = i;
            if (last - i < gcol[(size_t)i * W + j]) { gcol[(size_t)i * W + j] = last - i; grow[(size_t)i * W + j] = last; }   // strictly nearer: ties keep the upper one
        }
        if (any) { spfa_atomic_min(&sh_minmax[0], j); spfa_atomic_max(&sh_minmax[1], j); }
    }
    BP_BSYNC();
    {
        const int c0 = sh_minmax[0], c1 = sh_minmax[1];
        for (size_t v = tid; v < nv; v += nt) {
            const int i = (int)(v / W), j = (int)(v - (size_t)i * W);
            int bj = j, bi = i;
            if (!cs[v] && c1 >= c0) {
                const int g = gcol[v];
                const int R = g < BIG / 2 ? g : W;                         // a better column is at most g away
                int lo = j - R < c0 ? c0 : j - R, hi = j + R > c1 ? c1 : j + R;
                long long best = -1;
                for (int jj = lo; jj <= hi; jj++) {
                    const int gg = gcol[(size_t)i * W + jj];
                    if (gg >= BIG / 2) continue;
                    const long long cost = (long long)gg * gg + (long long)(j - jj) * (j - jj);
                    if (best < 0 || cost < best) { best = cost; bj = jj; }
                }
                bi = grow[(size_t)i * W + bj];
            }
            closest[v] = bi; closest[nv + v] = bj;
        }
    }
    BP_BSYNC();
}

// Block-cooperative.  scratch: >= 4 * H * W ints (the per-block scratch of k_bd_dist / k_bd_obs).  Writes the maps of `slot`.
BP_DEV void bd_build_room_maps(SpfaShared& S, const BdConst& C, const BdStatics& T, int slot, int* scratch, int* sh_minmax) {
    const BdParams& P = C.P;
    const int tid = BP_TID, nt = BP_NT;
    const int H = P.H, W = P.W, N = P.SMALL;
    const size_t nv = (size_t)H * W;
    uint8_t* cs = const_cast<uint8_t*>(bd_cspace(C, slot));
    uint8_t* thin = const_cast<uint8_t*>(bd_cspace_thin(C, slot));
    int32_t* closest = const_cast<int32_t*>(bd_closest(C, slot));
    uint8_t* sm = const_cast<uint8_t*>(bd_small_free(C, slot));
    float* ch3 = const_cast<float*>(bd_ch3_map(C, slot));
    int* gcol = scratch;                      // [nv] column distance (obstacle pass, then free-cell pass)
    int* grow = scratch + nv;                 // [nv] row index of the nearest free cell of the column
    // ---- 1. small obstacle map: 1 = wall / column / divider / corner pixel
    for (int i = tid; i < N * N; i += nt) sm[i] = 0;
    BP_BSYNC();
    for (int k = 0; k < P.n_static; k++) {
        if (T.st_npx[k] <= 0) continue;
        bd_fill_poly(sm, N, T.st_px[k], T.st_py[k], T.st_npx[k], 1);      // (ends with a block barrier)
    }
    // ---- 2. column pass of the dilation; outside the small window every pixel is an obstacle (padded room of ones)
    const int rmax = (int)floor(P.robot_radius * P.ppm);                       // the wider of the two disks
    const int i0 = P.small_i0, j0 = P.small_j0;
    for (size_t v = tid; v < nv; v += nt) gcol[v] = 0;
    BP_BSYNC();
    for (int c = tid; c < N; c += nt) {
        const int j = j0 + c;
        if (j < 0 || j >= W) continue;
        int d = 0;                                                             // the row above the window is an obstacle row
        for (int r = 0; r < N; r++) {
            const int i = i0 + r;
            d = sm[r * N + c] ? 0 : (d < rmax + 1 ? d + 1 : rmax + 1);
            if (i >= 0 && i < H) gcol[(size_t)i * W + j] = d;
        }
        d = 0;
        for (int r = N - 1; r >= 0; r--) {
            const int i = i0 + r;
            d = sm[r * N + c] ? 0 : (d < rmax + 1 ? d + 1 : rmax + 1);
            if (i >= 0 && i < H) { int& g = gcol[(size_t)i * W + j]; if (d < g) g = d; }
        }
    }
    BP_BSYNC();
    // ---- 3. the two configuration spaces
    {
        const int r_thin = (int)floor(P.robot_half_width * P.ppm);
        for (size_t v = tid; v < nv; v += nt) {
            const int i = (int)(v / W), j = (int)(v - (size_t)i * W);
            cs[v] = bd_dilated(gcol, H, W, i, j, rmax, rmax * rmax) ? 0 : 1;
            thin[v] = bd_dilated(gcol, H, W, i, j, r_thin, r_thin * r_thin) ? 0 : 1;
        }
    }
    BP_BSYNC();
    // ---- 4. feature transform of the configuration space: nearest free cell, ties -> smallest column, then smallest row
    bd_feature_transform(cs, H, W, closest, gcol, grow, sh_minmax);
    // ---- 5. small_free = 1 - small obstacle map
    for (int i = tid; i < N * N; i += nt) sm[i] = sm[i] ? 0 : 1;
    BP_BSYNC();
    // ---- 6. receptacle-source distance field, scaled (create_global_shortest_path_to_receptacle_map)
    {
        int pi, pj, si, sj;
        pos_to_pixel(P, P.recept_x, P.recept_y, &pi, &pj);
        si = closest[(size_t)pi * W + pj]; sj = closest[nv + (size_t)pi * W + pj];
        int* dist = scratch; int* q0 = scratch + nv; int* q1 = q0 + nv; int* mark = q1 + nv;
        spfa_relax(S, cs, H, W, si, sj, (float*)dist, q0, q1, mark);
        BP_BSYNC();
        const float inf = 2.0f * (float)nv;
        for (size_t v = tid; v < nv; v += nt) {
            float d = i2f(dist[v]);
            d = (d < inf) ? d : 0.0f;                                     // the solver's output convention: unreachable -> 0
            float a = d / (float)P.ppm;
            float b = (float)((double)a / P.sp_div);
            ch3[v] = b * (float)P.sp_scale;
        }
    }
    BP_BSYNC();
}

}  // namespace bd
}  // namespace bp
