// box-delivery-v0: block-level work items — shortest-path distances / robot paths on the configuration
// space, and the 224 x 224 x 4 observation.
//
// Follows BoxDeliveryEnv.shortest_path_distance (box_delivery_env.py:1266-1268, called per box at :651-655
// and :736-740), shortest_path (:1209-1264), generate_observation (:1045-1059), update_global_overhead_map
// (:1177-1207), get_local_map (:1078-1096), get_local_distance_map (:1098-1101),
// create_global_shortest_path_map (:1131-1138), create_global_shortest_path_to_receptacle_map (:1115-1129).
// Third-party primitives restated (pinned against the libraries by tests/test_oracle_boxdel_cpu.py through
// their numpy twins in oracle/box_delivery_oracle.py):
//   * cv2.fillPoly on int32 vertices: outline with the 8-connected LineIterator (left-to-right), interior
//     rows y in [y_min, y_max) spanning [ceil(x_left), floor(x_right)] of the 16.16 fixed-point edges;
//   * scipy.ndimage.rotate(order=0, reshape=True): cosdg / sindg, output shape from the rotated corners,
//     offset = in_centre - M @ out_centre, nearest sample floor(c + 0.5), outside [0, n-1] -> 0.
// Only the 224 x 224 window is ever produced: every global map is a pure function of the pixel.
//
// One thread block per item; the spfa distance field lives in a per-block scratch arena (L2-resident),
// never in a per-map output array: 2 N + 2 fields per env step would be 40 MB per env otherwise.
#pragma once
#include "boxdel_env.cuh"
#include "maze_raster.cuh"    // mz_sincosdg, mz_fma

namespace bp {
namespace bd {

// scratch layout per block (ints): dist[nv] q0[nv] q1[nv] mark[nv] small[SMALL*SMALL/4 + 1]
// exact-parents mode adds parents[nv] behind it (the serial solver's ring and flags live in q0 / q1 / mark until the path is extracted)
BP_HD size_t bd_scratch_ints(const BdParams& P) { return (size_t)(P.exact_parents ? 5 : 4) * P.H * P.W + (size_t)P.SMALL * P.SMALL / 4 + 64; }

// shortest_path_distance(box k, receptacle) is a function of the box origin and the room alone.  A box that has not moved since
// its last solve reuses the value: every box between the end of a step and the start of the next (final_dist of step t is
// init_dist of step t + 1), and the boxes the robot did not touch within a step.  Bitwise the same result as solving again.
BP_DEV bool bd_box_dist_cached(BdEnv* E, int k, int phase, bool write) {
    BdEnvCold* K = E->cold;
    if (!K->dist_ok[k]) return false;
    V2 o = bd_origin(E->body[1 + k]);
    if (o.x != K->dist_key[k][0] || o.y != K->dist_key[k][1]) return false;
    if (write) { if (phase == 0) K->init_dist[k] = K->dist_val[k]; else K->final_dist[k] = K->dist_val[k]; }
    return true;
}

// ------------------------------------------------------------------ distance / path items
// k < NB: shortest_path_distance(box k, receptacle) -> E->cold->init_dist[k] (phase 0) or E->cold->final_dist[k] (phase 1);
// k == NB: the robot's waypoints when the straight line was blocked (phase 0 only).
BP_DEV void bd_dist_item(SpfaShared& S, const BdConst& C, BdEnv* E, int k, int phase, int* scratch) {
    const BdParams& P = C.P;
    const int nv = P.H * P.W;
    int* dist = scratch; int* q0 = scratch + nv; int* q1 = q0 + nv; int* mark = q1 + nv;
    double sx, sy, gx, gy;
    int si, sj, ti, tj;
    if (k < P.NB) {
        const BdBody& B = E->body[1 + k];
        if (!B.alive || P.recept_visible) return;
        if (bd_box_dist_cached(E, k, phase, BP_TID == 0)) return;
        V2 o = bd_origin(B);
        sx = o.x; sy = o.y; gx = P.recept_x; gy = P.recept_y;
        int pi, pj;
        pos_to_pixel(P, sx, sy, &pi, &pj); closest_valid(C, E->map_slot, pi, pj, &si, &sj);
        pos_to_pixel(P, gx, gy, &pi, &pj); closest_valid(C, E->map_slot, pi, pj, &ti, &tj);
    } else {
        if (phase != 0 || !E->need_spfa) return;
        sx = E->init_pos[0]; sy = E->init_pos[1]; gx = E->tgt[0]; gy = E->tgt[1];
        si = E->src_i; sj = E->src_j; ti = E->tgt_i; tj = E->tgt_j;
    }
    int* parents = nullptr;
    if (P.exact_parents) {
        // the reference's own parent order (ties between equal-cost predecessors decide which staircase the path takes, and the
        // Douglas-Peucker / line-of-sight pruning downstream amplifies that): one thread restates its queue
        parents = scratch + (size_t)4 * nv + (size_t)P.SMALL * P.SMALL / 4 + 64;
        // (labels in a shared-memory window of the room were measured slower: the solve is bound by the dependent instruction
        // chain of one thread, and the shared memory costs resident blocks)
        spfa_serial_init(bd_cspace(C, E->map_slot), nv, (float*)dist, parents, (unsigned char*)mark, BP_TID, BP_NT);
        BP_BSYNC();
        if (BP_TID == 0) {
            if (P.border_blocked) spfa_serial_run<true>(bd_cspace(C, E->map_slot), P.H, P.W, si, sj, (float*)dist, parents, q0, (unsigned char*)mark);
            else spfa_serial_run<false>(bd_cspace(C, E->map_slot), P.H, P.W, si, sj, (float*)dist, parents, q0, (unsigned char*)mark);
        }
    } else {
        spfa_relax(S, bd_cspace(C, E->map_slot), P.H, P.W, si, sj, (float*)dist, q0, q1, mark);
    }
    BP_BSYNC();
    if (BP_TID == 0) {
        double wp[MAXWP][3];
        int n = bd_extract_path(C, bd_cspace(C, E->map_slot), parents, dist, si, sj, ti, tj, sx, sy, gx, gy, q0, q1, mark, wp, MAXWP);
        if (n == 0) { E->status |= BD_PATH_OVERFLOW; n = 2; wp[0][0] = sx; wp[0][1] = sy; wp[1][0] = gx; wp[1][1] = gy; }
        if (k < P.NB) {
            double d = 0.0;
            for (int i = 1; i < n; i++) d = d + dist2(wp[i - 1][0], wp[i - 1][1], wp[i][0], wp[i][1]);
            if (phase == 0) E->cold->init_dist[k] = d; else E->cold->final_dist[k] = d;
            E->cold->dist_key[k][0] = sx; E->cold->dist_key[k][1] = sy; E->cold->dist_val[k] = d; E->cold->dist_ok[k] = 1;
        } else {
            for (int i = 0; i < n; i++) { E->cold->path[i][0] = wp[i][0]; E->cold->path[i][1] = wp[i][1]; }
            E->n_path = n;
        }
    }
    BP_BSYNC();
}

// 1 when free cell v is visible from the receptacle cell (or v is not free)
BP_DEV int bd_cell_visible(const BdConst& C, int room, int v) {
    const BdParams& P = C.P;
    const uint8_t* cs = bd_cspace(C, room);
    if (!cs[v]) return 1;
    return line_free(cs, P.W, C.rooms[room].recept_ci, C.rooms[room].recept_cj, v / P.W, v % P.W) ? 1 : 0;
}

// ------------------------------------------------------------------ cv2.fillPoly (one convex polygon, <= 4 int vertices)
BP_DEV void bd_cv_line(unsigned char* img, int N, int x0, int y0, int x1, int y1, unsigned char code) {
    int dx = x1 - x0, dy = y1 - y0, sy = 1;
    if (dx < 0) { dx = -dx; dy = -dy; x0 = x1; y0 = y1; }
    if (dy < 0) { dy = -dy; sy = -1; }
    bool vert = dy > dx;
    int major = vert ? dy : dx, minor = vert ? dx : dy;
    int err = major - (minor + minor);
    int x = x0, y = y0;
    for (int i = 0; i <= major; i++) {
        if (x >= 0 && x < N && y >= 0 && y < N) img[y * N + x] = code;
        bool ng = err < 0;
        err += -(minor + minor) + (ng ? (major + major) : 0);
        if (vert) { y += sy; if (ng) x += 1; } else { x += 1; if (ng) y += sy; }
    }
}
BP_DEV void bd_fill_poly(unsigned char* img, int N, const int* px, const int* py, int n, unsigned char code) {
    const int tid = BP_TID, nt = BP_NT;
    for (int e = tid; e < n; e += nt) { int a = (e + n - 1) % n; bd_cv_line(img, N, px[a], py[a], px[e], py[e], code); }
    int ymin = py[0], ymax = py[0];
    for (int i = 1; i < n; i++) { ymin = py[i] < ymin ? py[i] : ymin; ymax = py[i] > ymax ? py[i] : ymax; }
    if (ymax > N) ymax = N;
    for (int y = ymin + tid; y < ymax; y += nt) {
        if (y < 0) continue;
        long long xs[SV]; int na = 0;
        for (int e = 0; e < n; e++) {
            int a = (e + n - 1) % n;
            int ax = px[a], ay = py[a], bx = px[e], by = py[e];
            if (ay == by) continue;
            long long num = (long long)(bx - ax) * 65536, den = by - ay;
            long long dxe = num / den;                 // C division truncates toward zero, as OpenCV's int64 division
            int y0 = ay < by ? ay : by, y1 = ay < by ? by : ay;
            long long x0 = (long long)(ay < by ? ax : bx) * 65536;
            if (y0 <= y && y < y1) xs[na++] = x0 + (long long)(y - y0) * dxe;
        }
        for (int i = 1; i < na; i++) { long long t = xs[i]; int j = i - 1; while (j >= 0 && xs[j] > t) { xs[j + 1] = xs[j]; j--; } xs[j + 1] = t; }
        for (int q = 0; q + 1 < na; q += 2) {
            long long x1 = (xs[q] + 65535) >> 16, x2 = xs[q + 1] >> 16;
            if (x1 < N && x2 >= 0) {
                if (x1 < 0) x1 = 0;
                if (x2 > N - 1) x2 = N - 1;
                for (long long x = x1; x <= x2; x++) img[y * N + (int)x] = code;
            }
        }
    }
    BP_BSYNC();
}

BP_DEV float bd_scale_sp(const BdParams& P, float d) {
    float a = d / (float)P.ppm;
    float b = (float)((double)a / P.sp_div);
    return b * (float)P.sp_scale;
}

struct BdObsShared { SpfaShared sp; int minbits[2]; int px[SV], py[SV]; };

// One env's observation: u8 [LW][LW][4] (HWC, as the reference stacks on axis 2).
BP_DEV void bd_obs_item(BdObsShared& S, const BdConst& C, const BdEnv* E, int* scratch, uint8_t* out) {
    const BdParams& P = C.P;
    const BdLayout& L = C.layouts[E->layout];
    const int tid = BP_TID, nt = BP_NT;
    const int nv = P.H * P.W;
    int* dist = scratch; int* q0 = scratch + nv; int* q1 = q0 + nv; int* mark = q1 + nv;
    unsigned char* small = (unsigned char*)(mark + nv);
    const int N = P.SMALL;
    const BdBody& R = E->body[0];
    // ---- robot-source distance field (create_global_shortest_path_map)
    {
        int pi, pj, si, sj;
        pos_to_pixel(P, R.px, R.py, &pi, &pj); closest_valid(C, E->map_slot, pi, pj, &si, &sj);
        spfa_relax(S.sp, bd_cspace(C, E->map_slot), P.H, P.W, si, sj, (float*)dist, q0, q1, mark);
        BP_BSYNC();
    }
    // ---- small overhead map: floor / obstacle, receptacle, boxes, robot chassis (update_global_overhead_map)
    { const uint8_t* sf = bd_small_free(C, E->map_slot); for (int i = tid; i < N * N; i += nt) small[i] = sf[i]; }        // static codes: 0 obstacle, 1 floor (3 goal area in area-clearing)
    BP_BSYNC();
    auto to_px = [&](double wx, double wy, int* x, int* y) {
        int ix = (int)(wx * P.ppm), iy = (int)(wy * P.ppm);       // astype(np.int32): truncation toward zero
        *x = ix + P.px_off; *y = N - (iy + P.px_off);
    };
    const int n_polys = 1 + P.NB + 1;
    for (int q = 0; q < n_polys; q++) {
        int n = 0; unsigned char code = 0;
        bool skip = false;
        if (tid == 0) {
            if (q == 0) {
                if (P.draw_recept) {
                    const BdStatics& T = bd_room(C, E);
                    n = T.st_nv[P.recept]; code = 3;
                    for (int v = 0; v < n; v++) to_px(T.st_v[P.recept][v][0], T.st_v[P.recept][v][1], &S.px[v], &S.py[v]);
                }
            } else if (q <= P.NB) {
                const BdBody& B = E->body[q];
                if (B.alive == 1) {
                    n = SV; code = (unsigned char)(E->completed[q - 1] ? P.box_done_code : P.box_code);
                    for (int v = 0; v < SV; v++) { V2 w = bd_local_to_world(B, L.box_lv[q - 1][v][0], L.box_lv[q - 1][v][1]); to_px(w.x, w.y, &S.px[v], &S.py[v]); }
                }
            } else {
                n = C.S.n_foot; code = (unsigned char)P.robot_code;
                for (int v = 0; v < n; v++) { V2 w = bd_local_to_world(R, C.S.rb_foot[v][0], C.S.rb_foot[v][1]); to_px(w.x, w.y, &S.px[v], &S.py[v]); }
            }
            S.minbits[0] = n; S.minbits[1] = code;
        }
        BP_BSYNC();
        n = S.minbits[0]; code = (unsigned char)S.minbits[1];
        skip = (n == 0);
        int px[SV], py[SV];
        for (int v = 0; v < SV; v++) { px[v] = S.px[v]; py[v] = S.py[v]; }
        BP_BSYNC();
        if (!skip) bd_fill_poly(small, N, px, py, n, code);
    }
    // ---- rotation geometry (get_local_map: crop CROP x CROP around the robot, rotate(90 - degrees(heading), order=0))
    const double deg = 90.0 - R.pa * (180.0 / BD_PI);
    double sd, cd; mz_sincosdg(deg, &sd, &cd);
    const double m00 = cd, m01 = sd, m10 = -sd, m11 = cd;
    const double iy = P.CROP, ix = P.CROP;
    double b0[4] = {0.0, m01 * ix, m00 * iy, m00 * iy + m01 * ix};
    double b1[4] = {0.0, m11 * ix, m10 * iy, m10 * iy + m11 * ix};
    double mn0 = b0[0], mx0 = b0[0], mn1 = b1[0], mx1 = b1[0];
    for (int i = 1; i < 4; i++) { mn0 = fmin(mn0, b0[i]); mx0 = fmax(mx0, b0[i]); mn1 = fmin(mn1, b1[i]); mx1 = fmax(mx1, b1[i]); }
    const int OH = (int)((mx0 - mn0) + 0.5), OW = (int)((mx1 - mn1) + 0.5);
    const double oc0 = (OH - 1) / 2.0, oc1 = (OW - 1) / 2.0, ic = (P.CROP - 1) / 2.0;
    const double off0 = ic - mz_fma(m00, oc0, m01 * oc1);
    const double off1 = ic - mz_fma(m10, oc0, m11 * oc1);
    const int r_start = OH / 2 - P.LW / 2, c_start = OW / 2 - P.LW / 2;
    const int pi = (int)floor(-R.py * P.ppm + P.H / 2.0), pj = (int)floor(R.px * P.ppm + P.W / 2.0);
    const int gi0 = pi - P.CROP / 2, gj0 = pj - P.CROP / 2;
    const float inf = 2.0f * (float)nv;
    // sample: global pixel of output (r, c) or -1
    auto src = [&](int r, int c) -> int {
        double i = (double)(r + r_start), j = (double)(c + c_start);
        double cc0 = (off0 + i * m00) + j * m01;
        double cc1 = (off1 + i * m10) + j * m11;
        if (cc0 < 0 || cc0 > P.CROP - 1 || cc1 < 0 || cc1 > P.CROP - 1) return -1;
        int gi = gi0 + (int)floor(cc0 + 0.5), gj = gj0 + (int)floor(cc1 + 0.5);
        if (gi < 0 || gj < 0 || gi >= P.H || gj >= P.W) return -1;
        return gi * P.W + gj;
    };
    auto ch2 = [&](int g) -> float { if (g < 0) return 0.0f; float d = i2f(dist[g]); return bd_scale_sp(P, d < inf ? d : 0.0f); };
    const float* ch3m = bd_ch3_map(C, E->map_slot);
    auto ch3 = [&](int g) -> float { if (g < 0) return 0.0f; return ch3m[g]; };
    // ---- pass 1: minima of the two distance channels over the window
    if (tid == 0) { S.minbits[0] = f2i(INFINITY); S.minbits[1] = f2i(INFINITY); }
    BP_BSYNC();
    {
        float m2 = INFINITY, m3 = INFINITY;
        for (int p = tid; p < P.LW * P.LW; p += nt) {
            int g = src(p / P.LW, p % P.LW);
            m2 = fminf(m2, ch2(g)); m3 = fminf(m3, ch3(g));
        }
        spfa_atomic_min(&S.minbits[0], f2i(m2)); spfa_atomic_min(&S.minbits[1], f2i(m3));   // non-negative floats: order-preserving bits
    }
    BP_BSYNC();
    const float min2 = i2f(S.minbits[0]), min3 = i2f(S.minbits[1]);
    // ---- pass 2: compose and store (4 bytes per pixel, coalesced)
    for (int p = tid; p < P.LW * P.LW; p += nt) {
        int g = src(p / P.LW, p % P.LW);
        unsigned char c0 = 0;
        if (g >= 0) {
            int gi = g / P.W, gj = g - gi * P.W;
            int li = gi - P.small_i0, lj = gj - P.small_j0;
            if (li >= 0 && lj >= 0 && li < N && lj < N) c0 = (unsigned char)(((float)small[li * N + lj] / 8.0f) * 255.0f);
        }
        unsigned char c1 = C.robot_chan[p];
        unsigned char c2 = (unsigned char)((ch2(g) - min2) * 255.0f);
        unsigned char c3 = (unsigned char)((ch3(g) - min3) * 255.0f);
        uint32_t w = (uint32_t)c0 | ((uint32_t)c1 << 8) | ((uint32_t)c2 << 16) | ((uint32_t)c3 << 24);
        ((uint32_t*)out)[p] = w;
    }
    BP_BSYNC();
}

}  // namespace bd
}  // namespace bp
