// Egocentric observation rasteriser for ship-ice-v0: one thread block per environment, the four
// 150x150 channels are composed in shared memory and written to HBM once, coalesced.
//
// Follows ShipIceEnv.generate_observation (ship_ice_env.py:378-409, egocentric branch) and the
// OccupancyGrid methods it calls (occupancy_map.py:37-65 compute_occ_img, :112-140
// ego_view_obstacle_map, :300-337 _compute_global_footprint, :379-410 ego_view_footprint, :413-433
// global_goal_dist_transform, :492-521 ego_view_goal_dist_transform, :524-554 global_orientation_map,
// :557-587 ego_view_orientation_map).  Third-party primitives restated bit-exactly:
//   * skimage.draw.polygon -> point_in_polygon of skimage/_shared/geometry.pyx (crossing test in
//     which vertices (|d| < 1e-12) and edges count as inside), evaluated in fp64 without FMA; the
//     quotient test (x0*y1 - x1*y0)/(y1 - y0) > 0 is evaluated as a sign test of numerator and
//     denominator, which is exact because an IEEE quotient of finite non-zero operands of this
//     magnitude cannot underflow to zero;
//   * cv2.line(thickness=1, LINE_8, shift=0) -> clipLine + the 8-connected LineIterator,
//     including the reference's uint16 cast of the end points (occupancy_map.py:546-547).
// Only the pixels of the 150x150 window are ever produced (the reference builds 1000x300 global
// maps and crops); the result is identical because every global map is a pure function of the pixel.
#pragma once
#include <math.h>
#include <stdint.h>
#include "sim_core.cuh"

#if defined(__CUDACC__) && !defined(BP_HOST_EMU)
#define BP_TID ((int)threadIdx.x)
#define BP_NT ((int)blockDim.x)
#define BP_BSYNC() __syncthreads()
#define BP_SMEM_ATOMIC_ADD(p, v) atomicAdd((p), (v))
#else
#define BP_TID 0
#define BP_NT 1
#define BP_BSYNC()
static inline int bp_emu_atomic_add(int* p, int v) { int o = *p; *p += v; return o; }
#define BP_SMEM_ATOMIC_ADD(p, v) bp_emu_atomic_add((p), (v))
#endif

namespace bp {

constexpr int OBS_WIN = 150;
constexpr int OBS_PIX = OBS_WIN * OBS_WIN;
constexpr int MAX_ROBOT_RAW = 32;
constexpr int MAX_CAND = 512;

struct ObsParams {
    int grid_h, grid_w;          // 1000, 300  (int(map/grid))
    double scale;                // m_to_pix_scale = occ_map_width / map_width = 25
    double vshift;               // local_window_v_shift = 2
    double goal_y;               // cfg.goal_y
    double g2m;                  // map_height / occ_map_height = 0.04
    double range_x, range_y;     // local_range = [12, 12]
    int n_robot_raw;
    double robot_raw[2 * MAX_ROBOT_RAW];
    double head[2], tail[2];
};

// skimage point_in_polygon on vertices (px[i], py[i]) = (vx[2i]*sx, vx[2i+1]*sy); != 0 means inside.
template <typename VT>
BP_DEV bool pip_inclusive(const VT* v, int n, double scale, double x, double y) {
    const double eps = 1e-12;
    unsigned r_cross = 0, l_cross = 0;
    double x1 = v[2 * (n - 1)] * scale - x;
    double y1 = v[2 * (n - 1) + 1] * scale - y;
    for (int i = 0; i < n; i++) {
        double x0 = v[2 * i] * scale - x;
        double y0 = v[2 * i + 1] * scale - y;
        if ((-eps < x0 && x0 < eps) && (-eps < y0 && y0 < eps)) return true;
        bool sr = (y0 > 0) != (y1 > 0);
        bool sl = (y0 < 0) != (y1 < 0);
        if (sr || sl) {
            double num = x0 * y1 - x1 * y0;
            double den = y1 - y0;
            bool pos = (num > 0 && den > 0) || (num < 0 && den < 0);
            bool ngt = (num > 0 && den < 0) || (num < 0 && den > 0);
            if (sr && pos) r_cross++;
            if (sl && ngt) l_cross++;
        }
        x1 = x0; y1 = y0;
    }
    if ((r_cross & 1) != (l_cross & 1)) return true;
    return (r_cross & 1) != 0;
}

// cv::clipLine on 64-bit points
BP_DEV bool clip_line(long long W, long long H, long long& x1, long long& y1, long long& x2, long long& y2) {
    long long right = W - 1, bottom = H - 1;
    int c1 = (x1 < 0) + (x1 > right) * 2 + (y1 < 0) * 4 + (y1 > bottom) * 8;
    int c2 = (x2 < 0) + (x2 > right) * 2 + (y2 < 0) * 4 + (y2 > bottom) * 8;
    if ((c1 & c2) == 0 && (c1 | c2) != 0) {
        long long a;
        if (c1 & 12) {
            a = c1 < 8 ? 0 : bottom;
            x1 += (long long)((double)(a - y1) * (double)(x2 - x1) / (double)(y2 - y1));
            y1 = a;
            c1 = (x1 < 0) + (x1 > right) * 2;
        }
        if (c2 & 12) {
            a = c2 < 8 ? 0 : bottom;
            x2 += (long long)((double)(a - y2) * (double)(x2 - x1) / (double)(y2 - y1));
            y2 = a;
            c2 = (x2 < 0) + (x2 > right) * 2;
        }
        if ((c1 & c2) == 0 && (c1 | c2) != 0) {
            if (c1) {
                a = c1 == 1 ? 0 : right;
                y1 += (long long)((double)(a - x1) * (double)(y2 - y1) / (double)(x2 - x1));
                x1 = a; c1 = 0;
            }
            if (c2) {
                a = c2 == 1 ? 0 : right;
                y2 += (long long)((double)(a - x2) * (double)(y2 - y1) / (double)(x2 - x1));
                x2 = a; c2 = 0;
            }
        }
    }
    return (c1 | c2) == 0;
}

// numpy float64 -> uint16 cast as executed on x86-64 (truncate to integer, keep the low 16 bits)
BP_DEV long long cast_u16(double x) { return (long long)((unsigned short)(long long)x); }

// Draw the orientation line (value 127) and the head pixel (255) into the window-local map.
BP_DEV void draw_orientation(unsigned char* loc, const ObsParams& O, double px, double py, double c, double s, int gi0, int gj0) {
    double ms = -s;
    double hx = O.head[0] * c + O.head[1] * ms + px, hy = O.head[0] * s + O.head[1] * c + py;
    double tx = O.tail[0] * c + O.tail[1] * ms + px, ty = O.tail[0] * s + O.tail[1] * c + py;
    long long x1 = cast_u16(hx * O.scale), y1 = cast_u16(hy * O.scale);
    long long x2 = cast_u16(tx * O.scale), y2 = cast_u16(ty * O.scale);
    long long hpx = x1, hpy = y1;
    const long long W = O.grid_w, H = O.grid_h;
    bool ok = true;
    if (!(x1 >= 0 && x1 < W && x2 >= 0 && x2 < W && y1 >= 0 && y1 < H && y2 >= 0 && y2 < H)) ok = clip_line(W, H, x1, y1, x2, y2);
    if (ok) {
        long long dx = x2 - x1, dy = y2 - y1;
        long long sx = 1, sy = 1;
        if (dx < 0) { dx = -dx; dy = -dy; x1 = x2; y1 = y2; }
        if (dy < 0) { dy = -dy; sy = -1; }
        bool vert = dy > dx;
        if (vert) { long long t = dx; dx = dy; dy = t; }
        long long err = dx - 2 * dy, plusDelta = 2 * dx, minusDelta = -2 * dy;
        long long x = x1, y = y1;
        for (long long i = 0; i <= dx; i++) {
            long long li = y - gi0, lj = x - gj0;
            if (li >= 0 && li < OBS_WIN && lj >= 0 && lj < OBS_WIN) loc[li * OBS_WIN + lj] = 127;
            bool m = err < 0;
            err += minusDelta + (m ? plusDelta : 0);
            if (vert) { y += sy; if (m) x += sx; } else { x += sx; if (m) y += sy; }
        }
    }
    // np.clip(head_pix, [0,0], [W-1,H-1]); global_orientation[head_y, head_x] = 1.0
    if (hpx > W - 1) hpx = W - 1;
    if (hpy > H - 1) hpy = H - 1;
    long long li = hpy - gi0, lj = hpx - gj0;
    if (li >= 0 && li < OBS_WIN && lj >= 0 && lj < OBS_WIN) loc[li * OBS_WIN + lj] = 255;
}

// Shared-memory working set of one block.
struct ObsSmem {
    unsigned char occ[OBS_PIX];
    unsigned char ori[OBS_PIX];
    double fpv[2 * MAX_ROBOT_RAW];      // in-grid robot vertices in metres (scaled on use)
    int cand[MAX_CAND];
    int n_cand, n_fp, cand_overflow;
};

// Compose the observation of one env.  verts/nverts/cen describe the obstacles as world-space
// polygons: after a step these are the hull vertices wv (get_obs_from_poly), at reset the raw trial
// vertices.  poly_stride = doubles between consecutive polygons.
template <typename VT, typename CT>
BP_DEV void render_obs(ObsSmem& S, const ObsParams& O, int n_obst, const VT* verts, int poly_stride, const int* nverts,
                       const CT* cenx, const CT* ceny, int cen_stride, double rx, double ry, double ra,
                       unsigned char* out /* [4][150][150] */) {
    const int tid = BP_TID, nt = BP_NT;
    double s, c; bp_sincos(ra, &s, &c);
    // window origin (occupancy_map.py:121-129): int() truncates toward zero
    const int window_x = (int)(rx * O.scale);
    const int window_y = (int)((ry + O.vshift) * O.scale);
    const int gi0 = window_y - OBS_WIN / 2, gj0 = window_x - OBS_WIN / 2;
    for (int i = tid; i < OBS_PIX; i += nt) { S.occ[i] = 0; S.ori[i] = 0; }
    if (tid == 0) {
        S.n_cand = 0; S.cand_overflow = 0;
        // footprint vertices: all configured vertices, those outside the grid are dropped (occupancy_map.py:320-330)
        int k = 0; double ms = -s;
        for (int i = 0; i < O.n_robot_raw; i++) {
            double vx = O.robot_raw[2 * i], vy = O.robot_raw[2 * i + 1];
            double wx = vx * c + vy * ms + rx, wy = vx * s + vy * c + ry;
            double gx = wx * O.scale, gy = wy * O.scale;
            if (gy < 0 || gy >= O.grid_h || gx < 0 || gx >= O.grid_w) continue;
            S.fpv[2 * k] = wx; S.fpv[2 * k + 1] = wy; k++;
        }
        S.n_fp = k;
    }
    BP_BSYNC();
    if (tid == 0) draw_orientation(S.ori, O, rx, ry, c, s, gi0, gj0);
    // candidate obstacles: pass the local-range test and can touch the window
    const double wl = (double)gj0 - 2.0, wr = (double)(gj0 + OBS_WIN) + 2.0, wb = (double)gi0 - 2.0, wt = (double)(gi0 + OBS_WIN) + 2.0;
    for (int b = tid; b < n_obst; b += nt) {
        double cx = cenx[(size_t)b * cen_stride], cy = ceny[(size_t)b * cen_stride];
        if (fabs(rx - cx) > O.range_x || fabs(ry - cy) > O.range_y) continue;
        const VT* v = verts + (size_t)b * poly_stride;
        int n = nverts[b];
        double l = INFINITY, r = -INFINITY, bt = INFINITY, tp = -INFINITY;
        for (int i = 0; i < n; i++) {
            double x = v[2 * i] * O.scale, y = v[2 * i + 1] * O.scale;
            l = fmin(l, x); r = fmax(r, x); bt = fmin(bt, y); tp = fmax(tp, y);
        }
        if (r < wl || l > wr || tp < wb || bt > wt) continue;
        int k = BP_SMEM_ATOMIC_ADD(&S.n_cand, 1);
        if (k < MAX_CAND) S.cand[k] = b; else S.cand_overflow = 1;
    }
    BP_BSYNC();
    int nc = S.n_cand < MAX_CAND ? S.n_cand : MAX_CAND;
    for (int k = 0; k < nc; k++) {
        int b = S.cand[k];
        const VT* v = verts + (size_t)b * poly_stride;
        int n = nverts[b];
        double rmin = INFINITY, rmax = -INFINITY, cmin = INFINITY, cmax = -INFINITY;
        for (int i = 0; i < n; i++) {
            double x = v[2 * i] * O.scale, y = v[2 * i + 1] * O.scale;
            cmin = fmin(cmin, x); cmax = fmax(cmax, x); rmin = fmin(rmin, y); rmax = fmax(rmax, y);
        }
        // skimage _polygon bounding box, clipped to the global image, then to the window
        long long minr = (long long)fmax(0.0, rmin), maxr = (long long)ceil(rmax);
        long long minc = (long long)fmax(0.0, cmin), maxc = (long long)ceil(cmax);
        if (maxr > O.grid_h - 1) maxr = O.grid_h - 1;
        if (maxc > O.grid_w - 1) maxc = O.grid_w - 1;
        if (minr < gi0) minr = gi0;
        if (minc < gj0) minc = gj0;
        if (maxr > gi0 + OBS_WIN - 1) maxr = gi0 + OBS_WIN - 1;
        if (maxc > gj0 + OBS_WIN - 1) maxc = gj0 + OBS_WIN - 1;
        if (maxr < minr || maxc < minc) continue;
        int bw = (int)(maxc - minc + 1), bh = (int)(maxr - minr + 1);
        for (int i = tid; i < bw * bh; i += nt) {
            int rr = (int)minr + i / bw, cc = (int)minc + i % bw;
            if (pip_inclusive(v, n, O.scale, (double)cc, (double)rr)) S.occ[(rr - gi0) * OBS_WIN + (cc - gj0)] = 255;
        }
    }
    BP_BSYNC();
    // footprint bounding box (draw.polygon without a shape argument)
    long long fminr = 0, fmaxr = -1, fminc = 0, fmaxc = -1;
    if (S.n_fp > 0) {
        double rmin = INFINITY, rmax = -INFINITY, cmin = INFINITY, cmax = -INFINITY;
        for (int i = 0; i < S.n_fp; i++) {
            double x = S.fpv[2 * i] * O.scale, y = S.fpv[2 * i + 1] * O.scale;
            cmin = fmin(cmin, x); cmax = fmax(cmax, x); rmin = fmin(rmin, y); rmax = fmax(rmax, y);
        }
        fminr = (long long)fmax(0.0, rmin); fmaxr = (long long)ceil(rmax);
        fminc = (long long)fmax(0.0, cmin); fmaxc = (long long)ceil(cmax);
    }
    for (int i = tid; i < OBS_PIX; i += nt) {
        int li = i / OBS_WIN, lj = i % OBS_WIN;
        int gi = gi0 + li, gj = gj0 + lj;
        bool inb = !(gi < 0 || gi >= O.grid_h || gj < 0 || gj >= O.grid_w);
        unsigned char fp = 0, edt = 255, ori = 0, occ = 0;
        if (inb) {
            fp = 127;
            if (gi >= fminr && gi <= fmaxr && gj >= fminc && gj <= fmaxc && pip_inclusive(S.fpv, S.n_fp, O.scale, (double)gj, (double)gi)) fp = 255;
            double dgoal = O.goal_y - gi * O.g2m;
            if (dgoal < 0) dgoal = 0;
            edt = (unsigned char)(int)((dgoal / O.goal_y) * 255.0);
            ori = S.ori[i];
            occ = S.occ[i];
        }
        out[i] = fp; out[OBS_PIX + i] = edt; out[2 * OBS_PIX + i] = ori; out[3 * OBS_PIX + i] = occ;
    }
    BP_BSYNC();
}

// ------------------------------------------------------------------ global (non-egocentric) observation
// ShipIceEnv.generate_observation with cfg.egocentric_obs = false (ship_ice_env.py:397-409).  In that mode the env builds its
// OccupancyGrid with 0.2 m cells (ship_ice_env.py:96: 200 x 60 cells for the 40 x 12 m map), and the observation is two
// channels on that cell grid, (x * 255).astype(u8):
//   0  concentration: compute_occ_img over ALL obstacles (no local range; occupancy_map.py:37-65) at m_to_pix_scale
//      (1000 x 300 pixels, union of the polygon fills), then compute_con_gridmap (:97-109) = skimage block_reduce(np.mean)
//      over int(0.2 * scale) = 5 x 5 pixel blocks (zero padded when the image is not a multiple of the block);
//   1  footprint: compute_ship_footprint_planner (:253-296): the configured ship vertices scaled to cells, those outside
//      the cell grid dropped, draw.polygon without a shape argument, value 1.0; all zeros when no vertex is inside.
// The 1000 x 300 binary image lives in shared memory as a bit mask (37.5 KB): groups of 32 threads take one floe each and
// OR the pixels its inclusive crossing test accepts; a block-wide pass then counts the bits of every cell and writes the
// 2 x 200 x 60 bytes of the env once.
struct GlobalObsParams {
    int raw_h, raw_w;            // int(map_height * scale), int(map_width * scale): 1000, 300
    int blk_y, blk_x;            // int(0.2 * scale): 5, 5
    int gh, gw;                  // int(map_height / 0.2), int(map_width / 0.2): 200, 60
    double cell_scale;           // gw / map_width (= gh / map_height, checked on the host): 5.0
};
constexpr int OBS_RAW_WORDS = 9600;          // 307 200 bits >= 1000 x 300
struct ObsGlobalSmem { unsigned int bits[OBS_RAW_WORDS]; double fpv[2 * MAX_ROBOT_RAW]; int n_fp; };

#if defined(__CUDACC__) && !defined(BP_HOST_EMU)
constexpr int OBS_FILL_GROUP = 32;
#define BP_SMEM_ATOMIC_OR(p, v) atomicOr((p), (v))
#else
constexpr int OBS_FILL_GROUP = 1;
#define BP_SMEM_ATOMIC_OR(p, v) (*(p) |= (v))
#endif

// skimage _polygon bounding box of the scaled polygon, clipped to an H x W image; false when empty
template <typename VT>
BP_DEV bool polygon_bbox(const VT* v, int n, double scale, int H, int W, int* minr, int* minc, int* bw, int* bh) {
    double rmin = INFINITY, rmax = -INFINITY, cmin = INFINITY, cmax = -INFINITY;
    for (int i = 0; i < n; i++) {
        double x = v[2 * i] * scale, y = v[2 * i + 1] * scale;
        cmin = fmin(cmin, x); cmax = fmax(cmax, x); rmin = fmin(rmin, y); rmax = fmax(rmax, y);
    }
    long long r0 = (long long)fmax(0.0, rmin), r1 = (long long)ceil(rmax);
    long long c0 = (long long)fmax(0.0, cmin), c1 = (long long)ceil(cmax);
    if (r1 > H - 1) r1 = H - 1;
    if (c1 > W - 1) c1 = W - 1;
    if (r1 < r0 || c1 < c0) return false;
    *minr = (int)r0; *minc = (int)c0; *bw = (int)(c1 - c0 + 1); *bh = (int)(r1 - r0 + 1);
    return true;
}

template <typename VT>
BP_DEV void render_obs_global(ObsGlobalSmem& S, const ObsParams& O, const GlobalObsParams& G, int n_obst, const VT* verts,
                              int poly_stride, const int* nverts, double rx, double ry, double ra, unsigned char* out /* [2][gh][gw] */) {
    const int tid = BP_TID, nt = BP_NT;
    const int nwords = (G.raw_h * G.raw_w + 31) / 32;
    for (int i = tid; i < nwords; i += nt) S.bits[i] = 0u;
    if (tid == 0) {
        double s, c; bp_sincos(ra, &s, &c);
        int k = 0; double ms = -s;
        for (int i = 0; i < O.n_robot_raw; i++) {
            double vx = O.robot_raw[2 * i], vy = O.robot_raw[2 * i + 1];
            double wx = vx * c + vy * ms + rx, wy = vx * s + vy * c + ry;
            double gx = wx * G.cell_scale, gy = wy * G.cell_scale;
            if (gy < 0 || gy >= G.gh || gx < 0 || gx >= G.gw) continue;
            S.fpv[2 * k] = wx; S.fpv[2 * k + 1] = wy; k++;
        }
        S.n_fp = k;
    }
    BP_BSYNC();
    const int lane = tid % OBS_FILL_GROUP, grp = tid / OBS_FILL_GROUP, ngrp = nt / OBS_FILL_GROUP;
    for (int b = grp; b < n_obst; b += ngrp) {
        const VT* v = verts + (size_t)b * poly_stride;
        const int n = nverts[b];
        int minr, minc, bw, bh;
        if (!polygon_bbox(v, n, O.scale, G.raw_h, G.raw_w, &minr, &minc, &bw, &bh)) continue;
        for (int i = lane; i < bw * bh; i += OBS_FILL_GROUP) {
            int rr = minr + i / bw, cc = minc + i % bw;
            if (pip_inclusive(v, n, O.scale, (double)cc, (double)rr)) {
                int bit = rr * G.raw_w + cc;
                BP_SMEM_ATOMIC_OR(&S.bits[bit >> 5], 1u << (bit & 31));
            }
        }
    }
    BP_BSYNC();
    const int ncell = G.gh * G.gw;
    const double inv_n = (double)(G.blk_y * G.blk_x);
    for (int i = tid; i < ncell; i += nt) {
        int gi = i / G.gw, gj = i % G.gw;
        int cnt = 0;
        for (int y = gi * G.blk_y; y < (gi + 1) * G.blk_y && y < G.raw_h; y++)
            for (int x = gj * G.blk_x; x < (gj + 1) * G.blk_x && x < G.raw_w; x++) {
                int bit = y * G.raw_w + x;
                cnt += (S.bits[bit >> 5] >> (bit & 31)) & 1u;
            }
        out[i] = (unsigned char)(int)(((double)cnt / inv_n) * 255.0);
        out[ncell + i] = 0;
    }
    BP_BSYNC();
    int minr, minc, bw, bh;
    // draw.polygon(r, c) without a shape: the vertices are inside the cell grid, and a pixel beyond the last vertex row /
    // column cannot pass the test, so clipping the box to the grid changes nothing and protects the buffer
    if (S.n_fp > 0 && polygon_bbox(S.fpv, S.n_fp, G.cell_scale, G.gh, G.gw, &minr, &minc, &bw, &bh)) {
        for (int i = tid; i < bw * bh; i += nt) {
            int rr = minr + i / bw, cc = minc + i % bw;
            if (pip_inclusive(S.fpv, S.n_fp, G.cell_scale, (double)cc, (double)rr)) out[ncell + rr * G.gw + cc] = 255;
        }
    }
    BP_BSYNC();
}

}  // namespace bp
