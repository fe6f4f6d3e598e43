// Egocentric observation of maze-NAMO-v0: one thread block per environment.
//
// Follows MazeNAMO.generate_observation (maze_NAMO_env.py:514-525) and OccupancyGrid.ego_view_map_maze
// (occupancy_map.py:142-202): four global 240x240 maps (robot footprint, movable obstacles, walls,
// goal-point distance) are cropped to a 288x288 window centred on the robot's pixel, rotated by
// (heading - pi/2) with scipy.ndimage.rotate(reshape=False, order=1) and cut to the centre 192x192;
// channels [footprint, obstacles, walls, distance], (x * 255).astype(uint8).
//
// Third-party primitives restated bit-exactly (the oracle pins them against the scipy in the image):
//   * skimage.draw.polygon -> pip_inclusive (raster.cuh);
//   * scipy.special.cosdg / sindg (cephes sindg.c): octant reduction in degrees + polevl;
//   * scipy.ndimage.rotate -> affine_transform -> NI_GeometricTransform with order 1, mode 'constant':
//       M = [[c, s], [-s, c]], offset = centre - M @ centre where numpy evaluates the 2x2 @ 2 product as
//       fma(M[h,0], v0, M[h,1] * v1); coordinate h = (offset[h] + i * M[h,0]) + j * M[h,1]; outside
//       [0, len-1] -> cval; weights w0 = 1 - (x - floor x), w1 = 1 - w0; value = sum over the 2x2 support
//       in row-major order of (v * w0[a]) * w1[b], accumulated from 0.
// The three binary channels live as bits of one byte per crop pixel in shared memory (81 KB); the
// distance channel is gathered from the per-handle constant map (L2 resident).  Walls and the
// distance map do not depend on the env state: they are built once per handle on the host.
#pragma once
#include "raster.cuh"

namespace bp {

constexpr int MZ_WIN = 192;             // local window (cells)
constexpr int MZ_CROP = 288;            // window + inflation (occupancy_map.py:158-160)
constexpr int MZ_PIX = MZ_WIN * MZ_WIN;
constexpr int MZ_OBS_C = 4;

struct MazeObsParams {
    int grid_h, grid_w;                 // 240, 240
    double scale;                       // m_to_pix_scale
    int n_robot_raw;
    double robot_raw[2 * MAX_ROBOT_RAW];
    const unsigned char* wall_map;      // [grid_h][grid_w] 1 = wall (compute_occ_img_walls)
    const double* dist_map;             // [grid_h][grid_w] normalised goal-point distance map
};

struct MazeObsSmem {
    unsigned char crop[MZ_CROP * MZ_CROP];   // bit0 footprint, bit1 movable obstacle, bit2 wall
    double fpv[2 * MAX_ROBOT_RAW];
    int n_fp;
};

BP_DEV double mz_polevl(double x, const double* c, int n) {
    double a = c[0];
    for (int i = 1; i <= n; i++) a = a * x + c[i];
    return a;
}
// cephes sindg / cosdg (scipy.special), degrees
BP_DEV void mz_sincosdg(double x, double* sn, double* cs) {
    const double sincof[6] = {1.58962301572218447952E-10, -2.50507477628503540135E-8, 2.75573136213856773549E-6,
                              -1.98412698295895384658E-4, 8.33333333332211858862E-3, -1.66666666666666307295E-1};
    const double coscof[7] = {1.13678171382044553091E-11, -2.08758833757683644217E-9, 2.75573155429816611547E-7,
                              -2.48015872936186303776E-5, 1.38888888888806666760E-3, -4.16666666666666348141E-2,
                              4.99999999999999999798E-1};
    const double PI180 = 1.74532925199432957692E-2;
    int ssign = 1, csign = 1;
    if (x < 0) { x = -x; ssign = -1; }
    double y = floor(x / 45.0);
    double z = floor(y * 0.0625);        // ldexp(y, -4)
    z = y - z * 16.0;                    // y - ldexp(z, 4)
    int j = (int)z;
    if (j & 1) { j += 1; y += 1.0; }
    j &= 7;
    if (j > 3) { ssign = -ssign; csign = -csign; j -= 4; }
    if (j > 1) csign = -csign;
    z = x - y * 45.0;
    z *= PI180;
    double zz = z * z;
    double ps = z + z * (zz * mz_polevl(zz, sincof, 5));
    double pc = 1.0 - zz * mz_polevl(zz, coscof, 6);
    double s_, c_;
    if (j == 1 || j == 2) { s_ = pc; c_ = ps; } else { s_ = ps; c_ = pc; }
    *sn = ssign < 0 ? -s_ : s_;
    *cs = csign < 0 ? -c_ : c_;
}

#if defined(__CUDACC__) && !defined(BP_HOST_EMU)
BP_DEV double mz_fma(double a, double b, double c) { return __fma_rn(a, b, c); }
#else
BP_DEV double mz_fma(double a, double b, double c) { return fma(a, b, c); }
#endif

// Rasterise polygon (verts in metres) into bit `bit` of the crop; draw.polygon bounding box, clipped to
// `clip_h x clip_w` when clip != 0 (the shape= argument), then to the crop window.
template <typename VT>
BP_DEV void mz_fill_poly(MazeObsSmem& S, const VT* v, int n, double scale, int clip, int grid_h, int grid_w, int g0i, int g0j, unsigned char bit) {
    const int tid = BP_TID, nt = BP_NT;
    double rmin = INFINITY, rmax = -INFINITY, cmin = INFINITY, cmax = -INFINITY;
    for (int i = 0; i < n; i++) {
        double x = v[2 * i] * scale, y = v[2 * i + 1] * scale;
        cmin = fmin(cmin, x); cmax = fmax(cmax, x); rmin = fmin(rmin, y); rmax = fmax(rmax, y);
    }
    long long minr = (long long)fmax(0.0, rmin), maxr = (long long)ceil(rmax);
    long long minc = (long long)fmax(0.0, cmin), maxc = (long long)ceil(cmax);
    if (clip) { if (maxr > grid_h - 1) maxr = grid_h - 1; if (maxc > grid_w - 1) maxc = grid_w - 1; }
    // pixels outside the global map never reach the crop (the crop copies in-bounds cells only)
    if (maxr > grid_h - 1) maxr = grid_h - 1;
    if (maxc > grid_w - 1) maxc = grid_w - 1;
    if (minr < g0i) minr = g0i;
    if (minc < g0j) minc = g0j;
    if (maxr > g0i + MZ_CROP - 1) maxr = g0i + MZ_CROP - 1;
    if (maxc > g0j + MZ_CROP - 1) maxc = g0j + MZ_CROP - 1;
    if (maxr < minr || maxc < minc) return;
    int bw = (int)(maxc - minc + 1), bh = (int)(maxr - minr + 1);
    for (int i = tid; i < bw * bh; i += nt) {
        int rr = (int)minr + i / bw, cc = (int)minc + i % bw;
        if (pip_inclusive(v, n, scale, (double)cc, (double)rr)) S.crop[(rr - g0i) * MZ_CROP + (cc - g0j)] |= bit;
    }
}

// verts / nverts: the movable obstacles as world-space polygons (hull vertices after a step, the raw
// layout vertices at reset: the reference keeps self.obstacles from before the settle steps).
template <typename VT>
BP_DEV void render_maze_obs(MazeObsSmem& S, const MazeObsParams& O, int n_obst, const VT* verts, int poly_stride, const int* nverts,
                            double rx, double ry, double ra, unsigned char* out /* [4][192][192] */) {
    const int tid = BP_TID, nt = BP_NT;
    double s, c; bp_sincos(ra, &s, &c);
    const int wx = (int)(rx * O.scale), wy = (int)(ry * O.scale);
    const int g0i = wy - MZ_CROP / 2, g0j = wx - MZ_CROP / 2;
    // crop of the wall map (zero padding outside the global map)
    for (int i = tid; i < MZ_CROP * MZ_CROP; i += nt) {
        int gi = g0i + i / MZ_CROP, gj = g0j + i % MZ_CROP;
        bool inb = !(gi < 0 || gi >= O.grid_h || gj < 0 || gj >= O.grid_w);
        S.crop[i] = (inb && O.wall_map[gi * O.grid_w + gj]) ? 4 : 0;
    }
    if (tid == 0) {
        // footprint vertices: cfg.robot.vertices rotated about the robot position; those outside the grid
        // are dropped (occupancy_map.py:358-365)
        int k = 0; double ms = -s;
        for (int i = 0; i < O.n_robot_raw; i++) {
            double vx = O.robot_raw[2 * i], vy = O.robot_raw[2 * i + 1];
            double wxx = vx * c + vy * ms + rx, wyy = vx * s + vy * c + ry;
            double gx = wxx * O.scale, gy = wyy * O.scale;
            if (gy < 0 || gy >= O.grid_h || gx < 0 || gx >= O.grid_w) continue;
            S.fpv[2 * k] = wxx; S.fpv[2 * k + 1] = wyy; k++;
        }
        S.n_fp = k;
    }
    BP_BSYNC();
    if (S.n_fp > 0) mz_fill_poly(S, S.fpv, S.n_fp, O.scale, 0, O.grid_h, O.grid_w, g0i, g0j, 1);
    BP_BSYNC();
    for (int b = 0; b < n_obst; b++) {
        mz_fill_poly(S, verts + (size_t)b * poly_stride, nverts[b], O.scale, 1, O.grid_h, O.grid_w, g0i, g0j, 2);
        BP_BSYNC();
    }
    // rotation: corrected_heading = heading - pi/2; degrees = corrected * (180 / pi)
    const double deg = (ra - 3.141592653589793 / 2) * (180 / 3.141592653589793);
    double sd, cd; mz_sincosdg(deg, &sd, &cd);
    const double m00 = cd, m01 = sd, m10 = -sd, m11 = cd;
    const double cen = (MZ_CROP - 1) / 2.0;
    const double off0 = cen - mz_fma(m00, cen, m01 * cen);
    const double off1 = cen - mz_fma(m10, cen, m11 * cen);
    const int h0 = (MZ_CROP - MZ_WIN) / 2;        // inflation // 2 = 48
    const double lim = (double)(MZ_CROP - 1);
    for (int q = tid; q < MZ_PIX / 4; q += nt) {
        int oi = (q * 4) / MZ_WIN, oj0 = (q * 4) % MZ_WIN;
        unsigned w4[MZ_OBS_C] = {0, 0, 0, 0};
        for (int k = 0; k < 4; k++) {
            int i = oi + h0, j = oj0 + k + h0;
            double c0 = (off0 + (double)i * m00) + (double)j * m01;
            double c1 = (off1 + (double)i * m10) + (double)j * m11;
            unsigned char px[MZ_OBS_C] = {0, 0, 0, 255};
            if (!(c0 < 0 || c0 > lim || c1 < 0 || c1 > lim)) {
                double f0 = floor(c0), f1 = floor(c1);
                double a0 = 1.0 - (c0 - f0), b0 = 1.0 - (c1 - f1);
                double a1 = 1.0 - a0, b1 = 1.0 - b0;
                int i0 = (int)f0, j0 = (int)f1;
                int i1 = i0 + 1, j1 = j0 + 1;
                if (i1 >= MZ_CROP) i1 = 2 * MZ_CROP - 2 - i1;
                if (j1 >= MZ_CROP) j1 = 2 * MZ_CROP - 2 - j1;
                unsigned char v00 = S.crop[i0 * MZ_CROP + j0], v01 = S.crop[i0 * MZ_CROP + j1];
                unsigned char v10 = S.crop[i1 * MZ_CROP + j0], v11 = S.crop[i1 * MZ_CROP + j1];
                double w00 = (1.0 * a0) * b0, w01 = (1.0 * a0) * b1, w10 = (1.0 * a1) * b0, w11 = (1.0 * a1) * b1;
                for (int ch = 0; ch < 3; ch++) {
                    unsigned char bit = (unsigned char)(1 << ch);
                    // (v * w0) * w1 with v in {0, 1}: a zero term adds +0.0, which leaves the sum unchanged
                    double t = 0.0;
                    if (v00 & bit) t += w00;
                    if (v01 & bit) t += w01;
                    if (v10 & bit) t += w10;
                    if (v11 & bit) t += w11;
                    px[ch] = (unsigned char)(int)(t * 255.0);
                }
                // distance channel: crop cells outside the global map hold 1.0
                int gi0 = g0i + i0, gi1 = g0i + i1, gj0 = g0j + j0, gj1 = g0j + j1;
                bool ri0 = gi0 >= 0 && gi0 < O.grid_h, ri1 = gi1 >= 0 && gi1 < O.grid_h;
                bool cj0 = gj0 >= 0 && gj0 < O.grid_w, cj1 = gj1 >= 0 && gj1 < O.grid_w;
                double d00 = (ri0 && cj0) ? O.dist_map[gi0 * O.grid_w + gj0] : 1.0;
                double d01 = (ri0 && cj1) ? O.dist_map[gi0 * O.grid_w + gj1] : 1.0;
                double d10 = (ri1 && cj0) ? O.dist_map[gi1 * O.grid_w + gj0] : 1.0;
                double d11 = (ri1 && cj1) ? O.dist_map[gi1 * O.grid_w + gj1] : 1.0;
                double t = 0.0;
                t += (d00 * a0) * b0; t += (d01 * a0) * b1; t += (d10 * a1) * b0; t += (d11 * a1) * b1;
                px[3] = (unsigned char)(int)(t * 255.0);
            }
            for (int ch = 0; ch < MZ_OBS_C; ch++) w4[ch] |= (unsigned)px[ch] << (8 * k);
        }
        for (int ch = 0; ch < MZ_OBS_C; ch++) ((unsigned*)(out + (size_t)ch * MZ_PIX))[q] = w4[ch];
    }
    BP_BSYNC();
}

}  // namespace bp
