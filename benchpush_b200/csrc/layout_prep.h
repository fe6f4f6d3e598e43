// Host-side preparation of immutable layout tables (run once per layout at load time).
//
// Turns the reference's trial description (ship_ice_env.py:186-216: obstacle vertex lists + centres,
// ship start pose) into the body tables the kernels read: convex hull (pymunk.Poly builds the hull
// of the given vertices), recentring on the hull centroid (sim_utils.py:136-153 create_polygon),
// edge normals, mass = density * (area + r*(perimeter + pi*r)), moment about the centroid
// (Chipmunk cpPolyShape mass info).
#pragma once
#include <algorithm>
#include <cmath>
#include <vector>

namespace bp_prep {

struct P2 { double x, y; };

// Strictly convex CCW hull starting at the lexicographically smallest vertex (x, then y):
// the vertex order pymunk.Poly.get_vertices() reports.
inline std::vector<P2> convex_hull(std::vector<P2> pts) {
    std::sort(pts.begin(), pts.end(), [](const P2& a, const P2& b) { return a.x < b.x || (a.x == b.x && a.y < b.y); });
    pts.erase(std::unique(pts.begin(), pts.end(), [](const P2& a, const P2& b) { return a.x == b.x && a.y == b.y; }), pts.end());
    size_t n = pts.size();
    if (n < 3) return pts;
    auto turn = [](const P2& o, const P2& a, const P2& b) { return (a.x - o.x) * (b.y - o.y) - (a.y - o.y) * (b.x - o.x); };
    std::vector<P2> h(2 * n);
    size_t k = 0;
    for (size_t i = 0; i < n; i++) {
        while (k >= 2 && turn(h[k - 2], h[k - 1], pts[i]) <= 0) k--;
        h[k++] = pts[i];
    }
    for (size_t i = n - 1, t = k + 1; i > 0; i--) {
        while (k >= t && turn(h[k - 2], h[k - 1], pts[i - 1]) <= 0) k--;
        h[k++] = pts[i - 1];
    }
    h.resize(k - 1);
    return h;
}

inline P2 centroid(const std::vector<P2>& v) {
    double sum = 0, sx = 0, sy = 0;
    size_t n = v.size();
    for (size_t i = 0; i < n; i++) {
        const P2& a = v[i]; const P2& b = v[(i + 1) % n];
        double cr = a.x * b.y - a.y * b.x;
        sum += cr; sx += (a.x + b.x) * cr; sy += (a.y + b.y) * cr;
    }
    double k = 1.0 / (3.0 * sum);
    return P2{sx * k, sy * k};
}
inline double area_with_radius(const std::vector<P2>& v, double r) {
    double area = 0, per = 0;
    size_t n = v.size();
    for (size_t i = 0; i < n; i++) {
        const P2& a = v[i]; const P2& b = v[(i + 1) % n];
        area += a.x * b.y - a.y * b.x;
        per += std::sqrt((a.x - b.x) * (a.x - b.x) + (a.y - b.y) * (a.y - b.y));
    }
    return r * (M_PI * std::fabs(r) + per) + area / 2.0;
}
// moment of a unit-mass polygon about `about` (the radius is ignored, as Chipmunk does)
inline double unit_moment(const std::vector<P2>& v, P2 about) {
    double s1 = 0, s2 = 0;
    size_t n = v.size();
    for (size_t i = 0; i < n; i++) {
        P2 a{v[i].x - about.x, v[i].y - about.y};
        P2 b{v[(i + 1) % n].x - about.x, v[(i + 1) % n].y - about.y};
        double cr = b.x * a.y - b.y * a.x;
        double dd = (a.x * a.x + a.y * a.y) + (a.x * b.x + a.y * b.y) + (b.x * b.x + b.y * b.y);
        s1 += cr * dd; s2 += cr;
    }
    return s1 / (6.0 * s2);
}

struct BodyPrep {
    std::vector<P2> verts;      // hull, body-local
    std::vector<P2> normals;    // normals[i] = normal of the edge ending at vertex i
    P2 cog;
    double m_inv, i_inv;
};

// dynamic == false: kinematic / static body (infinite mass, cog at the body origin)
inline BodyPrep prep_body(const std::vector<P2>& local_verts, double radius, double density, bool dynamic) {
    BodyPrep b;
    b.verts = convex_hull(local_verts);
    size_t n = b.verts.size();
    b.normals.resize(n);
    for (size_t i = 0; i < n; i++) {
        const P2& p0 = b.verts[(i + n - 1) % n]; const P2& p1 = b.verts[i];
        double ex = p1.x - p0.x, ey = p1.y - p0.y;          // rperp(e) = (e.y, -e.x)
        double len = std::sqrt(ey * ey + ex * ex);
        double inv = 1.0 / (len + 2.2250738585072014e-308);
        b.normals[i] = P2{ey * inv, -ex * inv};
    }
    if (dynamic) {
        b.cog = centroid(b.verts);
        double m = density * area_with_radius(b.verts, radius);
        double im = m * unit_moment(b.verts, b.cog);
        b.m_inv = 1.0 / m; b.i_inv = 1.0 / im;
    } else {
        b.cog = P2{0, 0}; b.m_inv = 0.0; b.i_inv = 0.0;
    }
    return b;
}

// sim_utils.py:136-153: vertices relative to the given centre are re-centred on the centroid of
// their hull before the shape is built; the body sits at `centre`.
inline std::vector<P2> recentre_on_hull_centroid(const std::vector<P2>& rel) {
    P2 c = centroid(convex_hull(rel));
    std::vector<P2> out(rel.size());
    for (size_t i = 0; i < rel.size(); i++) out[i] = P2{rel[i].x - c.x, rel[i].y - c.y};
    return out;
}

}  // namespace bp_prep
