// One physics substep for one environment (one warp): neighbour list, awake set, pair
// evaluation, arbiter cache, coloured sequential-impulse solve.  See sim_core.cuh for the design.
// Semantics follow Chipmunk2D's cpSpaceStep as used by the reference through pymunk
// (reference call sites ship_ice_env.py:218-219, 280-281).
#pragma once
#include "sim_core.cuh"

namespace bp {

BP_DEV unsigned pair_key(int i, int j) { return ((unsigned)i << 16) | (unsigned)j; }
BP_DEV double body_margin(const SimParams& P, const EnvScal* sc, int b) { return b == sc->robot ? P.margin_robot : P.margin_body; }
BP_DEV bool body_dynamic(const LayoutView& L, int b) { return L.minv[b] != 0.0; }

// ------------------------------------------------------------------ neighbour list
// Rebuild the candidate pair list from the current AABBs (sorted by (i,j), i<j; ballot compaction
// keeps the order), carry per-pair state (GJK hint, flags, live slot) over from the previous list,
// rebuild the adjacency and the live bitmask.
BP_DEVNI void rebuild_pairs(const EnvView& ev, const SimParams& P, const Dims& d, bool carry) {
    EnvScal* sc = ev.sc;
    const int lane = bp_lane();
    const int nb = sc->nbody;
    const int cur = sc->cur, nxt = carry ? 1 - cur : cur;
    const int old_n = carry ? sc->n_pairs : 0;
    unsigned* nkey = ev.pkey[nxt];
    for (int b = lane; b < nb; b += BP_NL) {
        ev.b0l[b] = ev.bbl[b]; ev.b0b[b] = ev.bbb[b]; ev.b0r[b] = ev.bbr[b]; ev.b0t[b] = ev.bbt[b];
        ev.deg[b] = 0;
    }
    bp_wsync();
    int np = 0;
    bool overflow = false;
    for (int i = 0; i < nb - 1; i++) {
        double mi = body_margin(P, sc, i);
        double il = ev.bbl[i] - mi, ib = ev.bbb[i] - mi, ir = ev.bbr[i] + mi, it = ev.bbt[i] + mi;
        for (int base = i + 1; base < nb; base += BP_NL) {
            int j = base + lane;
            bool hit = false;
            if (j < nb) {
                double mj = body_margin(P, sc, j);
                hit = (il <= ev.bbr[j] + mj && ev.bbl[j] - mj <= ir && ib <= ev.bbt[j] + mj && ev.bbb[j] - mj <= it);
            }
            int tot; int rk = bp_wrank(hit, &tot);
            if (hit) {
                if (np + rk < d.PMAX) nkey[np + rk] = pair_key(i, j); else overflow = true;
            }
            np += tot;
        }
    }
    if (bp_wany(overflow)) { if (lane == 0) bp_atomic_or((unsigned*)&sc->status, ST_PAIR_OVERFLOW); }
    if (np > d.PMAX) np = d.PMAX;
    for (int s = lane; s < d.LMAX; s += BP_NL) ev.ls_keep[s] = 0;
    bp_wsync();
    const unsigned* okey = ev.pkey[cur];
    for (int p = lane; p < np; p += BP_NL) {
        unsigned key = nkey[p];
        unsigned gj = 0; int lv = -1; unsigned char fl = 0;
        if (carry) {
            int lo = 0, hi = old_n - 1;
            while (lo <= hi) {
                int mid = (lo + hi) >> 1;
                unsigned k = okey[mid];
                if (k == key) { gj = ev.pgjk[cur][mid]; lv = ev.plive[cur][mid]; fl = ev.pflag[cur][mid] & 1; break; }
                if (k < key) lo = mid + 1; else hi = mid - 1;
            }
        }
        ev.pgjk[nxt][p] = gj; ev.plive[nxt][p] = lv; ev.pflag[nxt][p] = fl;
        if (lv >= 0) { ev.ls_pair[lv] = p; ev.ls_keep[lv] = 1; }
        bp_atomic_add(&ev.deg[key >> 16], 1);
        bp_atomic_add(&ev.deg[key & 0xFFFF], 1);
    }
    bp_wsync();
    if (carry) {
        // live slots whose pair fell out of the list (bodies now farther apart than the margins)
        for (int base = 0; base < d.LMAX; base += BP_NL) {
            int s = base + lane;
            bool drop = (s < d.LMAX) && ev.ls_pair[s] >= 0 && !ev.ls_keep[s];
            int tot; int rk = bp_wrank(drop, &tot);
            int nf = sc->n_free;
            if (drop) {
                if (!ev.ls_touch[s]) bp_atomic_add(&sc->n_cached, -1);
                ev.ls_pair[s] = -1; ev.ls_touch[s] = 0; ev.free_stack[nf + rk] = s;
            }
            bp_wsync();
            if (lane == 0) sc->n_free = nf + tot;
            bp_wsync();
        }
    }
    // adjacency offsets: exclusive scan of the degrees
    int acc = 0;
    for (int base = 0; base < nb; base += BP_NL) {
        int b = base + lane;
        int dg = (b < nb) ? ev.deg[b] : 0;
        int tot; int ex = bp_wscan_excl(dg, &tot);
        if (b < nb) { ev.adj_start[b] = acc + ex; ev.deg[b] = 0; }
        acc += tot;
    }
    if (lane == 0) ev.adj_start[nb] = acc;
    bp_wsync();
    for (int p = lane; p < np; p += BP_NL) {
        unsigned k = nkey[p]; int i = k >> 16, j = k & 0xFFFF;
        ev.adj[ev.adj_start[i] + bp_atomic_add(&ev.deg[i], 1)] = p;
        ev.adj[ev.adj_start[j] + bp_atomic_add(&ev.deg[j], 1)] = p;
    }
    int words = d.PMAX / 32 + 1;
    for (int w = lane; w < words; w += BP_NL) ev.livemask[w] = 0;
    bp_wsync();
    for (int s = lane; s < d.LMAX; s += BP_NL)
        if (ev.ls_pair[s] >= 0 && ev.ls_touch[s]) bp_atomic_or(&ev.livemask[ev.ls_pair[s] >> 5], 1u << (ev.ls_pair[s] & 31));
    if (lane == 0) { sc->n_pairs = np; sc->cur = nxt; sc->rebuild = 0; sc->dirty = 1; sc->n_apairs = 0; sc->stat_rebuild++; }
    bp_wsync();
}

// Append the pairs incident to body `bdy` that are not yet in the awake-pair list.
BP_DEV void add_body_pairs(const EnvView& ev, int bdy) {
    EnvScal* sc = ev.sc;
    const int lane = bp_lane();
    const int cur = sc->cur;
    int s0 = ev.adj_start[bdy], s1 = ev.adj_start[bdy + 1];
    int n = sc->n_apairs;
    for (int base = s0; base < s1; base += BP_NL) {
        int e = base + lane;
        bool take = false; int p = -1;
        if (e < s1) { p = ev.adj[e]; take = !(ev.pflag[cur][p] & 2); }
        int tot; int rk = bp_wrank(take, &tot);
        if (take) { ev.apairs[n + rk] = p; ev.pflag[cur][p] |= 2; }
        n += tot;
    }
    bp_wsync();
    if (lane == 0) sc->n_apairs = n;
    bp_wsync();
}
BP_DEVNI void build_apairs(const EnvView& ev) {
    EnvScal* sc = ev.sc;
    const int lane = bp_lane();
    const int cur = sc->cur;
    for (int p = lane; p < sc->n_pairs; p += BP_NL) ev.pflag[cur][p] &= ~2;
    if (lane == 0) sc->n_apairs = 0;
    bp_wsync();
    int na = sc->n_awake;
    for (int k = 0; k < na; k++) add_body_pairs(ev, ev.awake[k]);
}

// ------------------------------------------------------------------ pair evaluation
// Evaluate queue[0..nq): exact AABB filter (cpBBIntersects), narrowphase, arbiter update
// (cpArbiterUpdate: r1/r2 relative to the cogs, cached impulses matched by feature hash).
BP_DEVNI void eval_queue(const EnvView& ev, const LayoutView& L, const Dims& d, int nq) {
    EnvScal* sc = ev.sc;
    const int lane = bp_lane();
    const int cur = sc->cur;
    const int stamp = sc->stamp;
    long long gjk_it = 0;
    for (int base = 0; base < nq; base += BP_NL) {
        int q = base + lane;
        bool active = q < nq;
        int p = -1, a = 0, b = 0, slot = -1;
        Manifold m; m.count = 0;
        if (active) {
            p = ev.queue[q];
            unsigned key = ev.pkey[cur][p];
            a = key >> 16; b = key & 0xFFFF;
            slot = ev.plive[cur][p];
            bool bbhit = (ev.bbl[a] <= ev.bbr[b] && ev.bbl[b] <= ev.bbr[a] && ev.bbb[a] <= ev.bbt[b] && ev.bbb[b] <= ev.bbt[a]);
            if (bbhit && !(L.minv[a] == 0.0 && L.minv[b] == 0.0)) {
                unsigned gid = ev.pgjk[cur][p];
                int iters = 0, capped = 0;
                collide(ev, L, a, b, &gid, &m, &iters, &capped);
                ev.pgjk[cur][p] = gid;
                gjk_it += iters;
                if (capped) bp_atomic_or((unsigned*)&sc->status, ST_GJK_ITER);
            }
            ev.pflag[cur][p] |= 1;
        }
        bool touching = active && m.count > 0;
        bool need = touching && slot < 0;
        int tot; int rk = bp_wrank(need, &tot);
        int nfree = sc->n_free;
        bool fresh = false;
        if (need) {
            if (rk < nfree) {
                slot = ev.free_stack[nfree - 1 - rk];
                ev.plive[cur][p] = slot;
                ev.ls_pair[slot] = p; ev.ls_a[slot] = a; ev.ls_b[slot] = b;
                ev.ls_cnt[slot] = 0; ev.ls_touch[slot] = 0; ev.ls_first[slot] = 1;
                ev.ls_e[slot] = L.mat_e[a] * L.mat_e[b];
                ev.ls_u[slot] = L.mat_u[a] * L.mat_u[b];
                fresh = true;
            } else { bp_atomic_or((unsigned*)&sc->status, ST_LIVE_OVERFLOW); touching = false; }
        }
        bp_wsync();
        if (lane == 0) sc->n_free = nfree - (tot < nfree ? tot : nfree);
        if (active && slot >= 0) {
            bool was = !fresh && ev.ls_touch[slot];
            if (touching) {
                double jn_new[2] = {0, 0}, jt_new[2] = {0, 0};
                int oc = fresh ? 0 : ev.ls_cnt[slot];
                for (int i = 0; i < m.count; i++)
                    for (int j = 0; j < oc; j++)
                        if (m.hash[i] == ev.ls_hash[j * d.LMAX + slot]) { jn_new[i] = ev.ls_jn[j * d.LMAX + slot]; jt_new[i] = ev.ls_jt[j * d.LMAX + slot]; }
                for (int i = 0; i < m.count; i++) {
                    int o = i * d.LMAX + slot;
                    ev.ls_r1x[o] = m.p1[i].x - ev.px[a]; ev.ls_r1y[o] = m.p1[i].y - ev.py[a];
                    ev.ls_r2x[o] = m.p2[i].x - ev.px[b]; ev.ls_r2y[o] = m.p2[i].y - ev.py[b];
                    ev.ls_hash[o] = m.hash[i]; ev.ls_jn[o] = jn_new[i]; ev.ls_jt[o] = jt_new[i];
                }
                ev.ls_cnt[slot] = m.count;
                ev.ls_nx[slot] = m.n.x; ev.ls_ny[slot] = m.n.y;
                if (!was) {
                    // new arbiter, or CACHED -> FIRST_COLLISION (cpArbiterUpdate tail)
                    ev.ls_first[slot] = 1;
                    ev.ls_touch[slot] = 1;
                    bp_atomic_or(&ev.livemask[p >> 5], 1u << (p & 31));
                    if (!fresh) bp_atomic_add(&sc->n_cached, -1);
                    sc->dirty = 1;
                }
                ev.ls_stamp[slot] = stamp;
                // a touching pair with an awake endpoint wakes the other (dynamic) endpoint
                if ((!ev.is_awake[a] && body_dynamic(L, a)) || (!ev.is_awake[b] && body_dynamic(L, b))) sc->any_wake = 1;
            } else if (was) {
                // separated: the arbiter stays cached (contacts + impulses kept) for `persistence` steps
                ev.ls_touch[slot] = 0;
                ev.ls_stamp[slot] = stamp - 1;
                bp_atomic_and(&ev.livemask[p >> 5], ~(1u << (p & 31)));
                bp_atomic_add(&sc->n_cached, 1);
                sc->dirty = 1;
            }
        }
        bp_wsync();
    }
    if (gjk_it) { /* statistics only */
#if defined(__CUDACC__) && !defined(BP_HOST_EMU)
        atomicAdd((unsigned long long*)&sc->stat_gjk, (unsigned long long)gjk_it);
#else
        sc->stat_gjk += gjk_it;
#endif
    }
    if (lane == 0) sc->stat_pairs_eval += nq;
    bp_wsync();
}

// Gather the awake pairs [i0, n_apairs) that need evaluation into the queue.
BP_DEV int gather_queue(const EnvView& ev, int i0) {
    EnvScal* sc = ev.sc;
    const int lane = bp_lane();
    const int cur = sc->cur;
    const int n = sc->n_apairs;
    int nq = 0;
    for (int base = i0; base < n; base += BP_NL) {
        int i = base + lane;
        bool take = false; int p = -1;
        if (i < n) {
            p = ev.apairs[i];
            unsigned key = ev.pkey[cur][p];
            take = !(ev.pflag[cur][p] & 1) || ev.moved[key >> 16] || ev.moved[key & 0xFFFF];
        }
        int tot; int rk = bp_wrank(take, &tot);
        if (take) ev.queue[nq + rk] = p;
        nq += tot;
    }
    bp_wsync();
    return nq;
}

// Wake every asleep dynamic body that shares a touching arbiter with an awake body; evaluate
// its never-evaluated pairs; repeat until closed (complete contact components).
BP_DEVNI void wake_closure(const EnvView& ev, const LayoutView& L, const Dims& d, int nq_last) {
    EnvScal* sc = ev.sc;
    const int lane = bp_lane();
    int nq = nq_last;
    while (sc->any_wake) {
        bp_wsync();
        int first_new = sc->n_awake;
        int ap0 = sc->n_apairs;
        if (lane == 0) {
            sc->any_wake = 0;
            int na = sc->n_awake;
            const int cur = sc->cur;
            for (int q = 0; q < nq; q++) {
                int p = ev.queue[q];
                int slot = ev.plive[cur][p];
                if (slot < 0 || !ev.ls_touch[slot]) continue;
                unsigned key = ev.pkey[cur][p];
                int ab[2] = {(int)(key >> 16), (int)(key & 0xFFFF)};
                for (int t = 0; t < 2; t++) {
                    int x = ab[t];
                    if (!ev.is_awake[x] && body_dynamic(L, x)) { ev.is_awake[x] = 1; ev.moved[x] = 0; ev.awake[na++] = x; }
                }
            }
            sc->n_awake = na;
        }
        bp_wsync();
        int na = sc->n_awake;
        for (int k = first_new; k < na; k++) add_body_pairs(ev, ev.awake[k]);
        nq = gather_queue(ev, ap0);
        eval_queue(ev, L, d, nq);
    }
}

// Free cached (separated) arbiters older than `persistence` steps (cpSpaceArbiterSetFilter).
BP_DEV void expire_cached(const EnvView& ev, const SimParams& P, const Dims& d) {
    EnvScal* sc = ev.sc;
    if (sc->n_cached <= 0) return;
    const int lane = bp_lane();
    const int cur = sc->cur;
    for (int base = 0; base < d.LMAX; base += BP_NL) {
        int s = base + lane;
        bool drop = (s < d.LMAX) && ev.ls_pair[s] >= 0 && !ev.ls_touch[s] && (sc->stamp - ev.ls_stamp[s] >= P.persistence);
        int tot; int rk = bp_wrank(drop, &tot);
        int nf = sc->n_free;
        if (drop) { ev.plive[cur][ev.ls_pair[s]] = -1; ev.ls_pair[s] = -1; ev.free_stack[nf + rk] = s; }
        bp_wsync();
        if (lane == 0 && tot) { sc->n_free = nf + tot; sc->n_cached -= tot; }
        bp_wsync();
    }
}

// Enumerate touching arbiters in ascending pair id and order them by greedy colour.
BP_DEVNI void build_order(const EnvView& ev, const LayoutView& L, const Dims& d) {
    EnvScal* sc = ev.sc;
    const int lane = bp_lane();
    const int cur = sc->cur;
    const int words = (sc->n_pairs + 31) / 32;
    const int wpl = (words + BP_NL - 1) / BP_NL;
    int w0 = lane * wpl, w1 = w0 + wpl; if (w1 > words) w1 = words;
    int cnt = 0;
    for (int w = w0; w < w1; w++) cnt += __popc(ev.livemask[w]);
    int tot; int pos = bp_wscan_excl(cnt, &tot);
    if (tot > d.LMAX) tot = d.LMAX;
    for (int w = w0; w < w1; w++) {
        unsigned mk_ = ev.livemask[w];
        while (mk_) {
            int bit = __ffsll((long long)mk_) - 1; mk_ &= mk_ - 1;
            int p = w * 32 + bit;
            if (pos < d.LMAX) ev.colour[pos] = ev.plive[cur][p];     // colour[] doubles as the canonical-order scratch
            pos++;
        }
    }
    bp_wsync();
    // greedy colouring in canonical order (sequential by definition; runs only when the set changed)
    if (lane == 0) {
        int na = sc->n_awake;
        for (int k = 0; k < na; k++) ev.colmask[ev.awake[k]] = 0ull;
        int cnts[MAXCOL + 1];
        for (int c = 0; c <= MAXCOL; c++) cnts[c] = 0;
        int maxc = -1;
        for (int i = 0; i < tot; i++) {
            int s = ev.colour[i];
            int a = ev.ls_a[s], b = ev.ls_b[s];
            unsigned long long used = 0ull;
            bool da = body_dynamic(L, a), db = body_dynamic(L, b);
            if (da) used |= ev.colmask[a];
            if (db) used |= ev.colmask[b];
            int c = (~used == 0ull) ? MAXCOL : (__ffsll((long long)(~used)) - 1);
            if (c >= MAXCOL) { sc->status |= ST_COLOUR_OVERFLOW; c = MAXCOL - 1; }
            unsigned long long bit = 1ull << c;
            if (da) ev.colmask[a] |= bit;
            if (db) ev.colmask[b] |= bit;
            ev.free_stack[d.LMAX - 1 - i] = c;       // scratch: top of the free stack is unused while n_free + tot <= LMAX
            cnts[c]++; if (c > maxc) maxc = c;
        }
        int acc = 0;
        for (int c = 0; c <= maxc; c++) { ev.col_start[c] = acc; acc += cnts[c]; cnts[c] = ev.col_start[c]; }
        ev.col_start[maxc + 1] = acc;
        for (int i = 0; i < tot; i++) { int c = ev.free_stack[d.LMAX - 1 - i]; ev.order[cnts[c]++] = ev.colour[i]; }
        sc->n_order = tot; sc->n_colours = maxc + 1; sc->dirty = 0; sc->stat_recolour++;
    }
    bp_wsync();
}

// ------------------------------------------------------------------ solver pieces (cpArbiter.c)
BP_DEV void arb_prestep(const EnvView& ev, const LayoutView& L, const SimParams& P, const Dims& d, int s) {
    int a = ev.ls_a[s], b = ev.ls_b[s];
    V2 n = mk(ev.ls_nx[s], ev.ls_ny[s]);
    V2 tn = perp(n);
    double ma = L.minv[a], ia = L.iinv[a], mb = L.minv[b], ib = L.iinv[b];
    V2 delta = mk(ev.px[b] - ev.px[a], ev.py[b] - ev.py[a]);
    V2 va = mk(ev.vx[a], ev.vy[a]), vb = mk(ev.vx[b], ev.vy[b]);
    double wa = ev.vw[a], wb = ev.vw[b];
    int cnt = ev.ls_cnt[s];
    for (int i = 0; i < cnt; i++) {
        int o = i * d.LMAX + s;
        V2 r1 = mk(ev.ls_r1x[o], ev.ls_r1y[o]), r2 = mk(ev.ls_r2x[o], ev.ls_r2y[o]);
        double r1n = cross(r1, n), r2n = cross(r2, n);
        double kn = (ma + ia * r1n * r1n) + (mb + ib * r2n * r2n);
        double r1t = cross(r1, tn), r2t = cross(r2, tn);
        double kt = (ma + ia * r1t * r1t) + (mb + ib * r2t * r2t);
        ev.ls_nm[o] = 1.0 / kn; ev.ls_tm[o] = 1.0 / kt;
        double dist = dot((r2 - r1) + delta, n);
        ev.ls_bias[o] = -P.bias_coef * fmin(0.0, dist + P.slop) / P.dt;
        ev.ls_jb[o] = 0.0;
        V2 vr = (vb + perp(r2) * wb) - (va + perp(r1) * wa);
        ev.ls_bounce[o] = dot(vr, n) * ev.ls_e[s];
    }
}
BP_DEV void arb_warm(const EnvView& ev, const LayoutView& L, const Dims& d, int s, double dt_coef) {
    if (ev.ls_first[s]) return;
    int a = ev.ls_a[s], b = ev.ls_b[s];
    V2 n = mk(ev.ls_nx[s], ev.ls_ny[s]);
    double ma = L.minv[a], ia = L.iinv[a], mb = L.minv[b], ib = L.iinv[b];
    int cnt = ev.ls_cnt[s];
    for (int i = 0; i < cnt; i++) {
        int o = i * d.LMAX + s;
        V2 r1 = mk(ev.ls_r1x[o], ev.ls_r1y[o]), r2 = mk(ev.ls_r2x[o], ev.ls_r2y[o]);
        V2 j = rotv(n, mk(ev.ls_jn[o], ev.ls_jt[o])) * dt_coef;
        V2 nj = neg(j);
        ev.vx[a] = ev.vx[a] + nj.x * ma; ev.vy[a] = ev.vy[a] + nj.y * ma; ev.vw[a] += ia * cross(r1, nj);
        ev.vx[b] = ev.vx[b] + j.x * mb; ev.vy[b] = ev.vy[b] + j.y * mb; ev.vw[b] += ib * cross(r2, j);
    }
}
BP_DEV void arb_solve(const EnvView& ev, const LayoutView& L, const Dims& d, int s) {
    int a = ev.ls_a[s], b = ev.ls_b[s];
    V2 n = mk(ev.ls_nx[s], ev.ls_ny[s]);
    V2 tn = perp(n);
    double fr = ev.ls_u[s];
    double ma = L.minv[a], ia = L.iinv[a], mb = L.minv[b], ib = L.iinv[b];
    V2 va = mk(ev.vx[a], ev.vy[a]), vb = mk(ev.vx[b], ev.vy[b]);
    double wa = ev.vw[a], wb = ev.vw[b];
    V2 ba = mk(ev.bx[a], ev.by[a]), bb = mk(ev.bx[b], ev.by[b]);
    double bwa = ev.bw[a], bwb = ev.bw[b];
    int cnt = ev.ls_cnt[s];
    for (int i = 0; i < cnt; i++) {
        int o = i * d.LMAX + s;
        V2 r1 = mk(ev.ls_r1x[o], ev.ls_r1y[o]), r2 = mk(ev.ls_r2x[o], ev.ls_r2y[o]);
        double nMass = ev.ls_nm[o];
        V2 vb1 = ba + perp(r1) * bwa;
        V2 vb2 = bb + perp(r2) * bwb;
        V2 vr = (vb + perp(r2) * wb) - (va + perp(r1) * wa);
        double vbn = dot(vb2 - vb1, n);
        double vrn = dot(vr, n);
        double vrt = dot(vr, tn);
        double jbn = (ev.ls_bias[o] - vbn) * nMass;
        double jbnOld = ev.ls_jb[o];
        double jb = fmax(jbnOld + jbn, 0.0);
        double jn = -(ev.ls_bounce[o] + vrn) * nMass;
        double jnOld = ev.ls_jn[o];
        double jnAcc = fmax(jnOld + jn, 0.0);
        double jtMax = fr * jnAcc;
        double jt = -vrt * ev.ls_tm[o];
        double jtOld = ev.ls_jt[o];
        double jtAcc = clampd(jtOld + jt, -jtMax, jtMax);
        ev.ls_jb[o] = jb; ev.ls_jn[o] = jnAcc; ev.ls_jt[o] = jtAcc;
        V2 jbv = n * (jb - jbnOld), njbv = neg(jbv);
        ba = ba + njbv * ma; bwa += ia * cross(r1, njbv);
        bb = bb + jbv * mb; bwb += ib * cross(r2, jbv);
        V2 j = rotv(n, mk(jnAcc - jnOld, jtAcc - jtOld)), nj = neg(j);
        va = va + nj * ma; wa += ia * cross(r1, nj);
        vb = vb + j * mb; wb += ib * cross(r2, j);
    }
    // kinematic / static bodies have m_inv = i_inv = 0: their velocities are unchanged (x + 0), skip the stores
    if (ma != 0.0 || ia != 0.0) { ev.vx[a] = va.x; ev.vy[a] = va.y; ev.vw[a] = wa; ev.bx[a] = ba.x; ev.by[a] = ba.y; ev.bw[a] = bwa; }
    if (mb != 0.0 || ib != 0.0) { ev.vx[b] = vb.x; ev.vy[b] = vb.y; ev.vw[b] = wb; ev.bx[b] = bb.x; ev.by[b] = bb.y; ev.bw[b] = bwb; }
}

// ------------------------------------------------------------------ one substep (cpSpaceStep)
BP_DEVNI void substep(const EnvView& ev, const LayoutView& L, const SimParams& P, const Dims& d, double dt, double prev_dt) {
    EnvScal* sc = ev.sc;
    const int lane = bp_lane();
    if (lane == 0) { sc->stamp++; sc->any_wake = 0; }
    bp_wsync();
    // 1. integrate positions of the awake bodies (everything else has v = v_bias = 0 exactly)
    int na = sc->n_awake;
    for (int k = lane; k < na; k += BP_NL) {
        int b = ev.awake[k];
        double vx = ev.vx[b], vy = ev.vy[b], vw = ev.vw[b], bx = ev.bx[b], by = ev.by[b], bw = ev.bw[b];
        bool mv = (vx != 0.0 || vy != 0.0 || vw != 0.0 || bx != 0.0 || by != 0.0 || bw != 0.0);
        if (mv) {
            ev.px[b] = ev.px[b] + (vx + bx) * dt;
            ev.py[b] = ev.py[b] + (vy + by) * dt;
            ev.pa[b] = ev.pa[b] + (vw + bw) * dt;
            refresh_body(ev, L, b);
            double mg = body_margin(P, sc, b);
            if (ev.bbl[b] < ev.b0l[b] - mg || ev.bbr[b] > ev.b0r[b] + mg || ev.bbb[b] < ev.b0b[b] - mg || ev.bbt[b] > ev.b0t[b] + mg) sc->rebuild = 1;
            ev.moved_step[b] = 1;
            ev.bx[b] = 0.0; ev.by[b] = 0.0; ev.bw[b] = 0.0;
        }
        ev.moved[b] = mv ? 1 : 0;
    }
    bp_wsync();
    if (sc->rebuild) { rebuild_pairs(ev, P, d, true); build_apairs(ev); }
    // 2. collide: pairs with a moved endpoint (or never evaluated), then close the awake set
    int nq = gather_queue(ev, 0);
    eval_queue(ev, L, d, nq);
    wake_closure(ev, L, d, nq);
    expire_cached(ev, P, d);
    if (sc->dirty) build_order(ev, L, d);
    // 3. prestep (uses the pre-damping velocities for the bounce term)
    int no = sc->n_order;
    for (int i = lane; i < no; i += BP_NL) arb_prestep(ev, L, P, d, ev.order[i]);
    bp_wsync();
    // 4. velocity update: v = v * damping^dt (gravity = 0, no forces); dynamic bodies only
    na = sc->n_awake;
    for (int k = lane; k < na; k += BP_NL) {
        int b = ev.awake[k];
        if (body_dynamic(L, b)) {
            ev.vx[b] = ev.vx[b] * P.damp + 0.0; ev.vy[b] = ev.vy[b] * P.damp + 0.0; ev.vw[b] = ev.vw[b] * P.damp + 0.0;
        }
    }
    bp_wsync();
    // 5. warm start + iterations, colour by colour (same-colour arbiters touch disjoint dynamic bodies)
    double dt_coef = (prev_dt == 0.0) ? 0.0 : dt / prev_dt;
    int nc = sc->n_colours;
    for (int c = 0; c < nc; c++) {
        int c0 = ev.col_start[c], c1 = ev.col_start[c + 1];
        for (int i = c0 + lane; i < c1; i += BP_NL) arb_warm(ev, L, d, ev.order[i], dt_coef);
        bp_wsync();
    }
    for (int it = 0; it < P.iterations; it++) {
        for (int c = 0; c < nc; c++) {
            int c0 = ev.col_start[c], c1 = ev.col_start[c + 1];
            for (int i = c0 + lane; i < c1; i += BP_NL) arb_solve(ev, L, d, ev.order[i]);
            bp_wsync();
        }
    }
    // 6. arbiters solved this step are NORMAL next step
    for (int i = lane; i < no; i += BP_NL) ev.ls_first[ev.order[i]] = 0;
    if (lane == 0) { sc->stat_substeps++; sc->stat_live_sum += no; sc->stat_awake_sum += na; }
    bp_wsync();
}

// ------------------------------------------------------------------ awake-set maintenance
// Put to sleep every awake body whose complete contact component carries no velocity, no bias
// velocity, no cached impulse and does not touch the kinematic robot: the solver is then exactly a
// no-op for it.  Runs between env steps.
BP_DEVNI void sleep_pass(const EnvView& ev, const LayoutView& L, const Dims& d) {
    EnvScal* sc = ev.sc;
    const int lane = bp_lane();
    const int cur = sc->cur;
    int na = sc->n_awake;
    for (int k = lane; k < na; k += BP_NL) {
        int b = ev.awake[k];
        bool nz = (ev.vx[b] != 0.0 || ev.vy[b] != 0.0 || ev.vw[b] != 0.0 || ev.bx[b] != 0.0 || ev.by[b] != 0.0 || ev.bw[b] != 0.0);
        ev.act[b] = (nz || !body_dynamic(L, b)) ? 1 : 0;
    }
    bp_wsync();
    for (int s = lane; s < d.LMAX; s += BP_NL) {
        if (ev.ls_pair[s] < 0) continue;
        int a = ev.ls_a[s], b = ev.ls_b[s];
        bool seed = false;
        for (int i = 0; i < ev.ls_cnt[s]; i++) {
            int o = i * d.LMAX + s;
            if (ev.ls_jn[o] != 0.0 || ev.ls_jt[o] != 0.0 || ev.ls_jb[o] != 0.0) seed = true;
        }
        if (ev.ls_touch[s] && (!body_dynamic(L, a) || !body_dynamic(L, b))) seed = true;   // touches the robot
        if (seed) { ev.act[a] = 1; ev.act[b] = 1; }
    }
    bp_wsync();
    for (;;) {
        bool changed = false;
        for (int s = lane; s < d.LMAX; s += BP_NL) {
            if (ev.ls_pair[s] < 0 || !ev.ls_touch[s]) continue;
            int a = ev.ls_a[s], b = ev.ls_b[s];
            if (!body_dynamic(L, a) || !body_dynamic(L, b)) continue;
            if (ev.act[a] != ev.act[b]) { ev.act[a] = 1; ev.act[b] = 1; changed = true; }
        }
        bp_wsync();
        if (!bp_wany(changed)) break;
    }
    // free the arbiters of sleeping bodies, clear the evaluated flags of their pairs
    bool any_sleep = false;
    for (int base = 0; base < d.LMAX; base += BP_NL) {
        int s = base + lane;
        bool drop = false;
        if (s < d.LMAX && ev.ls_pair[s] >= 0) {
            int a = ev.ls_a[s], b = ev.ls_b[s];
            drop = (body_dynamic(L, a) && !ev.act[a]) || (body_dynamic(L, b) && !ev.act[b]);
        }
        int tot; int rk = bp_wrank(drop, &tot);
        int nf = sc->n_free;
        if (drop) {
            int p = ev.ls_pair[s];
            if (ev.ls_touch[s]) bp_atomic_and(&ev.livemask[p >> 5], ~(1u << (p & 31))); else bp_atomic_add(&sc->n_cached, -1);
            ev.plive[cur][p] = -1; ev.ls_pair[s] = -1; ev.ls_touch[s] = 0; ev.free_stack[nf + rk] = s;
        }
        bp_wsync();
        if (lane == 0 && tot) { sc->n_free = nf + tot; sc->dirty = 1; }
        bp_wsync();
    }
    for (int k = 0; k < na; k++) {
        int b = ev.awake[k];
        if (ev.act[b]) continue;
        any_sleep = true;
        int s0 = ev.adj_start[b], s1 = ev.adj_start[b + 1];
        for (int e = s0 + lane; e < s1; e += BP_NL) ev.pflag[cur][ev.adj[e]] &= ~1;
    }
    bp_wsync();
    // compact the awake list
    int nn = 0;
    for (int base = 0; base < na; base += BP_NL) {
        int k = base + lane;
        int b = (k < na) ? ev.awake[k] : -1;
        bool keep = (k < na) && ev.act[b];
        int tot; int rk = bp_wrank(keep, &tot);
        bp_wsync();
        if (k < na && !keep) { ev.is_awake[b] = 0; ev.moved[b] = 0; }
        if (keep) ev.awake[nn + rk] = b;
        nn += tot;
        bp_wsync();
    }
    if (lane == 0) sc->n_awake = nn;
    bp_wsync();
    if (any_sleep || sc->n_apairs == 0) build_apairs(ev);
}

// Initialise an environment from a layout: poses, zero velocities, everything awake, empty cache.
BP_DEVNI void init_env(const EnvView& ev, const LayoutView& L, const SimParams& P, const Dims& d, int layout_id) {
    EnvScal* sc = ev.sc;
    const int lane = bp_lane();
    const int nb = L.nbody;
    if (lane == 0) {
        sc->stamp = 0; sc->n_awake = nb; sc->n_pairs = 0; sc->cur = 0; sc->n_free = d.LMAX; sc->n_order = 0; sc->n_colours = 0;
        sc->dirty = 1; sc->status = 0; sc->layout = layout_id; sc->nbody = nb; sc->robot = nb - 1; sc->t = 0; sc->rebuild = 0;
        sc->n_apairs = 0; sc->n_cached = 0; sc->any_wake = 0; sc->total_work = 0.0;
    }
    for (int b = lane; b < nb; b += BP_NL) {
        double a = L.init[3 * b + 2];
        double s, c; bp_sincos(a, &s, &c);
        double cx = L.cog[2 * b], cy = L.cog[2 * b + 1];
        // cpBodySetPosition: p = position + rotate(cog)
        ev.px[b] = (c * cx + (-s) * cy) + L.init[3 * b];
        ev.py[b] = (s * cx + c * cy) + L.init[3 * b + 1];
        ev.pa[b] = a;
        ev.vx[b] = ev.vy[b] = ev.vw[b] = 0.0; ev.bx[b] = ev.by[b] = ev.bw[b] = 0.0;
        refresh_body(ev, L, b);
        ev.awake[b] = b; ev.is_awake[b] = 1; ev.moved[b] = 0; ev.moved_step[b] = 0; ev.act[b] = 1;
    }
    for (int s = lane; s < d.LMAX; s += BP_NL) { ev.free_stack[s] = d.LMAX - 1 - s; ev.ls_pair[s] = -1; ev.ls_touch[s] = 0; ev.ls_cnt[s] = 0; }
    bp_wsync();
    rebuild_pairs(ev, P, d, false);
    build_apairs(ev);
}

}  // namespace bp
