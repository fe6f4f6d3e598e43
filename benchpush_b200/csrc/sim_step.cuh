// One physics substep for one environment (one warp): neighbour list, awake set, staged pair
// evaluation, arbiter cache, coloured sequential-impulse solve.  See sim_core.cuh for the design.
// Semantics follow Chipmunk2D's cpSpaceStep as used by the reference through pymunk
// (reference call sites ship_ice_env.py:218-219, 280-281).
#pragma once
#include "sim_core.cuh"

namespace bp {

BP_DEV unsigned pair_key(int i, int j) { return ((unsigned)i << 16) | (unsigned)j; }
BP_DEV real body_margin(const SimParams& P, const EnvScal* sc, int b) { return b >= sc->robot ? P.margin_robot : P.margin_body; }
BP_DEV bool body_dynamic(const LayoutView& L, int b) { return L.minv[b] != RC(0.0); }

#if defined(BP_PHASE_TIMERS) && defined(__CUDACC__) && !defined(BP_HOST_EMU)
#define BP_TICK(var) long long var = clock64()
#define BP_TOCK(sc, k, var) do { if (bp_lane() == 0) (sc)->t_phase[k] += clock64() - var; var = clock64(); } while (0)
#else
#define BP_TICK(var)
#define BP_TOCK(sc, k, var)
#endif

// ------------------------------------------------------------------ neighbour list
// Tail of a neighbour-list build: np_ = ev.pair[nxt][0..np) holds the new list (sorted by key, per-pair
// state already carried over).  Re-links the live arbiter slots, frees the slots whose pair fell out
// of the list, rebuilds the adjacency and the live bitmask, clears the moved-since-build flags.
BP_DEV void rebuild_finish(const EnvView& ev, const Dims& d, int np, int nxt, bool carry) {
    EnvScal* sc = ev.sc;
    const int lane = bp_lane();
    const int nb = sc->nbody;
    PairRec* np_ = ev.pair[nxt];
    const int hw = sc->slot_hw;
    for (int s = lane; s < hw; s += BP_NL) ev.slot[s].keep = 0;
    for (int b = lane; b < nb; b += BP_NL) { ev.deg[b] = 0; ev.mvrb[b] = 0; }
    bp_wsync();
    for (int p = lane; p < np; p += BP_NL) {
        unsigned key = np_[p].key;
        int lv = np_[p].live;
        if (lv >= 0) { ev.slot[lv].pair = p; ev.slot[lv].keep = 1; }
        bp_atomic_add(&ev.deg[key >> 16], 1);
        bp_atomic_add(&ev.deg[key & 0xFFFF], 1);
    }
    bp_wsync();
    if (carry) {
        // live slots whose pair fell out of the list (bodies now farther apart than the margins)
        for (int base = 0; base < hw; base += BP_NL) {
            int s = base + lane;
            bool drop = (s < hw) && ev.slot[s].pair >= 0 && !ev.slot[s].keep;
            int tot; int rk = bp_wrank(drop, &tot);
            int nf = sc->n_free;
            if (drop) {
                if (!ev.slot[s].touch) bp_atomic_add(&sc->n_cached, -1);
                ev.slot[s].pair = -1; ev.slot[s].touch = 0; ev.free_stack[nf + rk] = s;
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
        unsigned k = np_[p].key; int i = k >> 16, j = k & 0xFFFF;
        ev.adj[ev.adj_start[i] + bp_atomic_add(&ev.deg[i], 1)] = p;
        ev.adj[ev.adj_start[j] + bp_atomic_add(&ev.deg[j], 1)] = p;
    }
    int words = carry ? (np > sc->n_pairs ? np : sc->n_pairs) / 32 + 1 : d.PMAX / 32 + 1;
    for (int w = lane; w < words; w += BP_NL) ev.livemask[w] = 0;
    bp_wsync();
    for (int s = lane; s < hw; s += BP_NL)
        if (ev.slot[s].pair >= 0 && ev.slot[s].touch) bp_atomic_or(&ev.livemask[ev.slot[s].pair >> 5], 1u << (ev.slot[s].pair & 31));
    if (lane == 0) { sc->n_pairs = np; sc->cur = nxt; sc->rebuild = 0; sc->dirty = 1; sc->n_apairs = 0; sc->stat_rebuild++; }
    bp_wsync();
}

// Rebuild the candidate pair list from the current AABBs (sorted by (i,j), i<j; ballot compaction
// keeps the order), carry per-pair state (GJK hint, flags, skip bound, live slot) over from the
// previous list, rebuild the adjacency and the live bitmask.
BP_DEVNI void rebuild_pairs(const EnvView& ev, const SimParams& P, const Dims& d, bool carry) {
    EnvScal* sc = ev.sc;
    const int lane = bp_lane();
    const int nb = sc->nbody;
    const int cur = sc->cur, nxt = carry ? 1 - cur : cur;
    const int old_n = carry ? sc->n_pairs : 0;
    PairRec* np_ = ev.pair[nxt];
    for (int b = lane; b < nb; b += BP_NL) ev.b0[b] = ev.bb[b];
    bp_wsync();
    int np = 0;
    bool overflow = false;
    for (int i = 0; i < nb - 1; i++) {
        real mi = body_margin(P, sc, i);
        BBox bi = ev.b0[i];
        real il = bi.l - mi, ib = bi.b - mi, ir = bi.r + mi, it = bi.t + mi;
        for (int base = i + 1; base < nb; base += BP_NL) {
            int j = base + lane;
            bool hit = false;
            if (j < nb) {
                real mj = body_margin(P, sc, j);
                BBox bj = ev.b0[j];
                hit = (il <= bj.r + mj && bj.l - mj <= ir && ib <= bj.t + mj && bj.b - mj <= it);
            }
            int tot; int rk = bp_wrank(hit, &tot);
            if (hit) {
                if (np + rk < d.PMAX) np_[np + rk].key = pair_key(i, j); else overflow = true;
            }
            np += tot;
        }
    }
    if (bp_wany(overflow)) { if (lane == 0) bp_atomic_or((unsigned*)&sc->status, ST_PAIR_OVERFLOW); }
    if (np > d.PMAX) np = d.PMAX;
    bp_wsync();
    const PairRec* op = ev.pair[cur];
    for (int p = lane; p < np; p += BP_NL) {
        unsigned key = np_[p].key;
        unsigned gj = 0; int lv = -1; unsigned fl = 0; real sk = RC(0.0);
        if (carry) {
            int lo = 0, hi = old_n - 1;
            while (lo <= hi) {
                int mid = (lo + hi) >> 1;
                unsigned k = op[mid].key;
                if (k == key) { gj = op[mid].gjk; lv = op[mid].live; fl = op[mid].flags & 1; sk = op[mid].skip; break; }
                if (k < key) lo = mid + 1; else hi = mid - 1;
            }
        }
        np_[p].gjk = gj; np_[p].live = lv; np_[p].flags = fl; np_[p].skip = sk;
    }
    bp_wsync();
    rebuild_finish(ev, d, np, nxt, carry);
}

// The same list, built incrementally.  A pair of two bodies whose AABBs did not change since the last
// build passes the margin test now iff it passed then, so only pairs with a moved endpoint (ev.mvrb,
// set wherever a body integrates) are re-tested: |moved| x nbody box tests instead of nbody^2 / 2.
// The surviving old pairs and the fresh ones are merged by key, so the result is the list
// rebuild_pairs() would produce, entry for entry.  Falls back to the full build when many bodies moved.
constexpr int INC_MAX_MOVED = 64, INC_MAX_FRESH = 512;
BP_DEVNI void rebuild_pairs_inc(const EnvView& ev, const SimParams& P, const Dims& d) {
    EnvScal* sc = ev.sc;
    const int lane = bp_lane();
    const int nb = sc->nbody;
    const int cur = sc->cur, nxt = 1 - cur;
    const int old_n = sc->n_pairs;
    const PairRec* op = ev.pair[cur];
    PairRec* np_ = ev.pair[nxt];
    int* mlist = ev.queue;                       // moved bodies, ascending
    unsigned* fresh = (unsigned*)ev.queue0;      // fresh keys in discovery order
    unsigned* fsort = (unsigned*)ev.adj;         // fresh keys, ascending (the adjacency is rebuilt below)
    int* srank = ev.adj + d.PMAX;                // [old_n + 1] survivors before old entry p
    const int max_m = INC_MAX_MOVED < d.PMAX ? INC_MAX_MOVED : d.PMAX;
    const int max_f = INC_MAX_FRESH < d.PMAX ? INC_MAX_FRESH : d.PMAX;
    int nm = 0;
    for (int base = 0; base < nb; base += BP_NL) {
        int b = base + lane;
        bool f = (b < nb) && ev.mvrb[b];
        int tot; int rk = bp_wrank(f, &tot);
        if (f && nm + rk < max_m) mlist[nm + rk] = b;
        nm += tot;
    }
    bp_wsync();
    if (nm > max_m || old_n + 1 > d.PMAX) { rebuild_pairs(ev, P, d, true); return; }
    for (int k = lane; k < nm; k += BP_NL) { int b = mlist[k]; ev.b0[b] = ev.bb[b]; }
    bp_wsync();
    // fresh pairs: every pair with a moved endpoint that passes the margin test now
    int nf = 0;
    for (int k = 0; k < nm; k++) {
        const int m = mlist[k];
        real mm = body_margin(P, sc, m);
        BBox bm = ev.b0[m];
        real il = bm.l - mm, ib = bm.b - mm, ir = bm.r + mm, it = bm.t + mm;
        for (int base = 0; base < nb; base += BP_NL) {
            int x = base + lane;
            bool hit = false;
            if (x < nb && x != m && !(x < m && ev.mvrb[x])) {      // a pair of two moved bodies is emitted by the smaller id
                real mx = body_margin(P, sc, x);
                BBox bx = ev.b0[x];
                hit = (il <= bx.r + mx && bx.l - mx <= ir && ib <= bx.t + mx && bx.b - mx <= it);
            }
            int tot; int rk = bp_wrank(hit, &tot);
            if (hit && nf + rk < max_f) fresh[nf + rk] = (x < m) ? pair_key(x, m) : pair_key(m, x);
            nf += tot;
        }
    }
    bp_wsync();
    if (nf > max_f) { rebuild_pairs(ev, P, d, true); return; }
    // sort the fresh keys (all distinct): rank = number of smaller keys
    for (int e = lane; e < nf; e += BP_NL) {
        unsigned key = fresh[e];
        int r = 0;
        for (int q = 0; q < nf; q++) r += (fresh[q] < key) ? 1 : 0;
        fsort[r] = key;
    }
    // survivors of the old list: neither endpoint moved
    int ns = 0;
    for (int base = 0; base < old_n; base += BP_NL) {
        int p = base + lane;
        int sv = 0;
        if (p < old_n) { unsigned key = op[p].key; sv = (!ev.mvrb[key >> 16] && !ev.mvrb[key & 0xFFFF]) ? 1 : 0; }
        int tot; int ex = bp_wscan_excl(sv, &tot);
        if (p < old_n) srank[p] = sv ? (ns + ex) : -(ns + ex) - 1;
        ns += tot;
    }
    if (lane == 0) srank[old_n] = -ns - 1;
    bp_wsync();
    const int np = ns + nf;
    if (np > d.PMAX) { rebuild_pairs(ev, P, d, true); return; }
    for (int p = lane; p < old_n; p += BP_NL) {
        int sr = srank[p];
        if (sr < 0) continue;
        PairRec q = op[p];
        int lo = 0, hi = nf;                       // fresh keys below q.key
        while (lo < hi) { int mid = (lo + hi) >> 1; if (fsort[mid] < q.key) lo = mid + 1; else hi = mid; }
        q.flags &= 1u;
        np_[sr + lo] = q;
    }
    for (int i = lane; i < nf; i += BP_NL) {
        unsigned key = fsort[i];
        int lo = 0, hi = old_n;                    // first old entry with key >= this key
        while (lo < hi) { int mid = (lo + hi) >> 1; if (op[mid].key < key) lo = mid + 1; else hi = mid; }
        int sr = srank[lo]; if (sr < 0) sr = -sr - 1;
        PairRec q; q.key = key; q.gjk = 0; q.live = -1; q.flags = 0; q.skip = RC(0.0);
        if (lo < old_n && op[lo].key == key) { q.gjk = op[lo].gjk; q.live = op[lo].live; q.flags = op[lo].flags & 1; q.skip = op[lo].skip; }
        np_[sr + i] = q;
    }
    bp_wsync();
    rebuild_finish(ev, d, np, nxt, true);
}

// Append the pairs incident to body `bdy` that are not yet in the awake-pair list.
BP_DEV void add_body_pairs(const EnvView& ev, int bdy) {
    EnvScal* sc = ev.sc;
    const int lane = bp_lane();
    PairRec* pr = ev.pair[sc->cur];
    int s0 = ev.adj_start[bdy], s1 = ev.adj_start[bdy + 1];
    int n = sc->n_apairs;
    for (int base = s0; base < s1; base += BP_NL) {
        int e = base + lane;
        bool take = false; int p = -1;
        if (e < s1) { p = ev.adj[e]; take = !(pr[p].flags & 2); }
        int tot; int rk = bp_wrank(take, &tot);
        if (take) { ev.apairs[n + rk] = p; ev.ap_gate[n + rk] = -RC(1.0); pr[p].flags |= 2; }
        n += tot;
    }
    bp_wsync();
    if (lane == 0) sc->n_apairs = n;
    bp_wsync();
}
BP_DEVNI void build_apairs(const EnvView& ev) {
    EnvScal* sc = ev.sc;
    const int lane = bp_lane();
    PairRec* pr = ev.pair[sc->cur];
    for (int p = lane; p < sc->n_pairs; p += BP_NL) pr[p].flags &= ~2u;
    if (lane == 0) sc->n_apairs = 0;
    bp_wsync();
    int na = sc->n_awake;
    for (int k = 0; k < na; k++) add_body_pairs(ev, ev.awake[k]);
}

// ------------------------------------------------------------------ staged pair evaluation
// Stage 1 (gather): walk the awake-pair list; a pair needs work when it was never evaluated, or an
// endpoint moved and the pair is touching / cached, or an endpoint moved and the accumulated motion
// D[a]+D[b] has consumed the separation bound recorded when it was last found apart (otherwise it
// is provably still apart, and Chipmunk would produce no arbiter for it either).  Exact AABB filter
// (cpBBIntersects) is applied here; survivors are compacted into the GJK queue.
// Stage 2 (GJK): all lanes run GJK on queued pairs; pairs whose cores overlap are compacted into the
// EPA queue, the others get their manifold + arbiter update (cpArbiterUpdate).
// Stage 3 (EPA): all lanes run EPA on the overlapping pairs, then manifold + arbiter update.
// Compaction between the stages keeps lanes on the same code path (GJK/EPA are long and divergent).

// Not-touching outcome for pair p: slot bookkeeping + separation bound.
BP_DEV void pair_apart_impl(const EnvView& ev, EnvScal* sc, PairRec* pr, int i, int p, int a, int b, real gap) {
    pr[p].flags |= 1;
    double g = gap * RC(0.999999) - RC(1e-9);
    pr[p].skip = ev.body[a].D + ev.body[b].D + g;
    ev.ap_gate[i] = sc->G + g;
    int slot = pr[p].live;
    if (slot >= 0 && ev.slot[slot].touch) {
        // separated: the arbiter stays cached (contacts + impulses kept) for `persistence` steps
        ev.slot[slot].touch = 0;
        ev.slot[slot].stamp = sc->stamp - 1;
        bp_atomic_and(&ev.livemask[p >> 5], ~(1u << (p & 31)));
        bp_atomic_add(&sc->n_cached, 1);
        sc->dirty = 1;
    }
}

BP_DEVNI void pair_apart(const EnvView& ev, EnvScal* sc, PairRec* pr, int i, int p, int a, int b, real gap) { pair_apart_impl(ev, sc, pr, i, p, a, b, gap); }
BP_DEVNI void pair_apart_b(const DevConst* C, char* base, int i, int p, int a, int b, real gap) {
    EnvView ev = view_s(base);
    pair_apart_impl(ev, ev.sc, ev.pair[ev.sc->cur], i, p, a, b, gap);
}

// Touching / not-touching outcome from a closest-feature result.  Warp-convergent (slot allocation).
BP_DEVNI void pair_finish(const EnvView& ev, const LayoutView& L, const Dims& d, PairRec* pr, bool active, int i, int p, int a, int b, const Closest& pts) {
    EnvScal* sc = ev.sc;
    const int lane = bp_lane();
    Manifold m; m.count = 0;
    int slot = -1;
    if (active) {
        make_manifold(ev, L, a, b, pts, &m);
        pr[p].gjk = pts.id;
        slot = pr[p].live;
    }
    bool touching = active && m.count > 0;
    bool need = touching && slot < 0;
    int tot; int rk = bp_wrank(need, &tot);
    int nfree = sc->n_free;
    bool fresh = false;
    if (need) {
        if (rk < nfree) {
            slot = ev.free_stack[nfree - 1 - rk];
            bp_atomic_max(&sc->slot_hw, slot + 1);
            pr[p].live = slot;
            SlotRec& S = ev.slot[slot];
            S.pair = p; S.a = a; S.b = b; S.cnt = 0; S.touch = 0; S.first = 1;
            S.e = L.mat_e[a] * L.mat_e[b];
            S.u = L.mat_u[a] * L.mat_u[b];
            fresh = true;
        } else { bp_atomic_or((unsigned*)&sc->status, ST_LIVE_OVERFLOW); touching = false; }
    }
    bp_wsync();
    if (lane == 0 && tot) sc->n_free = nfree - (tot < nfree ? tot : nfree);
    if (active) {
        if (touching) {
            SlotRec& S = ev.slot[slot];
            bool was = !fresh && S.touch;
            real jn_new[2] = {0, 0}, jt_new[2] = {0, 0};
            int oc = fresh ? 0 : S.cnt;
            for (int i = 0; i < m.count; i++)
                for (int j = 0; j < oc; j++)
                    if (m.hash[i] == S.c[j].hash) { jn_new[i] = S.c[j].jn; jt_new[i] = S.c[j].jt; }
            real ax = ev.body[a].px, ay = ev.body[a].py, bx = ev.body[b].px, by = ev.body[b].py;
            for (int i = 0; i < m.count; i++) {
                ContactRec& c = S.c[i];
                c.r1x = m.p1[i].x - ax; c.r1y = m.p1[i].y - ay;
                c.r2x = m.p2[i].x - bx; c.r2y = m.p2[i].y - by;
                c.hash = m.hash[i]; c.jn = jn_new[i]; c.jt = jt_new[i];
            }
            S.cnt = m.count;
            S.nx = m.n.x; S.ny = m.n.y;
            if (!was) {
                // new arbiter, or CACHED -> FIRST_COLLISION (cpArbiterUpdate tail)
                S.first = 1; S.touch = 1;
                bp_atomic_or(&ev.livemask[p >> 5], 1u << (p & 31));
                if (!fresh) bp_atomic_add(&sc->n_cached, -1);
                sc->dirty = 1;
            }
            S.stamp = sc->stamp;
            pr[p].flags |= 1;
            pr[p].skip = -RC(1.0);
            ev.ap_gate[i] = -RC(1.0);
            // a touching pair with an awake endpoint wakes the other (dynamic) endpoint
            if ((!ev.is_awake[a] && body_dynamic(L, a)) || (!ev.is_awake[b] && body_dynamic(L, b))) sc->any_wake = 1;
        } else {
            pair_apart(ev, sc, pr, i, p, a, b, pts.d - L.rad[a] - L.rad[b]);
        }
    }
    bp_wsync();
}

// Evaluate the awake pairs [i0, n_apairs).  Returns the number of pairs that reached the GJK stage
// (they are left in ev.queue for the wake pass).
BP_DEVNI int eval_pairs(const EnvView& ev, const LayoutView& L, const Dims& d, int i0) {
    EnvScal* sc = ev.sc;
    const int lane = bp_lane();
    PairRec* pr = ev.pair[sc->cur];
    const int n = sc->n_apairs;
    // ---- stage 1a: coarse gate scan over contiguous per-entry gates (no pair can need work while G < gate)
    const double G = sc->G;
    int n0 = 0;
    for (int base = i0; base < n; base += BP_NL) {
        int i = base + lane;
        bool pass = (i < n) && (G >= ev.ap_gate[i]);
        int tot; int rk = bp_wrank(pass, &tot);
        if (pass) ev.queue0[n0 + rk] = i;
        n0 += tot;
    }
    bp_wsync();
    // ---- stage 1b: exact need test + AABB filter on the survivors
    int nq = 0, n_eval = 0, n_skip = 0;
    for (int base = 0; base < n0; base += BP_NL) {
        int k = base + lane;
        bool take = false; int i = -1;
        if (k < n0) {
            i = ev.queue0[k];
            int p = ev.apairs[i];
            PairRec q = pr[p];
            int a = q.key >> 16, b = q.key & 0xFFFF;
            bool mv = ev.moved[a] || ev.moved[b];
            double dsum = ev.body[a].D + ev.body[b].D;
            bool need = !(q.flags & 1) || (mv && (q.skip < RC(0.0) || dsum >= q.skip));
            if (!need) {
                if (q.skip >= RC(0.0)) { ev.ap_gate[i] = G + (q.skip - dsum); if (mv) n_skip++; }
            } else {
                n_eval++;
                BBox A = ev.bb[a], B = ev.bb[b];
                bool bbhit = (A.l <= B.r && B.l <= A.r && A.b <= B.t && B.b <= A.t);
                if (bbhit && !(L.minv[a] == RC(0.0) && L.minv[b] == RC(0.0))) take = true;
                else {
                    real gap = fmax(fmax(A.l - B.r, B.l - A.r), fmax(A.b - B.t, B.b - A.t));
                    pair_apart(ev, sc, pr, i, p, a, b, bbhit ? RC(0.0) : gap);
                }
            }
        }
        int tot; int rk = bp_wrank(take, &tot);
        if (take) ev.queue[nq + rk] = i;
        nq += tot;
    }
    bp_wsync();
    // ---- stage 2
    int ne = 0, gjk_it = 0;
    for (int base = 0; base < nq; base += BP_NL) {
        int q = base + lane;
        bool active = q < nq, to_epa = false;
        int p = -1, a = 0, b = 0;
        Closest pts; pts.d = 0; pts.id = 0; pts.n = mk(0, 0);
        unsigned hint = 0;
        int i = -1;
        if (active) {
            i = ev.queue[q];
            p = ev.apairs[i];
            unsigned key = pr[p].key;
            a = key >> 16; b = key & 0xFFFF;
            int capped = 0;
            to_epa = gjk_only(body_poly(ev, L, a), body_poly(ev, L, b), bb_centre(ev.bb[a]), bb_centre(ev.bb[b]), pr[p].gjk, &pts, &hint, &gjk_it, &capped);
            if (capped) bp_atomic_or((unsigned*)&sc->status, ST_GJK_ITER);
        }
        int tot; int rk = bp_wrank(to_epa, &tot);
        if (to_epa) {
            if (ne + rk < d.LMAX) { ev.queue2[ne + rk] = i; ev.qhint[ne + rk] = hint; }
            else { bp_atomic_or((unsigned*)&sc->status, ST_LIVE_OVERFLOW); }
        }
        ne += tot;
        pair_finish(ev, L, d, pr, active && !to_epa, i, p, a, b, pts);
    }
    if (ne > d.LMAX) ne = d.LMAX;
    // ---- stage 3
    for (int base = 0; base < ne; base += BP_NL) {
        int q = base + lane;
        bool active = q < ne;
        int p = -1, a = 0, b = 0;
        Closest pts; pts.d = 0; pts.id = 0; pts.n = mk(0, 0);
        int i = -1;
        if (active) {
            i = ev.queue2[q];
            p = ev.apairs[i];
            unsigned key = pr[p].key;
            a = key >> 16; b = key & 0xFFFF;
            int capped = 0;
            pts = gjk_epa(body_poly(ev, L, a), body_poly(ev, L, b), ev.qhint[q], &capped);
            if (capped) bp_atomic_or((unsigned*)&sc->status, ST_GJK_ITER);
        }
        pair_finish(ev, L, d, pr, active, i, p, a, b, pts);
    }
    // statistics only
    int t1, t2, t3;
    bp_wscan_excl(n_eval, &t1); bp_wscan_excl(n_skip, &t2); bp_wscan_excl(gjk_it, &t3);
    if (lane == 0) { sc->stat_pairs_eval += t1; sc->stat_skipped += t2; sc->stat_gjk += t3; sc->stat_epa += ne; }
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
            const PairRec* pr = ev.pair[sc->cur];
            for (int q = 0; q < nq; q++) {
                int p = ev.apairs[ev.queue[q]];
                int slot = pr[p].live;
                if (slot < 0 || !ev.slot[slot].touch) continue;
                unsigned key = pr[p].key;
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
        nq = eval_pairs(ev, L, d, ap0);
    }
}

// Free cached (separated) arbiters older than `persistence` steps (cpSpaceArbiterSetFilter).
BP_DEVNI void expire_cached(const EnvView& ev, const SimParams& P, const Dims& d) {
    EnvScal* sc = ev.sc;
    const int lane = bp_lane();
    PairRec* pr = ev.pair[sc->cur];
    const int hw = sc->slot_hw;
    for (int base = 0; base < hw; base += BP_NL) {
        int s = base + lane;
        bool drop = (s < hw) && ev.slot[s].pair >= 0 && !ev.slot[s].touch && (sc->stamp - ev.slot[s].stamp >= P.persistence);
        int tot; int rk = bp_wrank(drop, &tot);
        int nf = sc->n_free;
        if (drop) { pr[ev.slot[s].pair].live = -1; ev.slot[s].pair = -1; ev.free_stack[nf + rk] = s; }
        bp_wsync();
        if (lane == 0 && tot) { sc->n_free = nf + tot; sc->n_cached -= tot; }
        bp_wsync();
    }
}

// Enumerate touching arbiters in ascending pair id and order them by greedy colour.
BP_DEVNI void build_order(const EnvView& ev, const LayoutView& L, const Dims& d) {
    EnvScal* sc = ev.sc;
    const int lane = bp_lane();
    const PairRec* pr = ev.pair[sc->cur];
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
            if (pos < d.LMAX) ev.colour[pos] = pr[p].live;            // colour[] doubles as the canonical-order scratch
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
            int a = ev.slot[s].a, b = ev.slot[s].b;
            unsigned long long used = 0ull;
            bool da = body_dynamic(L, a), db = body_dynamic(L, b);
            if (da) used |= ev.colmask[a];
            if (db) used |= ev.colmask[b];
            int c = (~used == 0ull) ? MAXCOL : (__ffsll((long long)(~used)) - 1);
            if (c >= MAXCOL) { sc->status |= ST_COLOUR_OVERFLOW; c = MAXCOL - 1; }
            unsigned long long bit = 1ull << c;
            if (da) ev.colmask[a] |= bit;
            if (db) ev.colmask[b] |= bit;
            ev.free_stack[d.LMAX - 1 - i] = c;       // scratch: the top of the free stack is unused while n_free + tot <= LMAX
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
// The three solver pieces work on references (one arbiter, its two bodies), so that the env block in global memory and the
// shared-memory resident copy (sim_res.cuh) run the same arithmetic.  Only the first 112 bytes of a BodyRec are touched.
BP_DEV void arb_prestep_core(SlotRec& S, const BodyRec& A, const BodyRec& B, const SimParams& P) {
    V2 n = mk(S.nx, S.ny);
    V2 tn = perp(n);
    real ma = A.minv, ia = A.iinv, mb = B.minv, ib = B.iinv;
    V2 delta = mk(B.px - A.px, B.py - A.py);
    V2 va = mk(A.vx, A.vy), vb = mk(B.vx, B.vy);
    real wa = A.vw, wb = B.vw;
    int cnt = S.cnt;
#pragma unroll 1
    for (int i = 0; i < cnt; i++) {
        ContactRec& c = S.c[i];
        V2 r1 = mk(c.r1x, c.r1y), r2 = mk(c.r2x, c.r2y);
        real r1n = cross(r1, n), r2n = cross(r2, n);
        real kn = (ma + ia * r1n * r1n) + (mb + ib * r2n * r2n);
        real r1t = cross(r1, tn), r2t = cross(r2, tn);
        real kt = (ma + ia * r1t * r1t) + (mb + ib * r2t * r2t);
        c.nm = RC(1.0) / kn; c.tm = RC(1.0) / kt;
        real dist = dot((r2 - r1) + delta, n);
        c.bias = -P.bias_coef * fmin(RC(0.0), dist + P.slop) / P.dt;
        c.jb = RC(0.0);
        V2 vr = (vb + perp(r2) * wb) - (va + perp(r1) * wa);
        c.bounce = dot(vr, n) * S.e;
    }
}
BP_DEV void arb_prestep(const EnvView& ev, const LayoutView& L, const SimParams& P, int s) {
    SlotRec& S = ev.slot[s];
    arb_prestep_core(S, ev.body[S.a], ev.body[S.b], P);
}
BP_DEV void arb_warm_core(const SlotRec& S, BodyRec& A, BodyRec& B, real dt_coef) {
    if (S.first) return;
    V2 n = mk(S.nx, S.ny);
    real ma = A.minv, ia = A.iinv, mb = B.minv, ib = B.iinv;
    int cnt = S.cnt;
#pragma unroll 1
    for (int i = 0; i < cnt; i++) {
        const ContactRec& c = S.c[i];
        V2 r1 = mk(c.r1x, c.r1y), r2 = mk(c.r2x, c.r2y);
        V2 j = rotv(n, mk(c.jn, c.jt)) * dt_coef;
        V2 nj = neg(j);
        A.vx = A.vx + nj.x * ma; A.vy = A.vy + nj.y * ma; A.vw += ia * cross(r1, nj);
        B.vx = B.vx + j.x * mb; B.vy = B.vy + j.y * mb; B.vw += ib * cross(r2, j);
    }
}
BP_DEV void arb_warm(const EnvView& ev, const LayoutView& L, int s, real dt_coef) {
    SlotRec& S = ev.slot[s];
    arb_warm_core(S, ev.body[S.a], ev.body[S.b], dt_coef);
}
BP_DEV void arb_solve_core(SlotRec& S, BodyRec& A, BodyRec& B) {
    V2 n = mk(S.nx, S.ny);
    V2 tn = perp(n);
    real fr = S.u;
    real ma = A.minv, ia = A.iinv, mb = B.minv, ib = B.iinv;
    V2 va = mk(A.vx, A.vy), vb = mk(B.vx, B.vy);
    real wa = A.vw, wb = B.vw;
    V2 ba = mk(A.bx, A.by), bb = mk(B.bx, B.by);
    real bwa = A.bw, bwb = B.bw;
    int cnt = S.cnt;
#pragma unroll 1
    for (int i = 0; i < cnt; i++) {
        ContactRec& c = S.c[i];
        V2 r1 = mk(c.r1x, c.r1y), r2 = mk(c.r2x, c.r2y);
        real nMass = c.nm;
        V2 vb1 = ba + perp(r1) * bwa;
        V2 vb2 = bb + perp(r2) * bwb;
        V2 vr = (vb + perp(r2) * wb) - (va + perp(r1) * wa);
        real vbn = dot(vb2 - vb1, n);
        real vrn = dot(vr, n);
        real vrt = dot(vr, tn);
        real jbn = (c.bias - vbn) * nMass;
        real jbnOld = c.jb;
        real jb = fmax(jbnOld + jbn, RC(0.0));
        real jn = -(c.bounce + vrn) * nMass;
        real jnOld = c.jn;
        real jnAcc = fmax(jnOld + jn, RC(0.0));
        real jtMax = fr * jnAcc;
        real jt = -vrt * c.tm;
        real jtOld = c.jt;
        real jtAcc = clampd(jtOld + jt, -jtMax, jtMax);
        c.jb = jb; c.jn = jnAcc; c.jt = jtAcc;
        V2 jbv = n * (jb - jbnOld), njbv = neg(jbv);
        ba = ba + njbv * ma; bwa += ia * cross(r1, njbv);
        bb = bb + jbv * mb; bwb += ib * cross(r2, jbv);
        V2 j = rotv(n, mk(jnAcc - jnOld, jtAcc - jtOld)), nj = neg(j);
        va = va + nj * ma; wa += ia * cross(r1, nj);
        vb = vb + j * mb; wb += ib * cross(r2, j);
    }
    // kinematic / static bodies have m_inv = i_inv = 0: their velocities are unchanged (x + 0), skip the stores
    if (ma != RC(0.0) || ia != RC(0.0)) { A.vx = va.x; A.vy = va.y; A.vw = wa; A.bx = ba.x; A.by = ba.y; A.bw = bwa; }
    if (mb != RC(0.0) || ib != RC(0.0)) { B.vx = vb.x; B.vy = vb.y; B.vw = wb; B.bx = bb.x; B.by = bb.y; B.bw = bwb; }
}
BP_DEV void arb_solve(const EnvView& ev, const LayoutView& L, int s) {
    SlotRec& S = ev.slot[s];
    arb_solve_core(S, ev.body[S.a], ev.body[S.b]);
}

// Steps 3-6 of a substep for ONE env: prestep, velocity update, warm start, iterations.
BP_DEVNI void solve_env(const EnvView& ev, const LayoutView& L, const SimParams& P, real dt, real prev_dt) {
    EnvScal* sc = ev.sc;
    const int lane = bp_lane();
    // 3. prestep (uses the pre-damping velocities for the bounce term)
    int no = sc->n_order;
    for (int i = lane; i < no; i += BP_NL) arb_prestep(ev, L, P, ev.order[i]);
    bp_wsync();
    // 4. velocity update: v = v * damping^dt (gravity = 0, no forces); dynamic bodies only
    int na = sc->n_awake;
    for (int k = lane; k < na; k += BP_NL) {
        int b = ev.awake[k];
        if (body_dynamic(L, b)) {
            BodyRec& R = ev.body[b];
            R.vx = R.vx * P.damp + RC(0.0); R.vy = R.vy * P.damp + RC(0.0); R.vw = R.vw * P.damp + RC(0.0);
        }
    }
    bp_wsync();
    // 5. warm start + iterations, colour by colour (same-colour arbiters touch disjoint dynamic bodies)
    real dt_coef = (prev_dt == RC(0.0)) ? RC(0.0) : dt / prev_dt;
    int nc = sc->n_colours;
    for (int c = 0; c < nc; c++) {
        int c0 = ev.col_start[c], c1 = ev.col_start[c + 1];
        for (int i = c0 + lane; i < c1; i += BP_NL) arb_warm(ev, L, ev.order[i], dt_coef);
        bp_wsync();
    }
    for (int it = 0; it < P.iterations; it++) {
        for (int c = 0; c < nc; c++) {
            int c0 = ev.col_start[c], c1 = ev.col_start[c + 1];
            for (int i = c0 + lane; i < c1; i += BP_NL) arb_solve(ev, L, ev.order[i]);
            bp_wsync();
        }
    }
    // 6. arbiters solved this step are NORMAL next step
    for (int i = lane; i < no; i += BP_NL) ev.slot[ev.order[i]].first = 0;
    if (lane == 0) { sc->stat_substeps++; sc->stat_live_sum += no; sc->stat_awake_sum += na; }
    bp_wsync();
}

// ------------------------------------------------------------------ one substep (cpSpaceStep)
BP_DEVNI void substep_rest(const EnvView& ev, const LayoutView& L, const SimParams& P, const Dims& d, real dt, real prev_dt);
BP_DEVNI void substep(const EnvView& ev, const LayoutView& L, const SimParams& P, const Dims& d, real dt, real prev_dt) {
    EnvScal* sc = ev.sc;
    const int lane = bp_lane();
    BP_TICK(tk);
    if (lane == 0) { sc->stamp++; sc->any_wake = 0; }
    bp_wsync();
    // 1. integrate positions of the awake bodies (everything else has v = v_bias = 0 exactly)
    int na = sc->n_awake;
    real dmax = RC(0.0);
    for (int k = lane; k < na; k += BP_NL) {
        int b = ev.awake[k];
        BodyRec& R = ev.body[b];
        real vx = R.vx, vy = R.vy, vw = R.vw, bx = R.bx, by = R.by, bw = R.bw;
        bool mv = (vx != RC(0.0) || vy != RC(0.0) || vw != RC(0.0) || bx != RC(0.0) || by != RC(0.0) || bw != RC(0.0));
        if (mv) {
            real dx = (vx + bx) * dt, dy = (vy + by) * dt, da = (vw + bw) * dt;
            R.px = R.px + dx; R.py = R.py + dy; R.pa = R.pa + da;
            // motion bound of any point of the body (rotation about the cog), rounded up
            real dD = (fabs(dx) + fabs(dy) + L.rbound[b] * fabs(da)) * RC(1.000001);
            R.D += dD;
            dmax = fmax(dmax, dD);
            refresh_body(ev, L, b);
            real mg = body_margin(P, sc, b);
            BBox c = ev.bb[b], o = ev.b0[b];
            if (c.l < o.l - mg || c.r > o.r + mg || c.b < o.b - mg || c.t > o.t + mg) sc->rebuild = 1;
            ev.moved_step[b] = 1; ev.mvrb[b] = 1;
            R.bx = RC(0.0); R.by = RC(0.0); R.bw = RC(0.0);
        }
        ev.moved[b] = mv ? 1 : 0;
    }
    dmax = bp_wmax(dmax);
    if (lane == 0) sc->G += RC(2.0) * dmax * RC(1.000001);
    bp_wsync();
    BP_TOCK(sc, 0, tk);
    substep_rest(ev, L, P, d, dt, prev_dt);
}
// Everything of a substep after the position update: neighbour-list rebuild, collide, wake closure, cache expiry, solve.
BP_DEVNI void substep_rest(const EnvView& ev, const LayoutView& L, const SimParams& P, const Dims& d, real dt, real prev_dt) {
    EnvScal* sc = ev.sc;
    BP_TICK(tk);
    if (sc->rebuild) { rebuild_pairs_inc(ev, P, d); build_apairs(ev); }
    BP_TOCK(sc, 1, tk);
    // 2. collide, then close the awake set over newly touching bodies
    int nq = eval_pairs(ev, L, d, 0);
    BP_TOCK(sc, 2, tk);
    wake_closure(ev, L, d, nq);
    if (sc->n_cached > 0) expire_cached(ev, P, d);
    if (sc->dirty) build_order(ev, L, d);
    BP_TOCK(sc, 3, tk);
    // 3.-6. prestep, velocity update, warm start, iterations
    solve_env(ev, L, P, dt, prev_dt);
    BP_TOCK(sc, 5, tk);
}

// ------------------------------------------------------------------ awake-set maintenance
// Put to sleep every awake body whose complete contact component carries no velocity, no bias
// velocity, no cached impulse and does not touch the kinematic robot: the solver is then exactly a
// no-op for it.  Runs between env steps.
BP_DEVNI void sleep_pass(const EnvView& ev, const LayoutView& L, const Dims& d) {
    EnvScal* sc = ev.sc;
    const int lane = bp_lane();
    PairRec* pr = ev.pair[sc->cur];
    int na = sc->n_awake;
    for (int k = lane; k < na; k += BP_NL) {
        int b = ev.awake[k];
        const BodyRec& R = ev.body[b];
        bool nz = (R.vx != RC(0.0) || R.vy != RC(0.0) || R.vw != RC(0.0) || R.bx != RC(0.0) || R.by != RC(0.0) || R.bw != RC(0.0));
        ev.act[b] = (nz || !body_dynamic(L, b)) ? 1 : 0;
    }
    bp_wsync();
    const int hw = sc->slot_hw;
    for (int s = lane; s < hw; s += BP_NL) {
        const SlotRec& S = ev.slot[s];
        if (S.pair < 0) continue;
        bool seed = false;
        for (int i = 0; i < S.cnt; i++)
            if (S.c[i].jn != RC(0.0) || S.c[i].jt != RC(0.0) || S.c[i].jb != RC(0.0)) seed = true;
        if (S.touch && (!body_dynamic(L, S.a) || !body_dynamic(L, S.b))) seed = true;   // touches the robot
        if (seed) { ev.act[S.a] = 1; ev.act[S.b] = 1; }
    }
    bp_wsync();
    for (;;) {
        bool changed = false;
        for (int s = lane; s < hw; s += BP_NL) {
            const SlotRec& S = ev.slot[s];
            if (S.pair < 0 || !S.touch) continue;
            if (!body_dynamic(L, S.a) || !body_dynamic(L, S.b)) continue;
            if (ev.act[S.a] != ev.act[S.b]) { ev.act[S.a] = 1; ev.act[S.b] = 1; changed = true; }
        }
        bp_wsync();
        if (!bp_wany(changed)) break;
    }
    // free the arbiters of sleeping bodies, clear the evaluated flags of their pairs
    bool any_sleep = false;
    for (int base = 0; base < hw; base += BP_NL) {
        int s = base + lane;
        bool drop = false;
        if (s < hw && ev.slot[s].pair >= 0) {
            int a = ev.slot[s].a, b = ev.slot[s].b;
            drop = (body_dynamic(L, a) && !ev.act[a]) || (body_dynamic(L, b) && !ev.act[b]);
        }
        int tot; int rk = bp_wrank(drop, &tot);
        int nf = sc->n_free;
        if (drop) {
            int p = ev.slot[s].pair;
            if (ev.slot[s].touch) bp_atomic_and(&ev.livemask[p >> 5], ~(1u << (p & 31))); else bp_atomic_add(&sc->n_cached, -1);
            pr[p].live = -1; ev.slot[s].pair = -1; ev.slot[s].touch = 0; ev.free_stack[nf + rk] = s;
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
        for (int e = s0 + lane; e < s1; e += BP_NL) pr[ev.adj[e]].flags &= ~1u;
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
// poses6 (optional, [nbody][6] = body position x, y, angle, vx, vy, w): start from an externally given state instead of the
// layout's (bp_*_set_bodies: parity fixtures taken in the middle of an episode).
BP_DEVNI void init_env(const EnvView& ev, const LayoutView& L, const SimParams& P, const Dims& d, int layout_id, const double* start = nullptr,
                       const double* poses6 = nullptr) {
    EnvScal* sc = ev.sc;
    const int lane = bp_lane();
    const int nb = L.nbody;
    if (lane == 0) {
        sc->stamp = 0; sc->n_awake = nb; sc->n_pairs = 0; sc->cur = 0; sc->n_free = d.LMAX; sc->n_order = 0; sc->n_colours = 0;
        sc->dirty = 1; sc->status = 0; sc->layout = layout_id; sc->nbody = nb; sc->robot = L.robot0; sc->t = 0; sc->rebuild = 0;
        sc->n_apairs = 0; sc->n_cached = 0; sc->any_wake = 0; sc->total_work = RC(0.0); sc->G = RC(0.0);
        sc->wall_hit = 0; sc->has_prev_dist = 0; sc->prev_dist = RC(0.0); sc->slot_hw = 0; sc->pad_hw = 0;
        int nf = 0;
        for (int b = 0; b < nb; b++) nf += body_dynamic(L, b) ? 0 : 1;
        sc->n_fixed = nf;
    }
    for (int b = lane; b < nb; b += BP_NL) {
        BodyRec& R = ev.body[b];
        // every shape of the kinematic robot (bodies >= robot0) starts at the robot's pose
        double ipd[6] = {0, 0, 0, 0, 0, 0};                  // the layout's pose (real) or an injected one (double)
        if (poses6) { for (int q = 0; q < 6; q++) ipd[q] = poses6[6 * b + q]; }
        else if (start && b >= L.robot0) { for (int q = 0; q < 3; q++) ipd[q] = start[q]; }
        else { for (int q = 0; q < 3; q++) ipd[q] = (double)L.init[3 * b + q]; }
        const real ip[6] = {(real)ipd[0], (real)ipd[1], (real)ipd[2], (real)ipd[3], (real)ipd[4], (real)ipd[5]};
        real a = ip[2];
        real s, c; bp_sincos(a, &s, &c);
        real cx = L.cog[2 * b], cy = L.cog[2 * b + 1];
        // cpBodySetPosition: p = position + rotate(cog)
        R.px = (c * cx + (-s) * cy) + ip[0];
        R.py = (s * cx + c * cy) + ip[1];
        R.pa = a;
        R.vx = R.vy = R.vw = RC(0.0); R.bx = R.by = R.bw = RC(0.0); R.D = RC(0.0);
        if (poses6) { R.vx = ip[3]; R.vy = ip[4]; R.vw = ip[5]; }
        R.minv = L.minv[b]; R.iinv = L.iinv[b]; R.pad0 = R.pad1 = RC(0.0);
        refresh_body(ev, L, b);
        ev.awake[b] = b; ev.is_awake[b] = 1; ev.moved[b] = 0; ev.moved_step[b] = 0; ev.act[b] = 1;
    }
    for (int s = lane; s < d.LMAX; s += BP_NL) { ev.free_stack[s] = d.LMAX - 1 - s; ev.slot[s].pair = -1; ev.slot[s].touch = 0; ev.slot[s].cnt = 0; }
    bp_wsync();
    rebuild_pairs(ev, P, d, false);
    build_apairs(ev);
}

}  // namespace bp
