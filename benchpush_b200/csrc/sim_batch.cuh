// Batched env step: ONE WARP ADVANCES A BATCH OF BK ENVIRONMENTS TOGETHER.
//
// Why (profiles/r01*): with one env per warp the ship-ice substep keeps ~5 of 32 lanes busy (a colour
// of the contact graph has 2-5 arbiters, 5-15 pairs need a narrowphase) and the kernel is bound by
// instruction fetch, not by HBM or the ALUs.  Environments are independent, so the work items of
// several envs can share a warp instruction: every hot phase of the substep below runs over the
// CONCATENATION of the per-env item lists of the batch (awake bodies, awake pairs, arbiters of one
// colour), one item per lane, so one fetched instruction serves up to 32 items of up to BK envs.
// Per-env results are bit-identical to the one-env-per-warp path (sim_step.cuh): every item
// executes the same arithmetic, arbiters of one env are still solved colour by colour in canonical
// order, and envs never share memory.  Rare per-env events (neighbour-list rebuild, wake closure,
// cache expiry, recolouring, sleep) reuse the warp-cooperative routines of sim_step.cuh, one env at
// a time.
// This header may be included more than once with different (BP_BK, BP_BATCH_NS): the library carries the batched step for
// BK = 8 (full GPUs) and BK = 4 (the small per-GPU batches of the 8-GPU strong-scaling point: more warps for the same envs;
// measured +22 % at 8 192 envs, equal at 16 384 — profiles/r01f_shipice_bk_small_batches.txt).  Results do not depend on BK.
#include "maze_env.cuh"

#ifndef BP_BK
#define BP_BK 8
#endif
#ifndef BP_BATCH_NS
#define BP_BATCH_NS bk_default
#endif

namespace bp {
namespace BP_BATCH_NS {

constexpr int BK = BP_BK;        // envs per warp batch
#ifndef BP_BQ
#define BP_BQ 256
#endif
#ifndef BP_BQE
#define BP_BQE 96
#endif
#ifndef BP_BWL
#define BP_BWL 512
#endif
constexpr int BQ = BP_BQ;        // per-warp narrowphase queue (entries)
constexpr int BQE = BP_BQE;      // per-warp EPA queue
constexpr int BWL = BP_BWL;      // merged solver work list (arbiters of the whole batch; a longer list is solved env by env)
static_assert(BQ >= 2 * BP_NL && BQE >= 2 * BP_NL && BQ % BP_NL == 0 && BQE % BP_NL == 0, "queues are filled a warp at a time");
constexpr int BCOL = 16;         // colours the merged solve handles (more -> per-env fallback)

struct BatchShared {             // one per warp, in shared memory
    char* base[BK];              // env blocks of the batch; nullptr = slot not stepping
    LayoutView L[BK];
    int env_id[BK];
    int off_awake[BK + 1], off_ap[BK + 1], off_ord[BK + 1];
    unsigned long long dmax_bits[BK];
    unsigned short ccnt[BK][BCOL];   // arbiters of env e with colour c
    unsigned short cpre[BK][BCOL];          // start of env e's segment inside the merged colour-c segment
    int off_col[BCOL + 1];
    int n_items;                 // merged list length, -1 = does not fit (per-env fallback)
    int n_col, wl_valid;
    int skip_res;                // resident step (sim_res.cuh): gate == -2 marks a touching pair that is evaluated from shared memory
    unsigned char flags[BK];     // bit0 yaw violated, bit1 boundary violated (ship_ice_env.py:284-290)
    int st_sub[BK], st_live[BK], st_awake[BK];   // per-step sums of the schedule's cost counters (EnvScal::stat_*), flushed once per env step
    unsigned wl[BWL];            // (e << 16) | slot, grouped by colour
    unsigned q0[BQ], q1[BQ];     // (e << 24) | index into apairs
    unsigned q2[BQE], q2hint[BQE];
};

BP_DEV int find_env(const int* off, int idx) {
    int e = 0;
#pragma unroll
    for (int j = 1; j < BK; j++) e += (idx >= off[j]) ? 1 : 0;
    return e;
}

// Touching / not-touching outcome for ONE lane's pair (no warp-collective inside: lanes of one call
// may belong to different envs).  Same state transitions as pair_finish (cpArbiterUpdate).
BP_DEVNI void pair_finish_lane(const DevConst* C, char* base, const LayoutView& L, int i, int p, int a, int b, const Closest& pts) {
    EnvView ev = view_s(base);
    EnvScal* sc = ev.sc;
    PairRec* pr = ev.pair[sc->cur];
    Manifold m; m.count = 0;
    make_manifold_b(C, base, L, a, b, pts, &m);
    pr[p].gjk = pts.id;
    int slot = pr[p].live;
    bool touching = m.count > 0;
    bool fresh = false;
    if (touching && slot < 0) {
        int k = bp_atomic_add(&sc->n_free, -1) - 1;
        if (k >= 0) {
            slot = ev.free_stack[k];
            bp_atomic_max(&sc->slot_hw, slot + 1);
            pr[p].live = slot;
            SlotRec& S = ev.slot[slot];
            S.pair = p; S.a = a; S.b = b; S.cnt = 0; S.touch = 0; S.first = 1;
            S.e = L.mat_e[a] * L.mat_e[b];
            S.u = L.mat_u[a] * L.mat_u[b];
            fresh = true;
        } else {
            bp_atomic_add(&sc->n_free, 1);
            bp_atomic_or((unsigned*)&sc->status, ST_LIVE_OVERFLOW);
            touching = false;
        }
    }
    if (touching) {
        SlotRec& S = ev.slot[slot];
        bool was = !fresh && S.touch;
        real jn_new[2] = {0, 0}, jt_new[2] = {0, 0};
        int oc = fresh ? 0 : S.cnt;
        for (int x = 0; x < m.count; x++)
            for (int j = 0; j < oc; j++)
                if (m.hash[x] == S.c[j].hash) { jn_new[x] = S.c[j].jn; jt_new[x] = S.c[j].jt; }
        real ax = ev.body[a].px, ay = ev.body[a].py, bx = ev.body[b].px, by = ev.body[b].py;
        for (int x = 0; x < m.count; x++) {
            ContactRec& c = S.c[x];
            c.r1x = m.p1[x].x - ax; c.r1y = m.p1[x].y - ay;
            c.r2x = m.p2[x].x - bx; c.r2y = m.p2[x].y - by;
            c.hash = m.hash[x]; c.jn = jn_new[x]; c.jt = jt_new[x];
        }
        S.cnt = m.count;
        S.nx = m.n.x; S.ny = m.n.y;
        if (!was) {
            S.first = 1; S.touch = 1;
            bp_atomic_or(&ev.livemask[p >> 5], 1u << (p & 31));
            if (!fresh) bp_atomic_add(&sc->n_cached, -1);
            sc->dirty = 1;
        }
        S.stamp = sc->stamp;
        pr[p].flags |= 1;
        pr[p].skip = -RC(1.0);
        ev.ap_gate[i] = -RC(1.0);
        if ((!ev.is_awake[a] && body_dynamic(L, a)) || (!ev.is_awake[b] && body_dynamic(L, b))) sc->any_wake = 1;
    } else {
        pair_apart_b(C, base, i, p, a, b, pts.d - L.rad[a] - L.rad[b]);
    }
}

// Wake every asleep dynamic body that shares a touching arbiter with an awake body, evaluate its
// pairs, repeat until closed.  Same closure as wake_closure(); it finds the candidates by scanning
// the awake-pair list (a touching pair with an asleep dynamic endpoint can only have appeared in
// this substep: the awake set is closed under "touching" at substep start).
BP_DEVNI void wake_closure_scan(const EnvView& ev, const LayoutView& L, const Dims& d) {
    EnvScal* sc = ev.sc;
    const int lane = bp_lane();
    while (sc->any_wake) {
        bp_wsync();
        const int first_new = sc->n_awake;
        const int ap0 = sc->n_apairs;
        const PairRec* pr = ev.pair[sc->cur];
        bp_wsync();
        if (lane == 0) sc->any_wake = 0;
        for (int bs = 0; bs < ap0; bs += BP_NL) {
            int i = bs + lane;
            bool hit = false; int a = 0, b = 0;
            if (i < ap0) {
                int p = ev.apairs[i];
                int slot = pr[p].live;
                if (slot >= 0 && ev.slot[slot].touch) {
                    unsigned key = pr[p].key;
                    a = (int)(key >> 16); b = (int)(key & 0xFFFF);
                    hit = (!ev.is_awake[a] && body_dynamic(L, a)) || (!ev.is_awake[b] && body_dynamic(L, b));
                }
            }
            unsigned m = bp_wballot(hit);
            while (m) {
                int src = __ffsll((long long)m) - 1; m &= m - 1;
                int xa = bp_wshfl(a, src), xb = bp_wshfl(b, src);
                if (lane == 0) {
                    int na = sc->n_awake;
                    int ab[2] = {xa, xb};
                    for (int t = 0; t < 2; t++) {
                        int x = ab[t];
                        if (!ev.is_awake[x] && body_dynamic(L, x)) { ev.is_awake[x] = 1; ev.moved[x] = 0; ev.awake[na++] = x; }
                    }
                    sc->n_awake = na;
                }
                bp_wsync();
            }
        }
        bp_wsync();
        int na = sc->n_awake;
        for (int k = first_new; k < na; k++) add_body_pairs(ev, ev.awake[k]);
        eval_pairs(ev, L, d, ap0);
    }
}

// EPA stage of the batched narrowphase: the queued pairs have overlapping cores.
BP_DEV void batch_flush_epa(const DevConst* C, BatchShared& S, int n2) {
    const int lane = bp_lane();
    for (int bs = 0; bs < n2; bs += BP_NL) {
        int q = bs + lane;
        if (q < n2) {
            unsigned ent = S.q2[q];
            int e = (int)(ent >> 24), i = (int)(ent & 0xFFFFFF);
            char* base = S.base[e];
            EnvView ev = view_s(base);
            EnvScal* sc = ev.sc;
            const LayoutView& L = S.L[e];
            PairRec* pr = ev.pair[sc->cur];
            int p = ev.apairs[i];
            unsigned key = pr[p].key;
            int a = (int)(key >> 16), b = (int)(key & 0xFFFF);
            int capped = 0;
            Closest pts = gjk_epa(body_poly(ev, L, a), body_poly(ev, L, b), S.q2hint[q], &capped);
            if (capped) bp_atomic_or((unsigned*)&sc->status, ST_GJK_ITER);
            pair_finish_lane(C, base, L, i, p, a, b, pts);
        }
    }
    bp_wsync();
}

// Concatenate the per-env colour classes into one work list per colour.
BP_DEVNI void batch_build_worklist(const DevConst* C, BatchShared& S) {
    const int lane = bp_lane();
    for (int t = lane; t < BK * BCOL; t += BP_NL) {
        int e = t / BCOL, c = t % BCOL, v = 0;
        if (S.base[e]) {
            EnvView ev = view_s(S.base[e]);
            if (c < ev.sc->n_colours) v = ev.col_start[c + 1] - ev.col_start[c];
        }
        S.ccnt[e][c] = (unsigned short)v;
    }
    bp_wsync();
    if (lane == 0) {
        int acc = 0, ncol = 0, tot_items = 0;
        bool fits = true;
        for (int e = 0; e < BK; e++) {
            S.off_ord[e] = tot_items;
            if (S.base[e]) {
                const EnvScal* sc = (const EnvScal*)S.base[e];
                if (sc->n_colours > BCOL) fits = false;
                if (sc->n_colours > ncol) ncol = sc->n_colours;
                tot_items += sc->n_order;
            }
        }
        S.off_ord[BK] = tot_items;
        if (tot_items > BWL || C->d.LMAX > 65535) fits = false;
        if (ncol > BCOL) ncol = BCOL;
        for (int c = 0; c < ncol; c++) {
            S.off_col[c] = acc;
            for (int e = 0; e < BK; e++) { S.cpre[e][c] = (unsigned short)acc; acc += S.ccnt[e][c]; }
        }
        S.off_col[ncol] = acc;
        S.n_col = ncol;
        S.n_items = fits ? acc : -1;
        S.wl_valid = 1;
    }
    bp_wsync();
    if (S.n_items < 0) return;
    const int tot = S.off_ord[BK];
    for (int j = lane; j < tot; j += BP_NL) {
        int e = find_env(S.off_ord, j), pos = j - S.off_ord[e];
        EnvView ev = view_s(S.base[e]);
        int nc = ev.sc->n_colours, c = 0;
        while (c + 1 < nc && ev.col_start[c + 1] <= pos) c++;
        S.wl[S.cpre[e][c] + (pos - ev.col_start[c])] = ((unsigned)e << 16) | (unsigned)ev.order[pos];
    }
    bp_wsync();
}

// Collide phase of a substep for every env of the batch with S.base[e] != nullptr: gate scan -> need test -> GJK -> (EPA) over
// the concatenated awake-pair lists.
BP_DEVNI void batch_collide(const DevConst* C, BatchShared& S) {
    const int lane = bp_lane();
    if (lane == 0) {
        int acc = 0;
        for (int e = 0; e < BK; e++) { S.off_ap[e] = acc; if (S.base[e]) acc += ((const EnvScal*)S.base[e])->n_apairs; }
        S.off_ap[BK] = acc;
    }
    bp_wsync();
    {
        const int total = S.off_ap[BK];
        int cursor = 0;
        while (cursor < total) {
            // 2a. coarse gate scan (no pair can need work while G < gate), fill q0
            int n0 = 0;
            while (cursor < total && n0 <= BQ - BP_NL) {
                int i = cursor + lane;
                bool pass = false; unsigned ent = 0;
                if (i < total) {
                    int e = find_env(S.off_ap, i), k = i - S.off_ap[e];
                    EnvView ev = view_s(S.base[e]);
                    double gt = ev.ap_gate[k];
                    pass = ev.sc->G >= gt && !(S.skip_res && gt == -RC(2.0));
                    ent = ((unsigned)e << 24) | (unsigned)k;
                }
                int t; int rk = bp_wrank(pass, &t);
                if (pass) S.q0[n0 + rk] = ent;
                n0 += t; cursor += BP_NL;
            }
            bp_wsync();
            // 2b. exact need test + AABB filter
            int n1 = 0;
            for (int bs = 0; bs < n0; bs += BP_NL) {
                int q = bs + lane;
                bool take = false; unsigned ent = 0;
                if (q < n0) {
                    ent = S.q0[q];
                    int e = (int)(ent >> 24), i = (int)(ent & 0xFFFFFF);
                    char* base = S.base[e];
                    EnvView ev = view_s(base);
                    EnvScal* sc = ev.sc;
                    const LayoutView& L = S.L[e];
                    PairRec* pr = ev.pair[sc->cur];
                    int p = ev.apairs[i];
                    PairRec pq = pr[p];
                    int a = (int)(pq.key >> 16), b = (int)(pq.key & 0xFFFF);
                    bool mv = ev.moved[a] || ev.moved[b];
                    double dsum = ev.body[a].D + ev.body[b].D;
                    bool need = !(pq.flags & 1) || (mv && (pq.skip < RC(0.0) || dsum >= pq.skip));
                    if (!need) {
                        if (pq.skip >= RC(0.0)) ev.ap_gate[i] = sc->G + (pq.skip - dsum);
                    } else {
                        BBox A = ev.bb[a], B = ev.bb[b];
                        bool bbhit = (A.l <= B.r && B.l <= A.r && A.b <= B.t && B.b <= A.t);
                        if (bbhit && !(L.minv[a] == RC(0.0) && L.minv[b] == RC(0.0))) {
                            take = true;
                            // the narrowphase below scans both polygons several times: start the loads now
                            bp_prefetch(ev.wv + (size_t)a * VMAX * 2, 3);
                            bp_prefetch(ev.wv + (size_t)b * VMAX * 2, 3);
                            if (pq.live >= 0) bp_prefetch(&ev.slot[pq.live], 2);
                        } else {
                            real gap = fmax(fmax(A.l - B.r, B.l - A.r), fmax(A.b - B.t, B.b - A.t));
                            pair_apart_b(C, base, i, p, a, b, bbhit ? RC(0.0) : gap);
                        }
                    }
                }
                int t; int rk = bp_wrank(take, &t);
                if (take) S.q1[n1 + rk] = ent;
                n1 += t;
            }
            bp_wsync();
            // 2c. GJK; pairs with overlapping cores go on to the EPA queue
            int n2 = 0;
            for (int bs = 0; bs < n1; bs += BP_NL) {
                int q = bs + lane;
                bool to_epa = false; unsigned ent = 0, hint = 0;
                if (q < n1) {
                    ent = S.q1[q];
                    int e = (int)(ent >> 24), i = (int)(ent & 0xFFFFFF);
                    char* base = S.base[e];
                    EnvView ev = view_s(base);
                    EnvScal* sc = ev.sc;
                    const LayoutView& L = S.L[e];
                    PairRec* pr = ev.pair[sc->cur];
                    int p = ev.apairs[i];
                    unsigned key = pr[p].key;
                    int a = (int)(key >> 16), b = (int)(key & 0xFFFF);
                    int capped = 0, iters = 0;
                    Closest pts; pts.d = 0; pts.id = 0; pts.n = mk(0, 0);
                    to_epa = gjk_only(body_poly(ev, L, a), body_poly(ev, L, b), bb_centre(ev.bb[a]), bb_centre(ev.bb[b]), pr[p].gjk, &pts, &hint, &iters, &capped);
                    if (capped) bp_atomic_or((unsigned*)&sc->status, ST_GJK_ITER);
                    if (!to_epa) pair_finish_lane(C, base, L, i, p, a, b, pts);
                }
                int t; int rk = bp_wrank(to_epa, &t);
                if (to_epa) { S.q2[n2 + rk] = ent; S.q2hint[n2 + rk] = hint; }
                n2 += t;
                if (n2 > BQE - BP_NL) { bp_wsync(); batch_flush_epa(C, S, n2); n2 = 0; }
            }
            bp_wsync();
            if (n2 > 0) batch_flush_epa(C, S, n2);
        }
    }
    bp_wsync();
}

// One physics substep (cpSpaceStep) for every env of the batch.
BP_DEVNI void batch_substep(const DevConst* C, BatchShared& S, real dt, real prev_dt) {
    const int lane = bp_lane();
    const SimParams& P = C->P;
    const Dims& d = C->d;
    // ---- 0. stamps, offsets of the concatenated awake list
    for (int e = lane; e < BK; e += BP_NL)
        if (S.base[e]) { EnvScal* sc = (EnvScal*)S.base[e]; sc->stamp++; sc->any_wake = 0; S.dmax_bits[e] = 0ull; }
    if (lane == 0) {
        int acc = 0;
        for (int e = 0; e < BK; e++) { S.off_awake[e] = acc; if (S.base[e]) acc += ((const EnvScal*)S.base[e])->n_awake; }
        S.off_awake[BK] = acc;
    }
    bp_wsync();
    // ---- 1. integrate positions of the awake bodies
    {
        const int tot = S.off_awake[BK];
        for (int i = lane; i < tot; i += BP_NL) {
            int e = find_env(S.off_awake, i), k = i - S.off_awake[e];
            char* base = S.base[e];
            EnvView ev = view_s(base);
            EnvScal* sc = ev.sc;
            const LayoutView& L = S.L[e];
            int b = ev.awake[k];
            BodyRec& R = ev.body[b];
            bp_prefetch(L.lv + (size_t)b * VMAX * 2, 2);
            real vx = R.vx, vy = R.vy, vw = R.vw, bx = R.bx, by = R.by, bw = R.bw;
            bool mv = (vx != RC(0.0) || vy != RC(0.0) || vw != RC(0.0) || bx != RC(0.0) || by != RC(0.0) || bw != RC(0.0));
            if (mv) {
                real dx = (vx + bx) * dt, dy = (vy + by) * dt, da = (vw + bw) * dt;
                R.px = R.px + dx; R.py = R.py + dy; R.pa = R.pa + da;
                real dD = (fabs(dx) + fabs(dy) + L.rbound[b] * fabs(da)) * RC(1.000001);
                R.D += dD;
                bp_atomic_max_u64(&S.dmax_bits[e], bp_d2bits(dD));
                refresh_body_b(C, base, L, b);
                real mg = body_margin(P, sc, b);
                BBox c = ev.bb[b], o = ev.b0[b];
                if (c.l < o.l - mg || c.r > o.r + mg || c.b < o.b - mg || c.t > o.t + mg) sc->rebuild = 1;
                ev.moved_step[b] = 1; ev.mvrb[b] = 1;
                R.bx = RC(0.0); R.by = RC(0.0); R.bw = RC(0.0);
            }
            ev.moved[b] = mv ? 1 : 0;
        }
    }
    bp_wsync();
    for (int e = lane; e < BK; e += BP_NL)
        if (S.base[e]) { EnvScal* sc = (EnvScal*)S.base[e]; sc->G += RC(2.0) * bp_bits2d(S.dmax_bits[e]) * RC(1.000001); }
    bp_wsync();
    // ---- rare: neighbour-list rebuild, one env at a time
    for (int e = 0; e < BK; e++) {
        char* base = S.base[e];
        if (base && ((const EnvScal*)base)->rebuild) {
            EnvView ev = view_s(base);
            rebuild_pairs_inc(ev, P, d);
            build_apairs(ev);
        }
    }
    batch_collide(C, S);
    // ---- rare per-env events: wake closure, cache expiry, recolouring
    for (int e = 0; e < BK; e++) {
        char* base = S.base[e];
        if (!base) continue;
        const EnvScal* sc = (const EnvScal*)base;
        if (sc->any_wake || sc->n_cached > 0 || sc->dirty) {
            EnvView ev = view_s(base);
            if (sc->any_wake) wake_closure_scan(ev, S.L[e], d);
            if (sc->n_cached > 0) expire_cached(ev, P, d);
            if (sc->dirty) { build_order(ev, S.L[e], d); if (lane == 0) S.wl_valid = 0; }
        }
    }
    bp_wsync();
    if (!S.wl_valid) batch_build_worklist(C, S);
    const real dt_coef = (prev_dt == RC(0.0)) ? RC(0.0) : dt / prev_dt;
    if (S.n_items < 0) {
        // the merged list does not fit: solve env by env (same arithmetic, same order)
        for (int e = 0; e < BK; e++)
            if (S.base[e]) { EnvView ev = view_s(S.base[e]); solve_env(ev, S.L[e], P, dt, prev_dt); }
        return;
    }
    // ---- 3. prestep over all arbiters of the batch
    const int n_items = S.n_items;
    for (int j = lane; j < n_items; j += BP_NL) {
        unsigned ent = S.wl[j];
        int e = (int)(ent >> 16);
        arb_prestep(view_s(S.base[e]), S.L[e], P, (int)(ent & 0xFFFF));
    }
    // ---- 4. velocity update of the awake dynamic bodies (the awake lists may have grown)
    if (lane == 0) {
        int acc = 0;
        for (int e = 0; e < BK; e++) { S.off_awake[e] = acc; if (S.base[e]) acc += ((const EnvScal*)S.base[e])->n_awake; }
        S.off_awake[BK] = acc;
    }
    bp_wsync();
    {
        const int tot = S.off_awake[BK];
        for (int i = lane; i < tot; i += BP_NL) {
            int e = find_env(S.off_awake, i), k = i - S.off_awake[e];
            EnvView ev = view_s(S.base[e]);
            int b = ev.awake[k];
            if (body_dynamic(S.L[e], b)) {
                BodyRec& R = ev.body[b];
                R.vx = R.vx * P.damp + RC(0.0); R.vy = R.vy * P.damp + RC(0.0); R.vw = R.vw * P.damp + RC(0.0);
            }
        }
    }
    bp_wsync();
    // ---- 5. warm start + iterations, colour by colour over the merged lists
    const int nc = S.n_col;
    for (int c = 0; c < nc; c++) {
        int c0 = S.off_col[c], c1 = S.off_col[c + 1];
        for (int j = c0 + lane; j < c1; j += BP_NL) {
            unsigned ent = S.wl[j];
            int e = (int)(ent >> 16);
            arb_warm(view_s(S.base[e]), S.L[e], (int)(ent & 0xFFFF), dt_coef);
        }
        bp_wsync();
    }
    for (int it = 0; it < P.iterations; it++) {
        for (int c = 0; c < nc; c++) {
            int c0 = S.off_col[c], c1 = S.off_col[c + 1];
            for (int j = c0 + lane; j < c1; j += BP_NL) {
                unsigned ent = S.wl[j];
                int e = (int)(ent >> 16);
                arb_solve(view_s(S.base[e]), S.L[e], (int)(ent & 0xFFFF));
            }
            bp_wsync();
        }
    }
    // ---- 6. arbiters solved this step are NORMAL next step
    for (int j = lane; j < n_items; j += BP_NL) {
        unsigned ent = S.wl[j];
        view_s(S.base[ent >> 16]).slot[ent & 0xFFFF].first = 0;
    }
    for (int e = lane; e < BK; e += BP_NL)
        if (S.base[e]) { const EnvScal* sc = (const EnvScal*)S.base[e]; S.st_sub[e]++; S.st_live[e] += sc->n_order; S.st_awake[e] += sc->n_awake; }
    bp_wsync();
}

// ShipIceEnv.step for the envs [env0, env0 + BK) of a handle (those >= n_envs are skipped).
// auto_reset: an env whose previous step terminated is reset instead (next trial) and returns the
// reset observation with reward 0, exactly like the one-env path in bp_api.cu.
// perm (optional): the step schedule, entry i = env id or -1 (k_schedule in bp_api.cu); the batch is perm[env0 .. env0 + BK).
BP_DEVNI void shipice_step_batch(const DevConst* C, BatchShared& S, char* mem, size_t env_bytes, int env0, const float* actions,
                                 const unsigned char* mask, const StepOut& out, int auto_reset, const int* perm = nullptr) {
    const int lane = bp_lane();
    const SimParams& P = C->P;
    const Dims& d = C->d;
    for (int e = 0; e < BK; e++) {
        int env = perm ? perm[env0 + e] : env0 + e;
        char* base = nullptr;
        if (env >= 0 && env < d.n_envs && !(mask && !mask[env])) {
            base = mem + env_bytes * (size_t)env;
            EnvView ev = view_s(base);
            EnvScal* sc = ev.sc;
            if (auto_reset && sc->done) {
                int ep = sc->episode + 1, first = sc->first_layout;
                int layout = (first + ep) % d.n_layouts;
                const double* sp = C->start_pose ? C->start_pose + 3 * (size_t)env : nullptr;
                if (C->snap && !sp) shipice_reset_from_snapshot(C, base, layout, env, out);
                else shipice_reset_env(ev, C->T, P, d, layout, env, out, sp);
                if (lane == 0) { sc->first_layout = first; sc->episode = ep; sc->done = 0; sc->obs_raw = 1; }
                base = nullptr;
            } else {
                LayoutView L = layout_view(C->T, d, sc->layout);
                if (lane == 0) S.L[e] = L;
                shipice_step_begin(ev, L, P, d, (real)actions[env]);
            }
        }
        if (lane == 0) { S.base[e] = base; S.env_id[e] = env; S.flags[e] = 0; S.st_sub[e] = 0; S.st_live[e] = 0; S.st_awake[e] = 0; }
        bp_wsync();
    }
    if (lane == 0) { S.wl_valid = 0; S.skip_res = 0; }
    bp_wsync();
    bool any = false;
    for (int e = 0; e < BK; e++) any = any || (S.base[e] != nullptr);
    if (any) {
        for (int k = 0; k < P.substeps; k++) {
            batch_substep(C, S, P.dt, P.dt);
            for (int e = lane; e < BK; e += BP_NL)
                if (S.base[e]) {
                    EnvView ev = view_s(S.base[e]);
                    S.flags[e] |= (unsigned char)env_substep_checks(ev, S.L[e], P, ev.sc->robot);
                }
            bp_wsync();
        }
    }
    for (int e = 0; e < BK; e++) {
        char* base = S.base[e];
        if (!base) continue;
        EnvView ev = view_s(base);
        int env = S.env_id[e];
        env_step_finish(ev, S.L[e], P, d, env, out, (int)S.flags[e]);
        if (lane == 0) {
            EnvScal* sc = ev.sc;
            sc->done = out.terminated[env]; sc->obs_raw = 0;
            sc->stat_substeps += S.st_sub[e]; sc->stat_live_sum += S.st_live[e]; sc->stat_awake_sum += S.st_awake[e];
        }
        bp_wsync();
    }
}

}  // namespace BP_BATCH_NS
}  // namespace bp
