// Resident env step: THE AWAKE WORLD OF A BATCH OF ENVIRONMENTS LIVES IN SHARED MEMORY FOR THE 400 SUBSTEPS OF AN ENV STEP.
//
// Replaces the per-env `for _ in range(steps): space.step(dt)` loop of the reference (ship_ice_env.py:280-290) like
// sim_batch.cuh does, with the same arithmetic and the same solver order (results are bit-identical), but a different home
// for the data.  Measured on B200 (profiles/r01g_*): the batched kernel re-streams ~57 KB of per-env state from HBM every
// substep (22.8 MB per env step, 112x the algorithmic bytes) and waits on it (long_scoreboard 22.6 of 30.5 cycles per issue),
// although a substep only works on the awake set: ~11 bodies and ~9 touching arbiters per env at concentration 0.5.
//
// Here one warp (= one thread block) owns RK envs and a pool in shared memory:
//   rb / rx  body records of the awake bodies (pose, velocities, bias velocities, 1/m, 1/I, cos/sin, motion bound; AABB,
//            neighbour-list limits, radius, centre of gravity)                                        224 B per body
//   rv       their world vertices                                                                      16 B per vertex
//   rs       the touching arbiters (cpArbiter: normal, <= 2 contacts with accumulated impulses)       272 B per arbiter
// The pool is filled from the env blocks with cp.async.bulk copies that complete on an mbarrier (TMA linear copies: one per
// body record, AABB, vertex ring and arbiter) and written back with bulk stores.  Per substep the position update, the
// narrowphase of the touching pairs (GJK / EPA / manifold) and the 10 solver sweeps run entirely on the pool; what the rest of
// the library reads between steps (poses, world vertices, AABBs, motion bounds, moved flags) is written through to the env
// block as it is produced (fire-and-forget stores), velocities and arbiters are written back when an env leaves the pool.
//
// Everything rare keeps running on the env block with the routines of sim_step.cuh / sim_batch.cuh: neighbour-list rebuild,
// the gate scan over NOT touching pairs (only when the env's motion accumulator has reached the smallest gate), wake
// closure, cache expiry, recolouring, sleep.  Each of them is bracketed by write-back and re-gather of the pool ("repack").
// An env whose awake set outgrows the pool is parked: its substep is completed from global memory, it rejoins when there is
// room, and an env that does not fit into an empty pool runs the old per-env path.  Results never depend on the path.
#pragma once
#include "smem_stage.cuh"

#ifndef BP_RES_NS
#error "define BP_RES_NS to the sim_batch.cuh instantiation the resident step builds on"
#endif
#ifndef BP_RES_PA
#define BP_RES_PA 96           // pool: awake bodies per warp
#endif
#ifndef BP_RES_PV
#define BP_RES_PV 1344         // pool: world vertices per warp
#endif
#ifndef BP_RES_PS
#define BP_RES_PS 88           // pool: touching arbiters per warp
#endif

#if defined(BP_RES_DEBUG) && defined(BP_HOST_EMU)
#include <cstdio>
#endif

namespace bp {
namespace res {
using namespace BP_RES_NS;

constexpr int RK = BK;
constexpr int R_PA = BP_RES_PA, R_PV = BP_RES_PV, R_PS = BP_RES_PS;
static_assert(R_PS <= BWL, "merged colour list too short");
enum ResState { RS_NONE = 0, RS_WAIT = 1, RS_RES = 2, RS_PARK = 3, RS_DONE = 4 };

struct alignas(16) RBody { real px, py, pa, vx, vy, vw, bx, by, bw, minv, iinv, rc, rs; double D; };   // = the head of a BodyRec (112 B in fp64, 64 B in fp32)
static_assert(sizeof(RBody) % 16 == 0 && sizeof(RBody) <= sizeof(BodyRec) && offsetof(RBody, D) == offsetof(BodyRec, D), "RBody must mirror the head of BodyRec");
struct alignas(16) RBodyX {
    real liml, limr, limb, limt;        // the body leaves its neighbour-list margin when its AABB crosses these
    BBox bb;                            // current AABB incl. radius
    real rbound, rad, cogx, cogy;
    unsigned short gid, voff;           // body id in the env block; first vertex in rv
    unsigned char nv, env, moved, pad0;
    real pad1;
};
static_assert(sizeof(RBodyX) % 16 == 0 && offsetof(RBodyX, bb) % 16 == 0, "RBodyX layout (bulk copies land on 16-byte boundaries)");
struct alignas(16) RSlot {
    SlotRec s;                          // bulk copy of the env block's arbiter (a, b stay body ids)
    unsigned short ka, kb, ai, env;     // pool indices of the two bodies, index into the env's awake-pair list, env of the batch
    unsigned gjk; int sid;              // GJK warm-start id of the pair; arbiter slot in the env block
};
static_assert(sizeof(RSlot) % 16 == 0 && sizeof(SlotRec) % 16 == 0, "RSlot layout");

struct ResShared {
    BatchShared B;                      // env blocks, layouts, narrowphase queues, merged colour list (wl) of sim_batch.cuh
    RBody rb[R_PA];
    RBodyX rx[R_PA];
    D2 rv[R_PV];
    RSlot rs[R_PS];
    unsigned long long mbar; unsigned mbar_phase; int pad0;
    char* base[RK];                     // env block of every claimed env that steps (B.base is the subset a scan works on)
    int st[RK], kdone[RK], a0[RK], na[RK], s0[RK], ns[RK], nvv[RK], robk[RK], robg[RK], stamp[RK], ncach[RK], ncol[RK];
    int cs[RK][BCOL + 1];               // colour ranges of the env's arbiters (relative to s0)
    double G[RK], gate_min[RK];
    unsigned char f_rebuild[RK], f_dirty[RK], f_scan[RK], f_need[RK];
    int a_tot, s_tot, v_tot, n_col, fin;
    int off_col[BCOL + 1];
};

BP_DEV BodyRec& as_body(RBody& r) { return *reinterpret_cast<BodyRec*>(&r); }
BP_DEV Poly res_poly(ResShared& S, const RBodyX& X) { Poly P; P.v = (const real*)(S.rv + X.voff); P.n = X.nv; return P; }
BP_DEV double bp_wmin_d(double v) { return -bp_wmax(-v); }
BP_DEV unsigned res_ring_bytes(int nv) { return (unsigned)((nv * sizeof(D2) + 15) / 16 * 16); }      // bulk copies move multiples of 16 bytes
static_assert(sizeof(BBox) % 16 == 0 && (VMAX * sizeof(D2)) % 16 == 0, "bulk-copy granularity");

#if defined(BP_RES_DEBUG) && defined(BP_HOST_EMU)
static void res_check(const DevConst* C, ResShared& S, int e, const char* tag) {
    if (!S.base[e]) return;
    EnvView ev = view_s(S.base[e]);
    EnvScal* sc = ev.sc;
    PairRec* pr = ev.pair[sc->cur];
    for (int p = 0; p < sc->n_pairs; p++) {
        int lv = pr[p].live;
        if (lv >= 0 && ev.slot[lv].pair != p) { fprintf(stderr, "CHECK[%s] e%d k%d: pair %d live=%d but slot.pair=%d\n", tag, e, S.kdone[e], p, lv, ev.slot[lv].pair); return; }
    }
    for (int s2 = 0; s2 < sc->slot_hw; s2++) {
        int p = ev.slot[s2].pair;
        if (p >= 0 && (p >= sc->n_pairs || pr[p].live != s2)) { fprintf(stderr, "CHECK[%s] e%d k%d: slot %d pair=%d but live=%d\n", tag, e, S.kdone[e], s2, p, p < sc->n_pairs ? pr[p].live : -99); return; }
    }
}
#define BP_RCHECK(e, tag) res_check(C, S, e, tag)
#else
#define BP_RCHECK(e, tag)
#endif

// ------------------------------------------------------------------ write-back / gather
// Write the body records and arbiters of resident env e back to its env block.
BP_DEVNI void res_flush_env(const DevConst* C, ResShared& S, int e) {
    const int lane = bp_lane();
    EnvView ev = view_s(S.base[e]);
    PairRec* pr = ev.pair[ev.sc->cur];
    bp_fence_async();
    bp_wsync();
    const int a0 = S.a0[e], na = S.na[e], s0 = S.s0[e], ns = S.ns[e];
    for (int k = lane; k < na; k += BP_NL) bp_bulk_s2g(&ev.body[S.rx[a0 + k].gid], &S.rb[a0 + k], (unsigned)sizeof(RBody));
    for (int j = lane; j < ns; j += BP_NL) {
        RSlot& T = S.rs[s0 + j];
        bp_bulk_s2g(&ev.slot[T.sid], &T.s, (unsigned)sizeof(SlotRec));
        pr[T.s.pair].gjk = T.gjk;
    }
    bp_bulk_commit_wait();
    bp_wsync();
    BP_RCHECK(e, "flush");
}
BP_DEVNI void res_flush_all(const DevConst* C, ResShared& S) {
    for (int e = 0; e < RK; e++)
        if (S.st[e] == RS_RES) res_flush_env(C, S, e);
}

// Gather env e into the pool at (a_cur, v_cur, s_cur).  Returns false (warp-uniform) when it does not fit.
BP_DEVNI bool res_gather_env(const DevConst* C, ResShared& S, int e, int a_cur, int v_cur, int s_cur) {
    const int lane = bp_lane();
    const SimParams& P = C->P;
    EnvView ev = view_s(S.base[e]);
    EnvScal* sc = ev.sc;
    const LayoutView& L = S.B.L[e];
    const int na = sc->n_awake, ns = sc->n_order, ncol = sc->n_colours;
    if (a_cur + na > R_PA || s_cur + ns > R_PS || ncol > BCOL || sc->n_apairs > 65535) return false;
    // vertex ring offsets
    int vacc = 0;
    for (int bs = 0; bs < na; bs += BP_NL) {
        int k = bs + lane;
        int nv = (k < na) ? L.nv[ev.awake[k]] : 0;
        int nva = (int)((nv * sizeof(D2) + 15) / 16 * 16 / sizeof(D2));      // vertex rings start and end on 16-byte boundaries
        int tot; int ex = bp_wscan_excl(nva, &tot);
        if (k < na) { RBodyX& X = S.rx[a_cur + k]; X.voff = (unsigned short)(v_cur + vacc + ex); X.nv = (unsigned char)nv; }
        vacc += tot;
    }
    if (v_cur + vacc > R_PV) return false;
    bp_fence_async();
    bp_wsync();
    // bulk copies: body record, AABB, vertex ring per awake body; the arbiter per touching pair
    unsigned mybytes = 0;
    for (int k = lane; k < na; k += BP_NL) mybytes += (unsigned)(sizeof(RBody) + sizeof(BBox)) + res_ring_bytes(S.rx[a_cur + k].nv);
    for (int j = lane; j < ns; j += BP_NL) mybytes += (unsigned)sizeof(SlotRec);
    bp_mbar_arrive_tx(&S.mbar, mybytes);
    for (int k = lane; k < na; k += BP_NL) {
        int b = ev.awake[k];
        RBodyX& X = S.rx[a_cur + k];
        bp_bulk_g2s(&S.rb[a_cur + k], &ev.body[b], (unsigned)sizeof(RBody), &S.mbar);
        bp_bulk_g2s(&X.bb, &ev.bb[b], (unsigned)sizeof(BBox), &S.mbar);
        bp_bulk_g2s(&S.rv[X.voff], ev.wv + (size_t)b * VMAX * 2, res_ring_bytes(X.nv), &S.mbar);
        real mg = body_margin(P, sc, b);
        BBox o = ev.b0[b];
        X.liml = o.l - mg; X.limr = o.r + mg; X.limb = o.b - mg; X.limt = o.t + mg;
        X.rbound = L.rbound[b]; X.rad = L.rad[b]; X.cogx = L.cog[2 * b]; X.cogy = L.cog[2 * b + 1];
        X.gid = (unsigned short)b; X.env = (unsigned char)e; X.moved = ev.moved[b]; X.pad0 = 0;
    }
    for (int j = lane; j < ns; j += BP_NL) {
        int sid = ev.order[j];
        RSlot& T = S.rs[s_cur + j];
        bp_bulk_g2s(&T.s, &ev.slot[sid], (unsigned)sizeof(SlotRec), &S.mbar);
        T.sid = sid; T.env = (unsigned short)e; T.ai = 0xFFFF;
    }
    for (int c = lane; c <= ncol; c += BP_NL) S.cs[e][c] = ev.col_start[c];
    bp_mbar_wait(&S.mbar, S.mbar_phase);
    bp_wsync();
    if (lane == 0) {
        S.mbar_phase ^= 1u;
        S.a0[e] = a_cur; S.na[e] = na; S.s0[e] = s_cur; S.ns[e] = ns; S.nvv[e] = vacc; S.ncol[e] = ncol;
        S.G[e] = sc->G; S.stamp[e] = sc->stamp; S.ncach[e] = sc->n_cached; S.robg[e] = sc->robot;
        S.f_rebuild[e] = 0; S.f_dirty[e] = 0;
    }
    // pool indices of the arbiters' bodies, warm-start ids
    const PairRec* pr = ev.pair[sc->cur];
    for (int j = lane; j < ns; j += BP_NL) {
        RSlot& T = S.rs[s_cur + j];
        int a = T.s.a, b = T.s.b, ka = 0, kb = 0;
        for (int k = 0; k < na; k++) {
            int g = S.rx[a_cur + k].gid;
            if (g == a) ka = a_cur + k;
            if (g == b) kb = a_cur + k;
        }
        T.ka = (unsigned short)ka; T.kb = (unsigned short)kb;
        T.gjk = pr[T.s.pair].gjk;
    }
    for (int k = lane; k < na; k += BP_NL)
        if (S.rx[a_cur + k].gid == sc->robot) S.robk[e] = a_cur + k;
    bp_wsync();
    // awake-pair list: touching pairs are evaluated from the pool (gate -2), the smallest gate of the others decides when
    // the env block's list has to be scanned again
    double gmin = INFINITY;
    const int nap = sc->n_apairs;
    for (int i = lane; i < nap; i += BP_NL) {
        int p = ev.apairs[i];
        int hit = -1;
        for (int j = 0; j < ns; j++)
            if (S.rs[s_cur + j].s.pair == p) hit = j;
        if (hit >= 0) { S.rs[s_cur + hit].ai = (unsigned short)i; ev.ap_gate[i] = -RC(2.0); }
        else {
            double g = ev.ap_gate[i];
            if (g == -RC(2.0)) { g = -RC(1.0); ev.ap_gate[i] = -RC(1.0); }      // was resident, is not any more: always evaluate
            gmin = fmin(gmin, g);
        }
    }
    gmin = bp_wmin_d(gmin);
    if (lane == 0) S.gate_min[e] = gmin;
    bp_wsync();
    return true;
}

// Smallest gate of the NOT resident awake pairs of env e (after a scan of the env block's list).
BP_DEVNI void res_gate_min(const DevConst* C, ResShared& S, int e) {
    const int lane = bp_lane();
    EnvView ev = view_s(S.base[e]);
    const int nap = ev.sc->n_apairs;
    double gmin = INFINITY;
    for (int i = lane; i < nap; i += BP_NL) {
        double g = ev.ap_gate[i];
        if (g != -RC(2.0)) gmin = fmin(gmin, g);
    }
    gmin = bp_wmin_d(gmin);
    if (lane == 0) S.gate_min[e] = gmin;
    bp_wsync();
}

// Merged colour lists over the resident envs: S.B.wl[off_col[c] ..) = pool indices of the arbiters of colour c.
BP_DEVNI void res_build_worklist(ResShared& S) {
    const int lane = bp_lane();
    if (lane == 0) {
        int acc = 0, ncol = 0;
        for (int e = 0; e < RK; e++)
            if (S.st[e] == RS_RES && S.ncol[e] > ncol) ncol = S.ncol[e];
        for (int c = 0; c < ncol; c++) {
            S.off_col[c] = acc;
            for (int e = 0; e < RK; e++) {
                S.B.cpre[e][c] = acc;
                if (S.st[e] == RS_RES && c < S.ncol[e]) acc += S.cs[e][c + 1] - S.cs[e][c];
            }
        }
        S.off_col[ncol] = acc;
        S.n_col = ncol;
    }
    bp_wsync();
    const int tot = S.s_tot;
    for (int q = lane; q < tot; q += BP_NL) {
        int e = S.rs[q].env;
        if (S.st[e] != RS_RES) continue;
        int pos = q - S.s0[e], nc = S.ncol[e], c = 0;
        while (c + 1 < nc && S.cs[e][c + 1] <= pos) c++;
        S.B.wl[S.B.cpre[e][c] + (pos - S.cs[e][c])] = (unsigned)q;
    }
    bp_wsync();
}

// ------------------------------------------------------------------ per-substep bookkeeping of one env
// ship_ice_env.py:284-290 on the pool (one lane): bit0 = yaw limit violated (yaw rate zeroed), bit1 = outside the channel.
BP_DEV int res_shipice_checks(ResShared& S, const SimParams& P, int e) {
    int f = 0;
    RBody& R = S.rb[S.robk[e]];
    const RBodyX& X = S.rx[S.robk[e]];
    real a = R.pa;
    if (a <= P.yaw_lo || a >= P.yaw_hi) { R.vw = RC(0.0); f |= 1; }
    real ox = R.px - (X.cogx * R.rc - X.cogy * R.rs);
    if (ox < 0 || ox > P.map_w) f |= 2;
    return f;
}

BP_DEVNI void res_finish_env(const DevConst* C, ResShared& S, int e, const StepOut& out) {
    EnvView ev = view_s(S.base[e]);
    int env = S.B.env_id[e];
    env_step_finish(ev, S.B.L[e], C->P, C->d, env, out, (int)S.B.flags[e]);
    if (bp_lane() == 0) { ev.sc->done = out.terminated[env]; ev.sc->obs_raw = 0; S.st[e] = RS_DONE; S.fin = 1; }
    bp_wsync();
}

// Env e just completed a substep from global memory (parked / evicted): substep checks, counter, finish.
BP_DEVNI void res_after_substep_global(const DevConst* C, ResShared& S, int e, const StepOut& out) {
    EnvView ev = view_s(S.base[e]);
    if (bp_lane() == 0) {
        S.B.flags[e] |= (unsigned char)env_substep_checks(ev, S.B.L[e], C->P, ev.sc->robot);
        S.kdone[e]++;
    }
    bp_wsync();
    if (S.kdone[e] >= C->P.substeps) res_finish_env(C, S, e, out);
}

// Env e runs its remaining substeps with the one-env path of sim_step.cuh (it does not fit into an empty pool).
BP_DEVNI void res_run_global(const DevConst* C, ResShared& S, int e, const StepOut& out) {
    EnvView ev = view_s(S.base[e]);
    const SimParams& P = C->P;
    while (S.st[e] != RS_DONE) {
        substep(ev, S.B.L[e], P, C->d, P.dt, P.dt);
        res_after_substep_global(C, S, e, out);
    }
}

// Re-pack the pool from the env blocks.  Precondition: every resident env is written back and its rare routines are done.
// phase 0: substep boundary (parked / waiting envs may join); 1: after the position update; 2: before the solve.
BP_DEVNI void res_repack(const DevConst* C, ResShared& S, const StepOut& out, int phase) {
    const int lane = bp_lane();
    const SimParams& P = C->P;
    int a_cur = 0, v_cur = 0, s_cur = 0;
    for (int e = 0; e < RK; e++) {
        const int st = S.st[e];
        bool cand = (st == RS_RES) || (phase == 0 && (st == RS_PARK || st == RS_WAIT));
        if (!cand) continue;
        {
            // the solver order has to describe the current set of touching arbiters (sleep_pass and the neighbour-list
            // rebuild free arbiters and only mark the order dirty)
            EnvView ev0 = view_s(S.base[e]);
            if (ev0.sc->dirty) build_order(ev0, S.B.L[e], C->d);
        }
        bool ok = res_gather_env(C, S, e, a_cur, v_cur, s_cur);
        if (ok) {
            a_cur += S.na[e]; v_cur += S.nvv[e]; s_cur += S.ns[e];
            if (lane == 0) S.st[e] = RS_RES;
            bp_wsync();
            continue;
        }
        EnvView ev = view_s(S.base[e]);
        if (st == RS_RES && phase != 0) {
            // evicted in the middle of a substep: complete it from the env block
            if (phase == 1) substep_rest(ev, S.B.L[e], P, C->d, P.dt, P.dt);
            else solve_env(ev, S.B.L[e], P, P.dt, P.dt);
            res_after_substep_global(C, S, e, out);
        }
        if (S.st[e] != RS_DONE) {
            if (phase == 0 && a_cur == 0 && s_cur == 0) res_run_global(C, S, e, out);
            else { if (lane == 0) S.st[e] = RS_PARK; bp_wsync(); }
        }
    }
    if (lane == 0) { S.a_tot = a_cur; S.v_tot = v_cur; S.s_tot = s_cur; }
    bp_wsync();
    res_build_worklist(S);
}

// ------------------------------------------------------------------ narrowphase of the touching pairs, on the pool
// The pair of pool arbiter q is found apart (pair_apart_impl of sim_step.cuh): the arbiter stays cached for `persistence`
// steps, the env block's pair entry gets its separation bound and gate.
BP_DEVNI void res_pair_apart(const DevConst* C, ResShared& S, int q, real gap) {
    RSlot& T = S.rs[q];
    const int e = T.env;
    EnvView ev = view_s(S.base[e]);
    EnvScal* sc = ev.sc;
    PairRec* pr = ev.pair[sc->cur];
    const int p = T.s.pair;
    pr[p].flags |= 1;
    double g = gap * RC(0.999999) - RC(1e-9);
    pr[p].skip = S.rb[T.ka].D + S.rb[T.kb].D + g;
    ev.ap_gate[T.ai] = S.G[e] + g;
    if (T.s.touch) {
        T.s.touch = 0;
        T.s.stamp = S.stamp[e] - 1;
        bp_atomic_and(&ev.livemask[p >> 5], ~(1u << (p & 31)));
        bp_atomic_add(&sc->n_cached, 1);
        bp_atomic_add(&S.ncach[e], 1);
        sc->dirty = 1;
        S.f_dirty[e] = 1;
    }
}

// Manifold + cpArbiterUpdate for pool arbiter q from the closest-feature result (pair_finish_lane of sim_batch.cuh for an
// arbiter that is already touching).
BP_DEVNI void res_pair_finish(const DevConst* C, ResShared& S, int q, const Closest& pts) {
    RSlot& T = S.rs[q];
    const int e = T.env;
    const LayoutView& L = S.B.L[e];
    const RBody& A = S.rb[T.ka]; const RBody& B = S.rb[T.kb];
    const RBodyX& XA = S.rx[T.ka]; const RBodyX& XB = S.rx[T.kb];
    Manifold m;
    make_manifold_core(res_poly(S, XA), res_poly(S, XB), XA.rad, XB.rad, L.ln + (size_t)XA.gid * VMAX * 2, L.ln + (size_t)XB.gid * VMAX * 2,
                       A.rc, A.rs, B.rc, B.rs, pts, &m);
    T.gjk = pts.id;
    if (m.count > 0) {
        SlotRec& R = T.s;
        real jn_new[2] = {0, 0}, jt_new[2] = {0, 0};
        int oc = R.cnt;
        for (int x = 0; x < m.count; x++)
            for (int j = 0; j < oc; j++)
                if (m.hash[x] == R.c[j].hash) { jn_new[x] = R.c[j].jn; jt_new[x] = R.c[j].jt; }
        real ax = A.px, ay = A.py, bx = B.px, by = B.py;
        for (int x = 0; x < m.count; x++) {
            ContactRec& c = R.c[x];
            c.r1x = m.p1[x].x - ax; c.r1y = m.p1[x].y - ay;
            c.r2x = m.p2[x].x - bx; c.r2y = m.p2[x].y - by;
            c.hash = m.hash[x]; c.jn = jn_new[x]; c.jt = jt_new[x];
        }
        R.cnt = m.count;
        R.nx = m.n.x; R.ny = m.n.y;
        R.stamp = S.stamp[e];
    } else {
        res_pair_apart(C, S, q, pts.d - XA.rad - XB.rad);
    }
}

BP_DEV void res_flush_epa(const DevConst* C, ResShared& S, int n2) {
    const int lane = bp_lane();
    for (int bs = 0; bs < n2; bs += BP_NL) {
        int t = bs + lane;
        if (t < n2) {
            int q = (int)S.B.q2[t];
            RSlot& T = S.rs[q];
            int capped = 0;
            Closest pts = gjk_epa(res_poly(S, S.rx[T.ka]), res_poly(S, S.rx[T.kb]), S.B.q2hint[t], &capped);
            if (capped) bp_atomic_or((unsigned*)&((EnvScal*)S.base[T.env])->status, ST_GJK_ITER);
            res_pair_finish(C, S, q, pts);
        }
    }
    bp_wsync();
}

BP_DEVNI void res_collide_touching(const DevConst* C, ResShared& S) {
    const int lane = bp_lane();
    const int tot = S.s_tot;
    // need test (a touching pair is re-evaluated when one of its bodies moved in this substep) + exact AABB filter
    int n1 = 0;
    for (int bs = 0; bs < tot; bs += BP_NL) {
        int q = bs + lane;
        bool take = false;
        if (q < tot) {
            RSlot& T = S.rs[q];
            if (S.st[T.env] == RS_RES && T.s.touch) {
                const RBodyX& XA = S.rx[T.ka]; const RBodyX& XB = S.rx[T.kb];
                if (XA.moved || XB.moved) {
                    BBox A = XA.bb, B = XB.bb;
                    bool bbhit = (A.l <= B.r && B.l <= A.r && A.b <= B.t && B.b <= A.t);
                    if (bbhit) take = true;
                    else res_pair_apart(C, S, q, fmax(fmax(A.l - B.r, B.l - A.r), fmax(A.b - B.t, B.b - A.t)));
                }
            }
        }
        int t; int rk = bp_wrank(take, &t);
        if (take) S.B.q1[n1 + rk] = (unsigned)q;
        n1 += t;
    }
    bp_wsync();
    // GJK; pairs with overlapping cores go on to the EPA queue
    int n2 = 0;
    for (int bs = 0; bs < n1; bs += BP_NL) {
        int t = bs + lane;
        bool to_epa = false; unsigned hint = 0; int q = 0;
        if (t < n1) {
            q = (int)S.B.q1[t];
            RSlot& T = S.rs[q];
            const RBodyX& XA = S.rx[T.ka]; const RBodyX& XB = S.rx[T.kb];
            int capped = 0, iters = 0;
            Closest pts; pts.d = 0; pts.id = 0; pts.n = mk(0, 0);
            to_epa = gjk_only(res_poly(S, XA), res_poly(S, XB), bb_centre(XA.bb), bb_centre(XB.bb), T.gjk, &pts, &hint, &iters, &capped);
            if (capped) bp_atomic_or((unsigned*)&((EnvScal*)S.base[T.env])->status, ST_GJK_ITER);
            if (!to_epa) res_pair_finish(C, S, q, pts);
        }
        int tt; int rk = bp_wrank(to_epa, &tt);
        if (to_epa) { S.B.q2[n2 + rk] = (unsigned)q; S.B.q2hint[n2 + rk] = hint; }
        n2 += tt;
        if (n2 > BQE - BP_NL) { bp_wsync(); res_flush_epa(C, S, n2); n2 = 0; }
    }
    bp_wsync();
    if (n2 > 0) res_flush_epa(C, S, n2);
}

// ------------------------------------------------------------------ one substep (cpSpaceStep) for the resident envs
BP_DEVNI void res_substep(const DevConst* C, ResShared& S, const StepOut& out) {
    const int lane = bp_lane();
    const SimParams& P = C->P;
    const Dims& d = C->d;
    const real dt = P.dt;
    // ---- 0. stamps
    for (int e = lane; e < RK; e += BP_NL)
        if (S.st[e] == RS_RES) { EnvScal* sc = (EnvScal*)S.base[e]; S.stamp[e]++; sc->stamp = S.stamp[e]; S.B.dmax_bits[e] = 0ull; }
    bp_wsync();
    // ---- 1. position update of the awake bodies; poses, vertices, AABBs are written through to the env block
    {
        const int tot = S.a_tot;
        for (int i = lane; i < tot; i += BP_NL) {
            RBodyX& X = S.rx[i];
            const int e = X.env;
            if (S.st[e] != RS_RES) continue;
            RBody& R = S.rb[i];
            EnvView ev = view_s(S.base[e]);
            const int b = X.gid;
            real vx = R.vx, vy = R.vy, vw = R.vw, bx = R.bx, by = R.by, bw = R.bw;
            bool mv = (vx != RC(0.0) || vy != RC(0.0) || vw != RC(0.0) || bx != RC(0.0) || by != RC(0.0) || bw != RC(0.0));
            if (mv) {
                real dx = (vx + bx) * dt, dy = (vy + by) * dt, da = (vw + bw) * dt;
                R.px = R.px + dx; R.py = R.py + dy; R.pa = R.pa + da;
                real dD = (fabs(dx) + fabs(dy) + X.rbound * fabs(da)) * RC(1.000001);
                R.D += dD;
                bp_atomic_max_u64(&S.B.dmax_bits[e], bp_d2bits(dD));
                // refresh_body_impl on the pool
                real s, c;
                bp_sincos(R.pa, &s, &c);
                R.rc = c; R.rs = s;
                real cx = X.cogx, cy = X.cogy;
                real tx = R.px - (cx * c - cy * s);
                real ty = R.py - (cx * s + cy * c);
                real ms = -s;
                const D2* lv2 = (const D2*)(S.B.L[e].lv + (size_t)b * VMAX * 2);
                D2* wv2 = (D2*)(ev.wv + (size_t)b * VMAX * 2);
                D2* sv2 = S.rv + X.voff;
                const int n = X.nv;
                real l = INFINITY, r = -INFINITY, bt = INFINITY, tp = -INFINITY;
#pragma unroll 2
                for (int v = 0; v < n; v++) {
                    D2 qv = lv2[v];
                    real wx = c * qv.x + ms * qv.y + tx;
                    real wy = s * qv.x + c * qv.y + ty;
                    D2 w; w.x = wx; w.y = wy;
                    sv2[v] = w; wv2[v] = w;
                    l = fmin(l, wx); r = fmax(r, wx); bt = fmin(bt, wy); tp = fmax(tp, wy);
                }
                real rad = X.rad;
                BBox box; box.l = l - rad; box.b = bt - rad; box.r = r + rad; box.t = tp + rad;
                X.bb = box; ev.bb[b] = box;
                if (box.l < X.liml || box.r > X.limr || box.b < X.limb || box.t > X.limt) { S.f_rebuild[e] = 1; ev.sc->rebuild = 1; }
                ev.moved_step[b] = 1; ev.mvrb[b] = 1;
                R.bx = RC(0.0); R.by = RC(0.0); R.bw = RC(0.0);
                BodyRec& GB = ev.body[b];
                GB.px = R.px; GB.py = R.py; GB.pa = R.pa; GB.rc = c; GB.rs = s; GB.D = R.D;
            }
            X.moved = mv ? 1 : 0;
            ev.moved[b] = mv ? 1 : 0;
        }
    }
    bp_wsync();
    for (int e = lane; e < RK; e += BP_NL)
        if (S.st[e] == RS_RES) { S.G[e] += RC(2.0) * bp_bits2d(S.B.dmax_bits[e]) * RC(1.000001); ((EnvScal*)S.base[e])->G = S.G[e]; }
    bp_wsync();
    // ---- rare: neighbour-list rebuild (env block), then re-pack
    {
        bool rb_any = false;
        for (int e = 0; e < RK; e++) rb_any = rb_any || (S.st[e] == RS_RES && S.f_rebuild[e]);
        if (rb_any) {
            res_flush_all(C, S);
            for (int e = 0; e < RK; e++)
                if (S.st[e] == RS_RES && S.f_rebuild[e]) {
                    EnvView ev = view_s(S.base[e]);
                    BP_RCHECK(e, "pre-rebuild");
                    rebuild_pairs_inc(ev, P, d);
                    build_apairs(ev);
                    BP_RCHECK(e, "rebuild");
                }
            res_repack(C, S, out, 1);
        }
    }
    // ---- 2a. collide: the touching pairs, on the pool
    res_collide_touching(C, S);
    // ---- 2b. collide: the other awake pairs, from the env block, only for envs whose accumulator reached their smallest gate
    {
        bool scan_any = false;
        for (int e = 0; e < RK; e++) {
            bool f = (S.st[e] == RS_RES) && (S.G[e] >= S.gate_min[e]);
            if (lane == 0) { S.f_scan[e] = f ? 1 : 0; S.B.base[e] = f ? S.base[e] : nullptr; }
            scan_any = scan_any || f;
        }
        bp_wsync();
        if (scan_any) {
            if (lane == 0) S.B.skip_res = 1;
            bp_wsync();
            batch_collide(C, S.B);
            for (int e = 0; e < RK; e++)
                if (S.f_scan[e]) {
                    BP_RCHECK(e, "scan");
                    EnvScal* sc = (EnvScal*)S.base[e];
                    if (lane == 0) { S.ncach[e] = sc->n_cached; if (sc->dirty) S.f_dirty[e] = 1; }
                    bp_wsync();
                    if (!S.f_dirty[e] && !sc->any_wake) res_gate_min(C, S, e);
                }
        }
    }
    bp_wsync();
    // ---- rare: wake closure, cache expiry, recolouring (env block), then re-pack
    {
        bool need_any = false;
        for (int e = 0; e < RK; e++) {
            bool f = false;
            if (S.st[e] == RS_RES) {
                f = S.f_dirty[e] != 0;
                if (S.f_scan[e]) { const EnvScal* sc = (const EnvScal*)S.base[e]; f = f || sc->any_wake || sc->dirty; }
            }
            if (lane == 0) S.f_need[e] = f ? 1 : 0;
            need_any = need_any || f;
        }
        bp_wsync();
        if (need_any) res_flush_all(C, S);
        for (int e = 0; e < RK; e++) {
            if (S.st[e] != RS_RES) continue;
            if (!S.f_need[e] && S.ncach[e] <= 0) continue;
            EnvView ev = view_s(S.base[e]);
            EnvScal* sc = ev.sc;
            BP_RCHECK(e, "pre-rare");
            if (S.f_need[e] && sc->any_wake) wake_closure_scan(ev, S.B.L[e], d);
            BP_RCHECK(e, "wake");
            if (sc->n_cached > 0) expire_cached(ev, P, d);
            BP_RCHECK(e, "expire");
            if (S.f_need[e] && sc->dirty) build_order(ev, S.B.L[e], d);
            if (lane == 0) S.ncach[e] = sc->n_cached;
            bp_wsync();
        }
        if (need_any) res_repack(C, S, out, 2);
    }
    // ---- 3. prestep over all resident arbiters
    {
        const int tot = S.s_tot;
        for (int q = lane; q < tot; q += BP_NL) {
            RSlot& T = S.rs[q];
            if (S.st[T.env] != RS_RES) continue;
            arb_prestep_core(T.s, as_body(S.rb[T.ka]), as_body(S.rb[T.kb]), P);
        }
    }
    // ---- 4. velocity update of the awake dynamic bodies
    {
        const int tot = S.a_tot;
        for (int i = lane; i < tot; i += BP_NL) {
            if (S.st[S.rx[i].env] != RS_RES) continue;
            RBody& R = S.rb[i];
            if (R.minv != RC(0.0)) { R.vx = R.vx * P.damp + RC(0.0); R.vy = R.vy * P.damp + RC(0.0); R.vw = R.vw * P.damp + RC(0.0); }
        }
    }
    bp_wsync();
    // ---- 5. warm start + iterations, colour by colour over the merged lists
    const int nc = S.n_col;
    for (int c = 0; c < nc; c++) {
        int c0 = S.off_col[c], c1 = S.off_col[c + 1];
        for (int j = c0 + lane; j < c1; j += BP_NL) {
            RSlot& T = S.rs[S.B.wl[j]];
            if (S.st[T.env] != RS_RES) continue;
            arb_warm_core(T.s, as_body(S.rb[T.ka]), as_body(S.rb[T.kb]), RC(1.0));
        }
        bp_wsync();
    }
    for (int it = 0; it < P.iterations; it++) {
        for (int c = 0; c < nc; c++) {
            int c0 = S.off_col[c], c1 = S.off_col[c + 1];
            for (int j = c0 + lane; j < c1; j += BP_NL) {
                RSlot& T = S.rs[S.B.wl[j]];
                if (S.st[T.env] != RS_RES) continue;
                arb_solve_core(T.s, as_body(S.rb[T.ka]), as_body(S.rb[T.kb]));
            }
            bp_wsync();
        }
    }
    // ---- 6. arbiters solved this step are NORMAL next step
    {
        const int tot = S.s_tot;
        for (int q = lane; q < tot; q += BP_NL)
            if (S.st[S.rs[q].env] == RS_RES) S.rs[q].s.first = 0;
    }
    bp_wsync();
}

// ShipIceEnv.step for the envs [env0, env0 + RK) of a handle (those >= n_envs are skipped); same contract as
// shipice_step_batch of sim_batch.cuh (auto reset included).
BP_DEVNI void shipice_step_resident(const DevConst* C, ResShared& S, char* mem, size_t env_bytes, int env0, const float* actions,
                                    const unsigned char* mask, const StepOut& out, int auto_reset, const int* perm = nullptr) {
    const int lane = bp_lane();
    const SimParams& P = C->P;
    const Dims& d = C->d;
    for (int e = 0; e < RK; e++) {
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
                if (lane == 0) S.B.L[e] = L;
                S.base[e] = base; S.kdone[e] = -1;
                BP_RCHECK(e, "pre-begin");
                shipice_step_begin(ev, L, P, d, (real)actions[env]);
                BP_RCHECK(e, "begin");
            }
        }
        if (lane == 0) {
            S.base[e] = base; S.B.base[e] = nullptr; S.B.env_id[e] = env; S.B.flags[e] = 0;
            S.st[e] = base ? RS_WAIT : RS_NONE; S.kdone[e] = 0; S.na[e] = 0; S.ns[e] = 0; S.f_rebuild[e] = 0; S.f_dirty[e] = 0; S.f_scan[e] = 0;
        }
        bp_wsync();
    }
    if (lane == 0) { S.B.skip_res = 1; S.fin = 0; S.a_tot = 0; S.s_tot = 0; S.v_tot = 0; S.n_col = 0; }
    bp_wsync();
    res_repack(C, S, out, 0);
    int since = 0;
    for (;;) {
        bool any_res = false, any_wait = false;
        for (int e = 0; e < RK; e++) {
            any_res = any_res || S.st[e] == RS_RES;
            any_wait = any_wait || S.st[e] == RS_PARK || S.st[e] == RS_WAIT;
        }
        if (!any_res && !any_wait) break;
        if (any_wait && (!any_res || S.fin || since >= 32)) {
            res_flush_all(C, S);
            if (lane == 0) S.fin = 0;
            bp_wsync();
            res_repack(C, S, out, 0);
            since = 0;
            continue;
        }
        res_substep(C, S, out);
        since++;
        for (int e = lane; e < RK; e += BP_NL)
            if (S.st[e] == RS_RES) { S.B.flags[e] |= (unsigned char)res_shipice_checks(S, P, e); S.kdone[e]++; }
        bp_wsync();
        for (int e = 0; e < RK; e++)
            if (S.st[e] == RS_RES && S.kdone[e] >= P.substeps) { res_flush_env(C, S, e); res_finish_env(C, S, e, out); }
    }
    if (lane == 0) S.B.skip_res = 0;
    bp_wsync();
}

}  // namespace res
}  // namespace bp
