// Batched single-source shortest paths on 8-connected boolean grids: one thread block per map.
//
// Replaces spfa.spfa (reference deps/spfa/src/main.cpp:10-116).  The reference relaxes edges with a
// sequential SPFA queue in fp32 (weights 1 and sqrtf(2), :14,:74).  Because fp32 addition is
// monotone, every label-correcting schedule converges to the same least fixed point
//     d[v] = min_u fl32(d[u] + w(u,v)),  d[source] = 0,
// so a frontier-parallel relaxation with atomicMin on the (non-negative, hence order-preserving)
// float bit patterns yields distances BIT-IDENTICAL to the reference.  Parents are chosen
// deterministically afterwards (first neighbour, in the reference's direction order, that attains
// the minimum); the reference's own tie-break depends on its queue order.
// Quirks kept: moves between two free cells are allowed regardless of the orthogonal neighbours
// (no corner-cutting check, :50-61); unreachable cells report 0 (:101); "infinity" = 2*H*W (:28);
// a blocked source reaches nothing.
#pragma once
#include <math.h>
#include <stdint.h>
#include "raster.cuh"   // BP_TID / BP_NT / BP_BSYNC

namespace bp {

#if defined(__CUDACC__) && !defined(BP_HOST_EMU)
BP_DEV int spfa_atomic_min(int* p, int v) { return atomicMin(p, v); }
BP_DEV int spfa_atomic_max(int* p, int v) { return atomicMax(p, v); }
BP_DEV int spfa_atomic_exch(int* p, int v) { return atomicExch(p, v); }
BP_DEV int spfa_atomic_add(int* p, int v) { return atomicAdd(p, v); }
BP_DEV int f2i(float f) { return __float_as_int(f); }
BP_DEV float i2f(int i) { return __int_as_float(i); }
#else
BP_DEV int spfa_atomic_min(int* p, int v) { int o = *p; if (v < o) *p = v; return o; }
BP_DEV int spfa_atomic_max(int* p, int v) { int o = *p; if (v > o) *p = v; return o; }
BP_DEV int spfa_atomic_exch(int* p, int v) { int o = *p; *p = v; return o; }
BP_DEV int spfa_atomic_add(int* p, int v) { int o = *p; *p += v; return o; }
BP_DEV int f2i(float f) { int i; memcpy(&i, &f, 4); return i; }
BP_DEV float i2f(int i) { float f; memcpy(&f, &i, 4); return f; }
#endif

struct SpfaShared { int n_next; };

// Relaxation only: on return dist_out holds the fp32 bit patterns of the least fixed point, unreachable cells = 2*H*W.
// scratch: q0[HW], q1[HW], mark[HW] (ints).  dist_out doubles as the working distance array.
BP_DEV void spfa_relax(SpfaShared& S, const uint8_t* map, int H, int W, int si, int sj, float* dist_out, int* q0, int* q1, int* mark) {
    const int tid = BP_TID, nt = BP_NT;
    const int nv = H * W;
    const float inf = 2.0f * (float)nv;
    const float s2 = sqrtf(2.0f);
    const int di[8] = {-1, -1, -1, 0, 1, 1, 1, 0};
    const int dj[8] = {-1, 0, 1, 1, 1, 0, -1, -1};
    int* dist = (int*)dist_out;
    const int inf_i = f2i(inf);
    for (int v = tid; v < nv; v += nt) { dist[v] = inf_i; mark[v] = 0; }
    BP_BSYNC();
    const int s = si * W + sj;
    if (tid == 0) { dist[s] = f2i(0.0f); q0[0] = s; S.n_next = 0; }
    BP_BSYNC();
    int n_cur = 1, iter = 1;
    while (n_cur > 0) {
        for (int idx = tid; idx < n_cur; idx += nt) {
            int u = q0[idx];
            if (!map[u]) continue;
            float du = i2f(*(volatile int*)&dist[u]);
            int ui = u / W, uj = u - ui * W;
#pragma unroll
            for (int k = 0; k < 8; k++) {
                int ip = ui + di[k], jp = uj + dj[k];
                if (ip < 0 || jp < 0 || ip >= H || jp >= W) continue;
                int v = ip * W + jp;
                if (!map[v]) continue;
                float nd = du + ((k & 1) ? 1.0f : s2);
                int ndi = f2i(nd);
                int old = spfa_atomic_min(&dist[v], ndi);
                if (ndi < old) {
                    if (spfa_atomic_exch(&mark[v], iter) != iter) { int pos = spfa_atomic_add(&S.n_next, 1); q1[pos] = v; }
                }
            }
        }
        BP_BSYNC();
        n_cur = S.n_next;
        BP_BSYNC();
        if (tid == 0) S.n_next = 0;
        int* t = q0; q0 = q1; q1 = t;
        iter++;
        BP_BSYNC();
    }
}

// The parent the kernels assign to cell v: the first neighbour u (reference direction order, u + dir[k] = v) with
// fl32(d[u] + w_k) == d[v]; -1 for the source, blocked and unreachable cells.  `dist` = raw output of spfa_relax.
BP_DEV int spfa_parent_of(const uint8_t* map, int H, int W, const int* dist, int s, int v) {
    const float inf = 2.0f * (float)(H * W);
    const float s2 = sqrtf(2.0f);
    const int di[8] = {-1, -1, -1, 0, 1, 1, 1, 0};
    const int dj[8] = {-1, 0, 1, 1, 1, 0, -1, -1};
    float d = i2f(dist[v]);
    if (v == s || !(d < inf) || !map[v]) return -1;
    int vi = v / W, vj = v - vi * W;
    for (int k = 0; k < 8; k++) {
        int ip = vi - di[k], jp = vj - dj[k];
        if (ip < 0 || jp < 0 || ip >= H || jp >= W) continue;
        int u = ip * W + jp;
        if (!map[u]) continue;
        float du = i2f(dist[u]);
        if (du < inf && du + ((k & 1) ? 1.0f : s2) == d) return u;
    }
    return -1;
}

// Reference-ORDER parents: the queue discipline of the reference solver restated literally (deps/spfa/src/main.cpp:62-85 — a
// FIFO label-correcting queue; a vertex pushed with a smaller label than the queue's front is swapped to the front; a parent
// is recorded at every strict improvement), run by ONE thread.  The parent of a vertex with several equal-cost predecessors
// is whichever improved it last in that processing order, which no order-free relaxation can reproduce; the env kernels use
// this only in their exact-parents mode (box distances / robot paths that must follow the reference's tie-breaks waypoint by
// waypoint).  dist[nv] receives the fp32 labels (2*H*W where unreachable), parents[nv] the predecessors (-1 = none).
// ring: at least nv + 2 ints (at most nv vertices are queued at a time, so the reference's 8*nv linear queue folds into a
// ring without changing which entries meet); inq: nv bytes.
// Split in two so that the O(nv) initialisation can be shared by the threads of a block (tid of nt) while the queue itself is
// run by one thread.
// Blocked cells get the label -1: no tentative label (>= 0) ever improves it, so the queue loop needs no look-up of the map.
BP_DEV void spfa_serial_init(const uint8_t* map, int nv, float* dist, int* parents, unsigned char* inq, int tid, int nt) {
    const float inf = 2.0f * (float)nv;
    for (int v = tid; v < nv; v += nt) { dist[v] = map[v] ? inf : -1.0f; parents[v] = -1; inq[v] = 0; }
}
// BORDER_BLOCKED: the caller guarantees that every cell of the first / last row and column is blocked (true for the padded room
// maps of the envs), so a queued cell has all eight neighbours inside the grid and the scan works on linear indices only.
template <bool BORDER_BLOCKED = false>
BP_DEV void spfa_serial_run(const uint8_t* map, int H, int W, int si, int sj, float* dist, int* parents, int* ring, unsigned char* inq) {
    const int nv = H * W;
    const int cap = nv + 2;
    const float s2 = sqrtf(2.0f);
    const int di[8] = {-1, -1, -1, 0, 1, 1, 1, 0};
    const int dj[8] = {-1, 0, 1, 1, 1, 0, -1, -1};
    const int s = si * W + sj;
    // the reference's head / tail only grow (queue[++tail] = v, u = queue[++head]); here they are slots of the ring and `live`
    // counts the entries between them (head < tail  <=>  live > 0)
    int head = 0, tail = 1, live = 1;
    if (!map[s]) return;                          // a blocked source has no edges
    dist[s] = 0.0f;
    ring[tail] = s;
    inq[s] = 1;
    while (live > 0) {
        if (++head == cap) head = 0;
        live--;
        const int u = ring[head];
        inq[u] = 0;
        const float du = dist[u];
        // neighbours in the reference's edge order; the labels are fetched up front (distinct cells: an update of one does not
        // change another's), the updates then happen in order
        int nbr[8]; float dv[8];
        if (BORDER_BLOCKED) {
#pragma unroll
            for (int k = 0; k < 8; k++) { nbr[k] = u + di[k] * W + dj[k]; dv[k] = dist[nbr[k]]; }
        } else {
            const int ui = u / W, uj = u - ui * W;
#pragma unroll
            for (int k = 0; k < 8; k++) {
                const int ip = ui + di[k], jp = uj + dj[k];
                const bool in = !(ip < 0 || jp < 0 || ip >= H || jp >= W);
                nbr[k] = in ? ip * W + jp : -1;
                dv[k] = in ? dist[nbr[k]] : -1.0f;
            }
        }
#pragma unroll
        for (int k = 0; k < 8; k++) {
            const int v = nbr[k];
            if (v < 0) continue;
            const float nd = du + ((k & 1) ? 1.0f : s2);
            if (nd < dv[k]) {
                parents[v] = u;
                dist[v] = nd;
                if (!inq[v]) {
                    if (++tail == cap) tail = 0;
                    live++;
                    ring[tail] = v;
                    inq[v] = 1;
                    int pf = head + 1; if (pf == cap) pf = 0;          // the entry that will be popped next
                    if (nd < dist[ring[pf]]) { ring[tail] = ring[pf]; ring[pf] = v; }
                }
            }
        }
    }
}
BP_DEV void spfa_serial_parents(const uint8_t* map, int H, int W, int si, int sj, float* dist, int* parents, int* ring, unsigned char* inq) {
    spfa_serial_init(map, H * W, dist, parents, inq, 0, 1);
    spfa_serial_run(map, H, W, si, sj, dist, parents, ring, inq);
    const float inf = 2.0f * (float)(H * W);
    for (int v = 0; v < H * W; v++) if (dist[v] < 0.0f) dist[v] = inf;      // blocked cells: unreachable, like the reference's labels
}

BP_DEV void spfa_block(SpfaShared& S, const uint8_t* map, int H, int W, int si, int sj, float* dist_out, int32_t* parents,
                       int* q0, int* q1, int* mark) {
    const int tid = BP_TID, nt = BP_NT;
    const int nv = H * W;
    const float inf = 2.0f * (float)nv;
    const float s2 = sqrtf(2.0f);
    const int di[8] = {-1, -1, -1, 0, 1, 1, 1, 0};
    const int dj[8] = {-1, 0, 1, 1, 1, 0, -1, -1};
    int* dist = (int*)dist_out;
    const int s = si * W + sj;
    spfa_relax(S, map, H, W, si, sj, dist_out, q0, q1, mark);
    // parents + output convention
    for (int v = tid; v < nv; v += nt) {
        float d = i2f(dist[v]);
        int par = -1;
        if (v != s && d < inf && map[v]) {
            int vi = v / W, vj = v - vi * W;
            for (int k = 0; k < 8 && par < 0; k++) {
                int ip = vi - di[k], jp = vj - dj[k];       // u such that u + dir[k] = v
                if (ip < 0 || jp < 0 || ip >= H || jp >= W) continue;
                int u = ip * W + jp;
                if (!map[u]) continue;
                float du = i2f(dist[u]);
                if (du < inf && du + ((k & 1) ? 1.0f : s2) == d) par = u;
            }
        }
        parents[v] = par;
    }
    BP_BSYNC();
    for (int v = tid; v < nv; v += nt) { float d = i2f(dist[v]); dist_out[v] = (d < inf) ? d : 0.0f; }
    BP_BSYNC();
}

}  // namespace bp
