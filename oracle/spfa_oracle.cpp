// ORACLE — TEST INFRASTRUCTURE ONLY (see chipmunk_oracle.cpp header for the usage rule).
//
// CPU restatement of the reference's grid shortest-path routine, deps/spfa/src/main.cpp:10-116:
// 8-connected moves between free cells (no corner-cutting check, :45-63), edge weights 1 and
// sqrtf(2) accumulated in fp32 (:14,:74), "infinity" = 2*H*W (:28), unreachable cells reported
// as 0 (:101), parents as raveled indices or -1 (:102).  PINNED: tests compare this port and the
// CUDA kernel against oracle/_ref/spfa (the reference source compiled in place) and against
// tests/golden/spfa_*.npz generated from it.
//
// The port keeps the reference's queue discipline (FIFO with the "smaller label to the front"
// swap, :80-85) so that parents match the reference exactly as well as distances.
#include <cmath>
#include <cstdint>
#include <vector>
#include <utility>

extern "C" void spfa_oracle(const uint8_t* map, int num_rows, int num_cols, int source_i, int source_j,
                            float* out_dists, int32_t* out_parents) {
    const float eps = 1e-6f;
    const int dirs[8][2] = {{-1, -1}, {-1, 0}, {-1, 1}, {0, 1}, {1, 1}, {1, 0}, {1, -1}, {0, -1}};
    const float s2 = std::sqrt(2.0f);
    const float dir_len[8] = {s2, 1, s2, 1, s2, 1, s2, 1};
    const int nv = num_rows * num_cols;
    const float inf = 2.0f * nv;
    std::vector<float> dists(nv, inf);
    std::vector<int32_t> parents(nv, -1);
    std::vector<int32_t> queue((size_t)8 * nv + 2, 0);
    std::vector<uint8_t> in_queue(nv, 0);
    int s = source_i * num_cols + source_j;
    int head = 0, tail = 0;
    dists[s] = 0;
    queue[++tail] = s;
    in_queue[s] = 1;
    while (head < tail) {
        int u = queue[++head];
        in_queue[u] = 0;
        if (!map[u]) continue;                       // a blocked source has no out-edges (:48-49)
        int ui = u / num_cols, uj = u % num_cols;
        for (int k = 0; k < 8; k++) {
            int ip = ui + dirs[k][0], jp = uj + dirs[k][1];
            if (ip < 0 || jp < 0 || ip >= num_rows || jp >= num_cols) continue;
            int vtx = ip * num_cols + jp;
            if (!map[vtx]) continue;
            float nd = dists[u] + dir_len[k];
            if (nd < dists[vtx]) {
                parents[vtx] = u;
                dists[vtx] = nd;
                if (!in_queue[vtx]) {
                    queue[++tail] = vtx;
                    in_queue[vtx] = 1;
                    if (dists[queue[tail]] < dists[queue[head + 1]]) std::swap(queue[tail], queue[head + 1]);
                }
            }
        }
    }
    for (int u = 0; u < nv; u++) {
        out_dists[u] = (dists[u] < inf - eps) * dists[u];
        out_parents[u] = parents[u];
    }
}
