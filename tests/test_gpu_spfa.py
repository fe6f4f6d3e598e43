"""CUDA spfa against the reference's own outputs (golden fixtures generated from
deps/spfa/src/main.cpp compiled in place) and against the oracle port at box-delivery size."""
import numpy as np
import pytest
import torch

from oracle import cpo

pytestmark = pytest.mark.gpu


def _valid_tree(d, p, m):
    H, W = d.shape
    idx = np.argwhere(p >= 0)
    pi, pj = np.divmod(p[p >= 0], W)
    assert np.all(np.maximum(np.abs(idx[:, 0] - pi), np.abs(idx[:, 1] - pj)) == 1)
    assert np.all(m[pi, pj] != 0)
    w = np.where((idx[:, 0] != pi) & (idx[:, 1] != pj), np.sqrt(np.float32(2.0)), np.float32(1.0)).astype(np.float32)
    assert np.array_equal((d[pi, pj] + w).astype(np.float32), d[p >= 0])


def test_spfa_matches_reference_golden(spfa_golden):
    from benchpush_b200.spfa import spfa
    for c in spfa_golden:
        d, p = spfa(c["map"], c["source"])
        assert np.array_equal(d, c["dists"])                     # bit-identical to the reference
        assert np.array_equal(p >= 0, c["parents"] >= 0)
        _valid_tree(d, p, c["map"])


def test_reference_order_parents_are_bit_identical(spfa_golden):
    """reference_parents=True (bp_spfa_batch_exact[_host]): the reference solver's queue restated by one thread per map —
    dists AND parents equal the reference's own outputs (golden fixtures from main.cpp compiled in place), and the oracle
    port at box-delivery size."""
    from benchpush_b200.spfa import spfa, spfa_batch
    for c in spfa_golden:
        d, p = spfa(c["map"], c["source"], reference_parents=True)
        assert np.array_equal(d, c["dists"]) and np.array_equal(p, c["parents"])
    rng = np.random.default_rng(8)
    B, H, W = 3, 430, 542
    maps = np.ones((B, H, W), dtype=np.uint8)
    maps[:, :14, :] = 0; maps[:, -14:, :] = 0; maps[:, :, :14] = 0; maps[:, :, -14:] = 0
    for b in range(B):
        for _ in range(10):
            i, j = rng.integers(20, H - 60), rng.integers(20, W - 60)
            maps[b, i:i + rng.integers(10, 50), j:j + rng.integers(10, 50)] = 0
    src = np.stack([rng.integers(15, 19, B), rng.integers(15, 19, B)], axis=1).astype(np.int32)
    for b in range(B):
        maps[b, src[b, 0], src[b, 1]] = 1
    d, p = spfa_batch(torch.as_tensor(maps, device="cuda"), torch.as_tensor(src, device="cuda"), reference_parents=True)
    d, p = d.cpu().numpy(), p.cpu().numpy()
    ties = 0
    d2, p2 = spfa_batch(torch.as_tensor(maps, device="cuda"), torch.as_tensor(src, device="cuda"))
    for b in range(B):
        od, op = cpo.spfa(maps[b], src[b])
        assert np.array_equal(d[b], od) and np.array_equal(p[b], op)
        ties += int((p2[b].cpu().numpy() != op).sum())
    assert np.array_equal(d2.cpu().numpy(), d)
    assert ties > 0            # the default parents do differ from the reference's on tie cells: that is what the exact mode is for


def test_spfa_batch_box_delivery_size():
    from benchpush_b200.spfa import spfa_batch
    rng = np.random.default_rng(5)
    B, H, W = 6, 430, 542                                          # box-delivery small_* grid
    maps = np.ones((B, H, W), dtype=np.uint8)
    maps[:, :14, :] = 0; maps[:, -14:, :] = 0; maps[:, :, :14] = 0; maps[:, :, -14:] = 0
    for b in range(B):
        for _ in range(12):
            i, j = rng.integers(20, H - 60), rng.integers(20, W - 60)
            maps[b, i:i + rng.integers(10, 50), j:j + rng.integers(10, 50)] = 0
    src = np.stack([rng.integers(15, 19, B), rng.integers(15, 19, B)], axis=1).astype(np.int32)
    for b in range(B):
        maps[b, src[b, 0], src[b, 1]] = 1
    d, p = spfa_batch(torch.as_tensor(maps, device="cuda"), torch.as_tensor(src, device="cuda"))
    d, p = d.cpu().numpy(), p.cpu().numpy()
    for b in range(B):
        od, op = cpo.spfa(maps[b], src[b])
        assert np.array_equal(d[b], od)
        assert np.array_equal(p[b] >= 0, op >= 0)
    _valid_tree(d[0], p[0], maps[0])
    # unreachable / blocked cells report 0 and -1 (main.cpp:101-102)
    assert np.all(d[maps == 0] == 0) and np.all(p[maps == 0] == -1)


def test_spfa_edge_cases():
    from benchpush_b200.spfa import spfa
    d, p = spfa(np.ones((1, 1), dtype=bool), (0, 0))
    assert d.shape == (1, 1) and d[0, 0] == 0 and p[0, 0] == -1
    m = np.ones((5, 5), dtype=bool); m[2, :] = False               # wall: the far side is unreachable
    d, p = spfa(m, (0, 0))
    assert np.all(d[3:] == 0) and np.all(p[3:] == -1) and d[1, 1] == np.float32(np.sqrt(np.float32(2)))
    m = np.ones((4, 4), dtype=bool); m[1, 1] = False               # blocked source reaches nothing
    d, p = spfa(m, (1, 1))
    assert np.all(d == 0) and np.all(p == -1)
    with pytest.raises(ValueError):
        spfa(np.ones((3, 3), dtype=bool), (3, 0))
