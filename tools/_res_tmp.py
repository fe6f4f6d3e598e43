import numpy as np, time, sys
from tests.conftest import SHIP_VERTS
from benchpush_b200.layouts import generate_experiment, pack_trials
from tests.emu.emu import Emu
nl = 5
packed = pack_trials(generate_experiment(nl, 0.5, seed=11))
E = 10
kw = dict(pmax=8 * (packed.max_floes + 1), lmax=2 * (packed.max_floes + 1))
A, B = Emu(E, packed, SHIP_VERTS, **kw), Emu(E, packed, SHIP_VERTS, **kw)
for e in range(E):
    A.reset(e, e % nl); B.reset(e, e % nl)
rng = np.random.default_rng(0)
big = int(np.argmax(packed.n_floes))
ta=tb=0
for k in range(int(sys.argv[1]) if len(sys.argv)>1 else 12):
    acts = rng.uniform(-1, 1, E).astype(np.float32)
    t0=time.time()
    for e0 in range(0, E, A.batch_size()):
        A.step_batch(e0, acts, True)
    t1=time.time()
    for e0 in range(0, E, B.batch_size()):
        B.step_resident(e0, acts, True)
    t2=time.time(); ta+=t1-t0; tb+=t2-t1
    ok=True
    for e in range(E):
        sa, sb = A.scal(e), B.scal(e)
        same = np.array_equal(A.bodies(e, big), B.bodies(e, big)) and A.out(e) == B.out(e)
        if not same:
            ok=False
            d = np.abs(A.bodies(e, big) - B.bodies(e, big)).max()
            print("step", k, "env", e, "MISMATCH maxdiff", d, A.out(e)["reward"], B.out(e)["reward"], sa["n_awake"], sb["n_awake"])
    print("step", k, "ok" if ok else "FAIL", "awake", [A.scal(e)["n_awake"] for e in range(E)])
    if not ok: break
print("batch %.2fs resident %.2fs" % (ta, tb))
