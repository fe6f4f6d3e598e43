import numpy as np, time, sys, os
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
for k in range(39):
    acts = rng.uniform(-1, 1, E).astype(np.float32)
    if k == 38:
        sys.stderr.write("=== step 38\n"); sys.stderr.flush()
    # only batch 0 (envs 0..7)
    A.step_batch(0, acts, True)
    if k < 38:
        devnull = os.open(os.devnull, os.O_WRONLY); saved = os.dup(2); os.dup2(devnull, 2)
    B.step_resident(0, acts, True)
    if k < 38:
        os.dup2(saved, 2); os.close(devnull); os.close(saved)
    for e in range(8):
        if not np.array_equal(A.bodies(e, big), B.bodies(e, big)):
            a, b = A.bodies(e, big), B.bodies(e, big)
            bad = np.where(np.abs(a - b).max(axis=1) > 0)[0]
            print("step", k, "env", e, "bodies differ:", bad[:20])
