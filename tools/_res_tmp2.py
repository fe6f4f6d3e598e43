import numpy as np, sys
from tests.conftest import SHIP_VERTS
from benchpush_b200.layouts import generate_experiment, pack_trials
from tests.emu.emu import Emu
nl = 5
packed = pack_trials(generate_experiment(nl, 0.5, seed=11))
E = 10
kw = dict(pmax=8 * (packed.max_floes + 1), lmax=2 * (packed.max_floes + 1))
A = Emu(E, packed, SHIP_VERTS, **kw)
for e in range(E): A.reset(e, e % nl)
rng = np.random.default_rng(0)
mode=sys.argv[1]
for k in range(40):
    acts = rng.uniform(-1, 1, E).astype(np.float32)
    for e0 in range(0, E, A.batch_size()):
        (A.step_batch if mode=="batch" else A.step_resident)(e0, acts, True)
    print(k, [A.out(e)["terminated"] for e in range(E)], flush=True)
