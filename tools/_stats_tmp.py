import numpy as np, time
from tests.conftest import SHIP_VERTS
from benchpush_b200.layouts import generate_experiment, pack_trials
from tests.emu.emu import Emu
NL=12
packed = pack_trials(generate_experiment(NL, 0.5, seed=99))
kw = dict(pmax=8 * (packed.max_floes + 1), lmax=2 * (packed.max_floes + 1))
E = Emu(NL, packed, SHIP_VERTS, **kw)
rng = np.random.default_rng(0)
rows=[]
t0=time.time()
for e in range(NL):
    E.reset(e, e)
    prev = E.scal(e)
    for k in range(60):
        E.step(e, float(rng.uniform(-1, 1)))
        s = E.scal(e)
        d = {x: s[x]-prev[x] for x in ("gjk","pairs_eval","live_sum","awake_sum","recolour","rebuild","substeps","epa","skipped")}
        n = max(d["substeps"],1)
        rows.append((e,k,d["awake_sum"]/n, d["live_sum"]/n, d["gjk"]/n, d["pairs_eval"]/n, d["epa"]/n, d["recolour"], d["rebuild"], s["n_apairs"], s["n_colours"], s["n_cached"]))
        prev = s
        if E.out(e)["terminated"]: break
print(time.time()-t0)
R=np.array(rows)
names=["env","k","awake","live","gjkit","eval","epa","recol","rebuild","apairs","colours","cached"]
for i,n in enumerate(names[2:],2):
    c=R[:,i]
    print("%-8s mean %.2f p50 %.2f p90 %.2f p99 %.2f max %.2f"%(n,c.mean(),np.percentile(c,50),np.percentile(c,90),np.percentile(c,99),c.max()))
print("episode lens", np.bincount(R[:,0].astype(int)))
