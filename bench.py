#!/usr/bin/env python
"""bench.py — env-steps/sec of ship-ice-v0 (BASELINE.json metric) on N B200s of one node.

    python bench.py --gpus 1 --steps K --warmup W            # our arm (CUDA, C ABI)
    python bench.py --impl reference --gpus 1 --steps K ...  # the restated reference on the host cores
    torchrun --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Workload (config.workload): BASELINE config 5 — ship-ice-v0, concentration 0.5 (~640 floes per env,
synthetic seeded layouts), 65 536 envs in total split evenly over the ranks (strong scaling, as the
metric is quoted "@65k envs on 1/2/4/8 B200"), random actions U(-1,1), next-step auto reset.
One "step" = one `step()` of every env of the rank: 400 physics substeps + work/reward + 4x150x150 u8
observation.  `value` = env steps actually simulated (reset calls are not counted as env steps, their
cost stays in the timed region) / device time, max over ranks.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np

METRIC = "env-steps/sec ship-ice-v0 @65k envs"
UNIT = "env-steps/s"
TOTAL_ENVS = 65536
CONC = 0.5
N_LAYOUTS = 100


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=4)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--envs", type=int, default=TOTAL_ENVS, help="total envs over all ranks (default: the metric's 65536)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-fp32", action="store_true", help="skip the secondary measurement of the fp32-state build (field fp32_state)")
    ap.add_argument("--cpu-seconds", type=float, default=15.0)
    ap.add_argument("--env", default="ship-ice", choices=["ship-ice", "box-delivery", "area-clearing", "maze"],
                    help="ship-ice (default, the BASELINE.json metric); box-delivery (BASELINE config 2: 4096 envs) and area-clearing (config 3: 16384 envs) print secondary lines")
    ap.add_argument("--obstacle-config", default="small_empty", choices=["small_empty", "small_columns", "large_columns", "large_divider"],
                    help="box-delivery room (cfg.env.obstacle_config); the column / divider rooms use a pool of --rooms layouts with their own maps")
    ap.add_argument("--rooms", type=int, default=64, help="with --host-rooms: size of the pool of host-built rooms")
    ap.add_argument("--host-rooms", action="store_true", help="column / divider rooms: build the maps on the host (cv2 / scipy) per pool entry instead of on the device at reset")
    ap.add_argument("--fast-parents", dest="exact_parents", action="store_false",
                    help="box-delivery / area-clearing: deterministic spfa parents from the parallel relaxation instead of the reference-order ones (default)")
    ap.add_argument("--desync", type=int, default=34,
                    help="spread the episode phases before the warm-up: env e is advanced (global id mod DESYNC) steps (1 = all envs start at reset)")
    return ap.parse_args()


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md: 6.65 TB/s)"


# ------------------------------------------------------------------------------------------ CPU arm (oracle)
def _cpu_worker(args):
    """One host process: one env of the workload stepped with the oracle for ~budget seconds (or nsteps env steps).
    ``pre`` untimed env steps come first, so that the timed steps sit at the same mix of episode phases as the GPU arm
    (an episode lasts >= 34 steps and gets more expensive as the ship enters the ice)."""
    idx, seed, budget, nsteps = args[:4]
    pre = args[4] if len(args) > 4 else 0
    from benchpush_b200.layouts import generate_experiment
    from oracle.ship_ice_oracle import ShipIceOracleEnv
    trials = generate_experiment(1, CONC, seed=99 + idx)
    env = ShipIceOracleEnv({"concentration": CONC}, experiment=trials, order_mode=1)
    env.reset()
    rng = np.random.default_rng(seed)
    for _ in range(pre):
        o, r, term, _, info = env.step(float(rng.uniform(-1, 1)))
        if term:
            env.reset()
    t0 = time.perf_counter()
    n = 0
    while True:
        o, r, term, _, info = env.step(float(rng.uniform(-1, 1)))
        n += 1
        if term:
            env.reset()
        if nsteps is not None:
            if n >= nsteps:
                break
        elif time.perf_counter() - t0 > budget:
            break
    return n, time.perf_counter() - t0


def _phase(i, cores, desync):
    """Untimed env steps of host core i before its timed sample: the cores are spread evenly over one episode."""
    return (i * max(desync, 1)) // max(cores, 1) if desync > 1 else 0


def cpu_sample(budget_s, nsteps=None, cores=None, desync=34):
    """Env-steps/s of the oracle with one env per host core, all cores busy at once."""
    import multiprocessing as mp
    cores = cores or (os.cpu_count() or 1)
    ctx = mp.get_context("spawn")
    with ctx.Pool(cores) as pool:
        res = pool.map(_cpu_worker, [(i, 1000 + i, budget_s, nsteps, _phase(i, cores, desync)) for i in range(cores)])
    steps = sum(r[0] for r in res)
    wall = max(r[1] for r in res)
    return steps / wall, cores, steps, wall


def run_reference(a):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    # each bench "step" = every host core advances its env by one env step (a bounded sample of the workload)
    total = a.warmup + a.steps
    import multiprocessing as mp
    ctx = mp.get_context("spawn")
    t_all = time.perf_counter()
    with ctx.Pool(cores) as pool:
        # every core: reset, (its episode-phase offset + the warm-up steps) untimed, then the timed steps on the same env
        res = pool.map(_cpu_worker, [(i, 2000 + i, 0.0, a.steps, _phase(i, cores, a.desync) + a.warmup) for i in range(cores)])
    steps = sum(r[0] for r in res)
    wall = max(r[1] for r in res)
    val = steps / wall
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": a.gpus, "steps": a.steps, "warmup": a.warmup,
        "ms_per_step": 1e3 * wall / max(a.steps, 1), "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": "ship-ice-v0 concentration 0.5 (synthetic seeded layouts), one env per host core, random actions",
                   "episode_phase": "core i advanced (i * %d // cores) + warm-up steps untimed before its timed steps (same spread over an episode as the GPU arm)" % a.desync,
                   "note": "the reference env cannot run here (pymunk/gymnasium absent): this is the restated reference (oracle/, fp64 C++ Chipmunk restatement + numpy rasters)"},
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": "%d env steps per core on %d cores (one env per core, incl. observation + reward)" % (a.steps, cores)},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------ clocks sampler
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.idx = gpu_index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.idx), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.rows.append(ln.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for k, nm in enumerate(names):
                if f[5 + k].lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------ our arm
def run_ours(a):
    import torch
    import torch.distributed as dist
    from benchpush_b200.envs import ShipIceVecEnv
    from benchpush_b200.layouts import generate_experiment, pack_trials
    from benchpush_b200.metrics import ShipIceMetric
    from benchpush_b200.parallel import shard_range

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", local))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    lo, hi = shard_range(a.envs, rank, world)
    E = hi - lo

    # layouts: every rank generates the same seeded set; trial selection is keyed by the GLOBAL env id
    packed = pack_trials(generate_experiment(N_LAYOUTS, CONC, seed=99))
    first = (np.arange(lo, hi) % N_LAYOUTS).astype(np.int32)
    env = ShipIceVecEnv(num_envs=E, cfg={"concentration": CONC}, layouts=packed, device=local, auto_reset=True, first_trial=first)
    env.check_status = False
    state_bytes = env.env_bytes()
    n_floes = packed.n_floes[first].astype(np.float64)
    sum_v = np.array([packed.n_verts[l, : packed.n_floes[l]].sum() for l in range(N_LAYOUTS)], dtype=np.float64)[first]
    # algorithmic bytes per env step (SURVEY.md §8d / BASELINE.md §3): B = 56 N + 8 sum(V) + 90 036
    alg_bytes_per_env = float(np.mean(56.0 * n_floes + 8.0 * sum_v + 90036.0))

    obs, info = env.reset()
    gen = torch.Generator(device=dev)
    gen.manual_seed(1234 + rank)
    # steady state: an episode lasts >= 34 steps and its cost grows as the ship enters the ice, so the envs are
    # first spread over the episode phases (untimed): env with global id g takes (g mod DESYNC) steps
    if a.desync > 1:
        phase = torch.arange(lo, hi, device=dev) % a.desync
        for r in range(a.desync - 1):
            act = (torch.rand(E, generator=gen, device=dev) * 2 - 1).float()
            env.step_async_nocheck(act, mask=(phase > r))
        torch.cuda.synchronize()
    total = a.warmup + a.steps
    actions = [(torch.rand(E, generator=gen, device=dev) * 2 - 1).float() for _ in range(total)]
    resets_in_step = torch.zeros(total + 1, dtype=torch.int64, device=dev)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def one_step(k):
        env.step_async_nocheck(actions[k])
        # terminations of this step are reset by the next step call (auto reset): not env steps
        resets_in_step[k + 1] = env._terminated.sum()

    for k in range(a.warmup):
        one_step(k)
    barrier()
    env.kernel_ms()
    env.set_profiling(True)
    launches0 = env.launch_count()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for k in range(a.warmup, total):
        one_step(k)
    e1.record()
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    ms = e0.elapsed_time(e1)
    km = env.kernel_ms()
    env.set_profiling(False)
    launches = env.launch_count() - launches0
    n_resets = int(resets_in_step[a.warmup: total].sum().item())
    env_steps = E * a.steps - n_resets
    status_bits = int(env._status.max().item())

    # ---- end to end through the public host-buffer API: actions H2D (pinned) + results D2H every step
    h_actions = [torch.empty(E, dtype=torch.float32).pin_memory().copy_(actions[k].cpu()) for k in range(a.warmup, total)]
    hb = env.step_host(h_actions[0].numpy(), want_obs="device")               # allocate host staging outside the timed region
    # episode scoring (ShipIceMetric) runs on the host results, inside the end-to-end region, as evaluate() does
    def host_info(hb):
        return {"state": torch.from_numpy(hb["state"]), "total_work": torch.from_numpy(hb["total_work"]),
                "trial_success": torch.from_numpy(hb["trial_success"]).bool()}
    metric = ShipIceMetric("random", ship_mass=env.cfg.ship.mass, goal=env.goal)
    metric.reset(host_info(hb))
    prev_done = torch.from_numpy(hb["terminated"]).bool().clone()
    barrier()
    e2e_l0 = env.launch_count()
    t0 = time.perf_counter()
    e2e_resets = 0
    for k in range(a.steps):
        # the step kernel AND the observation kernel run every step; the 4x150x150 image stays in the env's device tensor
        hb = env.step_host(h_actions[k].numpy(), want_obs="device")
        hi_ = host_info(hb)
        if bool(prev_done.any()):
            metric.reset(hi_, mask=prev_done)                                 # next-step auto reset: this call returned the reset state
        term = torch.from_numpy(hb["terminated"]).bool()
        metric.update(hi_, torch.from_numpy(hb["reward"]), term)
        prev_done = term.clone()
        e2e_resets += int(hb["terminated"].sum())
    torch.cuda.synchronize()
    t_e2e = time.perf_counter() - t0
    e2e_launches = env.launch_count() - e2e_l0
    assert e2e_launches == 2 * a.steps, "end-to-end region must launch the step and the observation kernel every step (%d launches in %d steps)" % (e2e_launches, a.steps)
    # terminations of step k are resets of step k+1; the last step's are not inside the window, the priming step's are
    e2e_steps = E * a.steps - e2e_resets
    h2d = E * 4
    d2h = E * (8 + 1 + 1 + 24 + 8 + 1 + 4)

    t = torch.tensor([ms, t_e2e, float(env_steps), float(e2e_steps), km["step_ms"], km["obs_ms"], float(launches), float(status_bits)],
                     dtype=torch.float64, device=dev)
    if world > 1:
        tmax = t.clone(); dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        tsum = t.clone(); dist.all_reduce(tsum, op=dist.ReduceOp.SUM)
    else:
        tmax, tsum = t, t
    ms_max, e2e_max = float(tmax[0]), float(tmax[1])
    value = float(tsum[2]) / (ms_max * 1e-3)
    e2e_value = float(tsum[3]) / e2e_max
    red = metric.reduce()                                                   # NCCL all-reduce of episode-metric sums (only collective)

    # ---- secondary measurement: the fp32-state build of the same sources (libbenchpush_b200_f32.so), same workload and timing
    # rules; reported next to the headline, which stays the fp64 build (bit-comparable with the oracle)
    fp32 = None
    if not a.no_fp32:
        try:
            env.close()
            del env
            torch.cuda.empty_cache()
            env32 = ShipIceVecEnv(num_envs=E, cfg={"concentration": CONC}, layouts=packed, device=local, auto_reset=True, first_trial=first, precision="fp32")
            env32.check_status = False
            env32.reset()
            g32 = torch.Generator(device=dev); g32.manual_seed(1234 + rank)
            if a.desync > 1:
                for r in range(a.desync - 1):
                    env32.step_async_nocheck((torch.rand(E, generator=g32, device=dev) * 2 - 1).float(), mask=(phase > r))
            resets32 = torch.zeros(total + 1, dtype=torch.int64, device=dev)
            for k in range(a.warmup):
                env32.step_async_nocheck(actions[k])
            barrier()
            f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            f0.record()
            for k in range(a.warmup, total):
                env32.step_async_nocheck(actions[k])
                resets32[k + 1] = env32._terminated.sum()
            f1.record()
            barrier()
            ms32 = f0.elapsed_time(f1)
            steps32 = E * a.steps - int(resets32[a.warmup: total].sum().item())
            t32 = torch.tensor([ms32, float(steps32)], dtype=torch.float64, device=dev)
            if world > 1:
                m32 = t32.clone(); dist.all_reduce(m32, op=dist.ReduceOp.MAX)
                s32 = t32.clone(); dist.all_reduce(s32, op=dist.ReduceOp.SUM)
            else:
                m32, s32 = t32, t32
            fp32 = {"value": float(s32[1]) / (float(m32[0]) * 1e-3), "unit": UNIT, "ms_per_step": float(m32[0]) / a.steps, "dtype": "f32 body state / contacts / solver, f64 motion bounds / work / outputs",
                    "env_steps_timed": int(s32[1]), "env_bytes": int(env32.env_bytes()),
                    "parity": "tolerance (tests/test_gpu_fp32.py): <= 1e-3 m / 1e-2 rad per env step vs the fp64 oracle, rasters bit-exact on the state, 200 seeded episodes statistically indistinguishable"}
            env32.close()
        except Exception as ex:
            fp32 = {"value": None, "error": repr(ex)}
            if world > 1:
                raise

    cpu = None
    if rank == 0 and not a.no_cpu_baseline:
        try:
            v, cores, n, wall = cpu_sample(a.cpu_seconds, desync=a.desync)
            cpu = {"value": v, "unit": UNIT, "cores": cores, "kind": "port",
                   "sample": "oracle (restated reference, fp64): one env per core on %d cores, %d env steps of the same workload in %.1f s, "
                             "cores spread over the episode phases like the GPU envs" % (cores, n, wall)}
        except Exception as ex:                                               # the GPU numbers stand on their own
            cpu = {"value": None, "unit": UNIT, "cores": 0, "kind": "port", "sample": "failed: %r" % (ex,)}

    if rank == 0:
        peak, peak_src = peaks()
        step_ms_avg = km["step_ms"] / max(km["step_launches"], 1)
        # algorithmic bytes of one launch = bytes per env step x the envs that STEPPED in it (auto-reset envs do not step)
        stepped_per_launch = env_steps / max(a.steps, 1)
        alg_bytes_launch = alg_bytes_per_env * stepped_per_launch
        achieved = alg_bytes_launch / (step_ms_avg * 1e-3) / 1e9 if step_ms_avg > 0 else 0.0
        traffic, traffic_src = None, None
        tp = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(tp):
            try:
                tj = json.load(open(tp))
                # the ncu capture was taken on the full-GPU kernel; a small per-GPU batch runs another instantiation: no figure
                if E >= int(tj.get("min_envs_per_gpu", 0)):
                    traffic = tj.get("dram_bytes_per_env_step") * stepped_per_launch
                    traffic_src = tj.get("source")
            except Exception:
                traffic = None
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": a.steps, "warmup": a.warmup,
            "ms_per_step": ms_max / a.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": "ship-ice-v0 concentration 0.5 (BASELINE config 5), %d envs total (%d per GPU), 400 substeps/step, "
                                   "4x150x150 u8 obs, U(-1,1) actions, next-step auto reset" % (a.envs, E),
                       "layouts": "%d synthetic seeded ice fields (benchpush_b200.layouts.generate_ice_field, seed 99), mean %.0f floes" % (N_LAYOUTS, float(n_floes.mean())),
                       "l2": "inputs_exceed_l2 (per-env state %.2f MB x %d envs per GPU)" % (state_bytes / 1e6, E),
                       "episode_phase": "steady state: env g advanced (g mod %d) steps before the warm-up, auto reset inside the timed region" % a.desync if a.desync > 1 else "all envs start at reset",
                       "env_steps_timed": int(tsum[2]), "resets_timed": n_resets, "env_status_bits": int(tmax[7])},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                         "traffic_source": traffic_src, "envs_stepped_per_launch": stepped_per_launch,
                         "kernel": "k_shipice_step", "kernel_ms": step_ms_avg, "obs_kernel_ms": km["obs_ms"] / max(km["obs_launches"], 1),
                         "algorithmic_bytes_per_env_step": alg_bytes_per_env, "peak_source": peak_src,
                         "note": "path is latency/issue bound by 400 sequential substeps (SURVEY.md §8d); HBM is the ceiling, not the limiter"},
            "cpu_baseline": cpu,
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "launches_in_region": int(e2e_launches),
                    "what": "bp_shipice_step_host_obs_device: actions from pinned host memory, step kernel + observation kernel every step, "
                            "reward/terminated/truncated/state/total_work/success/status read back every step and scored on the host (ShipIceMetric); "
                            "the 4x150x150 u8 images are rendered into the env's device tensor for a policy on the GPU and not copied"},
            "gpu_launches": int(tsum[6]),
            "episode_metrics": red,
            "fp32_state": fp32,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


# ------------------------------------------------------------------------------------------ box-delivery (secondary line)
BD_METRIC = "env-steps/sec box-delivery-v0 @4096 envs"
BD_CFG = {"agent": {"action_type": "heading"}}


def bd_cfg(obstacle_config):
    return {"agent": {"action_type": "heading"}, "env": {"obstacle_config": obstacle_config}}


def _bd_cpu_worker(args):
    idx, seed, budget, nsteps = args[:4]
    if len(args) > 4 and args[4] == "area-clearing":
        import random
        from oracle.area_clearing_oracle import AreaClearingOracleEnv, sample_layout as ac_layout
        env = AreaClearingOracleEnv(layouts=[ac_layout(random.Random(1 + idx))])
    else:
        from oracle.box_delivery_oracle import BoxDeliveryOracleEnv, sample_layout
        cfg = bd_cfg(args[5] if len(args) > 5 else "small_empty")
        env = BoxDeliveryOracleEnv(cfg=cfg, layouts=[sample_layout(np.random.RandomState(42 + idx), cfg)])
    env.reset()
    rng = np.random.default_rng(seed)
    t0 = time.perf_counter()
    n = 0
    while True:
        o, r, term, trunc, info = env.step(float(rng.uniform(-1, 1)))
        n += 1
        if term or trunc:
            env.reset()
        if (nsteps is not None and n >= nsteps) or (nsteps is None and time.perf_counter() - t0 > budget):
            break
    return n, time.perf_counter() - t0


def _bd_cpu(budget_s, nsteps=None, kind="box-delivery", oc="small_empty"):
    import multiprocessing as mp
    cores = os.cpu_count() or 1
    with mp.get_context("spawn").Pool(cores) as pool:
        res = pool.map(_bd_cpu_worker, [(i, 1000 + i, budget_s, nsteps, kind, oc) for i in range(cores)])
    steps = sum(r[0] for r in res)
    wall = max(r[1] for r in res)
    return steps / wall, cores, steps, wall


def run_boxdel(a):
    """BASELINE config 2: box-delivery-v0, small_empty room, 10 boxes, heading actions, spfa distance-map observation.
    One step = one step() of every env: action decode, position-control chain, ~800 physics substeps, rewards, 224x224x4 obs."""
    area = a.env == "area-clearing"
    n_envs = a.envs if a.envs != TOTAL_ENVS else (16384 if area else 4096)
    oc = a.obstacle_config
    metric = ("env-steps/sec area-clearing-v0 @%d envs" % n_envs) if area else ("env-steps/sec box-delivery-v0 @%d envs" % n_envs)
    wl = ("area-clearing-v0 (BASELINE config 3): clear_env, 10 boxes of 1 m, heading actions U(-1,1)" if area
          else "box-delivery-v0 (BASELINE config 2): %s room, %d boxes, heading actions U(-1,1)%s" % (
              oc, 10 if oc.startswith("small") else 20,
              "" if oc.endswith("_empty") else (", room maps rebuilt on the device at every reset" if not a.host_rooms else ", a pool of %d host-built rooms" % a.rooms)))
    wl += ", %s spfa parents" % ("reference-order" if a.exact_parents else "deterministic (--fast-parents)")
    if a.impl == "reference":
        if int(os.environ.get("RANK", "0")) != 0:
            return
        _bd_cpu(0.0, max(a.warmup, 1), a.env, oc)
        v, cores, n, wall = _bd_cpu(0.0, a.steps, a.env, oc)
        print(json.dumps({"impl": "reference", "metric": metric, "value": v, "unit": UNIT, "n_gpus": a.gpus, "steps": a.steps, "warmup": a.warmup,
                          "ms_per_step": 1e3 * wall / max(a.steps, 1), "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
                          "data": "synthetic", "config": {"workload": wl + "; one env per host core",
                                                           "note": "restated reference (oracle/ on the C++ Chipmunk restatement); the reference env cannot run here"},
                          "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "sample": "%d env steps per core on %d cores" % (a.steps, cores)},
                          "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}), flush=True)
        return
    import torch
    from benchpush_b200.envs import AreaClearingVecEnv, BoxDeliveryVecEnv
    torch.cuda.set_device(0)
    dev = torch.device("cuda", 0)
    env = AreaClearingVecEnv(num_envs=n_envs, device=0, exact_spfa_parents=a.exact_parents) if area else BoxDeliveryVecEnv(num_envs=n_envs, cfg=bd_cfg(oc), device=0, max_rooms=a.rooms, device_rooms=(False if a.host_rooms else None),
                                                                                         exact_spfa_parents=a.exact_parents)
    env.reset()
    gen = torch.Generator(device=dev); gen.manual_seed(1234)
    total = a.warmup + a.steps
    actions = [(torch.rand(n_envs, generator=gen, device=dev) * 2 - 1).float() for _ in range(total)]
    for k in range(a.warmup):
        env.step(actions[k])
    torch.cuda.synchronize()
    env.kernel_ms(); env.set_profiling(True)
    l0 = env.launch_count()
    sampler = ClockSampler(0); sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for k in range(a.warmup, total):
        env.step(actions[k])
    e1.record(); torch.cuda.synchronize()
    clocks = sampler.stop()
    ms = e0.elapsed_time(e1)
    km = env.kernel_ms(); env.set_profiling(False)
    launches = env.launch_count() - l0
    # end to end: actions from pinned host memory, rewards / flags / info read back every step
    h_act = [torch.empty(n_envs, dtype=torch.float32).pin_memory().copy_(actions[k].cpu()) for k in range(a.warmup, total)]
    env.step_host(h_act[0].numpy(), want_obs="device")
    torch.cuda.synchronize()
    l1 = env.launch_count()
    t0 = time.perf_counter()
    for k in range(a.steps):
        hb = env.step_host(h_act[k].numpy(), want_obs="device")      # every kernel of a step incl. the observation kernel
    t_e2e = time.perf_counter() - t0
    e2e_launches = env.launch_count() - l1
    assert e2e_launches == launches, "the end-to-end region must launch what the device-timed region launches (%d vs %d)" % (e2e_launches, launches)
    NB = env.num_boxes
    alg = 4 + 224 * 224 * 4 + 2 * (1 + NB) * 6 * 8 + 48           # action + observation + poses read and written + outputs
    move_ms = km["move_ms"] / max(km["steps"], 1)
    peak, peak_src = peaks()
    achieved = alg * n_envs / (move_ms * 1e-3) / 1e9 if move_ms > 0 else 0.0
    cpu = None
    if not a.no_cpu_baseline:
        v, cores, n, wall = _bd_cpu(a.cpu_seconds, None, a.env, oc)
        cpu = {"value": v, "unit": UNIT, "cores": cores, "kind": "port",
               "sample": "oracle (restated reference): one env per core on %d cores, %d env steps in %.1f s" % (cores, n, wall)}
    print(json.dumps({
        "metric": metric, "value": n_envs * a.steps / (ms * 1e-3), "unit": UNIT, "n_gpus": 1, "steps": a.steps, "warmup": a.warmup,
        "ms_per_step": ms / a.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": wl + ", %d envs, 224x224x4 u8 obs with spfa distance channels" % n_envs,
                   "l2": "inputs_exceed_l2 (per-env state %.0f KB x %d envs + %.0f MB spfa scratch)" % (env._L.bp_boxdel_env_bytes(env._h) / 1e3, n_envs, 592 * 3.8),
                   "env_status_bits": int(env._buf["info_i"][:, 3].max()), "boxes_delivered_or_cleared": int(env._buf["info_i"][:, 0].sum())},
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": None,
                     "kernel": "k_bd_move", "kernel_ms": move_ms, "obs_kernel_ms": km["obs_ms"] / max(km["steps"], 1),
                     "algorithmic_bytes_per_env_step": alg, "peak_source": peak_src,
                     "note": "latency bound: ~800 dependent physics substeps per env step, 1-3 contacts each"},
        "cpu_baseline": cpu, "clocks": clocks,
        "e2e": {"value": n_envs * a.steps / t_e2e, "unit": UNIT, "h2d_bytes_per_step": n_envs * 4, "d2h_bytes_per_step": n_envs * (8 + 1 + 1 + 64 + 16 + NB),
                "launches_in_region": int(e2e_launches),
                "what": "bp_boxdel_step_host_obs_device: actions from pinned host memory; every kernel of the step incl. the observation kernel; reward / flags / info / completion read back every step; the 224x224x4 u8 images are rendered into the env's device tensor for a policy on the GPU and not copied"},
        "gpu_launches": launches}), flush=True)


# ------------------------------------------------------------------------------------------ maze-NAMO (secondary line, BASELINE config 4)
def _maze_cpu_worker(args):
    idx, budget, nsteps = args
    from oracle.maze_oracle import MazeNamoOracleEnv
    from benchpush_b200.layouts import generate_maze_layouts
    env = MazeNamoOracleEnv(layouts=[generate_maze_layouts(64, seed=0)[idx % 64]], order_mode=1)
    env.reset()
    rng = np.random.default_rng(500 + idx)
    n = 0
    t0 = time.perf_counter()
    while True:
        obs, r, term, trunc, info = env.step(float(np.float32(rng.uniform(-1, 1))))
        n += 1
        if term or trunc:
            env.reset()
        if (nsteps is not None and n >= nsteps) or (nsteps is None and time.perf_counter() - t0 > budget):
            break
    return n, time.perf_counter() - t0


def _maze_cpu(budget_s, nsteps=None):
    import multiprocessing as mp
    cores = os.cpu_count() or 1
    with mp.get_context("spawn").Pool(cores) as pool:
        res = pool.map(_maze_cpu_worker, [(i, budget_s, nsteps) for i in range(cores)])
    steps = sum(r[0] for r in res)
    wall = max(r[1] for r in res)
    return steps / wall, cores, steps, wall


def run_maze(a):
    """BASELINE config 4: maze-NAMO-v0 (maze v1, 5 movable obstacles), 16 384 envs, random actions, next-step auto reset.
    One step = one step() of every env: 100 physics substeps, reward / info, 4x192x192 observation."""
    n_envs = a.envs if a.envs != TOTAL_ENVS else 16384
    metric = "env-steps/sec maze-NAMO-v0 @%d envs" % n_envs
    wl = "maze-NAMO-v0 (BASELINE config 4): maze v1, 5 movable obstacles, 256 seeded layouts, U(-1,1) actions, next-step auto reset"
    if a.impl == "reference":
        if int(os.environ.get("RANK", "0")) != 0:
            return
        _maze_cpu(0.0, max(a.warmup, 1))
        v, cores, n, wall = _maze_cpu(0.0, a.steps)
        print(json.dumps({"impl": "reference", "metric": metric, "value": v, "unit": UNIT, "n_gpus": a.gpus, "steps": a.steps, "warmup": a.warmup,
                          "ms_per_step": 1e3 * wall / max(a.steps, 1), "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
                          "data": "synthetic", "config": {"workload": wl + "; one env per host core",
                                                           "note": "restated reference (oracle/ on the C++ Chipmunk restatement); the reference env cannot run here"},
                          "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "sample": "%d env steps per core on %d cores" % (a.steps, cores)},
                          "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}), flush=True)
        return
    import torch
    from benchpush_b200.envs import MazeNamoVecEnv
    torch.cuda.set_device(0)
    dev = torch.device("cuda", 0)
    env = MazeNamoVecEnv(num_envs=n_envs, device=0, auto_reset=True, n_synthetic_layouts=256)
    env.check_status = False
    env.reset()
    gen = torch.Generator(device=dev); gen.manual_seed(1234)
    total = a.warmup + a.steps
    actions = [(torch.rand(n_envs, generator=gen, device=dev) * 2 - 1).float() for _ in range(total)]
    for k in range(a.warmup):
        env.step_async_nocheck(actions[k])
    torch.cuda.synchronize()
    env.set_profiling(True); env.kernel_ms()
    l0 = env.launch_count()
    sampler = ClockSampler(0); sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for k in range(a.warmup, total):
        env.step_async_nocheck(actions[k])
    e1.record(); torch.cuda.synchronize()
    clocks = sampler.stop()
    ms = e0.elapsed_time(e1)
    km = env.kernel_ms(); env.set_profiling(False)
    launches = env.launch_count() - l0
    h_act = [torch.empty(n_envs, dtype=torch.float32).pin_memory().copy_(actions[k].cpu()) for k in range(a.warmup, total)]
    env.step_host(h_act[0].numpy(), want_obs="device")
    torch.cuda.synchronize()
    l1 = env.launch_count()
    t0 = time.perf_counter()
    for k in range(a.steps):
        hb = env.step_host(h_act[k].numpy(), want_obs="device")
    t_e2e = time.perf_counter() - t0
    e2e_launches = env.launch_count() - l1
    step_ms = km["step_ms"] / max(km["step_launches"], 1)
    obs_ms = km["obs_ms"] / max(km["obs_launches"], 1)
    alg = 4 + 4 * 192 * 192 + 2 * 6 * 6 * 8 + 64              # action + observation + poses of robot and 5 obstacles read and written + outputs
    peak, peak_src = peaks()
    achieved = alg * n_envs / (step_ms * 1e-3) / 1e9 if step_ms > 0 else 0.0
    cpu = None
    if not a.no_cpu_baseline:
        v, cores, n, wall = _maze_cpu(a.cpu_seconds, None)
        cpu = {"value": v, "unit": UNIT, "cores": cores, "kind": "port",
               "sample": "oracle (restated reference): one env per core on %d cores, %d env steps in %.1f s" % (cores, n, wall)}
    d2h = sum(v.nbytes for k, v in hb.items() if isinstance(v, np.ndarray))
    print(json.dumps({
        "metric": metric, "value": n_envs * a.steps / (ms * 1e-3), "unit": UNIT, "n_gpus": 1, "steps": a.steps, "warmup": a.warmup,
        "ms_per_step": ms / a.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": wl + ", %d envs, 4x192x192 u8 obs" % n_envs,
                   "l2": "inputs_exceed_l2 (per-env state %.0f KB x %d envs)" % (env.env_bytes() / 1e3, n_envs),
                   "env_status_bits": int(env._status.max())},
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": None,
                     "kernel": "k_shipice_step (maze handle)", "kernel_ms": step_ms, "obs_kernel_ms": obs_ms,
                     "algorithmic_bytes_per_env_step": alg, "peak_source": peak_src,
                     "note": "latency bound: 100 dependent physics substeps per env step on a handful of bodies"},
        "cpu_baseline": cpu, "clocks": clocks,
        "e2e": {"value": n_envs * a.steps / t_e2e, "unit": UNIT, "h2d_bytes_per_step": n_envs * 4, "d2h_bytes_per_step": int(d2h),
                "launches_in_region": int(e2e_launches),
                "what": "bp_maze_step_host_obs_device: actions from pinned host memory, step kernel + observation kernel every step, reward / flags / state read back every step; images stay in the env's device tensor"},
        "gpu_launches": launches}), flush=True)


def main():
    a = parse()
    if a.env in ("box-delivery", "area-clearing"):
        run_boxdel(a)
    elif a.env == "maze":
        run_maze(a)
    elif a.impl == "reference":
        run_reference(a)
    else:
        run_ours(a)


if __name__ == "__main__":
    main()
