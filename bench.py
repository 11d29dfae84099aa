#!/usr/bin/env python3
"""bench.py - env steps/sec (= tactile frames/sec) of the batched tactile-RL engine, BASELINE.json config 2:
edge_follow-v0, UR5 + TacTip 128x128, 4096 parallel envs per B200 (weak scaling over --gpus).

    python bench.py --gpus 1 --steps 200 --warmup 10
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port P bench.py --gpus 8 ...
    python bench.py --impl reference ...      # the reference's CPU path on the host cores (unmodified PyBullet reference if the
                                              # box has it, else the CPU oracle port)

One "step" = one env step of every env: action encode -> TCP velocity control -> 24 physics substeps ->
reward/done -> tactile render -> auto-reset of finished envs (+ their first observation).

Numbers on the JSON line:
  value      device-resident throughput: actions already in HBM, obs/reward/done written to torch.cuda tensors;
             each step timed by its own CUDA-event pair on the launching stream, L2 flushed between steps.
  e2e        same metric through the public VecEnv API with HOST numpy actions and HOST numpy results
             (pinned H2D of actions, D2H of obs + reward + done inside the timed region every step).
  roofline   the raster kernel alone (tg_raster_only) against the measured HBM peak: algorithmic bytes per
             env-step (S*S obs + 64 B state = 16,448 at 128x128) x N / CUDA-event duration.
  secondary  BASELINE configs 3, 4, 5 at their per-GPU sizes, short runs of the same protocol (value, physics_ms, raster_ms and
             the raster's roofline fraction with SURVEY 8(d)'s 32,832 / 16,524 / 65,628 B per env-step).
  gather     with --gather (and by default when N > 1): the same steps with the optional collated batch - one packed in-place
             NCCL all-gather per step, overlapped with the next step - next to the plain number.
  cpu_baseline  the CPU oracle (oracle/, "port" - pybullet is not installable here) timed on one host core on a
             bounded sample of the same workload (rank 0, N = 1 only).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

# BASELINE.json configs 2..5 at their per-GPU sizes; alg_bytes = SURVEY.md 8(d)'s algorithmic bytes per env-step
WORKLOADS = {
    "edge": dict(env_id="edge_follow-v0", n=4096, img=128, max_steps=200, act_dim=2, extra=64, kernel="scan_setup_kernel + raster_scan_kernel",
                 modes={"movement_mode": "xy", "control_mode": "TCP_velocity_control", "noise_mode": "rand_height", "observation_mode": "tactile",
                        "reward_mode": "dense", "arm_type": "ur5", "tactile_sensor_name": "tactip"}, label="edge_follow-v0 ur5+tactip"),
    "surface": dict(env_id="surface_follow-v0", n=1024, img=128, max_steps=200, act_dim=3, extra=64 + 64 * 64 * 4, kernel="raster_hf_kernel",
                    modes={"movement_mode": "xyzRxRy", "control_mode": "TCP_velocity_control", "noise_mode": "simplex", "observation_mode": "tactile",
                           "reward_mode": "dense", "arm_type": "ur5", "tactile_sensor_name": "digit"}, label="surface_follow-v0 ur5+digit"),
    "push": dict(env_id="object_push-v0", n=8192, img=128, max_steps=1000, act_dim=2, extra=92 + 48, kernel="scan_setup_kernel + raster_scan_kernel",
                 modes={"movement_mode": "TyRz", "control_mode": "TCP_velocity_control", "rand_init_orn": False, "rand_obj_mass": False,
                        "traj_type": "simplex", "observation_mode": "tactile_and_feature", "reward_mode": "dense", "arm_type": "mg400",
                        "tactile_sensor_name": "digitac"}, label="object_push-v0 mg400+digitac"),
    "balance": dict(env_id="object_balance-v0", n=2048, img=256, max_steps=250, act_dim=2, extra=92, kernel="scan_setup_kernel + raster_scan_kernel",
                    modes={"movement_mode": "xy", "control_mode": "TCP_velocity_control", "object_mode": "pole", "rand_gravity": True,
                           "rand_embed_dist": True, "observation_mode": "tactile", "reward_mode": "dense", "arm_type": "ur5",
                           "tactile_sensor_name": "tactip"}, label="object_balance-v0 ur5+tactip"),
}


def workload(name, n=0):
    w = dict(WORKLOADS[name])
    if n:
        w["n"] = n
    w["name"] = name
    w["alg_bytes"] = w["img"] * w["img"] + w["extra"]
    w["text"] = "%s %dx%d, %d envs/GPU, max_steps %d, actions iid U(-0.25,0.25) seed 0" % (w["label"], w["img"], w["img"], w["n"], w["max_steps"])
    return w


def measured_hbm_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:  # noqa: BLE001
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler(threading.Thread):
    """nvidia-smi clocks + throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.rows, self._stop_evt = index, [], threading.Event()

    def run(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        while not self._stop_evt.is_set():
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q, "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([x.strip() for x in out.split(",")])
            except Exception:  # noqa: BLE001
                pass
            self._stop_evt.wait(0.2)

    def stop(self):
        self._stop_evt.set()
        self.join(timeout=3)
        clk = sorted(int(float(r[0])) for r in self.rows if r and r[0].replace(".", "").isdigit())
        mx = max([int(float(r[1])) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()] or [0])
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in self.rows if len(r) >= 7 for i in range(4) if r[3 + i].lower().startswith("active")})
        return {"sm_mhz": clk[len(clk) // 2] if clk else None, "sm_max_mhz": mx or None, "reasons": reasons, "samples": len(self.rows)}


# ------------------------------------------------------------------------------------------------ CPU arms
def make_oracle_env(w, seed):
    from oracle import oracle as O

    cls = {"object_balance-v0": O.ObjectBalanceOracle, "surface_follow-v0": O.SurfaceFollowOracle, "object_push-v0": O.ObjectPushOracle,
           "edge_follow-v0": O.EdgeFollowOracle}[w["env_id"]]
    return cls(image_size=w["img"], max_steps=w["max_steps"], seed=seed)


def make_cpu_env(w, seed, want_live):
    """-> (env, kind): the UNMODIFIED reference on pybullet when the box has it (tools/live_reference.py), else the oracle port"""
    if want_live:
        sys.path.insert(0, os.path.join(ROOT, "tools"))
        import live_reference as LR

        tg, pb, why = LR.probe()
        if tg is not None:
            env = LR.make_env(w["env_id"], w["modes"], [w["img"], w["img"]], w["max_steps"])
            env.seed(seed)
            return env, "reference"
    return make_oracle_env(w, seed), "port"


def cpu_rate(w, seconds, seed=0):
    """steps/s of the CPU path on ONE core: same env, same action distribution, auto-reset on done."""
    import numpy as np

    env, kind = make_cpu_env(w, seed, True)
    env.reset()
    rng = np.random.RandomState(seed)
    n, t0 = 0, time.perf_counter()
    while True:
        _, _, done, _ = env.step(rng.uniform(-0.25, 0.25, w["act_dim"]).astype(np.float32))
        n += 1
        if done:
            env.reset()
        if n % 50 == 0 and time.perf_counter() - t0 >= seconds:
            break
    return n / (time.perf_counter() - t0), n, kind


_WORKER = None


def _cpu_worker_init(name, n):
    global _WORKER
    import numpy as np

    w = workload(name, n)
    env, kind = make_cpu_env(w, os.getpid(), True)
    env.reset()
    _WORKER = (env, np.random.RandomState(os.getpid()), w, kind)


def _cpu_worker(seconds):
    """one sample on one core: steps completed in `seconds` by this worker's persistent env"""
    import numpy as np

    env, rng, w, kind = _WORKER
    n, t0 = 0, time.perf_counter()
    while True:
        _, _, done, _ = env.step(rng.uniform(-0.25, 0.25, w["act_dim"]).astype(np.float32))
        n += 1
        if done:
            env.reset()
        if n % 20 == 0 and time.perf_counter() - t0 >= seconds:
            break
    return n, time.perf_counter() - t0, kind


def run_reference(args):
    """--impl reference: the reference's own CPU implementation of the path on every host core, one env per process (what
    SubprocVecEnv does, sb3_helpers/rl_utils.py:17-30).  The unmodified reference on PyBullet is tried first
    (tools/live_reference.py); this image and the GPU pool have no pybullet wheel, so in practice the CPU oracle port runs."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import multiprocessing as mp

    from oracle import oracle as O

    O.build()
    w = workload(args.workload, args.envs)
    cores = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    per_step = 3.0  # seconds of CPU work per "step" sample
    vals, kind = [], "port"
    with mp.get_context("spawn").Pool(cores, initializer=_cpu_worker_init, initargs=(args.workload, args.envs)) as pool:
        for k in range(args.warmup + args.steps):
            t0 = time.perf_counter()
            res = pool.map(_cpu_worker, [per_step] * cores, chunksize=1)
            wall = time.perf_counter() - t0
            kind = res[0][2]
            if k >= args.warmup:
                vals.append(sum(r[0] for r in res) / wall)
    value = sum(vals) / len(vals)
    note = ("the UNMODIFIED reference (tactile_gym on pybullet, DIRECT mode), one env per process, all host cores" if kind == "reference"
            else "CPU oracle port (pybullet not importable on this box); one env per process, all host cores")
    line = {
        "impl": "reference", "metric": "env steps/sec (tactile frames/sec)", "value": value, "unit": "env-steps/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 / value,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": w["text"], "note": note},
        "cpu_baseline": {"value": value, "unit": "env-steps/s", "cores": cores, "kind": kind,
                         "sample": "%d samples x %.0f s on %d processes" % (args.steps, per_step, cores)},
        "e2e": {"value": value, "unit": "env-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------------ GPU arm
class Bench:
    def __init__(self, args):
        import numpy as np
        import torch
        import torch.distributed as dist

        self.np, self.torch, self.dist, self.args = np, torch, dist, args
        self.rank, self.world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        if not torch.cuda.is_available():
            raise SystemExit("bench.py needs a CUDA device (there is no CPU fallback)")
        torch.cuda.set_device(self.local)
        self.dev = torch.device("cuda", self.local)
        if self.world > 1:
            os.environ.setdefault("NCCL_DEBUG", "WARN")   # keep NCCL's version banner off stdout: one JSON line only
            dist.init_process_group("nccl", device_id=self.dev)
        self.flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=self.dev)   # > 126 MB L2

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize(self.dev)

    def timed_steps(self, fn, count, first=0):
        """per-step CUDA-event pairs on the launching stream, L2 flushed (untimed) between steps"""
        torch = self.torch
        evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(count)]
        for k in range(count):
            self.flush.fill_(k & 0xff)
            evs[k][0].record()
            fn(first + k)
            evs[k][1].record()
        torch.cuda.synchronize(self.dev)
        return sum(a.elapsed_time(b) for a, b in evs)

    def over_ranks(self, vals):
        """-> (max, min, median) lists over ranks of a list of floats"""
        torch = self.torch
        t = torch.tensor(list(vals), dtype=torch.float64, device=self.dev)
        if self.world == 1:
            v = t.tolist()
            return v, v, v
        g = torch.empty((self.world, t.numel()), dtype=torch.float64, device=self.dev)
        self.dist.all_gather_into_tensor(g, t)
        return g.max(0).values.tolist(), g.min(0).values.tolist(), g.median(0).values.tolist()

    def make(self, w):
        import tactile_gym_b200 as tg

        n = w["n"]
        env = tg.make_vec(w["env_id"], n, env_kwargs={"env_modes": w["modes"], "image_size": [w["img"], w["img"]], "max_steps": w["max_steps"]},
                          device=self.local)
        # global env index -> seed, so results do not depend on how many GPUs share the envs
        env.world.seed([1 + self.rank * n + i for i in range(n)])
        env.reset()
        return env

    def stagger(self, env, w):
        """episode phases uniform in [0, max_steps), as in a long-running job: every timed step carries its share of episode
        ends.  The artificial phase shift makes many envs finish before their pre-computed next episode exists (they complete
        it inline, tg_pipeline_stalls counts them): let the reset pipeline reach its steady state before warm-up."""
        torch, np = self.torch, self.np
        wd = env.world
        st = wd.get_state()
        st[:, 2 * wd.nb + 9] = np.random.RandomState(1000 + self.rank).randint(0, w["max_steps"], size=w["n"])   # the `steps` field
        wd.set_state(st)
        g0 = torch.Generator(device=self.dev); g0.manual_seed(12345 + self.rank)
        # (surface_follow's rebuild - 4,096 OpenSimplex points, IK, a ~110-substep move - spans ~75 launches)
        for _ in range({"push": 8, "surface": 120}.get(w["name"], 40)):
            wd.step((torch.rand((w["n"], wd.act_dim), device=self.dev, generator=g0) - 0.5) * 0.5)
        torch.cuda.synchronize(self.dev)

    def kernels_alone(self, env, acts, reps):
        """raster kernel alone (roofline) and physics kernel alone - after the stepping arms: physics_only steps without resets,
        which bunches the episode ends and would make following steps pay a burst of inline resets"""
        wd = env.world
        for _ in range(3):
            wd.raster_only()
        t_raster = self.timed_steps(lambda k: wd.raster_only(), reps) / reps
        t_phys = self.timed_steps(lambda k: wd.physics_only(acts[k % acts.shape[0]]), reps) / reps
        return t_raster, t_phys

    def secondary(self, name):
        """one of BASELINE configs 3 / 4 / 5 at its per-GPU size: a short run of the headline protocol"""
        torch = self.torch
        w = workload(name)
        K, W = self.args.secondary_steps, 3
        env = self.make(w)
        wd = env.world
        if self.args.phases == "staggered":
            self.stagger(env, w)
        gen = torch.Generator(device=self.dev); gen.manual_seed(self.rank)
        acts = (torch.rand((W + K, w["n"], wd.act_dim), device=self.dev, generator=gen) - 0.5) * 0.5
        for k in range(W):
            wd.step(acts[k])
        self.barrier()
        l0 = wd.launch_count()
        t_ms = self.timed_steps(lambda k: wd.step(acts[k]), K, W)
        launches = wd.launch_count() - l0
        t_raster, t_phys = self.kernels_alone(env, acts, 5)
        (t_ms, t_raster, t_phys), _, _ = self.over_ranks([t_ms, t_raster, t_phys])
        env.close()
        del env, acts
        torch.cuda.empty_cache()
        peak, _ = measured_hbm_peak()
        ach = w["n"] * w["alg_bytes"] / (t_raster * 1e-3) / 1e9
        return {"workload": w["text"], "value": w["n"] * self.world * K / (t_ms * 1e-3), "unit": "env-steps/s", "steps": K, "warmup": W,
                "ms_per_step": t_ms / K, "physics_ms": t_phys, "raster_ms": t_raster, "gpu_launches": int(launches),
                "roofline": {"kernel": w["kernel"], "bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
                             "algorithmic_bytes_per_launch": w["n"] * w["alg_bytes"], "launch_ms": t_raster}}

    def run(self):
        args, torch, np = self.args, self.torch, self.np
        w = workload(args.workload, args.envs)
        n, K, W = w["n"], args.steps, max(args.warmup, 3)
        env = self.make(w)
        wd = env.world
        if args.phases == "staggered":
            self.stagger(env, w)
        gen = torch.Generator(device=self.dev); gen.manual_seed(self.rank)
        acts = (torch.rand((W + K, n, wd.act_dim), device=self.dev, generator=gen) - 0.5) * 0.5
        for k in range(W):
            wd.step(acts[k])
        self.barrier()
        sampler = ClockSampler(self.local) if self.rank == 0 else None
        if sampler:
            sampler.start()
        l0 = wd.launch_count()
        t_ms = self.timed_steps(lambda k: wd.step(acts[k]), K, W)
        launches = wd.launch_count() - l0
        self.barrier()

        # the optional collated batch: one packed in-place all-gather per step on a side stream, overlapped with the next step
        gather = None
        if args.gather or self.world > 1:
            hs = [None]

            def step_gather(k):
                h = env.step_collated(acts[k])
                if hs[0] is not None:
                    env.collated_wait(hs[0])     # the consumer takes batch k-1 while step k runs
                hs[0] = h
            for k in range(W):
                step_gather(k)
            self.barrier()
            t_g = self.timed_steps(step_gather, K, W)
            env.collated_wait(hs[0])
            self.barrier()
            (t_g,), _, _ = self.over_ranks([t_g])
            slot = env._cb.slot.bytes
            gather = {"value": n * self.world * K / (t_g * 1e-3), "unit": "env-steps/s", "ms_per_step": t_g / K,
                      "bytes_per_rank": slot, "bytes_total": slot * self.world,
                      "how": "one in-place all_gather_into_tensor per step on a packed [obs|reward|done|feat] buffer the kernels write into; "
                             "double-buffered, issued on a side stream, consumed one step later"}

        # end to end through the VecEnv API: host numpy in, host numpy out
        a_host = acts.cpu().numpy()
        for k in range(W):      # same warm-up as the device-resident arm
            env.step(a_host[k])
        self.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        Ke = min(K, 200)
        e0.record()
        for k in range(Ke):
            env.step(a_host[W + k])
        e1.record()
        torch.cuda.synchronize(self.dev)
        t_e2e = e0.elapsed_time(e1)
        self.barrier()

        # the floor under e2e: this step's observation bytes over the box's pinned D2H path, all ranks copying at once (PCIe /
        # host memory, no kernels) - what the VecEnv API costs even with infinitely fast kernels
        t_floor = None
        if not env._oracle:
            pin = torch.empty(wd.obs.numel(), dtype=torch.uint8).pin_memory()
            src = wd.obs.view(-1)
            for _ in range(2):
                pin.copy_(src, non_blocking=True)
            self.barrier()
            f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            f0.record()
            for _ in range(10):
                pin.copy_(src, non_blocking=True)
            f1.record()
            torch.cuda.synchronize(self.dev)
            t_floor = f0.elapsed_time(f1) / 10
            (t_floor,), _, _ = self.over_ranks([t_floor])
            del pin
            self.barrier()

        t_raster, t_phys = self.kernels_alone(env, acts, 20)
        clocks = sampler.stop() if sampler else None     # sampled across all timed regions of the headline workload
        mx, mn, md = self.over_ranks([t_ms, t_e2e, t_raster, t_phys])
        t_ms, t_e2e, t_raster, t_phys = mx
        h2d, d2h, host_path = env.h2d_bytes_per_step, env.d2h_bytes_per_step, env._host_step
        lanes = int(wd.cfg.lanes_per_warp)
        env.close()
        del env, acts
        torch.cuda.empty_cache()

        sec = {}
        if args.workload == "edge" and args.secondary_steps > 0:
            for name in ("surface", "push", "balance"):
                sec[name] = self.secondary(name)

        if self.rank == 0:
            peak, peak_src = measured_hbm_peak()
            achieved = n * w["alg_bytes"] / (t_raster * 1e-3) / 1e9
            cpu = None
            if self.world == 1:
                from oracle import oracle as O

                O.build()
                rate, nsteps, kind = cpu_rate(w, args.cpu_seconds)
                cpu = {"value": rate, "unit": "env-steps/s", "cores": 1, "kind": kind,
                       "sample": "%d env steps of the same workload (one env, auto-reset) in %.1f s" % (nsteps, nsteps / rate)}
            traffic = None
            tp = os.path.join(ROOT, "profiles", "raster_traffic.json")
            if os.path.isfile(tp):
                rec = json.load(open(tp)).get(args.workload)
                if rec and rec.get("n_envs") == n and rec.get("image") == w["img"]:     # only for the launch shape that was captured
                    traffic = rec.get("dram_bytes_per_launch")
            line = {
                "metric": "env steps/sec (tactile frames/sec)", "value": n * self.world * K / (t_ms * 1e-3), "unit": "env-steps/s",
                "n_gpus": self.world, "steps": K, "warmup": W, "ms_per_step": t_ms / K, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": {"workload": w["text"], "global_envs": n * self.world,
                           "parallelism": "env-sharded x%d, no data-path collective" % self.world,
                           "l2": "256 MB buffer written between timed steps (L2 flush); per-step CUDA events",
                           "physics_ms": t_phys, "raster_ms": t_raster, "lanes_per_warp": lanes, "episode_phases": args.phases,
                           "per_rank_ms_per_step": {"min": mn[0] / K, "median": md[0] / K, "max": mx[0] / K}},
                "e2e": {"value": n * self.world * Ke / (t_e2e * 1e-3), "unit": "env-steps/s", "h2d_bytes_per_step": h2d,
                        "d2h_bytes_per_step": d2h, "steps": Ke,
                        "per_rank_ms_per_step": {"min": mn[1] / Ke, "median": md[1] / Ke, "max": mx[1] / Ke},
                        "floor": None if t_floor is None else {
                            "value": n * self.world / (t_floor * 1e-3), "unit": "env-steps/s", "ms_per_step": t_floor,
                            "d2h_GBps_aggregate": n * self.world * w["img"] * w["img"] / (t_floor * 1e-3) / 1e9,
                            "what": "the step's observation bytes copied device -> pinned host by all %d ranks at once, nothing else: "
                                    "the PCIe / host-memory ceiling of the numpy VecEnv API on this box" % self.world},
                        "path": ("TactileVecEnv.step (numpy in/out) -> tg_step_host: obs rendered + copied out in chunks, D2H overlapped"
                                 if host_path else "TactileVecEnv.step (numpy in/out) -> tg_step + torch copies")},
                "gpu_launches": int(launches),
                "roofline": {"kernel": w["kernel"], "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                             "traffic": traffic, "peak_source": peak_src, "algorithmic_bytes_per_launch": n * w["alg_bytes"], "launch_ms": t_raster},
                "cpu_baseline": cpu,
                "clocks": clocks,
            }
            if gather:
                line["gather"] = gather
            if sec:
                line["secondary"] = sec
            print(json.dumps(line))
        if self.world > 1:
            self.dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="edge", choices=sorted(WORKLOADS))
    ap.add_argument("--envs", type=int, default=0, help="envs per GPU (0: the workload's)")
    ap.add_argument("--phases", default="staggered", choices=["staggered", "sync"],
                    help="staggered: episode phases uniform in [0, max_steps) as in a long-running job, so every timed "
                         "step carries its share of episode resets; sync: all envs start their episode together")
    ap.add_argument("--cpu-seconds", type=float, default=12.0)
    ap.add_argument("--secondary-steps", type=int, default=10, help="timed steps of each secondary workload (configs 3/4/5); 0: skip")
    ap.add_argument("--gather", action="store_true", help="also time the steps with the collated-batch all-gather (default on when N > 1)")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    import __graft_entry__ as g

    if not os.path.isfile(g.LIB):
        g.build()
    Bench(args).run()


if __name__ == "__main__":
    main()
