#!/usr/bin/env python3
"""bench.py - env steps/sec (= tactile frames/sec) of the batched tactile-RL engine, BASELINE.json config 2:
edge_follow-v0, UR5 + TacTip 128x128, 4096 parallel envs per B200 (weak scaling over --gpus).

    python bench.py --gpus 1 --steps 200 --warmup 10
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port P bench.py --gpus 8 ...
    python bench.py --impl reference ...      # the CPU restatement (oracle port) on the host cores

One "step" = one env step of every env: action encode -> TCP velocity control -> 24 physics substeps ->
reward/done -> tactile render -> auto-reset of finished envs (+ their first observation).

Numbers on the JSON line:
  value      device-resident throughput: actions already in HBM, obs/reward/done written to torch.cuda tensors;
             each step timed by its own CUDA-event pair on the launching stream, L2 flushed between steps.
  e2e        same metric through the public VecEnv API with HOST numpy actions and HOST numpy results
             (pinned H2D of actions, D2H of obs + reward + done inside the timed region every step).
  roofline   the raster kernel alone (tg_raster_only) against the measured HBM peak: algorithmic bytes per
             env-step (S*S obs + 64 B state = 16,448 at 128x128) x N / CUDA-event duration.
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

MODES = {"movement_mode": "xy", "control_mode": "TCP_velocity_control", "noise_mode": "rand_height",
         "observation_mode": "tactile", "reward_mode": "dense", "arm_type": "ur5", "tactile_sensor_name": "tactip"}
ENV_ID, N_ENVS, IMG, MAX_STEPS = "edge_follow-v0", 4096, 128, 200
ALG_BYTES = IMG * IMG + 64
WORKLOAD = "edge_follow-v0 ur5+tactip %dx%d, %d envs/GPU, max_steps %d, actions iid U(-0.25,0.25) seed 0" % (IMG, IMG, N_ENVS, MAX_STEPS)


def select_workload(name):
    """default: BASELINE config 2 (the one the metric is quoted on).  The others are for the profiles, not the driver:
    'balance': config 5 per GPU (object_balance-v0 ur5+tactip 256x256, 16384 envs over 8 GPUs = 2048 envs/GPU);
    'surface': config 3 per GPU (surface_follow-v0 ur5+digit 128x128, 8192 envs over 8 GPUs = 1024 envs/GPU);
    'push': config 4 (object_push-v0 mg400+digitac 128x128, 8192 envs)."""
    global MODES, ENV_ID, N_ENVS, IMG, MAX_STEPS, ALG_BYTES, WORKLOAD
    if name == "surface":
        MODES = {"movement_mode": "xyzRxRy", "control_mode": "TCP_velocity_control", "noise_mode": "simplex", "observation_mode": "tactile",
                 "reward_mode": "dense", "arm_type": "ur5", "tactile_sensor_name": "digit"}
        ENV_ID, N_ENVS, IMG, MAX_STEPS = "surface_follow-v0", 1024, 128, 200
        ALG_BYTES = IMG * IMG + 64 + 64 * 64 * 4      # + the env's 64x64 f32 heights
        WORKLOAD = "surface_follow-v0 ur5+digit %dx%d, %d envs/GPU, max_steps %d, actions iid U(-0.25,0.25) seed 0" % (IMG, IMG, N_ENVS, MAX_STEPS)
    if name == "push":
        MODES = {"movement_mode": "TyRz", "control_mode": "TCP_velocity_control", "rand_init_orn": False, "rand_obj_mass": False,
                 "traj_type": "simplex", "observation_mode": "tactile_and_feature", "reward_mode": "dense", "arm_type": "mg400",
                 "tactile_sensor_name": "digitac"}
        ENV_ID, N_ENVS, IMG, MAX_STEPS = "object_push-v0", 8192, 128, 1000
        ALG_BYTES = IMG * IMG + 92 + 48               # + 12 f32 features
        WORKLOAD = "object_push-v0 mg400+digitac %dx%d, %d envs/GPU, max_steps %d, actions iid U(-0.25,0.25) seed 0" % (IMG, IMG, N_ENVS, MAX_STEPS)
    if name == "balance":
        MODES = {"movement_mode": "xy", "control_mode": "TCP_velocity_control", "object_mode": "pole", "rand_gravity": True,
                 "rand_embed_dist": True, "observation_mode": "tactile", "reward_mode": "dense", "arm_type": "ur5",
                 "tactile_sensor_name": "tactip"}
        ENV_ID, N_ENVS, IMG, MAX_STEPS = "object_balance-v0", 2048, 256, 250
        ALG_BYTES = IMG * IMG + 92
        WORKLOAD = "object_balance-v0 ur5+tactip %dx%d, %d envs/GPU, max_steps %d, actions iid U(-0.25,0.25) seed 0" % (IMG, IMG, N_ENVS, MAX_STEPS)


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


def make_oracle_env(seed):
    from oracle import oracle as O

    if ENV_ID == "object_balance-v0":
        return O.ObjectBalanceOracle(image_size=IMG, max_steps=MAX_STEPS, seed=seed)
    if ENV_ID == "surface_follow-v0":
        return O.SurfaceFollowOracle(image_size=IMG, max_steps=MAX_STEPS, seed=seed)
    if ENV_ID == "object_push-v0":
        return O.ObjectPushOracle(image_size=IMG, max_steps=MAX_STEPS, seed=seed)
    return O.EdgeFollowOracle(image_size=IMG, max_steps=MAX_STEPS, seed=seed)


ACT_DIM = {"edge_follow-v0": 2, "object_balance-v0": 2, "surface_follow-v0": 3, "object_push-v0": 2}


def cpu_port_rate(seconds, seed=0):
    """steps/s of the CPU oracle on ONE core: same env, same action distribution, auto-reset on done."""
    import numpy as np

    env = make_oracle_env(seed)
    env.reset()
    rng = np.random.RandomState(seed)
    n, t0 = 0, time.perf_counter()
    while True:
        _, _, done, _ = env.step(rng.uniform(-0.25, 0.25, ACT_DIM[ENV_ID]).astype(np.float32))
        n += 1
        if done:
            env.reset()
        if n % 50 == 0 and time.perf_counter() - t0 >= seconds:
            break
    return n / (time.perf_counter() - t0), n


_WORKER_ENV = None


def _cpu_worker_init(workload="edge"):
    global _WORKER_ENV
    import numpy as np

    select_workload(workload)

    env = make_oracle_env(os.getpid())
    env.reset()
    _WORKER_ENV = (env, np.random.RandomState(os.getpid()))


def _cpu_worker(seconds):
    """one sample on one core: steps completed in `seconds` by this worker's persistent env"""
    import numpy as np

    env, rng = _WORKER_ENV
    n, t0 = 0, time.perf_counter()
    while True:
        _, _, done, _ = env.step(rng.uniform(-0.25, 0.25, ACT_DIM[ENV_ID]).astype(np.float32))
        n += 1
        if done:
            env.reset()
        if n % 20 == 0 and time.perf_counter() - t0 >= seconds:
            break
    return n, time.perf_counter() - t0


def run_reference(args):
    """--impl reference: the reference's CPU path.  pybullet cannot be installed in this image (no wheel in
    /opt/wheelhouse, no network), so this times the CPU oracle port on every host core instead."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import multiprocessing as mp

    from oracle import oracle as O

    O.build()
    cores = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    per_step = 3.0  # seconds of CPU work per "step" sample
    vals = []
    with mp.get_context("spawn").Pool(cores, initializer=_cpu_worker_init, initargs=(args.workload,)) as pool:
        for k in range(args.warmup + args.steps):
            t0 = time.perf_counter()
            res = pool.map(_cpu_worker, [per_step] * cores, chunksize=1)
            wall = time.perf_counter() - t0
            if k >= args.warmup:
                vals.append(sum(r[0] for r in res) / wall)
    value = sum(vals) / len(vals)
    line = {
        "impl": "reference", "metric": "env steps/sec (tactile frames/sec)", "value": value, "unit": "env-steps/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 / value,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": WORKLOAD, "note": "CPU oracle port (pybullet not installable here); one env per process, all host cores"},
        "cpu_baseline": {"value": value, "unit": "env-steps/s", "cores": cores, "kind": "port",
                         "sample": "%d samples x %.0f s on %d processes" % (args.steps, per_step, cores)},
        "e2e": {"value": value, "unit": "env-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="edge", choices=["edge", "balance", "surface", "push"])
    ap.add_argument("--envs", type=int, default=0, help="envs per GPU (0: the workload's)")
    ap.add_argument("--phases", default="staggered", choices=["staggered", "sync"],
                    help="staggered: episode phases uniform in [0, max_steps) as in a long-running job, so every timed "
                         "step carries its share of episode resets; sync: all envs start their episode together")
    ap.add_argument("--cpu-seconds", type=float, default=12.0)
    args = ap.parse_args()
    select_workload(args.workload)
    if not args.envs:
        args.envs = N_ENVS
    if args.impl == "reference":
        return run_reference(args)

    import numpy as np
    import torch
    import torch.distributed as dist

    import __graft_entry__ as g

    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (there is no CPU fallback)")
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("NCCL_DEBUG", "WARN")   # keep NCCL's version banner off stdout: one JSON line only
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    if not os.path.isfile(g.LIB):
        g.build()
    import tactile_gym_b200 as tg

    n = args.envs
    K, W = args.steps, max(args.warmup, 3)
    # global env index -> seed, so results do not depend on how many GPUs share the envs
    env = tg.make_vec(ENV_ID, n, env_kwargs={"env_modes": MODES, "image_size": [IMG, IMG], "max_steps": MAX_STEPS}, device=local)
    env.world.seed([1 + rank * n + i for i in range(n)])
    env.reset()
    w = env.world
    dev = w.device
    if args.phases == "staggered":
        st = w.get_state()
        st[:, 2 * w.nb + 9] = np.random.RandomState(1000 + rank).randint(0, MAX_STEPS, size=n)   # the `steps` field
        w.set_state(st)
        # the artificial phase shift makes many envs finish before their pre-computed next episode exists (they complete
        # it inline, tg_pipeline_stalls counts them): let the reset pipeline reach its steady state before warm-up
        g0 = torch.Generator(device=dev); g0.manual_seed(12345 + rank)
        for _ in range(40):
            w.step((torch.rand((n, w.act_dim), device=dev, generator=g0) - 0.5) * 0.5)
        torch.cuda.synchronize(dev)
    gen = torch.Generator(device=dev); gen.manual_seed(rank)
    acts = (torch.rand((W + K, n, w.act_dim), device=dev, generator=gen) - 0.5) * 0.5
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)   # > 126 MB L2

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def timed_steps(fn, count, first):
        """per-step CUDA-event pairs on the launching stream, L2 flushed (untimed) between steps"""
        evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(count)]
        for k in range(count):
            flush.fill_(k & 0xff)
            evs[k][0].record()
            fn(first + k)
            evs[k][1].record()
        torch.cuda.synchronize(dev)
        return sum(a.elapsed_time(b) for a, b in evs)

    for k in range(W):
        w.step(acts[k])
    barrier()
    sampler = ClockSampler(local) if rank == 0 else None
    if sampler:
        sampler.start()
    l0 = w.launch_count()
    t_ms = timed_steps(lambda k: w.step(acts[k]), K, W)
    launches = w.launch_count() - l0
    barrier()

    # end to end through the VecEnv API: host numpy in, host numpy out
    a_host = acts.cpu().numpy()
    for k in range(W):      # same warm-up as the device-resident arm
        env.step(a_host[k])
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    Ke = min(K, 200)
    e0.record()
    for k in range(Ke):
        env.step(a_host[W + k])
    e1.record()
    torch.cuda.synchronize(dev)
    t_e2e = e0.elapsed_time(e1)
    barrier()

    # raster kernel alone (roofline) and physics kernel alone - last: physics_only steps without resets, which bunches the
    # episode ends and would make the next steps pay a burst of inline resets
    for _ in range(3):
        w.raster_only()
    t_raster = timed_steps(lambda k: w.raster_only(), 20, 0) / 20
    t_phys = timed_steps(lambda k: w.physics_only(acts[k % (W + K)]), 20, 0) / 20
    clocks = sampler.stop() if sampler else None     # sampled across all three timed regions

    tt = torch.tensor([t_ms, t_e2e, t_raster, t_phys], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    t_ms, t_e2e, t_raster, t_phys = [float(x) for x in tt.tolist()]

    if rank == 0:
        peak, peak_src = measured_hbm_peak()
        achieved = n * ALG_BYTES / (t_raster * 1e-3) / 1e9
        cpu = None
        if world == 1:
            from oracle import oracle as O

            O.build()
            rate, nsteps = cpu_port_rate(args.cpu_seconds)
            cpu = {"value": rate, "unit": "env-steps/s", "cores": 1, "kind": "port",
                   "sample": "%d env steps of the same workload (one env, auto-reset) in %.1f s" % (nsteps, nsteps / rate)}
        traffic = None
        tp = os.path.join(ROOT, "profiles", "raster_traffic.json")
        if os.path.isfile(tp):
            rec = json.load(open(tp)).get(args.workload)
            if rec and rec.get("n_envs") == n and rec.get("image") == IMG:     # only for the launch shape that was captured
                traffic = rec.get("dram_bytes_per_launch")
        line = {
            "metric": "env steps/sec (tactile frames/sec)", "value": n * world * K / (t_ms * 1e-3), "unit": "env-steps/s",
            "n_gpus": world, "steps": K, "warmup": W, "ms_per_step": t_ms / K, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": WORKLOAD, "global_envs": n * world, "parallelism": "env-sharded x%d, no data-path collective" % world,
                       "l2": "256 MB buffer written between timed steps (L2 flush); per-step CUDA events",
                       "physics_ms": t_phys, "raster_ms": t_raster, "lanes_per_warp": int(w.cfg.lanes_per_warp), "episode_phases": args.phases},
            "e2e": {"value": n * world * Ke / (t_e2e * 1e-3), "unit": "env-steps/s", "h2d_bytes_per_step": env.h2d_bytes_per_step,
                    "d2h_bytes_per_step": env.d2h_bytes_per_step, "steps": Ke,
                    "path": "TactileVecEnv.step (numpy in/out) -> tg_step_host: obs rendered + copied out in chunks, D2H overlapped" if env._host_step else "TactileVecEnv.step (numpy in/out) -> tg_step + torch copies"},
            "gpu_launches": int(launches),
            "roofline": {"kernel": "raster_hf_kernel" if ENV_ID.startswith("surface") else "raster_kernel", "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic, "peak_source": peak_src, "algorithmic_bytes_per_launch": n * ALG_BYTES, "launch_ms": t_raster},
            "cpu_baseline": cpu,
            "clocks": clocks,
        }
        print(json.dumps(line))
    env.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
