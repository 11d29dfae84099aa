#!/usr/bin/env python3
"""The reference's examples/demo_edge_env.py on the batched engine: one env through the gym.Env surface, then 4,096 through the
VecEnv surface (needs a B200; there is no CPU fallback).

    python examples/demo_edge_env.py
"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import tactile_gym_b200 as tg  # noqa: E402

env_modes = {
    "movement_mode": "xy",                 # "xy" | "xyz" | "xyRz" | "xyzRz"
    "control_mode": "TCP_velocity_control",
    "noise_mode": "rand_height",
    "observation_mode": "tactile",         # "oracle" | "tactile"
    "reward_mode": "dense",                # "dense" | "sparse"
    "arm_type": "ur5",                     # "ur5" | "mg400"
    "tactile_sensor_name": "tactip",       # "tactip" | "digit" | "digitac"
}


def main():
    # --- one env, the gym.Env surface (old-gym API: reset() -> obs dict, step(a) -> obs, reward, done, info)
    env = tg.make("edge_follow-v0", max_steps=250, image_size=[128, 128], env_modes=env_modes)
    env.seed(0)
    obs = env.reset()
    print("observation:", {k: (v.shape, v.dtype) for k, v in obs.items()}, "action space:", env.action_space.shape)
    ret = 0.0
    for _ in range(50):
        obs, reward, done, info = env.step(np.array([0.25, 0.0], dtype=np.float32))
        ret += reward
        if done:
            obs = env.reset()
    print("return over 50 steps: %.3f" % ret)
    env.close()

    # --- 4,096 envs, the VecEnv surface SB3 uses (numpy in / numpy out, auto-reset, terminal_observation + episode infos)
    n = 4096
    venv = tg.make_vec("edge_follow-v0", n, seed=1, env_kwargs={"env_modes": env_modes, "image_size": [128, 128], "max_steps": 200})
    obs = venv.reset()
    rng = np.random.RandomState(0)
    t0 = time.perf_counter()
    steps = 100
    for _ in range(steps):
        obs, rew, done, infos = venv.step(rng.uniform(-0.25, 0.25, (n, 2)).astype(np.float32))
    dt = time.perf_counter() - t0
    print("%d envs x %d steps in %.2f s -> %.2f M env-steps/s through numpy" % (n, steps, dt, n * steps / dt / 1e6))
    venv.close()


if __name__ == "__main__":
    main()
