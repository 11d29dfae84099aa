#!/usr/bin/env python3
"""Where tactile_gym/sb3_helpers/train_agent.py builds its SubprocVecEnv, build a TactileVecEnv instead.
stable_baselines3 is not part of this repository's environment; this script shows the wiring and exits cleanly without it.

    python examples/train_ppo_sb3.py
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import tactile_gym_b200 as tg  # noqa: E402

# the reference's rl_params_ppo for edge_follow-v0 (tactile_gym/sb3_helpers/params/edge_follow_params.py)
rl_params = {
    "env_name": "edge_follow-v0", "max_ep_len": 200, "image_size": [128, 128], "n_envs": 4096, "n_stack": 1, "seed": 1,
    "env_modes": {"movement_mode": "xy", "control_mode": "TCP_velocity_control", "noise_mode": "rand_height", "observation_mode": "tactile",
                  "reward_mode": "dense", "arm_type": "ur5", "tactile_sensor_name": "digitac"},
}


def main():
    try:
        from stable_baselines3 import PPO
        from stable_baselines3.common.vec_env import VecFrameStack, VecTransposeImage
    except ImportError:
        print("stable_baselines3 is not installed: nothing to train with (the env side of the wiring is tg.make_vec below)")
        return
    env = tg.make_vec(rl_params["env_name"], rl_params["n_envs"], seed=rl_params["seed"],
                      env_kwargs={"env_modes": rl_params["env_modes"], "image_size": rl_params["image_size"], "max_steps": rl_params["max_ep_len"]})
    env = VecTransposeImage(VecFrameStack(env, rl_params["n_stack"]))
    model = PPO("MultiInputPolicy", env, n_steps=32, batch_size=4096, verbose=1)
    model.learn(total_timesteps=rl_params["n_envs"] * 32 * 10)
    env.close()


if __name__ == "__main__":
    main()
