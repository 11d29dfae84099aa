#!/usr/bin/env python3
"""The env construction arguments the reference's own training set-ups use (tests/golden/reference_params.json): every
`rl_params_ppo` / `rl_params_sac` dict of tactile_gym/sb3_helpers/params/*_params.py - env id, max_ep_len, image_size, env_modes -
evaluated from the files' source (ast; the modules themselves import kornia / stable_baselines3, which are not installed).
tests/test_host.py builds the engine's task description from each of them: what `train_agent.py` passes must be accepted as is.
Run in the build container only (needs /root/reference)."""
import ast
import json
import os
import sys

REF = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))
OUT = sys.argv[2] if len(sys.argv) > 2 else os.path.join(os.path.dirname(HERE), "tests", "golden", "reference_params.json")


def main():
    pdir = os.path.join(REF, "tactile_gym", "sb3_helpers", "params")
    out = []
    for fn in sorted(os.listdir(pdir)):
        if not fn.endswith("_params.py"):
            continue
        tree = ast.parse(open(os.path.join(pdir, fn)).read())
        for node in tree.body:
            if isinstance(node, ast.Assign) and isinstance(node.targets[0], ast.Name) and node.targets[0].id in ("rl_params_ppo", "rl_params_sac"):
                ns = {}
                exec(compile(ast.Module(body=[node], type_ignores=[]), fn, "exec"), {"int": int, "float": float}, ns)
                d = ns[node.targets[0].id]
                out.append({"file": fn, "dict": node.targets[0].id, "env_name": d["env_name"], "max_ep_len": d["max_ep_len"],
                            "image_size": d["image_size"], "env_modes": d["env_modes"], "n_stack": d.get("n_stack"), "n_envs": d.get("n_envs")})
    # the demo scripts (examples/demo_*_env.py): the env_modes / image_size / max_steps literals inside main()
    ids = {"EdgeFollowEnv": "edge_follow-v0", "SurfaceFollowAutoEnv": "surface_follow-v0", "SurfaceFollowGoalEnv": "surface_follow-v1",
           "SurfaceFollowVertEnv": "surface_follow-v2", "ObjectRollEnv": "object_roll-v0", "ObjectPushEnv": "object_push-v0",
           "ObjectBalanceEnv": "object_balance-v0"}
    edir = os.path.join(REF, "examples")
    for fn in sorted(os.listdir(edir)):
        if not (fn.startswith("demo_") and fn.endswith("_env.py")):
            continue
        tree = ast.parse(open(os.path.join(edir, fn)).read())
        main_fn = next((n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name == "main"), None)
        if main_fn is None:
            continue
        ns = {}
        for node in main_fn.body:
            if isinstance(node, ast.Assign) and isinstance(node.targets[0], ast.Name) and node.targets[0].id in ("env_modes", "image_size", "max_steps"):
                exec(compile(ast.Module(body=[node], type_ignores=[]), fn, "exec"), {"int": int, "float": float}, ns)
        cls = next((n.func.id for n in ast.walk(main_fn) if isinstance(n, ast.Call) and isinstance(n.func, ast.Name) and n.func.id in ids), None)
        if cls and "env_modes" in ns:
            out.append({"file": "examples/" + fn, "dict": "env_modes", "env_name": ids[cls], "max_ep_len": ns.get("max_steps", 250),
                        "image_size": ns.get("image_size", [128, 128]), "env_modes": ns["env_modes"], "n_stack": None, "n_envs": None})
    json.dump(out, open(OUT, "w"), indent=1, sort_keys=True)
    print("wrote", OUT, len(out), "parameter sets")


if __name__ == "__main__":
    main()
