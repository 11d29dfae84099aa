"""SB3-VecEnv-shaped front end of the batched engine.

Stands where stable_baselines3's SubprocVecEnv stands in the reference's training set-up
(tactile_gym/sb3_helpers/rl_utils.py:17-35): same surface (num_envs, observation_space, action_space,
reset, step_async / step_wait / step, seed, close, get_attr / set_attr / env_method, env_is_wrapped),
numpy in / numpy out, auto-reset with infos[i]["terminal_observation"], Monitor-style
infos[i]["episode"] = {"r", "l", "t"}.  `step_tensor` is the device-resident path (no host copies).
"""
import time

import numpy as np

from . import spaces
from .engine import (TactileWorld, edge_follow_config, object_balance_config, object_push_config, object_roll_config, surface_follow_config,
                     surface_follow_goal_config, surface_follow_vert_config)

try:  # pragma: no cover
    from stable_baselines3.common.vec_env import VecEnv as _VecEnvBase
except Exception:  # noqa: BLE001
    _VecEnvBase = object

CONFIG_BUILDERS = {"edge_follow-v0": edge_follow_config, "object_balance-v0": object_balance_config,
                   "surface_follow-v0": surface_follow_config, "surface_follow-v1": surface_follow_goal_config, "surface_follow-v2": surface_follow_vert_config, "object_push-v0": object_push_config, "object_roll-v0": object_roll_config}


class _Info(dict):
    """info dict of an env that did not finish.  They are reused from step to step (4,096 fresh dicts per step cost more host
    time than the kernels); one that a wrapper or callback wrote into registers itself and is emptied before the next step, so
    nothing leaks from one step into the next."""
    __slots__ = ("_dirty",)

    def _touch(self):
        self._dirty.append(self)

    def __setitem__(self, k, v):
        self._touch(); dict.__setitem__(self, k, v)

    def __delitem__(self, k):
        self._touch(); dict.__delitem__(self, k)

    def update(self, *a, **kw):
        self._touch(); dict.update(self, *a, **kw)

    def setdefault(self, k, d=None):
        self._touch(); return dict.setdefault(self, k, d)

    def pop(self, *a):
        self._touch(); return dict.pop(self, *a)

    def popitem(self):
        self._touch(); return dict.popitem(self)

    def clear(self):
        dict.clear(self)

    def __ior__(self, other):
        self._touch(); dict.update(self, other); return self


class TactileVecEnv(_VecEnvBase):
    def __init__(self, env_id, n_envs, seed=None, env_kwargs=None, device=0, lanes_per_warp=0, copy_chunks=0, rng=None):
        kw = dict(env_kwargs or {})
        if env_id not in CONFIG_BUILDERS:
            raise NotImplementedError("%s is not built yet in tactile_gym_b200" % env_id)
        env_modes = kw.get("env_modes")
        if env_modes is None:
            raise ValueError("env_kwargs['env_modes'] is required")
        if kw.get("show_gui") or kw.get("show_tactile"):
            raise ValueError("show_gui / show_tactile are not available in the batched engine")
        self.observation_mode = env_modes.get("observation_mode", "tactile")
        if self.observation_mode not in ("oracle", "tactile") and not (self.observation_mode == "tactile_and_feature" and env_id in ("object_push-v0", "object_roll-v0", "surface_follow-v1", "surface_follow-v2")):
            raise NotImplementedError("observation_mode %r is not built for %s" % (self.observation_mode, env_id))
        image_size = kw.get("image_size", [64, 64])
        max_steps = kw.get("max_steps", 250)
        built = CONFIG_BUILDERS[env_id](env_modes, image_size, max_steps, n_envs, lanes_per_warp=lanes_per_warp)
        cfg, keep, draw = built if len(built) == 3 else (built[0], built[1], None)
        self.world = TactileWorld(cfg, keep, device=device, draw_fn=draw, rng=rng)
        self.num_envs = n_envs
        # tg_step_host needs the standby reset pipeline (episodes of >= 2 steps); copy_chunks < 0 forces the torch copy path
        self.copy_chunks = copy_chunks
        self._host_step = copy_chunks >= 0 and max_steps >= 2
        S = int(image_size[0])
        # "oracle": the task's state vector instead of the image (get_oracle_obs); nothing is rendered in that mode
        self._oracle = self.observation_mode == "oracle"
        sp = {}
        if self._oracle:
            self.world.bind_oracle_obs()
            self._noracle = self.world.n_oracle
            sp["oracle"] = spaces.Box(low=-np.inf, high=np.inf, shape=(self._noracle,), dtype=np.float32)
        else:
            sp["tactile"] = spaces.Box(low=0, high=255, shape=(S, S, 1), dtype=np.uint8)
        self._with_feat = self.observation_mode == "tactile_and_feature"
        self._nfeat = self.world.nfeat
        if self._with_feat:
            sp["extended_feature"] = spaces.Box(low=-np.inf, high=np.inf, shape=(self._nfeat,), dtype=np.float32)
        self.observation_space = spaces.Dict(sp)
        self.action_space = spaces.Box(low=-0.25, high=0.25, shape=(self.world.act_dim,), dtype=np.float32)
        self.metadata = {"render.modes": ["rgb_array"]}
        self.render_mode = "rgb_array"
        self._actions = None
        self._ep_ret = np.zeros(n_envs, dtype=np.float64)
        self._ep_len = np.zeros(n_envs, dtype=np.int64)
        self._t0 = time.time()
        torch = self.world.torch
        self._pin_actions = torch.zeros((n_envs, self.world.act_dim), dtype=torch.float32).pin_memory()
        # two pinned observation buffers, used alternately: the arrays handed out are views (no 64 MB host copy);
        # an observation stays valid until the step after next
        self._pin_obs2 = [torch.zeros((1 if self._oracle else n_envs, S, S, 1), dtype=torch.uint8).pin_memory() for _ in range(2)]
        self._pin_oracle = torch.zeros((n_envs, self.world.oracle_obs.shape[1]), dtype=torch.float32).pin_memory() if self._oracle else None
        self._flip = 0
        self._pin_obs = self._pin_obs2[0]
        self._pin_rew = torch.zeros(n_envs, dtype=torch.float32).pin_memory()
        self._pin_done = torch.zeros(n_envs, dtype=torch.uint8).pin_memory()
        self._pin_feat = torch.zeros((n_envs, 12), dtype=torch.float32).pin_memory() if self._with_feat else None
        self._dirty_infos = []
        self._blank_infos = [_Info() for _ in range(n_envs)]
        for d in self._blank_infos:
            d._dirty = self._dirty_infos
        # terminal observations of the envs that finish in a step: gathered on the device, copied out through a pinned stage
        # that grows (geometrically) to the largest number of simultaneous episode ends seen
        self._pin_idx = torch.zeros(n_envs, dtype=torch.int64).pin_memory()
        self._term_dev = self._term_stage = None
        self._grow_term_stage(min(n_envs, 64))
        # tg_step_host delivers the finished envs' terminal observations compacted, with the step's own copies: a pinned stage
        # of `cap` rows (copied every step whatever the count, so it follows the number of episode ends really seen)
        self._tstage = None
        self._tstage_used = False
        self._grow_tstage(3 * n_envs // max(int(max_steps), 1))    # three times a steady state's episode ends per step
        if _VecEnvBase is not object:  # pragma: no cover - needs stable_baselines3
            # SB3's wrappers read reset_infos / _seeds / _options / render_mode off the base class
            self._sb3_init()
        if seed is not None:
            self.seed(seed)
        self.h2d_bytes_per_step = self._pin_actions.numel() * 4

    @property
    def d2h_bytes_per_step(self):
        """bytes a step copies device -> host: observations, rewards, dones, features, and the terminal-observation stage"""
        b = (self._pin_oracle.numel() * 4 if self._oracle else self._pin_obs.numel()) + self._pin_rew.numel() * 4 + self._pin_done.numel()
        b += self._pin_feat.numel() * 4 if self._with_feat else 0
        if self._host_step and self._tstage is not None:
            b += self._tstage[0].numel() + self._tstage[1].numel() * 4 + (self._tstage[2].numel() * 4 if self._tstage[2] is not None else 0)
        return b

    def _sb3_init(self):  # pragma: no cover - needs stable_baselines3
        import inspect

        kw = {}
        if "render_mode" in inspect.signature(_VecEnvBase.__init__).parameters:
            kw["render_mode"] = self.render_mode
        try:
            _VecEnvBase.__init__(self, self.num_envs, self.observation_space, self.action_space, **kw)
        except TypeError:
            _VecEnvBase.__init__(self, self.num_envs, self.observation_space, self.action_space)

    # ---------------------------------------------------------------- VecEnv API (host numpy)
    def seed(self, seed=None):
        return self.world.seed(seed)

    def _next_obs_buffer(self):
        self._flip ^= 1
        self._pin_obs = self._pin_obs2[self._flip]
        return self._pin_obs

    def _obs_dict(self):
        if self._oracle:
            return {"oracle": self._pin_oracle.numpy()[:, : self._noracle].copy()}
        o = {"tactile": self._pin_obs.numpy()}
        if self._with_feat:
            o["extended_feature"] = self._pin_feat.numpy()[:, : self._nfeat].copy()
        return o

    def reset(self, seed=None, options=None):
        """seed / options: accepted for gymnasium-era callers (seed re-seeds env i with seed + i first)"""
        if seed is not None:
            self.seed(seed)
        self.world.reset(render=not self._oracle)
        if self._oracle:
            self._pin_oracle.copy_(self.world.oracle_obs, non_blocking=True)
        else:
            self._next_obs_buffer().copy_(self.world.obs, non_blocking=True)
        if self._with_feat:
            self._pin_feat.copy_(self.world.feat, non_blocking=True)
        self.world.torch.cuda.synchronize(self.world.device)
        self._ep_ret[:] = 0
        self._ep_len[:] = 0
        return self._obs_dict()

    def step_async(self, actions):
        torch = self.world.torch
        self._pin_actions.copy_(torch.from_numpy(np.ascontiguousarray(actions, dtype=np.float32).reshape(self.num_envs, -1)))
        if self._host_step:
            # one C-ABI call with the pinned host buffers: chunked raster, D2H overlapped on the library's copy stream
            self.world.step_host(self._pin_actions, None if self._oracle else self._next_obs_buffer(), self._pin_rew, self._pin_done,
                                 h_feat=self._pin_feat, h_oracle=self._pin_oracle, want_terminal_obs=True, chunks=self.copy_chunks, term=self._tstage)
            self._tstage_used = self._tstage is not None
            return
        self._tstage_used = False
        a = self._pin_actions.to(self.world.device, non_blocking=True)
        self.world.step(a, want_terminal_obs=True)
        if self._oracle:
            self._pin_oracle.copy_(self.world.oracle_obs, non_blocking=True)
        else:
            self._next_obs_buffer().copy_(self.world.obs, non_blocking=True)
        self._pin_rew.copy_(self.world.reward, non_blocking=True)
        self._pin_done.copy_(self.world.done, non_blocking=True)
        if self._with_feat:
            self._pin_feat.copy_(self.world.feat, non_blocking=True)

    def _grow_tstage(self, k):
        torch = self.world.torch
        if self._oracle:
            return
        cap = min(self.num_envs, max(32, 1 << int(max(k, 1) - 1).bit_length()))
        if self._tstage is not None and self._tstage[0].shape[0] >= cap:
            return
        S = self.world.S
        self._tstage = (torch.zeros((cap, S, S, 1), dtype=torch.uint8).pin_memory(), torch.zeros(1 + cap, dtype=torch.int32).pin_memory(),
                        torch.zeros((cap, 12), dtype=torch.float32).pin_memory() if self._with_feat else None)

    def _grow_term_stage(self, k):
        torch = self.world.torch
        if self._oracle:
            return
        if self._term_stage is not None and self._term_stage.shape[0] >= k:
            return
        cap = min(self.num_envs, max(64, 1 << int(k - 1).bit_length()))
        S = self.world.S
        self._term_dev = torch.zeros((cap, S, S, 1), dtype=torch.uint8, device=self.world.device)
        self._term_stage = torch.zeros((cap, S, S, 1), dtype=torch.uint8).pin_memory()

    def step_wait(self):
        torch = self.world.torch
        torch.cuda.current_stream(self.world.device).synchronize()
        rew = self._pin_rew.numpy().copy()
        done = self._pin_done.numpy().view(np.bool_).copy()
        self._ep_ret += rew
        self._ep_len += 1
        # envs that did not finish share their (empty) info dict from step to step: 4096 fresh dicts per step cost more
        # host time than the kernels; finished envs get a fresh dict
        while self._dirty_infos:
            self._dirty_infos.pop().clear()
        infos = list(self._blank_infos)
        idx = np.flatnonzero(done)
        if idx.size:
            k = idx.size
            now = round(time.time() - self._t0, 6)
            fast = (not self._oracle) and self._tstage_used and k <= self._tstage[0].shape[0] and int(self._tstage[1][0]) == k
            if not fast:
                self._grow_term_stage(k)
                didx = self._pin_idx[:k]
                didx.copy_(torch.from_numpy(idx))
                didx = didx.to(self.world.device, non_blocking=True)
            if self._oracle:
                term = self.world.term_oracle_obs[didx].cpu().numpy()[:, : self._noracle]
                for j, i in enumerate(idx):
                    infos[i] = {"terminal_observation": {"oracle": term[j]},
                                "episode": {"r": float(self._ep_ret[i]), "l": int(self._ep_len[i]), "t": now}}
            elif fast:
                # they came with the step (tg_step_host's compacted terminal rows, ascending env index like `idx`)
                term = self._tstage[0].numpy()[:k].copy()
                tfeat = self._tstage[2].numpy()[:k].copy() if self._with_feat else None
                for j, i in enumerate(idx):
                    info = {"terminal_observation": {"tactile": term[j]},
                            "episode": {"r": float(self._ep_ret[i]), "l": int(self._ep_len[i]), "t": now}}
                    if self._with_feat:
                        info["terminal_observation"]["extended_feature"] = tfeat[j][: self._nfeat]
                    infos[i] = info
            else:
                # (more episode ends than the stage holds, or the torch-copy path) gather the finished envs' terminal
                # observations on the device, one small pinned copy out
                if self._tstage_used:
                    self._grow_tstage(2 * k)       # the next steps carry them along
                term = self._term_stage[:k]
                torch.index_select(self.world.term_obs, 0, didx, out=self._term_dev[:k])
                term.copy_(self._term_dev[:k], non_blocking=True)
                tfeat = None
                if self._with_feat:
                    tfeat = self.world.term_feat[didx].cpu().numpy()
                torch.cuda.current_stream(self.world.device).synchronize()
                term = term.numpy().copy()
                for j, i in enumerate(idx):
                    info = {"terminal_observation": {"tactile": term[j]},
                            "episode": {"r": float(self._ep_ret[i]), "l": int(self._ep_len[i]), "t": now}}
                    if self._with_feat:
                        info["terminal_observation"]["extended_feature"] = tfeat[j][: self._nfeat]
                    infos[i] = info
            self._ep_ret[idx] = 0
            self._ep_len[idx] = 0
        return self._obs_dict(), rew, done, infos

    def step(self, actions):
        self.step_async(actions)
        return self.step_wait()

    # ---------------------------------------------------------------- device-resident path
    def step_tensor(self, actions):
        """actions: cuda float32 [N, act_dim] -> (obs, reward, done) cuda tensors; nothing synchronises."""
        return self.world.step(actions)

    def step_collated(self, actions, group=None):
        """Device-resident step whose results are all-gathered over the process group into ONE collated batch (the job
        SubprocVecEnv's pipes do in the reference, sb3_helpers/rl_utils.py:17-30): the kernels write into this rank's slot of
        a packed buffer, one in-place NCCL all_gather_into_tensor runs on a side stream, and the call returns at once with a
        handle; `collated_wait(handle)` gives (obs [G, N, S, S, 1], reward [G, N], done [G, N], feat [G, N, 12] | None) in
        global env order, valid until the step after next (two buffers alternate, so the gather overlaps the next step)."""
        from .distributed import CollatedBatch
        from . import _lib as L

        if getattr(self, "_cb", None) is None:
            self._cb = CollatedBatch(self.num_envs, self.world.S, L.TG_PUSH_NFEAT if self.world.nfeat else 0, self.world.device, group=group)
        self.world.step(actions, out=self._cb.local_views())
        return self._cb.gather()

    def collated_wait(self, handle):
        return self._cb.wait(handle)

    def close(self):
        self.world.close()

    def get_attr(self, attr_name, indices=None):
        n = self.num_envs if indices is None else len(np.atleast_1d(indices))
        return [getattr(self, attr_name) for _ in range(n)]

    def set_attr(self, attr_name, value, indices=None):
        setattr(self, attr_name, value)

    def env_method(self, method_name, *args, indices=None, **kwargs):
        n = self.num_envs if indices is None else len(np.atleast_1d(indices))
        return [getattr(self, method_name)(*args, **kwargs) for _ in range(n)]

    def env_is_wrapped(self, wrapper_class, indices=None):
        n = self.num_envs if indices is None else len(np.atleast_1d(indices))
        return [False] * n

    def get_images(self):
        return [np.repeat(o, 3, axis=2) for o in self.world.obs.cpu().numpy()]

    def render(self, mode="rgb_array"):
        return self.get_images()[0]
