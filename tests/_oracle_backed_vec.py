"""CPU stand-in for tg.make_vec built on the oracle, to dry-run the GPU child scripts' own logic (indices, shapes, tolerances)."""
import numpy as np
from oracle import oracle as O


class FakeWorld:
    def __init__(self, envs, nb, act_dim):
        self.envs, self.nb, self.act_dim = envs, nb, act_dim
        self.draws = None
        self.round = 0

    def set_draws(self, d):
        self.draws = d

    def get_state(self):
        st = np.zeros((len(self.envs), 2 * self.nb + 26))
        for i, e in enumerate(self.envs):
            st[i, :self.nb] = e.s.q[:self.nb]; st[i, self.nb:2 * self.nb] = e.s.qd[:self.nb]
            st[i, 2 * self.nb + 9] = e.steps; st[i, 2 * self.nb + 10] = e.last_reset_substeps
        return st

    def pipeline_error(self):
        return False


class FakeVec:
    def __init__(self, env_id, n, env_kwargs=None, **kw):
        m = env_kwargs["env_modes"]; S = env_kwargs["image_size"][0]
        self.mode = m["observation_mode"]
        if env_id == "surface_follow-v2":
            mk = lambda: O.SurfaceFollowOracle(image_size=S, arm=m["arm_type"], sensor=m["tactile_sensor_name"], movement_mode="xRz", variant="vert",
                                               noise_mode=m["noise_mode"], render=self.mode == "tactile")
            act = 2
        else:
            mk = lambda: O.EdgeFollowOracle(image_size=S, arm=m["arm_type"], sensor=m["tactile_sensor_name"], movement_mode=m["movement_mode"],
                                            control_mode=m["control_mode"])
            act = 4
        self.envs = [mk() for _ in range(n)]
        self.world = FakeWorld(self.envs, self.envs[0].m.ndof, act)

    def _obs(self, imgs):
        if self.mode == "oracle":
            return {"oracle": np.array([e.oracle_obs() for e in self.envs], dtype=np.float32)}
        return {"tactile": np.array(imgs)}

    def reset(self):
        imgs = []
        for i, e in enumerate(self.envs):
            d = self.world.draws[i, 0]
            imgs.append(e.reset(draws=(d[0], d[1])))
        return self._obs(imgs)

    def step(self, act):
        imgs, rew, done = [], [], []
        for i, e in enumerate(self.envs):
            o, r, d, _ = e.step(act[i]); imgs.append(o); rew.append(r); done.append(d)
        return self._obs(imgs), np.array(rew, dtype=np.float32), np.array(done), [{} for _ in self.envs]

    def close(self):
        pass
