"""TEST INFRASTRUCTURE: sim objects with the surface of marl_soccer_b200.host_api.HostBufferSim backed by
the CPU checkers (oracle or host harness), injected through the `_sim_factory` test seam to exercise
the Python drop-in classes' host logic without a GPU.  Never used by the product."""
from __future__ import annotations

import numpy as np

import hostsim_lib as H
import oracle_lib as O
import parity_util as P

P.add_batch_api(H.HostSim)


class OracleBackedSim:
    def __init__(self, n, config, seed):
        self.n = n
        self._v = O.OracleVec(n, config, seed=seed)
        self.score = np.zeros((n, 2), np.int32)

    def reset(self, mode=0, seed=None, mask=None):
        return self._v.reset(mode, seed=seed, mask=mask)

    def step(self, actions, auto_reset=True):
        # the score of the step (before auto-reset) is read from the states when auto-reset is off;
        # with auto-reset on, finished envs report the counters captured just before the reset
        obs, rew, done, goal = self._v.step(actions, auto_reset=False)
        st = self._v.get_states()
        self.score = np.array([s["score"] for s in st], np.int32).reshape(self.n, 2)
        if auto_reset:
            for i in np.nonzero(done)[0]:
                e = self._v.env(int(i))
                obs[i] = e.reset(O.MODE_FULL_RANDOM)
        return obs, rew.astype(np.float32), done, goal

    def counters(self):
        st = self._v.get_states()
        return (np.array([s["score"] for s in st], np.int32).reshape(self.n, 2),
                np.array([s["steps"] for s in st], np.int32))

    def get_state(self, i):
        return P.oracle_to_dev_state(self._v.env(i).get_state())

    def close(self):
        pass


class HostSimBacked(H.HostSim):
    def __init__(self, n, config, seed):
        super().__init__(n, config, seed=seed)
        self.score = np.zeros((n, 2), np.int32)

    def counters(self):
        st = [self.get_state(i) for i in range(self.n)]
        return (np.array([[s.score[0], s.score[1]] for s in st], np.int32),
                np.array([s.steps for s in st], np.int32))

    def close(self):
        pass
