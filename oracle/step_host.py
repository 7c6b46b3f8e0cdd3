"""TEST INFRASTRUCTURE ONLY: ctypes wrapper of oracle/_build/libpgd_step_host.so, the host (g++) build of the
role-per-warp step (pgdrive_b200/csrc/pgd_step.cuh).  Same interface as oracle.Oracle so that tests
can roll the two side by side.  Nothing under pgdrive_b200/ imports this."""
import ctypes as C
import os
import subprocess

import numpy as np

from pgdrive_b200 import cabi

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "_build", "libpgd_step_host.so")
_lib = None


def lib():
    global _lib
    if _lib is None:
        subprocess.check_call(["make", "-C", HERE, "_build/libpgd_step_host.so"], stdout=subprocess.DEVNULL)
        L = C.CDLL(LIB)
        vp = C.c_void_p
        L.sth_create.restype = vp
        L.sth_create.argtypes = [C.POINTER(cabi.PgdTables), C.POINTER(cabi.PgdConfig), C.c_int32]
        L.sth_destroy.argtypes = [vp]
        L.sth_set_envs_per_cta.argtypes = [vp, C.c_int32]
        L.sth_reset.argtypes = [vp, vp, vp, C.c_int32, vp, vp]
        L.sth_step.argtypes = [vp, vp, vp, vp, vp, vp]
        L.sth_get_state.argtypes = [vp, C.c_int32, vp]
        _lib = L
    return _lib


class HostStep:
    def __init__(self, T, num_envs, roles=4, envs_per_cta=32, **cfg):
        self.L = lib()
        self.tables, self._keep = cabi.pack_tables(T)
        self.cfg = cabi.make_config(num_envs, **cfg)
        self.n = num_envs
        self.h = self.L.sth_create(C.byref(self.tables), C.byref(self.cfg), int(roles))
        self.L.sth_set_envs_per_cta(self.h, int(envs_per_cta))
        self.obs_dim = cabi.obs_dim(self.cfg)
        self.obs = np.zeros((num_envs, self.obs_dim), np.float32)
        self.reward = np.zeros(num_envs, np.float32)
        self.done = np.zeros(num_envs, np.uint8)
        self.info = np.zeros(num_envs, cabi.INFO_DT)

    def close(self):
        if self.h:
            self.L.sth_destroy(self.h)
            self.h = None

    def reset(self, env_ids, episode_ids):
        ids = np.ascontiguousarray(list(env_ids), np.int32)
        eps = np.ascontiguousarray(list(episode_ids), np.int32)
        self.L.sth_reset(self.h, ids.ctypes.data, eps.ctypes.data, len(ids), self.obs.ctypes.data,
                         self.info.ctypes.data)
        return self.obs

    def step(self, actions):
        a = np.ascontiguousarray(actions, np.float32).reshape(self.n, 2)
        self.L.sth_step(self.h, a.ctypes.data, self.obs.ctypes.data, self.reward.ctypes.data, self.done.ctypes.data,
                        self.info.ctypes.data)
        return self.obs, self.reward, self.done, self.info

    def get_state(self, env):
        s = np.zeros(1, cabi.ENV_STATE_DT)
        self.L.sth_get_state(self.h, int(env), s.ctypes.data)
        return s
