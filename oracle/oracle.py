"""ctypes wrapper of the CPU oracle (oracle/pgd_oracle.c).  TEST INFRASTRUCTURE: imported only by
tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs."""
import ctypes as C
import os
import subprocess
import threading

import numpy as np

from pgdrive_b200 import cabi

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "_build", "libpgd_oracle.so")


def build(force=False):
    """make decides what is stale (sources and the shared headers are its prerequisites)."""
    subprocess.check_call(["make", "-C", HERE] + (["-B"] if force else []), stdout=subprocess.DEVNULL)
    return LIB


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(LIB)
        vp, i32 = C.c_void_p, C.c_int32
        L.orc_create.restype = vp
        L.orc_create.argtypes = [C.POINTER(cabi.PgdTables), C.POINTER(cabi.PgdConfig)]
        L.orc_destroy.argtypes = [vp]
        L.orc_set_call_index.argtypes = [vp, C.c_uint32]
        L.orc_set_fast.argtypes = [vp, i32]
        L.orc_reset.argtypes = [vp, i32, i32, vp, vp]
        L.orc_step.argtypes = [vp, i32, vp, vp, vp, vp, vp]
        L.orc_step_range.argtypes = [vp, i32, i32, vp, vp, vp, vp, vp]
        L.orc_get_state.argtypes = [vp, i32, vp]
        L.orc_observe.argtypes = [vp, i32, vp, vp]
        L.orc_set_state.argtypes = [vp, i32, vp]
        L.orc_lane_local.argtypes = [vp, C.c_float, C.c_float, vp]
        L.orc_lane_position.argtypes = [vp, C.c_float, C.c_float, vp]
        L.orc_lane_heading_at.argtypes = [vp, C.c_float]
        L.orc_lane_heading_at.restype = C.c_float
        L.orc_ray_rect.argtypes = [C.c_float] * 9
        L.orc_ray_rect.restype = C.c_float
        _lib = L
    return _lib


class Oracle:
    """N independent environments stepped one after the other on the CPU."""
    def __init__(self, T, num_envs, fast=False, **cfg):
        """``fast``: queries through the maps' bucket grids instead of over every primitive -- same results
        (tests/test_oracle_golden.py); used when the oracle is timed as the CPU baseline."""
        self.L = lib()
        self.tables, self._keep = cabi.pack_tables(T)
        self.cfg = cabi.make_config(num_envs, **cfg)
        self.n = num_envs
        self.h = self.L.orc_create(C.byref(self.tables), C.byref(self.cfg))
        if not self.h:
            raise ValueError("an episode needs more vehicle slots / trigger groups than the simulator has")
        if fast:
            self.L.orc_set_fast(self.h, 1)
        self.obs_dim = cabi.obs_dim(self.cfg)
        self.obs = np.zeros((num_envs, self.obs_dim), np.float32)
        self.reward = np.zeros(num_envs, np.float32)
        self.done = np.zeros(num_envs, np.uint8)
        self.info = np.zeros(num_envs, cabi.INFO_DT)
        self.calls = 0  # API calls so far, like pgd_abi.cu's call_index (one key component of the lidar noise)

    def _count_call(self):
        self.calls += 1
        self.L.orc_set_call_index(self.h, self.calls)

    def close(self):
        if self.h:
            self.L.orc_destroy(self.h)
            self.h = None

    def reset(self, env_ids, episode_ids):
        self._count_call()
        for e, ep in zip(env_ids, episode_ids):
            self.L.orc_reset(self.h, int(e), int(ep), self.obs[e].ctypes.data, self.info[e:e + 1].ctypes.data)
        return self.obs

    def step(self, actions, threads=1):
        self._count_call()
        a = np.ascontiguousarray(actions, np.float32).reshape(self.n, 2)
        args = (a.ctypes.data, self.obs.ctypes.data, self.reward.ctypes.data, self.done.ctypes.data,
                self.info.ctypes.data)
        if threads <= 1:
            self.L.orc_step_range(self.h, 0, self.n, *args)
        else:  # ctypes drops the GIL during the call
            cuts = np.linspace(0, self.n, threads + 1).astype(int)
            ts = [
                threading.Thread(target=self.L.orc_step_range, args=(self.h, int(cuts[i]), int(cuts[i + 1])) + args)
                for i in range(threads)
            ]
            [t.start() for t in ts]
            [t.join() for t in ts]
        return self.obs, self.reward, self.done, self.info

    def observe(self, env):
        """Observation of the current state of ``env`` (no stepping, no mutation)."""
        obs = np.zeros(self.obs_dim, np.float32)
        info = np.zeros(1, cabi.INFO_DT)
        self.L.orc_observe(self.h, int(env), obs.ctypes.data, info.ctypes.data)
        return obs, info[0]

    def get_state(self, env):
        s = np.zeros(1, cabi.ENV_STATE_DT)
        self.L.orc_get_state(self.h, int(env), s.ctypes.data)
        return s

    def set_state(self, env, s):
        s = np.ascontiguousarray(s)
        self.L.orc_set_state(self.h, int(env), s.ctypes.data)
