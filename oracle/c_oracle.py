"""ctypes loader for oracle/point_oracle.c (test infrastructure only)."""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(_HERE, "_build", "libpoint_oracle.so")
        if not os.path.exists(path):
            subprocess.check_call(["make", "-C", _HERE], stdout=subprocess.DEVNULL)
        _LIB = ctypes.CDLL(path)
    return _LIB


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def physics_step(state, action):
    lib().po_physics_step(ctypes.c_int64(state.shape[0]), _p(state), _p(action))


def pos(state):
    out = np.zeros((state.shape[0], 2))
    lib().po_pos(ctypes.c_int64(state.shape[0]), _p(state), _p(out))
    return out


def obs(state, goal):
    out = np.zeros((state.shape[0], 14), np.float32)
    lib().po_obs(ctypes.c_int64(state.shape[0]), _p(state), _p(goal), _p(out))
    return out


def vec_step(state, action, goal, prev_pos, elapsed, time_limit, terminate_on_goal):
    n = state.shape[0]
    o = np.zeros((n, 14), np.float32)
    rew = np.zeros(n)
    reach = np.zeros(n, np.uint8)
    done = np.zeros(n, np.uint8)
    trunc = np.zeros(n, np.uint8)
    lib().po_vec_step(ctypes.c_int64(n), _p(state), _p(action), _p(goal), _p(prev_pos), _p(elapsed),
                      ctypes.c_int32(time_limit or 0), ctypes.c_int(int(terminate_on_goal)),
                      _p(o), _p(rew), _p(reach), _p(done), _p(trunc))
    return o, rew, reach.astype(bool), done.astype(bool), trunc.astype(bool)
