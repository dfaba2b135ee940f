"""TEST INFRASTRUCTURE ONLY -- unpickling shim for the absent `chumpy` package.

Only `shapedirs` inside the SMPL pickle is a chumpy object; the hot path never
reads it, it just has to unpickle.
"""
import sys
import types
import numpy as np


class Ch:
    def __setstate__(self, state):
        self.__dict__.update(state if isinstance(state, dict) else {"x": state})

    def __array__(self, dtype=None, copy=None):
        a = np.asarray(self.__dict__.get("x"))
        return a.astype(dtype) if dtype is not None else a


ch = types.ModuleType("chumpy.ch")
ch.Ch = Ch
sys.modules["chumpy.ch"] = ch
