"""robotdynamics.jl_b200 — B200-native batched dynamics / Jacobian evaluation for the RobotDynamics.jl hot path.

The directory name contains a dot, so it cannot be imported with a plain `import` statement; the repo root ships
`rdb200.py`, which loads this directory as the package `rdb200`:

    import rdb200 as rd
    model = rd.Cartpole(); dmodel = rd.DiscretizedDynamics(model, rd.RK4)
    rd.jacobian_(rd.StaticReturn(), rd.B200(), dmodel, J, y, Z)

Importing the package loads librdb200.so (built by csrc/build.py); there is no CPU fallback.
"""
from . import _abi
from ._abi import (AOS, SOA, F32, F64, EULER, RK2 as RK2_CODE, Context, ModelHandle, PinnedArray, RegisteredArray, RDBError,
                   NotImplementedModelError, LIB_PATH, context)
from .api import *  # noqa: F401,F403
from .api import (InPlace, StaticReturn, ForwardAD, FiniteDifference, UserDefined, B200, Euler, RK2, RK3, RK4,
                  QuatRotation, UnitQuaternion, MRP, RodriguesParam)

_abi.lib()   # fail at import time, loudly, if the CUDA extension is missing
__version__ = "0.1.0"
