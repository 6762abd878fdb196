"""flow_factory_b200 - B200-native (sm_100a) rollout engine behind Flow-Factory's SD3.5 adapter API.

Only the rollout hot path lives here: the trajectory sampler, the MMDiT forward as hand-written
tcgen05/TMEM/TMA kernels, and the fused Euler/SDE + log-prob step, exposed through the C ABI in include/ffb200.h.
"""
from .scheduler import (FlowMatchEulerDiscreteSDEScheduler, SDESchedulerOutput, calculate_shift, make_step_coef,
                        set_scheduler_timesteps)
from .weights import EngineConfig

__all__ = ["FlowMatchEulerDiscreteSDEScheduler", "SDESchedulerOutput", "calculate_shift", "make_step_coef",
           "set_scheduler_timesteps", "EngineConfig", "RolloutEngine"]


def __getattr__(name):
    if name == "RolloutEngine":
        from .engine import RolloutEngine
        return RolloutEngine
    if name in ("B200SD3_5Adapter", "SD3_5Sample"):
        from . import adapter
        return getattr(adapter, name)
    raise AttributeError(name)
