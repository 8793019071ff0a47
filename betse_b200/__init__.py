"""betse_b200 — B200-native (sm_100a, fp64) implementation of BETSE's per-timestep
tissue-update loop (the body of Simulator._run_sim_core_loop, betse/science/sim.py:1132-1390)
behind a C ABI (include/betse_b200.h).  See DESIGN.md and INTEGRATION.md."""
from .capi import BetseB200Error  # noqa: F401

__all__ = ["BetseB200Error", "TissueEngine"]


def __getattr__(name):
    if name == "TissueEngine":
        from .engine import TissueEngine
        return TissueEngine
    raise AttributeError(name)
