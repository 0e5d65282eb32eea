"""bayesiannetwork_b200 -- batched loopy belief propagation (Pearl pi/lambda) on B200.

Host-side mirror of the reference's ``bn::inference::belief_propagation`` for the one hot path
this repository accelerates.  All arithmetic runs in hand-written sm_100a CUDA kernels behind the
C ABI of ``include/bnbp.h`` (``lib/libbnbp.so``); there is no CPU fallback.
"""
from .flat import EvidenceBatch, FlatNetwork  # noqa: F401

__all__ = ["FlatNetwork", "EvidenceBatch", "BeliefPropagation", "loopy_belief_propagation"]


def __getattr__(name):
    if name in ("BeliefPropagation", "loopy_belief_propagation", "BnbpError"):
        from . import engine
        return getattr(engine, name)
    raise AttributeError(name)
