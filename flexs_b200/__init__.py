"""flexs_b200 — B200-native virtual-screen hot path behind the FLEXS plugin API.

``import flexs_b200 as flexs`` gives the reference's surface for that path:
``Landscape`` / ``Model`` / ``LandscapeAsModel`` / ``Ensemble`` / ``Explorer`` / ``evaluate`` /
``baselines.models.{CNN,MLP}`` / ``baselines.explorers.{Adalead,CbAS,CMAES,DynaPPO}`` /
``utils.sequence_utils`` / ``landscapes.{TFBinding,AdditiveAAVPackaging}``.  Scoring and training run in hand-written sm_100a kernels inside
``libflexs_b200.so`` (C ABI: include/flexs_b200.h); importing the package needs no GPU, using a
surrogate does (there is no CPU fallback).
"""
from flexs_b200 import types  # noqa: F401
from flexs_b200.landscape import Landscape  # noqa: F401
from flexs_b200.model import LandscapeAsModel, Model  # noqa: F401
from flexs_b200.ensemble import Ensemble  # noqa: F401
from flexs_b200.explorer import Explorer  # noqa: F401
from flexs_b200 import utils  # noqa: F401
from flexs_b200 import baselines, evaluate, landscapes  # noqa: F401

__version__ = "0.1.0"
