"""smoe_stereo_b200 -- B200-native (sm_100a) RAFT-Stereo correlation hot path of SMoE-Stereo.

Public surface = the reference's for this path (core/corr.py, sampler/):
    CorrBlock1D, CorrBlockFast1D, PytorchAlternateCorrBlock1D, CorrSampler, corr_sampler
plus ``install()`` to make them the implementations an unmodified reference checkout uses.
"""
from . import _lib                                             # noqa: F401
from . import corr_sampler                                     # noqa: F401
from .corr import (AlternateCorrBlock, CorrBlock1D, CorrBlockFast1D, CorrSampler,  # noqa: F401
                   FusedPyramid, PytorchAlternateCorrBlock1D, set_default_output_ring,
                   set_default_volume_dtype)
from .dropin import install, uninstall                        # noqa: F401
from .sharding import shard_bounds, gather_lookup, ShardedCorrBlock1D   # noqa: F401

__version__ = "0.1.0"
