"""msd_b200 -- B200-native (sm_100a) restoration hot path of ajammoussi/Multi-Spectral-Satellite-Image-Denoising.

Import as ``msd_b200`` (the repo-root alias package) -- the directory name required by the build contract contains
hyphens.  Sub-packages mirror the reference's ``src`` layout for the hot path only:
    msd_b200.data      AddSensorNoise
    msd_b200.models    SatMAERestoration (+ parameter-holder modules)
    msd_b200.training  CombinedLoss, SSIMLoss, MetricsTracker, calculate_*, evaluate_model
"""
from . import _lib  # noqa: F401
from ._lib import MsdError, LIB_PATH  # noqa: F401

__version__ = "0.1.0"
