"""tracking_b200 -- B200-native foreground-extraction hot path of USTC-Computer-Vision/tracking.

Everything that computes lives in CUDA kernels behind the C ABI of include/bgsb200.h
(tracking_b200/libbgsb200.so, built in-tree by tracking_b200._build).  This package is the
Python host-side mirror of the reference's plugin interface; the C++ drop-in adapters are in
tracking_b200/adapters/.  There is no CPU fallback.
"""
from .bgs import (ALGOS, USTC_BGS, AdaptiveBackgroundLearning, AdaptiveSelectiveBackgroundLearning, DPAdaptiveMedianBGS, DPMeanBGS, DPPratiMediodBGS, DPWrenGABGS,  # noqa: F401
                  DPZivkovicAGMMBGS, FrameDifferenceBGS, SigmaDeltaBGS,
                  MixtureOfGaussianV2BGS, StaticFrameDifferenceBGS, WeightedMovingMeanBGS,
                  WeightedMovingVarianceBGS, pinned_empty, process_fanout)
from .capi import BgsbError, kernel_launch_count  # noqa: F401

__all__ = ["FrameDifferenceBGS", "StaticFrameDifferenceBGS", "WeightedMovingMeanBGS", "WeightedMovingVarianceBGS", "AdaptiveBackgroundLearning", "AdaptiveSelectiveBackgroundLearning", "DPZivkovicAGMMBGS", "DPAdaptiveMedianBGS", "DPMeanBGS", "DPWrenGABGS", "DPPratiMediodBGS", "SigmaDeltaBGS",
           "MixtureOfGaussianV2BGS", "USTC_BGS", "ALGOS", "process_fanout", "pinned_empty", "BgsbError", "kernel_launch_count"]
