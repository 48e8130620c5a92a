"""reevr_b200 — B200-native partitioned-convolution engine behind the FFTConvolver /
TwoStageFFTConvolver surface of tiagolr/reevr (include/b200conv.h)."""
from .convolver import B200ConvError, Engine, FFTConvolver, StereoConvolver, TwoStageFFTConvolver  # noqa: F401

__all__ = ["Engine", "FFTConvolver", "TwoStageFFTConvolver", "StereoConvolver", "B200ConvError"]
