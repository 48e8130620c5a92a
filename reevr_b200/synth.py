"""Synthetic signals of SURVEY.md §8(d), shared by bench.py and the tests so that the CPU
reference and the GPU engine see identical bytes."""
import numpy as np


def synth_ir(n_taps: int, channel: int = 0) -> np.ndarray:
    """Gaussian noise x exponential decay reaching -60 dB at the last tap, peak-normalised to 1
    (keeps the tail above the reference's 1e-6 trim threshold, FFTConvolver.cpp:103-106)."""
    rng = np.random.default_rng(4321 + channel)
    g = rng.standard_normal(n_taps)
    tau = n_taps / np.log(1000.0)
    h = g * np.exp(-np.arange(n_taps) / tau)
    h /= np.max(np.abs(h))
    return h.astype(np.float32)


def synth_input(n: int, channel: int = 0) -> np.ndarray:
    """White Gaussian noise, sigma = 0.25."""
    rng = np.random.default_rng(1234 + channel)
    return (0.25 * rng.standard_normal(n)).astype(np.float32)
