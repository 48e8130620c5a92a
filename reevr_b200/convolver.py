"""Python mirror of the reference's convolver surface on top of the C ABI (tests / bench glue).

Class and method names follow the reference:
  FFTConvolver          libs/FFTConvolver/FFTConvolver.h:62-80         init / process / clear / reset
  TwoStageFFTConvolver  libs/FFTConvolver/TwoStageFFTConvolver.h:65-83 init / process / reset / clear
  StereoConvolver       src/dsp/StereoConvolver.h:20-30                prepare / loadImpulse / process / reset / clear
`Engine` is the multi-channel handle underneath (one launch set for C channels).
The C++ drop-in classes for the JUCE host live in include/FFTConvolver.h and
include/TwoStageFFTConvolver.h; this module exists because the test-suite and bench are Python.
"""
from __future__ import annotations

import ctypes as C
from typing import Sequence

import numpy as np

from . import _lib


class B200ConvError(RuntimeError):
    pass


def _ptr_array(arrs: Sequence[np.ndarray]):
    arr = (C.c_void_p * len(arrs))()
    for i, a in enumerate(arrs):
        arr[i] = a.ctypes.data
    return arr


class Engine:
    """C mono convolvers in one handle (b200conv_t)."""

    def __init__(self, n_channels: int = 1, device: int = 0, max_batch_blocks: int = 0,
                 shard_rank: int = 0, shard_count: int = 1, cmac_variant: int = 0, lib=None):
        self._l = lib or _lib.default()
        cfg = _lib.Config(n_channels, device, max_batch_blocks, shard_rank, shard_count, cmac_variant)
        self._h = self._l.b200conv_create(C.byref(cfg))
        if not self._h:
            raise B200ConvError("b200conv_create failed")
        self.n_channels = n_channels
        self._reduce_cb = None

    # -- helpers ---------------------------------------------------------------------------
    def _check(self, rc: int, what: str, allow_einval: bool = False) -> int:
        if rc == 0 or (allow_einval and rc == -1):
            return rc
        raise B200ConvError(f"{what} failed ({rc}): {self._l.b200conv_last_error(self._h).decode()}")

    def _irs(self, irs):
        irs = [np.ascontiguousarray(a, dtype=np.float32) for a in irs]
        if len(irs) != self.n_channels:
            raise ValueError("need one IR per channel")
        keep = [a if a.size else np.zeros(1, np.float32) for a in irs]
        lens = (C.c_size_t * len(irs))(*[a.size for a in irs])
        return keep, _ptr_array(keep), lens

    # -- IR load ---------------------------------------------------------------------------
    def init_uniform(self, block: int, irs) -> bool:
        keep, ptrs, lens = self._irs(irs)
        return self._check(self._l.b200conv_init_uniform(self._h, block, ptrs, lens), "init_uniform", True) == 0

    def init_twostage(self, head: int, tail: int, irs) -> bool:
        keep, ptrs, lens = self._irs(irs)
        return self._check(self._l.b200conv_init_twostage(self._h, head, tail, ptrs, lens), "init_twostage", True) == 0

    @staticmethod
    def _shape_params(autogain=True, reverse=False, trim_left=0.0, trim_right=0.0, gain=1.0, lut=None, srate=48000.0,
                      clip=True, attack=0.0, decay=0.0):
        keep = None if lut is None else np.ascontiguousarray(lut, dtype=np.float64)
        sp = _lib.IrShapeParams(int(autogain), int(reverse), trim_left, trim_right, gain,
                                keep.ctypes.data if keep is not None else None, float(srate), int(clip), attack, decay)
        return sp, keep

    def init_twostage_shaped(self, head: int, tail: int, raw_irs, **shape) -> bool:
        """IR shaping (Impulse::recalcImpulse subset) on the device, partition spectra built from the device-resident
        taps (b200conv_init_twostage_shaped); raw_irs: equally long raw channels."""
        keep, ptrs, lens = self._irs(raw_irs)
        sp, lut = self._shape_params(**shape)
        return self._check(self._l.b200conv_init_twostage_shaped(self._h, head, tail, ptrs, keep[0].size, C.byref(sp)),
                           "init_twostage_shaped", True) == 0

    def init_uniform_shaped(self, block: int, raw_irs, **shape) -> bool:
        keep, ptrs, lens = self._irs(raw_irs)
        sp, lut = self._shape_params(**shape)
        return self._check(self._l.b200conv_init_uniform_shaped(self._h, block, ptrs, keep[0].size, C.byref(sp)),
                           "init_uniform_shaped", True) == 0

    def init_stages(self, blocks, offsets, irs) -> bool:
        keep, ptrs, lens = self._irs(irs)
        b = (C.c_size_t * len(blocks))(*blocks)
        o = (C.c_size_t * len(offsets))(*offsets)
        return self._check(self._l.b200conv_init_stages(self._h, len(blocks), b, o, ptrs, lens), "init_stages", True) == 0

    # -- processing ------------------------------------------------------------------------
    def process(self, xs) -> list:
        """xs: C arrays of equal length (host). Returns C float32 arrays."""
        xs = [np.ascontiguousarray(a, dtype=np.float32) for a in xs]
        n = xs[0].size
        n_in = getattr(self, "_n_in", None) or self.n_channels
        n_out = getattr(self, "_n_out", None) or self.n_channels
        if any(a.size != n for a in xs) or len(xs) != n_in:
            raise ValueError("need one equally long input per (routed) input channel")
        ys = [np.empty(max(n, 1), np.float32)[:n] for _ in range(n_out)]
        if n:
            self._check(self._l.b200conv_process(self._h, _ptr_array(xs), _ptr_array(ys), n), "process")
        return ys

    def prime(self, xs) -> None:
        """Feeds history through the convolver in one batched call without producing output
        (IR hot-swap warm-up, src/PluginProcessor.cpp:1695-1750)."""
        xs = [np.ascontiguousarray(a, dtype=np.float32) for a in xs]
        if xs[0].size:
            self._check(self._l.b200conv_prime(self._h, _ptr_array(xs), xs[0].size), "prime")

    def process_xfade(self, new: "Engine", xs, alpha0: float, alpha_step: float) -> list:
        """self = outgoing convolver, `new` = incoming one: device-side crossfade of both outputs
        (src/PluginProcessor.cpp:1800-1830)."""
        xs = [np.ascontiguousarray(a, dtype=np.float32) for a in xs]
        n = xs[0].size
        n_out = getattr(self, "_n_out", None) or self.n_channels
        ys = [np.empty(max(n, 1), np.float32)[:n] for _ in range(n_out)]
        if n:
            self._check(self._l.b200conv_process_xfade(self._h, new._h, _ptr_array(xs), _ptr_array(ys), n,
                                                       alpha0, alpha_step), "process_xfade")
        return ys

    def process_into(self, in_ptrs, out_ptrs, n: int) -> None:
        """Raw host pointers (ctypes arrays of void*), e.g. pinned staging buffers."""
        self._check(self._l.b200conv_process(self._h, in_ptrs, out_ptrs, n), "process")

    def process_device(self, in_ptr: int, in_stride: int, out_ptr: int, out_stride: int, n: int, sync: bool = False):
        self._check(self._l.b200conv_process_device(self._h, in_ptr, in_stride, out_ptr, out_stride, n, int(sync)),
                    "process_device")

    def process_sliced(self, xs, ys, slice_rank: int, slice_count: int) -> None:
        """Time-slice sharding (b200conv_process_sliced): xs / ys are the WHOLE call's host arrays; only this
        rank's slice of every ys[c] is written."""
        n = xs[0].size
        self._check(self._l.b200conv_process_sliced(self._h, _ptr_array(xs), _ptr_array(ys), n, slice_rank, slice_count),
                    "process_sliced")

    def process_sliced_into(self, in_ptrs, out_ptrs, n: int, slice_rank: int, slice_count: int) -> None:
        self._check(self._l.b200conv_process_sliced(self._h, in_ptrs, out_ptrs, n, slice_rank, slice_count), "process_sliced")

    def process_device_sliced(self, in_ptr: int, in_stride: int, out_ptr: int, out_stride: int, n: int,
                              slice_rank: int, slice_count: int, sync: bool = False):
        self._check(self._l.b200conv_process_device_sliced(self._h, in_ptr, in_stride, out_ptr, out_stride, n,
                                                           slice_rank, slice_count, int(sync)), "process_device_sliced")

    def chain_configure(self, srate: float, lowcut_hz: float = 20.0, lowcut_slope: int = 0, highcut_hz: float = 20000.0,
                        highcut_slope: int = 0, predelay: int = 0, width: float = 1.0, drygain: float = 1.0,
                        wetgain: float = 1.0, true_stereo: bool = True) -> None:
        """The send / wet chain of REEVRAudioProcessor::processBlock on the device (b200conv_chain_configure)."""
        cfg = _lib.ChainConfig(srate, lowcut_hz, lowcut_slope, highcut_hz, highcut_slope, predelay, width, drygain, wetgain,
                               int(true_stereo))
        self._check(self._l.b200conv_chain_configure(self._h, C.byref(cfg)), "chain_configure")

    def chain_process(self, dryL, dryR, ysend=None, yrev=None):
        """(outL, outR) = drygain * dry + wetgain * width(yrev * mixdown(convolvers(predelay(filters(dry * ysend)))))"""
        xs = [np.ascontiguousarray(a, dtype=np.float32) for a in (dryL, dryR)]
        n = xs[0].size
        env = [None if e is None else np.ascontiguousarray(e, dtype=np.float32) for e in (ysend, yrev)]
        ys = [np.empty(max(n, 1), np.float32)[:n] for _ in range(2)]
        if n:
            self._check(self._l.b200conv_chain_process(
                self._h, _ptr_array(xs), env[0].ctypes.data if env[0] is not None else None,
                env[1].ctypes.data if env[1] is not None else None, _ptr_array(ys), n), "chain_process")
        return ys[0], ys[1]

    def clear(self):
        self._check(self._l.b200conv_clear(self._h), "clear")

    def reset(self):
        self._check(self._l.b200conv_reset(self._h), "reset")

    # -- introspection ---------------------------------------------------------------------
    def stages(self) -> list:
        out = []
        for s in range(self._l.b200conv_num_stages(self._h)):
            info = _lib.StageInfo()
            self._l.b200conv_stage(self._h, s, C.byref(info))
            out.append(dict(block=info.block, partitions=info.partitions, tap_offset=info.tap_offset,
                            p_begin=info.p_begin, p_end=info.p_end))
        return out

    def ir_len(self, c: int = 0) -> int:
        return int(self._l.b200conv_ir_len(self._h, c))

    @property
    def launch_count(self) -> int:
        return int(self._l.b200conv_launch_count(self._h))

    def last_sweep_variant(self) -> int:
        """22 / 26 packed-FMA batched sweep, 40 tensor-core sweep, 100..108 streaming forms."""
        return int(self._l.b200conv_last_sweep_variant(self._h))

    @property
    def stream(self) -> int:
        return int(self._l.b200conv_stream(self._h) or 0)

    def set_option(self, name: str, value: int):
        """A/B switches of the engine: "rt" (one-launch real-time path), "fft512" (register-resident B = 512 FFTs), "tc" (tensor-core sweep for long launch groups)."""
        self._check(self._l.b200conv_set_option(self._h, name.encode(), int(value)), "set_option")

    def set_timing(self, on: bool):
        self._l.b200conv_set_timing(self._h, int(on))

    def last_timing(self) -> dict:
        a, b, c, n = C.c_float(), C.c_float(), C.c_float(), C.c_int()
        self._l.b200conv_last_timing(self._h, C.byref(a), C.byref(b), C.byref(c), C.byref(n))
        return dict(cmac_ms=a.value, fft_ms=b.value, ifft_ms=c.value, cmac_launches=n.value)

    def set_reduce(self, fn):
        """fn(dev_ptr:int, n_floats:int, stream:int) -> int (0 = ok); sums the buffer into shard 0."""
        def tramp(_user, ptr, n, stream):
            try:
                return int(fn(int(ptr), int(n), int(stream or 0)) or 0)
            except Exception:  # never raise through the C ABI
                import traceback
                traceback.print_exc()
                return -1
        self._reduce_cb = _lib.REDUCE_FN(tramp)
        self._l.b200conv_set_reduce(self._h, self._reduce_cb, None)

    def set_routing(self, in_map, mix):
        """Convolver c reads input in_map[c]; output o = sum_c mix[o][c] * y_c (computed on the device).
        `mix` is an (n_out, C) array.  set_routing(None, None) removes the routing."""
        if in_map is None:
            self._check(self._l.b200conv_set_routing(self._h, 0, None, 0, None), "set_routing")
            self._n_in = self._n_out = None
            return
        mix = np.ascontiguousarray(mix, dtype=np.float32)
        n_out, Cc = mix.shape
        assert Cc == self.n_channels and len(in_map) == self.n_channels
        n_in = max(in_map) + 1
        im = (C.c_int * Cc)(*in_map)
        self._check(self._l.b200conv_set_routing(self._h, n_in, im, n_out,
                                                 mix.ctypes.data_as(C.POINTER(C.c_float))), "set_routing")
        self._n_in, self._n_out = n_in, n_out

    def p2p_attach(self, allgather, mode: int = 0, host_barrier=None):
        """Enables the fused multi-GPU path (slot exchange over peer memory, see b200conv.h).
        `allgather(blob: bytes) -> list[bytes]` must return every shard's blob in rank order.
        mode 0 = CUDA IPC (one process per GPU), 1 = raw pointers (all shards in this process).
        `host_barrier` (callable -> 0) replaces the flag kernel in the CPU emulation build."""
        if host_barrier is not None:
            def tramp(_user):
                try:
                    return int(host_barrier() or 0)
                except Exception:
                    import traceback
                    traceback.print_exc()
                    return -1
            self._barrier_cb = _lib.BARRIER_FN(tramp)
            self._l.b200conv_p2p_set_host_barrier(self._h, self._barrier_cb, None)
        blobs = allgather(self.p2p_export(mode))
        self.p2p_import(blobs)

    def p2p_export(self, mode: int = 0) -> bytes:
        n = self._l.b200conv_p2p_blob_size(self._h)
        blob = C.create_string_buffer(n)
        self._check(self._l.b200conv_p2p_export(self._h, blob, mode), "p2p_export")
        return blob.raw

    def p2p_import(self, blobs) -> None:
        """blobs: every shard's exported blob, in rank order."""
        data = b"".join(blobs)
        joined = C.create_string_buffer(data, len(data))
        self._check(self._l.b200conv_p2p_import(self._h, joined), "p2p_import")

    def p2p_detach(self):
        """Leave the slot-exchange path (every shard must do the same); the reduce hook takes over again."""
        self._check(self._l.b200conv_p2p_detach(self._h), "p2p_detach")

    def p2p_set_input_broadcast(self, on: bool = True):
        """Host-pointer calls: only shard 0 uploads the input; the peers receive it over NVLink."""
        self._check(self._l.b200conv_p2p_set_input_broadcast(self._h, int(on)), "p2p_set_input_broadcast")

    def close(self):
        if getattr(self, "_h", None):
            self._l.b200conv_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def ir_decay_eq(ir, lut, srate: float, device: int = 0, lib=None) -> np.ndarray:
    """Device version of Impulse::applyDecay (src/dsp/Impulse.cpp:602-648): returns the shaped IR."""
    lib = lib or _lib.default()
    buf = np.array(ir, dtype=np.float32, copy=True)
    lut = np.ascontiguousarray(lut, dtype=np.float64)
    if lut.size != 2049:
        raise ValueError("lut needs 2049 entries (4096-point STFT)")
    rc = lib.b200conv_ir_decay_eq(device, buf.ctypes.data, buf.size, lut.ctypes.data, float(srate))
    if rc != 0:
        raise B200ConvError(f"b200conv_ir_decay_eq failed ({rc})")
    return buf


def ir_shape(raw_irs, device: int = 0, lib=None, **shape):
    """Device version of the Impulse::recalcImpulse subset (b200conv_ir_shape); returns the shaped channels."""
    lib = lib or _lib.default()
    raws = [np.ascontiguousarray(a, dtype=np.float32) for a in raw_irs]
    n = raws[0].size
    outs = [np.empty(max(n, 1), np.float32) for _ in raws]
    sp, lut = Engine._shape_params(**shape)
    m = C.c_size_t(0)
    rc = lib.b200conv_ir_shape(device, _ptr_array(raws), len(raws), n, C.byref(sp), _ptr_array(outs), C.byref(m))
    if rc != 0:
        raise B200ConvError(f"b200conv_ir_shape failed ({rc})")
    return [o[:m.value].copy() for o in outs]


class FFTConvolver:
    """Uniform partitioned convolver, one channel (reference: FFTConvolver.h:62-80)."""

    def __init__(self, **kw):
        self._e = Engine(1, **kw)

    def init(self, blockSize: int, ir) -> bool:
        return self._e.init_uniform(blockSize, [ir])

    def process(self, x) -> np.ndarray:
        return self._e.process([x])[0]

    def clear(self):
        self._e.clear()

    def reset(self):
        self._e.reset()


class TwoStageFFTConvolver(FFTConvolver):
    """Head/tail convolver, one channel (reference: TwoStageFFTConvolver.h:65-83)."""

    def init(self, headBlockSize: int, tailBlockSize: int, ir) -> bool:  # type: ignore[override]
        return self._e.init_twostage(headBlockSize, tailBlockSize, [ir])


class StereoConvolver:
    """LL/RR (+LR/RL in quad mode) convolvers of src/dsp/StereoConvolver.{h,cpp} as ONE handle."""

    def __init__(self, **kw):
        self._kw = kw
        self._e = None
        self.isQuad = False
        self.headBlockSize = 0
        self.tailBlockSize = 0
        self.size = 0

    def prepare(self, samplesPerBlock: int):          # StereoConvolver.cpp:8-20
        self.size = samplesPerBlock
        h = 1
        while h < samplesPerBlock:
            h *= 2
        self.headBlockSize = h
        self.tailBlockSize = max(8192, 2 * h)

    def loadImpulse(self, irLL, irRR, irLR=None, irRL=None):   # StereoConvolver.cpp:22-31
        self.isQuad = irLR is not None and irRL is not None
        irs = [irLL, irRR] + ([irLR, irRL] if self.isQuad else [])
        if self._e is None or self._e.n_channels != len(irs):
            self._e = Engine(len(irs), **self._kw)
        return self._e.init_twostage(self.headBlockSize, self.tailBlockSize, irs)

    def process(self, dataL, dataR):                   # StereoConvolver.cpp:33-42
        """Returns (bufferLL, bufferRR[, bufferLR, bufferRL])."""
        xs = [dataL, dataR] + ([dataL, dataR] if self.isQuad else [])
        return tuple(self._e.process(xs))

    def enable_device_mixdown(self, true_stereo: bool = True):
        """feed {L, R} once and get the wet {L, R} back, mixed on the device as
        src/PluginProcessor.cpp:1833-1838 does on the host (L = LL + RL, R = RR + LR when quad & true stereo)."""
        if self.isQuad:
            ts = 1.0 if true_stereo else 0.0
            self._e.set_routing([0, 1, 0, 1], [[1, 0, 0, ts], [0, 1, ts, 0]])
        else:
            self._e.set_routing([0, 1], [[1, 0], [0, 1]])

    def process_mixed(self, dataL, dataR):
        """(wetL, wetR) after enable_device_mixdown()."""
        return tuple(self._e.process([dataL, dataR]))

    def clear(self):
        if self._e:
            self._e.clear()

    def reset(self):
        if self._e:
            self._e.reset()
