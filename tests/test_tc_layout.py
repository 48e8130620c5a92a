"""Index arithmetic of the tensor-core sweep (reevr_b200/csrc/kernels_tc.cuh) on the CPU: the geometry, the pre-swizzled
time-line layout and the Toeplitz tile images are rebuilt here with the header's own inline functions (compiled by g++
through tests/cpp/tc_layout_shim.cpp) and pushed through a float64 model of what k_tc_sweep / k_tc_merge_y do with them
— tile n0, K chunk c -> strip plane c % 2, rows n0 + n + c / 2; A image row m = part * 64 + i, column jj -> H[i + Q - 32 c - jj]
— and the result must be the sweep FFTConvolver.cpp:176-187 defines, y[t] = sum_p H[p] x[t - p], packed DC / Nyquist entry
included.  The kernels themselves (TMA, tcgen05, tensor memory) are covered on the GPU by tests/test_tc_sweep.py."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def shim(tmp_path_factory):
    so = str(tmp_path_factory.mktemp("tc") / "libtc_layout.so")
    cuda_inc = os.path.join(os.environ.get("CUDA_HOME", "/usr/local/cuda"), "include")
    cmd = ["g++", "-O1", "-std=c++17", "-shared", "-fPIC", "-I", cuda_inc, os.path.join(ROOT, "tests", "cpp", "tc_layout_shim.cpp"), "-o", so]
    out = subprocess.run(cmd, capture_output=True, text=True)
    assert out.returncode == 0, out.stderr
    lib = C.CDLL(so)
    lib.tc_xt_index.restype = C.c_ulonglong
    lib.tc_xt_index.argtypes = [C.c_longlong, C.c_int, C.c_int, C.c_longlong, C.c_int, C.c_int]
    lib.tc_sw128.restype = C.c_uint
    lib.tc_sw128.argtypes = [C.c_uint, C.c_uint]
    return lib


def geom(lib, P, nb):
    out = (C.c_int * 5)()
    lib.tc_geom(P, nb, out)
    return dict(zip(("Q", "nchunk", "nseg", "ntile", "rows"), out))


def consts(lib):
    out = (C.c_int * 8)()
    lib.tc_consts(out)
    return dict(zip(("R", "N", "strip_rows", "strip_bytes", "a_tile_bytes", "max_chunks", "flush", "smem"), out))


def test_geometry_limits(shim):
    k = consts(shim)
    assert (k["R"], k["N"], k["strip_rows"], k["strip_bytes"], k["a_tile_bytes"]) == (64, 128, 144, 144 * 128, 16384)
    assert 8 * k["strip_bytes"] + 4 * k["a_tile_bytes"] + 1024 == k["smem"] <= 227 * 1024
    assert k["strip_bytes"] % 1024 == 0                       # strips stay aligned to the 1024-byte swizzle atom
    g = geom(shim, 938, 112608)                               # the metric shape
    assert g == dict(Q=960, nchunk=32, nseg=1760, ntile=14, rows=14 * 128 + 16)
    assert (g["nchunk"] - 1) // 2 + k["N"] <= k["strip_rows"]  # the largest row shift stays inside the strip
    assert shim.tc_geom_ok(961, 5000, 512) == 1 and shim.tc_geom_ok(962, 5000, 512) == 0
    assert shim.tc_geom_ok(100, 5000, 48) == 0 and shim.tc_geom_ok(1, 1, 32) == 1
    assert geom(shim, 1, 1) == dict(Q=0, nchunk=2, nseg=1, ntile=1, rows=144)


def test_swizzle_is_a_permutation_of_each_row(shim):
    for r in range(16):
        offs = sorted(shim.tc_sw128(r, e) for e in range(32))
        assert offs == [r * 128 + 4 * e for e in range(32)]
    # the time-line image of a row equals the shared-memory image of the same row index modulo 8
    rows = 144
    for R in (0, 5, 8, 131):
        base = shim.tc_xt_index(0, 0, 0, R, 0, rows) & ~31
        for jj in range(32):
            assert (shim.tc_xt_index(0, 0, 0, R, jj, rows) - base) * 4 == shim.tc_sw128(R % 8, jj) - (R % 8) * 128


@pytest.mark.parametrize("P,nb", [(100, 300), (938, 200), (1, 70), (65, 8300)])
def test_float64_model_of_the_tiled_sweep(shim, P, nb):
    rng = np.random.default_rng(P * 1000 + nb)
    k = consts(shim)
    g = geom(shim, P, nb)
    Q, rows = g["Q"], g["rows"]
    lines = 2                                                  # line 0 is the packed DC / Nyquist entry
    H = rng.standard_normal((lines, P)) + 1j * rng.standard_normal((lines, P))
    xrow0 = Q + 2
    x = rng.standard_normal((lines, xrow0 + nb)) + 1j * rng.standard_normal((lines, xrow0 + nb))
    x[:, :xrow0 - (P - 1)] = np.nan                            # rows the FFMA sweep never reads: must not leak in
    # k_tc_split_x: tau = row - (xrow0 - Q); rows outside [xrow0 - (P - 1), xrow0 + nb) read as zero
    Xt = np.full(lines * 4 * 2 * rows * 32, np.nan)
    for line in range(lines):
        for tau in range(rows * 64):
            row = xrow0 - Q + tau
            v = x[line, row] if xrow0 - (P - 1) <= row < xrow0 + nb else 0.0
            R, e, jj = tau >> 6, (tau >> 5) & 1, tau & 31
            Xt[shim.tc_xt_index(line, 0, e, R, jj, rows)] = v.real if v else 0.0     # hi planes carry the value,
            Xt[shim.tc_xt_index(line, 1, e, R, jj, rows)] = 0.0                       # lo planes zero in this model
            Xt[shim.tc_xt_index(line, 2, e, R, jj, rows)] = v.imag if v else 0.0
            Xt[shim.tc_xt_index(line, 3, e, R, jj, rows)] = 0.0
    assert not np.isnan(Xt).any()
    unsw = np.array([[shim.tc_sw128(r, e) // 4 for e in range(32)] for r in range(8)])   # float offset inside an 8-row atom

    def strip_rows(line, pl, e, r0, count):                   # rows r0 .. r0+count of one plane, un-swizzled
        out = np.empty((count, 32))
        for i in range(count):
            R = r0 + i
            base = shim.tc_xt_index(line, pl, e, R, 0, rows) & ~31
            out[i] = Xt[base + unsw[R % 8] - (R % 8) * 32]
        return out

    y = np.zeros((lines, nb), complex)
    for line in range(lines):
        for nt in range(g["ntile"]):
            D = np.zeros((2, 128, k["N"]))                     # [D | D2]
            for c in range(g["nchunk"]):
                A = np.zeros((128, 32))                        # k_tc_build_a (hi image, un-swizzled view)
                for m in range(128):
                    part, i = m >> 6, m & 63
                    for jj in range(32):
                        p = i + Q - (32 * c + jj)
                        if 0 <= p < P:
                            A[m, jj] = H[line, p].imag if part else H[line, p].real
                e, q = c & 1, c >> 1
                for comp in range(2):
                    Bm = strip_rows(line, comp * 2, e, nt * k["N"] + q, k["N"])       # [n][jj]
                    D[comp] += A @ Bm.T
            for n in range(k["N"]):                            # k_tc_merge_y
                for i in range(64):
                    t = 64 * (nt * k["N"] + n) + i
                    if t < nb:
                        d0, d1, e0, e1 = D[0, i, n], D[0, 64 + i, n], D[1, i, n], D[1, 64 + i, n]
                        y[line, t] = complex(d0, e1) if line == 0 else complex(d0 - e1, d1 + e0)
    for line in range(lines):
        for t in list(range(min(nb, 70))) + [nb - 1, nb // 2]:
            hs, xs = H[line], x[line, xrow0 + t - np.arange(P)]
            ref = complex(np.sum(hs.real * xs.real), np.sum(hs.imag * xs.imag)) if line == 0 else np.sum(hs * xs)
            assert abs(y[line, t] - ref) <= 1e-9 * max(1.0, abs(ref)), (line, t)


def test_bench_flop_model_matches_the_kernel_geometry(shim):
    """bench.py's tensor roofline counts tiles x K chunks x 24 MMAs x 2*128*128*8 executed flops per launch; the tile and
    chunk counts it derives from (P, blocks per launch) must be the kernel's own (make_geom)."""
    k = consts(shim)
    for P, nb, C, B in [(938, 112608, 2, 512), (938, 14077, 2, 512), (100, 4608, 2, 64), (961, 8192, 4, 256)]:
        g = geom(shim, P, nb)
        q = (max(P - 1, 0) + 63) // 64 * 64                                # bench.py: sweep_roofline()
        nchunk = q // 32 + 2
        ntile = -(-(-(-nb // 64)) // 128)
        assert (nchunk, ntile) == (g["nchunk"], g["ntile"])
        per_stage_mmas = 16 + 8                                              # hi image: 4 k-steps x 2 time lines x 2 products, lo image: x 1
        flop = C * B * ntile * nchunk * per_stage_mmas * 2.0 * 128 * k["N"] * 8
        if (P, nb, C, B) == (938, 112608, 2, 512):
            assert flop == 2886218022912.0                                   # the figure in profiles/r02_bench_n1.json
