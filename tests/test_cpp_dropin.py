"""Compiles tests/cpp/dropin_test.cpp against the C++ drop-in headers and runs it, linked to
the CPU emulation of the C ABI (always) and to the CUDA library (gpu)."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _build_and_run(libdir, libname, exe, src=None, expect="ALL OK"):
    src = src or os.path.join(ROOT, "tests", "cpp", "dropin_test.cpp")
    cmd = ["g++", "-O2", "-std=c++17", "-pthread", "-I", os.path.join(ROOT, "include"), src, "-o", exe,
           "-L", libdir, f"-l:{libname}", f"-Wl,-rpath,{libdir}"]
    out = subprocess.run(cmd, capture_output=True, text=True)
    assert out.returncode == 0, out.stderr
    run = subprocess.run([exe], capture_output=True, text=True, timeout=600)
    print(run.stdout, run.stderr)
    assert run.returncode == 0, run.stdout + run.stderr
    assert expect in run.stdout


def test_dropin_headers_on_emulation(tmp_path):
    from tests.emu.build_emu import build
    lib = build()
    _build_and_run(os.path.dirname(lib), os.path.basename(lib), str(tmp_path / "dropin_emu"))


@pytest.mark.gpu
def test_dropin_headers_on_gpu(tmp_path):
    from reevr_b200 import _lib
    _build_and_run(os.path.dirname(_lib.LIB_PATH), os.path.basename(_lib.LIB_PATH), str(tmp_path / "dropin_gpu"))


def test_example_offline_render_on_emulation(tmp_path):
    from tests.emu.build_emu import build
    lib = build()
    _build_and_run(os.path.dirname(lib), os.path.basename(lib), str(tmp_path / "render_emu"),
                   src=os.path.join(ROOT, "examples", "offline_render.cpp"), expect="rendered")


@pytest.mark.gpu
def test_example_offline_render_on_gpu(tmp_path):
    from reevr_b200 import _lib
    _build_and_run(os.path.dirname(_lib.LIB_PATH), os.path.basename(_lib.LIB_PATH), str(tmp_path / "render_gpu"),
                   src=os.path.join(ROOT, "examples", "offline_render.cpp"), expect="rendered")


def test_example_plugin_callback_on_emulation(tmp_path):
    from tests.emu.build_emu import build
    lib = build()
    _build_and_run(os.path.dirname(lib), os.path.basename(lib), str(tmp_path / "callback_emu"),
                   src=os.path.join(ROOT, "examples", "plugin_callback.cpp"), expect="rendered")


@pytest.mark.gpu
def test_example_plugin_callback_on_gpu(tmp_path):
    from reevr_b200 import _lib
    _build_and_run(os.path.dirname(_lib.LIB_PATH), os.path.basename(_lib.LIB_PATH), str(tmp_path / "callback_gpu"),
                   src=os.path.join(ROOT, "examples", "plugin_callback.cpp"), expect="rendered")
