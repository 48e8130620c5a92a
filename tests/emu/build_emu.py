"""TEST INFRASTRUCTURE: compiles engine.cu + kernels.cuh with g++ against tests/emu/cuda_emu.h."""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
CSRC = os.path.join(ROOT, "reevr_b200", "csrc")
LIB = os.path.join(HERE, "libb200conv_emu.so")
DEPS = [os.path.join(CSRC, "engine.cu"), os.path.join(CSRC, "irshape.cu"), os.path.join(CSRC, "kernels.cuh"),
        os.path.join(CSRC, "kernels_stream.cuh"), os.path.join(CSRC, "kernels_fft512.cuh"), os.path.join(CSRC, "kernels_rt.cuh"),
        os.path.join(CSRC, "kernels_chain.cuh"),
        os.path.join(HERE, "cuda_emu.h"), os.path.join(ROOT, "include", "b200conv.h")]


def build(force=False):
    if not force and os.path.exists(LIB) and all(os.path.getmtime(d) <= os.path.getmtime(LIB) for d in DEPS):
        return LIB
    cmd = ["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-DPC_EMULATE", "-I", HERE, "-x", "c++",
           os.path.join(CSRC, "engine.cu"), os.path.join(CSRC, "irshape.cu"), "-o", LIB]
    out = subprocess.run(cmd, capture_output=True, text=True)
    if out.returncode != 0:
        raise RuntimeError("emu build failed:\n" + out.stdout + out.stderr)
    return LIB


if __name__ == "__main__":
    print(build(force=True))
