"""build.py — compile libpheniqs_b200.so in-tree with nvcc for sm_100a (cross-compiles without a GPU)."""
from __future__ import annotations

import os
import shutil
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
SOURCES = ("api.cu", "kernels.cu", "pack.cu")
HEADERS = ("kernels.cuh", "pack.cuh", "spec.hpp", "json.hpp", "report.hpp", os.path.join("..", "..", "include", "pheniqs_b200.h"))
OUTPUT = os.path.join(HERE, "libpheniqs_b200.so")


def nvcc_command(verbose: bool = False):
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    host = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
    command = [nvcc, "-ccbin", host, "-std=c++17", "-O3", "-lineinfo",
               "-gencode", "arch=compute_100a,code=sm_100a",
               "-Xcompiler", "-fPIC", "-shared"]
    if verbose:
        command += ["-Xptxas", "-v"]
    command += [os.path.join(HERE, "csrc", s) for s in SOURCES] + ["-o", OUTPUT]
    return command


def stale() -> bool:
    if not os.path.exists(OUTPUT):
        return True
    built = os.path.getmtime(OUTPUT)
    return any(os.path.getmtime(os.path.join(HERE, "csrc", f)) > built for f in SOURCES + HEADERS)


def build(force: bool = False, verbose: bool = False) -> str:
    if force or stale():
        subprocess.run(nvcc_command(verbose), check=True)
    return OUTPUT


if __name__ == "__main__":
    print(build(force=True, verbose=True))
