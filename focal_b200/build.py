"""Build libfocal_b200.so (the product: C ABI of include/focal_b200.h) and libfocal_bringup.so (hardware probes and
micro-benchmarks used by tests/test_gpu_probe.py and tools/) in-tree with nvcc for sm_100a -- no torch headers.

    python -m focal_b200.build [--force]
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG_DIR, "csrc")
LIB_PATH = os.path.join(PKG_DIR, "libfocal_b200.so")
BRINGUP_PATH = os.path.join(PKG_DIR, "libfocal_bringup.so")
SOURCES = [os.path.join(CSRC, "focal_b200.cu")]
BRINGUP_SOURCES = [os.path.join(CSRC, "bringup.cu")]
HEADERS = sorted(os.path.join(CSRC, n) for n in os.listdir(CSRC) if n.endswith((".cuh", ".h"))) + [
    os.path.join(os.path.dirname(PKG_DIR), "include", "focal_b200.h")]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "-Xptxas=-v",
    "-Xcompiler", "-fPIC", "-shared",
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: cannot build libfocal_b200.so")


def _stale(out: str, sources) -> bool:
    if not os.path.exists(out):
        return True
    t = os.path.getmtime(out)
    return any(os.path.getmtime(s) > t for s in list(sources) + HEADERS)


def needs_build() -> bool:
    return _stale(LIB_PATH, SOURCES) or _stale(BRINGUP_PATH, BRINGUP_SOURCES)


def build_variant(tag: str, defines: dict) -> str:
    """Experiment builds (tools/variant_bench.py): libfocal_b200_<tag>.so with extra -D knobs."""
    out = os.path.join(PKG_DIR, f"libfocal_b200_{tag}.so")
    cmd = [_nvcc()] + NVCC_FLAGS + [f"-D{k}={v}" for k, v in defines.items()] + ["-o", out] + SOURCES
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
        raise RuntimeError(f"nvcc failed building variant {tag}")
    return out


def _compile(out: str, sources, log_name: str, verbose: bool) -> None:
    cmd = [_nvcc()] + NVCC_FLAGS + ["-o", out] + list(sources)
    res = subprocess.run(cmd, capture_output=True, text=True)
    log = res.stdout + res.stderr
    with open(os.path.join(PKG_DIR, log_name), "w") as fh:
        fh.write(" ".join(cmd) + "\n" + log)
    if res.returncode != 0:
        sys.stderr.write(log)
        raise RuntimeError(f"nvcc failed building {os.path.basename(out)}")
    if verbose:
        print(log)


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile whatever is out of date; the two libraries build concurrently."""
    from concurrent.futures import ThreadPoolExecutor
    jobs = []
    if force or _stale(LIB_PATH, SOURCES):
        jobs.append((LIB_PATH, SOURCES, "build.log"))
    if force or _stale(BRINGUP_PATH, BRINGUP_SOURCES):
        jobs.append((BRINGUP_PATH, BRINGUP_SOURCES, "build_bringup.log"))
    if jobs:
        with ThreadPoolExecutor(max_workers=len(jobs)) as ex:
            for f in [ex.submit(_compile, out, src, log, verbose) for out, src, log in jobs]:
                f.result()
    return LIB_PATH


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
