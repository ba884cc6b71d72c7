"""Build libdml_b200.so in-tree with nvcc for sm_100a (no JIT cache: the .so travels with the repo).

    python disentangled-multimodal-learning_b200/build.py [--force]
"""
from __future__ import annotations

import concurrent.futures as cf
import hashlib
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIBDIR = os.path.join(HERE, "lib")
LIB = os.path.join(LIBDIR, "libdml_b200.so")
# test-only library (never loaded by the product path): the legacy mma.sync cross-check kernels of csrc/test_only/ plus a
# second build of the tcgen05 backward with its debug knobs compiled in (-DDML_TEST_KNOBS)
TEST_LIB = os.path.join(LIBDIR, "libdml_b200_test.so")
TEST_KNOB_SOURCES = ["deform_attn_tc_bwd.cu"]
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC",
         "--expt-relaxed-constexpr", "-Xptxas", "-v"]


def sources():
    return sorted(f for f in os.listdir(CSRC) if f.endswith(".cu"))


def test_sources():
    d = os.path.join(CSRC, "test_only")
    return sorted(os.path.join("test_only", f) for f in os.listdir(d) if f.endswith(".cu"))


def _digest():
    h = hashlib.sha256()
    for root in (CSRC, os.path.join(CSRC, "test_only"), os.path.join(os.path.dirname(HERE), "include")):
        for f in sorted(os.listdir(root)):
            if f.endswith((".cu", ".cuh", ".h")):
                h.update(f.encode())
                with open(os.path.join(root, f), "rb") as fh:
                    h.update(fh.read())
    h.update(" ".join(FLAGS).encode())
    return h.hexdigest()


def _compile(job):
    src, extra, tag = job
    stem = os.path.basename(src)[:-3] + tag
    obj = os.path.join(LIBDIR, stem + ".o")
    cmd = [NVCC, *FLAGS, *extra, "-c", os.path.join(CSRC, src), "-o", obj]
    r = subprocess.run(cmd, capture_output=True, text=True)
    with open(os.path.join(LIBDIR, stem + ".ptxas.log"), "w") as f:
        f.write(r.stderr)
    if r.returncode != 0:
        raise RuntimeError(f"nvcc failed on {src}:\n{r.stdout}\n{r.stderr}")
    return obj


def build(force: bool = False, verbose: bool = True) -> str:
    os.makedirs(LIBDIR, exist_ok=True)
    stamp = os.path.join(LIBDIR, "build.sha256")
    dig = _digest()
    if (not force and os.path.exists(LIB) and os.path.exists(TEST_LIB) and os.path.exists(stamp)
            and open(stamp).read().strip() == dig):
        return LIB
    prod = [(s_, [], "") for s_ in sources()]
    knobs = ["-DDML_TEST_KNOBS"] + (["-DDML_TRACE"] if os.environ.get("DML_B200_TRACE") else [])   # clock64() stamps for scripts/trace_dkv.py
    test = [(s_, [], "") for s_ in test_sources()] + [(s_, knobs, "_knobs") for s_ in TEST_KNOB_SOURCES]
    with cf.ThreadPoolExecutor(max_workers=os.cpu_count() or 4) as ex:
        objs_all = list(ex.map(_compile, prod + test))
    objs, test_objs = objs_all[:len(prod)], objs_all[len(prod):]
    for lib, oo in ((LIB, objs), (TEST_LIB, test_objs)):
        cmd = [NVCC, "-shared", "-o", lib, *oo, "-gencode", "arch=compute_100a,code=sm_100a", "-lcudart"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    with open(stamp, "w") as f:
        f.write(dig)
    if verbose:
        print(f"built {LIB} from {len(objs)} translation units")
    return LIB


if __name__ == "__main__":
    build(force="--force" in sys.argv)
