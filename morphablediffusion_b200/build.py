"""Builds libmdiff.so (hand-written sm_100a CUDA behind a C ABI) in-tree with nvcc."""
import os
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent
CSRC = ROOT / "csrc"
OUT = ROOT / "_lib"
LIB = OUT / "libmdiff.so"
# development variants: MD_BUILD_FLAGS="-DMD_KPROF" MD_BUILD_TAG=kprof builds _lib/kprof/libmdiff.so beside the product
if os.environ.get("MD_BUILD_TAG"):
    OUT = OUT / os.environ["MD_BUILD_TAG"]
    LIB = OUT / "libmdiff.so"

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-lineinfo",
    "-Xcompiler", "-fPIC,-fvisibility=hidden", "--expt-relaxed-constexpr", "-cudart", "static",
]


def sources():
    return sorted(CSRC.glob("*.cu"))


def needs_build():
    if not LIB.exists():
        return True
    t = LIB.stat().st_mtime
    deps = list(CSRC.glob("*.cu")) + list(CSRC.glob("*.cuh")) + list(CSRC.glob("*.h")) + \
        [ROOT.parent / "include" / "mdiff.h", Path(__file__)]
    return any(d.stat().st_mtime > t for d in deps)


def build(force=False, verbose=False):
    if not force and not needs_build():
        return LIB
    OUT.mkdir(parents=True, exist_ok=True)
    objs = []
    procs = []
    for src in sources():
        obj = OUT / (src.stem + ".o")
        objs.append(obj)
        deps = [src] + list(CSRC.glob("*.cuh")) + list(CSRC.glob("*.h")) + [ROOT.parent / "include" / "mdiff.h"]
        if not force and obj.exists() and all(obj.stat().st_mtime > d.stat().st_mtime for d in deps):
            continue
        cmd = ["nvcc", *NVCC_FLAGS, *os.environ.get("MD_BUILD_FLAGS", "").split(), "-c", str(src), "-o", str(obj)]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    for src, p in procs:
        out, _ = p.communicate()
        if verbose or p.returncode != 0:
            sys.stderr.write(out)
        if p.returncode != 0:
            raise RuntimeError(f"nvcc failed on {src}")
    cmd = ["nvcc", "-shared", "-o", str(LIB), *map(str, objs), "-cudart", "static", "-ldl"]
    subprocess.check_call(cmd)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
