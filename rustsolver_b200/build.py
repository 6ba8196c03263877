"""In-tree build of the engine library and (separately) the test oracle.

`build_engine()` compiles rustsolver_b200/csrc/* into rustsolver_b200/libb200cfr.so for sm_100a
with nvcc (cross-compiles without a GPU).  `build_oracle()` compiles oracle/cfr_oracle.c into
oracle/liborc.so with gcc; the oracle is test infrastructure and is never linked into the engine.
"""
from __future__ import annotations

import os
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
CSRC = ROOT / "rustsolver_b200" / "csrc"
ENGINE_SO = ROOT / "rustsolver_b200" / "libb200cfr.so"
ORACLE_SO = ROOT / "oracle" / "liborc.so"

ENGINE_SOURCES = ["kernels.cu", "street_kernel.cu", "indexer_kernel.cu", "abstraction_kernels.cu", "histogram_kernel.cu", "engine.cu", "plan.cpp", "poker.cpp", "game.cpp", "hand_indexer.cpp",
                  "trainer.cpp", "host_api.cpp"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC,-Wall,-O3"]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if cand and (os.path.sep not in cand or os.path.exists(cand)):
            return cand
    return "nvcc"


def _stale(target: Path, deps) -> bool:
    if not target.exists():
        return True
    t = target.stat().st_mtime
    return any(Path(d).stat().st_mtime > t for d in deps)


def build_engine(force: bool = False, verbose: bool = False) -> Path:
    deps = [CSRC / s for s in ENGINE_SOURCES] + list(CSRC.glob("*.h")) + list(CSRC.glob("*.cuh")) + \
        list((ROOT / "include").glob("*.h"))
    if not force and not _stale(ENGINE_SO, deps):
        return ENGINE_SO
    objdir = ROOT / "build" / "engine"
    objdir.mkdir(parents=True, exist_ok=True)
    objs = []
    procs = []
    for s in ENGINE_SOURCES:
        src = CSRC / s
        obj = objdir / (s + ".o")
        objs.append(str(obj))
        if not force and not _stale(obj, [src] + [d for d in deps if d.suffix in (".h", ".cuh")]):
            continue
        cmd = [_nvcc(), *NVCC_FLAGS, "-x", "cu" if s.endswith(".cu") else "c++", "-c", str(src), "-o", str(obj)]
        if s.endswith(".cu"):
            cmd.insert(1, "-Xptxas=-v" if verbose else "-Xptxas=-O3")
        procs.append((s, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    for s, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0:
            raise RuntimeError(f"nvcc failed on {s}:\n{out}")
        if verbose and out.strip():
            print(out)
    link = [_nvcc(), "-shared", "-o", str(ENGINE_SO), *objs, "-gencode", "arch=compute_100a,code=sm_100a",
            "-cudart", "static", "-ldl", "-lpthread"]
    r = subprocess.run(link, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stdout}")
    return ENGINE_SO


def build_oracle(force: bool = False) -> Path:
    src = ROOT / "oracle" / "cfr_oracle.c"
    if not force and not _stale(ORACLE_SO, [src, ROOT / "oracle" / "abstraction_oracle.c"]):
        return ORACLE_SO
    # -ffp-contract=off: no FMA contraction, the fp32 restatements must keep the reference's operation order
    cmd = ["gcc", "-O3", "-march=native", "-std=gnu11", "-ffp-contract=off", "-fopenmp", "-shared", "-fPIC", "-Wall",
           str(src), str(ROOT / "oracle" / "abstraction_oracle.c"), "-o", str(ORACLE_SO), "-lm"]
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"gcc failed on the oracle:\n{r.stdout}")
    return ORACLE_SO


if __name__ == "__main__":
    force = "--force" in sys.argv
    print(build_engine(force=force, verbose="-v" in sys.argv))
    print(build_oracle(force=force))
