"""Build recipe for libbevgen_b200.so (sm_100a only, in-tree so the .so travels with gpurun snapshots)."""
import os
import subprocess
import sys
from pathlib import Path

HERE = Path(__file__).resolve().parent
CSRC = HERE / "csrc"
LIB = HERE / "libbevgen_b200.so"
SOURCES = ["c_api.cu", "gemm_tc.cu", "elementwise.cu", "vq.cu", "transformer.cu", "decode.cu", "decode_persistent.cu", "attn_fused.cu", "conv_halo.cu", "conv_fused.cu", "conv_fused2.cu", "conv_fused3.cu", "conv_small.cu", "gemm_pair.cu", "maskgit.cu"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "--use_fast_math=false",
              "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden"]


def _nvcc():
    for c in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if c and (os.path.isabs(c) and os.path.exists(c) or not os.path.isabs(c)):
            return c
    return "nvcc"


def needs_build() -> bool:
    if not LIB.exists():
        return True
    t = LIB.stat().st_mtime
    deps = list(CSRC.glob("*.cu")) + list(CSRC.glob("*.cuh")) + [HERE.parent / "include" / "bevgen_b200.h"]
    return any(d.stat().st_mtime > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> Path:
    if not force and not needs_build():
        return LIB
    objs = []
    flags = [f for f in NVCC_FLAGS if not f.startswith("--use_fast_math")]
    procs = []
    (HERE / "build").mkdir(exist_ok=True)
    for s in SOURCES:
        o = HERE / "build" / (s + ".o")
        cmd = [_nvcc(), *flags, "-c", str(CSRC / s), "-o", str(o)]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
        procs.append((s, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
        objs.append(str(o))
    for s, p in procs:
        out, _ = p.communicate()
        if verbose or p.returncode:
            print(out, file=sys.stderr)
        if p.returncode:
            raise RuntimeError(f"nvcc failed on {s}")
    cmd = [_nvcc(), "-shared", "-o", str(LIB), *objs, "-Xcompiler", "-fPIC", "-cudart", "static"]
    subprocess.check_call(cmd)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
