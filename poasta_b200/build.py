"""Build the C-ABI shared library in-tree: ``python -m poasta_b200.build``.

nvcc cross-compiles sm_100a without a GPU. The resulting ``poasta_b200/libposta_b200.so`` is
git-ignored but travels to the GPU box with the repo snapshot.
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT = os.path.join(HERE, "libposta_b200.so")
CLI = os.path.join(HERE, "poasta")
SOURCES = ["capi_graph.cpp", "align.cu"]
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
         "-Xcompiler", "-fPIC,-Wall,-pthread", "-Xptxas", "-v"]


def _stale(out, deps):
    if not os.path.exists(out):
        return True
    t = os.path.getmtime(out)
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(HERE, "..", "include", "poasta_b200.h")]
    if force or _stale(OUT, deps):
        cmd = [NVCC] + FLAGS + ["-shared", "-o", OUT] + [os.path.join(CSRC, s) for s in SOURCES]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if verbose or r.returncode:
            sys.stderr.write(r.stdout + r.stderr)
        if r.returncode:
            raise RuntimeError("nvcc failed building libposta_b200.so")
        with open(os.path.join(HERE, "ptxas_info.txt"), "w") as f:
            f.write(r.stdout + r.stderr)
    synth_so = os.path.join(HERE, "libsynth.so")
    synth_src = os.path.join(CSRC, "synth.c")
    if force or _stale(synth_so, [synth_src]):
        r = subprocess.run(["gcc", "-O2", "-fPIC", "-shared", "-o", synth_so, synth_src], capture_output=True, text=True)
        if r.returncode:
            sys.stderr.write(r.stdout + r.stderr)
            raise RuntimeError("gcc failed building libsynth.so")
    cli_src = os.path.join(CSRC, "cli_main.cpp")
    if os.path.exists(cli_src) and (force or _stale(CLI, deps + [OUT])):
        cmd = ["g++", "-O2", "-std=c++17", "-o", CLI, cli_src, "-L" + HERE, "-lposta_b200",
               "-Wl,-rpath,$ORIGIN", "-pthread", "-lz"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode:
            sys.stderr.write(r.stdout + r.stderr)
            raise RuntimeError("g++ failed building the poasta CLI")
    return OUT


if __name__ == "__main__":
    build(force="--force" in sys.argv, verbose=True)
    print(OUT)
