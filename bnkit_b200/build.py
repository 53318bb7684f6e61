"""Build libbnkgpu.so in-tree with nvcc for sm_100a (no JIT cache: the .so travels with the repo)."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT = os.path.join(HERE, "libbnkgpu.so")
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
COMMON = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC", "-Xcompiler", "-ffp-contract=off"]
# tables.cu must not contract a*b+c into an FMA (bit-exact P(t) / log P(t)); the sweeps may.
UNITS = [
    ("tables.cu", ["-fmad=false"]),
    ("sweeps.cu", []),
    ("sweeps_fast.cu", []),
    ("sweeps_warp.cu", []),
    ("wide.cu", []),
    ("wide_general.cu", []),
    ("cherry.cu", []),
    ("traceback.cu", []),
    ("indel_tables.cu", ["-fmad=false"]),
    ("indel.cu", []),
    ("bep.cu", []),
    ("bep_reach.cu", []),
    ("bep_api.cu", []),
    ("api.cu", []),
    ("multi.cu", []),
    ("subst_model.cpp", ["-fmad=false"]),
    ("tree_plan.cpp", []),
]


def _nvcc():
    for c in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if c and (os.path.isabs(c) and os.path.exists(c) or not os.path.isabs(c)):
            return c
    raise RuntimeError("nvcc not found")


def sources():
    return [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(HERE, "..", "include", "bnkgpu.h")]


def needs_build():
    if not os.path.exists(OUT):
        return True
    t = os.path.getmtime(OUT)
    return any(os.path.getmtime(s) > t for s in sources())


def build(force=False, verbose=False):
    if not force and not needs_build():
        return OUT
    nvcc = _nvcc()
    objdir = os.path.join(HERE, "build")
    os.makedirs(objdir, exist_ok=True)
    objs = []
    procs = []
    for src, extra in UNITS:
        obj = os.path.join(objdir, src + ".o")
        cmd = [nvcc] + ARCH + COMMON + extra + ["-c", os.path.join(CSRC, src), "-o", obj]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
            print(" ".join(cmd), flush=True)
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
        objs.append(obj)
    failed = False
    for src, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0:
            failed = True
            sys.stderr.write("nvcc failed for %s:\n%s\n" % (src, out))
        elif verbose or "warning" in out:
            sys.stderr.write(out)
    if failed:
        raise RuntimeError("libbnkgpu build failed")
    cmd = [nvcc] + ARCH + ["-shared", "-o", OUT] + objs + ["-lcudart", "-ldl"]
    subprocess.run(cmd, check=True)
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
