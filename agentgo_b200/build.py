"""Build the CUDA extension in-tree: agentgo_b200/libagentgo_b200.so (sm_100a).

nvcc cross-compiles without a GPU.  The .so is git-ignored but travels with gpurun.
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
INCLUDE = os.path.join(HERE, "..", "include")
LIB = os.path.join(HERE, "libagentgo_b200.so")
GTP = os.path.join(HERE, "agentgo_gtp")
SOURCES = ["host_board.cpp", "host_layout.cpp", "engine.cu", "playout_thread.cu", "playout_warp.cu", "playout_group.cu",
           "diag.cu", "capi.cpp", "nccl_group.cpp"]
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
COMMON = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC", "-I", CSRC, "-I", INCLUDE]


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def _all_deps():
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)]
    deps.append(os.path.join(INCLUDE, "agentgo_b200.h"))
    deps.append(os.path.abspath(__file__))
    return deps


def build(force=False, verbose=False, debug=False):
    """debug=True builds libagentgo_b200_dbg.so with -DAGB_CHECK_BOUNDS (checked shared-memory
    indexing in the warp kernel); select it at run time with AGB_LIB=<path>."""
    deps = _all_deps()
    lib = LIB.replace(".so", "_dbg.so") if debug else LIB
    if force or _stale(lib, deps):
        objs = []
        objdir = os.path.join(HERE, "build", "dbg" if debug else "rel")
        os.makedirs(objdir, exist_ok=True)
        procs = []
        for src in SOURCES:
            obj = os.path.join(objdir, src + ".o")
            objs.append(obj)
            cmd = [NVCC] + ARCH + COMMON + (["-DAGB_CHECK_BOUNDS"] if debug else []) + \
                  ["-x", "cu", "-c", os.path.join(CSRC, src), "-o", obj]
            if verbose:
                cmd.insert(1, "-Xptxas=-v")
            procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
        failed = False
        for src, p in procs:
            out, _ = p.communicate()
            if p.returncode != 0:
                failed = True
                sys.stderr.write("nvcc failed on %s:\n%s\n" % (src, out))
            elif verbose and out:
                sys.stderr.write(out)
        if failed:
            raise RuntimeError("CUDA build failed")
        cmd = [NVCC] + ARCH + ["-shared", "-o", lib] + objs + ["-lcudart", "-ldl"]
        subprocess.check_call(cmd)
    if debug:
        return lib
    gtp_src = os.path.join(CSRC, "gtp_main.cpp")
    if os.path.exists(gtp_src) and (force or _stale(GTP, deps + [LIB])):
        cmd = ["g++", "-O2", "-std=c++17", "-I", CSRC, "-I", INCLUDE, gtp_src,
               os.path.join(CSRC, "gtp_adapter.cpp"), "-o", GTP,
               "-L", HERE, "-lagentgo_b200", "-Wl,-rpath,$ORIGIN", "-ldl", "-pthread"]
        subprocess.check_call(cmd)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv, debug="--debug" in sys.argv))
