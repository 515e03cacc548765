"""Debug build of the library with the phase-trace stamps compiled in (-DPMV_ATTN_TRACE) for scripts/attn_trace*.py:
writes scripts/bin/libpmv_b200_trace.so (git-ignored).  The product library (pmv_b200/libpmv_b200.so) is never built with it."""
import concurrent.futures as cf, glob, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "portrait-mode-video_b200", "csrc")
OUT = os.path.join(ROOT, "scripts", "bin")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC",
         "--expt-relaxed-constexpr", "-DPMV_ATTN_TRACE"]


def one(src):
    obj = os.path.join(OUT, "obj", os.path.basename(src)[:-3] + ".o")
    subprocess.run(["nvcc", *FLAGS, "-c", src, "-o", obj], check=True, capture_output=True)
    return obj


os.makedirs(os.path.join(OUT, "obj"), exist_ok=True)
with cf.ThreadPoolExecutor(8) as ex:
    objs = list(ex.map(one, sorted(glob.glob(os.path.join(CSRC, "*.cu")))))
lib = os.path.join(OUT, "libpmv_b200_trace.so")
subprocess.run(["nvcc", "-shared", "-o", lib, *objs, "-gencode", "arch=compute_100a,code=sm_100a", "-lcudart"], check=True)
print(lib)
