"""Build libqt_b200.so (hand-written sm_100a kernels + C ABI) in-tree with nvcc.

    python quantized-training_b200/build.py [--force]

The .so lands next to the Python package (quantized_training/_lib/) so that it travels
with the repo snapshot to the GPU box; it is git-ignored.  nvcc cross-compiles without a GPU.
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT_DIR = os.path.join(HERE, "quantized_training", "_lib")
LIB = os.path.join(OUT_DIR, "libqt_b200.so")
SOURCES = ["qt_format.cc", "qt_lut.cc", "qt_fq.cu", "qt_codes.cu", "qt_block.cu", "qt_block_flat.cu", "qt_block_cols.cu", "qt_block_tile.cu", "qt_gemm.cu", "qt_fused.cu"]
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC", "-shared", "-ccbin", "/usr/bin/g++",
]


def _stale():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(HERE, "..", "include", "qt_b200.h")]
    return any(os.path.getmtime(d) > t for d in deps)


def build_extension(force=False, verbose=False):
    """Compile every source to an object file (in parallel, only the stale ones) and link the shared library."""
    if not force and not _stale():
        return LIB
    from concurrent.futures import ThreadPoolExecutor

    os.makedirs(OUT_DIR, exist_ok=True)
    obj_dir = os.path.join(HERE, "build")
    os.makedirs(obj_dir, exist_ok=True)
    headers = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".h", ".cuh"))] + \
              [os.path.join(HERE, "..", "include", "qt_b200.h")]
    newest_header = max(os.path.getmtime(h) for h in headers)
    compile_flags = [f for f in FLAGS if f != "-shared"]

    def compile_one(src):
        path = os.path.join(CSRC, src)
        obj = os.path.join(obj_dir, src + ".o")
        if not force and os.path.exists(obj) and os.path.getmtime(obj) > max(os.path.getmtime(path), newest_header):
            return obj, None
        cmd = [NVCC] + compile_flags + (["-Xptxas", "-v"] if verbose else []) + ["-c", "-o", obj, path]
        return obj, subprocess.run(cmd, capture_output=True, text=True)

    with ThreadPoolExecutor(max_workers=len(SOURCES)) as pool:
        results = list(pool.map(compile_one, SOURCES))
    for obj, r in results:
        if r is not None and r.returncode != 0:
            sys.stderr.write(r.stdout + r.stderr)
            raise RuntimeError("nvcc failed building libqt_b200.so")
        if r is not None and verbose:
            sys.stderr.write(r.stderr)
    r = subprocess.run([NVCC, "-shared", "-ccbin", "/usr/bin/g++", "-o", LIB] + [obj for obj, _ in results],
                       capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
        raise RuntimeError("nvcc failed linking libqt_b200.so")
    return LIB


if __name__ == "__main__":
    print(build_extension(force="--force" in sys.argv, verbose="-v" in sys.argv))
