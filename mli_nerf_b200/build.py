"""Build recipe for libmli_b200.so (hand-written CUDA for sm_100a behind the C ABI of include/mli_b200.h).

nvcc cross-compiles without a GPU; the .so is written in-tree (mli_nerf_b200/libmli_b200.so) so that it travels
with the repo snapshot to the GPU box.  No torch dependency: the library links only against cudart.
"""
import hashlib
import os
import shutil
import subprocess
import sys

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
SRC_DIR = os.path.join(PKG_DIR, "csrc")
OBJ_DIR = os.path.join(PKG_DIR, "build")
LIB_PATH = os.path.join(PKG_DIR, "libmli_b200.so")
SOURCES = ["capi.cu", "hashgrid.cu", "gemm_simt.cu", "gemm_tcgen05.cu", "heads_fused.cu", "sdf_trunk.cu", "tcl_ops.cu", "rowdot.cu", "rays_sampling.cu",
           "geometry.cu", "composite.cu", "losses.cu", "visibility.cu", "peer.cu"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-Xcompiler", "-fPIC",
              "--expt-relaxed-constexpr", "-Xcudafe", "--diag_suppress=177"]


def _nvcc():
    exe = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(exe):
        raise RuntimeError("nvcc not found: libmli_b200.so cannot be built (there is no CPU fallback)")
    return exe


def _stamp(paths):
    h = hashlib.sha256()
    for p in sorted(paths):
        h.update(p.encode())
        h.update(open(p, "rb").read())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def build(force=False, verbose=False):
    """Compile every .cu for sm_100a and link libmli_b200.so.  Incremental per object file."""
    os.makedirs(OBJ_DIR, exist_ok=True)
    headers = [os.path.join(SRC_DIR, f) for f in os.listdir(SRC_DIR) if f.endswith((".h", ".cuh"))]
    headers.append(os.path.join(PKG_DIR, "..", "include", "mli_b200.h"))
    nvcc = _nvcc()
    objs, relink = [], force or not os.path.exists(LIB_PATH)
    procs = []
    for src in SOURCES:
        path = os.path.join(SRC_DIR, src)
        obj = os.path.join(OBJ_DIR, src.replace(".cu", ".o"))
        stamp_file = obj + ".stamp"
        stamp = _stamp([path] + headers)
        objs.append(obj)
        if not force and os.path.exists(obj) and os.path.exists(stamp_file) and open(stamp_file).read() == stamp:
            continue
        cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-c", path, "-o", obj]
        procs.append((src, stamp_file, stamp, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT)))
        relink = True
    for src, stamp_file, stamp, proc in procs:
        out, _ = proc.communicate()
        if proc.returncode != 0:
            sys.stderr.write(out.decode())
            raise RuntimeError(f"nvcc failed on {src}")
        if verbose:
            sys.stderr.write(out.decode())
        open(stamp_file, "w").write(stamp)
    if relink:
        cmd = [nvcc, "-Wno-deprecated-gpu-targets", "-shared", "-o", LIB_PATH] + objs + ["-lcudart"]
        subprocess.check_call(cmd)
    return LIB_PATH


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
