"""Build libb200_train.so in-tree with nvcc for sm_100a (cross-compiles without a GPU)."""
import os
import subprocess
import sys
import hashlib

HERE = os.path.dirname(os.path.abspath(__file__))
SOURCES = ["gemm_tcgen05.cu", "gemm_tcgen05_2cta.cu", "elementwise.cu", "attention.cu", "attention_tc.cu", "attention_fwd64.cu", "attention_fwd_ts.cu", "comm.cu", "c_api.cu"]
HEADERS = ["ptx.cuh", "common.h", "../../include/b200_train.h"]
LIB = os.path.join(HERE, "libb200_train.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-Xcompiler", "-fPIC",
         "--use_fast_math=false"] if False else ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3",
                                                   "-std=c++17", "-Xcompiler", "-fPIC"]


def _stamp():
    h = hashlib.sha256()
    for f in SOURCES + HEADERS + ["build.py"]:
        with open(os.path.join(HERE, f), "rb") as fh:
            h.update(fh.read())
    return h.hexdigest()


def _cublas_dirs():
    dirs = ["/usr/local/cuda/lib64"]
    try:
        import nvidia.cublas  # torch's bundled copy (already resident in a torch process)
        dirs.insert(0, os.path.join(os.path.dirname(nvidia.cublas.__file__), "lib"))
    except Exception:
        pass
    return dirs


def build(force=False, verbose=False):
    stamp_file = LIB + ".stamp"
    stamp = _stamp()
    if not force and os.path.exists(LIB) and os.path.exists(stamp_file) and open(stamp_file).read() == stamp:
        return LIB
    objs = []
    procs = []
    for src in SOURCES:
        obj = os.path.join(HERE, src.replace(".cu", ".o"))
        cmd = [NVCC] + FLAGS + ["-c", os.path.join(HERE, src), "-o", obj]
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
        objs.append(obj)
    for src, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0:
            raise RuntimeError(f"nvcc failed on {src}:\n{out}")
        if verbose and out.strip():
            print(out)
    link = [NVCC, "-shared", "-o", LIB] + objs + ["-L/usr/local/cuda/lib64", "-lcublasLt"]
    for d in _cublas_dirs():
        link += ["-Xlinker", "-rpath", "-Xlinker", d]
    r = subprocess.run(link, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stdout}")
    with open(stamp_file, "w") as f:
        f.write(stamp)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
