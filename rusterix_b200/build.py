"""In-tree builds: the CUDA product library (nvcc, sm_100a only) and the CPU oracle (g++).
`python -m rusterix_b200.build` builds both; __graft_entry__.build() calls build_all()."""
import os
import shutil
import subprocess
import sys

_HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(_HERE)
CSRC = os.path.join(_HERE, "csrc")
# experiments: RX_BUILD_TAG=tma RX_NVCC_EXTRA=-DRX_TMA_STORE=1 builds librxcuda_tma.so next to the shipped library
# (load it with RXC_LIB=<path>, rusterix_b200/_lib.py)
_TAG = os.environ.get("RX_BUILD_TAG", "")
LIB = os.path.join(_HERE, "librxcuda%s.so" % (("_" + _TAG) if _TAG else ""))

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-std=c++17", "-lineinfo",
    # Rust never contracts a*b+c (SURVEY T-fma): contraction is off for the whole library and
    # FMAs are written explicitly where the reference has mul_add.  Denormals, division and
    # square roots stay IEEE (-ftz=false -prec-div=true -prec-sqrt=true are the defaults).
    "-fmad=false",
    "-Xcompiler", "-fPIC,-O2,-Wall",
    "-shared",
    "-cudart", "shared",
    # libcudart.so.12: the toolkit copy or the one the venv's torch loads
    "-Xlinker", "-rpath=/usr/local/cuda/lib64",
    "-Xlinker", "-rpath=/opt/prime-rl/.venv/lib/python3.12/site-packages/nvidia/cuda_runtime/lib",
]


def _newer(target, sources):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(s) > t for s in sources)


def cuda_sources():
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cu"))


def build_cuda(force=False, verbose=False):
    """One object per .cu (compiled in parallel, only when the source or a header is newer), then one link."""
    from concurrent.futures import ThreadPoolExecutor

    srcs = cuda_sources()
    headers = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    headers.append(os.path.join(ROOT, "include", "rxcuda.h"))
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    extra = os.environ.get("RX_NVCC_EXTRA", "").split()  # experiments only, e.g. -DRX_RASTER_MIN_BLOCKS=4
    objdir = os.path.join(CSRC, "_obj" + (("_" + _TAG) if _TAG else ""))
    os.makedirs(objdir, exist_ok=True)
    tag = os.path.join(objdir, "flags.txt")   # objects built with other flags are stale
    flags_now = " ".join(NVCC_FLAGS + extra)
    if not os.path.exists(tag) or open(tag).read() != flags_now:
        force = True
    compile_flags = [f for f in NVCC_FLAGS if f != "-shared"]

    def compile_one(src):
        obj = os.path.join(objdir, os.path.basename(src)[:-3] + ".o")
        if not force and not _newer(obj, [src] + headers):
            return obj, ""
        cmd = [nvcc] + compile_flags + extra + (["-Xptxas", "-v"] if verbose else []) + ["-I", os.path.join(ROOT, "include"), "-c", "-o", obj, src]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("nvcc failed:\n" + r.stdout + r.stderr)
        return obj, r.stdout + r.stderr

    with ThreadPoolExecutor(max_workers=max(1, len(srcs))) as ex:
        results = list(ex.map(compile_one, srcs))
    objs = [o for o, _ in results]
    if verbose:
        print("".join(t for _, t in results))
    if force or _newer(LIB, objs):
        r = subprocess.run([nvcc] + NVCC_FLAGS + ["-o", LIB] + objs + ["-ldl"], capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("link failed:\n" + r.stdout + r.stderr)
    open(tag, "w").write(flags_now)
    return LIB


def build_oracle(force=False):
    odir = os.path.join(ROOT, "oracle")
    if force:
        subprocess.run(["make", "-C", odir, "clean"], check=True, capture_output=True)
    r = subprocess.run(["make", "-C", odir], capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("oracle build failed:\n" + r.stdout + r.stderr)
    return os.path.join(odir, "_build", "librxoracle.so")


def build_all(force=False, verbose=False):
    return build_cuda(force, verbose), build_oracle(force)


if __name__ == "__main__":
    print(build_all(force="--force" in sys.argv, verbose="-v" in sys.argv))
