"""Build libicf_b200.so in-tree with nvcc for sm_100a (no torch headers: the library is a plain C-ABI).

Every source is compiled to an object file in parallel, then linked.  A fingerprint of the sources (sha256 over csrc/ and
include/icf.h) is stored next to the library; ``icf_b200.lib.load`` refuses a library whose fingerprint does not match the
tree, so a stale binary can never serve the tests."""
import hashlib
import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
OUT = os.path.join(HERE, "icf_b200", "libicf_b200.so")
STAMP = os.path.join(HERE, "icf_b200", "libicf_b200.srchash")
OBJ = os.path.join(HERE, "build")
SOURCES = ["icf_api.cu", "icf_elementwise.cu", "icf_conv_simt.cu", "icf_conv_tc.cu", "icf_conv_ws.cu", "icf_wgrad_px8.cu",
           "icf_conv_sc.cu", "icf_conv_cm.cu", "icf_finetune_scm.cu", "icf_spectro.cu"]


def source_hash() -> str:
    h = hashlib.sha256()
    files = sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cu", ".cuh", ".h")))
    for path in files + [os.path.join(ROOT, "include", "icf.h")]:
        h.update(os.path.basename(path).encode())
        with open(path, "rb") as f:
            h.update(f.read())
    flags = "fast" if os.environ.get("ICF_FAST_MATH") else "precise"
    h.update((flags + ("+instr" if os.environ.get("ICF_WS_INSTRUMENT") else "")).encode())
    return h.hexdigest()


def built_hash() -> str:
    try:
        with open(STAMP) as f:
            return f.read().strip()
    except OSError:
        return ""


def needs_build():
    return not os.path.exists(OUT) or built_hash() != source_hash()


def find_nvcc():
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    return nvcc if os.path.exists(nvcc) else shutil.which("nvcc")


def build(force=False, verbose=False):
    if not force and not needs_build():
        return OUT
    nvcc = find_nvcc()
    if nvcc is None:
        raise RuntimeError("nvcc not found: cannot build libicf_b200.so")
    os.makedirs(OBJ, exist_ok=True)
    common = [nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
              "--use_fast_math" if os.environ.get("ICF_FAST_MATH") else "-DICF_PRECISE_MATH",
              "-Xcompiler", "-fPIC,-O2", "-I", os.path.join(ROOT, "include"), "-I", CSRC]
    if os.environ.get("ICF_WS_INSTRUMENT"):
        common.insert(1, "-DICF_WS_INSTRUMENT")
    if verbose:
        common.insert(1, "-Xptxas=-v")

    def compile_one(src):
        obj = os.path.join(OBJ, src.replace(".cu", ".o"))
        r = subprocess.run(common + ["-c", os.path.join(CSRC, src), "-o", obj], capture_output=True, text=True)
        return src, obj, r

    with ThreadPoolExecutor(max_workers=min(len(SOURCES), os.cpu_count() or 1)) as ex:
        results = list(ex.map(compile_one, SOURCES))
    for src, obj, r in results:
        if r.returncode != 0:
            sys.stderr.write(r.stdout + r.stderr)
            raise RuntimeError(f"nvcc failed compiling {src}")
        if verbose:
            sys.stderr.write(r.stderr)
    r = subprocess.run([nvcc, "-shared", "-o", OUT] + [obj for _, obj, _ in results], capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
        raise RuntimeError("nvcc failed linking libicf_b200.so")
    with open(STAMP, "w") as f:
        f.write(source_hash())
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
