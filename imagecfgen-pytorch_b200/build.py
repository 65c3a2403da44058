"""Build libicf_b200.so in-tree with nvcc for sm_100a (no torch headers: the library is a plain C-ABI)."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
OUT = os.path.join(HERE, "icf_b200", "libicf_b200.so")
SOURCES = ["icf_api.cu", "icf_elementwise.cu", "icf_conv_simt.cu", "icf_conv_tc.cu", "icf_conv_ws.cu", "icf_wgrad_px8.cu", "icf_conv_sc.cu", "icf_finetune_scm.cu"]


def needs_build():
    if not os.path.exists(OUT):
        return True
    t = os.path.getmtime(OUT)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(ROOT, "include", "icf.h")]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    if not force and not needs_build():
        return OUT
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    cmd = [nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
           "--use_fast_math" if os.environ.get("ICF_FAST_MATH") else "-DICF_PRECISE_MATH",
           "-Xcompiler", "-fPIC,-O2", "-shared", "-I", os.path.join(ROOT, "include"), "-I", CSRC,
           "-o", OUT] + [os.path.join(CSRC, s) for s in SOURCES]
    if os.environ.get("ICF_WS_INSTRUMENT"):
        cmd.insert(1, "-DICF_WS_INSTRUMENT")
    if verbose:
        cmd.insert(1, "-Xptxas=-v")
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
        raise RuntimeError("nvcc failed building libicf_b200.so")
    if verbose:
        sys.stderr.write(r.stderr)
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
