"""TEST INFRASTRUCTURE ONLY -- recipe that compiles the REFERENCE's own compiled code for the path.

The only compiled code the reference has on the path is the SPN scan
(`/root/reference/dmb/ops/spn/src/gaterecurrent2dnoind_kernel.cu:130-685`: plain CUDA C, no torch types; its
pybind11/torch wrapper `gaterecurrent2dnoind_cuda.cpp` is NOT used).  This compiles that file, from where it lies,
with `oracle/spn_ref_shim.cu` (our C-ABI dispatch around its launchers) into `oracle/_ref/libspn_ref.so`:

    python oracle/build_ref.py            # needs /root/reference (build container); no-op with a message otherwise

`oracle/_ref/` is git-ignored (no reference code enters the history) but travels to the GPU box with the snapshot,
where `tests/test_gpu_spn_ref.py` uses it as the checker for `csrc/scans.cu`.  Nothing in the product imports it.
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF_SRC = "/root/reference/dmb/ops/spn/src"
OUT_DIR = os.path.join(HERE, "_ref")
OUT = os.path.join(OUT_DIR, "libspn_ref.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")


def build_ref(force=False):
    """Returns the path of the built library, or None when the reference tree is absent."""
    kernel = os.path.join(REF_SRC, "gaterecurrent2dnoind_kernel.cu")
    shim = os.path.join(HERE, "spn_ref_shim.cu")
    if not os.path.exists(kernel):
        return OUT if os.path.exists(OUT) else None
    os.makedirs(OUT_DIR, exist_ok=True)
    if not force and os.path.exists(OUT) and os.path.getmtime(OUT) >= max(os.path.getmtime(kernel), os.path.getmtime(shim)):
        return OUT
    cmd = [NVCC, "-gencode", "arch=compute_100a,code=sm_100a", "-O2", "-shared", "-Xcompiler", "-fPIC", "-w",
           "-I", REF_SRC, kernel, shim, "-o", OUT]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode:
        sys.stderr.write(" ".join(cmd) + "\n" + r.stdout + r.stderr)
        raise RuntimeError("reference SPN kernel did not compile")
    return OUT


if __name__ == "__main__":
    print(build_ref(force="-f" in sys.argv))
