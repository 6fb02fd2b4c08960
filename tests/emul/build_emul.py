"""TEST INFRASTRUCTURE ONLY.  Compiles kagnn_b200/csrc/backward.cu as serial HOST code (g++, -DKAGNN_HOST_CHECK, see
kagnn_b200/csrc/launch.cuh and host_check.h) into tests/emul/_build/libkagnn_bwd_hostcheck.so, so that the CPU test-suite can
check the index arithmetic of the backward kernels in a container without a GPU.  Never loaded by the product."""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
SRC = os.path.join(ROOT, "kagnn_b200", "csrc", "backward.cu")
OUT_DIR = os.path.join(HERE, "_build")
OUT = os.path.join(OUT_DIR, "libkagnn_bwd_hostcheck.so")


def build() -> str:
    os.makedirs(OUT_DIR, exist_ok=True)
    deps = [SRC, os.path.join(HERE, "host_check.h"), os.path.join(ROOT, "kagnn_b200", "csrc", "launch.cuh"),
            os.path.join(ROOT, "include", "kagnn_b200.h")]
    if os.path.exists(OUT) and all(os.path.getmtime(OUT) >= os.path.getmtime(d) for d in deps):
        return OUT
    cmd = ["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-DKAGNN_HOST_CHECK", "-I", HERE, "-x", "c++", SRC, "-o", OUT]
    subprocess.run(cmd, check=True)
    return OUT


if __name__ == "__main__":
    print(build())
