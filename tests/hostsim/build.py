"""Build the TEST-ONLY host instantiation of the device templates (see hostsim.cpp)."""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
CSRC = os.path.join(ROOT, "mallorn_astrophysics_b200", "csrc")
OUT = os.path.join(HERE, "libhostsim.so")


def build(force=False):
    srcs = [os.path.join(HERE, "hostsim.cpp")] + [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    if not force and os.path.exists(OUT) and all(os.path.getmtime(OUT) >= os.path.getmtime(s) for s in srcs):
        return OUT
    cmd = ["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-x", "c++", "-I", CSRC,
           os.path.join(HERE, "hostsim.cpp"), "-o", OUT, "-lm"]
    subprocess.check_call(cmd)
    return OUT


if __name__ == "__main__":
    print(build(force=True))
