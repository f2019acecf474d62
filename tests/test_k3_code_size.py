"""Build property of the K3 scalar passes, checked where nvcc cross-compiles (no GPU): with the team helpers shared
(MLRN_SHARED_HELPERS, team.cuh) no k_pass kernel is larger than 9 k instructions (a kernel's count includes its copy of
every helper it calls: ptxas places called device functions inside the calling kernel's code).  Inlined, the TDE kernel was 13.3 k instructions (213 KB) and ran 2.1x slower."""
import os
import re
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "mallorn_astrophysics_b200", "csrc")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
CUOBJDUMP = os.path.join(os.path.dirname(NVCC), "cuobjdump")


def _functions(obj):
    """{mangled name: instruction count} from cuobjdump -sass."""
    sass = subprocess.run([CUOBJDUMP, "-sass", obj], capture_output=True, text=True, check=True).stdout
    out, name = {}, None
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            name = m.group(1)
            out[name] = 0
        elif name and re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+\S", line):
            out[name] += 1
    return out


@pytest.mark.skipif(not (os.path.exists(NVCC) and os.path.exists(CUOBJDUMP)), reason="needs nvcc and cuobjdump")
def test_k3_kernels_keep_their_code_size(tmp_path):
    obj = str(tmp_path / "passes.o")
    log = subprocess.run([NVCC, "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-Xptxas", "-v",
                          "-c", os.path.join(CSRC, "passes.cu"), "-o", obj], capture_output=True, text=True, cwd=CSRC)
    assert log.returncode == 0, log.stderr[-2000:]
    fns = _functions(obj)
    kernels = {k: v for k, v in fns.items() if "k_pass" in k}
    assert len(kernels) == 4, sorted(fns)
    assert max(kernels.values()) <= 9000, kernels                 # 16 B each; the inlined TDE kernel was 13.3 k (213 KB)
    frames = [int(x) for x in re.findall(r"(\d+) bytes stack frame", log.stderr)]
    assert frames and max(frames) <= 256, frames
