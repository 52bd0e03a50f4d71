"""The numerics contract depends on a ptxas behaviour -fmad=false does not control (packed mul -> add contraction,
DESIGN.md s4): audit the SASS of the built library on every CPU test run (cuobjdump ships with the toolkit)."""
import os
import shutil
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))


@pytest.mark.skipif(shutil.which("cuobjdump") is None, reason="cuobjdump not on PATH")
def test_no_contracted_fma_in_the_encode_kernels():
    import sass_audit
    lib = os.path.join(ROOT, "tileconv_b200", "libtcb200.so")
    problems, summary = sass_audit.audit(lib)
    kernels = [n for n in summary if "encode_tiles_kernel" in n]
    assert len(kernels) >= 26
    assert problems == []
    # the audit must be looking at real code: the cluster kernels do contain exact-constant FFMA2s and separate FMULs / FADDs
    iterative_bc1 = next(n for n in kernels if "ILi1ELb0ELi2ELb0E" in n)
    assert summary[iterative_bc1]["FFMA2"] > 0 and summary[iterative_bc1]["FMUL"] > 100 and summary[iterative_bc1]["FADD"] > 100


@pytest.mark.skipif(shutil.which("cuobjdump") is None or shutil.which("nvcc") is None, reason="CUDA toolkit not on PATH")
def test_the_audit_catches_a_contraction(tmp_path):
    """A kernel named like ours with a packed multiply feeding a packed add -- the pattern ptxas fuses regardless of
    -fmad=false -- must be flagged."""
    import sass_audit
    src = tmp_path / "bad.cu"
    src.write_text(r'''
#include <cuda_runtime.h>
__global__ void encode_tiles_kernel_bad(const float2* a, const float2* b, float2* c, float* s)
{
  float2 x = a[threadIdx.x], y = b[threadIdx.x];
  float2 p = __fmul2_rn(x, y);
  c[threadIdx.x] = __fadd2_rn(p, y);          // contracted to FFMA2 by ptxas 12.9
}
''')
    lib = tmp_path / "bad.so"
    subprocess.run(["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-fmad=false", "-shared", "-Xcompiler", "-fPIC",
                    "-o", str(lib), str(src)], check=True, capture_output=True)
    problems, _ = sass_audit.audit(str(lib))
    assert problems and "FFMA2" in problems[0]
