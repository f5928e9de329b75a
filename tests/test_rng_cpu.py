"""CPU: pq3d_b200/rng.py (the tensor restatement tests use to replay dropout masks) agrees bit for bit with the
host-callable RNG functions in pq3d_b200/csrc/ptx.cuh that the kernels use."""
import os
import shutil
import subprocess

import pytest
import torch

from pq3d_b200 import rng

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = r'''
#include <cstdio>
#include <cstdlib>
#include "ptx.cuh"
int main(int argc, char** argv) {
  const uint32_t seed = strtoul(argv[1], 0, 10), site = strtoul(argv[2], 0, 10);
  const float p = atof(argv[3]);
  const uint32_t key = pq3d::drop_key(seed, site), th = pq3d::drop_threshold(p);
  for (uint32_t i = 0; i < 4096; ++i) putchar(pq3d::drop_keep(key, i * 2654435761u, th) ? '1' : '0');
  return 0;
}
'''


@pytest.mark.skipif(shutil.which("nvcc") is None and not os.path.exists("/usr/local/cuda/bin/nvcc"), reason="needs nvcc")
def test_rng_restatement_matches_header(tmp_path):
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    src = tmp_path / "rng_host.cu"
    src.write_text(SRC)
    exe = tmp_path / "rng_host"
    subprocess.run([nvcc, "-O1", "-I", os.path.join(ROOT, "pq3d_b200", "csrc"), str(src), "-o", str(exe)], check=True,
                   capture_output=True)
    for seed, site, p in ((12345, 0, 0.1), (4000000000, 37, 0.6), (7, 70, 0.5)):
        out = subprocess.run([str(exe), str(seed), str(site), str(p)], check=True, capture_output=True, text=True).stdout
        idx = (torch.arange(4096, dtype=torch.int64) * 2654435761) & 0xFFFFFFFF
        keep = rng.keep_mask(seed, site, idx, p)
        assert "".join("1" if k else "0" for k in keep.tolist()) == out
        assert abs(keep.float().mean().item() - (1 - p)) < 0.05
