#!/bin/bash
mkdir -p gpurun_out
for pdl in 1 0; do echo "=== PQ3D_PDL=$pdl"; PQ3D_PDL=$pdl timeout 600 python tools/chainbench.py 2>&1 | tail -20; done | tee gpurun_out/chainbench.log
