set -x
(timeout 600 python -m pytest tests/test_aux_gpu.py -m gpu -q -s) > gpurun_out/r2c_aux.log 2>&1; tail -25 gpurun_out/r2c_aux.log
(timeout 1500 python -m pytest tests/test_parity_matched_gpu.py -m gpu -q -s) > gpurun_out/r2c_matched.log 2>&1; grep -E "^c[1-4]|passed|failed|Error|assert" gpurun_out/r2c_matched.log | cut -c1-600 | tail -40
