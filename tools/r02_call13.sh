#!/bin/bash
# row f4 on the B200: the crypto tests, the new k_fast layout test, crypto throughput
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_crypto.py tests/test_gpu_parity.py -m gpu -q -x --timeout 600 -k "crypto or aes or pkzip or both_table_layouts or encrypt" 2>&1 | tail -5
timeout 600 python tools/gpu_crypto_bench.py 2>&1 | tail -12
