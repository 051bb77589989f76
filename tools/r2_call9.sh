#!/bin/bash
# round 2, GPU call 9 (1 GPU): the one GPU test that changed after call 8
O=gpurun_out/r2c9; mkdir -p $O
timeout 400 python -m pytest tests/test_gpu_hooked.py -m gpu -q -k "matmul or add_two" --timeout=300 > $O/pytest.txt 2>&1; echo "pytest rc=$?"; tail -4 $O/pytest.txt
