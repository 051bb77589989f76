#!/bin/bash
# round 2, GPU call 11 (1 GPU): two hooked pairs through the shared pool after the pool-header change (version 3: pid namespace)
O=gpurun_out/r2c11; mkdir -p $O
timeout 200 python -m pytest tests/test_gpu_hooked.py -m gpu -q -k "add_two or oversubscribed_pair" --timeout=150 > $O/pytest.txt 2>&1; echo "pytest rc=$?"; tail -4 $O/pytest.txt
