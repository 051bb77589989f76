#!/bin/bash
# round 2, GPU call 10 (1 GPU): the hooked-application tests after the last hook.c changes
# (cap reservation, atomic flags, capture-aware cuMemFreeAsync)
O=gpurun_out/r2c10; mkdir -p $O
timeout 330 python -m pytest tests/test_gpu_hooked.py -m gpu -q -k "add_two or graph or cooperative or stream_ordered or uvm_mode or no_swap" --timeout=200 > $O/pytest.txt 2>&1; echo "pytest rc=$?"; tail -4 $O/pytest.txt
