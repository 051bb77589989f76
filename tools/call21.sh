#!/bin/bash
set -u
OUT=gpurun_out/call21
mkdir -p $OUT
timeout 120 python -m pytest tests/test_gpu_engine.py::test_host_io_upload_and_readback_without_the_gpu tests/test_gpu_hooked.py::test_loading_client_does_not_take_the_gpu -m gpu -q -x > $OUT/pytest_new.txt 2>&1; echo "pytest(new) rc=$?" | tee $OUT/summary.txt
tail -15 $OUT/pytest_new.txt | cut -c1-300 | tee -a $OUT/summary.txt
