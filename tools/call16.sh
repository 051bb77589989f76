#!/bin/bash
set -u
OUT=gpurun_out/call16
mkdir -p $OUT
echo "== new GPU tests first" | tee $OUT/summary.txt
timeout 400 python -m pytest tests/test_gpu_engine.py::test_host_io_upload_and_readback_without_the_gpu tests/test_gpu_hooked.py::test_loading_client_does_not_take_the_gpu -m gpu -q > $OUT/pytest_new.txt 2>&1; echo "pytest(new) rc=$?" | tee -a $OUT/summary.txt
tail -40 $OUT/pytest_new.txt | cut -c1-400 | tee -a $OUT/summary.txt
echo "== whole GPU suite" | tee -a $OUT/summary.txt
timeout 900 python -m pytest tests -m gpu -q -x > $OUT/pytest_gpu.txt 2>&1; echo "pytest(gpu) rc=$?" | tee -a $OUT/summary.txt
tail -15 $OUT/pytest_gpu.txt | cut -c1-400 | tee -a $OUT/summary.txt
echo "== smoke" | tee -a $OUT/summary.txt
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 | tee -a $OUT/summary.txt
echo "== config #3 shape: matmul(ones, ones), 10% of HBM, TQ 5: ours, reference" | tee -a $OUT/summary.txt
timeout 400 python bench.py --kind matmul --pattern ones --hbm-fraction 0.1 --tq 5 --steps 3 --warmup 3 > $OUT/matmul_ours.json 2> $OUT/matmul_ours.err; echo "rc=$?" | tee -a $OUT/summary.txt
tail -1 $OUT/matmul_ours.json | cut -c1-2500 | tee -a $OUT/summary.txt
timeout 400 python bench.py --impl reference --kind matmul --pattern ones --hbm-fraction 0.1 --tq 5 --steps 3 --warmup 3 > $OUT/matmul_ref.json 2> $OUT/matmul_ref.err; echo "rc=$?" | tee -a $OUT/summary.txt
tail -1 $OUT/matmul_ref.json | cut -c1-2500 | tee -a $OUT/summary.txt
