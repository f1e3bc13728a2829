#!/bin/bash
# padded replay of the other model families, triclinic GPU builder, then the whole GPU suite
mkdir -p gpurun_out
timeout 300 python profiles/debug_padded.py > gpurun_out/debug_padded.txt 2>&1; grep -E "==|FAILED|<<<" gpurun_out/debug_padded.txt | head -40
timeout 900 python -m pytest tests/test_gpu_store.py tests/test_gpu_builder.py -q > gpurun_out/tests_store.log 2>&1
echo "pytest exit $?" >> gpurun_out/tests_store.log
grep -E "FAILED|ERROR|passed|failed|pytest exit" gpurun_out/tests_store.log | tail -15
rm -f gpurun_out/parity_errors.txt
timeout 1200 python -m pytest tests -m gpu -q --deselect tests/test_gpu_store.py --deselect tests/test_gpu_builder.py > gpurun_out/tests_gpu_rest.log 2>&1
echo "pytest rest exit $?" >> gpurun_out/tests_gpu_rest.log
grep -E "FAILED|ERROR|passed|failed|pytest rest exit" gpurun_out/tests_gpu_rest.log | tail -15
