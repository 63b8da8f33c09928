#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_hostpipe.py tests/test_gpu_threads.py -m gpu -q -x > gpurun_out/pytest_hostpipe.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_hostpipe.log
