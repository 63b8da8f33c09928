#!/bin/bash
# compute-sanitizer over the small parity cases (memcheck: out-of-bounds / misaligned; racecheck: shared-memory hazards;
# synccheck: barrier misuse)
mkdir -p gpurun_out
for tool in memcheck racecheck synccheck; do
  timeout 900 compute-sanitizer --tool $tool --error-exitcode 9 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/sanitize_$tool.log 2>&1
  echo "rc=$?" >> gpurun_out/sanitize_$tool.log
done
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_device.py -q -m gpu -k "packed_kernels and cfg3" > gpurun_out/sanitize_memcheck_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/sanitize_memcheck_pytest.log
# the sliced host entry points (three streams) and the reveal kernel
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_hostpipe.py tests/test_gpu_parity.py -q -m gpu -k "524288 or reconstruct" > gpurun_out/sanitize_memcheck_hostpipe.log 2>&1; echo "rc=$?" >> gpurun_out/sanitize_memcheck_hostpipe.log
