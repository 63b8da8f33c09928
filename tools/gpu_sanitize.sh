#!/bin/bash
# compute-sanitizer over the small parity cases (memcheck: out-of-bounds / misaligned; racecheck: shared-memory hazards;
# synccheck: barrier misuse).  Round 2 adds the paired-tile, run-time-shaped and any-share-count kernels, the snapshot
# transpose and the fused encode + mask.
mkdir -p gpurun_out
for tool in memcheck racecheck synccheck; do
  timeout 900 compute-sanitizer --tool $tool --error-exitcode 9 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/sanitize_$tool.log 2>&1
  echo "rc=$?" >> gpurun_out/sanitize_$tool.log
done
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_device.py -q -m gpu -k "packed_kernels and cfg3" > gpurun_out/sanitize_memcheck_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/sanitize_memcheck_pytest.log
# the sliced host entry points (three streams) and the reveal kernel
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_hostpipe.py tests/test_gpu_parity.py -q -m gpu -k "524288 or reconstruct" > gpurun_out/sanitize_memcheck_hostpipe.log 2>&1; echo "rc=$?" >> gpurun_out/sanitize_memcheck_hostpipe.log
# round 2 kernels: run-time-shaped share-gen (two shapes), additive split with n = 9 / 16, snapshot transpose, fused encode + mask
timeout 1200 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_device.py tests/test_gpu_parity.py tests/test_gpu_codec.py tests/test_gpu_federated.py -q -m gpu \
  -k "(runtime_shaped and (shape0 or shape3) and 2305843009213693951) or (additive_generate and (9-433 or 16-433)) or snapshot_transpose_matches or fused_encode_mask" \
  > gpurun_out/sanitize_memcheck_r02.log 2>&1; echo "rc=$?" >> gpurun_out/sanitize_memcheck_r02.log
timeout 1200 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_device.py -q -m gpu \
  -k "(runtime_shaped and shape0 and 2305843009213693951) or (packed_kernels and cfg3 and tensor)" > gpurun_out/sanitize_racecheck_r02.log 2>&1; echo "rc=$?" >> gpurun_out/sanitize_racecheck_r02.log
# the paired-tile kernel with the share count at run time (even and odd t, two share groups), the reveal kernel with bulk
# stores (aligned and unaligned outputs), the single-launch varint codec
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_device.py tests/test_gpu_codec.py -q -m gpu \
  -k "(run_time_share_count and (shape0 or shape6 or shape9)) or (reveal_many_tiles and cfg4) or varint" \
  > gpurun_out/sanitize_memcheck_r02b.log 2>&1; echo "rc=$?" >> gpurun_out/sanitize_memcheck_r02b.log
timeout 1500 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_device.py tests/test_gpu_codec.py -q -m gpu \
  -k "(run_time_share_count and shape6 and 0-0) or (reveal_many_tiles and cfg3 and 0) or varint_known or varint_encode_decode" \
  > gpurun_out/sanitize_racecheck_r02b.log 2>&1; echo "rc=$?" >> gpurun_out/sanitize_racecheck_r02b.log
# the fused mask -> share kernel (masks added in place in the staging buffer)
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_device.py -q -m gpu \
  -k "mask_share_generate_matches and (cfg3 or cfg4) and (full or chacha)" > gpurun_out/sanitize_memcheck_masked.log 2>&1; echo "rc=$?" >> gpurun_out/sanitize_memcheck_masked.log
timeout 1500 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_device.py -q -m gpu \
  -k "mask_share_generate_matches and cfg3 and full" > gpurun_out/sanitize_racecheck_masked.log 2>&1; echo "rc=$?" >> gpurun_out/sanitize_racecheck_masked.log
timeout 1500 compute-sanitizer --tool synccheck --error-exitcode 9 python -m pytest tests/test_gpu_device.py -q -m gpu \
  -k "mask_share_generate_matches and cfg5 and full" > gpurun_out/sanitize_synccheck_masked.log 2>&1; echo "rc=$?" >> gpurun_out/sanitize_synccheck_masked.log
# the paired-tile share-gen -> clerk-sum kernel (TMEM accumulation over participants, drains, two participants per step), the
# keystream-constant variants of the mask / mask-combine / additive kernels
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_device.py tests/test_gpu_parity.py -q -m gpu \
  -k "share_generate_combine or tmem_accumulation or (chacha_mask and 2305843009213693951) or (additive_generate and 3-2305843009213693951)" \
  > gpurun_out/sanitize_memcheck_fused2.log 2>&1; echo "rc=$?" >> gpurun_out/sanitize_memcheck_fused2.log
timeout 1500 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_device.py -q -m gpu \
  -k "share_generate_combine or (tmem_accumulation and cfg3)" > gpurun_out/sanitize_racecheck_fused2.log 2>&1; echo "rc=$?" >> gpurun_out/sanitize_racecheck_fused2.log
# the paired reveal kernel (16-byte share loads, two accumulators, bulk stores)
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_device.py tests/test_gpu_parity.py -q -m gpu \
  -k "reveal_many_tiles or (reconstruct_generic and (shape0 or shape3 or shape5 or shape9))" > gpurun_out/sanitize_memcheck_reveal2.log 2>&1; echo "rc=$?" >> gpurun_out/sanitize_memcheck_reveal2.log
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_device.py -q -m gpu \
  -k "reveal_many_tiles and cfg5" > gpurun_out/sanitize_racecheck_reveal2.log 2>&1; echo "rc=$?" >> gpurun_out/sanitize_racecheck_reveal2.log
(echo "compute-sanitizer on a B200 (tools/gpu_sanitize.sh), round 2"; for f in gpurun_out/sanitize_*.log; do echo "== $(basename $f)"; grep -v "^$" $f | tail -6; done) > gpurun_out/r02_sanitizer.txt
