#!/bin/bash
# GPU job: the host-buffer leg of bench.py under different client-thread counts and schedules
mkdir -p gpurun_out
for cfg in "1 phases" "2 phases" "3 phases" "2 pipelined" "3 pipelined" "4 pipelined" "6 pipelined"; do
  set -- $cfg
  timeout 300 python bench.py --steps 3 --e2e-threads $1 --e2e-mode $2 --no-configs45 --no-round-sweep --no-cpu-baseline > gpurun_out/e2e_$1_$2.json 2> gpurun_out/e2e_$1_$2.err
  python -c "
import json; d=json.load(open('gpurun_out/e2e_$1_$2.json'))['e2e']; print('threads $1 $2: %.4g el/s  (%.2f ms per step)  pageable %.3g  link %s' % (d['value'], 4e10/d['value'], d['pageable_value'], d['link']['duplex_each_way_GBps']))"
done
