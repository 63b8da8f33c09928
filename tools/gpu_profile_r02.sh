#!/bin/bash
# round-2 evidence run on one B200: launch list of a bench step, full ncu captures of the dominant kernels at the bench
# configuration, and the clocks during a 20-step bench (the driver's command line)
mkdir -p gpurun_out
Q="--no-e2e --no-cpu-baseline --no-round-sweep --no-configs45"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r02_launches.csv \
    python bench.py --steps 2 --warmup 3 $Q > gpurun_out/r02_ncu_launches.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:packed_share_tc2 -s 3 -c 1 -o gpurun_out/r02_k2 \
    python bench.py --steps 1 --warmup 3 $Q > gpurun_out/r02_ncu_k2.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:combine_kernel -s 30 -c 1 -o gpurun_out/r02_k3 \
    python bench.py --steps 1 --warmup 3 $Q > gpurun_out/r02_ncu_k3.log 2>&1
# the shape-generic kernels on k=3 / t=3 / n=7: the paired-tile kernel with the share count at run time (20 rounds), the
# run-time-shaped kernel (12 rounds)
for v in tc2n:20:packed_share_tc2 tcg:12:packed_share_tcg; do
  name=${v%%:*}; rest=${v#*:}; rounds=${rest%%:*}; rx=${rest#*:}
  ROUNDS=$rounds timeout 900 ncu --set full --clock-control none --import-source on -k regex:$rx -c 1 -o gpurun_out/r02_$name -f \
    python - > gpurun_out/r02_ncu_$name.log 2>&1 <<'PY'
import hashlib, os, torch, sda_b200
from sda_b200 import params
ctx = sda_b200.Context(0, rng_rounds=int(os.environ["ROUNDS"]))
s = params.LinearSecretSharingScheme.PackedShamir(3, 7, 3, params.P61, params.ROOT_ORDER_31, params.ROOT_ORDER_41)
P, dim = 64, 10_000_000
sec = torch.empty((P, dim), dtype=torch.int64, device="cuda"); ctx.synth_fill_dev(3, params.P61, 0, P * dim, sec)
out = torch.empty((P, 7, s.batches(dim)), dtype=torch.int64, device="cuda")
seeds = b"".join(hashlib.sha256(b"%d" % i).digest() for i in range(P))
ctx.share_generate_dev(s, sec, dim, P, dim, seeds, out); ctx.synchronize()
PY
done
# the fused share-gen -> clerk-sum kernel (paired tiles) on config #5's shape, 256 participants x 10M
timeout 600 ncu --set full --clock-control none --import-source on -k regex:packed_share_combine_tc2 -c 1 -o gpurun_out/r02_fused2 -f \
  python - > gpurun_out/r02_ncu_fused2.log 2>&1 <<'PY'
import hashlib, torch, sda_b200
from sda_b200 import params
ctx = sda_b200.Context(0)
s = params.config5()
P, dim = 256, 10_000_000
sec = torch.empty((P, dim), dtype=torch.int64, device="cuda"); ctx.synth_fill_dev(3, params.P61, 0, P * dim, sec)
out = torch.empty((s.output_size(), s.batches(dim)), dtype=torch.int64, device="cuda")
seeds = b"".join(hashlib.sha256(b"%d" % i).digest() for i in range(P))
ctx.share_generate_combine_dev(s, sec, dim, P, dim, seeds, out); ctx.synchronize()
print(ctx.last_kernel())
PY
# reveal (config #4's 9 x [2M] and 7 clerks x [3.33M]) and the varint codec
timeout 600 ncu --set full --clock-control none --import-source on -k regex:reveal_tc -s 6 -c 2 -o gpurun_out/r02_reveal -f \
  python tools/kernel_bench.py --only packed_reconstruct > gpurun_out/r02_ncu_reveal.log 2>&1
timeout 600 ncu --set full --clock-control none -k regex:varint -s 4 -c 2 -o gpurun_out/r02_codec -f \
  python tools/kernel_bench.py --only varint > gpurun_out/r02_ncu_codec.log 2>&1
timeout 600 python bench.py --steps 20 --warmup 5 --no-configs45 > gpurun_out/r02_bench_20.json 2> gpurun_out/r02_bench_20.err
python -c "
import json; d=json.load(open('gpurun_out/r02_bench_20.json')); print('20 steps: frac %.4f ms %.3f value %.4g e2e %.4g' % (d['roofline']['frac'], d['roofline']['ms_per_launch'], d['value'], d['e2e']['value']), d['clocks'])"
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,power.limit,memory.total --format=csv > gpurun_out/r02_smi.txt 2>&1
lscpu | head -20 > gpurun_out/r02_lscpu.txt 2>&1
