#!/bin/bash
# GPU job: -m gpu suite, kernel table, and ncu --set full of the reveal kernels and the codec
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
timeout 900 python tools/kernel_bench.py 2>/dev/null > gpurun_out/r02_kernels.jsonl
python tools/kernels_md.py gpurun_out/r02_kernels.jsonl > gpurun_out/r02_kernels_table.md; grep -i "packed_\|varint" gpurun_out/r02_kernels_table.md
timeout 600 ncu --set full --clock-control none --import-source on -k regex:reveal_tc -s 6 -c 2 -o gpurun_out/r02_reveal -f \
  python tools/kernel_bench.py --only packed_reconstruct > gpurun_out/ncu_reveal.log 2>&1
timeout 600 ncu --set full --clock-control none -k regex:varint -s 4 -c 2 -o gpurun_out/r02_codec -f \
  python tools/kernel_bench.py --only varint > gpurun_out/ncu_codec.log 2>&1
ls -la gpurun_out/*.ncu-rep
