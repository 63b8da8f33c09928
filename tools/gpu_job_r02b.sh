#!/bin/bash
# GPU job: the whole -m gpu suite, the share-generation and reveal lines of the kernel table for the library and for build
# variants, and the share-gen kernel's cycle counts (ncu) per variant.
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
timeout 600 python tools/kernel_bench.py --only packed_ 2>/dev/null > gpurun_out/kb_base.jsonl
python tools/kernels_md.py gpurun_out/kb_base.jsonl | grep -i "packed_"
for v in "$@"; do
  SDA_B200_LIB=$PWD/sda_b200/variants/lib_$v.so timeout 600 python tools/kernel_bench.py --only "packed_share cfg" 2>/dev/null > gpurun_out/kb_$v.jsonl
  echo "== $v"; python tools/kernels_md.py gpurun_out/kb_$v.jsonl | grep -i "tensor cores"
done
RX=packed_share_tc2 bash tools/gpu_variants_ncu.sh n6 base "$@"
