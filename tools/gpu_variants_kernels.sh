#!/bin/bash
# usage: bash tools/gpu_variants_kernels.sh name1 name2 ...   (variants built by tools/build_variant.sh)
# GPU job: kernel-table lines of the templated shapes for the library and for build variants, then the share-gen cycle counts
mkdir -p gpurun_out
for v in base "$@"; do
  lib=$PWD/sda_b200/variants/lib_$v.so; [ $v = base ] && lib=$PWD/sda_b200/libsda_b200.so
  SDA_B200_LIB=$lib timeout 600 python tools/kernel_bench.py --only "packed_share cfg" 2>/dev/null > gpurun_out/kb_$v.jsonl
  echo "== $v"; python tools/kernels_md.py gpurun_out/kb_$v.jsonl | grep -i "tensor cores" | cut -c1-140
done
RX=packed_share_tc2 bash tools/gpu_variants_ncu.sh n8 base "$@"
