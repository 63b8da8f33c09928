for v in base revpf revmb16; do
  lib=$PWD/sda_b200/variants/lib_$v.so; [ $v = base ] && lib=$PWD/sda_b200/libsda_b200.so
  SDA_B200_LIB=$lib timeout 300 python tools/kernel_bench.py --only packed_reconstruct 2>/dev/null > gpurun_out/kb_$v.jsonl
  echo "== $v"; python tools/kernels_md.py gpurun_out/kb_$v.jsonl | grep packed_
done
bash tools/gpu_sanitize.sh; cat gpurun_out/r02_sanitizer.txt
