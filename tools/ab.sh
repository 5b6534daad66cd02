#!/bin/bash
# usage (GPU box): tools/ab.sh "variant1 variant2 ..." "workload1 workload2 ..." [steps]  ->  one line per (variant, workload)
# A/B of build/lib_<variant>.so builds (tools/build_variant.sh) through bench.py's encode leg only (whole-stream parity on).
# "product" = the in-tree library; "product16" = the same with $HB_FORCE_WORKERS=16.
steps=${3:-50}
for v in $1; do
  for w in $2; do
    lib=build/lib_$v.so; fw=0
    [ "$v" = "product" ] && lib=huffman-gpu_b200/libhuffb200.so
    [ "$v" = "product16" ] && lib=huffman-gpu_b200/libhuffb200.so && fw=16
    HB_FORCE_WORKERS=$fw HB_LIB=$lib python bench.py --workload $w --steps $steps --no-cpu --no-e2e --no-pipeline --no-per-config 2>gpurun_out/ab_err.txt | python -c "
import json,sys
for line in sys.stdin:
    if line.startswith('{'):
        l=json.loads(line)
        print('$v', '$w', 'GB/s %.0f' % l['value'], 'frac %.3f' % l['roofline']['frac'], l['detail']['kernel_variant'], 'parity', l['parity'].get('whole_stream_vs_cpu_vlc_encode'), l['parity']['shard_windows_vs_cpu_oracle'], l.get('parity_error'))
"
  done
done
