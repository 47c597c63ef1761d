#!/bin/bash
# Development aid: A/B of library variants (scripts/build_variant.sh) inside one gpurun call: alternates the
# variants so that box-to-box and thermal differences cancel.   usage: scripts/ab.sh "<bench args>" name1 name2 ...
args=$1; shift
for rep in 1 2 3; do
  for v in "$@"; do
    WXB_LIB=$PWD/whisperx_b200/_variants/lib_$v.so python bench.py --only-main $args 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read())
print('%-10s overlapped %.4f  serial %.4f  dp %.4f ms' % ('$v', d['ms_per_step'], d['value_serial']['ms_per_step'], d['roofline']['kernel_ms']))"
  done
done
