#!/bin/bash
# Development aid: instructions per frame of the recurrence loops for every K (from SASS).
# usage: scripts/sass_stats.sh [extra nvcc flags]
set -e
cd "$(dirname "$0")/.."
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -fmad=false -std=c++17 -I include "$@" -cubin -o /tmp/wxb_inst.cubin whisperx_b200/csrc/wxb_inst.cu
for K in 1 2 3 4 5 6 7 8; do
  cuobjdump -sass -fun "_ZN3wxb12align_kernelILi${K}ELb0ELb0EEEvNS_6ParamsEiii" /tmp/wxb_inst.cubin | grep -E "^\s+/\*[0-9a-f]{4,5}\*/" | sed -E 's#/\* 0x[0-9a-f]+ \*/##' > /tmp/k$K.sass
  python3 - $K <<'PY'
import re, sys
K=sys.argv[1]
ins=[]
for l in open(f'/tmp/k{K}.sass'):
    m=re.match(r'\s*/\*([0-9a-f]{4,5})\*/\s*(.*?);', l)
    if m: ins.append((int(m.group(1),16), m.group(2).strip()))
addr={a:i for i,(a,_) in enumerate(ins)}
out=[]
for i,(a,t) in enumerate(ins):
    m=re.search(r'BRA(?:\.U)?\s+(?:!?U?P\d,\s*)?(0x[0-9a-f]+)', t)
    if m:
        tgt=int(m.group(1),16)
        if tgt<a and tgt in addr:
            body=ins[addr[tgt]:i+1]
            n=sum(1 for _,x in body if 'SHFL.UP' in x)
            if n: out.append(f"{len(body)}/{n}={len(body)/n:.1f}")
ldl=sum(1 for _,x in ins if 'LDL' in x or 'STL' in x)
print(f"K={K} total={len(ins)} loops(instr/frames)={out} local={ldl}")
PY
done
