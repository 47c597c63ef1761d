#!/bin/bash
# Development aid: build libwxb_align.so from the csrc/ of a given git revision (or the working tree with "WORK")
# into whisperx_b200/_variants/lib_<name>.so, for A/B runs inside one gpurun call (select with WXB_LIB=...).
# usage: scripts/build_variant.sh <name> <rev|WORK> [extra nvcc flags]
set -e
cd "$(dirname "$0")/.."
name=$1; rev=$2; shift 2
tmp=$(mktemp -d)
mkdir -p $tmp/whisperx_b200/csrc $tmp/include whisperx_b200/_variants
if [ "$rev" = "WORK" ]; then
  cp whisperx_b200/csrc/* $tmp/whisperx_b200/csrc/; cp include/wxb_align.h $tmp/include/
else
  for f in $(git ls-tree --name-only $rev whisperx_b200/csrc/); do git show $rev:$f > $tmp/$f; done
  git show $rev:include/wxb_align.h > $tmp/include/wxb_align.h
fi
FLAGS="-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -fmad=false -std=c++17 -Xcompiler -fPIC -I $tmp/include $@"
nvcc $FLAGS -c $tmp/whisperx_b200/csrc/wxb_inst.cu -o $tmp/inst.o &
nvcc $FLAGS -c $tmp/whisperx_b200/csrc/wxb_align.cu -o $tmp/align.o &
wait
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o whisperx_b200/_variants/lib_$name.so $tmp/inst.o $tmp/align.o
rm -rf $tmp
echo whisperx_b200/_variants/lib_$name.so
