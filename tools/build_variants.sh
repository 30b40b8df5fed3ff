#!/bin/bash
# Build alternative libbsq.so variants for A/B runs (tools/kab.py): name=flags pairs on the command line, e.g.
#   tools/build_variants.sh base="-DBSQ_REGION_V1" ldg256="-DBSQ_LDG256"
# Output: variants/libbsq_<name>.so (git-ignored; travels to the GPU box with the snapshot).
set -e
cd "$(dirname "$0")/.."
mkdir -p variants
for kv in "$@"; do
  name="${kv%%=*}"; flags="${kv#*=}"
  ( nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 --shared -Xcompiler -fPIC -Wno-deprecated-declarations $flags \
      -o variants/libbsq_$name.so biscuit_b200/csrc/bsq_align.cu biscuit_b200/csrc/bsq_index_build.cu biscuit_b200/csrc/bsq_pileup.cu && echo "built $name ($flags)" ) &
done
wait
