#!/bin/bash
# Build tuning variants of the CUDA library: tools/build_variants.sh "BK:MINB BK:MINB ..."
# -> benchpush_b200/csrc/variants/lib_k<BK>_b<MINB>.so   (development aid; *.so is git-ignored)
cd "$(dirname "$0")/../benchpush_b200/csrc" || exit 1
mkdir -p variants
for v in $1; do
  k=${v%%:*}; b=${v##*:}
  make -B BK=$k MINB=$b OUT=variants/lib_k${k}_b${b}.so > variants/build_k${k}_b${b}.log 2>&1 &
done
wait
grep -H "k_shipice_step" -A2 variants/build_*.log | grep -E "registers" 
ls -la variants/*.so
