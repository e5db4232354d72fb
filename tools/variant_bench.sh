#!/bin/bash
# GPU box: time every prebuilt variant (tools/build_variants.py) on the bench workloads; hit buffers are compared with
# the FIRST variant's.   tools/variant_bench.sh [cases...]
cd "$(dirname "$0")/.."
LIB=bvhtracer_b200/lib
cp $LIB/libbvht_cuda.so /tmp/libbvht_cuda.default.so
rm -rf /tmp/variant_ref; mkdir -p /tmp/variant_ref
for so in ${VARIANTS:-$(ls $LIB/variants/*.so)}; do
  name=$(basename $so .so); name=${name#libbvht_cuda_}
  cp $so $LIB/libbvht_cuda.so
  echo "== $name"
  python tools/variant_bench.py "$@" 2>&1 | tail -n 12
done
cp /tmp/libbvht_cuda.default.so $LIB/libbvht_cuda.so
