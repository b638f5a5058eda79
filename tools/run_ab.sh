#!/bin/bash
# usage: tools/run_ab.sh "<ab_lib args>" variant...   ("base" = the regular library)
cd "$(dirname "$0")/.."
args=$1; shift
for v in "$@"; do
  echo "== $v"
  if [ "$v" = base ]; then lib=""; else lib="$PWD/openvdb_b200/variants/libvdbrt_$v.so"; fi
  VDBRT_LIBRARY=$lib timeout 600 python tools/ab_lib.py $args < /dev/null 2>&1 | tail -8
done
