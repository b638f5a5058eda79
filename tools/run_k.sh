#!/bin/bash
cd "$(dirname "$0")/.."
for v in base long5; do
  echo "== $v"
  if [ "$v" = base ]; then lib=""; else lib="$PWD/openvdb_b200/variants/libvdbrt_$v.so"; fi
  VDBRT_LIBRARY=$lib timeout 600 python tools/sweep13.py < /dev/null 2>&1 | tail -9
done
