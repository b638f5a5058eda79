#!/bin/bash
cd "$(dirname "$0")/.."
timeout 900 python -m pytest tests -m gpu -x -q -k "fog or volume or cli or kats" 2>&1 | tail -3
for v in base prev; do
  echo "== $v"
  if [ "$v" = base ]; then lib=""; else lib="$PWD/openvdb_b200/variants/libvdbrt_$v.so"; fi
  VDBRT_LIBRARY=$lib timeout 600 python tools/fog_ab.py c3 c5 < /dev/null 2>&1 | grep -v "wave 0\|1/8" | tail -6
done
