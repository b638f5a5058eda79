#!/bin/bash
# compute-sanitizer memcheck over the smoke run and the long-ray tests (suspend / resume, rounds)
cd "$(dirname "$0")/.."
timeout 600 compute-sanitizer --tool memcheck --print-limit 5 python __graft_entry__.py smoke 2>&1 | tail -6
timeout 900 compute-sanitizer --tool memcheck --print-limit 5 python -m pytest tests/test_gpu_long_rays.py -x -q 2>&1 | tail -6
