#!/bin/bash
cd "$(dirname "$0")/.."
timeout 900 python -m pytest tests -m gpu -x -q -k "sched or long_rays or config4 or hygiene or iterations" 2>&1 | tail -3
tools/run_ab.sh "c2 c4 c4share8" base
