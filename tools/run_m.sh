#!/bin/bash
cd "$(dirname "$0")/.."
timeout 900 python -m pytest tests -m gpu -x -q -k "long_rays or sched or shared_film or parity" 2>&1 | tail -3
tools/run_ab.sh "c2share8" base
