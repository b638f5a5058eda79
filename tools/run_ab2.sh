#!/bin/bash
cd "$(dirname "$0")/.."
timeout 900 python -m pytest tests -m gpu -x -q -k "not fog and not volume and not fullsize" 2>&1 | tail -3
tools/run_ab.sh "c2 c4 c2share8 c4share8" base prev
