#!/bin/bash
cd "$(dirname "$0")/.."
timeout 900 python -m pytest tests -m gpu -x -q -k "not fog and not volume and not fullsize" 2>&1 | tail -3
timeout 900 python tools/sweep12.py < /dev/null 2>&1 | tail -30
