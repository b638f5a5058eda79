#!/bin/bash
cd "$(dirname "$0")/.."
VDBRT_DEBUG_TILES=1 timeout 600 python tools/diag_phases.py < /dev/null 2>&1 | tail -12
