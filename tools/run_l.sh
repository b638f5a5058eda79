#!/bin/bash
cd "$(dirname "$0")/.."
timeout 600 python tools/sweep15.py < /dev/null 2>&1 | tail -14
