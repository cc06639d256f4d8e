#!/usr/bin/env bash
# run quick_perf against every experimental library under build/
for lib in build/lib_*.so; do echo "== $lib"; CATS_B200_LIB=$lib timeout 120 python scripts/quick_perf.py "$@" 2>&1 | tail -4; done
