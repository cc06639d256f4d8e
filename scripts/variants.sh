#!/usr/bin/env bash
# run the variant perf probe against every experimental library under build/
for lib in build/lib_*.so; do CATS_B200_LIB=$lib timeout 200 python scripts/r3_variant_perf.py "$@" 2>&1 | grep -v "^$" | tail -8; done
