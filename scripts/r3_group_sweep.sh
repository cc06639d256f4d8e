for g in 1 2 4 8 12 14; do CATS_GROUP=$g python scripts/r3_variant_perf.py s8 2>&1 | sed "s/^/G=$g /"; done
