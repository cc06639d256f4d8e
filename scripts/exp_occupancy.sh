for lib in build/lib_*.so; do echo "== $lib"; CATS_B200_LIB=$lib timeout 120 python scripts/quick_perf.py 16384 1 2>&1 | tail -1; done
