for h in 0 1 2 3 5 7; do echo "hints $h"; LZ4B200_V6_HINTS=$h timeout 300 python tools/k1_probe.py --size-mib 4096 --tunings 60 --reps 2 2>&1 | tail -1; done
