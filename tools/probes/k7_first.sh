mkdir -p gpurun_out
timeout 400 python -m pytest tests -m gpu -x -q --timeout 150 -k "chain_kernel" > gpurun_out/t_k7.log 2>&1; echo rc=$? >> gpurun_out/t_k7.log
tail -3 gpurun_out/t_k7.log
P="python tools/k1_probe.py --size-mib 512 --frame-mib 2 --block 256k --linked --kinds text --tunings 0 --reps 1"
echo "== k7 linked 2 MiB x256, no block checksum"; LZ4B200_K7_DEBUG=1 timeout 120 $P --no-block-checksum 2>&1 | tail -2 | cut -c1-520
