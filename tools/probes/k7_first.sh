mkdir -p gpurun_out
timeout 300 python -m pytest tests -m gpu -x -q --timeout 120 -k "chain_kernel_variants and k7" > gpurun_out/t_k7.log 2>&1; echo rc=$? >> gpurun_out/t_k7.log
tail -5 gpurun_out/t_k7.log
P="python tools/k1_probe.py --size-mib 512 --frame-mib 2 --block 256k --linked --kinds text --tunings 0 --reps 1"
echo "== k7 no block checksum"; LZ4B200_K7_DEBUG=1 timeout 120 $P --no-block-checksum 2>&1 | tail -3 | cut -c1-400
