#!/bin/bash
# ncu capture of the K7 chain kernel on 256 linked frames of 2 MiB (after the same command has run plain)
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
P="python tools/k1_probe.py --size-mib 512 --frame-mib 2 --block 256k --linked --kinds text --tunings 0 --reps 1 --no-block-checksum"
$P > gpurun_out/k7_plain.log 2>&1 || { echo "plain failed"; tail -5 gpurun_out/k7_plain.log; exit 1; }
tail -1 gpurun_out/k7_plain.log | cut -c1-300
ncu --set full --clock-control none --import-source on -k regex:decode_chain_k7 -s 1 -c 1 -o gpurun_out/k7_dev $P > gpurun_out/ncu_k7.log 2>&1
ls -la gpurun_out/k7_dev.ncu-rep
