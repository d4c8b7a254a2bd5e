#!/bin/bash
# Round-2 ncu evidence: launch list of the bench step, full captures of every kernel of the path.
cd "$(dirname "$0")/../.."
B="python bench.py --steps 2 --warmup 1 --skip-configs --skip-cpu-baseline --e2e-steps 1"
$B > gpurun_out/plain_bench.log 2>&1 || { echo "plain bench failed"; tail -5 gpurun_out/plain_bench.log; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches.csv $B > gpurun_out/ncu_l.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:decode_blocks_v6 -s 3 -c 1 -o gpurun_out/r02_k1_v6 $B > gpurun_out/ncu_v6.log 2>&1
ncu --set full --clock-control none -k regex:xxh32_frames -s 3 -c 1 -o gpurun_out/r02_k3 $B > gpurun_out/ncu_k3.log 2>&1
# K4 (linked frames), K6 alternative, K5 (size pre-pass), K2 (stored), the Update-path kernels
C="python tools/bench_configs.py 0.25 linked-64KiB"
$C > gpurun_out/plain_c.log 2>&1 && ncu --set full --clock-control none -k regex:decode_chain_pipe -s 1 -c 1 -o gpurun_out/r02_k4 $C > gpurun_out/ncu_k4.log 2>&1
LZ4B200_CHAIN_KERNEL=k6 $C > gpurun_out/plain_c6.log 2>&1 && LZ4B200_CHAIN_KERNEL=k6 ncu --set full --clock-control none -k regex:decode_chain_k6 -s 1 -c 1 -o gpurun_out/r02_k6 $C > gpurun_out/ncu_k6.log 2>&1
D="python tools/bench_configs.py 0.25 4MiB-blocks/random"
$D > gpurun_out/plain_d.log 2>&1 && ncu --set full --clock-control none -k regex:copy_stored -s 1 -c 1 -o gpurun_out/r02_k2 $D > gpurun_out/ncu_k2.log 2>&1
E="python tools/probes/k5_probe.py"
$E > gpurun_out/plain_e.log 2>&1 && ncu --set full --clock-control none -k regex:size_blocks -c 1 -o gpurun_out/r02_k5 $E > gpurun_out/ncu_k5.log 2>&1
F="python tools/stream_probe.py 8 4 1"
$F > gpurun_out/plain_f.log 2>&1 && ncu --set full --clock-control none -k regex:stream_block_kernel -s 20 -c 1 -o gpurun_out/r02_stream_block $F > gpurun_out/ncu_sb.log 2>&1
ls -la gpurun_out/*.ncu-rep
