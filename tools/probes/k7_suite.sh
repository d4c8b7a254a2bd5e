mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q --timeout 300 > gpurun_out/t_all.log 2>&1; echo rc=$? >> gpurun_out/t_all.log
tail -6 gpurun_out/t_all.log
for a in "32 7 0" "32 4 1" "32 6 1"; do echo "== stream_probe $a"; timeout 200 python tools/stream_probe.py $a 2>&1 | tail -4; done
echo "== stream_probe 32 7 0 (K7 off)"; LZ4B200_STREAM_K7=0 timeout 200 python tools/stream_probe.py 32 7 0 2>&1 | tail -2
