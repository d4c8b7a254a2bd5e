cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q --timeout 300 -k "update or Update or read_ahead or stream or cli or vector or lifetime or k3 or xxh" > gpurun_out/t_upd.log 2>&1; echo rc=$? >> gpurun_out/t_upd.log
tail -3 gpurun_out/t_upd.log
bash tools/probes/update_dbg.sh 2>&1 | grep -v "lz4ada update"
