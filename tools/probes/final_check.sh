cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q --timeout 300 > gpurun_out/final_tests.log 2>&1; echo "tests rc=$?"; tail -2 gpurun_out/final_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
bash tools/probes/update_cli_probe.sh > gpurun_out/final_cli.log 2>&1; echo "cli rc=$?"; grep -E "^==|MB/s|ok" gpurun_out/final_cli.log | grep -E "keep|update:|ok" | cut -c1-170 | head -40
