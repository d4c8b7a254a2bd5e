#!/bin/bash
# Round-2 final single-GPU evidence: the whole GPU test suite, smoke, the default bench (both arms), the CLI / Update
# probe, and an ncu capture of the chain kernel K7.
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q --timeout 300 > gpurun_out/final_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/final_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 900 python bench.py > gpurun_out/final_bench.json 2> gpurun_out/final_bench.err; echo "bench rc=$?"
timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/final_bench_ref.json 2> gpurun_out/final_bench_ref.err; echo "ref rc=$?"
bash tools/probes/update_cli_probe.sh > gpurun_out/final_cli.log 2>&1; echo "cli rc=$?"; grep -E "^==|MB/s|ok" gpurun_out/final_cli.log | cut -c1-220
bash tools/probes/k7_ncu.sh
