#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -40 > gpurun_out/r2g_pytest_all.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/r2g_smoke.txt 2>&1
timeout 600 python bench.py > gpurun_out/r2g_bench.json 2> gpurun_out/r2g_bench.err
tail -5 gpurun_out/r2g_pytest_all.txt; tail -2 gpurun_out/r2g_smoke.txt; cut -c1-400 gpurun_out/r2g_bench.json
