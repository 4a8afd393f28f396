#!/bin/bash
cd /root/repo; mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -6
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 900 python bench.py > gpurun_out/r02_bench_n1_v3.json 2> gpurun_out/r02_bench_n1_v3.err; tail -c 600 gpurun_out/r02_bench_n1_v3.json; tail -3 gpurun_out/r02_bench_n1_v3.err
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02_bench_ref_v3.json 2>&1; tail -c 400 gpurun_out/r02_bench_ref_v3.json
