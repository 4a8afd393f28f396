#!/bin/bash
cd /root/repo; mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -5
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r02_bench_n1_v4.json 2> gpurun_out/r02_bench_n1_v4.err; tail -2 gpurun_out/r02_bench_n1_v4.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02_bench_ref_v4.json 2>/dev/null; head -c 600 gpurun_out/r02_bench_ref_v4.json
