#!/bin/bash
# final evidence of the round with the final tree: step breakdown, launch list of the bench command, ncu --set full of one step
cd /root/repo; mkdir -p gpurun_out
timeout 300 python tools/profile_step.py --out gpurun_out/r02_step_breakdown_v2.txt > gpurun_out/profile_step.log 2>&1; head -8 gpurun_out/r02_step_breakdown_v2.txt
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/r02_launches_v2.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-render --no-extra --no-graph > gpurun_out/launches_bench.log 2>&1; tail -2 gpurun_out/launches_bench.log | cut -c1-200
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:'k_field|k_march|k_composite|k_msg|k_grad|k_flat' -o gpurun_out/r02_step_main_v2 -f python tools/ncu_step.py > gpurun_out/ncu_step.log 2>&1; tail -2 gpurun_out/ncu_step.log
timeout 300 python tools/graph_offsets.py --out gpurun_out/r02_graph_offsets_final.txt > /dev/null 2>&1; head -20 gpurun_out/r02_graph_offsets_final.txt
