#!/bin/bash
# final evidence of the round with the final tree
cd /root/repo; mkdir -p gpurun_out
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r02_bench_n1_v5.json 2> gpurun_out/r02_bench_n1_v5.err; tail -c 300 gpurun_out/r02_bench_n1_v5.json | head -c 300; echo
timeout 300 python tools/profile_step.py --out gpurun_out/r02_step_breakdown_v3.txt > gpurun_out/profile_step.log 2>&1; head -12 gpurun_out/r02_step_breakdown_v3.txt
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/r02_launches_v3.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-render --no-extra --no-graph > gpurun_out/launches_bench.log 2>&1; tail -2 gpurun_out/launches_bench.log | cut -c1-200
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:'k_dec_conv_tc|k_dec_wgrad|k_msg_adam_sum' -o gpurun_out/r02_step_dec_v3 -f python tools/ncu_step.py > gpurun_out/ncu_step.log 2>&1; tail -2 gpurun_out/ncu_step.log
timeout 300 python tools/graph_offsets.py --out gpurun_out/r02_graph_offsets_v3.txt > /dev/null 2>&1; head -22 gpurun_out/r02_graph_offsets_v3.txt
NSIG_LIB=tools/scratch/libs/libnsig_trace.so NSIG_DEC_PDL=0 timeout 300 python tools/dec_trace.py > gpurun_out/r02_decoder_phase_trace_tc.txt 2>&1; tail -32 gpurun_out/r02_decoder_phase_trace_tc.txt | cut -c1-140
