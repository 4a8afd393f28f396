#!/bin/bash
cd /root/repo; mkdir -p gpurun_out
timeout 300 python tools/graph_offsets.py --out gpurun_out/r02_graph_offsets_all.txt 2>&1 | tail -70
echo ==== few
timeout 300 python tools/graph_offsets.py --names nsig_grad_check_update_scale,nsig_near_far_from_aabb,nsig_field_forward,nsig_msg_adam_step --out gpurun_out/r02_graph_offsets_few.txt 2>&1 | tail -12
