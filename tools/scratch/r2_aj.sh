#!/bin/bash
cd /root/repo; mkdir -p gpurun_out
N="nsig_march_rays_train,nsig_near_far_from_aabb,nsig_field_forward,nsig_field_backward_masks,nsig_decoder_forward,nsig_decoder_backward,nsig_msg_adam_step,nsig_msg_table_sum,nsig_composite_train_blend_fwd,nsig_composite_train_blend_bwd,nsig_wtmk_loss_fwd"
timeout 600 python tools/graph_offsets.py --march-ahead 0 --out gpurun_out/off_ma0.txt > /dev/null 2>gpurun_out/off0.err; tail -2 gpurun_out/off0.err
timeout 600 python tools/graph_offsets.py --march-ahead 1 --out gpurun_out/off_ma1.txt > /dev/null 2>gpurun_out/off1.err; tail -2 gpurun_out/off1.err
