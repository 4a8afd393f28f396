#!/bin/bash
cd /root/repo; mkdir -p gpurun_out
for c in 1 2; do
NSIG_ADAM_CTAS_PER_SM=$c timeout 300 python tools/graph_offsets.py --out gpurun_out/r02_graph_offsets_lookahead_adam$c.txt 2>&1 | tail -19
done
