cd $GRAFT_REPO_ROOT
run() { timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $1 bench.py --gpus 2 --steps 30 --warmup 5 --no-render 2>gpurun_out/err_$1.log | tail -1 | python -c "
import sys,json; d=json.loads(sys.stdin.read()); print(sys.argv[1], 'ms/step', round(d['ms_per_step'],4), 'e2e', round(d['e2e']['ms_per_step'],4), 'spr', round(d['config']['mean_samples_per_ray'],1), d['config'].get('exchange'))" "$2"; grep nsig gpurun_out/err_$1.log | head -3; }
NSIG_DIAG_SAME_RAYS=1 run 29522 "same_rays_kernel_exchange"
NSIG_DIAG_SAME_RAYS=1 NSIG_AR_NCCL=1 run 29523 "same_rays_nccl"
run 29525 "default_kernel_exchange"
NSIG_AR_NCCL=1 run 29526 "default_nccl"
