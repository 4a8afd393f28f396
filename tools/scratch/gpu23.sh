cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
export NSIG_NO_SIDE_STREAMS=1 NSIG_DEC_NO_SIDE=1
timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:'^k_(field|march|composite|msg_adam$|msg_table)' -f -o gpurun_out/step_v5_main python tools/ncu_step.py > gpurun_out/ncu_step_main.log 2>&1
timeout 600 ncu --set full --clock-control none --profile-from-start off -k regex:'^k_dec_(conv|wgrad)' --launch-skip 2 -c 4 -f -o gpurun_out/step_v5_dec python tools/ncu_step.py > gpurun_out/ncu_step_dec.log 2>&1
python tools/ncu_summary.py gpurun_out/step_v5_main.ncu-rep > gpurun_out/step_v5_main.txt
python tools/ncu_summary.py gpurun_out/step_v5_dec.ncu-rep > gpurun_out/step_v5_dec.txt
ncu -i gpurun_out/step_v5_main.ncu-rep --page raw --csv > gpurun_out/step_v5_main_raw.csv
ncu -i gpurun_out/step_v5_main.ncu-rep --page source --csv -k regex:k_field_bwd > gpurun_out/step_v5_bwd_source.csv 2>/dev/null
ls -la gpurun_out/ | head -30
rm -f gpurun_out/step_v5.ncu-rep
