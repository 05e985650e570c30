# ncu launch list + one --set full capture of the tensor-core weight-gradient kernel, training step (eager launches) -> gpurun_out/
set -x
CMD="python tools/train_step_times.py 8 1"
$CMD > gpurun_out/train_plain.log 2>&1 || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/train_launches.csv $CMD > gpurun_out/train_ncu_list.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:wgrad_tc_kernel --launch-skip 40 --launch-count 1 -f -o gpurun_out/train_wgrad_tc $CMD > gpurun_out/train_ncu_wgrad.log 2>&1
tail -3 gpurun_out/train_plain.log
ls -la gpurun_out/train_wgrad_tc.ncu-rep
