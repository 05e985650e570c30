# ncu launch list of the training step (tools/train_step_times.py: 3 warm-up + 1 + 1 steps) -> gpurun_out/
set -x
CMD="python tools/train_step_times.py 8 1"
$CMD > gpurun_out/train_plain.log 2>&1 || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/train_launches.csv $CMD > gpurun_out/train_ncu_list.log 2>&1
tail -2 gpurun_out/train_plain.log
