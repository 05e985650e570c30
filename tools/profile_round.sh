set -x
CMD="python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-other-arms --no-graph"
$CMD > gpurun_out/r36_plain.log 2> gpurun_out/r36_plain.err || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/r36_launches_x3.csv $CMD > gpurun_out/r36_ncu_list.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:^conv_tc_kernel --launch-count 1 -f -o gpurun_out/r36_k2_x3 $CMD > gpurun_out/r36_ncu_k2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:^gdn_x3_kernel --launch-count 1 -f -o gpurun_out/r36_gdn_x3 $CMD > gpurun_out/r36_ncu_gdn.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:gm_likelihood --launch-count 1 -f -o gpurun_out/r36_lik $CMD > gpurun_out/r36_ncu_lik.log 2>&1
CMDB="python bench.py --precision bf16 --steps 2 --warmup 1 --no-cpu-baseline --no-other-arms --no-graph"
$CMDB > gpurun_out/r36_plain_bf16.log 2> gpurun_out/r36_plain_bf16.err || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/r36_launches_bf16.csv $CMDB > gpurun_out/r36_ncu_list_bf16.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:conv_tc_kernel --launch-count 1 -f -o gpurun_out/r36_k2_bf16 $CMDB > gpurun_out/r36_ncu_k2_bf16.log 2>&1
ls -la gpurun_out/r32*
