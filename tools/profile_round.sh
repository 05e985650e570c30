# Round-2 evidence (GPU box): launch list + ncu --set full captures of the kernels the roofline lines name.
#   bash tools/profile_round.sh        -> gpurun_out/r2_*.csv / *.ncu-rep / *.txt   (summaries are then copied to profiles/)
set -x
CMD="python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-other-arms --no-graph --no-train-step --no-scalable --no-strong --no-config3"
$CMD > gpurun_out/r2_plain.log 2> gpurun_out/r2_plain.err || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/r2_launches_bf16x3.csv $CMD > gpurun_out/r2_ncu_list.log 2>&1
for spec in "k2:^conv_tc_kernel:0" "last:last_scatter_x3_kernel:0" "gdn:^gdn_ts_kernel:0" "lik:gm_likelihood:0" "first:first_fused_x3_kernel:0" "d3:^conv_tc_kernel:15"; do
  IFS=: read name re skip <<< "$spec"
  ncu --set full --clock-control none --import-source on -k regex:$re --launch-skip $skip --launch-count 1 -f -o gpurun_out/r2_ncu_$name $CMD > gpurun_out/r2_ncu_$name.log 2>&1
  python tools/ncu_summary.py gpurun_out/r2_ncu_$name.ncu-rep > gpurun_out/r2_ncu_full_$name.txt 2>&1
done
ls -la gpurun_out/r2_*
