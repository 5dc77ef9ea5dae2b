set -x
cd $GRAFT_REPO_ROOT
python tools/ncu_train_step.py w32_coco 32 > gpurun_out/plain.log 2>&1 || exit 1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r2_train_launches_final.csv python tools/ncu_train_step.py w32_coco 32 > gpurun_out/ncu_list.log 2>&1
for k in conv3x3_tf32_small_kernel conv3x3_tf32_tc5_kernel wgrad_kernel wgrad32_flat_kernel gemm_tf32_tc5_kernel; do
  timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:^$k -s 12 -c 1 -f -o gpurun_out/final_$k python tools/ncu_train_step.py w32_coco 32 > gpurun_out/ncu_$k.log 2>&1
done
ls -la gpurun_out/*.ncu-rep
