# Round-end evidence run: full GPU test-suite, smoke, the training bench line, the ncu launch list of one eager training
# step and `ncu --set full` captures of the two kernels the training roofline names (each after the plain run exited 0).
set -x
cd $GRAFT_REPO_ROOT
python -m pytest tests -q -m gpu --timeout 900 2>&1 | tail -2 > gpurun_out/final_tests.log
python __graft_entry__.py smoke 2>&1 | tail -1 >> gpurun_out/final_tests.log
python bench.py --workload train --steps 30 --warmup 3 > gpurun_out/final_train.json 2> gpurun_out/final_train.err
python tools/ncu_train_step.py w32_coco 32 > gpurun_out/plain.log 2>&1 || exit 1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r2_train_launches_final.csv python tools/ncu_train_step.py w32_coco 32 > gpurun_out/ncu_list.log 2>&1
for k in conv3x3_tf32_small_kernel wgrad_tc5_kernel; do
  timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:$k -s 12 -c 1 -f -o gpurun_out/final_$k python tools/ncu_train_step.py w32_coco 32 > gpurun_out/ncu_$k.log 2>&1
done
ls -la gpurun_out/*.ncu-rep
