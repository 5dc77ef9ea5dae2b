# Round-end evidence run: full GPU test-suite, smoke, the default (inference) bench line, the training bench line.
# (The ncu launch list / --set full captures under profiles/ were taken with tools/ncu_train_step.py:
#   ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file L.csv python tools/ncu_train_step.py w32_coco 32
#   ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:<kernel> -s 12 -c 1 -o R python tools/ncu_train_step.py w32_coco 32)
cd $GRAFT_REPO_ROOT
python -m pytest tests -q -m gpu --timeout 900 2>&1 | tail -2 > gpurun_out/final_tests.log
python __graft_entry__.py smoke 2>&1 | tail -1 >> gpurun_out/final_tests.log
python bench.py --workload train --steps 30 --warmup 3 > gpurun_out/final_train.json 2> gpurun_out/final_train.err
python bench.py > gpurun_out/final_infer.json 2> gpurun_out/final_infer.err
