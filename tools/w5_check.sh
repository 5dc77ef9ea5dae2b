# scratch driver: training tests + the eager per-shape profile of one step
cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_train_kernels_gpu.py tests/test_train_step_gpu.py -q -x 2>&1 | tail -3 > gpurun_out/w5.log
timeout 600 python tools/bench_train.py > gpurun_out/train_shapes.log 2>&1
