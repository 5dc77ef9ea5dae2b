# scratch driver: compute-sanitizer memcheck over the training conv / wgrad parity cases
cd $GRAFT_REPO_ROOT
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 7 --print-limit 5 python -m pytest tests/test_train_kernels_gpu.py -q -x -k "conv_forward and 0-0.004" > gpurun_out/memcheck.log 2>&1
echo "rc=$?" >> gpurun_out/memcheck.log
