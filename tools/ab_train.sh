# A/B of two builds of the library on the same box: gpurun_ab/lib_old.so vs gpurun_ab/lib_new.so, alternating
# (boxes differ by 1-3 %, more than most single changes).  Build both here, then `gpurun -- bash tools/ab_train.sh`.
cd $GRAFT_REPO_ROOT
cp rsgnet_b200/librsg_b200.so /tmp/lib_keep.so
for r in 1 2; do
  for v in old new; do
    cp gpurun_ab/lib_$v.so rsgnet_b200/librsg_b200.so
    python bench.py --workload train --steps 20 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('$v', round(d['value'],1), round(d['ms_per_step'],3), d['clocks']['sm_mhz'])"
  done
done > gpurun_out/ab.log 2>&1
cp /tmp/lib_keep.so rsgnet_b200/librsg_b200.so
timeout 600 python -m pytest tests/test_train_kernels_gpu.py -q -x 2>&1 | tail -1 >> gpurun_out/ab.log
timeout 300 python tools/bench_train_gemm.py 4 5 2>&1 | tail -2 >> gpurun_out/ab.log
