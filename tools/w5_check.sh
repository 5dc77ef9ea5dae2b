# scratch driver for the tcgen05 weight-gradient kernel: parity cases, micro-benchmark, the training bench line
cd $GRAFT_REPO_ROOT
{
timeout 600 python -m pytest tests/test_train_kernels_gpu.py -q -x -k "conv_forward" -s 2>&1 | grep -E "precise=0|passed|failed|rror" | tail -30
timeout 300 python tools/bench_train_gemm.py 1 2 3 6 7 9 10 2>&1 | tail -7
timeout 600 python bench.py --workload train --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/w5_bench.json 2> gpurun_out/w5_bench.err
python -c "
import json; d=json.load(open('gpurun_out/w5_bench.json')); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['gpu_launches']/20, d.get('parity'))"
} > gpurun_out/w5.log 2>&1
