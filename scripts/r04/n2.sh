#!/bin/bash
# Two B200s: the NCCL shard tests and the 2-rank bench lines of configs[1] / [3] / [4]
O=gpurun_out/r04f; mkdir -p $O
python -m pytest tests/test_gpu_shard.py -m gpu -q > $O/tests_n2.log 2>&1; tail -n 2 $O/tests_n2.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541"
$TR bench.py --gpus 2 --no-cpu-baseline > $O/bench_n2.json 2> $O/bench_n2.err
$TR bench.py --gpus 2 --workload c4 > $O/bench_c4_n2.json 2> $O/bench_c4_n2.err
$TR bench.py --gpus 2 --workload c5 > $O/bench_c5_n2.json 2> $O/bench_c5_n2.err
tail -c 300 $O/bench_n2.json | head -c 300; echo
