#!/bin/bash
# Eight B200s: the bench line of configs[1] (short form: no CPU baseline, no matching table, no batch sweep)
O=gpurun_out/r04f; mkdir -p $O
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29551 bench.py --gpus 8 --no-cpu-baseline --no-matching --no-sweep > $O/bench_n8.json 2> $O/bench_n8.err
tail -c 400 $O/bench_n8.json
