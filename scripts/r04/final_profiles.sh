#!/bin/bash
# Round-end evidence on ONE B200: gpu tests, smoke, the bench lines of every workload, the reference arm, the ncu launch list of a
# bench run and one ncu --set full capture of every kernel of a step (+ the one-launch matcher).
O=gpurun_out/r04f; mkdir -p $O
python -m pytest tests -m gpu -q > $O/tests.log 2>&1; tail -n 2 $O/tests.log
python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; tail -n 1 $O/smoke.log
python bench.py > $O/bench_n1.json 2> $O/bench_n1.err
python bench.py --impl reference --steps 3 --warmup 1 > $O/bench_ref.json 2> $O/bench_ref.err
for w in c1 c3 c4 c5; do python bench.py --workload $w > $O/bench_$w.json 2> $O/bench_$w.err; done
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-sweep --no-matching > $O/ncu_launch_bench.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:^k_ -s 50 -c 26 -f -o $O/all_full python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-sweep --no-matching > $O/ncu_all.log 2>&1
ls -la $O
