#!/bin/bash
# ncu --set full of ONE kernel of a bench step (never a bench value). usage: ncu_one.sh <kernel regex> <tag> [skip] [count]
O=gpurun_out/r04; mkdir -p $O
ncu --set full --clock-control none --import-source on -k regex:$1 -s ${3:-8} -c ${4:-1} -f -o $O/$2 python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-sweep --no-matching > $O/$2_ncu.log 2>&1
ls -la $O/$2.ncu-rep
