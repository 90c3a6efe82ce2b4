#!/bin/bash
# compute-sanitizer memcheck + racecheck of the final build on small problems (scripts/exp/sanitize_small.py)
O=gpurun_out/r04f; mkdir -p $O
timeout 900 compute-sanitizer --tool memcheck python scripts/exp/sanitize_small.py > $O/sanitizer_memcheck.log 2>&1; tail -n 2 $O/sanitizer_memcheck.log
timeout 900 compute-sanitizer --tool racecheck python scripts/exp/sanitize_small.py > $O/sanitizer_racecheck.log 2>&1; tail -n 2 $O/sanitizer_racecheck.log
