#!/bin/bash
# round 2: graph replay + wide-config evidence: GPU suite, bench with the config-4 sub-cube leg, sampled slices
O=gpurun_out; mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -x -q --durations=8 > $O/r2c_pytest.log 2>&1; echo "pytest exit $?" >> $O/r2c_pytest.log; tail -25 $O/r2c_pytest.log
timeout 400 python bench.py --steps 50 --no-cpu-baseline > $O/r2c_bench_n1.json 2> $O/r2c_bench_n1.err; tail -c 1800 $O/r2c_bench_n1.json; tail -5 $O/r2c_bench_n1.err
for c in c2 c4 c5; do timeout 300 python tools/sampled_slices.py --config $c --keep 4 > $O/r2c_sampled_$c.json 2> $O/r2c_sampled_$c.err; tail -c 900 $O/r2c_sampled_$c.json; tail -3 $O/r2c_sampled_$c.err; done
