#!/bin/bash
O=gpurun_out; mkdir -p $O
timeout 600 python -m pytest tests -m gpu -x -q > $O/r1f_pytest.log 2>&1; echo "pytest exit $?" >> $O/r1f_pytest.log; tail -3 $O/r1f_pytest.log
timeout 200 python bench.py --steps 50 --no-cpu-baseline > $O/r1f_bench_n1.json 2> $O/r1f_bench_n1.err; tail -c 300 $O/r1f_bench_n1.json
