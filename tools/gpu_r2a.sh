#!/bin/bash
# round 2, first GPU pass: GPU parity suite (incl. the device-resident permute cases), smoke(), short bench
O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q > $O/r2a_pytest.log 2>&1; echo "pytest exit $?" >> $O/r2a_pytest.log; tail -5 $O/r2a_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/r2a_smoke.log 2>&1; echo "smoke exit $?" >> $O/r2a_smoke.log; tail -3 $O/r2a_smoke.log
timeout 300 python bench.py --steps 50 --no-cpu-baseline > $O/r2a_bench_n1.json 2> $O/r2a_bench_n1.err; tail -c 400 $O/r2a_bench_n1.json
