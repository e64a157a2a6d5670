#!/bin/bash
O=gpurun_out; mkdir -p $O
timeout 100 python -m pytest tests/test_gpu_parity.py -x -q -k "first_model or contract_from_host or slice_blocks" > $O/r1j_pytest_subset.log 2>&1; tail -1 $O/r1j_pytest_subset.log
timeout 60 python bench.py --steps 30 --no-cpu-baseline > $O/r1j_bench_n1.json 2> $O/r1j_bench_n1.err; tail -c 250 $O/r1j_bench_n1.json
