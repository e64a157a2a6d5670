#!/bin/bash
O=gpurun_out; mkdir -p $O
for gs in 4 8 2 0; do echo "TSB_GRAPH_STREAMS=$gs"; TSB_GRAPH_STREAMS=$gs timeout 300 python tools/time_phases.py 2>&1 | tail -4; done
echo "TSB_PDL=0"; TSB_PDL=0 timeout 300 python tools/time_phases.py 2>&1 | tail -4
