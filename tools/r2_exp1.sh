#!/bin/bash
# Scratch experiment batch (round 2): exit-wait variant A/B, batch shapes for single-step launches, host gather breakdown, microbenchmarks.
cd "$(dirname "$0")/.."
L=$PWD/q-learning_b200
child() { python tools/step_latency_ab.py child "$@"; }
echo "== single step, default lib (wait_group.read at exit)"
for n in 4096 65536; do echo -n "n=$n: "; child $n $((n*16)); done
echo -n "n=4096 cap 1M: "; child 4096 1048576
echo -n "n=4096 QLC_EPC=14: "; QLC_EPC=14 child 4096 1048576
echo -n "n=4096 QLC_EPC=12: "; QLC_EPC=12 child 4096 1048576
echo -n "n=4096 cfg5: "; QLC_ADVANCE_CFG=5 child 4096 1048576
echo -n "n=4096 cfg5 EPC=7: "; QLC_ADVANCE_CFG=5 QLC_EPC=7 child 4096 1048576
echo "== single step, full wait at exit"
for n in 4096 65536; do echo -n "n=$n: "; QLC_LIB=$L/libqlcuda_waitfull.so child $n $((n*16)); done
echo "== parity subset with the default lib"
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_round2.py -m gpu -x -q 2>&1 | tail -3
echo "== host gather breakdown"
python tools/host_gather_breakdown.py 32
echo "== widen rate"
tools/microbench/widen_rate
QLC_HOST_THREADS=8 tools/microbench/widen_rate
echo "== mma rate"
tools/microbench/mma_rate
echo "== r2_latency default"
python tools/r2_latency.py
echo "== r2_latency QLC_GATHER_SLICES=2"
QLC_GATHER_SLICES=2 python tools/r2_latency.py 2>&1 | head -4
