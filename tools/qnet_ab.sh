#!/bin/bash
# A/B the shifted-window conv kernels: QLC_QNET_IMPL = number of convs on the new path (0 = thread-gather GEMMs only)
for impl in 3 4; do
  echo "== impl $impl"
  QLC_QNET_IMPL=$impl timeout 300 python -m pytest tests/test_qnet.py -x -q -k forward 2>&1 | tail -4
done
for impl in 3 4; do
  echo "== bench impl $impl"
  QLC_QNET_IMPL=$impl timeout 120 python tools/qnet_bench.py 2>&1 | tail -4
done
