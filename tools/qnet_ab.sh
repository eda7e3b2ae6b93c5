#!/bin/bash
# A/B the Q-network paths: QLC_QNET_IMPL = number of layers on the shifted-window / plane-layout path
# (0 = thread-gathered im2col GEMMs only ... 4 = everything, the default)
for impl in 0 1 2 3 4; do
  echo "== impl $impl"
  QLC_QNET_IMPL=$impl timeout 300 python -m pytest tests/test_qnet.py -x -q -k forward 2>&1 | tail -1
  QLC_QNET_IMPL=$impl timeout 120 python tools/qnet_bench.py 2>&1 | tail -2
done
