#!/bin/bash
# per-kernel launch times of one Q-network forward (ncu serialises and cold-caches: use the SHARES)
N=${1:-4096}
timeout 300 python tools/qnet_once.py $N > gpurun_out/qnet_plain.log 2>&1 || exit 1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/qnet_launches.csv python tools/qnet_once.py $N > gpurun_out/qnet_ncu.log 2>&1
python - <<'PY'
import csv
lines = open('gpurun_out/qnet_launches.csv').read().splitlines()
i = [k for k, l in enumerate(lines) if l.startswith('"ID"')][0]
rows = list(csv.DictReader(lines[i:]))
for r in rows[-9:]:
    print("%-90s grid %-14s block %-12s %10.1f us" % (r['Kernel Name'][:90], r['Grid Size'], r['Block Size'], float(r['Metric Value'].replace(',', '')) / 1e3))
PY
