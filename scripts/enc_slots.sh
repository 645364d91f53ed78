#!/bin/bash
cd "$(dirname "$0")/.."
for sl in 4 5 6; do
  echo "== CNS_TC_SLOTS=$sl"
  CNS_TC_SLOTS=$sl CNS_PRECISION=f16 python scripts/enc_edge_time.py 2>&1 | tail -1
  CNS_TC_SLOTS=$sl CNS_PROBE_PRECISION=f16 python scripts/split_probe.py 512 300 33 7 4 700 2>&1 | tail -1
done
