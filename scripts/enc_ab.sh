#!/bin/bash
# A-B timing of the encoder edge kernels on the bench workload (f16 and fp32 modes) + a parity spot check
cd "$(dirname "$0")/.."
for p in f16 fp32; do CNS_PRECISION=$p python scripts/enc_edge_time.py 2>&1 | tail -1; done
python scripts/split_probe.py 512 300 33 2>&1 | tail -1
CNS_PROBE_PRECISION=f16 python scripts/split_probe.py 512 300 33 2>&1 | tail -1
