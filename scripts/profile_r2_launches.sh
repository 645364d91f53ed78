#!/bin/bash
# Launch lists only (final code): the bench step in f16 mode and the servo (e2e) path. Graph replay off for the
# profiled processes (CNS_NO_GRAPHS=1) so ncu sees plain kernel launches; each command first exits 0 without ncu.
cd "$(dirname "$0")/.."
O=gpurun_out
export CNS_NO_GRAPHS=1
FWD='regex:enc_|gemm_tc|edgeconv|pool_head|graphnorm|linear_kernel|prep_|l1_|sorted_|graph_tiles|graph_ptr|backbone|csr_|scan_|frame_|mark_visible'
python bench.py --steps 2 --warmup 1 --precision f16 --no-extras --no-cpu-baseline > /dev/null 2>&1 || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -k "$FWD" -c 300 --csv --log-file $O/launches_r2_f16.csv \
    python bench.py --steps 2 --warmup 1 --precision f16 --no-extras --no-cpu-baseline > /dev/null 2>&1
python scripts/servo_host_probe.py > /dev/null 2>&1 || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -k "$FWD" -s 100 -c 400 --csv --log-file $O/launches_r2_e2e.csv \
    python scripts/servo_host_probe.py > /dev/null 2>&1
ls -la $O/launches_r2_*.csv
