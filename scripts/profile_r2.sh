#!/bin/bash
# Round-2 profiling pass (run on the GPU box after the un-profiled commands have exited 0): ncu launch lists of the bench
# step in f16 and fp32 mode and of one training update, and `--set full` captures of the dominant kernels.
# Graph replay is switched off for the profiled processes (CNS_NO_GRAPHS=1): ncu then sees plain kernel launches.
cd "$(dirname "$0")/.."
O=gpurun_out
export CNS_NO_GRAPHS=1
FWD='regex:enc_|gemm_tc|edgeconv|pool_head|graphnorm|linear_kernel|prep_|l1_|sorted_|graph_tiles|graph_ptr|backbone|csr_|scan_|frame_mask'
for p in f16 fp32; do
  python bench.py --steps 2 --warmup 1 --precision $p --no-extras --no-cpu-baseline > /dev/null 2>&1 || exit 1
  ncu --metrics gpu__time_duration.sum --clock-control none -k "$FWD" -c 300 --csv --log-file $O/launches_r2_$p.csv \
      python bench.py --steps 2 --warmup 1 --precision $p --no-extras --no-cpu-baseline > /dev/null 2>&1
done
python scripts/train_step.py 256 2 512 0 > $O/train_step.log 2>&1 || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -k 'regex:^(?!.*(ap_|gb_|sel_|segment_reduce|pack_|compose))' -c 1500 --csv \
    --log-file $O/launches_r2_train.csv python scripts/train_step.py 256 2 512 0 > /dev/null 2>&1
CNS_PRECISION=f16 ncu --set full --clock-control none --import-source on -k regex:enc_edge_tc -s 3 -c 1 -o $O/r2_enc_edge_f16 -f python scripts/enc_edge_time.py > /dev/null 2>&1
CNS_PRECISION=fp32 ncu --set full --clock-control none --import-source on -k regex:enc_edge_split -s 3 -c 1 -o $O/r2_enc_edge_split -f python scripts/enc_edge_time.py > /dev/null 2>&1
CNS_PRECISION=f16 ncu --set full --clock-control none --import-source on -k regex:backbone_fused -s 3 -c 1 -o $O/r2_backbone_fused -f python scripts/enc_edge_time.py > /dev/null 2>&1
CNS_PRECISION=fp32 ncu --set full --clock-control none -k regex:gemm_tc_split -s 14 -c 7 -o $O/r2_gemm_tc_split -f python scripts/enc_edge_time.py > /dev/null 2>&1
ls -la $O/*.ncu-rep $O/launches_r2_*.csv
