#!/bin/bash
# Round artefacts in one GPU call (run from the repo root on the GPU box; bench.py must have exited 0 without ncu before):
#   1. ncu launch list of the bench command (per-launch durations)    -> gpurun_out/r02_launches.csv
#   2. ncu --set full of the heaviest kernels at full size             -> gpurun_out/r02_full_*.ncu-rep
# Numbers printed under ncu are never bench values.
set -x
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/r02_launches.csv \
    python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu --no-decode > gpurun_out/r02_launches.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on \
    --kernel-name regex:"k_pair_hist_pipe|k_pack_tiles|k_small_groups_packed|k_newline_write|k_newline_count|k_gather_rows32|k_record_stats_names|k_qname_tokens|k_scatter_key_windows|k_gather_narrow" \
    --launch-count 12 -o gpurun_out/r02_full_a -f python tools/profile_phases.py > gpurun_out/r02_full_a.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on --kernel-name regex:"k_radix_scatter_key|k_radix_hist" --launch-skip 40 --launch-count 2 \
    -o gpurun_out/r02_full_b -f python tools/profile_phases.py > gpurun_out/r02_full_b.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on --kernel-name regex:"k_decode_tiles" --launch-skip 2 --launch-count 1 \
    -o gpurun_out/r02_full_c -f python tools/profile_decode.py > gpurun_out/r02_full_c.log 2>&1
ls -la gpurun_out/r02_*
