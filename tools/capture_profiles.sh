#!/bin/bash
# Round artefacts in one GPU call (run from the repo root on the GPU box):
#   1. bench.py (N=1, default flags) and the reference arm           -> gpurun_out/final_bench_n1.json, final_bench_ref.json
#   2. per-kernel event timings of one encode / one decode            -> gpurun_out/final_phases.json, final_decode.json
#   3. ncu launch list of the bench command (per-launch durations)    -> gpurun_out/final_launches.csv
#   4. ncu --set full of the heaviest kernels at full size             -> gpurun_out/final_full_*.ncu-rep
# Numbers printed under ncu are never bench values.
set -x
python bench.py > gpurun_out/final_bench_n1.json 2> gpurun_out/final_bench_n1.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/final_bench_ref.json 2> gpurun_out/final_bench_ref.err
python tools/profile_phases.py > gpurun_out/final_phases.json 2> gpurun_out/final_phases.err
python tools/profile_decode.py > gpurun_out/final_decode.json 2> gpurun_out/final_decode.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/final_launches.csv \
    python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu --no-decode > gpurun_out/final_launches.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on \
    --kernel-name regex:"k_pair_hist_tiles|k_pack_tiles|k_small_groups|k_newline_write|k_newline_count|k_gather_rows32|k_record_stats_names|k_qname_tokens" \
    --launch-count 9 -o gpurun_out/final_full_a -f python tools/profile_phases.py > gpurun_out/final_full_a.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on --kernel-name regex:"k_radix_scatter" --launch-skip 20 --launch-count 1 \
    -o gpurun_out/final_full_b -f python tools/profile_phases.py > gpurun_out/final_full_b.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on --kernel-name regex:"k_decode_tiles" --launch-skip 2 --launch-count 1 \
    -o gpurun_out/final_full_c -f python tools/profile_decode.py > gpurun_out/final_full_c.log 2>&1
ls -la gpurun_out/final_*
