#!/bin/bash
# Round artefacts in one GPU call (run from the repo root on the GPU box; bench.py must have exited 0 without ncu before):
#   1. diagnostics of the other configs / components (event timings, no profiler)   -> gpurun_out/r02_*.json
#   2. ncu launch list of the bench command (per-launch durations)                   -> gpurun_out/r02_launches.csv
#   3. ncu --set full of the heaviest kernels at full size                            -> gpurun_out/r02_full_*.ncu-rep
# Numbers printed under ncu are never bench values.
set -x
[ "$1" = "captureonly" ] || {
python tools/profile_phases.py > gpurun_out/r02_phases.json 2> gpurun_out/r02_phases.err
python tools/profile_phases.py 50000000 QNAME casava 100 > gpurun_out/r02_config3.json 2> gpurun_out/r02_config3.err
python tools/profile_phases.py 20000000 None illumina 100 2.2 > gpurun_out/r02_config4_p22.json 2> gpurun_out/r02_config4_p22.err
python tools/profile_ont.py > gpurun_out/r02_ont.json 2> gpurun_out/r02_ont.err
python tools/profile_decode.py > gpurun_out/r02_decode.json 2> gpurun_out/r02_decode.err
python tools/profile_testfeed.py > gpurun_out/r02_testfeed.json 2> gpurun_out/r02_testfeed.err
python tools/profile_sort_adversarial.py > gpurun_out/r02_sort_adversarial.json 2> gpurun_out/r02_sort_adversarial.err
python tools/bench_layouts.py > gpurun_out/r02_layouts.json 2> gpurun_out/r02_layouts.err
}
[ "$1" = "nocapture" ] && exit 0
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/r02_launches.csv \
    python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu --no-decode > gpurun_out/r02_launches.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on \
    --kernel-name regex:"k_pair_hist_pipe|k_pack_tiles|k_small_groups_packed|k_newline_scan1|k_gather_rows32|k_record_stats_names|k_qname_tokens|k_pairs_apply|k_pairs_regroup|k_gather_narrow|k_gather_items" \
    --launch-count 14 -o gpurun_out/r02_full_a -f python tools/profile_phases.py > gpurun_out/r02_full_a.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on --kernel-name regex:"k_radix_scatter|k_radix_hist" --launch-skip 40 --launch-count 2 \
    -o gpurun_out/r02_full_b -f python tools/profile_phases.py > gpurun_out/r02_full_b.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on --kernel-name regex:"k_decode_tiles" --launch-skip 2 --launch-count 1 \
    -o gpurun_out/r02_full_c -f python tools/profile_decode.py > gpurun_out/r02_full_c.log 2>&1
ls -la gpurun_out/r02_*
