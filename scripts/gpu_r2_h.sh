#!/bin/bash
# round 2, GPU call H: baseline of HEAD after the container was re-created: GPU suite, default bench, launch list
# (duration only), ncu --set full of one launch of every analyzer kernel
OUT=gpurun_out; mkdir -p $OUT
timeout 1200 python -m pytest tests -m gpu -q --deselect tests/test_gpu_multi.py 2>&1 | tail -25 > $OUT/r2h_tests.log
tail -6 $OUT/r2h_tests.log
timeout 900 python bench.py > $OUT/r2h_bench_4k.json 2> $OUT/r2h_bench.err
tail -c 300 $OUT/r2h_bench.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file $OUT/r2h_launches.csv \
    python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --no-ba --no-plugin > $OUT/r2h_ncu_b.log 2>&1
python scripts/launch_summary.py $OUT/r2h_launches.csv > $OUT/r2h_launch_summary.txt; cat $OUT/r2h_launch_summary.txt
timeout 900 ncu --set full --clock-control none --import-source on \
    -k "regex:gray_l1_tma|l2_l3_tma|pad_border|min_eig_kernel|nms_candidates|greedy_suppress|compact_top|select_rank|lk10_kernel|lk10_template|lk_compact|pnp_lm|raycast_resident" \
    -s 195 -c 15 -o $OUT/r2h_prof -f python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --no-ba --no-plugin > $OUT/r2h_ncu_full.log 2>&1
ls -la $OUT | tail -8
