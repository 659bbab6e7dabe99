#!/bin/bash
# round 2, GPU call M: the build with two detector + two LK streams: default bench line (all blocks), launch list,
# ncu --set full of one launch of each kernel, LK track error by skip for both motions
OUT=gpurun_out; mkdir -p $OUT
timeout 900 python bench.py > $OUT/r2m_bench_4k.json 2> $OUT/r2m_bench.err
tail -c 300 $OUT/r2m_bench.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > $OUT/r2m_ref_4k.json 2>> $OUT/r2m_bench.err
timeout 600 python scripts/lk_track_error_by_skip.py > $OUT/r2m_lk_track_error_by_skip.txt 2>> $OUT/r2m_bench.err
cat $OUT/r2m_lk_track_error_by_skip.txt
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file $OUT/r2m_launches.csv \
    python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --no-ba --no-plugin > $OUT/r2m_ncu_b.log 2>&1
python scripts/launch_summary.py $OUT/r2m_launches.csv > $OUT/r2m_launch_summary.txt; cat $OUT/r2m_launch_summary.txt
timeout 900 ncu --set full --clock-control none --cache-control none --import-source on \
    -k "regex:gray_l1_tma|l2_l3_tma|pad_border|min_eig_kernel|nms_candidates|greedy_suppress|compact_top|select_rank|lk10_kernel|lk10_template|lk_compact|pnp_lm|raycast_resident" \
    -s 195 -c 15 -o $OUT/r2m_prof -f python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --no-ba --no-plugin > $OUT/r2m_ncu_full.log 2>&1
for c in 1080p 720p; do
  timeout 600 python bench.py --config $c --no-ba --no-plugin > $OUT/r2m_bench_$c.json 2>> $OUT/r2m_bench.err
done
ls -la $OUT | tail -8
