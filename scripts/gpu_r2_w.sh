#!/bin/bash
# round 2, GPU call W: compute-sanitizer (memcheck, racecheck, synccheck) over the small streaming / schedule tests,
# cold ncu --set full of every analyzer kernel of the final build
OUT=gpurun_out; mkdir -p $OUT
T="tests/test_gpu_analyze.py::test_streaming_analyzer_matches_pairwise tests/test_gpu_analyze.py::test_streaming_with_preset_keypoints_on_the_borders tests/test_gpu_analyze.py::test_lk_work_queue_schedules_agree tests/test_gpu_shard.py::test_sharded_pass_equals_unsharded_pass"
for tool in memcheck racecheck synccheck; do
  timeout 900 compute-sanitizer --tool $tool --error-exitcode 9 --print-limit 20 python -m pytest $T -m gpu -q -x > $OUT/r2w_sanitizer_$tool.log 2>&1
  echo "$tool rc=$?"; grep -E "ERROR SUMMARY|passed|failed|RACECHECK SUMMARY|hazard" $OUT/r2w_sanitizer_$tool.log | tail -4
done
timeout 900 ncu --set full --clock-control none --import-source on \
    -k "regex:gray_l1_tma|l2_l3_tma|pad_border|init_cell_max|min_eig_kernel|nms_candidates|greedy_suppress|compact_top|select_rank|lk10_kernel|lk10_template|lk_compact|pnp_lm|raycast_resident" \
    -s 210 -c 16 -o $OUT/r2w_prof -f python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --no-ba --no-plugin > $OUT/r2w_ncu_full.log 2>&1
ls -la $OUT | tail -5
