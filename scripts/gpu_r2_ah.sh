#!/bin/bash
# round 2, GPU call AH: cold ncu --set full of one launch of every analyzer kernel of the final build
OUT=gpurun_out; mkdir -p $OUT
timeout 900 ncu --set full --clock-control none --import-source on \
    -k "regex:gray_l1_tma|l2_l3_tma|pad_border|init_cell_max|min_eig_kernel|nms_candidates|greedy_suppress|compact_top|select_rank|spatial_order|lk10_kernel|lk10_template|lk_compact|pnp_lm|raycast_resident" \
    -s 221 -c 17 -o $OUT/r2ah_prof -f python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --no-ba --no-plugin > $OUT/r2ah_ncu_full.log 2>&1
ls -la $OUT/r2ah_prof.ncu-rep
