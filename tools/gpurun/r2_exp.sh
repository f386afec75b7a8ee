mkdir -p gpurun_out
for v in 16 6 3 2; do echo "== bwd1 per_sm $v"; MSD_BWD1_PER_SM=$v timeout 300 python tools/bn_bench.py --iters 10 2>&1 | grep -E "bn_act_bwd"; done | tee gpurun_out/exp_bwd1_grid.txt
for v in 4 2 1; do echo "== colsum per_sm $v"; MSD_COLSUM_PER_SM=$v timeout 300 python tools/bn_bench.py --iters 10 2>&1 | grep -E "bn_sums"; done | tee -a gpurun_out/exp_bwd1_grid.txt
