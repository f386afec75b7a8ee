mkdir -p gpurun_out
for i in 1 2; do python tools/gemm_bench.py --only conv4; MSD_DBG_BO=256 python tools/gemm_bench.py --only conv4; done 2>&1 | tee gpurun_out/exp_dx0.txt
