mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_train.py -q -p no:cacheprovider ${PYTEST_ARGS:-} > gpurun_out/t_train.log 2>&1; echo "train rc=$?"; tail -n 25 gpurun_out/t_train.log
