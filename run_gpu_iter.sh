mkdir -p gpurun_out
timeout 600 python -m pytest ${PYTEST_FILES:-tests/test_gpu_scene.py} -q -p no:cacheprovider ${PYTEST_ARGS:-} > gpurun_out/t_iter.log 2>&1; echo "rc=$?"; tail -n 30 gpurun_out/t_iter.log
