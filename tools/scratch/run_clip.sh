cd /root/repo
timeout 900 python -m pytest tests/test_gpu_linear.py tests/test_gpu_fullwidth.py tests/test_gpu_models.py tests/test_gpu_native_entry_points.py -x -q 2>&1 | tail -5
timeout 300 python tools/clip_gemm_bench.py 100 2>&1 | grep gemm
for d in 0 1; do SCB_LINEAR_TMA_STORE=$d timeout 300 python tools/model_bench.py 2>&1 | tail -2 | sed "s/^/TMA_STORE=$d /"; done
