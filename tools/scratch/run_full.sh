cd /root/repo
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
DECODE_BENCH_GRAPH_ONLY=1 timeout 300 python tools/decode_bench.py 5800 25 3 2>&1 | grep "R="
timeout 300 python tools/model_bench.py 2>&1 | tail -2
timeout 900 python bench.py > gpurun_out/bench_r02_h.json 2> gpurun_out/bench_r02_h.err; cut -c1-300 gpurun_out/bench_r02_h.json
