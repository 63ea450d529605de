#!/bin/bash
# Round profile pass (run under gpurun, one GPU).  Outputs land in gpurun_out/; summaries are copied
# into profiles/ by hand (see profiles/README.md).
#   bash tools/profile_round.sh [mega]      # "mega": also the (slow, ~10 min) full capture of the persistent decode kernel
set -u
O=gpurun_out
mkdir -p $O
RED="--steps 1 --warmup 0 --frames 200 --caption-tokens 32 --answer-tokens 64 --no-cpu-baseline"
# 1. launch list of a reduced bench command (same kernels, same stage order; SCB_NO_GRAPH=1 so that every decode
#    step is an ordinary launch ncu can see), then the same command untimed by ncu
SCB_NO_GRAPH=1 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 30000 --csv \
  --log-file $O/bench_reduced_launches.csv python bench.py $RED > $O/bench_reduced_under_ncu.log 2>&1
timeout 300 python bench.py $RED > $O/bench_reduced.json 2> $O/bench_reduced.err
# 2. one prefill layer: CTA-pair tcgen05 GEMMs + fused attention (full sets, source view)
timeout 300 ncu --set full --clock-control none --import-source on -k regex:linear_tcgen05\|attention_tcgen05 --launch-skip 40 --launch-count 5 \
  -f -o $O/prefill_layer_full python tools/decode_probe.py 1 5800 1 > /dev/null 2>&1
# 3. k-means merge shape and retrieval: per-kernel times + DRAM bytes
timeout 200 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:kmeans\|cosine_topk \
  --csv --log-file $O/kmeans_topk_launches.csv python tools/kbench.py kmeans topk > /dev/null 2>&1
if [ "${1:-}" = "mega" ]; then
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:decode_mega_kernel --launch-skip 2 --launch-count 1 \
    -f -o $O/decode_mega_full python tools/decode_probe.py 4 5800 1 > $O/decode_mega_under_ncu.log 2>&1
fi
ls -la $O | tail -12
