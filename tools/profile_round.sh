#!/bin/bash
# Round profile pass (run under gpurun, one GPU).  Outputs land in gpurun_out/; summaries are copied
# into profiles/ by hand (see profiles/README.md).
set -u
O=gpurun_out
mkdir -p $O
RED="--steps 1 --warmup 0 --frames 200 --caption-tokens 32 --answer-tokens 64 --no-cpu-baseline"
# 1. launch list of a reduced bench command (same kernels, same stage order), then the same command untimed by ncu
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 20000 --csv \
  --log-file $O/bench_reduced_launches.csv python bench.py $RED > $O/bench_reduced_under_ncu.log 2>&1
timeout 300 python bench.py $RED > $O/bench_reduced.json 2> $O/bench_reduced.err
# 2. DRAM traffic of one decode step (1 sequence, ctx 5800), per kernel
timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none \
  --cache-control none --csv --log-file $O/decode_traffic_r1.csv python tools/decode_probe.py 2 5800 1 > /dev/null 2>&1
# 3. full captures: one prefill layer (GEMMs + attention), one decode layer (GEMVs + attention)
K='regex:gemv_mma|linear_tcgen05|attention_tcgen05|decode_attn_fused'
timeout 300 ncu --set full --clock-control none --import-source on -k "$K" --launch-skip 50 --launch-count 5 \
  -f -o $O/prefill_layer_full python tools/decode_probe.py 2 5800 1 > /dev/null 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k "$K" --launch-skip 200 --launch-count 5 \
  -f -o $O/decode_layer_full python tools/decode_probe.py 2 5800 1 > /dev/null 2>&1
ls -la $O | tail -12
