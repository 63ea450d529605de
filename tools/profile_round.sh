#!/bin/bash
# Round profile pass (run under gpurun, one GPU).  Outputs land in gpurun_out/; summaries are copied
# into profiles/ by hand (see profiles/README.md).
#   bash tools/profile_round.sh [mega]      # "mega": also the (slow) full captures of the persistent decode kernel
set -u
O=gpurun_out
mkdir -p $O
RED="--steps 1 --warmup 0 --frames 200 --caption-tokens 32 --answer-tokens 64 --no-cpu-baseline"
# 1. launch list of a reduced bench command (same kernels, same stage order; SCB_NO_GRAPH=1 so that every decode
#    step is an ordinary launch ncu can see), then the same command untimed by ncu
SCB_NO_GRAPH=1 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 40000 --csv \
  --log-file $O/bench_reduced_launches.csv python bench.py $RED > $O/bench_reduced_under_ncu.log 2>&1
timeout 300 python bench.py $RED > $O/bench_reduced.json 2> $O/bench_reduced.err
python tools/ncu_summary.py $O/bench_reduced_launches.csv > $O/bench_reduced_summary.txt 2>&1
# 2. encode pass (100 frames: the bench's pass size) + one prefill: per-kernel times, tensor-pipe activity, DRAM bytes
timeout 600 ncu --metrics gpu__time_duration.sum,sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_tensor.sum,dram__bytes_read.sum,dram__bytes_write.sum -k regex:linear_tcgen05\|attention_tcgen05\|norm_\|im2col\|pool2x2\|rope_append\|gemv_mma \
  --clock-control none -c 1200 --csv --log-file $O/encode_prefill_kernels.csv python tools/model_bench.py ncu > $O/model_bench_under_ncu.log 2>&1
# 3. k-means merge shape and retrieval: per-kernel times + DRAM bytes
timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:kmeans\|cosine_topk \
  --csv --log-file $O/kmeans_topk_launches.csv python tools/kbench.py kmeans topk > /dev/null 2>&1
if [ "${1:-}" = "mega" ]; then
  # one decode_mega_kernel launch, 3 sequences (the answer + 2 summaries batch of the bench), full sets + source
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:decode_mega_kernel --launch-skip 2 --launch-count 1 \
    -f -o $O/decode_mega_r3_full python tools/decode_probe.py 4 5800 3 > $O/decode_mega_under_ncu.log 2>&1
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:peer_allgather\|mean_pool\|cosine_topk --launch-count 3 \
    -f -o $O/topk_full python tools/kbench.py topk_small > /dev/null 2>&1
fi
ls -la $O | tail -14
