#!/bin/bash
# ncu evidence of one DDPM-256 sparse step @1.2 % (engine launched eagerly between cudaProfilerStart/Stop: bench.py --ncu).
# Usage (on the GPU box, via gpurun): bash tools/profile_step.sh <tag>     -> gpurun_out/<tag>_*.csv|txt|ncu-rep
tag=${1:-r01x}
mkdir -p gpurun_out
NCU="timeout 280 ncu --profile-from-start off --clock-control none"
$NCU --metrics gpu__time_duration.sum --csv --log-file gpurun_out/${tag}_launches_engine_step.csv python bench.py --ncu > gpurun_out/${tag}_ncu.log 2>&1
$NCU --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum -k regex:tile_conv --csv --log-file gpurun_out/${tag}_tileconv_dram_traffic_step.csv python bench.py --ncu >> gpurun_out/${tag}_ncu.log 2>&1
$NCU --set full --import-source on -k regex:tile_conv_tc5 -s 20 -c 6 -o gpurun_out/${tag}_tc5_full -f python bench.py --ncu >> gpurun_out/${tag}_ncu.log 2>&1
$NCU --set full --import-source on -k regex:attention_kernel -c 2 -o gpurun_out/${tag}_attn_full -f python bench.py --ncu >> gpurun_out/${tag}_ncu.log 2>&1
timeout 200 python tools/trace_graph.py --detail > gpurun_out/${tag}_graph_timeline.txt 2>&1
tail -2 gpurun_out/${tag}_graph_timeline.txt
