# ncu evidence (one B200).  Kernels are launched as plain stream launches (--no-graph, no programmatic dependent launch):
# ncu serialises and replays every kernel anyway, and cannot replay kernel nodes of the captured multi-lane graph.
# Only the timed steps are profiled (bench.py brackets them with cudaProfilerStart/Stop).
T=${1:-r1k}
mkdir -p gpurun_out /tmp/rep
export CTX_CONV_PDL=0
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed,lts__throughput.avg.pct_of_peak_sustained_elapsed,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed
ncu --profile-from-start off --metrics $M --clock-control none -c 400 --csv --log-file gpurun_out/${T}_launches.csv python bench.py --quick --no-graph --steps 2 --warmup 3 > gpurun_out/${T}_ncu_launch.log 2>&1
ncu --profile-from-start off --metrics $M --clock-control none -c 400 --csv --log-file gpurun_out/${T}_launches_512.csv python bench.py --quick --no-graph --size 512 --batch 16 --precision fp16 --steps 2 --warmup 3 > gpurun_out/${T}_ncu_launch_512.log 2>&1
ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:'conv_tc_kernel|conv_halo|conv_stem2' -c 16 -f -o /tmp/rep/conv python bench.py --quick --no-graph --steps 1 --warmup 3 > gpurun_out/${T}_ncu_conv.log 2>&1
ncu -i /tmp/rep/conv.ncu-rep --page raw --csv > gpurun_out/${T}_conv_raw.csv 2>/dev/null
ncu --profile-from-start off --set full --clock-control none -k regex:"attention_tc|proj_kv_tc" -c 2 -f -o /tmp/rep/attn python bench.py --quick --no-graph --steps 1 --warmup 3 > gpurun_out/${T}_ncu_attn.log 2>&1
ncu -i /tmp/rep/attn.ncu-rep --page raw --csv > gpurun_out/${T}_attn_raw.csv 2>/dev/null
tail -n 3 gpurun_out/${T}_ncu_launch.log gpurun_out/${T}_ncu_conv.log gpurun_out/${T}_ncu_attn.log
wc -l gpurun_out/${T}_*.csv
