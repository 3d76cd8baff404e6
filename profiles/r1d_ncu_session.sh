# ncu captures of the dominant kernels (one GPU).  Reports stay on the box (too large); the raw / source pages are
# exported as CSV into gpurun_out/.
mkdir -p gpurun_out /tmp/rep
M="gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,sm__throughput.avg.pct_of_peak_sustained_elapsed,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,lts__t_sector_hit_rate.pct,lts__throughput.avg.pct_of_peak_sustained_elapsed,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum,smsp__cycles_active.avg,sm__cycles_elapsed.max,launch__grid_size,launch__registers_per_thread,lts__t_bytes.sum,sm__inst_executed_pipe_xu.sum"
python bench.py --steps 5 --warmup 3 --layers gpurun_out/r1d_layers.json --no-cpu-baseline > gpurun_out/r1d_bench_nocpu.json 2> gpurun_out/r1d_bench.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r1d_launches.csv python bench.py --quick --steps 2 --warmup 3 > gpurun_out/r1d_ncu_launch.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:conv_tc_kernel -c 20 -f -o /tmp/rep/conv python bench.py --quick --no-graph --steps 1 --warmup 3 > gpurun_out/r1d_ncu_conv.log 2>&1
ncu -i /tmp/rep/conv.ncu-rep --page raw --csv > gpurun_out/r1d_conv_raw.csv 2>/dev/null
ncu --set full --clock-control none --import-source on -k regex:attention -c 1 -f -o /tmp/rep/attn python bench.py --quick --no-graph --steps 1 --warmup 3 > gpurun_out/r1d_ncu_attn.log 2>&1
ncu -i /tmp/rep/attn.ncu-rep --page raw --csv > gpurun_out/r1d_attn_raw.csv 2>/dev/null
ncu -i /tmp/rep/attn.ncu-rep --page source --csv > gpurun_out/r1d_attn_source.csv 2>/dev/null
ncu --set full --clock-control none -k 'regex:select_candidates|class_nms|image_select|detect_forward' -c 8 -f -o /tmp/rep/post python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/r1d_ncu_post.log 2>&1
ncu -i /tmp/rep/post.ncu-rep --page raw --csv > gpurun_out/r1d_post_raw.csv 2>/dev/null
ls -la /tmp/rep gpurun_out
