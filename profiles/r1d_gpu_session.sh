mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/r1d_pytest.txt
python bench.py --steps 20 --warmup 5 --layers gpurun_out/r1d_layers.json > gpurun_out/r1d_bench.json 2> gpurun_out/r1d_bench.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r1d_launches.csv python bench.py --quick --steps 2 --warmup 3 > gpurun_out/r1d_ncu_launch.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:conv_tc_kernel -c 16 -f -o gpurun_out/r1d_conv_full python bench.py --quick --no-graph --steps 1 --warmup 3 > gpurun_out/r1d_ncu_conv.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:attention -c 2 -f -o gpurun_out/r1d_attn_full python bench.py --quick --no-graph --steps 1 --warmup 3 > gpurun_out/r1d_ncu_attn.log 2>&1
ncu --set full --clock-control none --import-source on -k 'regex:select_candidates|class_nms|image_select|detect_forward' -c 8 -f -o gpurun_out/r1d_post_full python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/r1d_ncu_post.log 2>&1
cat gpurun_out/r1d_pytest.txt; cat gpurun_out/r1d_bench.json; tail -3 gpurun_out/r1d_bench.err
ls -la gpurun_out
