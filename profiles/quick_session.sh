mkdir -p gpurun_out
python -m pytest tests/test_gpu_net.py -m gpu -x -q -k "conv or tiling or fused or lanes" 2>&1 | tail -4
rm -f gpurun_out/q_autotune.log
CTX_AUTOTUNE_LOG=gpurun_out/q_autotune.log python bench.py --quick --steps 20 --warmup 5 --layers gpurun_out/q_layers.json 2>&1 | tail -1
python bench.py --steps 5 --warmup 3 --no-cpu-baseline --layers gpurun_out/q_layers.json 2>&1 | tail -1 | cut -c1-200
