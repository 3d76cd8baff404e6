mkdir -p gpurun_out
python -m pytest tests/test_gpu_net.py -m gpu -x -q -k "tiling or fused or lanes" 2>&1 | tail -4
rm -f gpurun_out/q_autotune.log
CTX_AUTOTUNE_LOG=gpurun_out/q_autotune.log python bench.py --quick --steps 20 --warmup 5 2>&1 | tail -1
