# Round-1 GPU session (one B200): parity tests, the contract bench line, the 512x512 config, ncu launch list (per-launch
# durations + DRAM bytes of two timed steps) and --set full captures of the dominant kernels exported as CSV.
T=${1:-r1k}
mkdir -p gpurun_out /tmp/rep
python -m pytest tests -m gpu -x -q 2>&1 | tail -6 > gpurun_out/${T}_pytest.txt; cat gpurun_out/${T}_pytest.txt
CTX_AUTOTUNE_LOG=gpurun_out/${T}_autotune.log python bench.py --steps 20 --warmup 5 --layers gpurun_out/${T}_layers.json > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err
cat gpurun_out/${T}_bench.json; tail -5 gpurun_out/${T}_bench.err
python bench.py --size 512 --batch 16 --precision fp16 --steps 10 --warmup 3 --no-cpu-baseline --layers gpurun_out/${T}_layers_512.json > gpurun_out/${T}_bench_512_fp16_b16.json 2>> gpurun_out/${T}_bench.err
cat gpurun_out/${T}_bench_512_fp16_b16.json
