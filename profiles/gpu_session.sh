mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/r1j_pytest.txt
cat gpurun_out/r1j_pytest.txt
CTX_AUTOTUNE_LOG=gpurun_out/r1j_autotune.log python bench.py --steps 20 --warmup 5 --no-cpu-baseline --layers gpurun_out/r1j_layers.json > gpurun_out/r1j_bench.json 2> gpurun_out/r1j_bench.err
cat gpurun_out/r1j_bench.json; tail -5 gpurun_out/r1j_bench.err
for v in "CTX_LANES=1 CTX_AUTOTUNE=0"; do
  echo "== $v"; env $v python bench.py --quick --steps 20 --warmup 5 2>&1 | tail -1
done | tee gpurun_out/r1j_variants.txt
ncu --metrics gpu__time_duration.sum --clock-control none -k 'regex:conv_tc|attention|maxpool|softmax|kv_project|q_project|nchw' -c 300 --csv --log-file gpurun_out/r1j_launches.csv python bench.py --quick --steps 1 --warmup 3 > gpurun_out/r1j_ncu_launch.log 2>&1
