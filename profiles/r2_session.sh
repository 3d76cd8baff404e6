# Round-2 GPU session (one B200): parity tests, smoke(), the contract bench line with per-layer tables.
T=${1:-r2h}
mkdir -p gpurun_out
( time python -m pytest tests -m gpu -x -q ) 2>&1 | tail -8 > gpurun_out/${T}_pytest.txt; cat gpurun_out/${T}_pytest.txt
python __graft_entry__.py smoke > gpurun_out/${T}_smoke.txt 2>&1; tail -4 gpurun_out/${T}_smoke.txt
CTX_AUTOTUNE_LOG=gpurun_out/${T}_autotune.log python bench.py --steps 20 --warmup 5 --layers gpurun_out/${T}_layers.json > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err
cut -c1-400 gpurun_out/${T}_bench.json; tail -5 gpurun_out/${T}_bench.err
