# same-box A/B of the softmax reference window of the attention kernel (CTX_ATTN_TAU / CTX_ATTN_REFUP; 12 / 0 = round-1 behaviour)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_net.py -m gpu -x -q -k "attention or transformer or forward" 2>&1 | tail -5 > gpurun_out/tau_pytest.txt; cat gpurun_out/tau_pytest.txt
for r in 1 2; do for v in "12 0" "15 6" "15 9"; do
  set -- $v
  echo "TAU=$1 REFUP=$2" | tee -a gpurun_out/tau_ab.txt
  CTX_ATTN_TAU=$1 CTX_ATTN_REFUP=$2 timeout 300 python profiles/dev/attn_ab.py 2>&1 | tail -1 | tee -a gpurun_out/tau_ab.txt
done; done
