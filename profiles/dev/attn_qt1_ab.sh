# same-box A/B of the attention kernel: one q-tile per CTA, two CTAs per SM (CTX_ATTN_QT1=1, default) vs two q-tiles per CTA (=0)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_net.py -m gpu -x -q -k "attention or transformer or forward" 2>&1 | tail -5 > gpurun_out/qt1_pytest.txt; cat gpurun_out/qt1_pytest.txt
for r in 1 2; do for v in 0 1; do
  CTX_ATTN_QT1=$v timeout 300 python profiles/dev/attn_ab.py 2>&1 | tail -2 | tee -a gpurun_out/qt1_ab.txt
done; done
