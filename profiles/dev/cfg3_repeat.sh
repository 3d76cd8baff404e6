# BASELINE config 3 (512x512 fp16 B=16, linear soft-NMS) alone, several times: e2e with DetectPost on its own stream vs behind the forward
mkdir -p gpurun_out
for r in 1 2 3; do for ps in own main; do
  CTX_BENCH_POST_STREAM=$ps timeout 300 python bench.py --size 512 --batch 16 --precision fp16 --nms linear --steps 10 --warmup 3 --no-cpu-baseline --no-extra 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('post_stream=$ps value %.0f e2e %.0f (%.2f ms/step, serial %.2f)' % (d['value'], d['e2e']['value'], d['e2e']['ms_per_step'], d['e2e']['serial_ms_per_step']))"
done; done
