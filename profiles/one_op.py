"""Run single ops of the compiled program inside a profiler range (for ncu --profile-from-start off).
usage: one_op.py <precision> <batch> <size> <op name substring> [reps]"""
import os, sys, types
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
prec, B, size, pat = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), sys.argv[4]
reps = int(sys.argv[5]) if len(sys.argv) > 5 else 2
dev = torch.device('cuda', 0)
w = bench.Workload(size, prec, B, dev, 'hard', seed=0)
eng = w.eng
eng.load_input(w.x_dev)
eng.run_range(0, eng.num_ops)
torch.cuda.synchronize()
idx = [i for i, l in enumerate(eng.layers) if pat in l[0]]
print('ops', [(i, eng.layers[i][0]) for i in idx])
torch.cuda.profiler.start()
for i in idx:
    for _ in range(reps):
        eng.run_range(i, i + 1)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
