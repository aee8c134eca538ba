import sys, time, cProfile, pstats, io
sys.path.insert(0, '.')
import torch
from trackmpnn_b200 import synth
from trackmpnn_b200.train_engine import TrainBatch
dev = torch.device('cuda:0')
chunks = []
for i in range(32):
    ts = synth.train_chunk_timestamps(3000 + i, 5, 2)
    X, y = synth.make_sequence(3000 + i, None, 40, 'kitti', timestamps=ts)
    chunks.append((torch.from_numpy(X).to(dev), torch.from_numpy(y).to(dev)))
for k in range(3):
    torch.cuda.synchronize(); t0 = time.perf_counter(); b = TrainBatch(chunks, dev); torch.cuda.synchronize(); print('build ms', 1e3 * (time.perf_counter() - t0), b.builder)
pr = cProfile.Profile(); pr.enable(); b = TrainBatch(chunks, dev); torch.cuda.synchronize(); pr.disable()
s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats('cumulative').print_stats(22); print(s.getvalue()[:3800])
