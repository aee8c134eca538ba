import sys, time, cProfile, pstats, io
sys.path.insert(0, '.')
sys.argv=['bench.py']
import torch, bench
from trackmpnn_b200 import synth
from trackmpnn_b200.models.track_mpnn import TrackMPNN
a = bench.parse()
dev = torch.device('cuda:0')
torch.manual_seed(5)
model = TrackMPNN('2d', 3, 64, 0, 'diff').to(dev).train()
opt = torch.optim.Adam(model.parameters(), lr=1e-4, weight_decay=5e-4)
def chunk(seed):
    ts = synth.train_chunk_timestamps(seed, 5, 2)
    return synth.make_sequence(seed, None, 40, 'kitti', timestamps=ts)
dd = [(torch.from_numpy(X).to(dev), torch.from_numpy(y).to(dev)) for X, y in (chunk(1000+i) for i in range(8))]
for X, y in dd[:3]: bench.train_chunk_cuda(model, opt, X, y)
torch.cuda.synchronize()
t0=time.perf_counter()
for X, y in dd[3:]: bench.train_chunk_cuda(model, opt, X, y)
torch.cuda.synchronize(); print('ms/chunk', (time.perf_counter()-t0)/5*1e3)
pr = cProfile.Profile(); pr.enable()
for X, y in dd[3:]: bench.train_chunk_cuda(model, opt, X, y)
torch.cuda.synchronize(); pr.disable()
s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats('cumulative').print_stats(45); print(s.getvalue()[:7000])
