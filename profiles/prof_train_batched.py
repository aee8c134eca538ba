"""cProfile + CUDA-event timing of one batched training step (32 KITTI-shaped chunks)."""
import sys, time, cProfile, pstats, io
sys.path.insert(0, '.')
import torch
from trackmpnn_b200 import synth
from trackmpnn_b200.models.track_mpnn import TrackMPNN
from trackmpnn_b200.train_engine import TrainBatch, batch_loss
dev = torch.device('cuda:0')
torch.manual_seed(5)
model = TrackMPNN('2d', 3, 64, 0, 'diff').to(dev).train()
opt = torch.optim.Adam(model.parameters(), lr=1e-4, weight_decay=5e-4)
chunks = []
for i in range(32):
    ts = synth.train_chunk_timestamps(3000 + i, 5, 2)
    X, y = synth.make_sequence(3000 + i, None, 40, 'kitti', timestamps=ts)
    chunks.append((torch.from_numpy(X).to(dev), torch.from_numpy(y).to(dev)))
t0 = time.perf_counter(); batch = TrainBatch(chunks, dev); torch.cuda.synchronize(); print('build s', time.perf_counter() - t0)
def step():
    opt.zero_grad(); l = batch_loss(model, batch); l.backward(); opt.step()
for _ in range(3): step()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
t0 = time.perf_counter(); e0.record(); step(); e1.record(); t_host = time.perf_counter() - t0; torch.cuda.synchronize()
print('host enqueue ms', 1e3 * t_host, 'gpu ms', e0.elapsed_time(e1))
torch.cuda.profiler.start(); step(); torch.cuda.synchronize(); torch.cuda.profiler.stop()   # ncu --profile-from-start off
pr = cProfile.Profile(); pr.enable(); step(); torch.cuda.synchronize(); pr.disable()
s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats('tottime').print_stats(22); print(s.getvalue()[:4500])
ts = []
for _ in range(8):
    torch.cuda.synchronize(); t0 = time.perf_counter(); step(); torch.cuda.synchronize(); ts.append(1e3 * (time.perf_counter() - t0))
print('step ms (8 more, synchronised):', ' '.join(f'{t:.1f}' for t in ts))
print('cuda mallocs', torch.cuda.memory_stats()['num_alloc_retries'], torch.cuda.memory_stats()['num_device_alloc'], 'reserved GB', torch.cuda.memory_reserved() / 1e9)
