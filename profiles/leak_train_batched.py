import sys, time
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
batch = TrainBatch(chunks, dev)
print('after build: alloc GB', torch.cuda.memory_allocated()/1e9, 'reserved', torch.cuda.memory_reserved()/1e9)
for i in range(14):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    opt.zero_grad(); l = batch_loss(model, batch); l.backward(); opt.step(); del l
    torch.cuda.synchronize()
    print(i, f'{1e3*(time.perf_counter()-t0):7.1f} ms  alloc {torch.cuda.memory_allocated()/1e9:.2f} GB  reserved {torch.cuda.memory_reserved()/1e9:.2f}  max {torch.cuda.max_memory_allocated()/1e9:.2f}  device_allocs {torch.cuda.memory_stats()["num_device_alloc"]}')
