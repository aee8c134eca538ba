"""Batched training step (32 KITTI-shaped chunks) with and without attention heads: CUDA-event time per step."""
import sys
sys.path.insert(0, '.')
import torch
from trackmpnn_b200 import synth
from trackmpnn_b200.models.track_mpnn import TrackMPNN
from trackmpnn_b200.train_engine import TrainBatch, batch_loss
dev = torch.device('cuda:0')
chunks = []
for i in range(32):
    ts = synth.train_chunk_timestamps(3000 + i, 5, 2)
    X, y = synth.make_sequence(3000 + i, None, 40, 'kitti', timestamps=ts)
    chunks.append((torch.from_numpy(X).to(dev), torch.from_numpy(y).to(dev)))
batch = TrainBatch(chunks, dev)
for heads in (0, 2, 0, 2):
    torch.manual_seed(5)
    model = TrackMPNN('2d', 3, 64, heads, 'diff').to(dev).train()
    opt = torch.optim.Adam(model.parameters(), lr=1e-4, weight_decay=5e-4)

    def step():
        opt.zero_grad(); l = batch_loss(model, batch); l.backward(); opt.step(); return l
    for _ in range(3):
        step()
    torch.cuda.synchronize()
    ts = []
    for _ in range(7):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); l = step(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    ts.sort()
    print(f'heads {heads}: median step {ts[3]:.2f} ms  ({32e3 / ts[3]:.0f} chunks/s)  loss {float(l):.4f}  '
          f'grad W_att {"-" if heads == 0 else float(model.factor_grus[0].gat[0].W_att.grad.abs().max())}')
