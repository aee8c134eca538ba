"""Batched training step (32 KITTI-shaped chunks): eager vs captured as ONE CUDA graph (forward, losses, backward through
autograd into the flat gradient buffer, Adam with capturable=True)."""
import sys, time, traceback
sys.path.insert(0, '.')
import torch
from trackmpnn_b200 import synth, parallel
from trackmpnn_b200.models.track_mpnn import TrackMPNN
from trackmpnn_b200.train_engine import TrainBatch, batch_loss, GraphedTrainStep
dev = torch.device('cuda:0')
torch.manual_seed(5)
model = TrackMPNN('2d', 3, 64, 0, 'diff').to(dev).train()
chunks = []
for i in range(32):
    ts = synth.train_chunk_timestamps(3000 + i, 5, 2)
    X, y = synth.make_sequence(3000 + i, None, 40, 'kitti', timestamps=ts)
    chunks.append((torch.from_numpy(X).to(dev), torch.from_numpy(y).to(dev)))
batch = TrainBatch(chunks, dev)
step = GraphedTrainStep(model, batch, lr=1e-4, weight_decay=5e-4)
for mode in ('eager', 'graph'):
    try:
        fn = step.eager if mode == 'eager' else step.replay
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        ts = []
        for _ in range(8):
            t0 = time.perf_counter(); fn(); torch.cuda.synchronize(); ts.append(1e3 * (time.perf_counter() - t0))
        print(mode, 'step ms:', ' '.join(f'{t:.2f}' for t in ts), 'loss', float(step.loss))
    except Exception:
        traceback.print_exc()
