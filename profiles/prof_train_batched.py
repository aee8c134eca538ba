"""One batched training step (32 KITTI-shaped chunks) for the profilers: eager launches of exactly what GraphedTrainStep
replays (flat gradient buffer, tcgen05 forward / backward contractions), bracketed by cudaProfilerStart / Stop
(`ncu --profile-from-start off`), plus CUDA-event timing of the eager and the graph-replayed step."""
import sys, time
sys.path.insert(0, '.')
import torch
from trackmpnn_b200 import synth
from trackmpnn_b200.models.track_mpnn import TrackMPNN
from trackmpnn_b200.train_engine import TrainBatch, GraphedTrainStep
dev = torch.device('cuda:0')
torch.manual_seed(5)
model = TrackMPNN('2d', 3, 64, 0, 'diff').to(dev).train()
chunks = []
for i in range(32):
    ts = synth.train_chunk_timestamps(3000 + i, 5, 2)
    X, y = synth.make_sequence(3000 + i, None, 40, 'kitti', timestamps=ts)
    chunks.append((torch.from_numpy(X).to(dev), torch.from_numpy(y).to(dev)))
t0 = time.perf_counter(); batch = TrainBatch(chunks, dev); torch.cuda.synchronize(); print('build s', time.perf_counter() - t0)
step = GraphedTrainStep(model, batch, lr=1e-4, weight_decay=5e-4)
for _ in range(3): step.eager()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
t0 = time.perf_counter(); e0.record(); step.eager(); e1.record(); t_host = time.perf_counter() - t0; torch.cuda.synchronize()
print('eager: host enqueue ms', 1e3 * t_host, 'gpu ms', e0.elapsed_time(e1))
torch.cuda.profiler.start(); step.eager(); torch.cuda.synchronize(); torch.cuda.profiler.stop()   # ncu --profile-from-start off
if '--no-graph' not in sys.argv:
    step.capture()
    ts = []
    for _ in range(8):
        torch.cuda.synchronize(); t0 = time.perf_counter(); step.replay(); torch.cuda.synchronize(); ts.append(1e3 * (time.perf_counter() - t0))
    print('graph replay step ms:', ' '.join(f'{t:.2f}' for t in ts), 'edge rows', batch.edge_rows)
