"""Role timeline of the tcgen05 edge kernel: builds a -DTMPNN_TC_TRACE copy of the library (not the shipped one),
runs the C3 engine a few frames and dumps CTA 0's clock64() stamps of one steady-state launch.

    python profiles/trace_tc.py build            # here (nvcc) -> build/libtmpnn_trace.so
    TMPNN_LIB=build/libtmpnn_trace.so python profiles/trace_tc.py run   # on the GPU box -> gpurun_out/tc_trace.npy
    python profiles/trace_tc.py show gpurun_out/tc_trace.npy
slots: producer warp 0: 0 loop top, 1 x images free, 2 x part written, 3 h images free, 4 h part written + fenced;
MMA issuer: 5 accumulators free, 6 all producers arrived, 7 MMAs issued; epilogue warp 0: 8 loop top, 9 accumulators
ready, 10 previous state read, 11 gates done (TMEM released), 12 stores done (h images released), 13 head done."""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def build():
    import __graft_entry__ as G
    srcs = [os.path.join(G.CSRC, s) for s in G.SOURCES]
    os.makedirs(os.path.join(ROOT, 'build'), exist_ok=True)
    out = os.path.join(ROOT, 'build', 'libtmpnn_trace.so')
    subprocess.check_call(['nvcc'] + G.NVCC_FLAGS + ['-DTMPNN_TC_TRACE', '-I', os.path.join(ROOT, 'include')] + srcs + ['-o', out])
    print(out)


def run():
    import ctypes as C
    import numpy as np, torch
    sys.argv = ['bench.py']
    import bench
    from trackmpnn_b200 import _lib as L, synth
    from trackmpnn_b200.engine import TrackEngine
    from trackmpnn_b200.models.track_mpnn import TrackMPNN
    a = bench.parse(); a.frames = 10
    dev = torch.device('cuda:0')
    torch.manual_seed(5)
    model = TrackMPNN('2d', synth.num_categories(a.dataset), 64, 0, 'diff').to(dev).eval()
    eng = TrackEngine(model, bench.make_sequences(a, 0, a.seqs_per_gpu), cur_win_size=a.win, use_cuda_graph=False)
    eng.run(max_ticks=8); torch.cuda.synchronize()
    cap = 2048
    buf = torch.zeros((cap, 16), dtype=torch.int64, device=dev)
    kern = os.environ.get('TC_KERNEL', 'pre')
    eng.tensor_kernel = kern
    f = getattr(L.lib(), 'tmpnn_debug_set_tc3_trace' if kern == 'pre' else 'tmpnn_debug_set_tc_trace')
    f.argtypes = [C.c_void_p, C.c_int]; f(buf.data_ptr(), cap)
    eng._tick(flip=False); torch.cuda.synchronize()   # one more frame, traced (later launches overwrite earlier ones)
    f(None, 0)
    np.save(os.path.join(ROOT, 'gpurun_out', 'tc_trace.npy'), buf.cpu().numpy())
    print('saved', int((buf[:, 0] != 0).sum()), 'iterations')


def show3(path):
    """Timeline of the re-staged kernel (mp_step_tc3.cu): producer 2 loop top, 3 h images free, 4 h part written
    (-> arrive full), 0 before / 1 after the wait for the other stage's x images; issuer 5-7; epilogue teams 8-13."""
    import numpy as np
    t = np.load(path).astype(np.float64)
    n = int((t[:, 9] != 0).sum())
    t = t[:n]
    lo, hi = n // 4, 3 * n // 4
    seg = t[lo:hi]
    print(f'{n} tiles; steady-state cycles per tile {(seg[-1, 9] - seg[0, 9]) / (len(seg) - 1):.0f}')
    def d(a, b, shift=0):
        x = seg[:, b] - seg[:, a] if shift == 0 else seg[shift:, b] - seg[:-shift, a]
        return f'{np.mean(x):7.0f} (p10 {np.percentile(x, 10):6.0f} p90 {np.percentile(x, 90):6.0f})'
    print('producer: wait h free (gates of tile-2) ', d(2, 3))
    print('producer: h part + cp.async wait + fence', d(3, 4))
    print('producer: wait x free (stores of tile-1)', d(0, 1))
    print('producer: issue x copies -> next top    ', d(1, 2, 1))
    print('issuer  : wait tmem free                ', d(7, 5, 1))
    print('issuer  : wait full                     ', d(5, 6))
    print('issuer  : issue 36 MMA                  ', d(6, 7))
    print('full arrive -> issuer sees              ', d(4, 6))
    print('MMA     : issued -> epilogue sees done  ', d(7, 9))
    print('epilogue: wait done                     ', d(8, 9))
    print('epilogue: gates + stores (4 chunks)     ', d(9, 12))
    print('epilogue: head                          ', d(12, 13))
    print('epilogue: team tile -> next team tile   ', d(8, 8, 2))
    print('hfree -> producer sees (tile+2)         ', d(12, 3, 2))
    t0 = t[lo, 2]
    for i in range(lo, lo + 6):
        print(i, ' '.join(f'{(v - t0):8.0f}' for v in t[i, :14]))


def show(path):
    import numpy as np
    t = np.load(path).astype(np.float64)
    n = int((t[:, 9] != 0).sum())
    t = t[:n]
    t0 = t[0, 0]
    lo, hi = n // 4, 3 * n // 4          # steady state
    seg = t[lo:hi]
    per_tile = (seg[-1, 9] - seg[0, 9]) / (len(seg) - 1)
    print(f'{n} tiles; steady-state cycles per tile {per_tile:.0f}')
    def d(a, b, shift=0):
        x = seg[:, b] - seg[:, a] if shift == 0 else seg[shift:, b] - seg[:-shift, a]
        return f'{np.mean(x):7.0f} (p10 {np.percentile(x, 10):6.0f} p90 {np.percentile(x, 90):6.0f})'
    print('producer: wait x free      ', d(0, 1))
    print('producer: x part           ', d(1, 2))
    print('producer: wait h free      ', d(2, 3))
    print('producer: h part + fence   ', d(3, 4))
    print('producer: arrive -> next   ', d(4, 0, 1))
    print('issuer  : wait tmem free   ', d(4, 5))
    print('issuer  : wait full        ', d(5, 6))
    print('issuer  : issue            ', d(6, 7))
    print('MMA     : issue -> done(E) ', d(7, 9))
    print('epilogue: wait done        ', d(8, 9))
    print('epilogue: read h_prev      ', d(9, 10))
    print('epilogue: gates            ', d(10, 11))
    print('epilogue: stores           ', d(11, 12))
    print('epilogue: head             ', d(12, 13))
    print('hand-over E hfree -> P sees', d(12, 3, 2))
    for i in range(lo, lo + 6):
        print(i, ' '.join(f'{(v - t0):8.0f}' for v in t[i, :14]))


if __name__ == '__main__':
    {'build': build, 'run': run, 'show3': lambda: show3(sys.argv[2])}.get(sys.argv[1], lambda: show(sys.argv[2]))()
