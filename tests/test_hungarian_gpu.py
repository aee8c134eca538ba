"""GPU: the device assignment solver against scipy.optimize.linear_sum_assignment -- the reference's own
(un-vendored) dependency for --hungarian (utils/graph.py:18,86) -- on tie-heavy matrices shaped like the
reference's cost matrices (100.0 fillers, costs in [0, 1], saturated values), wide, tall and square."""
import numpy as np
import pytest
import torch
from scipy.optimize import linear_sum_assignment

pytestmark = pytest.mark.gpu


def _solve(C):
    from trackmpnn_b200 import _lib as L
    b, nr, nc = C.shape
    dev = torch.device('cuda:0')
    c = torch.from_numpy(np.ascontiguousarray(C, dtype=np.float32)).to(dev)
    out = torch.full((b, max(nr, 1)), -7, dtype=torch.int32, device=dev)
    nbytes = int(L.lib().tmpnn_lsap_scratch_bytes(b, nr, nc))
    scratch = torch.empty((nbytes + 7) // 8 + 1, dtype=torch.int64, device=dev)
    L.call('tmpnn_lsap_solve', L.ptr(c), nr, nc, b, L.ptr(out), L.ptr(scratch), L.stream())
    return out.cpu().numpy()


@pytest.mark.parametrize('nr,nc', [(1, 1), (3, 7), (7, 3), (12, 12), (40, 17), (17, 40), (64, 33), (90, 90), (150, 40)])
@pytest.mark.parametrize('kind', ['reference-like', 'integer-ties', 'saturated'])
def test_lsap_matches_scipy(nr, nc, kind):
    rs = np.random.RandomState(nr * 1000 + nc + len(kind))
    batch = 24
    if kind == 'reference-like':      # sparse bipartite graph: most pairs have no edge (100.0), the rest 1 - p
        C = np.full((batch, nr, nc), 100.0, np.float32)
        mask = rs.rand(batch, nr, nc) < 0.35
        C[mask] = rs.rand(int(mask.sum())).astype(np.float32)
    elif kind == 'integer-ties':      # small integer costs: many equal-cost optima
        C = rs.randint(0, 4, size=(batch, nr, nc)).astype(np.float32)
    else:                             # scores saturated at 0 / 1 plus fillers
        C = rs.choice(np.array([0.0, 1.0, 100.0], np.float32), size=(batch, nr, nc), p=[0.2, 0.3, 0.5])
    got = _solve(C)
    for b in range(batch):
        r, c = linear_sum_assignment(C[b])
        want = -np.ones(nr, np.int64)
        want[r] = c
        np.testing.assert_array_equal(got[b, :nr], want, err_msg=f'{kind} {nr}x{nc} matrix {b}')
