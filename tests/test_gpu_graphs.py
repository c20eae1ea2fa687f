"""GPU: the device graph builder (csrc/graph_build.cu through graphs.ball_pairs_device / multi_level_ball_graph on
CUDA tensors) against the golden edge lists / attributes produced by the reference's own generators
(G6: SquareMeshGenerator + sklearn pairwise_distances; G7: RandomMultiMeshGenerator).  pytest -m gpu."""
import os

import numpy as np
import pytest
import torch

from graph_pde_b200 import graphs
from tests.helpers import GOLDEN

pytestmark = pytest.mark.gpu
DEV = 'cuda:0'


def test_ball_pairs_device_matches_sklearn_golden():
    g = np.load(os.path.join(GOLDEN, 'g6_ball_graphs.npz'))
    for key in [k for k in g.files if k.startswith('ei/')]:
        s, r = key[3:].split('_')
        s, r = int(s), float(r)
        grid = graphs.square_grid(s, DEV, torch.float64)
        th = (torch.arange(s * s, dtype=torch.float64) * 0.01).to(DEV)
        ei, ea = graphs.ball_pairs_device(grid, grid, r, theta_a=th, theta_b=th, with_attr=True)
        np.testing.assert_array_equal(ei.cpu().numpy(), g[key].astype(np.int64))
        np.testing.assert_allclose(ea.cpu().numpy(), g['ea/' + key[3:]], atol=1e-6)
        # and the lattice-stencil generator used for the big meshes agrees with it on the device
        assert torch.equal(graphs.ball_connectivity(s, r, DEV), ei)


def test_multilevel_graph_on_device_matches_reference_golden():
    g7 = np.load(os.path.join(GOLDEN, 'g7_multilevel_graph.npz'))
    torch.manual_seed(0)
    g = graphs.multi_level_ball_graph(int(g7['s']), [int(v) for v in g7['m']], list(g7['ri']), list(g7['rx']),
                                      theta=torch.from_numpy(g7['theta']), device=DEV)
    for name, key in (('edge_index_mid', 'e_mid'), ('edge_index_down', 'e_down'), ('edge_index_up', 'e_up'),
                      ('edge_index_range', 'r_mid'), ('edge_index_down_range', 'r_down'),
                      ('edge_index_up_range', 'r_up')):
        assert np.array_equal(getattr(g, name).cpu().numpy(), g7[key].astype(np.int64)), name
    for name, key in (('edge_attr_mid', 'a_mid'), ('edge_attr_down', 'a_down'), ('edge_attr_up', 'a_up')):
        assert np.allclose(getattr(g, name).cpu().numpy(), g7[key], atol=1e-6), name


def test_ball_pairs_device_large_and_rectangular():
    """241^2 points (3.4e9 candidate pairs, no N x N matrix) against the exact lattice rule; a rectangular
    (bipartite) call against dense float64 distances."""
    s, r = 241, 0.05
    grid = graphs.square_grid(s, DEV, torch.float64)
    ei = graphs.ball_pairs_device(grid, grid, r)
    ref = graphs.ball_connectivity(s, r, DEV)
    # distances == r (lattice ties, e.g. offset (0,12) at r*(s-1) = 12) are rounding dependent in float64
    # (SURVEY H3): the device kernel must contain every strict-interior edge and nothing beyond the ties-in set
    strict = graphs.ball_connectivity(s, r, DEV, ties_in=False)
    key = lambda e: e[0] * (s * s) + e[1]      # noqa: E731
    k_dev, k_in, k_all = key(ei), key(strict), key(ref)
    assert bool(torch.isin(k_in, k_dev).all()) and bool(torch.isin(k_dev, k_all).all())
    assert bool((k_dev[1:] > k_dev[:-1]).all())                      # source-major, destination ascending
    gen = torch.Generator().manual_seed(3)
    pa, pb = torch.rand(700, 2, generator=gen, dtype=torch.float64), torch.rand(1300, 2, generator=gen, dtype=torch.float64)
    e2 = graphs.ball_pairs_device(pa.to(DEV), pb.to(DEV), 0.07, src_base=5, dst_base=1000)
    d = torch.cdist(pa, pb)
    ref2 = torch.nonzero(d <= 0.07).t()
    assert torch.equal(e2.cpu(), torch.stack([ref2[0] + 5, ref2[1] + 1000]))
