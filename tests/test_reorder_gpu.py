"""GPU parity of the PBT policy-batch reorder (mlb_reorder_chunks / mlb_gather_rows_clip) --
bit-exact against (i) the fixtures generated from the unmodified reference source on the
reference's own four KAT vectors (tests/golden/reorder_chunks.npz, reference
tests/test_rollouts.py:58-81), (ii) the oracle on random assignments incl. empty policies, ragged
sizes and one-policy inputs, and (iii) the round-trip property the reference's test asserts."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = 'cuda:0'
G = os.path.join(os.path.dirname(__file__), 'golden')


def _run(mlb, a, P, C, B):
    from madrona_learn_b200.pbt_reorder import _compute_reorder_chunks
    tp, ts = _compute_reorder_chunks(torch.from_numpy(a.astype(np.int32)).to(DEV), P, C, B)
    torch.cuda.synchronize()
    return tp.cpu().numpy(), ts.cpu().numpy()


def test_reference_kat_vectors(mlb):
    z = np.load(os.path.join(G, 'reorder_chunks.npz'))
    for i in range(4):
        a = z[f'v{i}_in']
        P, C = 6, 4
        B = a.size // C + P - 1
        tp, ts = _run(mlb, a, P, C, B)
        np.testing.assert_array_equal(tp, z[f'v{i}_to_policy'])
        np.testing.assert_array_equal(ts, z[f'v{i}_to_sim'])


@pytest.mark.parametrize('S,P,C,seed', [(13, 6, 4, 0), (1000, 7, 16, 1), (4097, 33, 64, 2), (100_000, 128, 256, 3),
                                        (65_536, 4, 1024, 4), (300, 1, 32, 5), (1 << 20, 64, 2048, 6), (5, 9, 2, 7)])
def test_matches_oracle_and_roundtrips(mlb, S, P, C, seed):
    from oracle import layouts
    from madrona_learn_b200.pbt_reorder import reorder_state_for
    rng = np.random.default_rng(seed)
    present = rng.random(P) < 0.8                     # some policies get no agents at all
    present[rng.integers(P)] = True
    a = rng.choice(np.nonzero(present)[0], size=S).astype(np.int32)
    B = S // C + P
    tp, ts = _run(mlb, a, P, C, B)
    rtp, rts = layouts.compute_reorder_chunks(a, P, C, B)
    np.testing.assert_array_equal(tp, rtp)
    np.testing.assert_array_equal(ts, rts)
    # round trip through the two gathers: to_sim(to_policy(x)) == x  (tests/test_rollouts.py:36-56)
    st = reorder_state_for(torch.from_numpy(a).to(DEV), P, C)
    x = torch.randn(S, 5, device=DEV)
    pol = st.to_policy(x)
    assert pol.shape == (B, C, 5)
    back = st.to_sim(pol)
    assert torch.equal(back, x)
    # every chunk holds agents of a single policy (what the grouped policy GEMM relies on)
    ids = torch.from_numpy(a).to(DEV).float().unsqueeze(1)
    chunks = st.to_policy(ids)[..., 0]
    valid = st.to_policy_idxs < S
    lo = torch.where(valid, chunks, torch.full_like(chunks, 1e9)).min(dim=1).values
    hi = torch.where(valid, chunks, torch.full_like(chunks, -1e9)).max(dim=1).values
    nonempty = valid.any(dim=1)
    assert torch.equal(lo[nonempty], hi[nonempty])
