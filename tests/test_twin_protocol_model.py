"""The mbarrier protocol of fresco_attn_twin_kernel, checked without a GPU: tools/twin_protocol_model.py restates which
barrier every role waits on / arrives at and with which phase parity (the expressions of attn_tcgen05.cu) and runs the
roles under random interleavings with asynchronous TMA / tensor-core engines.  A lost arrival or an aliased parity wait
shows up as a deadlock, a premature one as a data hazard."""
import os
import sys

import pytest

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools"))
import twin_protocol_model as M  # noqa: E402


@pytest.mark.parametrize("nb,stages,split,fold", [(3, 4, 1, True), (3, 4, 2, True), (2, 2, 1, False), (2, 4, 2, False),
                                                  (2, 4, 1, False)])
def test_no_deadlock_no_hazard(nb, stages, split, fold):
    """head_dim 40 (three score regions, folded row sum, one / two threads per row), head_dim 80 (two regions, ring of 2),
    head_dim 64 (two regions, ring of 4): every tile count from 1 (no steady state) to well past the ring depth."""
    for n_tiles in (1, 2, 3, 4, 5, 8, 13):
        for seed in range(4):
            M.simulate(n_tiles, nb=nb, stages=stages, split=split, fold=fold, seed=seed, rescale_prob=0.3)


def test_model_catches_the_aliased_epilogue_wait():
    """The protocol as first written waited for the last P V with a parity wait on bar_pv.  With three score regions that
    barrier may be two completions behind when a warp reaches the epilogue, and the wait passes on the stale phase; the
    kernel now has a barrier that completes exactly once (bar_fin).  The model must see the difference."""
    caught = 0
    for seed in range(40):
        try:
            M.simulate(4, nb=3, stages=4, split=1, fold=True, seed=seed, final_on_pv=True)
        except M.Hazard as e:
            assert "epilogue" in str(e)
            caught += 1
    assert caught > 0
    for seed in range(40):
        M.simulate(4, nb=3, stages=4, split=1, fold=True, seed=seed)
