"""GPU parity of the most recently added graph families (Net-as-node `Dag` programs, nonlinear biquads): same bar as
tests/test_gpu_jit.py (bit-exact vs the oracle through the C ABI), kept in a file that sorts after the parity tests."""
import numpy as np
import pytest

from test_gpu_jit import WIDER, run_case

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", sorted(WIDER))
def test_wider_graph_matches_oracle(name):
    b, g, o = run_case(WIDER[name], 40, 2000 + 61)
    assert g.shape == o.shape and np.isfinite(g).all() and np.abs(o).max() > 1e-4
    bad = int((g != o).sum())
    assert bad == 0, (name, bad, float(np.abs(g - o).max()), b.classes()[0]["signature"])
