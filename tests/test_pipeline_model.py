"""CPU: the mbarrier protocol of the CTA-pair conv kernel (csrc/conv_tc2.cu) replayed by the
discrete-event model tools/pipeline_model.py under random schedules: no deadlock, no stage / accumulator overwritten
while still read, every consumer sees the data it expects."""
import importlib.util
import os

import pytest

_p = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools", "pipeline_model.py")
_s = importlib.util.spec_from_file_location("pipeline_model", _p)
P = importlib.util.module_from_spec(_s)
_s.loader.exec_module(P)


@pytest.mark.parametrize("tiles,iters", [(1, 1), (1, 9), (2, 18), (5, 18), (6, 4)])
def test_conv_pair_protocol(tiles, iters):
    for seed in range(12):
        assert P.run_conv_pair(tiles, iters, seed)
