"""CPU: the mbarrier protocols of the two bring-up kernels (csrc/conv_halo.cu, csrc/conv_tc2.cu) replayed by the
discrete-event model tools/pipeline_model.py under random schedules: no deadlock, no stage / accumulator overwritten
while still read, every consumer sees the data it expects.  A deliberately broken configuration must be caught."""
import importlib.util
import os

import pytest

_p = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools", "pipeline_model.py")
_s = importlib.util.spec_from_file_location("pipeline_model", _p)
P = importlib.util.module_from_spec(_s)
_s.loader.exec_module(P)


@pytest.mark.parametrize("items,kchunks,w_stages", [(1, 1, 2), (1, 2, 3), (4, 1, 2), (5, 2, 3), (3, 4, 4), (7, 1, 4)])
def test_conv_halo_protocol(items, kchunks, w_stages):
    for seed in range(12):
        assert P.run_conv_halo(items, kchunks, w_stages, seed)


@pytest.mark.parametrize("tiles,iters", [(1, 1), (1, 9), (2, 18), (5, 18), (6, 4)])
def test_conv_pair_protocol(tiles, iters):
    for seed in range(12):
        assert P.run_conv_pair(tiles, iters, seed)


def test_model_catches_a_broken_protocol():
    # one activation stage: the look-ahead load of chunk j+1 waits for chunk j, whose remaining weight taps the same
    # producer thread has not issued yet
    with pytest.raises(AssertionError, match="deadlock"):
        P.run_conv_halo(3, 1, 3, 0, a_stages=1)
    # a producer that refills weight stages without waiting for the tensor core: data hazard, not a deadlock
    with pytest.raises(AssertionError):
        P.run_conv_halo(3, 2, 2, 1, skip_w_empty_wait=True)
