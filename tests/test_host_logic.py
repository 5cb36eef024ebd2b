"""Host-side logic that needs no GPU: the bookkeeping of host-fed results, the workload's sharding invariance, the
CPU-affinity helper."""
import importlib
import types

import numpy as np
import pytest


@pytest.fixture(scope="module")
def streams_mod():
    return importlib.import_module("rtmodt_b200.streams")


def test_step_result_knows_when_its_slot_has_moved_on(streams_mod):
    """Two slots: the result of step i lives in slot i % 2 until step i + 2 is enqueued.  A stale result must not wait
    on the slot's event (it now belongs to the later step) and must refuse to be read."""
    waited = []
    slot = dict(done=types.SimpleNamespace(synchronize=lambda: waited.append(1)), status=types.SimpleNamespace(numpy=lambda: np.zeros(2, np.int32)))
    feeder = types.SimpleNamespace(k=1, depth=2)
    res = streams_mod.StepResult(feeder, slot, 0)              # step 0, one step enqueued so far
    assert not res.stale and res.wait() is res and waited == [1]
    feeder.k = 2                                               # step 1 enqueued: slot 1, step 0 still owns slot 0
    assert not res.stale
    feeder.k = 3                                               # step 2 enqueued: slot 0 is its now
    assert res.stale
    res.wait()
    assert waited == [1], "a stale result waited on the later step's event"
    with pytest.raises(Exception, match="overwritten by step 2"):
        res.detections()


def test_workload_streams_do_not_depend_on_the_batch_they_are_generated_in():
    """SURVEY 8d 'identical head tensors': a stream's planted cells and its zones depend on its global id only."""
    import torch
    wl_mod = importlib.import_module("rtmodt_b200.workload")
    a = wl_mod.PostBackboneWorkload(3, 2, first_stream=4, device="cpu", dtype=torch.float32, num_objects=5)
    b = wl_mod.PostBackboneWorkload(1, 2, first_stream=5, device="cpu", dtype=torch.float32, num_objects=5)
    for f in range(2):
        for ta, tb in zip(a.heads[f], b.heads[f]):
            assert torch.equal(ta[1], tb[0])
    assert a.zones[1] == b.zones[0]


def test_cpu_binding_helper_degrades_quietly_without_nvml():
    sharding = importlib.import_module("rtmodt_b200.sharding")
    n = sharding.bind_process_to_gpu(0)
    assert n is None or n > 0
