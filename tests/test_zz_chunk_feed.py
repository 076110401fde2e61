"""ChunkFeed (soap_b200/halo_tasks.py): pageable host buffers -- what SOAP's SharedArray windows are -- are
page-locked in place and uploaded asynchronously; the chunk that arrives is the chunk that was sent."""

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_chunk_feed_registers_pageable_buffers_and_uploads_them():
    import torch

    from soap_b200.halo_tasks import ChunkFeed

    rng = np.random.default_rng(7)
    pos = rng.uniform(0.0, 10.0, size=(1 << 18, 3))
    mass = rng.uniform(0.5, 1.5, size=1 << 18).astype(np.float32)
    data = {1: {"Coordinates": pos, "Masses": mass}}
    halos = {"cofp": torch.zeros((4, 3), dtype=torch.float64)}
    feed = ChunkFeed(0)
    assert not torch.from_numpy(pos).is_pinned()
    token = feed.register(data)
    assert len(token) == 2
    assert torch.from_numpy(pos).is_pinned() and torch.from_numpy(mass).is_pinned()
    ticket = feed.upload({t: {k: torch.from_numpy(v) for k, v in d.items()} for t, d in data.items()}, halos)
    dev, h_dev = feed.wait(ticket)
    torch.cuda.synchronize()
    assert dev[1]["Coordinates"].is_cuda and h_dev["cofp"].is_cuda
    assert np.array_equal(dev[1]["Coordinates"].cpu().numpy(), pos)
    assert np.array_equal(dev[1]["Masses"].cpu().numpy(), mass)
    feed.unregister(token)
    assert token == [] and not torch.from_numpy(pos).is_pinned()
