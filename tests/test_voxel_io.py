"""the fp16 feature-grid loader (host side only; runs without a GPU)."""
import os
import sys

import numpy as np
import pytest
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pixie_b200 import voxel_io as V  # noqa: E402


def _write(tmp, name, arr):
    d = os.path.join(tmp, name)
    os.makedirs(d, exist_ok=True)
    np.save(os.path.join(d, V.FEATURE_FILE), arr)
    return d


def test_load_is_bitwise_and_matches_reference_dataset_layout(tmp_path):
    rng = np.random.default_rng(0)
    a = rng.normal(0, 0.05, size=(6, 6, 6, 8)).astype(np.float16)
    d = _write(str(tmp_path), "obj0", a)
    t = V.load_feature_grid(os.path.join(d, V.FEATURE_FILE), pin=False)
    assert t.dtype == torch.float16 and tuple(t.shape) == (1, 6, 6, 6, 8)
    assert np.array_equal(t.numpy()[0], a)
    # what the reference dataset hands to the network (my_data.py:163, 221): float32 (C, D, H, W) of the same values
    ref = torch.from_numpy(a.astype(np.float32)).permute(3, 0, 1, 2)
    assert torch.equal(t[0].permute(3, 0, 1, 2).to(torch.float32), ref)


def test_rejects_wrong_dtype_and_shape(tmp_path):
    d = _write(str(tmp_path), "bad", np.zeros((4, 4, 4, 8), np.float32))
    with pytest.raises(TypeError):
        V.load_feature_grid(os.path.join(d, V.FEATURE_FILE), pin=False)
    d2 = _write(str(tmp_path), "bad2", np.zeros((4, 5, 4, 8), np.float16))
    with pytest.raises(ValueError):
        V.load_feature_grid(os.path.join(d2, V.FEATURE_FILE), pin=False)
    with pytest.raises(FileNotFoundError):
        V.load_mask(os.path.join(str(tmp_path), "nope.npy"))


def test_scene_stream_cycles_buffers(tmp_path, monkeypatch):
    monkeypatch.setattr(torch.cuda, "is_available", lambda: False)
    dirs = [_write(str(tmp_path), f"o{i}", np.full((4, 4, 4, 8), i, np.float16)) for i in range(5)]
    seen = []
    for d, t in V.scene_stream(dirs, n_buffers=2):
        seen.append((d, float(t[0, 0, 0, 0, 0]), t.data_ptr()))
    assert [s[1] for s in seen] == [0.0, 1.0, 2.0, 3.0, 4.0]
    assert seen[0][2] == seen[2][2] == seen[4][2] and seen[1][2] == seen[3][2] and seen[0][2] != seen[1][2]


def test_scene_stream_waits_for_the_consumer_before_rewriting(tmp_path, monkeypatch):
    """A recycled pinned buffer is rewritten only after the `h2d_done` event its last user recorded (the consumer's
    host->device copy is asynchronous)."""
    monkeypatch.setattr(torch.cuda, "is_available", lambda: False)

    class Ev:
        def __init__(self):
            self.waited = False

        def synchronize(self):
            self.waited = True

    dirs = [_write(str(tmp_path), f"o{i}", np.full((4, 4, 4, 8), i, np.float16)) for i in range(5)]
    items = []
    for item in V.scene_stream(dirs, n_buffers=2):
        # by the time scene i is produced, the generator must have waited for scene i-2 (same buffer)
        if len(items) >= 2:
            assert items[-2].h2d_done.waited
            assert not items[-1].h2d_done.waited
        item.h2d_done = Ev()
        items.append(item)
    assert [float(it.tensor[0, 0, 0, 0, 0]) for it in items[-2:]] == [3.0, 4.0]
