"""CPU: batching logic of gif_b200.inference.get_images_from_flame_params (generic_utils.py:33-55) with a stand-in model."""
import numpy as np
import torch

from gif_b200.inference import get_images_from_flame_params


class _Stub(torch.nn.Module):
    def __init__(self):
        super().__init__()
        self.w = torch.nn.Parameter(torch.ones(1))
        self.calls = []

    def forward(self, x, pose, step=0, alpha=1, input_indices=None):
        self.calls.append((x.shape[0], step, alpha, None if pose is None else pose.shape[0], self.training))
        img = x[:, :3] * 3.0 + input_indices.view(-1, 1, 1, 1).float() * 0.0
        return [img * 0, img]


def test_batches_of_16_clamped_concatenated_in_order():
    m = _Stub().train()
    n = 37
    x = np.random.RandomState(0).uniform(-1, 1, (n, 6, 4, 4)).astype(np.float32)
    idx = np.arange(n, dtype=np.int64)
    pose = np.zeros((n, 3), dtype=np.float32)
    out = get_images_from_flame_params(x, pose, m, step=2, alpha=1, input_indices=idx)
    assert [c[0] for c in m.calls] == [16, 16, 5] and all(c[1:4] == (2, 1, c[0]) and c[4] is False for c in m.calls)
    assert tuple(out.shape) == (n, 3, 4, 4) and out.device.type == "cpu"
    assert torch.equal(out, torch.clamp(torch.from_numpy(x[:, :3]) * 3.0, -1, 1))
    assert m.training            # restored
    out2 = get_images_from_flame_params(torch.from_numpy(x), None, m, 2, 1, torch.from_numpy(idx), batch_size=8)
    assert torch.equal(out, out2) and m.calls[-1][3] is None
