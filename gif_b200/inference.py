"""EMA-generator inference loop of the evaluation path: ``get_images_from_flame_params``
(my_utils/generic_utils.py:33-55, called every 500 iterations for FID at train.py:267-272 and by the plot scripts) --
same signature and result; batches go up through pinned host memory with non-blocking copies and the clamped images come
back into one pre-allocated pinned tensor, so the device never waits for the host between the 16-image batches."""
import numpy as np
import torch


def get_images_from_flame_params(flame_params, pose, model, step, alpha, input_indices, batch_size=16, device=None):
    """flame_params (N, ...) numpy / tensor: the generator input (condition maps (N,6,H,W) in rendered-condition mode);
    pose (N, .) or None; input_indices (N,) -> images (N,3,R,R) on the CPU, clamped to [-1, 1]."""
    was_training = getattr(model, "training", False)
    if hasattr(model, "eval"):
        model.eval()
    as_t = lambda a: a if torch.is_tensor(a) else torch.from_numpy(np.asarray(a))
    flame_params, input_indices = as_t(flame_params), as_t(input_indices)
    pose = None if pose is None else as_t(pose)
    if device is None:
        p = next(iter(model.parameters()), None) if hasattr(model, "parameters") else None
        device = p.device if p is not None else torch.device("cuda" if torch.cuda.is_available() else "cpu")
    device = torch.device(device)
    pin = device.type == "cuda"
    if pin:
        flame_params, input_indices = flame_params.pin_memory(), input_indices.pin_memory()
        pose = None if pose is None else pose.pin_memory()
    out = None
    n = flame_params.shape[0]
    with torch.no_grad():
        for b0 in range(0, n, batch_size):
            sl = slice(b0, b0 + batch_size)
            x = flame_params[sl].to(device, non_blocking=True)
            idx = input_indices[sl].to(device, non_blocking=True)
            ps = None if pose is None else pose[sl].to(device, non_blocking=True)
            img = torch.clamp(model(x, ps, step=step, alpha=alpha, input_indices=idx)[-1], -1, 1)
            if out is None:
                out = torch.empty((n,) + tuple(img.shape[1:]), dtype=img.dtype, pin_memory=pin)
            out[sl].copy_(img, non_blocking=True)
    if pin:
        torch.cuda.current_stream(device).synchronize()
    if was_training and hasattr(model, "train"):
        model.train()
    return out
