"""Order-stable digests of tensors / state_dicts shared by the golden generator and the tests."""
import hashlib

import numpy as np
import torch


def digest(*arrays) -> str:
    h = hashlib.sha256()
    for a in arrays:
        if isinstance(a, torch.Tensor):
            a = a.detach().cpu().contiguous().numpy()
        a = np.ascontiguousarray(a)
        h.update(str(a.dtype).encode())
        h.update(str(a.shape).encode())
        h.update(a.tobytes())
    return h.hexdigest()


def state_digest(sd) -> str:
    h = hashlib.sha256()
    for k in sd:
        h.update(k.encode())
        h.update(digest(sd[k]).encode())
    return h.hexdigest()
