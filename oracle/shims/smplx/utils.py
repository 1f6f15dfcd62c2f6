import numpy as np
import torch


class Struct:
    def __init__(self, **kwargs):
        for k, v in kwargs.items():
            setattr(self, k, v)


def to_tensor(array, dtype=torch.float32):
    if torch.is_tensor(array):
        return array.to(dtype)
    return torch.tensor(array, dtype=dtype)


def to_np(array, dtype=np.float32):
    if "scipy.sparse" in str(type(array)):
        array = array.todense()
    if hasattr(array, "r") and not isinstance(array, np.ndarray):
        array = array.r
    return np.array(array, dtype=dtype)
