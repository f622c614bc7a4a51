"""Device-buffer helpers for the GPU parity tests (torch is only the allocator here)."""
import numpy as np
import torch

from meepoembedding_b200 import _capi as capi

DEV = "cuda:0"


def dkeys(keys: np.ndarray) -> torch.Tensor:
    return torch.from_numpy(np.ascontiguousarray(keys).view(np.int64)).to(DEV)


def drows(rows: np.ndarray, dtype: str) -> torch.Tensor:
    """Host rows/grads (fp32, or bf16 carried as uint16 bits) -> device tensor of the table dtype."""
    if dtype == "f32":
        return torch.from_numpy(np.ascontiguousarray(rows, dtype=np.float32)).to(DEV)
    return torch.from_numpy(np.ascontiguousarray(rows).view(np.int16)).to(DEV).view(torch.bfloat16)


def hrows(rows: torch.Tensor, dtype: str) -> np.ndarray:
    """Device rows -> host array in the oracle's representation."""
    if dtype == "f32":
        return rows.cpu().numpy()
    return rows.view(torch.int16).cpu().numpy().view(np.uint16)


def gpu_foi(t, keys: np.ndarray, dtype: str, insert=True):
    k = dkeys(keys)
    rows, st = (t.find_or_insert if insert else t.lookup)(k)
    torch.cuda.synchronize()
    return hrows(rows, dtype), st.cpu().numpy()


def gpu_apply(t, keys: np.ndarray, grads: np.ndarray, dtype: str):
    t.apply_gradients(dkeys(keys), drows(grads, dtype))
    torch.cuda.synchronize()
