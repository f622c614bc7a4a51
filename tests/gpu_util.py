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


def gpu_export(t, delta=False):
    """(keys, rows, state, scores, steps) of a CUDA table as host arrays, sorted by key (delta: the
    incremental export, which marks the returned tuples clean)."""
    n = t.export_delta_size() if delta else t.export_size()
    keys = torch.empty(n, dtype=torch.int64, device=DEV)
    rows = torch.empty((n, t.row_bytes), dtype=torch.uint8, device=DEV)
    state = torch.empty((n, max(t.state_bytes, 1)), dtype=torch.uint8, device=DEV)
    scores = torch.empty(n, dtype=torch.int64, device=DEV)
    steps = torch.empty(n, dtype=torch.int32, device=DEV)
    if not (delta and n == 0):
        got = t.export_buffers(keys, rows, state if t.state_bytes else None, scores, steps, max_n=n, delta=delta)
        assert got == n
    rdt = np.float32 if t.dtype == capi.F32 else np.uint16
    return (keys.cpu().numpy().view(np.uint64), rows.cpu().numpy().view(rdt).reshape(n, t.dim),
            state.cpu().numpy()[:, :t.state_bytes].copy().view(np.float32).reshape(n, t.state_bytes // 4),
            scores.cpu().numpy().view(np.uint64), steps.cpu().numpy().view(np.uint32))
