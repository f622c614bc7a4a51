import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
os.environ.setdefault("MEEPO_PEER_TIMEOUT_MS", "5000")
from meepoembedding_b200 import Table, product_library
from util import make_keys, table_kwargs
lib = product_library()
world, dim = 2, 128
kw = table_kwargs(dim=dim, capacity=1 << 14)
tables = [Table(lib=lib, device=0, **kw) for r in range(world)]
blobs = b"".join(t.peer_prepare(r, world, 6000, 0) for r, t in enumerate(tables))
for t in tables: t.peer_attach(blobs)
streams = [torch.cuda.Stream() for _ in range(world)]
rng = np.random.default_rng(1)
per = [make_keys(rng, n, 5000, dup_frac=0.4, invalid=True) for n in (257, 4096)]
dk = [torch.from_numpy(k.view(np.int64)).cuda() for k in per]
rows = [torch.zeros((k.size, dim), device="cuda") for k in per]
st = [torch.full((k.size,), 99, dtype=torch.uint8, device="cuda") for k in per]
torch.cuda.synchronize()
for r in range(world):
    tables[r].sharded_find_or_insert(dk[r], rows[r], st[r], stream=streams[r].cuda_stream)
torch.cuda.synchronize()
for r in range(world):
    s = st[r].cpu().numpy()
    own = np.array([lib.owner(int(k), world) for k in per[r]])
    print("rank", r, "n", per[r].size, "status hist", np.bincount(s, minlength=5)[:5], "unique", np.unique(per[r]).size)
    for o in range(world):
        m = own == o
        print("   owner", o, "elements", m.sum(), "status hist", np.bincount(s[m], minlength=5)[:5])
    bad = np.where(s == 0)[0]
    if bad.size:
        uk, first = np.unique(per[r], return_index=True)
        print("   first bad idx", bad[:10], "n bad unique", np.unique(per[r][bad]).size)
for t in tables:
    print(t.stats())
