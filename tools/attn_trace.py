"""Timeline of block 0 of the attention backward kernel (pg_debug_set_trace): per role, average cycles spent in each
wait / work segment of a tile.  Development tool."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pytorch_generative_b200 import _lib as L

dev = torch.device("cuda:0")
N, S, H, D = int(os.environ.get("PG_N", 64)), 1024, 8, 64
P = N * S
qkv = torch.randn(P, 3 * H * D, device=dev).bfloat16()
q, k, v = qkv[:, :H * D], qkv[:, H * D:2 * H * D], qkv[:, 2 * H * D:]
o = torch.empty(P, H * D, device=dev, dtype=torch.bfloat16)
lse = torch.empty(N, H, S, device=dev)
do = torch.randn(P, H * D, device=dev).bfloat16()
dqkv = torch.empty_like(qkv)
delta = torch.empty(N, H, S, device=dev)
dq_acc = torch.zeros(P, H * D, device=dev)
L.causal_attn_fwd(q, k, v, o, lse, N, S, H, D, D, False)
bwd = lambda: L.causal_attn_bwd(q, k, v, o, do, lse, delta, dq_acc, dqkv[:, :H * D], dqkv[:, H * D:2 * H * D], dqkv[:, 2 * H * D:], N, S, H,
                                D, D, False)
bwd(); torch.cuda.synchronize()
trace = torch.zeros(4 * 4096, dtype=torch.int64, device=dev)
lib = L.load()
lib.pg_debug_set_trace.argtypes = [ctypes.c_void_p]
lib.pg_debug_set_trace.restype = None
lib.pg_debug_set_trace(trace.data_ptr())
bwd(); torch.cuda.synchronize()
lib.pg_debug_set_trace(None)
t = trace.cpu().view(4, 4096)
names = {0: ("producer", 2, ["wait q_empty", "-> wait do_empty"]),
         1: ("mma", 6, ["wait p_full", "issue S(next), dV", "wait ds_full", "issue dK, dP(next)", "wait dq_empty", "issue dQ + loop"]),
         2: ("drain", 2, ["wait dq_full", "drain work"]),
         3: ("softmax w0", 8, ["wait s_full", "LDTM+exp (A)", "wait p_free", "store P+fence+arrive", "wait dp_full", "LDTM+dS (B)",
                               "wait ds_free", "store dS+fence+arrive+loop"])}
for r, (name, per, labels) in names.items():
    x = t[r]
    n = int((x != 0).sum())
    tiles = n // per
    if tiles < 3:
        print(name, "no data"); continue
    x = x[:tiles * per].view(tiles, per).double()
    total = (x[-1, 0] - x[0, 0]) / (tiles - 1)
    print(f"== {name}: {tiles} tiles, {total:.0f} cycles / tile")
    for j in range(per):
        nxt = x[:, j + 1] if j + 1 < per else torch.cat((x[1:, 0], x[-1:, 0]))
        d = (nxt - x[:, j])[:-1]
        print(f"   {labels[j]:32s} mean {d.mean():8.0f}   median {d.median():8.0f}   max {d.max():8.0f}")
