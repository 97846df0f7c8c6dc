#!/usr/bin/env python
"""Reference points for the roofline (GPU box): HBM throughput of a pure copy (what MEASURED_PEAKS.json holds), a pure
fill, and a stream with the step's read:write mix (1 byte read per 2.2 written), all over buffers >> L2."""
import json, torch
dev = torch.device("cuda:0")
n = 1 << 30  # 1 Gi floats = 4 GiB
a = torch.empty(n, dtype=torch.float32, device=dev).fill_(1.0)
b = torch.empty(n, dtype=torch.float32, device=dev)


def best(fn, bytes_moved, reps=8):
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(); e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return bytes_moved / (min(ts) * 1e-3) / 1e9


out = {"copy_gbs": best(lambda: b.copy_(a), 8 * n), "fill_gbs": best(lambda: b.fill_(2.0), 4 * n),
       "read_gbs": best(lambda: a.sum(), 4 * n)}
# read n/2.2 floats, write n floats: b[:n] = a[:m].repeat-like via index_select is not a stream; use two ops on one stream
m = int(n / 2.2)
out["mix_1r_2p2w_gbs"] = best(lambda: (b.fill_(3.0), a[:m].sum()), 4 * n + 4 * m)
print(json.dumps(out))
