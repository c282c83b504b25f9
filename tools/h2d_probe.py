"""H2D copy rate from pinned memory: one stream vs two concurrent streams (is 44 GB/s the link or one copy engine?)."""
import json
import torch
n = 1 << 29  # 4 GiB of complex128 = 2^28 amps; use bytes
host = torch.empty(4 << 30, dtype=torch.uint8, pin_memory=True)
dev = torch.empty(4 << 30, dtype=torch.uint8, device='cuda')
def timed(fn, reps=3):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
one = timed(lambda: dev.copy_(host, non_blocking=True))
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
half = host.numel() // 2
def two():
    cur = torch.cuda.current_stream()
    s1.wait_stream(cur); s2.wait_stream(cur)
    with torch.cuda.stream(s1): dev[:half].copy_(host[:half], non_blocking=True)
    with torch.cuda.stream(s2): dev[half:].copy_(host[half:], non_blocking=True)
    cur.wait_stream(s1); cur.wait_stream(s2)
two_ms = timed(two)
def chunks():
    c = host.numel() // 16
    for k in range(16): dev[k*c:(k+1)*c].copy_(host[k*c:(k+1)*c], non_blocking=True)
ch_ms = timed(chunks)
d2h = timed(lambda: host.copy_(dev, non_blocking=True))
gb = host.numel() / 1e9
print(json.dumps({'h2d_one_stream_gbs': gb / one * 1e3, 'h2d_two_streams_gbs': gb / two_ms * 1e3, 'h2d_16_chunks_gbs': gb / ch_ms * 1e3, 'd2h_gbs': gb / d2h * 1e3}))
