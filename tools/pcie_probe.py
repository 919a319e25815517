"""PCIe probe: H2D alone, D2H alone, both at once on two streams (pinned memory).  Explains the e2e floor."""
import json, sys, torch
gb = float(sys.argv[1]) if len(sys.argv) > 1 else 2.0
n = int(gb * (1 << 30))
h_in = torch.empty(n, dtype=torch.uint8, pin_memory=True); h_out = torch.empty(n, dtype=torch.uint8, pin_memory=True)
d_in = torch.empty(n, dtype=torch.uint8, device="cuda"); d_out = torch.ones(n, dtype=torch.uint8, device="cuda")
s1 = torch.cuda.Stream(); s2 = torch.cuda.Stream()
def timed(fn, reps=3):
    best = 1e9
    for _ in range(reps):
        torch.cuda.synchronize()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        s1.wait_event(e0); s2.wait_event(e0)
        fn()
        e_a = torch.cuda.Event(); e_b = torch.cuda.Event(); e_a.record(s1); e_b.record(s2)
        torch.cuda.current_stream().wait_event(e_a); torch.cuda.current_stream().wait_event(e_b)
        e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best
def h2d():
    with torch.cuda.stream(s1): d_in.copy_(h_in, non_blocking=True)
def d2h():
    with torch.cuda.stream(s2): h_out.copy_(d_out, non_blocking=True)
def both():
    h2d(); d2h()
def chunked_both(k=32):
    c = n // k
    for i in range(k):
        with torch.cuda.stream(s1): d_in[i*c:(i+1)*c].copy_(h_in[i*c:(i+1)*c], non_blocking=True)
        with torch.cuda.stream(s2): h_out[i*c:(i+1)*c].copy_(d_out[i*c:(i+1)*c], non_blocking=True)
t1 = timed(h2d); t2 = timed(d2h); t3 = timed(both); t4 = timed(chunked_both)
print(json.dumps({"GiB_each": gb, "h2d_ms": t1, "h2d_GBps": n / t1 / 1e6, "d2h_ms": t2, "d2h_GBps": n / t2 / 1e6,
                  "both_ms": t3, "both_aggregate_GBps": 2 * n / t3 / 1e6, "chunked_both_ms": t4,
                  "duplex_speedup_vs_serial": (t1 + t2) / t3}))
