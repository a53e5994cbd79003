"""Does a host<->device copy overlap with an HBM-saturating kernel stream on this GPU?  (explains the pipelined e2e figure of bench.py)"""
import torch

n = 1 << 27  # 1 GiB of fp64
h_in = torch.empty(n, dtype=torch.float64).pin_memory()
h_out = torch.empty(n, dtype=torch.float64).pin_memory()
d_in = torch.empty(n, dtype=torch.float64, device="cuda")
d_out = torch.ones(n, dtype=torch.float64, device="cuda")
a = torch.ones(n, dtype=torch.float64, device="cuda")
b = torch.ones(n, dtype=torch.float64, device="cuda")
s_k, s_up, s_dn = torch.cuda.Stream(), torch.cuda.Stream(), torch.cuda.Stream()


def kernels(reps=100):
    with torch.cuda.stream(s_k):
        for _ in range(reps):
            torch.add(a, b, out=a)  # 24 B per element: HBM bound


def ev():
    return torch.cuda.Event(enable_timing=True)


def timed(label, do_k, do_c):
    torch.cuda.synchronize()
    k0, k1, c0, c1, c2, c3 = ev(), ev(), ev(), ev(), ev(), ev()
    if do_k:
        k0.record(s_k); kernels(); k1.record(s_k)
    if do_c:
        with torch.cuda.stream(s_up):
            c0.record(s_up); d_in.copy_(h_in, non_blocking=True); c1.record(s_up)
        with torch.cuda.stream(s_dn):
            c2.record(s_dn); h_out.copy_(d_out, non_blocking=True); c3.record(s_dn)
    torch.cuda.synchronize()
    msg = label + ":"
    if do_k:
        msg += f" kernels {k0.elapsed_time(k1):7.1f} ms"
    if do_c:
        msg += f" | H2D 1 GiB {c0.elapsed_time(c1):6.1f} ms, D2H 1 GiB {c2.elapsed_time(c3):6.1f} ms"
    print(msg, flush=True)


for _ in range(2):
    timed("kernels alone", True, False)
    timed("copies alone ", False, True)
    timed("both at once ", True, True)
