"""PCIe probe for the end-to-end path: pinned-host copies H2D, D2H and both at once (two streams, the two copy engines)."""
import time
import torch
n_up, n_dn = 25 << 20, 30 << 20
h_up = torch.empty(n_up, dtype=torch.uint8).pin_memory(); h_dn = torch.empty(n_dn, dtype=torch.uint8).pin_memory()
d_up = torch.empty(n_up, dtype=torch.uint8, device="cuda"); d_dn = torch.empty(n_dn, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def run(up, dn, reps=50):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(reps):
        if up:
            with torch.cuda.stream(s1): d_up.copy_(h_up, non_blocking=True)
        if dn:
            with torch.cuda.stream(s2): h_dn.copy_(d_dn, non_blocking=True)
    torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / reps
    return dt
for name, up, dn in (("H2D 25 MiB", 1, 0), ("D2H 30 MiB", 0, 1), ("both", 1, 1)):
    dt = run(up, dn)
    b = (n_up if up else 0) + (n_dn if dn else 0)
    print("%-12s %.3f ms  %.1f GB/s" % (name, dt * 1e3, b / dt / 1e9))
# chunked pipeline: 8 slices, H2D slice k+1 while D2H slice k
def pipelined(slices=8, reps=30):
    cu, cd = n_up // slices, n_dn // slices
    ev = [torch.cuda.Event() for _ in range(slices)]
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(reps):
        for k in range(slices):
            with torch.cuda.stream(s1):
                d_up[k * cu:(k + 1) * cu].copy_(h_up[k * cu:(k + 1) * cu], non_blocking=True); ev[k].record(s1)
            with torch.cuda.stream(s2):
                s2.wait_event(ev[k]); h_dn[k * cd:(k + 1) * cd].copy_(d_dn[k * cd:(k + 1) * cd], non_blocking=True)
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / reps
dt = pipelined(); print("pipelined 8 slices (H2D k+1 || D2H k): %.3f ms  %.1f GB/s combined" % (dt * 1e3, (n_up + n_dn) / dt / 1e9))
