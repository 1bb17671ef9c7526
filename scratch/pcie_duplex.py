"""Ceiling of the end-to-end figure: pinned-memory PCIe bandwidth of this box, one direction at a time and both at once (1 GB messages)."""
import torch, time
n = 1007144768 // 8
h_in = torch.empty(n, dtype=torch.float64).pin_memory(); h_out = torch.empty(n, dtype=torch.float64).pin_memory()
d_in = torch.empty(n, dtype=torch.float64, device="cuda"); d_out = torch.empty(n, dtype=torch.float64, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def run(h2d, d2h, reps=8):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(reps):
        if h2d:
            with torch.cuda.stream(s1): d_in.copy_(h_in, non_blocking=True)
        if d2h:
            with torch.cuda.stream(s2): h_out.copy_(d_out, non_blocking=True)
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps
run(True, True, 2)
for name, a, b in (("H2D alone", True, False), ("D2H alone", False, True), ("both at once", True, True)):
    t = run(a, b)
    print("%-13s %.2f ms per 1.007 GB message -> %.1f GB/s per direction" % (name, t * 1e3, n * 8 / t / 1e9))
