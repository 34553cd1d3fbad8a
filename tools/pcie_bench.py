"""tools/pcie_bench.py -- what the host link gives on this box: pinned H2D, D2H, and both at once (two streams), 256 MiB
transfers, so that the host-tensor (`e2e`) numbers of bench.py can be read against their own roofline."""
import torch, time
n = 256 << 20
h_in, h_out = torch.empty(n, dtype=torch.uint8).pin_memory(), torch.empty(n, dtype=torch.uint8).pin_memory()
d_in, d_out = torch.empty(n, dtype=torch.uint8, device="cuda"), torch.empty(n, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()

def run(up, down, reps=10):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(reps):
        if up:
            with torch.cuda.stream(s1): d_in.copy_(h_in, non_blocking=True)
        if down:
            with torch.cuda.stream(s2): h_out.copy_(d_out, non_blocking=True)
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / reps

for _ in range(2): run(True, True)
tu, td, tb = run(True, False), run(False, True), run(True, True)
print(f"H2D {n / tu / 1e9:.1f} GB/s   D2H {n / td / 1e9:.1f} GB/s   both at once: {n / tb / 1e9:.1f} GB/s each way ({2 * n / tb / 1e9:.1f} GB/s total)")
