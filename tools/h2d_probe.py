"""PCIe H2D probe: copy granularity and stream concurrency (developer tool)."""
import time, torch
dev = torch.device("cuda", 0)
N = 90_316_800
src = torch.empty(N, dtype=torch.uint8).pin_memory(); dst = torch.empty(N, dtype=torch.uint8, device=dev)
src2 = torch.empty(N, dtype=torch.uint8, pin_memory=True)
def timeit(fn, n=10):
    fn(); torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(n): fn()
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / n
for name, s in (("pin_memory()", src), ("empty(pin_memory=True)", src2)):
    dt = timeit(lambda: dst.copy_(s, non_blocking=True)); print(f"{name}: single 90MB {dt*1e3:.3f} ms {N/dt/1e9:.1f} GB/s")
for chunks in (4, 32, 128):
    sv, dv = src.chunk(chunks), dst.chunk(chunks)
    dt = timeit(lambda: [d.copy_(s, non_blocking=True) for s, d in zip(sv, dv)]); print(f"{chunks} chunks 1 stream {dt*1e3:.3f} ms {N/dt/1e9:.1f} GB/s")
for ns in (2, 4):
    streams = [torch.cuda.Stream() for _ in range(ns)]
    sv, dv = src.chunk(32), dst.chunk(32)
    def f():
        for i, (s, d) in enumerate(zip(sv, dv)):
            with torch.cuda.stream(streams[i % ns]): d.copy_(s, non_blocking=True)
    dt = timeit(f); print(f"32 chunks {ns} streams {dt*1e3:.3f} ms {N/dt/1e9:.1f} GB/s")
import os; print("cpus", os.cpu_count()); os.system("nvidia-smi topo -m | head -8; cat /sys/bus/pci/devices/*/numa_node 2>/dev/null | sort | uniq -c | head")
