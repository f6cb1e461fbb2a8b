"""Where do the pinned host buffers of the e2e leg live?  Concurrent H2D + D2H bandwidth of 1 GiB pinned buffers allocated
(a) as bench.py does at N = 1 (no binding) and (b) after binding the process to the CPUs NVML lists as local to GPU 0."""
import glob, os, time
import torch

def nodes():
    out = {}
    for d in sorted(glob.glob("/sys/devices/system/node/node[0-9]*")):
        out[os.path.basename(d)] = open(d + "/cpulist").read().strip()
    return out

def gpu_cpus():
    import pynvml
    pynvml.nvmlInit()
    h = pynvml.nvmlDeviceGetHandleByIndex(0)
    words = pynvml.nvmlDeviceGetCpuAffinity(h, (os.cpu_count() + 63) // 64)
    return {64 * w + b for w, word in enumerate(words) for b in range(64) if (word >> b) & 1}

def bw(tag):
    n = 1 << 28
    hin = torch.empty(n, dtype=torch.float32).pin_memory(); hin.fill_(1.0)
    hout = torch.empty(n, dtype=torch.float32).pin_memory(); hout.fill_(0.0)
    din = torch.empty(n, dtype=torch.float32, device="cuda"); dout = torch.ones(n, dtype=torch.float32, device="cuda")
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    for mode in ("h2d", "d2h", "both"):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        for _ in range(5):
            if mode in ("h2d", "both"):
                with torch.cuda.stream(s1): din.copy_(hin, non_blocking=True)
            if mode in ("d2h", "both"):
                with torch.cuda.stream(s2): hout.copy_(dout, non_blocking=True)
        torch.cuda.synchronize(); dt = time.perf_counter() - t0
        print(f"{tag} {mode}: {5 * n * 4 / dt / 1e9:.1f} GB/s per direction", flush=True)

print("numa nodes:", nodes())
print("allowed cpus:", sorted(os.sched_getaffinity(0)))
try:
    g = gpu_cpus(); print("GPU 0 local cpus (NVML):", sorted(g)[:4], "...", len(g))
except Exception as e:
    g = set(); print("nvml:", e)
bw("unbound")
both = g & os.sched_getaffinity(0)
print("local and allowed:", len(both))
if both and both != os.sched_getaffinity(0):
    os.sched_setaffinity(0, both); bw("bound-local")
other = os.sched_getaffinity(0) - g
