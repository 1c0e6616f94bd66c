"""Peer-to-peer copy bandwidth between the GPUs of the box (torch copy_ of 1 GiB, best of 5) + topology."""
import subprocess, time, torch
n = torch.cuda.device_count()
print(subprocess.run(["nvidia-smi", "topo", "-m"], capture_output=True, text=True).stdout[:1500])
for a in range(min(n, 2)):
    for b in range(n):
        if a == b: continue
        x = torch.empty(1 << 30, dtype=torch.uint8, device="cuda:%d" % a)
        y = torch.empty(1 << 30, dtype=torch.uint8, device="cuda:%d" % b)
        best = 1e9
        for _ in range(5):
            torch.cuda.synchronize(a); torch.cuda.synchronize(b)
            t0 = time.perf_counter(); y.copy_(x); torch.cuda.synchronize(a); torch.cuda.synchronize(b)
            best = min(best, time.perf_counter() - t0)
        print("copy %d -> %d: %.1f GB/s (can_access_peer %s)" % (a, b, (1 << 30) / best / 1e9, torch.cuda.can_device_access_peer(a, b)))
