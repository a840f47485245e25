"""Does a chunked H2D -> kernel -> D2H pipeline overlap on this box?  (design probe for the host-vector path)"""
import time
import torch

n, K = 7880599, 8
xh = torch.empty(n, dtype=torch.float64).pin_memory(); yh = torch.empty(n, dtype=torch.float64).pin_memory()
xh.copy_(torch.arange(n, dtype=torch.float64))
xd = torch.empty(n, dtype=torch.float64, device="cuda"); yd = torch.empty(n, dtype=torch.float64, device="cuda")
big = torch.empty(3 * 10**8, dtype=torch.float64, device="cuda")          # stand-in for a 0.5 ms kernel: 8 x 0.07 ms
su, sc, sd = torch.cuda.Stream(), torch.cuda.Stream(), torch.cuda.Stream()
b = [n * c // K for c in range(K + 1)]
chunk = big.numel() // K // 4

def serial():
    with torch.cuda.stream(sc):
        xd.copy_(xh, non_blocking=True)
        for c in range(K):
            big[c * chunk:(c + 1) * chunk].mul_(1.0000001)
        torch.add(xd, 1.0, out=yd)
        yh.copy_(yd, non_blocking=True)

def piped(order):
    evu = [torch.cuda.Event() for _ in range(K)]; evk = [torch.cuda.Event() for _ in range(K)]
    if order == "up-first":
        for c in range(K):
            with torch.cuda.stream(su):
                xd[b[c]:b[c + 1]].copy_(xh[b[c]:b[c + 1]], non_blocking=True); evu[c].record(su)
    for c in range(K):
        if order != "up-first":
            with torch.cuda.stream(su):
                xd[b[c]:b[c + 1]].copy_(xh[b[c]:b[c + 1]], non_blocking=True); evu[c].record(su)
        with torch.cuda.stream(sc):
            sc.wait_event(evu[c])
            big[c * chunk:(c + 1) * chunk].mul_(1.0000001)
            torch.add(xd[b[c]:b[c + 1]], 1.0, out=yd[b[c]:b[c + 1]]); evk[c].record(sc)
        with torch.cuda.stream(sd):
            sd.wait_event(evk[c])
            yh[b[c]:b[c + 1]].copy_(yd[b[c]:b[c + 1]], non_blocking=True)

def t(fn, reps=10):
    fn(); torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(reps): fn()
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / reps * 1e3

print("serial ms", t(serial))
print("pipelined interleaved ms", t(lambda: piped("interleaved")))
print("pipelined uploads-first ms", t(lambda: piped("up-first")))
assert torch.equal(yh, xh + 1.0)
