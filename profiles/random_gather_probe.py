import torch, time
torch.cuda.init()
for elsz, dt in ((8, torch.int64), (4, torch.int32)):
    N = (17 << 30) // elsz
    x = torch.empty(N, dtype=dt, device='cuda'); x.zero_()
    M = 1 << 28
    idx = torch.randint(0, N, (M,), device='cuda')
    for rep in range(3):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        y = x[idx]
        torch.cuda.synchronize(); dt_ = time.perf_counter() - t0
    print("random %dB gathers from 17GB: %.1f G/s (%.1f ms for %d)" % (elsz, M / dt_ / 1e9, dt_ * 1e3, M))
    del x, y, idx
