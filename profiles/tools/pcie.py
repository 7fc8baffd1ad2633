import torch, time
dev = torch.device("cuda:0")
for mb in (16, 64, 512):
    n = mb << 20
    h1 = torch.empty(n, dtype=torch.uint8).pin_memory(); h2 = torch.empty(n, dtype=torch.uint8).pin_memory()
    d1 = torch.empty(n, dtype=torch.uint8, device=dev); d2 = torch.empty(n, dtype=torch.uint8, device=dev)
    s1 = torch.cuda.Stream(); s2 = torch.cuda.Stream()
    def run(up, down, reps=10):
        torch.cuda.synchronize(); t = time.perf_counter()
        for _ in range(reps):
            if up:
                with torch.cuda.stream(s1): d1.copy_(h1, non_blocking=True)
            if down:
                with torch.cuda.stream(s2): h2.copy_(d2, non_blocking=True)
        torch.cuda.synchronize(); return (time.perf_counter() - t) / reps
    run(1, 1, 2)
    a = run(1, 0); b = run(0, 1); c = run(1, 1)
    print(f"{mb} MB: H2D {n/a/1e9:.1f} GB/s, D2H {n/b/1e9:.1f} GB/s, both {n/c/1e9:.1f} GB/s each way")
# pageable for comparison
h = torch.empty(512 << 20, dtype=torch.uint8); d = torch.empty(512 << 20, dtype=torch.uint8, device=dev)
torch.cuda.synchronize(); t = time.perf_counter(); d.copy_(h); torch.cuda.synchronize(); print("pageable H2D GB/s", h.numel() / (time.perf_counter() - t) / 1e9)
