"""Device -> host copy of a result array: fresh page-locked staging buffer vs pageable destination (first use and reuse)."""
import time, torch
torch.cuda.init(); torch.zeros(1, device="cuda")
for mb in (4, 16, 64, 256, 1024):
    n = mb * (1 << 20) // 8
    x = torch.zeros(n, dtype=torch.float64, device="cuda"); torch.cuda.synchronize()
    t0 = time.perf_counter(); h = torch.empty(n, dtype=torch.float64, pin_memory=True); t_alloc = time.perf_counter() - t0
    t0 = time.perf_counter(); h.copy_(x, non_blocking=True); torch.cuda.synchronize(); t_copy = time.perf_counter() - t0
    t0 = time.perf_counter(); y = x.cpu(); t_page = time.perf_counter() - t0
    t0 = time.perf_counter(); y = x.cpu(); t_page2 = time.perf_counter() - t0
    del h; t0 = time.perf_counter(); h = torch.empty(n, dtype=torch.float64, pin_memory=True); t_realloc = time.perf_counter() - t0
    print(f"{mb:5d} MB: pinned alloc {t_alloc*1e3:8.1f} ms + copy {t_copy*1e3:7.1f} ms (re-alloc from cache {t_realloc*1e3:6.2f} ms) | pageable .cpu() {t_page*1e3:8.1f} ms, again {t_page2*1e3:8.1f} ms")
    del h, x, y
