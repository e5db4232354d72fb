"""Do small host->device copies on one stream queue behind large device->host copies on another?  (B200, PCIe 5)"""
import torch
dev = torch.device("cuda")
big_d = torch.empty(32 << 20, dtype=torch.uint8, device=dev)
big_h = torch.empty(32 << 20, dtype=torch.uint8).pin_memory()
small_h = torch.zeros(2048, dtype=torch.uint8).pin_memory()
small_d = torch.empty(2048, dtype=torch.uint8, device=dev)
work = torch.empty(64 << 20, dtype=torch.uint8, device=dev)
sa, sb = torch.cuda.Stream(), torch.cuda.Stream()


def loop_b(n, with_h2d, with_memset):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(sb):
        e0.record()
        for _ in range(n):
            if with_h2d:
                for _ in range(6):
                    small_d.copy_(small_h, non_blocking=True)
            if with_memset:
                small_d.zero_()
            work.add_(1)            # ~64 MB r+w kernel, ~25 us
        e1.record()
    return e0, e1


for traffic in (False, True):
    for with_h2d in (False, True):
        torch.cuda.synchronize()
        if traffic:
            with torch.cuda.stream(sa):
                for _ in range(40):
                    for k in range(12):
                        big_h[k * (2 << 20):(k + 1) * (2 << 20) + (700 << 10)].copy_(big_d[k * (2 << 20):(k + 1) * (2 << 20) + (700 << 10)], non_blocking=True)
        e0, e1 = loop_b(50, with_h2d, False)
        torch.cuda.synchronize()
        print(f"D2H traffic on another stream: {traffic!s:5}  6 small H2D per iteration: {with_h2d!s:5}  -> {e0.elapsed_time(e1) / 50 * 1e3:8.1f} us per iteration")
