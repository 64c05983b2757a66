"""What a store-only / store-mostly stream reaches on this GPU (the roofline of stem / bicubic / head, which write 8-30 x
what they read), beside the copy figure of MEASURED_PEAKS.json:  python tools/write_peak.py"""
import torch


def timeit(fn, reps=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best * 1e-3


def main():
    dev = torch.device("cuda", 0)
    for mb in (400, 2048):
        n = mb * (1 << 20) // 4
        a = torch.empty(n, device=dev)
        b = torch.empty(n, device=dev)
        t = timeit(lambda: a.zero_())
        print(f"{mb:5d} MiB  memset (zero_)          {n * 4 / t / 1e9:8.1f} GB/s written")
        t = timeit(lambda: a.fill_(1.5))
        print(f"{mb:5d} MiB  fill_ (store kernel)    {n * 4 / t / 1e9:8.1f} GB/s written")
        t = timeit(lambda: b.copy_(a))
        print(f"{mb:5d} MiB  copy_                   {2 * n * 4 / t / 1e9:8.1f} GB/s read + written")
        small = a[: n // 8]
        t = timeit(lambda: torch.add(small.view(-1, 1).expand(-1, 8), 1.0, out=b.view(-1, 8)))
        print(f"{mb:5d} MiB  1 read : 8 written      {(n * 4 + n // 8 * 4) / t / 1e9:8.1f} GB/s read + written")


if __name__ == "__main__":
    main()
