"""Measured HBM bandwidth of a write-only and a copy pass (torch ops, CUDA events, best of 10): the yard-stick
for the store-bound layers (first convolution, K = 80 FC layer), whose traffic is almost all writes."""
import torch

def best(fn, n=10):
    ts = []
    for _ in range(n):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return min(ts)

def main():
    n = 1 << 31                                   # 4 GiB of bf16: far beyond the 126 MB L2
    x = torch.empty(n, dtype=torch.bfloat16, device='cuda')
    y = torch.empty(n, dtype=torch.bfloat16, device='cuda')
    x.zero_(); y.zero_(); torch.cuda.synchronize()
    w = best(lambda: x.fill_(1.0))
    print('write-only  (fill_)   %.1f GB/s' % (2 * n / w / 1e6))
    w = best(lambda: torch.cuda.current_stream().synchronize() or x.zero_())
    print('write-only  (memset)  %.1f GB/s' % (2 * n / w / 1e6))
    c = best(lambda: y.copy_(x))
    print('copy        (copy_)   %.1f GB/s read+write' % (4 * n / c / 1e6))

if __name__ == '__main__':
    main()
