#!/usr/bin/env python
"""Measured HBM ceilings for a pure WRITE stream (what K1 is: 1.22 GB written, 0.14 GB read per launch)
next to the read+write copy figure the roofline uses.  torch is only the timer / allocator here.
  python tools/hbm_write_peak.py > profiles/rNN_hbm_write_peak.txt"""
import torch

def timeit(fn, n=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / n

def main():
    nbytes = 4 << 30
    x = torch.empty(nbytes // 8, dtype=torch.float64, device="cuda")
    y = torch.empty_like(x)
    t_fill = timeit(lambda: x.fill_(1.0))
    t_zero = timeit(lambda: x.zero_())
    t_copy = timeit(lambda: y.copy_(x))
    t_read = timeit(lambda: x.sum())
    print(f"device: {torch.cuda.get_device_name(0)}")
    print(f"fill_  (write only, 4 GiB):        {t_fill:.3f} ms  {nbytes / t_fill / 1e6:.0f} GB/s")
    print(f"zero_  (memset, 4 GiB):            {t_zero:.3f} ms  {nbytes / t_zero / 1e6:.0f} GB/s")
    print(f"copy_  (read + write, 2 x 4 GiB):  {t_copy:.3f} ms  {2 * nbytes / t_copy / 1e6:.0f} GB/s")
    print(f"sum    (read only, 4 GiB):         {t_read:.3f} ms  {nbytes / t_read / 1e6:.0f} GB/s")

if __name__ == "__main__":
    main()
