"""How fast can ONE output file take bytes from memory on this box?  (The command line's limit: every label has one
output file, detect.py:295-298, and the page-cache write path of a file is serialised by its inode lock.)
Writes 1 GiB with a single write(), with k parallel pwrite() slices and through a shared mmap filled by k threads.
    python tools/write_probe.py [directory]        (CPU only)"""
import mmap
import os
import sys
import time
from concurrent.futures import ThreadPoolExecutor

import numpy as np

N = 1 << 30
d = sys.argv[1] if len(sys.argv) > 1 else "/tmp"
path = os.path.join(d, "rd_write_probe.bin")
buf = np.random.default_rng(0).integers(0, 255, N, dtype=np.uint8)
mv = memoryview(buf)


def t_write():
    with open(path, "wb", buffering=0) as f:
        t = time.perf_counter()
        f.write(mv)
        return time.perf_counter() - t


def t_pwrite(k):
    fd = os.open(path, os.O_WRONLY | os.O_CREAT | os.O_TRUNC)
    step = N // k
    t = time.perf_counter()
    with ThreadPoolExecutor(k) as ex:
        list(ex.map(lambda i: os.pwrite(fd, mv[i * step:(i + 1) * step], i * step), range(k)))
    dt = time.perf_counter() - t
    os.close(fd)
    return dt


def t_mmap(k):
    fd = os.open(path, os.O_RDWR | os.O_CREAT | os.O_TRUNC)
    os.ftruncate(fd, N)
    t = time.perf_counter()
    m = mmap.mmap(fd, N)
    dst = np.frombuffer(m, np.uint8)
    step = N // k
    with ThreadPoolExecutor(k) as ex:
        list(ex.map(lambda i: np.copyto(dst[i * step:(i + 1) * step], buf[i * step:(i + 1) * step]), range(k)))
    del dst
    m.close()
    dt = time.perf_counter() - t
    os.close(fd)
    return dt


print("1 GiB into one file under %s (%d cores)" % (d, os.cpu_count()))
for name, fn in (("write() x1", t_write), ("pwrite x4", lambda: t_pwrite(4)), ("pwrite x8", lambda: t_pwrite(8)),
                 ("mmap fill x1", lambda: t_mmap(1)), ("mmap fill x8", lambda: t_mmap(8))):
    dt = min(fn() for _ in range(2))
    print("%-14s %.2f GB/s" % (name, N / dt / 1e9), flush=True)
    os.unlink(path)
