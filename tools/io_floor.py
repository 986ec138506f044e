"""What the box's page cache can do for one output file: single-thread and 4-thread buffered pwrite of
2.2 GB into a fresh file, and a 4-thread memcpy into a MAP_SHARED mapping (the two ways to land bytes)."""
import mmap, os, threading, time
import numpy as np
N = 2_200_000_000
buf = np.random.default_rng(0).integers(0, 255, 64 << 20, dtype=np.uint8).tobytes()
path = os.path.join(os.environ.get("TMPDIR", "/tmp"), "io_floor.bin")
def fresh():
    if os.path.exists(path): os.unlink(path)
    fd = os.open(path, os.O_RDWR | os.O_CREAT | os.O_TRUNC, 0o644); os.ftruncate(fd, N); return fd
def pw(fd, lo, hi):
    o = lo
    while o < hi:
        n = min(len(buf), hi - o); os.pwrite(fd, buf[:n], o); o += n
for nt in (1, 4):
    fd = fresh(); t = time.time()
    th = [threading.Thread(target=pw, args=(fd, N * k // nt, N * (k + 1) // nt)) for k in range(nt)]
    [x.start() for x in th]; [x.join() for x in th]
    dt = time.time() - t; os.close(fd)
    print(f"buffered pwrite, {nt} thread(s): {dt:.3f} s  {N/dt/1e9:.2f} GB/s", flush=True)
for nt in (4, 8, 16):
    fd = fresh(); mm = mmap.mmap(fd, N); view = np.frombuffer(mm, dtype=np.uint8); src = np.frombuffer(buf, dtype=np.uint8)
    def mc(lo, hi):
        o = lo
        while o < hi:
            n = min(len(src), hi - o); view[o:o + n] = src[:n]; o += n
    t = time.time()
    th = [threading.Thread(target=mc, args=(N * k // nt, N * (k + 1) // nt)) for k in range(nt)]
    [x.start() for x in th]; [x.join() for x in th]
    dt = time.time() - t
    print(f"memcpy into MAP_SHARED mapping, {nt} threads: {dt:.3f} s  {N/dt/1e9:.2f} GB/s", flush=True)
    del view; mm.close(); os.close(fd)
for nt in (8, 16):
    fd = fresh(); t = time.time()
    th = [threading.Thread(target=pw, args=(fd, N * k // nt, N * (k + 1) // nt)) for k in range(nt)]
    [x.start() for x in th]; [x.join() for x in th]
    dt = time.time() - t; os.close(fd)
    print(f"buffered pwrite, {nt} thread(s): {dt:.3f} s  {N/dt/1e9:.2f} GB/s", flush=True)
os.unlink(path)
