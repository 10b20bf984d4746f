import time, ctypes as C, numpy as np, torch, threading
cudart = C.CDLL("libcudart.so.12")
n = 256 << 20
a = np.ones(n, np.uint8); a[::4096] = 2
torch.cuda.init(); d = torch.empty(n, dtype=torch.uint8, device="cuda")
for rep in range(3):
    t0 = time.perf_counter(); rc = cudart.cudaHostRegister(C.c_void_p(a.ctypes.data), C.c_size_t(n), C.c_uint(0)); t1 = time.perf_counter()
    t2 = time.perf_counter(); rc2 = cudart.cudaHostUnregister(C.c_void_p(a.ctypes.data)); t3 = time.perf_counter()
    print("register %d MiB: %.2f ms rc %d, unregister %.2f ms rc %d" % (n >> 20, (t1 - t0) * 1e3, rc, (t3 - t2) * 1e3, rc2))
# pageable copy rate
t = torch.from_numpy(a)
for rep in range(2):
    torch.cuda.synchronize(); t0 = time.perf_counter(); d.copy_(t); torch.cuda.synchronize(); t1 = time.perf_counter()
    print("pageable H2D %.1f GB/s" % (n / (t1 - t0) / 1e9))
h = torch.empty(n, dtype=torch.uint8)
for rep in range(2):
    torch.cuda.synchronize(); t0 = time.perf_counter(); h.copy_(d); torch.cuda.synchronize(); t1 = time.perf_counter()
    print("pageable D2H %.1f GB/s" % (n / (t1 - t0) / 1e9))
p = torch.empty(n, dtype=torch.uint8).pin_memory()
for nt in (1, 2, 4, 8):
    parts = np.array_split(np.arange(n), nt)
    def work(i):
        lo, hi = parts[i][0], parts[i][-1] + 1
        p.numpy()[lo:hi] = a[lo:hi]
    t0 = time.perf_counter()
    th = [threading.Thread(target=work, args=(i,)) for i in range(nt)]
    [x.start() for x in th]; [x.join() for x in th]
    t1 = time.perf_counter()
    print("memcpy to pinned, %d threads: %.1f GB/s" % (nt, n / (t1 - t0) / 1e9))
