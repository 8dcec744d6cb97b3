import torch, time
n = 1400*1024*1024
h_in = torch.empty(n, dtype=torch.uint8).pin_memory()
h_out = torch.empty(n, dtype=torch.uint8).pin_memory()
d_in = torch.empty(n, dtype=torch.uint8, device='cuda')
d_out = torch.empty(n, dtype=torch.uint8, device='cuda')
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def t(fn, reps=3):
    best=1e9
    for _ in range(reps):
        torch.cuda.synchronize(); t0=time.perf_counter(); fn(); torch.cuda.synchronize(); best=min(best,time.perf_counter()-t0)
    return best
a=t(lambda: d_in.copy_(h_in, non_blocking=True))
b=t(lambda: h_out.copy_(d_out, non_blocking=True))
def both():
    with torch.cuda.stream(s1): d_in.copy_(h_in, non_blocking=True)
    with torch.cuda.stream(s2): h_out.copy_(d_out, non_blocking=True)
c=t(both)
print("H2D %.1f GB/s  D2H %.1f GB/s  both: %.1f GB/s each way (%.1f ms)"%(n/a/1e9, n/b/1e9, n/c/1e9, c*1e3))
# chunked 8.9MB
ch = 65536*136
def chunked():
    with torch.cuda.stream(s1):
        for o in range(0, n, ch): d_in[o:o+ch].copy_(h_in[o:o+ch], non_blocking=True)
d=t(chunked)
print("H2D chunked 8.9MB: %.1f GB/s"%(n/d/1e9))
