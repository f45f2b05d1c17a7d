"""H2D copy rate of one 51.4 MB minibatch (16384 x 784 fp32) from page-locked host memory: default pinned vs write-combined,
one stream vs the copy split over two streams.  python tools/h2d_probe.py"""
import ctypes as C, time
from cuda import cudart


def chk(r):
    if isinstance(r, tuple):
        err, rest = r[0], r[1:]
    else:
        err, rest = r, ()
    assert err == cudart.cudaError_t.cudaSuccess, err
    return rest[0] if len(rest) == 1 else rest


n = 16384 * 784 * 4
chk(cudart.cudaSetDevice(0))
dev = chk(cudart.cudaMalloc(n))
s0 = chk(cudart.cudaStreamCreate()); s1 = chk(cudart.cudaStreamCreate())
e0 = chk(cudart.cudaEventCreate()); e1 = chk(cudart.cudaEventCreate())
for name, flags in (("pinned", cudart.cudaHostAllocDefault), ("write-combined", cudart.cudaHostAllocWriteCombined)):
    host = chk(cudart.cudaHostAlloc(n, flags))
    C.memset(host, 1, n)
    for split in (1, 2, 4):
        for it in range(2):
            chk(cudart.cudaEventRecord(e0, s0))
            reps = 20
            for r in range(reps):
                if split == 1:
                    chk(cudart.cudaMemcpyAsync(dev, host, n, cudart.cudaMemcpyKind.cudaMemcpyHostToDevice, s0))
                else:
                    part = n // split
                    chk(cudart.cudaStreamWaitEvent(s1, e0, 0))
                    for k in range(split):
                        chk(cudart.cudaMemcpyAsync(dev + k * part, host + k * part, part,
                                                   cudart.cudaMemcpyKind.cudaMemcpyHostToDevice, s0 if k % 2 == 0 else s1))
                    chk(cudart.cudaEventRecord(e1, s1)); chk(cudart.cudaStreamWaitEvent(s0, e1, 0))
            chk(cudart.cudaEventRecord(e1, s0)); chk(cudart.cudaEventSynchronize(e1))
            ms = chk(cudart.cudaEventElapsedTime(e0, e1))
        print("%-15s split %d: %.1f GB/s (%.3f ms per 51.4 MB)" % (name, split, reps * n / ms / 1e6, ms / reps), flush=True)
    chk(cudart.cudaFreeHost(host))
