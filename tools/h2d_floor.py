"""Host-to-device copy bandwidth of the config-2 batch (742 MB) from pinned memory: the floor of the
end-to-end number.  python tools/h2d_test.py"""
import ctypes, time
import torch
rt = ctypes.CDLL("libcudart.so.12")
n = 742_319_760
torch.cuda.init(); torch.zeros(1).cuda()
d = torch.empty(n, dtype=torch.uint8, device="cuda")
for flags, name in ((0, "default"), (4, "write-combined"), (1, "portable")):
    p = ctypes.c_void_p()
    assert rt.cudaHostAlloc(ctypes.byref(p), ctypes.c_size_t(n), ctypes.c_uint(flags)) == 0
    ctypes.memset(p, 1, n)
    s = torch.cuda.Stream()
    for chunks in (1, 12):
        best = 1e9
        for _ in range(4):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            sz = n // chunks
            for c in range(chunks):
                rt.cudaMemcpyAsync(ctypes.c_void_p(d.data_ptr() + c * sz), ctypes.c_void_p(p.value + c * sz), ctypes.c_size_t(sz), ctypes.c_int(1), ctypes.c_void_p(s.cuda_stream))
            s.synchronize()
            best = min(best, time.perf_counter() - t0)
        print(f"{name:15s} chunks={chunks:2d}: {n / best / 1e9:.1f} GB/s ({best*1e3:.2f} ms)", flush=True)
    rt.cudaFreeHost(p)
