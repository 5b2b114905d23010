"""Does a 2D (pitched) D2H copy of 4 KiB rows run slower than a 1D copy of the same bytes?  (The multichannel stream's
host path downloads K outputs per channel from a device slot with a different pitch.)  usage: python tools/r2_memcpy2d.py"""
import ctypes as C, time
import torch

rt = C.CDLL("libcudart.so.12")
rt.cudaMemcpy2DAsync.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_size_t, C.c_size_t, C.c_int, C.c_void_p]
rt.cudaMemcpyAsync.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_int, C.c_void_p]
rows, K, OP = 1 << 18, 1024, 1048
d = torch.randn(rows, OP, device="cuda")
h = torch.empty(rows, OP, dtype=torch.float32, pin_memory=True)
hx = torch.empty(rows, K, dtype=torch.float32, pin_memory=True)
D2H, H2D = 2, 1

def t(fn, reps=5):
    fn(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps * 1e3

gb = rows * K * 4 / 1e9
for name, fn in (
    ("D2H 1D, contiguous 1 GiB", lambda: rt.cudaMemcpyAsync(hx.data_ptr(), d.data_ptr(), rows * K * 4, D2H, None)),
    ("D2H 2D, 4096-byte rows, device pitch 4192 -> host pitch 4192", lambda: rt.cudaMemcpy2DAsync(h.data_ptr(), OP * 4, d.data_ptr(), OP * 4, K * 4, rows, D2H, None)),
    ("D2H 2D, 4096-byte rows, device pitch 4192 -> host pitch 4096", lambda: rt.cudaMemcpy2DAsync(hx.data_ptr(), K * 4, d.data_ptr(), OP * 4, K * 4, rows, D2H, None)),
    ("D2H 2D, 4096-byte rows, device pitch 4096 -> host pitch 4192", lambda: rt.cudaMemcpy2DAsync(h.data_ptr(), OP * 4, d.data_ptr(), K * 4, K * 4, rows, D2H, None)),
    ("H2D 1D, contiguous 1 GiB", lambda: rt.cudaMemcpyAsync(d.data_ptr(), hx.data_ptr(), rows * K * 4, H2D, None)),
    ("H2D 2D, host pitch 4096 -> device pitch 4192", lambda: rt.cudaMemcpy2DAsync(d.data_ptr(), OP * 4, hx.data_ptr(), K * 4, K * 4, rows, H2D, None)),
):
    ms = t(fn)
    print(f"{name:66s}: {ms:7.2f} ms  {gb / ms * 1e3:6.1f} GB/s")
