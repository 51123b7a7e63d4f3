"""D2H bandwidth of cudaMemcpy2DAsync for the row shapes the result download uses (n rows, width w of pitch p)."""
import ctypes as C
import time

import torch

rt = C.CDLL("libcudart.so.12") if True else None
torch.cuda.init()
torch.zeros(1, device="cuda")
n = 262144
for pitch in (1008, 1024):
    dev = torch.empty(n * pitch + 4096, dtype=torch.uint8, device="cuda")
    host = torch.empty(n * pitch + 4096, dtype=torch.uint8, pin_memory=True)
    for w in (96, 512, 544, 576, 640, 768, pitch):
        for off_kind in ("right", "left"):
            off = pitch - w if off_kind == "right" else 0
            def run():
                rc = rt.cudaMemcpy2DAsync(C.c_void_p(host.data_ptr() + off), C.c_size_t(pitch), C.c_void_p(dev.data_ptr() + off),
                                          C.c_size_t(pitch), C.c_size_t(w), C.c_size_t(n), C.c_int(2), C.c_void_p(0))
                assert rc == 0, rc
            run(); torch.cuda.synchronize()
            t0 = time.perf_counter()
            for _ in range(5):
                run()
            torch.cuda.synchronize()
            dt = (time.perf_counter() - t0) / 5
            print(f"pitch {pitch} width {w} {off_kind}-aligned: {n * w / dt * 1e-9:.1f} GB/s payload, {dt * 1e3:.2f} ms")
