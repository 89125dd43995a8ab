"""How fast can pageable memory be made DMA-able?  cudaHostRegister + direct cudaMemcpyAsync + cudaHostUnregister versus the
library's bounce path, for one 52 MB part (what one of 8 ranks pulls at C4) and for the whole 417 MB; registration split over
threads.  Measurement aid for the upload design."""
import ctypes, json, os, sys, threading, time
import numpy as np
import torch

rt = ctypes.CDLL("libcudart.so.12") if os.path.exists("/usr/local/cuda/lib64/libcudart.so.12") else ctypes.CDLL(torch.utils.cpp_extension.CUDA_HOME + "/lib64/libcudart.so")
rt.cudaHostRegister.argtypes = [ctypes.c_void_p, ctypes.c_size_t, ctypes.c_uint]
rt.cudaHostUnregister.argtypes = [ctypes.c_void_p]
torch.cuda.init(); torch.zeros(1, device="cuda")
out = []
for mb in (52, 417):
    n = mb * (1 << 20) // 8
    a = np.random.default_rng(0).standard_normal(n)
    d = torch.empty(n, dtype=torch.float64, device="cuda")
    for nthreads in (1, 4, 16):
        for rep in range(3):
            a2 = a.copy()            # fresh pageable pages every time
            ptr = a2.ctypes.data
            page = 4096
            base = (ptr + page - 1) // page * page
            size = (ptr + n * 8 - base) // page * page
            bounds = [base + (size // page) * t // nthreads * page for t in range(nthreads + 1)]
            rcs = [0] * nthreads
            def reg(t):
                rcs[t] = rt.cudaHostRegister(ctypes.c_void_p(bounds[t]), bounds[t + 1] - bounds[t], 0)
            t0 = time.perf_counter()
            th = [threading.Thread(target=reg, args=(t,)) for t in range(nthreads)]
            [x.start() for x in th]; [x.join() for x in th]
            t_reg = time.perf_counter() - t0
            t0 = time.perf_counter()
            d.copy_(torch.from_numpy(a2), non_blocking=True); torch.cuda.synchronize()
            t_cp = time.perf_counter() - t0
            t0 = time.perf_counter()
            for t in range(nthreads):
                rt.cudaHostUnregister(ctypes.c_void_p(bounds[t]))
            t_un = time.perf_counter() - t0
        out.append({"MB": mb, "threads": nthreads, "register_ms": t_reg * 1e3, "copy_ms": t_cp * 1e3, "unregister_ms": t_un * 1e3, "rc": rcs})
        print(json.dumps(out[-1]), flush=True)
json.dump(out, open("gpurun_out/gpu_hostreg.json", "w"), indent=1)
