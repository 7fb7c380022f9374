import os, sys, time
sys.path.insert(0, os.getcwd())
import numpy as np, torch
import bench
from oracle import oracle as O
print("affinity", len(os.sched_getaffinity(0)), "cpu_count", os.cpu_count())
for f in ("/sys/fs/cgroup/cpu.max", "/sys/fs/cgroup/cpu/cpu.cfs_quota_us", "/sys/fs/cgroup/cpu/cpu.cfs_period_us"):
    try: print(f, open(f).read().strip())
    except Exception as e: print(f, "n/a")
bench.L, bench.C, bench.H, bench.D = 32,4096,32,128
kv = bench.synth_kv_torch(256, "cpu", 4321)
bits = kv.view(torch.int16).numpy().view(np.uint16).reshape(32,2,256,4096)
kb,vb=O.make_bins("lmsys/longchat-7b-16k")
import ctypes
libc=ctypes.CDLL("libc.so.6"); libc.mallopt(-3, 1<<30); libc.mallopt(-1, ctypes.c_int(2**31-1))
for nt in (8,16,32,64,128):
    O.set_threads(nt)
    best=9
    for rep in range(3):
        t0=time.perf_counter(); enc=O.encode_chunk(bits,0,kb,vb,0); O.decode_chunk(enc,0,kb,vb,0); best=min(best,time.perf_counter()-t0)
    t0=time.perf_counter(); sym,mx=O.quantize(bits,0,kb,vb); t1=time.perf_counter(); c=O.cdf(sym); t2=time.perf_counter(); bs,ln=O.encode_group(c,sym,0,256,0); t3=time.perf_counter()
    out=np.zeros(sym.shape,np.uint8); t3b=time.perf_counter(); O.decode_group(c,bs,ln,out,0,256,0); t4=time.perf_counter(); O.dequantize(out,mx,0,kb,vb,0); t5=time.perf_counter()
    print(nt, "chunk s", round(best,3), dict(quant=round(t1-t0,3),cdf=round(t2-t1,3),enc=round(t3-t2,3),dec=round(t4-t3b,3),deq=round(t5-t4,3)))
