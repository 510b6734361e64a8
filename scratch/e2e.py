import sys, time, importlib, ctypes, numpy as np, torch
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
import bench
mh = importlib.import_module("markov-huffman-coding_b200")
n=1<<30
tc=bench.ipsum_transition_counts()
d_in=torch.empty(n,dtype=torch.uint8,device='cuda'); mh.synth_markov(tc,bench.SEED,65536,0,d_in.data_ptr(),n,0)
h_in=torch.empty(n,dtype=torch.uint8,pin_memory=True); h_in.copy_(d_in); torch.cuda.synchronize()
cap=n+n//8+4096
h_s=torch.empty(cap+1,dtype=torch.uint8,pin_memory=True); h_b=torch.empty(n,dtype=torch.uint8,pin_memory=True)
s=mh.Session(n)
lib=mh._lib; ol=ctypes.c_uint64(0); dl=ctypes.c_uint64(0)
a,b,c=h_in.numpy(),h_s.numpy(),h_b.numpy()
for i in range(4):
    t0=time.perf_counter(); tab=ctypes.c_void_p()
    rc=lib.mh_session_compress(s._h,a.ctypes.data,n,1,b.ctypes.data,b.size,ctypes.byref(ol),ctypes.byref(tab)); t1=time.perf_counter()
    rc2=lib.mh_session_decompress(s._h,tab,b.ctypes.data,ol.value,c.ctypes.data,n,ctypes.byref(dl)); t2=time.perf_counter()
    lib.mh_table_destroy(tab)
    print(rc,rc2,'compress %.2f ms  decompress %.2f ms'%((t1-t0)*1e3,(t2-t1)*1e3))
