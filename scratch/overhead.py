import sys, time, importlib, numpy as np, torch
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
import bench
mh = importlib.import_module("markov-huffman-coding_b200")
sharding = importlib.import_module("markov-huffman-coding_b200.sharding")
dev=torch.device('cuda',0); n=1<<30
tc=bench.ipsum_transition_counts()
d_in=torch.empty(n,dtype=torch.uint8,device=dev); mh.synth_markov(tc,bench.SEED,65536,0,d_in.data_ptr(),n,0)
cap=n+n//8+4096
d_pay=torch.empty(cap,dtype=torch.uint8,device=dev); d_out=torch.empty(n,dtype=torch.uint8,device=dev)
d_counts=torch.zeros(65536,dtype=torch.int64,device=dev); d_r1=torch.zeros(4,dtype=torch.int64,device=dev); d_r2=torch.zeros(4,dtype=torch.int64,device=dev)
h_counts=torch.empty(65536,dtype=torch.int64,pin_memory=True); h_res=torch.empty(8,dtype=torch.int64,pin_memory=True)
ws=mh.Workspace(n,cap); book=dectab=None
T={}
def tick(name,t0):
    t=time.perf_counter(); T[name]=T.get(name,0)+(t-t0); return t
for it in range(8):
    if it==3: T.clear()
    torch.cuda.synchronize(); t=time.perf_counter()
    mh.gpu_histogram(d_in.data_ptr(),n,0x20,1,d_counts.data_ptr(),ws,0); h_counts.copy_(d_counts,non_blocking=True); t=tick('hist_launch',t)
    torch.cuda.current_stream().synchronize(); t=tick('hist_wait',t)
    ac=h_counts.numpy().view(np.uint64).reshape(1,65536); g=sharding.global_counts(ac); t=tick('np_sum',t)
    prov=mh.CodingProvider.from_counts_array(g,1); t=tick('from_counts',t)
    if book is None: book,dectab=mh.Codebook(prov),mh.DecodeTable(prov)
    book.update(prov,0); t=tick('book_update',t)
    mh.gpu_encode(d_in.data_ptr(),n,0x20,book,0,d_pay.data_ptr(),cap,d_r1.data_ptr(),ws,0); h_res[:4].copy_(d_r1,non_blocking=True); t=tick('enc_launch',t)
    torch.cuda.current_stream().synchronize(); t=tick('enc_wait',t)
    bits=int(h_res[0])
    dectab.update(prov,0); t=tick('dec_update',t)
    mh.gpu_decode(d_pay.data_ptr(),0,bits,0x20,dectab,d_out.data_ptr(),n,d_r2.data_ptr(),ws,0); h_res[4:].copy_(d_r2,non_blocking=True); t=tick('dec_launch',t)
    torch.cuda.current_stream().synchronize(); t=tick('dec_wait',t)
for k,v in T.items(): print(f'{k:12s} {v/5*1e3:8.3f} ms')
print('total', sum(T.values())/5*1e3)
