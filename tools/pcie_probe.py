import torch, time
torch.cuda.set_device(0)
n=131080192//4
h=torch.empty(n,dtype=torch.float32).pin_memory(); d=torch.empty(n,dtype=torch.float32,device='cuda')
ho=torch.empty(65372160//4,dtype=torch.float32).pin_memory(); do=torch.empty(65372160//4,dtype=torch.float32,device='cuda')
def bw(fn,bytes_,it=10):
    fn(); torch.cuda.synchronize(); t=time.perf_counter()
    for _ in range(it): fn()
    torch.cuda.synchronize(); return bytes_*it/(time.perf_counter()-t)/1e9
print("H2D one copy GB/s", bw(lambda: d.copy_(h,non_blocking=True), n*4))
print("D2H one copy GB/s", bw(lambda: ho.copy_(do,non_blocking=True), ho.numel()*4))
s1,s2=torch.cuda.Stream(),torch.cuda.Stream()
def both():
    with torch.cuda.stream(s1): d.copy_(h,non_blocking=True)
    with torch.cuda.stream(s2): ho.copy_(do,non_blocking=True)
print("H2D+D2H concurrent: GB/s H2D-equivalent", bw(both, n*4))
def chunks8():
    c=n//8
    for i in range(8):
        with torch.cuda.stream(s1): d[i*c:(i+1)*c].copy_(h[i*c:(i+1)*c],non_blocking=True)
        with torch.cuda.stream(s2): ho[i*(ho.numel()//8):(i+1)*(ho.numel()//8)].copy_(do[i*(ho.numel()//8):(i+1)*(ho.numel()//8)],non_blocking=True)
print("8 chunks concurrent: GB/s H2D-equivalent", bw(chunks8, n*4))
