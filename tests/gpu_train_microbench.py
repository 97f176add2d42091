"""Microbenchmarks behind the training-step choices (run on a B200: python tests/gpu_train_microbench.py): device-side
pixel draws (permutation vs first-occurrence-distinct draws vs with-replacement) and fused Adam over 48 tensors vs one
flat tensor."""
import torch, time
dev="cuda"
gen=torch.Generator(device=dev); gen.manual_seed(0)
def t(fn, n=200):
    for _ in range(20): fn()
    torch.cuda.synchronize()
    e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
    t0=time.perf_counter(); e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1)/n*1e3, (time.perf_counter()-t0)/n*1e6
N=1024; M=640000
def a(): return torch.randperm(M, device=dev, generator=gen)[:N]
def b():
    c = torch.randint(0, M, (2*N,), device=dev, generator=gen)
    vals, order = torch.sort(c, stable=True)
    dup_sorted = torch.zeros(2*N, dtype=torch.bool, device=dev); dup_sorted[1:] = vals[1:] == vals[:-1]
    dup = torch.empty_like(dup_sorted); dup[order] = dup_sorted
    keep = ~dup
    pos = torch.cumsum(keep, 0) - 1
    slot = torch.where(keep & (pos < N), pos, N)
    out = torch.empty(N+1, dtype=torch.int64, device=dev); out[slot] = c
    return out[:N]
def c_(): return torch.randint(0, M, (N,), device=dev, generator=gen)
print("randperm us (gpu, wall):", t(a)); print("distinct-draw us:", t(b)); print("randint us:", t(c_))
x=b(); print(torch.unique(x).numel())
# adam cost
ps=[torch.nn.Parameter(torch.randn(256,256,device=dev)) for _ in range(48)]
for p in ps: p.grad=torch.randn_like(p)
o=torch.optim.Adam(ps, lr=1e-3, fused=True)
print("fused adam 48 tensors us:", t(o.step))
fp=torch.nn.Parameter(torch.randn(1191688,device=dev)); fp.grad=torch.randn_like(fp)
o2=torch.optim.Adam([fp], lr=1e-3, fused=True)
print("fused adam flat us:", t(o2.step))
