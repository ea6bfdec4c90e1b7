import torch, time
dev='cuda'
H=torch.randn(32,128,1024).to(torch.bfloat16).pin_memory()
d=torch.empty_like(H, device=dev)
s=torch.cuda.Stream()
for n in (1,2):
    torch.cuda.synchronize()
    e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(s):
        e0.record()
        for _ in range(20): d.copy_(H, non_blocking=True)
        e1.record()
    torch.cuda.synchronize()
    ms=e0.elapsed_time(e1)/20
    print(f"H2D {H.numel()*2/1e6:.1f} MB: {ms:.3f} ms -> {H.numel()*2/ms/1e6:.1f} GB/s")
