"""Is a cuBLAS SGEMM faster than cuDNN's fp32 convolution for the 1x1 convolutions of ResNet-50 at bs=256?"""
import torch
import torch.nn.functional as F

dev = torch.device("cuda:0")
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
torch.backends.cudnn.benchmark = True


def t(fn, n=10):
    for _ in range(3):
        fn()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / n


tot_c = tot_m = 0.0
for (cin, cout, hw, count) in ((64, 64, 56, 1), (64, 256, 56, 3), (256, 64, 56, 2), (256, 128, 56, 1), (128, 512, 28, 4),
                               (512, 128, 28, 3), (512, 256, 28, 1), (256, 1024, 14, 6), (1024, 256, 14, 5),
                               (1024, 512, 14, 1), (512, 2048, 7, 3), (2048, 512, 7, 2)):
    x = torch.randn(256, cin, hw, hw, device=dev)
    w = torch.randn(cout, cin, 1, 1, device=dev) * 0.05
    w2 = w.view(cout, cin)
    with torch.no_grad():
        tc = t(lambda: F.conv2d(x, w))
        tm = t(lambda: torch.matmul(w2, x.flatten(2)))
        y1, y2 = F.conv2d(x, w), torch.matmul(w2, x.flatten(2)).view(256, cout, hw, hw)
    fl = 2.0 * 256 * hw * hw * cin * cout
    tot_c += tc * count
    tot_m += tm * count
    print(f"{cin:5d}->{cout:5d} @{hw:3d}: conv2d {tc:.3f} ms ({fl/tc/1e9:.1f} TF/s)  matmul {tm:.3f} ms ({fl/tm/1e9:.1f} TF/s)  "
          f"rel diff {float((y1-y2).norm()/y1.norm()):.1e}", flush=True)
print(f"per full forward: conv2d {tot_c:.2f} ms, matmul {tot_m:.2f} ms")
# 3x3 for reference
for (c, hw) in ((64, 56), (128, 28), (256, 14), (512, 7)):
    x = torch.randn(256, c, hw, hw, device=dev)
    w = torch.randn(c, c, 3, 3, device=dev) * 0.05
    with torch.no_grad():
        tc = t(lambda: F.conv2d(x, w, padding=1))
    print(f"3x3 {c}->{c} @{hw}: {tc:.3f} ms ({2.0*256*hw*hw*c*c*9/tc/1e9:.1f} TF/s)")
