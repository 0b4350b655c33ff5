"""Times the fp32 calibration forward of a torchvision model (NCHW vs channels_last, TF32 off/on)."""
import sys, time, torch, torchvision
name = sys.argv[1] if len(sys.argv) > 1 else "resnet50"
bs = int(sys.argv[2]) if len(sys.argv) > 2 else 256
dev = torch.device("cuda:0")
torch.backends.cudnn.benchmark = True
torch.manual_seed(0)
model = getattr(torchvision.models, name)(weights=None).eval().to(dev)
x = torch.randn(bs, 3, 224, 224, device=dev)
def t(model, x, n=5):
    with torch.no_grad():
        for _ in range(3): model(x)
        torch.cuda.synchronize(); t0 = time.perf_counter()
        for _ in range(n): model(x)
        torch.cuda.synchronize()
    return (time.perf_counter() - t0) / n * 1e3
for tf32 in (False, True):
    torch.backends.cudnn.allow_tf32 = tf32
    torch.backends.cuda.matmul.allow_tf32 = tf32
    print(f"tf32={tf32} NCHW: {t(model, x):.1f} ms")
    m2 = model.to(memory_format=torch.channels_last)
    print(f"tf32={tf32} NHWC: {t(m2, x.contiguous(memory_format=torch.channels_last)):.1f} ms")
    model = model.to(memory_format=torch.contiguous_format)
