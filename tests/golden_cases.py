"""Seeded input builders shared by tests/golden/make_golden.py (which feeds them to the
reference) and by the tests (which feed them to the oracle and to the CUDA path).
Only inputs live here -- expected outputs are the committed tests/golden/*.npz files."""
import numpy as np
import torch
import torch.nn as nn


def _gen(seed):
    return torch.Generator().manual_seed(seed)


def quantizer_inputs():
    d = torch.tensor(0.25)
    ties = d * torch.tensor([-2.5, -1.5, -0.5, 0.5, 1.5, 2.5, 7.5, 8.5, 100.0, -100.0, 0.0, -0.0,
                             3.4999, 3.5001, -3.4999, -3.5001])
    lam_edge = torch.tensor([0.1, -0.1, 0.10000001, -0.10000001, 0.225, 0.22500001, -0.225, 0.35, 0.3499999,
                             float(np.float32(0.1) + np.float32(0.125))])
    rnd = torch.randn(4096, generator=_gen(11)) * 0.7
    return {
        "ties": (torch.cat([ties, lam_edge]), d, 8, 0.1),
        "rand": (rnd, torch.tensor(0.0371), 4, 0.05),
        "k1": (torch.randn(512, generator=_gen(12)), torch.tensor(0.5), 1, 0.3),
    }


def _problem(seed, N, d, m, relu=False, xq_noise=0.0, zero_xq=(), zero_x=(), wscale=0.1):
    g = _gen(seed)
    W = torch.randn(N, d, generator=g) * wscale
    X = torch.randn(m, d, generator=g)
    if relu:
        mix = torch.randn(d, d, generator=g) * 0.3 + torch.eye(d)
        X = torch.relu(X @ mix)
    Xq = X + xq_noise * torch.randn(m, d, generator=g) if xq_noise else X.clone()
    if relu:
        Xq = torch.relu(Xq)
    for t in zero_xq:
        Xq[:, t] = 0
    for t in zero_x:
        X[:, t] = 0
    return W, X.contiguous(), Xq.contiguous()


def greedy_inputs():
    cases = {}

    def add(tag, mode, K, lam, delta, **kw):
        W, X, Xq = _problem(**kw)
        cases[tag] = dict(W=W, X=X, Xq=Xq, mode=mode, K=K, lam=lam, delta=torch.tensor(delta))

    add("g_small", "msq", 8, 0.0, 0.04, seed=21, N=8, d=16, m=32)
    add("g_xq", "msq", 8, 0.0, 0.035, seed=22, N=33, d=20, m=50, xq_noise=0.05, zero_xq=(3, 19), zero_x=(7,))
    add("g_soft", "soft", 8, 0.02, 0.03, seed=23, N=16, d=24, m=37)
    add("g_hard", "hard", 8, 0.02, 0.03, seed=24, N=16, d=24, m=40, xq_noise=0.02)
    add("g_relu", "msq", 2, 0.0, 0.11, seed=25, N=40, d=70, m=130, relu=True, xq_noise=0.03)
    add("g_wide", "msq", 4, 0.0, 0.06, seed=26, N=130, d=100, m=64, relu=True, xq_noise=0.02, zero_xq=(0,))
    return cases


def layer_inputs():
    cases = {}

    def add(tag, reg, lam, pct, groups, K=8, step=1.16 / 8, **kw):
        W, X, Xq = _problem(**kw)
        if groups > 1:  # X holds groups * d columns (quantize_neural_net.py:337-338)
            g = _gen(kw["seed"] + 1000)
            X = torch.relu(torch.randn(X.shape[0], W.shape[1] * groups, generator=g))
            Xq = torch.relu(X + 0.02 * torch.randn(X.shape, generator=g))
        cases[tag] = dict(W=W, X=X, Xq=Xq, reg=reg, lam=lam, pct=pct, groups=groups, K=K, step=step)

    add("l_msq", None, 0.1, 1, 1, seed=31, N=24, d=40, m=96, relu=True, xq_noise=0.02)
    add("l_pct", None, 0.1, 0.9, 1, seed=32, N=20, d=36, m=70, xq_noise=0.02)
    add("l_l1", "L1", 0.01, 1, 1, seed=33, N=24, d=40, m=96, relu=True, xq_noise=0.02)
    add("l_l0", "L0", 0.01, 1, 1, seed=34, N=24, d=40, m=96, relu=True, xq_noise=0.02)
    add("l_3bit", None, 0.1, 1, 1, K=4, step=1.16 / 4, seed=35, N=17, d=33, m=45, xq_noise=0.02)
    add("l_grp", None, 0.1, 1, 2, seed=36, N=8, d=12, m=52)
    add("l_grp4_l1", "L1", 0.005, 1, 4, seed=37, N=16, d=9, m=44)
    return cases


def conv_inputs():
    cases = {}

    def add(tag, shape, kernel, padding, dilation, p, seed, stride=(1, 1), groups=1):
        g = _gen(seed)
        a = torch.randn(*shape, generator=g)
        q = a + 0.01 * torch.randn(*shape, generator=g)
        cases[tag] = dict(inp_a=a, inp_q=q, kernel=kernel, padding=padding, dilation=dilation, stride=stride,
                          groups=groups, p=p, np_seed=seed)

    add("c_3x3", (3, 4, 9, 9), (3, 3), (1, 1), (1, 1), 0.25, 41)
    add("c_1x1_s2", (2, 6, 8, 8), (1, 1), (0, 0), (1, 1), 0.25, 42, stride=(2, 2))
    add("c_7x7", (2, 3, 30, 30), (7, 7), (3, 3), (1, 1), 0.25, 43, stride=(2, 2))
    add("c_dil", (2, 2, 13, 11), (3, 2), (2, 1), (2, 2), 0.5, 44)
    add("c_full", (2, 3, 6, 6), (3, 3), (0, 0), (1, 1), 1, 45)
    add("c_grp", (2, 8, 8, 8), (3, 3), (1, 1), (1, 1), 0.25, 46, groups=4)
    return cases


def tiny_cnn(seed=0):
    torch.manual_seed(seed)
    return nn.Sequential(
        nn.Conv2d(3, 8, 3, padding=1), nn.ReLU(), nn.MaxPool2d(2),
        nn.Conv2d(8, 8, 3, padding=1, groups=2), nn.ReLU(),
        nn.Conv2d(8, 16, 1, stride=2), nn.ReLU(),
        nn.AdaptiveAvgPool2d(1), nn.Flatten(), nn.Linear(16, 10),
    ).eval()


def image_batches(n, batch, size, seed):
    g = _gen(seed)
    return [(torch.randn(batch, 3, size, size, generator=g), torch.zeros(batch, dtype=torch.long)) for _ in range(n)]


def network_inputs():
    cases = {}
    for tag, reg, lam, ignore in (("n_msq", None, 0.1, []), ("n_l1", "L1", 0.002, [0])):
        cases[tag] = dict(model=tiny_cnn(0), batch=6, bits=4, scalar=1.16, reg=reg, lam=lam, p=0.25, ignore=ignore,
                          np_seed=5, loader=lambda: image_batches(4, 6, 16, 51),
                          probe=image_batches(1, 5, 16, 52)[0][0])
    return cases


def bn_cnn(seed=0):
    """Conv/BN pairs with and without a conv bias, non-trivial running statistics; some exact zeros in the weights."""
    torch.manual_seed(seed)
    net = nn.Sequential(
        nn.Conv2d(3, 6, 3, padding=1, bias=False), nn.BatchNorm2d(6), nn.ReLU(),
        nn.Conv2d(6, 8, 3, padding=1, bias=True), nn.BatchNorm2d(8, eps=1e-3), nn.ReLU(),
        nn.Conv2d(8, 8, 1), nn.ReLU(),                       # no BN behind it: must stay untouched
        nn.AdaptiveAvgPool2d(1), nn.Flatten(), nn.Linear(8, 5),
    ).eval()
    g = _gen(seed + 100)
    for mod in net:
        if isinstance(mod, nn.BatchNorm2d):
            mod.running_mean = torch.randn(mod.num_features, generator=g) * 0.3
            mod.running_var = torch.rand(mod.num_features, generator=g) + 0.5
            mod.weight.data = torch.randn(mod.num_features, generator=g)
            mod.bias.data = torch.randn(mod.num_features, generator=g) * 0.2
    net[0].weight.data[net[0].weight.data.abs() < 0.05] = 0.0
    net[-1].bias.data[:2] = 0.0
    return net


def labelled_loader(n=23, batch=5, size=8, classes=5, seed=61):
    g = _gen(seed)
    x = torch.randn(n, 3, size, size, generator=g)
    y = torch.randint(0, classes, (n,), generator=g)
    return torch.utils.data.DataLoader(torch.utils.data.TensorDataset(x, y), batch_size=batch, shuffle=False)


def config1_inputs():
    """BASELINE.json configs[0]: Linear 1024->1024, m=2048 Gaussian inputs, 4-bit, scalar 1.16."""
    torch.manual_seed(0)
    W = nn.Linear(1024, 1024).weight.data.clone()
    X = torch.randn(2048, 1024, generator=_gen(1))
    return dict(W=W, X=X, K=8, step=1.16 / 8)


def checksum(*tensors):
    """CRC32 of the raw bytes of the input tensors (as a float, npz-friendly), stored beside the
    golden outputs so a test can tell 'inputs regenerated differently on this machine' from
    'wrong answer'.  Bitwise, hence independent of any floating-point summation order."""
    import zlib
    crc = 0
    for t in tensors:
        crc = zlib.crc32(t.detach().cpu().contiguous().numpy().tobytes(), crc)
    return float(crc)
