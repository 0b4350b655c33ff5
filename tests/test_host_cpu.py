"""CPU-side checks: the C-ABI library loads and exports every symbol include/gpfq_b200.h
declares (no compute without a GPU), the host logic of the orchestrator mirror, and the
world_size-2 sharding exchange over gloo."""
import ctypes
import os
import re
import sys

import numpy as np
import pytest
import torch
import torch.nn as nn

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    import quantized_neural_nets_b200._lib as L
    header = open(os.path.join(ROOT, "include", "gpfq_b200.h")).read()
    declared = set(re.findall(r"\b(gpfq_[a-z0-9_]+)\s*\(", header))
    assert declared, "no declarations parsed"
    dll = ctypes.CDLL(L.LIB_PATH)
    for name in declared:
        assert hasattr(dll, name), f"{name} declared in gpfq_b200.h but not exported"
    assert declared == set(L.SIGNATURES), "ctypes binding and header disagree"
    assert L.lib.gpfq_abi_version() == 7


def test_workspace_and_argument_validation_without_gpu():
    import quantized_neural_nets_b200._lib as L
    assert L.lib.gpfq_workspace_bytes(L.SOLVER_DIRECT, 512, 256, 200960) >= 512 * 200960 * 4
    assert L.lib.gpfq_workspace_bytes(L.SOLVER_GRAM_F64, 512, 256, 200960) >= 3 * 256 * 256 * 8
    assert L.lib.gpfq_workspace_bytes(L.SOLVER_GRAM_F64, 512, 4608, 768) == 0      # d too large for the Gram form
    # invalid arguments are rejected before anything touches the device
    rc = L.lib.gpfq_solve_f32(0, None, 4, None, None, 8, 4, 4, 8, 3, 2, None, 8, 0, 0.0, 0, None, 4, None, None, None,
                              None, 8, None, 0, None)
    assert rc != 0 and b"neuron range" in L.lib.gpfq_last_error()
    rc = L.lib.gpfq_quantize_f32(None, None, 4, None, 8, 7, 0.0, 0, None)
    assert rc != 0 and b"mode" in L.lib.gpfq_last_error()


def test_no_cpu_fallback():
    import quantized_neural_nets_b200 as qb
    if torch.cuda.is_available():
        pytest.skip("checks the no-GPU behaviour")
    with pytest.raises(RuntimeError):
        qb.StepAlgorithm._msq(0.1, torch.zeros(4), 8, 0.0)
    with pytest.raises(RuntimeError):
        qb.QuantizeNeuralNet(nn.Sequential(nn.Linear(4, 4)), "x", 2, [], 4, 4, [], 1.16, 1.16, 1, 1, None, 0.1, 0.25,
                             False, torch.device("cpu"))


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "quantized_neural_nets_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in text.replace("the oracle", ""), f"{f} mentions the oracle"


def test_extract_layers_matches_oracle_walk():
    import torchvision
    from oracle import gpfq_oracle as orc
    from quantized_neural_nets_b200.utils import extract_layers
    for name, count in (("resnet18", 21), ("alexnet", 8), ("resnet50", 54), ("vgg16", 16), ("mobilenet_v2", None)):
        m = getattr(torchvision.models, name)(weights=None)
        a, b = [], []
        extract_layers(m, a)
        orc.extract_layers(m, b)
        assert len(a) == len(b) and all(x is y for x, y in zip(a, b)), name
        if count is not None:
            assert len(a) == count, (name, len(a))


def test_neuron_slices_partition():
    from quantized_neural_nets_b200.sharding import neuron_slice
    for N in (1, 7, 64, 1000, 4096):
        for world in (1, 2, 3, 4, 8):
            cover = []
            for r in range(world):
                n0, n1 = neuron_slice(N, 1, world=world, rank=r)
                assert 0 <= n0 <= n1 <= N
                cover += list(range(n0, n1))
            assert cover == list(range(N))


def test_neuron_slices_of_grouped_layers_cover_whole_groups():
    """Grouped convolutions are partitioned by whole conv groups (the batched grouped solver
    rejects ranges that cut a group)."""
    from quantized_neural_nets_b200.sharding import neuron_slice, slice_rows
    for N, groups in ((96, 3), (960, 960), (64, 4), (30, 5), (24, 2)):
        per_group = N // groups
        for world in (1, 2, 3, 4, 8):
            cover = []
            for r in range(world):
                n0, n1 = neuron_slice(N, groups, world=world, rank=r)
                assert n0 % per_group == 0 and n1 % per_group == 0 and n1 - n0 <= slice_rows(N, groups, world)
                cover += list(range(n0, n1))
            assert cover == list(range(N))


def test_pack_unpack_round_trip():
    from quantized_neural_nets_b200 import sharding as sh
    g = torch.Generator().manual_seed(3)
    N, d, world = 10, 7, 4
    Q = torch.randn(N, d, generator=g)
    e2 = torch.rand(N, generator=g, dtype=torch.float64)
    r2 = torch.rand(N, generator=g, dtype=torch.float64)
    per = sh.rows_per_rank(N, world)
    bufs = [sh.pack_slice(Q, e2, r2, *sh.neuron_slice(N, 1, world=world, rank=r), per) for r in range(world)]
    Q2, e2b, r2b = sh.unpack_all(torch.cat(bufs), N, d)
    assert torch.equal(Q, Q2) and torch.equal(e2, e2b) and torch.equal(r2, r2b)


def _gloo_worker(rank, world, port, out):
    import torch.distributed as dist
    from quantized_neural_nets_b200 import sharding as sh
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    g = torch.Generator().manual_seed(5)
    N, d = 11, 6
    Qtrue = torch.randn(N, d, generator=g)
    etrue = torch.rand(N, generator=g, dtype=torch.float64)
    rtrue = torch.rand(N, generator=g, dtype=torch.float64)
    n0, n1 = sh.neuron_slice(N, 1)
    Q = torch.zeros(N, d); e = torch.zeros(N, dtype=torch.float64); r = torch.zeros(N, dtype=torch.float64)
    Q[n0:n1], e[n0:n1], r[n0:n1] = Qtrue[n0:n1], etrue[n0:n1], rtrue[n0:n1]
    Qf, ef, rf = sh.gather_layer(Q, e, r, n0, n1)
    ok = torch.equal(Qf, Qtrue) and torch.equal(ef, etrue) and torch.equal(rf, rtrue)
    # sharded calibration forward: each rank contributes the calibration rows of its own images
    m_local, dd = 5, 3
    Xtrue = torch.randn(world * m_local, dd, generator=g)
    Xqtrue = torch.randn(world * m_local, dd, generator=g)
    Xf, Xqf = sh.gather_inputs(Xtrue[rank * m_local:(rank + 1) * m_local], Xqtrue[rank * m_local:(rank + 1) * m_local])
    ok = ok and torch.equal(Xf, Xtrue) and torch.equal(Xqf, Xqtrue) and Xf.stride(0) == 1 and Xf.stride(1) % 4 == 0
    out[rank] = bool(ok)
    dist.destroy_process_group()


def test_all_gather_world2_gloo():
    import torch.multiprocessing as mp
    port = 29500 + (os.getpid() % 2000)
    with mp.Manager() as mgr:
        out = mgr.dict()
        mp.spawn(_gloo_worker, args=(2, port, out), nprocs=2, join=True)
        assert dict(out) == {0: True, 1: True}


def test_model_utils_match_reference_fixture():
    """fusion_layers_inplace / eval_sparsity / test_accuracy against outputs of the reference's own functions
    (utils.py:96-130, :133-159, :54-73; fixture written by tests/golden/make_golden.py::model_utils)."""
    import golden_cases as gc
    import quantized_neural_nets_b200 as qb
    g = np.load(os.path.join(ROOT, "tests", "golden", "model_utils.npz"))
    net = gc.bn_cnn(0)
    probe = gc.image_batches(1, 4, 8, 62)[0][0]
    with torch.no_grad():
        np.testing.assert_array_equal(net(probe).numpy(), g["logits_before"])
    assert qb.eval_sparsity(net) == g["sparsity"]
    np.testing.assert_array_equal(qb.test_accuracy(net, gc.labelled_loader(), torch.device("cpu"), topk=(1, 3)), g["topk"])
    qb.fusion_layers_inplace(net, torch.device("cpu"))
    for name, t in net.state_dict().items():
        np.testing.assert_array_equal(t.numpy(), g["fused_" + name.replace(".", "_")], err_msg=name)
    with torch.no_grad():
        np.testing.assert_array_equal(net(probe).numpy(), g["logits_after"])
    # layer indices are unchanged by the fusion (the BN modules stay in the graph)
    layers = []
    qb.extract_layers(net, layers)
    assert [type(l) for l in layers] == [nn.Conv2d, nn.Conv2d, nn.Conv2d, nn.Linear]


def test_packed_format_oracle_round_trip():
    from oracle import gpfq_oracle as orc
    rng = np.random.default_rng(0)
    for K, reg in ((8, None), (4, "L1"), (8, "L0"), (1, None), (64, None)):
        top = K + 1 if reg == "L0" else K
        bits = orc.packed_bits(K, reg)
        assert (1 << bits) >= 2 * top + 1 > (1 << (bits - 1))
        lv = rng.integers(-top, top + 1, size=1003)
        packed = orc.pack_levels(lv, K, reg)
        assert packed.dtype == np.uint8 and packed.size == (1003 + 7) // 8 * bits
        words = packed.reshape(-1, bits).astype(np.uint64)
        word = sum(words[:, b] << np.uint64(8 * b) for b in range(bits))
        back = np.stack([(word >> np.uint64(i * bits)) & np.uint64((1 << bits) - 1) for i in range(8)], axis=1)
        np.testing.assert_array_equal(back.ravel()[:1003].astype(np.int64) - top, lv)
    import quantized_neural_nets_b200._lib as L
    assert [L.lib.gpfq_packed_bits(K, m) for K, m in ((8, 0), (4, 1), (8, 2), (1, 0), (64, 3))] == [5, 4, 5, 2, 8]


def test_patch_index_draw_consumes_numpy_stream_like_the_reference():
    """SaveInputConv2d._draw makes one vectorised draw; the reference makes one np.random.choice per image
    (quantize_neural_net.py:340-345).  Same indices, same generator state afterwards."""
    from quantized_neural_nets_b200.quantize_neural_net import SaveInputConv2d
    for B, L, p in ((256, 1024, 0.25), (7, 25, 0.25), (5, 49, 1), (3, 1, 0.25), (16, 3136, 0.1)):
        keep = int(p * L + 1 if p != 1 else p * L)
        np.random.seed(11)
        want = np.concatenate([np.random.choice(np.arange(L * i, L * (i + 1)), size=keep) for i in range(B)])
        tail_want = np.random.rand(3)
        np.random.seed(11)
        got = SaveInputConv2d(3, 1, 0, 1, 1, p)._draw(B, L)
        tail_got = np.random.rand(3)
        np.testing.assert_array_equal(got, want)
        np.testing.assert_array_equal(tail_got, tail_want)
        assert got.dtype == want.dtype


def test_forward_fusion_pass_structure():
    """The torch.fx pass behind fuse_forward=True (forward_fusion.py): every BatchNorm2d becomes one fused site with
    the right activation bounds and residual wiring, the traced module shares the network's Conv2d / Linear
    objects (so the capture hooks still fire), and on CPU tensors the sites fall through to the modules' own
    forward, i.e. the function is unchanged."""
    import torchvision
    from quantized_neural_nets_b200.forward_fusion import fuse_inference_forward, FusedBNAct
    inf = float("inf")
    expect = {"resnet18": {(0.0, inf): 17, (-inf, inf): 3}, "mobilenet_v2": {(0.0, 6.0): 35, (-inf, inf): 17}}
    for name, kinds in expect.items():
        torch.manual_seed(0)
        model = getattr(torchvision.models, name)(weights=None).eval()
        for mod in model.modules():
            if isinstance(mod, nn.BatchNorm2d):
                mod.running_mean.normal_(0, 0.2)
                mod.running_var.uniform_(0.5, 1.5)
        fused, sites = fuse_inference_forward(model)
        got = {}
        for mod in fused.modules():
            if isinstance(mod, FusedBNAct):
                got[(mod.lo, mod.hi)] = got.get((mod.lo, mod.hi), 0) + 1
        assert got == kinds and sites == sum(kinds.values()), (name, got)
        convs = [m for m in model.modules() if isinstance(m, (nn.Conv2d, nn.Linear))]
        assert all(a is b for a, b in zip(convs, [m for m in fused.modules() if isinstance(m, (nn.Conv2d, nn.Linear))]))
        seen = []
        handle = convs[5].register_forward_hook(lambda m, i, o: seen.append(1))
        x = torch.randn(2, 3, 64, 64)
        with torch.no_grad():
            assert torch.equal(model(x), fused(x))
        handle.remove()
        assert len(seen) == 2
    # residual wiring: BN -> add(identity) -> ReLU collapses into one site that takes the identity as 2nd input
    block = torchvision.models.resnet18(weights=None).eval().layer1[0]
    fused, sites = fuse_inference_forward(block)
    calls = [n for n in fused.graph.nodes if n.op == "call_module" and "_gpfq_fused_bn_" in str(n.target)]
    assert sites == 2 and [len(n.args) for n in calls] == [1, 2]
    assert not any(n.op == "call_function" for n in fused.graph.nodes)          # the add is gone


def test_pointwise_convs_as_gemm_patches_and_restores():
    import torchvision
    from quantized_neural_nets_b200.forward_fusion import pointwise_convs_as_gemm
    torch.manual_seed(0)
    model = torchvision.models.resnet50(weights=None).eval()
    x = torch.randn(1, 3, 64, 64)
    with torch.no_grad():
        want = model(x)
    with pointwise_convs_as_gemm(model) as n:
        assert n == 33                                   # 16 bottlenecks x (conv1, conv3) + layer1.0.downsample; stride-2 / 3x3 / 7x7 untouched
        assert sum('forward' in m.__dict__ for m in model.modules()) == 33
        with torch.no_grad():
            assert torch.equal(model(x), want)           # CPU tensors take Conv2d.forward
    assert not any('forward' in m.__dict__ for m in model.modules())


@pytest.mark.parametrize("name", ["r01_bench_n1_final4.json", "r02b_bench_n1_final.json"])
def test_committed_bench_line_has_the_contract_keys(name):
    """The JSON lines bench.py printed on the B200 for the final trees of both rounds (profiles/) carry every key of the
    bench contract: metric / value / e2e / gpu_launches / clocks / roofline / cpu_baseline."""
    import json
    text = open(os.path.join(ROOT, "profiles", name)).read()
    j = json.loads(next(line for line in text.splitlines() if line.startswith("{")))
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                "vs_baseline", "dtype", "data", "config", "e2e", "gpu_launches", "clocks", "roofline", "cpu_baseline"):
        assert key in j, key
    assert j["n_gpus"] == 1 and j["dtype"] == "f32" and j["higher_is_better"] is True and j["vs_baseline"] is None
    assert "workload" in j["config"] and "model" not in j["config"]
    assert abs(j["value"] - j["units_per_step"] / (j["ms_per_step"] * 1e-3)) < 1e-6 * j["value"]
    e2e = j["e2e"]
    assert e2e["unit"] == j["unit"] and e2e["h2d_bytes_per_step"] > 0 and e2e["d2h_bytes_per_step"] > 0
    assert e2e["value"] != j["value"]
    assert j["gpu_launches"] > 0
    assert not set(j["clocks"]["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}
    roof = j["roofline"]
    assert roof["bound"] in ("hbm", "tensor") and roof["unit"] in ("GB/s", "TFLOP/s")
    assert abs(roof["frac"] - roof["achieved"] / roof["peak"]) < 1e-9 and roof["traffic"] is not None
    cpu = j["cpu_baseline"]
    assert cpu["kind"] in ("reference", "port") and cpu["cores"] >= 1 and cpu["value"] > 0 and cpu["sample"]
