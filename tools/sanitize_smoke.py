"""Small end-to-end exercise of every kernel in libgpfq_b200 (for compute-sanitizer); checks results against the
CPU oracle.  The direct-solver variant is chosen with GPFQ_RESIDENT (and GPFQ_RESIDENT_CLUSTER / GPFQ_RESIDENT_TN) in the environment."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np, torch
import golden_cases as gc
import quantized_neural_nets_b200 as qb
from quantized_neural_nets_b200 import step_algorithm as sa
from oracle import gpfq_oracle as orc

DEV = torch.device("cuda:0")
ok = True
for (N, d, m, reg, lam) in [(70, 75, 300, None, 0.0), (40, 100, 130, "L1", 0.003), (33, 40, 700, "L0", 0.003)]:
    W, X, Xq = gc._problem(seed=N, N=N, d=d, m=m, relu=True, xq_noise=0.02, zero_xq=(3,))
    Qo, erro, relo, _, _ = orc.quantize_layer(W, X, Xq, m, 1.16 / 8, 8, 1, reg, lam, 1, False)
    delta = orc.layer_step_size(W, 1.16 / 8, 8, 1, reg, lam)
    for solver in (0, 2, 1):
        Q, e2, r2 = sa.quantize_layer_impl(W.to(DEV), X.to(DEV), Xq.to(DEV), m, 1.16 / 8, 8, 1, reg, lam, 1, False, DEV,
                                           solver=solver, return_partials=True)
        agree = (orc.level_index(Q.cpu(), delta, reg, lam) == orc.level_index(Qo, delta, reg, lam)).float().mean().item()
        rel = float((e2.sum() / r2.sum()).sqrt())
        good = agree >= 0.999 and abs(rel - float(relo)) <= 1e-3 * float(relo)
        ok &= good
        print(f"N={N} d={d} m={m} reg={reg} solver={solver}: agree {agree:.5f} rel {rel:.5f} (oracle {float(relo):.5f}) {'ok' if good else 'FAIL'}")
# adder path (U export) + conv capture + tiny network
W, X, Xq = gc._problem(seed=5, N=20, d=40, m=90, xq_noise=0.02)
Q, err, rel, adder, rel_adder = qb.StepAlgorithm._quantize_layer(W.to(DEV), X.to(DEV), Xq.to(DEV), 90, 1.16 / 8, 8, 1, None, 0.1, 1, False, DEV)
Qo, erro, relo, addero, _ = orc.quantize_layer(W, X, Xq, 90, 1.16 / 8, 8, 1, None, 0.1, 1, False)
good = torch.equal(Q.cpu(), Qo) and np.allclose(adder.cpu().numpy(), addero.numpy(), atol=1e-5)
ok &= good
print("adder path", "ok" if good else "FAIL")
c = gc.network_inputs()["n_msq"]
np.random.seed(c["np_seed"])
qnn = qb.QuantizeNeuralNet(c["model"].to(DEV), "tiny", c["batch"], c["loader"](), 4, 4, [], 1.16, 1.16, 1, 1, None, 0.1, 0.25, False, DEV)
qnn.quantize_network()
torch.cuda.synchronize()
print("tiny network ok")
sys.exit(0 if ok else 1)
