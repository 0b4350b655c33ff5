"""Multi-GPU consistency check (launch with torchrun, one process per GPU):
  1. neuron-sharded solve + Q all-gather gives bit-identical weights to the unsharded run and on every rank;
  2. with the calibration forward sharded over ranks as well, weights stay identical across ranks and agree with
     the unsharded run up to rounding-tie flips.
usage: torchrun --nproc-per-node 2 tools/dist_check.py [model] [batch]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist
import torchvision
import quantized_neural_nets_b200 as qb

model_name = sys.argv[1] if len(sys.argv) > 1 else "resnet18"
batch = int(sys.argv[2]) if len(sys.argv) > 2 else 16
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
rank, world = dist.get_rank(), dist.get_world_size()
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
torch.manual_seed(0)
model = getattr(torchvision.models, model_name)(weights=None).eval().to(dev)
n_layers = len([m for m in model.modules() if isinstance(m, (torch.nn.Conv2d, torch.nn.Linear))])
g = torch.Generator().manual_seed(1)
batches = [(torch.randn(batch, 3, 224, 224, generator=g), None) for _ in range(n_layers)]


def run(**kw):
    np.random.seed(0)
    q = qb.QuantizeNeuralNet(model, model_name, batch, batches, 4, 4, [], 1.16, 1.16, 1, 1, None, 0.1, 0.25, False, dev, **kw)
    q.quantize_network()
    torch.cuda.synchronize()
    return [l.weight.data.clone() for l in q.quantized_network_layers], [float(r) for (_, _, r) in q.layer_log]


def same_on_all_ranks(ws):
    ok = True
    for w in ws:
        ref = w.clone()
        dist.broadcast(ref, src=0)
        ok &= bool(torch.equal(ref, w))
    t = torch.tensor([int(ok)], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MIN)
    return bool(t.item())


w_single, rel_single = run(process_group=False)
w_shard, rel_shard = run()
w_fwd, rel_fwd = run(shard_forward=True, gram_reduce=False)
w_red, rel_red = run(shard_forward=True, gram_reduce=True)
ok1 = all(torch.equal(a, b) for a, b in zip(w_single, w_shard)) and same_on_all_ranks(w_shard)
ok2 = same_on_all_ranks(w_fwd)
agree = [float((a == b).float().mean()) for a, b in zip(w_single, w_fwd)]
# free-running comparison: the first layer sees bit-identical inputs (pure im2col of the images); later layers
# see activations that differ by cuDNN batch-size effects and by upstream tie flips, which GPFQ amplifies, so
# only the relative errors are expected to stay close there
ok3 = agree[0] == 1.0 and max(abs(a - b) / a for a, b in zip(rel_single, rel_fwd)) < 5e-2
# 3. same, but the Gram-eligible layers all-reduce d x d Gram matrices instead of all-gathering their inputs
ok4 = same_on_all_ranks(w_red)
agree_red = [float((a == b).float().mean()) for a, b in zip(w_single, w_red)]
ok5 = agree_red[0] >= 0.999 and max(abs(a - b) / a for a, b in zip(rel_single, rel_red)) < 5e-2
if rank == 0:
    print(f"gram-reduce identical on all ranks: {ok4}; per-layer agreement with unsharded: {[round(a, 4) for a in agree_red]}")
    print("per-layer rel err (gram-reduce):", [round(a, 5) for a in rel_red])
if rank == 0:
    print(f"world={world} model={model_name} batch={batch}")
    print(f"neuron-sharded == unsharded, identical on all ranks: {ok1}")
    print(f"sharded-forward identical on all ranks: {ok2}; min per-layer weight agreement with unsharded: {min(agree):.6f}, "
          f"mean {sum(agree)/len(agree):.6f}")
    print("per-layer agreement:", [round(a, 4) for a in agree])
    print("per-layer rel err (unsharded):", [round(a, 5) for a in rel_single])
    print("per-layer rel err (sharded fwd):", [round(a, 5) for a in rel_fwd])
dist.destroy_process_group()
sys.exit(0 if (ok1 and ok2 and ok3 and ok4 and ok5) else 1)
