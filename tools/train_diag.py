"""Per-parameter gradient error of the native training path AND of torch fp32 autograd, both against torch fp64
autograd on the CPU (GPU box): shows how much of the native error is the conditioning of the train-mode BatchNorm chain.
usage: python tools/train_diag.py [B 1 X Y Z] ; writes gpurun_out/r2_train_grad_vs_fp64.json"""
import json
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from test_gpu_parity import _train_pair
from params.losses.dice_spvPA import Dice_spvPA

shape = tuple(int(v) for v in sys.argv[1:6]) if len(sys.argv) > 5 else (2, 1, 64, 64, 16)
ref, nat = _train_pair(True)
import copy
ref32 = copy.deepcopy(ref)
ref = ref.double()
g = torch.Generator().manual_seed(5)
x = torch.randn(shape, generator=g)
y = (torch.rand((shape[0], 1) + shape[2:], generator=g) > 0.7).float()
crit = Dice_spvPA(to_onehot_y=True, softmax=True)
out_r = ref(x.double()); loss_r = crit(out_r, y.double()); loss_r.backward()
out_3 = ref32(x); loss_3 = crit(out_3, y); loss_3.backward()
p3 = dict(ref32.named_parameters())
out_n = nat(x.cuda()); loss_n = crit(out_n, y.cuda()); loss_n.backward()
print("loss", loss_r.item(), loss_n.item(), "logits err", (out_n[0].cpu().double() - out_r[0]).abs().max().item())
pr, pn = dict(ref.named_parameters()), dict(nat.named_parameters())
rows = []
for n, p in pr.items():
    gr, gn = p.grad, pn[n].grad.cpu().double()
    rows.append(((gn - gr).abs().max().item() / (gr.abs().max().item() + 1e-30), gr.abs().max().item(), n,
                 ((gn * gr).sum() / (gr * gr).sum()).item() if gr.abs().sum() > 0 else float("nan"),
                 (p3[n].grad.double() - gr).abs().max().item() / (gr.abs().max().item() + 1e-30)))
rows = [r for r in rows if r[1] > 1e-9]
for r in sorted(rows, reverse=True)[:30]:
    print(f"rel_err {r[0]:.3e}  torch_fp32_rel_err {r[4]:.3e}  gmax {r[1]:.3e}  proj {r[3]:.5f}  {r[2][-60:]}")

import statistics
ratio = [r[0] / max(r[4], 1e-12) for r in rows]
summary = {"what": "max-abs gradient error per parameter tensor relative to the tensor's largest fp64 gradient entry; native (split-bf16 "
                   "activations, fp32 accumulation) and torch fp32 CPU autograd, both against torch fp64 CPU autograd; dropout 0",
           "shape": list(shape), "tensors": len(rows),
           "native_rel_err": {"max": max(r[0] for r in rows), "median": statistics.median(r[0] for r in rows)},
           "torch_fp32_rel_err": {"max": max(r[4] for r in rows), "median": statistics.median(r[4] for r in rows)},
           "native_over_fp32": {"max": max(ratio), "median": statistics.median(ratio)},
           "projection_min_max": [min(r[3] for r in rows), max(r[3] for r in rows)],
           "loss_fp64": loss_r.item(), "loss_native": loss_n.item(), "loss_fp32": loss_3.item(),
           "worst": [{"param": r[2], "native": r[0], "torch_fp32": r[4], "gmax": r[1], "proj": r[3]} for r in sorted(rows, reverse=True)[:12]]}
os.makedirs("gpurun_out", exist_ok=True)
tag = "x".join(str(v) for v in shape)
json.dump(summary, open(f"gpurun_out/r2_train_grad_vs_fp64_{tag}.json", "w"), indent=1)
print(json.dumps({k: summary[k] for k in ("native_rel_err", "torch_fp32_rel_err", "native_over_fp32", "projection_min_max")}))
