"""Short driver for ncu: W warm-up training steps + K profiled steps of the bench workload (no CPU legs)."""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from visiontransformer_b200.ce.classes import LightningViTModel  # noqa: E402
from visiontransformer_b200.optim import FusedAdam  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=64)
ap.add_argument("--dropout", type=float, default=0.1)
ap.add_argument("--warmup", type=int, default=2)
ap.add_argument("--steps", type=int, default=1)
ap.add_argument("--mode", default="train", choices=["train", "infer"])
args = ap.parse_args()
DROP = args.dropout
dev = torch.device("cuda:0")
torch.manual_seed(0)
m = LightningViTModel(17, 16, 768, 12, 12, hidden_dropout_prob=DROP, attention_probs_dropout_prob=DROP).to(dev).train()
opt = FusedAdam(m, lr=1e-5)
x = torch.rand(args.batch, 3, 224, 224, device=dev)
y = torch.randint(0, 17, (args.batch, 256, 256), device=dev)
for i in range(args.warmup + args.steps):
    if i == args.warmup:
        torch.cuda.synchronize()
        torch.cuda.profiler.start()   # process-wide (backward launches come from the autograd thread): ncu --profile-from-start off
    if args.mode == "train":
        loss = m.training_step((x, y), i)
        loss.backward()
        opt.step()
        opt.zero_grad()
    else:
        with torch.no_grad():
            m.model.predict_mask(x)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("done", float(loss) if args.mode == "train" else "")
