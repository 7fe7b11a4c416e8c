"""torchrun --nproc-per-node N tests/dp_check_nccl.py : data-parallel equivalence on real GPUs (NCCL).
N ranks each train on their shard of a global batch; rank 0 also trains a single-process replica on the whole
batch; losses and post-step weights must agree (CE: averaged grads; PAEDTrainer: global-batch loss, summed grads)."""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import vitseg_oracle as O  # noqa: E402  (synthetic inputs + seeded weights only)
from visiontransformer_b200.ce.classes import LightningViTModel  # noqa: E402
from visiontransformer_b200.dp import DataParallel, shard_batch  # noqa: E402
from visiontransformer_b200.paed.classes import PAEDTrainer  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
ok = True


def build(cls, C, seed):
    cfg = O.OracleConfig(num_classes=C, patch_size=16, hidden_size=128, num_hidden_layers=2, num_attention_heads=2)
    sd = O.seeded_state_dict(cfg, seed, head_gain=4.0)
    m = cls(C, 16, 128, 2, 2, hidden_dropout_prob=0.0, attention_probs_dropout_prob=0.0)
    m.load_state_dict(O.to_module_state_dict(sd, "model."), strict=True)
    return m.to(dev).train()


def compare(tag, cls, C, batch, make_opt, steps=3):
    global ok
    m = build(cls, C, 5)
    dp = DataParallel(m, make_opt(m))
    losses = []
    for i in range(steps):
        losses.append(dp.step(shard_batch(batch, rank, world), i).item())
    if cls is LightningViTModel:   # mean of shard means == global mean
        t = torch.tensor(losses, device=dev)
        dist.all_reduce(t)
        losses = (t / world).tolist()
    if rank == 0:
        ref = build(cls, C, 5)
        opt = make_opt(ref)
        rl = []
        for i in range(steps):
            loss = ref.training_step(batch, i)
            loss.backward()
            opt.step()
            opt.zero_grad(set_to_none=True)
            rl.append(loss.item())
        werr = max(((a - b).abs().max() / b.abs().max().clamp_min(1e-12)).item()
                   for (_, a), (_, b) in zip(m.named_parameters(), ref.named_parameters()))
        lerr = max(abs(a - b) / abs(b) for a, b in zip(losses, rl))
        good = lerr < 2e-3 and werr < 2e-2
        ok = ok and good
        print(f"{'PASS' if good else 'FAIL'} {tag}: dp{world} losses {losses} vs single {rl}; max rel loss err {lerr:.2e}, "
              f"max rel weight err {werr:.2e}", flush=True)


B = 4 * world
x = O.synthetic_images(B, 224, seed=1).to(dev)
y = O.synthetic_labels(B, 17, seed=2).to(dev)
# plain SGD: Adam's g/sqrt(v) turns bf16-level gradient noise on near-zero gradients into O(lr) weight differences,
# which says nothing about the all-reduce; SGD keeps the comparison linear in the gradients.
compare("CE", LightningViTModel, 17, (x, y), lambda m: torch.optim.SGD(m.parameters(), lr=0.05))
masks, se, si = [t.to(dev) for t in O.synthetic_binary_targets(B, 224, seed=3)]
compare("PAEDTrainer", PAEDTrainer, 1, (x, masks, se, si), lambda m: torch.optim.SGD(m.parameters(), lr=0.05))
# sharded inference: gathered masks == single-process masks
m = build(LightningViTModel, 17, 5).eval()
dp = DataParallel(m)
got = dp.predict_masks(x, gather=True)
if rank == 0:
    want = m.model.predict_mask(x)
    same = bool((got == want).all())
    ok = ok and same
    print(f"{'PASS' if same else 'FAIL'} sharded inference masks equal", flush=True)
dist.barrier()
dist.destroy_process_group()
if rank == 0 and not ok:
    sys.exit(1)
