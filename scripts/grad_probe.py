"""Gradient / Adam-trajectory probe at the benchmarked shape (scripts/, not part of the product)."""
import os, pickle, sys, tempfile
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from oracle import r4r_oracle as O
import reviews4rec_b200 as R
from reviews4rec_b200 import ops
from reviews4rec_b200.optim import FusedAdam
from reviews4rec_b200.synthetic import SyntheticReader
V, B = 50001, 64
hp = {"model_type": "deepconn", "latent_size": 10, "word_embed_size": 300, "dropout": 0.0, "total_users": 1_000_000,
      "total_items": 100_000, "lr": 0.002, "weight_decay": 1e-6, "input_length": 1000, "batch_size": B}
P = O.init_params(hp, V, seed=5)
tmp = tempfile.mkdtemp()
pickle.dump(np.zeros((V, 300), dtype=np.float32), open(os.path.join(tmp, "word2vec.pkl"), "wb"), 2)
hp["data_dir"] = tmp
batches = SyntheticReader(hp, B, 12, V, seed=1234).batches
mode = sys.argv[1] if len(sys.argv) > 1 else "exact"
ops.set_conv_mode(mode)
model = R.DeepCoNN(hp); model.load_state_dict(P); model = model.cuda().train()
opt = FusedAdam(model.parameters(), lr=hp["lr"], weight_decay=hp["weight_decay"])
oopt = None
names = dict(model.named_parameters())
for step, (data, y) in enumerate(batches):
    out_ref, se, grads = O.grads_of(P, data, y, hp, True)
    model.zero_grad()
    out = model([None if d is None else d.cuda() for d in data])
    R.MSELoss(hp)(out, y.cuda()).backward()
    line = ["step %d rating err %.2e" % (step, float((out.detach().cpu() - out_ref).abs().max()))]
    for k in ("user_conv.convs.0.weight", "user_conv.convs.0.bias", "user_conv.fc.weight", "fm.V", "global_bias"):
        g, r = names[k].grad.cpu(), grads[k]
        line.append("%s: |g|max %.2e err %.2e nz %d/%d" % (k.replace("user_conv.", "u."), float(r.abs().max()), float((g - r).abs().max()),
                                                            int((g != 0).sum()), int((r != 0).sum())))
    print("  ".join(line))
    if oopt is None:
        oopt = O.AdamState(P, [k for k in P if k not in O.frozen_keys(P)], hp["lr"], hp["weight_decay"])
    oopt.step(grads)
    opt.step()
    w, wr = names["user_conv.convs.0.weight"].detach().cpu(), P["user_conv.convs.0.weight"]
    print("     after Adam: conv W max |diff| %.3e (|W|max %.3e)  #elements off by > 1e-4: %d" % (
        float((w - wr).abs().max()), float(wr.abs().max()), int(((w - wr).abs() > 1e-4).sum())))
