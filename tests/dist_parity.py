"""Launched by torchrun (one rank per GPU): trains every sharded model type on the reference's golden
batches, each rank taking a 1/P slice of every batch, and checks the re-assembled final state_dict and
the train MSE against the single-process reference run stored in tests/golden (same global batch)."""
import argparse
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--transport", default="nccl", choices=["nccl", "p2p"])
    ap.add_argument("--models", default="deepconn,deepconn++,NARRE,transnet++,MF_dot")
    args = ap.parse_args()
    import faulthandler
    faulthandler.dump_traceback_later(150, exit=True)        # a hung collective must not burn GPU minutes
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    import reviews4rec_b200 as R
    from reviews4rec_b200 import sharded as S
    from reviews4rec_b200.optim import FusedAdam
    from reviews4rec_b200.utils import init_transnet_optim
    from reviews4rec_b200.train import transnet_step
    from tests.helpers import assert_close, golden_batches, golden_state, load_golden
    from tests.test_gpu_models import build

    tr = S.Transport()
    runs = [(mt, "exact") for mt in args.models.split(",")] + [("deepconn", "f16"), ("NARRE", "f16")]
    for mt, mode in runs:
        z, dims = load_golden(mt)
        model, hp = build(mt, z, dims, mode=mode)           # full reference-shaped parameters, identical on all ranks
        wtr = None
        if args.transport == "p2p":
            wtr = S.P2PTransport.for_word_table(dims["V"], dims["E"])
        S.shard_model(model, tr, word_transport=wtr, agree_cap=True)      # 5-rating batches split 3 + 2
        crit = R.MSELoss(hp)
        is_tn = mt.startswith("transnet")
        opt = init_transnet_optim(hp, model, FusedAdam) if is_tn else FusedAdam(model.parameters(), lr=hp["lr"], weight_decay=hp["weight_decay"])
        model.train()
        se_sum = torch.zeros((), device="cuda", dtype=torch.float64)
        n_tot = 0
        for data, y in golden_batches(z, dims, "cuda"):
            n_tot += int(y.shape[0])
            d, yy = S.shard_batch(data, y, rank, world)
            # local SUMS scaled by P / B_global -> global-batch mean after the rank average (a rank's slice may be
            # empty when the batch is smaller than the world: sums stay well defined, means would be NaN)
            wgt = float(world) / float(y.shape[0])
            if is_tn:
                # restated step (train.transnet_step) with rank-averaged replicated gradients
                src, sfm, tgt = [list(x) for x in __import__("reviews4rec_b200.train", fromlist=["x"])._transnet_param_groups(model, hp)]
                out = model(d)
                lt = crit(out[1], yy, return_mean=False).sum() * wgt
                lx = ((model.source.ir - model.target.ir) ** 2).sum() * wgt       # out[2] = its local mean (TransNet.py:118-122)
                se = crit(out[0], yy, return_mean=False)
                g_t = torch.autograd.grad(lt, tgt, retain_graph=True, allow_unused=True)
                g_s = torch.autograd.grad(lx, src, retain_graph=True, allow_unused=True)
                g_f = torch.autograd.grad(se.sum() * wgt, sfm, allow_unused=True)
                for params, grads, o in ((tgt, g_t, opt[2]), (src, g_s, opt[0]), (sfm, g_f, opt[1])):
                    for p, g in zip(params, grads):
                        p.grad = g
                    S.allreduce_dense_grads(model)
                    o.step()
                    for p in params:
                        p.grad = None
                se_sum += se.detach().double().sum()
            else:
                model.zero_grad()
                out = model(d)
                se = crit(out, yy, return_mean=False)
                se_sum += se.detach().double().sum()
                (se.sum() * wgt).backward()
                S.allreduce_dense_grads(model)
                opt.step()
        dist.all_reduce(se_sum)
        sd = S.gather_state_dict(model)
        ref = golden_state(z, "final")
        if rank == 0:
            mse = float(se_sum) / n_tot
            want = float(z["metric.MSE_sum"]) / n_tot if is_tn else float(z["metric.MSE"])
            assert abs(mse - want) <= 1e-4 * max(1.0, abs(want)) + (0 if is_tn else 5e-5), "%s MSE %.6f vs %.6f" % (mt, mse, want)
            assert set(sd) == set(ref), set(sd) ^ set(ref)
            for k in (ref if mode == "exact" else []):      # fast modes: MSE parity only (half-precision conv operands)
                atol = hp["lr"] * dims["NB"] if (mt == "NARRE" and k.startswith("attention_scorer_") and k.endswith(".3.bias")) else 4e-6
                assert_close(sd[k], ref[k], rtol=1e-4, atol=atol, msg="%s final.%s" % (mt, k))
            print("dist_parity[%s, %s, %s, world %d]: MSE %.6f (reference %.6f), %d tensors match" % (mt, mode, args.transport, world, mse, want, len(ref) if mode == "exact" else 0), flush=True)
        dist.barrier()
    captured_prefetch_check(rank, world, tr)
    if rank == 0:
        print("DIST_PARITY_OK", flush=True)
    torch.cuda.synchronize()
    sys.stdout.flush()
    os._exit(0)                                              # skip NCCL / symmetric-memory teardown


def captured_prefetch_check(rank, world, tr):
    """train.CapturedStep with the word lookup of step k+1 prefetched on a forked graph branch (next_data=...) must
    train exactly like the eager sharded loop on the same ranks: 8-rating global batches split evenly, two epochs."""
    import reviews4rec_b200 as R
    from reviews4rec_b200 import sharded as S
    from reviews4rec_b200.optim import FusedAdam
    from reviews4rec_b200.train import CapturedStep
    from tests.helpers import assert_close, golden_batches, load_golden
    from tests.test_gpu_models import build
    for mt, mode in (("deepconn", "exact"), ("deepconn++", "f16"), ("NARRE", "f16")):
        z, dims = load_golden(mt)
        gb = golden_batches(z, dims, "cuda")
        cat = lambda a, b: None if a is None else torch.cat([a, b[:3]])
        big = [([cat(d0, d1) for d0, d1 in zip(gb[i][0], gb[(i + 1) % len(gb)][0])], torch.cat([gb[i][1], gb[(i + 1) % len(gb)][1][:3]]))
               for i in range(len(gb))]                         # 8 ratings per global batch
        local = [S.shard_batch(d, y, rank, world) for d, y in big]
        local = [([None if t is None else t.contiguous() for t in d], y.contiguous()) for d, y in local]
        finals = []
        for captured in (False, True):
            model, hp = build(mt, z, dims, mode=mode)
            S.shard_model(model, tr)
            model.train()
            crit = R.MSELoss(hp)
            opt = FusedAdam(model.parameters(), lr=hp["lr"], weight_decay=hp["weight_decay"], capturable=captured)
            se = torch.zeros(1, device="cuda")
            if captured:
                steps = [CapturedStep(model, crit, opt, d, y, se, dist.group.WORLD, float(world), next_data=local[(i + 1) % len(local)][0])
                         for i, (d, y) in enumerate(local)]
                steps[0].prime()
                for _ in range(2):
                    for st in steps:
                        st.replay()
            else:
                for _ in range(2):
                    for d, y in local:
                        model.zero_grad()
                        out = model(d)
                        e = crit(out, y, return_mean=False)
                        se += e.detach().sum()
                        torch.mean(e).backward()
                        S.allreduce_dense_grads(model)
                        opt.step()
            torch.cuda.synchronize()
            dist.all_reduce(se)
            finals.append((S.gather_state_dict(model), float(se)))
        if rank == 0:
            (a, sa), (b, sb) = finals
            assert abs(sa - sb) <= 1e-5 * abs(sa), (mt, sa, sb)
            for k in a:
                atol = hp["lr"] * 6 if (mt == "NARRE" and k.startswith("attention_scorer_") and k.endswith(".3.bias")) else 4e-6
                assert_close(b[k], a[k], rtol=1e-4, atol=atol, msg="%s captured+prefetch vs eager: %s" % (mt, k))
            print("dist_parity[%s, %s, world %d]: captured steps with prefetched lookups == eager sharded loop (SE sum %.6f)" % (mt, mode, world, sb), flush=True)
        dist.barrier()


if __name__ == "__main__":
    main()
