"""CPU checks of the benchmark's host side: the synthetic Amazon-shaped batches follow SURVEY.md 8(d), and the
contract figures bench.py divides by are the ones SURVEY.md derives."""
import numpy as np
import torch


def test_synthetic_batches_follow_the_survey_spec():
    from reviews4rec_b200.synthetic import SyntheticReader
    hp = {"model_type": "deepconn", "total_users": 1000000, "total_items": 100000, "input_length": 1000}
    V, B = 50001, 512
    r = SyntheticReader(hp, B, 2, V, seed=1234)
    (data, y), (data2, _) = r.batches
    assert [d is None for d in data] == [True, True, True, False, False, False, False]        # slots DeepCoNN does not read
    ud, idoc, uid, iid = data[3], data[4], data[5], data[6]
    assert ud.dtype == torch.int64 and tuple(ud.shape) == (B, 1000) and y.dtype == torch.float32
    assert int(ud.min()) >= 0 and int(ud.max()) < V and int(uid.max()) < hp["total_users"] and int(iid.max()) < hp["total_items"]
    docs = torch.cat([ud, idoc]).numpy()
    lens = np.where((docs != 0).any(1), 1000 - np.argmax((docs != 0)[:, ::-1], axis=1), 0)
    assert 230 <= np.median(lens) <= 380                       # log-normal, median 300
    assert 0.55 <= float((lens < 1000).mean()) <= 0.95         # most documents carry a padding tail
    assert 0.50 <= float((docs == 0).mean()) <= 0.70           # ~60 % of all positions are the padding token
    body = docs[docs != 0]
    assert float((body == 1).mean()) > 5 * float((body == 1000).mean())        # Zipf head
    assert set(np.unique(y.numpy())) <= {1.0, 2.0, 3.0, 4.0, 5.0} and float((y == 5).float().mean()) > 0.4
    assert not torch.equal(ud, data2[3])                       # batches differ
    r2 = SyntheticReader(hp, B, 1, V, seed=1234)
    assert torch.equal(r2.batches[0][0][3], ud)                # seeded


def test_full_length_option_only_removes_the_padding():
    """bench.py --full-length: the same seeded tokens, no padding tail; the default workload is untouched."""
    from reviews4rec_b200.synthetic import SyntheticReader
    hp = {"model_type": "deepconn", "total_users": 1000, "total_items": 100, "input_length": 1000}
    a = SyntheticReader(hp, 64, 1, 5001, seed=7).batches[0][0][3]
    b = SyntheticReader(dict(hp, synthetic_full_length=True), 64, 1, 5001, seed=7).batches[0][0][3]
    assert int((b == 0).sum()) == 0 and tuple(b.shape) == tuple(a.shape)
    assert bool((a == 0).any())
    live = a != 0
    assert torch.equal(a[live], b[live])                        # where the default keeps a token it is the same token


def test_contract_figures():
    import bench
    hp = bench.model_hp("deepconn")
    assert bench.algorithmic_bytes_per_rating(hp) == 2416024
    assert bench.algorithmic_bytes_per_rating(bench.model_hp("NARRE")) == 4834840
    assert bench.algorithmic_bytes_per_rating(bench.model_hp("transnet++")) == 3624160
    assert bench.adam_stream_bytes_per_step(bench.model_hp("deepconn++")) == (1000002 + 100002) * 4 * 7      # 30.8 MB
    assert bench.adam_stream_bytes_per_step(bench.model_hp("NARRE")) == (500002 + 50002) * 11 * 4 * 7       # 169 MB
    assert bench.towers_of(hp) == (2, 1000) and bench.towers_of(bench.model_hp("NARRE")) == (20, 200)
    assert bench.V_WORDS == 50001 and hp["total_users"] == 1000000 and hp["total_items"] == 100000
    assert hp["word_embed_size"] == 300 and hp["input_length"] == 1000 and hp["latent_size"] == 10


def test_reference_arm_does_not_load_the_product():
    """`bench.py --impl reference` must not import reviews4rec_b200 (whose __init__ maps libr4r_b200.so): it loads
    the synthetic generator by path and runs only the oracle."""
    import subprocess
    import sys
    code = ("import sys, bench; S = bench.load_synthetic(); import oracle.r4r_oracle; "
            "bad = [m for m in sys.modules if m.startswith('reviews4rec_b200')]; "
            "assert not bad, bad; assert hasattr(S, 'SyntheticReader'); print('clean')")
    import os
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, "-c", code], cwd=root, capture_output=True, text=True)
    assert out.returncode == 0 and "clean" in out.stdout, out.stderr
