"""CPU-only checks of the host layer: the C-ABI library loads and exports every symbol the header
declares, the drop-in modules construct with the reference's state_dict layout, and the ops refuse
CPU tensors (no fallback)."""
import ctypes
import os
import pickle
import re
import tempfile

import pytest
import torch

from tests.helpers import MODEL_TYPES, ROOT, golden_hp, golden_state, load_golden


def header_symbols():
    src = open(os.path.join(ROOT, "include", "r4r_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(r4r_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    from reviews4rec_b200 import _lib
    syms = header_symbols()
    assert len(syms) >= 19
    raw = ctypes.CDLL(_lib.LIB_PATH)
    for s in syms:
        assert hasattr(raw, s), "libr4r_b200.so does not export %s" % s
    assert set(syms) == set(_lib.SIGNATURES), set(syms) ^ set(_lib.SIGNATURES)
    assert _lib.lib.r4r_abi_version() == _lib.ABI_VERSION == 5


def test_argument_errors_are_reported_without_a_gpu():
    from reviews4rec_b200 import _lib
    rc = _lib.lib.r4r_word_gather_f32(None, 10, 4, None, 3, None, None)
    assert rc == -1 and b"null" in _lib.lib.r4r_last_error()
    assert _lib.lib.r4r_conv_wpack_bytes(300, 100) == 3 * 38 * (64 + 48) * 16
    assert _lib.lib.r4r_conv_wpack_bytes(300, 1000) == -1
    with pytest.raises(RuntimeError):
        _lib.call("r4r_shadow_build", None, 1, 1, None, 8, 0, None)


@pytest.mark.parametrize("mt", MODEL_TYPES)
def test_modules_construct_with_reference_state_dict(mt):
    import reviews4rec_b200 as R
    z, dims = load_golden(mt)
    tmp = tempfile.mkdtemp()
    with open(os.path.join(tmp, "word2vec.pkl"), "wb") as f:
        pickle.dump(torch.randn(dims["V"], dims["E"]).tolist(), f, 2)
    hp = golden_hp(mt, dims)
    hp["data_dir"] = tmp
    cls = {"deepconn": R.DeepCoNN, "deepconn++": R.DeepCoNN, "NARRE": R.NARRE, "transnet": R.TransNet,
           "transnet++": R.TransNet}.get(mt, R.MF)
    m = cls(hp)
    ref = golden_state(z, "init")
    sd = m.state_dict()
    assert set(sd) == set(ref)
    for k in ref:
        assert tuple(sd[k].shape) == tuple(ref[k].shape), k
    m.load_state_dict(ref)
    frozen = [n for n, p in m.named_parameters() if not p.requires_grad]
    assert frozen == [n for n in sd if n.endswith("word2vec.weight")]       # finding 2: only the word table
    assert "TextCNN" in str(m) or mt in ("bias_only", "MF_dot", "MF")


def test_ops_refuse_cpu_tensors():
    from reviews4rec_b200 import ops
    import reviews4rec_b200 as R
    with pytest.raises(RuntimeError, match="CUDA tensors only"):
        ops.rows_gather(torch.randn(5, 3), torch.zeros(2, dtype=torch.int64))
    with pytest.raises(RuntimeError, match="CUDA tensors only"):
        R.MSELoss({})(torch.randn(3), torch.randn(3))
    with pytest.raises(RuntimeError, match="CUDA tensors only"):
        ops.linear(torch.randn(3, 4), torch.randn(2, 4), None)
    z, dims = load_golden("MF_dot")
    tmp = tempfile.mkdtemp()
    hp = golden_hp("MF_dot", dims)
    m = R.MF(hp)
    with pytest.raises(RuntimeError, match="CUDA tensors only"):
        m([None] * 5 + [torch.zeros(2, dtype=torch.int64), torch.zeros(2, dtype=torch.int64)])


def test_product_does_not_import_the_oracle():
    pkg = os.path.join(ROOT, "reviews4rec_b200")
    for dp, _, fns in os.walk(pkg):
        for fn in fns:
            if fn.endswith(".py"):
                src = open(os.path.join(dp, fn)).read()
                assert "oracle" not in src, "%s mentions the oracle" % fn
