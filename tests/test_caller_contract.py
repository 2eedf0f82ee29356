"""Drop-in conformance against the reference's CALLERS (SURVEY.md section 4 "drop-in test"): what
MedTok/tokenizer.py:124-126,199,230-238 and train_MedTok.py:213-231,281-289 pass to and read from the quantizer.
The facts are extracted from the unmodified sources by oracle/make_contract.py into tests/golden/caller_contract.json;
when /root/reference is present (authoring container) the fixture is re-derived and must be current."""
import ast
import inspect
import json
import os

import pytest

from conftest import GOLDEN_DIR, ROOT

CONTRACT = json.load(open(os.path.join(GOLDEN_DIR, "caller_contract.json")))


def test_fixture_is_current_with_the_reference():
    if not os.path.exists("/root/reference/MedTok/tokenizer.py"):
        pytest.skip("reference sources not on this box")
    from oracle import make_contract
    assert json.loads(json.dumps(make_contract.extract(), sort_keys=True)) == CONTRACT


def test_constructor_and_forward_signatures_accept_the_callers_arguments():
    from medtok_b200.vector_quantization_soft_one_new import VectorQuantizer
    ctor = inspect.signature(VectorQuantizer.__init__).parameters
    for kw in CONTRACT["ctor_kwargs"]:                       # tokenizer.py:124-126 passes these by keyword
        assert kw in ctor, f"VectorQuantizer.__init__ lacks the keyword {kw!r}"
    required = [n for n, p in ctor.items() if n != "self" and p.default is inspect.Parameter.empty]
    assert set(required) <= set(CONTRACT["ctor_kwargs"]), "a required constructor argument the caller never passes"
    fwd = [p for n, p in inspect.signature(VectorQuantizer.forward).parameters.items() if n != "self"]
    assert len(fwd) >= CONTRACT["forward_positional_args"]   # tokenizer.py:199 passes six positionals
    required_fwd = [p for p in fwd if p.default is inspect.Parameter.empty]
    assert len(required_fwd) <= CONTRACT["forward_positional_args"]
    # the import path differs only in the package name
    assert CONTRACT["quantizer_import_module"] == ["MedTok.vector_quantization_soft_one_new"]
    import medtok_b200.vector_quantization_soft_one_new  # noqa: F401


def test_forward_returns_every_key_the_callers_read():
    """static: the dict built by forward() (plus the token entries it merges in) names every key"""
    src = open(os.path.join(ROOT, "medtok_b200", "vector_quantization_soft_one_new.py")).read()
    literals = {n.value for n in ast.walk(ast.parse(src)) if isinstance(n, ast.Constant) and isinstance(n.value, str)}
    for key in CONTRACT["train_keys"] + CONTRACT["tokenizer_eval_keys"]:
        assert key in literals, f"{key!r} is read by the reference's callers but never produced"
    for key, pos in CONTRACT["train_loss_tuple_positions"].items():
        assert max(pos) <= 1 and key in literals          # [0] = vq loss, [1] = commitment loss


def test_hard_quantizer_surface():
    """NormEMAVectorQuantizer: ctor signature of norm_ema_quantizer.py:141-143 and the attributes callers touch"""
    from medtok_b200 import norm_ema_quantizer as m
    names = list(inspect.signature(m.NormEMAVectorQuantizer.__init__).parameters)[1:]
    assert names == ["n_embed", "embedding_dim", "beta", "decay", "eps", "statistic_code_usage", "kmeans_init",
                     "codebook_init_path"]
    for attr in ("reset_cluster_size", "forward"):
        assert hasattr(m.NormEMAVectorQuantizer, attr)
    for name in ("l2norm", "ema_inplace", "norm_ema_inplace", "sample_vectors", "kmeans", "EmbeddingEMA",
                 "NormEMAVectorQuantizer", "VectorQuantizer", "CVectorQuantiser", "FeaturePool"):
        assert hasattr(m, name), name


def test_unsupported_embedding_dims_are_rejected_at_construction():
    """D must be a multiple of 16 and <= 256 (tcgen05 operand layout): the constructors say so instead of the first forward"""
    from medtok_b200 import _lib
    from medtok_b200.norm_ema_quantizer import NormEMAVectorQuantizer
    from medtok_b200.vector_quantization_soft_one_new import VectorQuantizer
    with pytest.raises(_lib.MedTokVQError, match="multiple of 16"):
        NormEMAVectorQuantizer(64, 8, 0.25)                       # tokenizer.py's unused default codebook_embed_dim=8
    with pytest.raises(_lib.MedTokVQError, match="multiple of 16"):
        VectorQuantizer(n_e=96, e_dim=24, beta=0.25, entropy_loss_ratio=0.0, l2_norm=True, show_usage=False, split=[24, 24])
    NormEMAVectorQuantizer(64, 16, 0.25)
