"""Import the UNMODIFIED reference quantizer modules: from /root/reference in the authoring container, from the
byte-identical copies under oracle/_ref/ (oracle/build_ref.py) on the GPU box, where only bench.py's CPU legs
(`--impl reference`, `cpu_baseline`) use them.  Recipe from SURVEY.md Appendix C: skip MedTok/__init__.py (it pulls dgl) by
registering a stub package, and shim two unused names that transformers>=5 dropped."""
import os
import sys
import types

REF_ROOT = os.environ.get("MEDTOK_REFERENCE", "/root/reference")
if not os.path.isfile(os.path.join(REF_ROOT, "MedTok", "norm_ema_quantizer.py")):
    # GPU box: the byte-identical copies staged by oracle/build_ref.py (git-ignored, travel with the tree)
    REF_ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")


def available() -> bool:
    return os.path.isfile(os.path.join(REF_ROOT, "MedTok", "norm_ema_quantizer.py"))


def load():
    """Returns (hard_module, soft_module) = the reference's norm_ema_quantizer and
    vector_quantization_soft_one_new modules, imported by path, untouched."""
    if not available():
        raise RuntimeError(f"reference not present at {REF_ROOT}")
    if "MedTok" not in sys.modules or not hasattr(sys.modules["MedTok"], "__path__"):
        pkg = types.ModuleType("MedTok")
        pkg.__path__ = [os.path.join(REF_ROOT, "MedTok")]
        sys.modules["MedTok"] = pkg
    import transformers.modeling_utils as mu
    for name in ("get_parameter_device", "get_parameter_dtype"):   # imported but unused by the reference
        if not hasattr(mu, name):
            setattr(mu, name, lambda m: None)
    import MedTok.norm_ema_quantizer as hard
    import MedTok.vector_quantization_soft_one_new as soft
    return hard, soft
