"""Stage the UNMODIFIED reference quantizer files under oracle/_ref/ so that they travel to the GPU box
(test / bench infrastructure only; oracle/_ref/ is git-ignored -- no reference source enters the history).

    python oracle/build_ref.py        # needs /root/reference (authoring container)

Copies, byte for byte, the two pure-Python files of the hot path
    MedTok/norm_ema_quantizer.py, MedTok/vector_quantization_soft_one_new.py
and records their sha256 in oracle/_ref/MANIFEST.json.  `oracle/ref_loader.py` then imports them by path behind the
stub package of SURVEY.md Appendix C (the package's own __init__ pulls dgl).  bench.py --impl reference times these
unmodified modules on the box's host cores (kind "reference"); without oracle/_ref it falls back to the ATen port.
"""
import hashlib
import json
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("MEDTOK_REFERENCE", "/root/reference")
OUT = os.path.join(HERE, "_ref")
FILES = ["MedTok/norm_ema_quantizer.py", "MedTok/vector_quantization_soft_one_new.py"]


def build() -> str | None:
    if not all(os.path.isfile(os.path.join(REF, f)) for f in FILES):
        return None                                   # GPU box / no reference checkout: keep what travelled with the tree
    manifest = {}
    for f in FILES:
        dst = os.path.join(OUT, f)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        shutil.copyfile(os.path.join(REF, f), dst)
        manifest[f] = hashlib.sha256(open(dst, "rb").read()).hexdigest()
    json.dump({"source": "mims-harvard/MedTok (unmodified copies)", "sha256": manifest}, open(os.path.join(OUT, "MANIFEST.json"), "w"),
              indent=1)
    return OUT


if __name__ == "__main__":
    print(build())
    sys.exit(0)
