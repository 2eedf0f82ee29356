"""Extract, from the UNMODIFIED reference sources, what its callers require of the quantizer module
(test infrastructure; writes tests/golden/caller_contract.json):

  * MedTok/tokenizer.py   -- the keyword arguments of the `VectorQuantizer(...)` construction (:124-126), the positional
                             arity of the `self.quantize(...)` call (:199) and every `quantized_result['key']` read (:230-238)
  * train_MedTok.py       -- every `quantized_result['key']` read and the loss-tuple positions indexed (:215-231,281-289)

    python oracle/make_contract.py            # needs /root/reference
"""
import ast
import json
import os
import sys

REF = os.environ.get("MEDTOK_REFERENCE", "/root/reference")
OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "caller_contract.json")


def extract(ref=REF):
    tok = ast.parse(open(os.path.join(ref, "MedTok", "tokenizer.py")).read())
    trn = ast.parse(open(os.path.join(ref, "train_MedTok.py")).read())
    ctor_kwargs, call_arity, imports = None, None, []
    for node in ast.walk(tok):
        if isinstance(node, ast.ImportFrom) and any(a.name == "VectorQuantizer" for a in node.names):
            imports.append(node.module)
        if isinstance(node, ast.Call) and isinstance(node.func, ast.Name) and node.func.id == "VectorQuantizer":
            ctor_kwargs = [k.arg for k in node.keywords]
        if (isinstance(node, ast.Call) and isinstance(node.func, ast.Attribute) and node.func.attr == "quantize"
                and len(node.args) > 1):
            call_arity = len(node.args)

    def keys_read(tree, var):
        keys, tuple_pos = set(), {}
        for node in ast.walk(tree):
            if isinstance(node, ast.Subscript) and isinstance(node.value, ast.Name) and node.value.id == var \
                    and isinstance(node.slice, ast.Constant) and isinstance(node.slice.value, str):
                keys.add(node.slice.value)
            # quantized_result['x_loss'][i]
            if isinstance(node, ast.Subscript) and isinstance(node.value, ast.Subscript) \
                    and isinstance(node.value.value, ast.Name) and node.value.value.id == var \
                    and isinstance(node.slice, ast.Constant) and isinstance(node.slice.value, int):
                k = node.value.slice.value
                tuple_pos.setdefault(k, set()).add(node.slice.value)
        return sorted(keys), {k: sorted(v) for k, v in tuple_pos.items()}

    tk, _ = keys_read(tok, "quantized_result")
    rk, rpos = keys_read(trn, "quantized_result")
    return {"source": "mims-harvard/MedTok: MedTok/tokenizer.py, train_MedTok.py (unmodified)",
            "quantizer_import_module": imports, "ctor_kwargs": ctor_kwargs, "forward_positional_args": call_arity,
            "tokenizer_eval_keys": tk, "train_keys": rk, "train_loss_tuple_positions": rpos}


if __name__ == "__main__":
    c = extract()
    json.dump(c, open(OUT, "w"), indent=1, sort_keys=True)
    print(json.dumps(c, indent=1))
    sys.exit(0)
