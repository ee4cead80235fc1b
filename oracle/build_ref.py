#!/usr/bin/env python
"""oracle/build_ref.py -- vendor the reference's OWN implementation of the path into oracle/_ref (git-ignored).

TEST INFRASTRUCTURE ONLY (same rule as the rest of oracle/): only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs may import what lands in oracle/_ref.

The reference is pure Python: "building" it is copying the handful of modules that make up the path, unmodified,
from the read-only checkout (REF_PATH, default /root/reference) so that they travel to the GPU box with the snapshot
(/root/reference does not exist there).  Nothing is edited; two import-time dependencies that are absent from the
image and irrelevant to the path are satisfied by stubs written next to the copies:

    termcolor   (utils.py:5, coloured log prefixes)          -> stubs/termcolor.py: colored(s, *a, **k) = s
    models.models_utils.truncate_number (int_opt_layer.py)   -> only if the OPT block is imported; not copied

Copied (relative to the reference root):
    quant/{__init__,quantizer,int_linear,int_matmul,omni_norm,utils,omniquant}.py
    models/{int_llama_layer,int_opt_layer,int_falcon_layer,transformation,models_utils}.py      utils.py
Run by __graft_entry__.build() when the checkout is present; a no-op (keeps what is there) otherwise.
"""
from __future__ import annotations

import hashlib
import json
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
DEST = os.path.join(HERE, "_ref")
FILES = ["quant/__init__.py", "quant/quantizer.py", "quant/int_linear.py", "quant/int_matmul.py",
         "quant/omni_norm.py", "quant/utils.py", "quant/omniquant.py",
         "models/int_llama_layer.py", "models/int_opt_layer.py", "models/int_falcon_layer.py",
         "models/transformation.py", "models/models_utils.py", "utils.py"]
STUB_TERMCOLOR = '"""Stub written by oracle/build_ref.py: the reference\'s utils.py imports termcolor for log colours."""\n' \
                 "def colored(text, *args, **kwargs):\n    return text\n"


def build(ref_path: str | None = None, verbose: bool = False) -> bool:
    """Returns True when oracle/_ref is usable (freshly copied or already there)."""
    ref = ref_path or os.environ.get("REF_PATH", "/root/reference")
    if not os.path.isdir(ref):
        ok = os.path.exists(os.path.join(DEST, "quant", "quantizer.py"))
        if verbose:
            print(f"[oracle/_ref] {ref} not present; {'keeping the existing copy' if ok else 'no copy available'}")
        return ok
    manifest = {}
    for rel in FILES:
        src = os.path.join(ref, rel)
        dst = os.path.join(DEST, rel)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        shutil.copyfile(src, dst)
        manifest[rel] = hashlib.sha256(open(src, "rb").read()).hexdigest()
    os.makedirs(os.path.join(DEST, "stubs"), exist_ok=True)
    with open(os.path.join(DEST, "stubs", "termcolor.py"), "w") as f:
        f.write(STUB_TERMCOLOR)
    with open(os.path.join(DEST, "MANIFEST.json"), "w") as f:
        json.dump({"source": ref, "sha256": manifest}, f, indent=1)
    if verbose:
        print(f"[oracle/_ref] copied {len(FILES)} unmodified reference modules from {ref}")
    return True


def paths() -> list:
    """sys.path entries that make `import quant.quantizer` etc. resolve to the vendored reference."""
    return [DEST, os.path.join(DEST, "stubs")]


def available() -> bool:
    return os.path.exists(os.path.join(DEST, "quant", "quantizer.py"))


if __name__ == "__main__":
    sys.exit(0 if build(verbose=True) else 1)
