"""Authoring-container script: which state-dict keys the UNMODIFIED reference keeps in a checkpoint.

Runs the reference's `filter_state_dict_to_trainable` (open_flamingo/train/train_utils.py:299-334, imported as is
from /root/reference) on this repo's drop-in Flamingo (whose parameter names are the reference's, see
tests/test_host_logic_cpu.py) for three configurations and commits the surviving key lists as
tests/golden/checkpoint_keys.json.

    python tests/golden/make_golden_checkpoint_keys.py
"""
import importlib.util
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
for p in (ROOT, os.path.join(ROOT, "tests"), HERE):
    if p not in sys.path:
        sys.path.insert(0, p)

from helpers_golden import load  # noqa: E402
from test_host_logic_cpu import _product  # noqa: E402


def main():
    spec = importlib.util.spec_from_file_location("ref_train_utils", "/root/reference/open_flamingo/train/train_utils.py")
    ref = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ref)
    out = []
    for every, freeze in ((1, True), (2, True), (2, False)):
        model, _, _ = _product(load(f"flamingo_every{every}"), every, freeze_lm_embeddings=freeze)
        kept = ref.filter_state_dict_to_trainable(model, model.state_dict())
        out.append({"fixture": f"flamingo_every{every}", "every": every, "freeze_lm_embeddings": freeze, "keys": sorted(kept)})
    with open(os.path.join(HERE, "checkpoint_keys.json"), "w") as f:
        json.dump(out, f, indent=0)
    print("wrote checkpoint_keys.json:", [len(c["keys"]) for c in out])


if __name__ == "__main__":
    main()
