"""Test infrastructure: copy the reference's two example packages VERBATIM into ``oracle/_ref/examples/``.

    python oracle/copy_ref_examples.py            (container; needs /root/reference)

``oracle/_ref/`` is git-ignored (reference sources never enter this repo's history) but travels to the
GPU box with the snapshot, where ``tests/test_dropin_gpu.py`` constructs the reference's OWN
``A1Conditional`` / ``AbbPushBox`` classes on this package through the ``shifu`` namespace.  Only the
files the two tasks import are copied; nothing is edited."""
import hashlib
import os
import shutil
import sys

REF = os.environ.get("SHIFU_REFERENCE", "/root/reference")
HERE = os.path.dirname(os.path.abspath(__file__))
DST = os.path.join(HERE, "_ref", "examples")
FILES = ("a1_conditional/a1_conditional.py", "a1_conditional/task_config.py",
         "abb_pushbox_vision/a_prior_stage.py", "abb_pushbox_vision/task_config.py")


def main() -> int:
    src_root = os.path.join(REF, "examples")
    if not os.path.isdir(src_root):
        print(f"{src_root} not found: nothing copied (the drop-in tests will be skipped)")
        return 0
    for rel in FILES:
        dst = os.path.join(DST, rel)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        shutil.copyfile(os.path.join(src_root, rel), dst)
    for pkg in ("", "a1_conditional", "abb_pushbox_vision"):
        init = os.path.join(DST, pkg, "__init__.py")
        if not os.path.exists(init):
            open(init, "w").close()
    with open(os.path.join(DST, "MANIFEST.txt"), "w") as f:
        for rel in FILES:
            with open(os.path.join(DST, rel), "rb") as g:
                f.write(f"{hashlib.sha256(g.read()).hexdigest()}  {rel}\n")
    print(f"copied {len(FILES)} files to {DST}")
    return 0


if __name__ == "__main__":
    sys.exit(main())
