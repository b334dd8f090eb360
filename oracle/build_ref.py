"""ORACLE SCAFFOLDING — recipe for `oracle/_ref/`: the UNMODIFIED reference, importable on the GPU box.

The reference is pure Python with no setup.py / pyproject (it cannot be pip-installed) and imports the un-vendored
`dynamic_network_architectures`; `/root/reference` does not exist on the GPU box.  This script copies the few files of
the hot path byte for byte from where they lie under /root/reference into `oracle/_ref/reference/` (git-ignored, NOT
gpurun-ignored: it travels with the snapshot like a built .so, and never enters the history) together with the 3-file
DNA shim of oracle/dna_shim.  `bench.py --impl reference` and the `cpu_baseline` leg import the
network and the losses from there (`kind: "reference"`); tests never read it.

    python oracle/build_ref.py            # in the build container (needs /root/reference)
"""
import hashlib
import json
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
DEST = os.path.join(HERE, "_ref")
REFERENCE_ROOT = os.environ.get("RESENC_REFERENCE_ROOT", "/root/reference")
# the hot path + the trainer's loss table; nothing else of the reference is needed to time it
FILES = [
    "builders/__init__.py", "builders/build_network_from_config.py", "builders/encoder.py", "builders/decoder.py",
    "builders/resblocks.py", "builders/simple_conv_blocks.py", "builders/utils.py",
    "training/losses/losses.py", "inference/helpers.py",
    # the reference trainer itself (tests/test_reference_trainer.py runs BaseTrainer.train() unchanged on the drop-in)
    "train.py",
]


def available():
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "builders"))


def built():
    return os.path.exists(os.path.join(DEST, "MANIFEST.json"))


def build(force=False):
    if not available():
        return False
    if built() and not force:
        return True
    root = os.path.join(DEST, "reference")
    if os.path.isdir(DEST):
        shutil.rmtree(DEST)
    manifest = {}
    for rel in FILES:
        src = os.path.join(REFERENCE_ROOT, rel)
        if not os.path.exists(src):
            if rel.endswith("__init__.py"):
                os.makedirs(os.path.dirname(os.path.join(root, rel)), exist_ok=True)
                open(os.path.join(root, rel), "w").close()
                continue
            raise FileNotFoundError(src)
        dst = os.path.join(root, rel)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        shutil.copyfile(src, dst)
        manifest[rel] = hashlib.sha1(open(src, "rb").read()).hexdigest()
    shutil.copytree(os.path.join(HERE, "dna_shim"), os.path.join(DEST, "dna_shim"),
                    ignore=shutil.ignore_patterns("__pycache__"))
    with open(os.path.join(DEST, "MANIFEST.json"), "w") as f:
        json.dump({"source": REFERENCE_ROOT, "sha1": manifest}, f, indent=1)
    return True


if __name__ == "__main__":
    ok = build(force="--force" in sys.argv)
    print("oracle/_ref:", "built" if ok else f"skipped ({REFERENCE_ROOT} not present)")
