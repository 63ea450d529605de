"""Stage the hot-path files of the reference into oracle/_ref/ (TEST INFRASTRUCTURE — git-ignored, never product).

/root/reference exists only in the build container; the GPU box gets a snapshot of this repo.  The few pure-Python
reference files the checker needs there travel as an UNMODIFIED, git-ignored copy under oracle/_ref/ (listed in
.gitignore, not in .gpurunignore) — the Python counterpart of compiling a C reference into oracle/_ref/:

  * bench.py --impl reference / cpu_baseline: the reference's own LlavaQwenForCausalLM (+ HF transformers) and
    utiles.weighted_kmeans_feature / torch_kmeans timed on the host cores (`kind: "reference"`);
  * tests/test_gpu_dropin_driver.py: the driver's own `updating_memory_buffer` and
    `longva_inference_with_embedding_multi_modal` (inference_streaming_longva_v2.py:164-378) AST-extracted and run
    over dropin/.

Run by __graft_entry__.build() whenever /root/reference is present:   python -m oracle.stage_ref
"""
from __future__ import annotations

import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
DEST = os.path.join(HERE, "_ref")
REFERENCE = os.environ.get("STREAMCHAT_REFERENCE", "/root/reference")
FILES = ["inference_streaming_longva_v2.py", "utiles.py", "inference_streamchat_v0.3.sh", "LICENSE"]
PACKAGES = ["longva", "torch_kmeans", "memory_bank"]


def staged() -> bool:
    return os.path.exists(os.path.join(DEST, "inference_streaming_longva_v2.py"))


def stage(reference: str = REFERENCE, dest: str = DEST) -> bool:
    """Copy the files (sources only, no bytecode).  Returns False when the reference tree is absent."""
    if not os.path.isdir(reference):
        return False
    os.makedirs(dest, exist_ok=True)
    for fn in FILES:
        src = os.path.join(reference, fn)
        if os.path.exists(src):
            shutil.copyfile(src, os.path.join(dest, fn))
    for pkg in PACKAGES:
        src, dst = os.path.join(reference, pkg), os.path.join(dest, pkg)
        if os.path.isdir(dst):
            shutil.rmtree(dst)
        shutil.copytree(src, dst, ignore=shutil.ignore_patterns("__pycache__", "*.pyc", "*.so", "*.bin", "*.pt"))
    with open(os.path.join(dest, "README"), "w") as f:
        f.write("Unmodified copy of hmxiong/StreamChat hot-path sources, staged by oracle/stage_ref.py from "
                f"{reference}.\nGit-ignored test infrastructure; never imported by streamchat_b200/.\n")
    return True


if __name__ == "__main__":
    ok = stage()
    print("staged reference into", DEST if ok else "(nothing: reference tree absent)")
    sys.exit(0)
