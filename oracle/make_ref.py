"""Recipe for oracle/_ref/: the UNMODIFIED reference files of the verify path, copied byte for byte from
/root/reference so that they can travel to the GPU box (which has no /root/reference) and be timed there as the
CPU baseline (`bench.py --impl reference`, `cpu_baseline.kind = "reference"`).

TEST / BASELINE INFRASTRUCTURE ONLY.  oracle/_ref/ is an OUTPUT directory: git-ignored (no reference source ever
enters the history), not gpurun-ignored.  Nothing under speculative-decoding_b200/ imports it.

Copied: utils/{logits_processor,caching,printing}.py and sampling/*.py (speculative_generate and the modules its
package __init__ imports).  A MANIFEST with the sha256 of every file is written next to them; ref_arm.py checks it.

  python oracle/make_ref.py [--src /root/reference]
"""
import argparse
import hashlib
import json
import os
import shutil

HERE = os.path.dirname(os.path.abspath(__file__))
DST = os.path.join(HERE, "_ref")
FILES = ["utils/logits_processor.py", "utils/caching.py", "utils/printing.py", "sampling/__init__.py",
         "sampling/base_decoding.py", "sampling/speculative_decoding.py", "sampling/codec_base_decoding.py",
         "sampling/codec_speculative_decoding.py"]


def make(src="/root/reference"):
    if not os.path.isdir(os.path.join(src, "utils")):
        return False
    manifest = {}
    for rel in FILES:
        s, d = os.path.join(src, rel), os.path.join(DST, rel)
        os.makedirs(os.path.dirname(d), exist_ok=True)
        shutil.copyfile(s, d)
        manifest[rel] = hashlib.sha256(open(d, "rb").read()).hexdigest()
    json.dump({"source": src, "files": manifest}, open(os.path.join(DST, "MANIFEST.json"), "w"), indent=1)
    return True


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--src", default="/root/reference")
    ok = make(ap.parse_args().src)
    print("oracle/_ref written" if ok else "reference tree not found: oracle/_ref left as it is")
