"""Generates tests/golden/velped_ref.json from the UNMODIFIED reference (run HERE, where /root/reference exists):
the output of oracle/_ref/velped_ref = tests/cpp/velped_driver.cpp linked with the reference's own
VelocityGenerator.o / PedestrianHandler.o (oracle/Makefile).  The full output (about 430 KB of hex words) is pinned by
its SHA-256 and line count; per-section digests and the first lines are kept so that a mismatch can be localised on a
box that has no reference tree."""
import hashlib
import json
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))


def digest(text):
    lines = text.strip().split("\n")
    sections = {}
    for l in lines:
        tag, kind = l.split(" ")[:2]
        sections.setdefault(f"{tag}_{kind}", hashlib.sha256()).update((l + "\n").encode())
    return {"lines": len(lines), "sha256": hashlib.sha256(text.encode()).hexdigest(),
            "sections": {k: v.hexdigest() for k, v in sorted(sections.items())},
            "head": lines[:12]}


if __name__ == "__main__":
    subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "ref"], stdout=subprocess.DEVNULL)
    out = subprocess.run([os.path.join(ROOT, "oracle", "_ref", "velped_ref")], capture_output=True, text=True, check=True).stdout
    with open(os.path.join(HERE, "velped_ref.json"), "w") as f:
        json.dump(digest(out), f, indent=1)
    print("wrote velped_ref.json:", digest(out)["lines"], "lines")
