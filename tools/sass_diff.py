#!/usr/bin/env python
"""Compares the SASS of two builds kernel by kernel (cuobjdump -sass on .so / .cubin files): lists kernels whose
instruction streams differ, appeared or disappeared.  Used to show that a host-side or generator change left the
default kernels byte-identical when no GPU is at hand to re-run the parity suite.

    python tools/sass_diff.py old/libcsmc.so classicalspinmc.jl_b200/libcsmc.so
"""
import collections
import re
import subprocess
import sys


def load(path):
    if path.endswith(".sass"):
        text = open(path).read()
    else:
        text = subprocess.run(["cuobjdump", "-sass", path], capture_output=True, text=True, check=True).stdout
    kernels, name = collections.defaultdict(list), None
    for line in text.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            name = m.group(1)
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(.*?);", line)
        if m and name:
            kernels[name].append(m.group(1).strip())
    return kernels


def main():
    a, b = load(sys.argv[1]), load(sys.argv[2])
    changed = sorted(k for k in a if k in b and a[k] != b[k])
    print(f"{len(a)} kernels before, {len(b)} after, {sum(1 for k in a if k in b and a[k] == b[k])} identical")
    for k in changed:
        print(f"changed: {k} ({len(a[k])} -> {len(b[k])} instructions)")
    for k in sorted(set(b) - set(a)):
        print(f"new:     {k} ({len(b[k])} instructions)")
    for k in sorted(set(a) - set(b)):
        print(f"gone:    {k}")
    return 1 if changed else 0


if __name__ == "__main__":
    sys.exit(main())
