#!/usr/bin/env python3
"""Compare the SASS of two builds of libssb.so kernel by kernel (opcode histograms and resource usage), to show that a
change meant for one code path left the measured kernels alone.

    python tools/sass_compare.py OLD.so NEW.so [kernel-name-substring ...]

Instruction operands that only move with the size of the kernel parameters (constant-bank offsets) are ignored."""
import re
import subprocess
import sys
from collections import Counter


def kernels(path):
    sass = subprocess.run(["cuobjdump", "-sass", path], capture_output=True, text=True, check=True).stdout
    out, name = {}, None
    for line in sass.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            name = m.group(1)
            out[name] = []
            continue
        if name is None:
            continue
        text = re.sub(r"/\*[0-9a-f]{4,}\*/", "", line)
        text = re.sub(r"/\* 0x[0-9a-f]+ \*/", "", text).strip()
        if text and not text.startswith("."):
            out[name].append(text)
    return out


def resources(path):
    txt = subprocess.run(["cuobjdump", "-res-usage", path], capture_output=True, text=True, check=True).stdout
    out, name = {}, None
    for line in txt.splitlines():
        m = re.match(r"\s*Function (\S+):", line)
        if m:
            name = m.group(1)
        elif name and "REG:" in line:
            out[name] = " ".join(tok for tok in line.split() if tok.split(":")[0] in ("REG", "STACK", "SHARED"))
            name = None
    return out


def opcodes(body):
    return Counter((tok.split()[1] if tok.startswith("@") else tok.split()[0]) for tok in body if tok)


def short(name):
    m = re.match(r"_ZN3ssb(\d+)", name)
    return name[len(m.group(0)):len(m.group(0)) + int(m.group(1))] + ("<1>" if "ILb1E" in name else "<0>" if "ILb0E" in name else "") if m else name


def main():
    old_path, new_path, wanted = sys.argv[1], sys.argv[2], sys.argv[3:]
    old, new, r_old, r_new = kernels(old_path), kernels(new_path), resources(old_path), resources(new_path)
    by_short_old = {short(k): k for k in old}
    for k in sorted(new, key=short):
        s = short(k)
        if wanted and not any(w in s for w in wanted):
            continue
        ko = by_short_old.get(s) or by_short_old.get(s.replace("<0>", ""))
        if ko is None:
            print(f"{s:28s} NEW        {len(new[k]):7d} instructions   {r_new.get(k, '')}")
            continue
        a, b = opcodes(old[ko]), opcodes(new[k])
        diff = {op: (a[op], b[op]) for op in set(a) | set(b) if a[op] != b[op]}
        verdict = "identical opcode histogram" if not diff else "differs: " + ", ".join(f"{op} {x}->{y}" for op, (x, y) in sorted(diff.items()))
        print(f"{s:28s} {len(old[ko]):7d} -> {len(new[k]):7d} instructions   {r_old.get(ko, '')} -> {r_new.get(k, '')}   {verdict}")


if __name__ == "__main__":
    main()
