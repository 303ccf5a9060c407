#!/usr/bin/env python3
"""Summarise the SASS opcode mix of kernels in an object/.so: tools/sass_mix.py <file> [name-substring]."""
import collections, re, subprocess, sys

def main():
    path, pat = sys.argv[1], (sys.argv[2] if len(sys.argv) > 2 else "")
    out = subprocess.run(["cuobjdump", "-sass", path], capture_output=True, text=True).stdout
    name, mix = None, {}
    for line in out.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            name = m.group(1); mix[name] = collections.Counter(); continue
        m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if m and name:
            mix[name][m.group(1).split(".")[0]] += 1
    for name, c in mix.items():
        if pat not in name: continue
        tot = sum(c.values())
        fp64 = sum(v for k, v in c.items() if k in ("DFMA", "DMUL", "DADD", "DSETP", "DMNMX"))
        print(f"{name}\n  total={tot} fp64={fp64} " + " ".join(f"{k}:{v}" for k, v in c.most_common(28)))

main()
