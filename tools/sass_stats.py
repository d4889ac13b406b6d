#!/usr/bin/env python3
"""Per-kernel SASS opcode histogram and loop-body sizes (backward-branch spans) of a .so/.cubin.

usage: tools/sass_stats.py <file> [kernel-name-substring]
"""
import re, subprocess, sys, collections

def main():
    path = sys.argv[1]
    want = sys.argv[2] if len(sys.argv) > 2 else ""
    txt = subprocess.run(["cuobjdump", "-sass", path], capture_output=True, text=True).stdout
    funcs = re.split(r"\n\s*Function : ", txt)[1:]
    for f in funcs:
        name = f.split("\n", 1)[0].strip()
        if want not in name:
            continue
        ins = re.findall(r"/\*([0-9a-f]{4,5})\*/\s+(.*?);", f)
        ops = collections.Counter()
        for addr, text in ins:
            t = text.strip()
            if t.startswith("@"):
                t = t.split(None, 1)[1]
            ops[t.split()[0].split(".")[0]] += 1
        print(f"== {name}: {len(ins)} instructions")
        print("   " + ", ".join(f"{k}:{v}" for k, v in ops.most_common(16)))
        for addr, text in ins:
            m = re.search(r"BRA\s+(?:\S+\s+)?0x([0-9a-f]+)", text)
            if m and int(m.group(1), 16) < int(addr, 16):
                lo, hi = int(m.group(1), 16), int(addr, 16)
                body = [t for a, t in ins if lo <= int(a, 16) <= hi]
                c = collections.Counter()
                for t in body:
                    t = t.strip()
                    if t.startswith("@"):
                        t = t.split(None, 1)[1]
                    c[t.split()[0].split(".")[0]] += 1
                print(f"   loop 0x{lo:x}..0x{hi:x}: {len(body)} instr: " + ", ".join(f"{k}:{v}" for k, v in c.most_common(12)))

if __name__ == "__main__":
    main()
