"""Dumps the SASS of one kernel from an object file / shared library and prints the innermost loops with instruction
counts (a loop = a backward branch; its body = the instructions between the target and the branch).
usage: python tools/sass_loop.py <file.o|.so> <kernel name regex> [--full]"""
import re, subprocess, sys

def kernels(path):
    out = subprocess.run(["cuobjdump", "-sass", path], capture_output=True, text=True).stdout
    cur, name = None, None
    res = {}
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            name = m.group(1); res[name] = []; continue
        m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", line)
        if m and name:
            res[name].append((int(m.group(1), 16), m.group(2).strip()))
    return res

def main():
    path, pat = sys.argv[1], sys.argv[2]
    full = "--full" in sys.argv
    for name, ins in kernels(path).items():
        if not re.search(pat, name):
            continue
        print("==", name, len(ins), "instructions")
        addr = {a: i for i, (a, _) in enumerate(ins)}
        for i, (a, t) in enumerate(ins):
            m = re.search(r"BRA(?:\.U)?\S*\s+(?:\S+,\s*)?(0x[0-9a-f]+)", t)
            if m:
                tgt = int(m.group(1), 16)
                if tgt <= a and tgt in addr:
                    body = ins[addr[tgt]:i + 1]
                    dp = sum(1 for _, x in body if re.match(r"(@!?U?P\d+\s+)?(DADD|DMUL|DFMA|DSETP)", x))
                    print("  loop 0x%04x..0x%04x: %d instructions, %d fp64" % (tgt, a, len(body), dp))
                    if full:
                        for aa, x in body:
                            print("    /*%04x*/ %s" % (aa, x))

if __name__ == "__main__":
    main()
