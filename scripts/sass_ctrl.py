"""Decode the scheduling control fields (stall count, yield, scoreboard set / wait masks) of sm_100a SASS.

usage: python scripts/sass_ctrl.py <cubin-or-so> <mangled kernel name> [regex marking the region start] [n instructions]

cuobjdump prints every 128-bit instruction as two 64-bit words; bits 105..125 hold, low to high:
stall(4) yield(1) write-barrier(3) read-barrier(3) wait-mask(6) reuse(4). A wait mask on a branch at the top of a
loop is how a "deferred" global load turns into a per-iteration stall (stall_long_sb on a BRA in ncu's source page).
"""
import re
import subprocess
import sys


def decode(path, fun):
    txt = subprocess.run(["cuobjdump", "-sass", "-fun", fun, path], capture_output=True, text=True).stdout.split("\n")
    ins, i = [], 0
    while i < len(txt):
        m = re.match(r"\s+/\*([0-9a-f]+)\*/\s+(.*?);\s+/\* (0x[0-9a-f]+) \*/", txt[i])
        m2 = re.match(r"\s+/\* (0x[0-9a-f]+) \*/", txt[i + 1]) if m and i + 1 < len(txt) else None
        if m and m2:
            w = (int(m2.group(1), 16) << 64) | int(m.group(3), 16)
            c = (w >> 105) & ((1 << 21) - 1)
            ins.append(dict(addr=int(m.group(1), 16), text=m.group(2).strip(), stall=c & 15, yld=(c >> 4) & 1,
                            wbar=(c >> 5) & 7, rbar=(c >> 8) & 7, wait=(c >> 11) & 63))
            i += 2
        else:
            i += 1
    return ins


def show(ins, lo, hi):
    for x in ins[lo:hi]:
        bar = lambda v: "-" if v == 7 else str(v)
        print("%05x %-64s st=%2d y=%d wb=%s rb=%s wait=%s" % (x["addr"], x["text"][:64], x["stall"], x["yld"], bar(x["wbar"]),
                                                             bar(x["rbar"]), format(x["wait"], "06b")))


if __name__ == "__main__":
    ins = decode(sys.argv[1], sys.argv[2])
    pat = re.compile(sys.argv[3]) if len(sys.argv) > 3 else None
    n = int(sys.argv[4]) if len(sys.argv) > 4 else 80
    print(len(ins), "instructions")
    if pat is None:
        show(ins, 0, len(ins))
    else:
        for k, x in enumerate(ins):
            if pat.search(x["text"]):
                print("---- match at %05x" % x["addr"])
                show(ins, max(0, k - 6), k + n)
