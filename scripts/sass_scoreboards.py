#!/usr/bin/env python
"""Lists the SASS of one kernel with the scoreboard fields of every instruction decoded from the 128-bit encoding
(write barrier, read barrier, wait mask, stall count): which loads share a scoreboard with which consumer.
usage: python scripts/sass_scoreboards.py build/sgd.o <mangled-kernel-name-substring> [first last]"""
import re
import subprocess
import sys


def main():
    obj, pat = sys.argv[1], sys.argv[2]
    lo, hi = (int(sys.argv[3]), int(sys.argv[4])) if len(sys.argv) > 4 else (0, 1 << 30)
    out = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True).stdout
    on, k, pend = False, 0, None
    for line in out.splitlines():
        if "Function :" in line:
            on = pat in line
            k = 0
            if on:
                print(line.strip())
            continue
        if not on:
            continue
        m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?)\s*/\* (0x[0-9a-f]{16}) \*/", line)
        if m:
            pend = (m.group(1), m.group(2))
            continue
        m = re.match(r"\s*/\* (0x[0-9a-f]{16}) \*/", line)
        if m and pend:
            hi64 = int(m.group(1), 16)
            ctrl = hi64 >> 41
            stall, yld, wb, rb, wait = ctrl & 15, (ctrl >> 4) & 1, (ctrl >> 5) & 7, (ctrl >> 8) & 7, (ctrl >> 11) & 63
            if lo <= k < hi:
                print("%4d %s W%s R%s wait=%s st=%2d  %s" % (k, pend[0], "-" if wb == 7 else wb, "-" if rb == 7 else rb,
                                                            "".join(str(b) if wait >> b & 1 else "." for b in range(6)), stall, pend[1]))
            k += 1
            pend = None


if __name__ == "__main__":
    main()
