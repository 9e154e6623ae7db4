#!/usr/bin/env python
"""Source lines behind ranges of SASS instructions of one kernel (to read profiles/source_blocks.py output).

    cuobjdump -xelf all lib/libjagged_sm100a.so        # -> jg_select.sm_100a.cubin ...
    nvdisasm -g jg_select.sm_100a.cubin > select.sass
    python profiles/sass_lines.py select.sass <mangled kernel name substring> 230-280 475-560 ...
"""
import re
import sys

lines = open(sys.argv[1]).read().splitlines()
start = next(i for i, l in enumerate(lines) if l.startswith(".text.") and sys.argv[2] in l and l.rstrip().endswith(":"))
out, cur = [], None
for l in lines[start + 1:]:
    if l.startswith("//--------------------- "):
        break
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = "%s:%s" % (m.group(1).split("/")[-1], m.group(2))
    elif re.match(r"\s+/\*[0-9a-f]{4}\*/", l):
        out.append(cur)
print(len(out), "SASS instructions")
for rng in sys.argv[3:]:
    a, b = (int(x) for x in rng.split("-"))
    seen = []
    for c in out[a:b + 1]:
        if c not in seen:
            seen.append(c)
    print(rng, " ".join(seen))
