#!/usr/bin/env python
"""Build-time guard: ptxas 12.9 has emitted LDGSTS (cp.async with an L2 cache hint) whose uniform
descriptor register was odd-numbered; the warp then traps with "illegal instruction".  Fail if any
LDGSTS/LDG/STG in the library uses an odd `desc[URn]`, or any LDGSTS takes a uniform
shared-memory offset (`[Rn+URm]`: in the hinted form that offset overwrites the low policy word)."""
import re, subprocess, sys
lib = sys.argv[1]
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
bad = [l.strip() for l in sass.splitlines()
       if re.search(r"desc\[UR(\d*[13579])\]", l) or re.search(r"LDGSTS\S* \[R\d+\+UR\d+", l)]
if bad:
    print(f"check_sass: {len(bad)} instruction(s) with an odd uniform descriptor register, e.g.\n  {bad[0]}")
    sys.exit(1)
print("check_sass: ok")
