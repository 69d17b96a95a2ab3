#!/usr/bin/env python3
"""Build-time guard: ptxas 12.9 has emitted LDGSTS (cp.async with an L2 cache hint) whose uniform
descriptor register was odd-numbered; the warp then traps with "illegal instruction".  Fail if any
LDGSTS/LDG/STG in the library uses an odd `desc[URn]`, or any LDGSTS takes a uniform
shared-memory offset (`[Rn+URm]`: in the hinted form that offset overwrites the low policy word).
The guard must not pass vacuously: a missing or failing cuobjdump, or SASS without the chain
kernels and their LDGSTS, fails the build too."""
import re, shutil, subprocess, sys
lib = sys.argv[1]
if not shutil.which("cuobjdump"):
    sys.exit("check_sass: cuobjdump not found on PATH")
r = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True)
if r.returncode != 0:
    sys.exit(f"check_sass: cuobjdump failed ({r.returncode}): {r.stderr[-300:]}")
sass = r.stdout
for needed in ("chain_stream_kernel", "chain_persistent_kernel", "LDGSTS"):
    if needed not in sass:
        sys.exit(f"check_sass: {needed} not found in the SASS of {lib}: the guard would be vacuous")
bad = [l.strip() for l in sass.splitlines()
       if re.search(r"desc\[UR(\d*[13579])\]", l) or re.search(r"LDGSTS\S* \[R\d+\+UR\d+", l)]
if bad:
    print(f"check_sass: {len(bad)} instruction(s) with an odd uniform descriptor register, e.g.\n  {bad[0]}")
    sys.exit(1)
print("check_sass: ok")
