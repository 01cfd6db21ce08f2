"""Count R2UR (vector -> uniform register moves) that sit directly in front of UTCHMMA instructions, per kernel of an
object file: descriptors / TMEM addresses that ptxas failed to keep in uniform registers show up here as a chain on
the MMA issue path (measured cost at cfg3: 9.2 -> 9.5 ms).   python tools/sass_r2ur_check.py py_psnode_b200/_lib/*.o"""
import re, subprocess, sys
for obj in sys.argv[1:]:
    txt = subprocess.run(["cuobjdump", "-sass", obj], stdout=subprocess.PIPE, text=True).stdout
    for f in re.split(r'\n\s*Function : ', txt)[1:]:
        name = f.split('\n')[0]
        ops = [m.group(2) for m in (re.search(r'/\*([0-9a-f]{4,5})\*/\s+(.*?);', l) for l in f.split('\n')) if m]
        mma = [i for i, o in enumerate(ops) if 'UTCHMMA' in o]
        if not mma:
            continue
        near = sum(1 for i, o in enumerate(ops) if 'R2UR' in o and any(0 < j - i <= 14 for j in mma))
        short = subprocess.run(["c++filt", name], stdout=subprocess.PIPE, text=True).stdout.strip()
        print(f"{near:4d} R2UR before MMAs  {len(mma):3d} UTCHMMA  {len(ops):5d} instr  {short[:110]}")
