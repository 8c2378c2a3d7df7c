"""Print the instruction mix of the innermost loops of one kernel (cuobjdump -sass)."""
import re, subprocess, sys, collections
so, pat = sys.argv[1], sys.argv[2]
txt = subprocess.run(['cuobjdump', '-sass', so], capture_output=True, text=True).stdout
funcs = re.split(r'\n\s*Function : ', txt)
for f in funcs[1:]:
    name = f.split('\n')[0]
    if not re.search(pat, name):
        continue
    ins = re.findall(r'/\*([0-9a-f]{4,5})\*/\s+(.*?);', f)
    addr = [int(a, 16) for a, _ in ins]
    print('==', name, len(ins), 'instructions')
    loops = []
    for a, t in ins:
        m = re.search(r'BRA(?:\.U)?\S*\s+(?:!?U?P\d,\s*)?`?\(?\.?L_x_\d+\)?|BRA\S*\s.*0x([0-9a-f]+)', t)
        m2 = re.search(r'0x([0-9a-f]+)', t) if 'BRA' in t else None
        if m2:
            tgt = int(m2.group(1), 16)
            if tgt <= int(a, 16):
                loops.append((tgt, int(a, 16)))
    for lo, hi in sorted(set(loops), key=lambda x: x[1] - x[0]):
        body = [t for a, t in ins if lo <= int(a, 16) <= hi]
        c = collections.Counter(re.sub(r'^@!?U?P\d+\s+', '', t).split()[0].split('.')[0] for t in body)
        print(f'loop {lo:#x}-{hi:#x}: {len(body)} instr:', dict(c.most_common()))
    if len(sys.argv) > 3:
        for a, t in ins:
            print(a, t)
