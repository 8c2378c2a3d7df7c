"""Print the SASS of one kernel between two addresses, with the register-file read count of every instruction (operand
registers that are not served by the reuse cache / uniform registers / immediates; 64-bit operands of D* count twice)."""
import re, subprocess, sys
so, pat, lo, hi = sys.argv[1], sys.argv[2], int(sys.argv[3], 16), int(sys.argv[4], 16)
txt = subprocess.run(['cuobjdump', '-sass', so], capture_output=True, text=True).stdout
for f in re.split(r'\n\s*Function : ', txt)[1:]:
    if not re.search(pat, f.split('\n')[0]):
        continue
    total = 0
    prev_reuse = {}
    for a, t in re.findall(r'/\*([0-9a-f]{4,5})\*/\s+(.*?);', f):
        if not (lo <= int(a, 16) <= hi):
            continue
        body = re.sub(r'^@!?U?P\d+\s+', '', t)
        op = body.split()[0]
        args = body[len(op):].split(',')
        srcs = [x.strip() for x in args[1:]]
        wide = op.startswith(('DFMA', 'DMUL', 'DADD'))
        reads = 0
        for slot, s in enumerate(srcs):
            m = re.match(r'^[-|~!]*\|?(R\d+)', s)
            if not m or m.group(1) == 'RZ':
                continue
            if prev_reuse.get(slot) == m.group(1):
                continue                      # served by the operand-reuse cache
            reads += 2 if wide else 1
        if op.startswith('IMAD.WIDE') and len(srcs) >= 3 and re.match(r'^R\d+', srcs[2]) :
            reads += 1                        # 64-bit addend
        prev_reuse = {slot: re.match(r'^[-|~!]*\|?(R\d+)', s).group(1) for slot, s in enumerate(srcs) if '.reuse' in s and re.match(r'^[-|~!]*\|?(R\d+)', s)}
        total += reads
        print('%s  %-60s reads=%d' % (a, t[:60], reads))
    print('total register reads', total, '-> cycles at 2 per cycle:', total / 2)
    break
