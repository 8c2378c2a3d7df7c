"""Critical path of a loop body in the SASS of one kernel: longest register-dependency (RAW) chain through ONE trip of
the innermost loop that holds at least `min_fp64` FP64 instructions, including the loop-carried dependence (the chain is
followed through two consecutive trips and the per-trip increment reported).  Latencies (cycles, measured on B200 where
noted): DFMA/DMUL/DADD 8 (scripts/micro/dfma_operands.cu), MUFU 22, LDS 30, SHFL 24, LDG 300, everything else 5.
    python scripts/sass_chain.py lib.so <kernel regex> [min_fp64]"""
import re, subprocess, sys, collections
so, pat = sys.argv[1], sys.argv[2]
min_fp64 = int(sys.argv[3]) if len(sys.argv) > 3 else 40
LAT = collections.OrderedDict([('DFMA', 8), ('DMUL', 8), ('DADD', 8), ('DSETP', 8), ('MUFU', 22), ('LDS', 30), ('SHFL', 24), ('LDG', 300), ('LDC', 30)])
def lat(op):
    for k, v in LAT.items():
        if op.startswith(k): return v
    return 5
txt = subprocess.run(['cuobjdump', '-sass', so], capture_output=True, text=True).stdout
for f in re.split(r'\n\s*Function : ', txt)[1:]:
    name = f.split('\n')[0]
    if not re.search(pat, name): continue
    ins = re.findall(r'/\*([0-9a-f]{4,5})\*/\s+(.*?);', f)
    best = None
    for a, t in ins:
        if 'BRA' in t:
            m = re.search(r'0x([0-9a-f]+)', t)
            if m and int(m.group(1), 16) <= int(a, 16):
                lo, hi = int(m.group(1), 16), int(a, 16)
                body = [(x, y) for x, y in ins if lo <= int(x, 16) <= hi]
                nd = sum(1 for _, y in body if re.sub(r'^@!?U?P\d+\s+', '', y).startswith(('DFMA', 'DMUL', 'DADD')))
                if nd >= min_fp64 and (best is None or len(body) < len(best)): best = body
    if best is None:
        print(name, 'no loop'); continue
    def regs_of(tok, wide):
        out = []
        for m in re.finditer(r'\b(U?R|U?P)(\d+)\b', tok):
            k, n = m.group(1), int(m.group(2))
            out.append((k, n))
            if wide and k == 'R': out.append((k, n + 1))
        return out
    def run(trips):
        ready = collections.defaultdict(float); who = {}
        end = 0.0; endi = None; t_issue = 0.0
        for trip in range(trips):
            for a, t in best:
                b = re.sub(r'^@!?U?P\d+\s+', '', t); pred = re.match(r'^@!?(U?P\d+)', t)
                op = b.split()[0]; args = [x.strip() for x in b[len(op):].split(',')]
                wide = op.startswith(('DFMA', 'DMUL', 'DADD', 'DSETP')) or '.64' in op or 'WIDE' in op
                is_store = op.startswith(('ST', 'BRA', 'BAR', 'EXIT', 'BSYNC', 'BSSY', 'ISETP', 'DSETP', 'FSETP', 'PLOP3', 'WARPSYNC', 'NANOSLEEP', 'RED', 'ATOM'))
                ndst = 0 if op.startswith(('ST', 'BRA', 'BAR', 'EXIT', 'BSYNC', 'BSSY', 'WARPSYNC', 'NANOSLEEP', 'RED')) else (2 if re.match(r'^(ISETP|DSETP|FSETP|PLOP3)', op) else 1)
                dsts = []; srcs = []
                for i, x in enumerate(args):
                    (dsts if i < ndst else srcs).extend(regs_of(x, wide and not x.startswith(('P', 'UP', '!P'))))
                if pred: srcs.extend(regs_of(pred.group(1), False))
                start = 0.0; src_of = None
                for r in srcs:
                    if r[1] == 255 or (r[0] in ('P', 'UP') and r[1] == 7): continue   # RZ / PT
                    if ready[r] > start: start, src_of = ready[r], who.get(r)
                fin = start + lat(op)
                for r in dsts:
                    if r[1] == 255 or (r[0] in ('P', 'UP') and r[1] == 7): continue
                    ready[r] = fin; who[r] = (trip, a, op, src_of, start)
                if fin > end: end, endi = fin, (trip, a, op, src_of, start)
        return end, endi
    def run_inorder(trips):
        """one warp alone, in-order issue: an instruction issues when the previous one has issued (FP64: 2 cycles of the
        issue port, 3 with three distinct register operands; others 1) and its operands are ready"""
        ready = collections.defaultdict(float); t = 0.0; marks = []
        for trip in range(trips):
            for a, tx in best:
                b = re.sub(r'^@!?U?P\d+\s+', '', tx); pred = re.match(r'^@!?(U?P\d+)', tx)
                op = b.split()[0]; args = [x.strip() for x in b[len(op):].split(',')]
                fp64 = op.startswith(('DFMA', 'DMUL', 'DADD', 'DSETP'))
                wide = fp64 or '.64' in op or 'WIDE' in op
                ndst = 0 if op.startswith(('ST', 'BRA', 'BAR', 'EXIT', 'BSYNC', 'BSSY', 'WARPSYNC', 'NANOSLEEP', 'RED')) else (2 if re.match(r'^(ISETP|DSETP|FSETP|PLOP3)', op) else 1)
                dsts = []; srcs = []
                for i, x in enumerate(args):
                    (dsts if i < ndst else srcs).extend(regs_of(x, wide and not x.startswith(('P', 'UP', '!P'))))
                if pred: srcs.extend(regs_of(pred.group(1), False))
                start = t
                for r in srcs:
                    if r[1] == 255 or (r[0] in ('P', 'UP') and r[1] == 7): continue
                    start = max(start, ready[r])
                nreg = len(set(r for x in args[ndst:] for r in regs_of(x, False) if r[0] == 'R' and r[1] != 255 and '.reuse' not in x))
                cost = max(2, min(3, nreg)) if fp64 else 1
                t = start + cost
                for r in dsts:
                    if r[1] == 255 or (r[0] in ('P', 'UP') and r[1] == 7): continue
                    ready[r] = start + lat(op)
            marks.append(t)
        return marks
    mk = run_inorder(4)
    print('one warp alone, in-order issue model: %.0f cycles per trip' % (mk[3] - mk[2]))
    e1, _ = run(2); e2, last = run(3)
    print('%s: %d instructions per trip; critical path per trip (steady state) %.0f cycles' % (name[:70], len(best), e2 - e1))
    # walk the chain back
    chain = []; cur = last
    while cur is not None and len(chain) < 400:
        chain.append(cur); cur = cur[3]
    c = collections.Counter(x[2].split('.')[0] for x in chain if x[0] == 2)
    print('ops on the chain in the last trip:', dict(c.most_common()))
    if '-v' in sys.argv:
        for x in reversed([y for y in chain if y[0] == 2]): print('  %s %-12s starts at %.0f' % (x[1], x[2], x[4]))
