"""Offline issue-cost model of the K1 inner loop (one step pair), read from the SASS.

Measured on B200 (profiles/r02_dfma_operands.txt, r02_probe_k1_ablation.log): a DFMA / DMUL / DADD occupies the issue
port of its SM sub-partition for max(2, number of 64-bit REGISTER operands not served by the operand-reuse cache)
cycles — 3 for a DFMA with three distinct register operands, 2 when one of them is an immediate, a uniform register or
a `.reuse`d operand of the previous instruction — and every other instruction for 1 cycle.  The sum over the loop body
reproduces the measured cycles per step pair of the shipped kernel and its ablations within 1-3 %.

    python scripts/k1_cost_model.py [nvcc flags ...]        compiles heun_single_balanced.cu with the flags given
    python scripts/k1_cost_model.py --so lib.so             reads an existing library
"""
import re, subprocess, sys, os, tempfile, collections

def functions(path):
    txt = subprocess.run(['cuobjdump', '-sass', path], capture_output=True, text=True).stdout
    for f in re.split(r'\n\s*Function : ', txt)[1:]:
        yield f.split('\n')[0], re.findall(r'/\*([0-9a-f]{4,5})\*/\s+(.*?);', f)

def inner_loop(ins):
    best = None
    for a, t in ins:
        if 'BRA' in t:
            m = re.search(r'0x([0-9a-f]+)', t)
            if m and int(m.group(1), 16) <= int(a, 16):
                lo, hi = int(m.group(1), 16), int(a, 16)
                body = [(x, y) for x, y in ins if lo <= int(x, 16) <= hi]
                nd = sum(1 for _, y in body if re.sub(r'^@!?U?P\d+\s+', '', y).startswith(('DFMA', 'DMUL', 'DADD')))
                if nd >= 40 and (best is None or len(body) < len(best)):
                    best = body
    return best

def cost(body, verbose=False):
    tot = 0; hist = collections.Counter(); other = 0; prev = {}
    for a, t in body:
        b = re.sub(r'^@!?U?P\d+\s+', '', t); op = b.split()[0]
        srcs = [x.strip() for x in b[len(op):].split(',')][1:]
        regs = {}
        for slot, s in enumerate(srcs):
            m = re.match(r'^[-|~!]*\|?(R\d+)', s)
            if m and m.group(1) != 'RZ':
                regs[slot] = m.group(1)
        if op.startswith(('DFMA', 'DMUL', 'DADD')):
            n = sum(1 for slot, r in regs.items() if prev.get(slot) != r)
            hist[n] += 1; c = max(2, n)
        else:
            other += 1; c = 1
        tot += c
        if verbose:
            print('%s %d  %s' % (a, c, t[:70]))
        prev = {slot: r for slot, r in regs.items() if '.reuse' in srcs[slot]}
    return tot, dict(hist), other

if __name__ == '__main__':
    args = sys.argv[1:]
    verbose = '-v' in args
    args = [a for a in args if a != '-v']
    pat = 'heun_single_balanced_kernelILb1ELb1ELb0'
    if args and args[0] == '--so':
        path = args[1]
        if len(args) > 2: pat = args[2]
    else:
        root = os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', 'magpy_b200', 'csrc')
        path = os.path.join(tempfile.mkdtemp(), 'k1.cubin')
        subprocess.check_call(['nvcc', '-O3', '-std=c++17', '-gencode', 'arch=compute_100a,code=sm_100a', '-cubin', '-o', path,
                               os.path.join(root, 'heun_single_balanced.cu')] + args)
    for name, ins in functions(path):
        if re.search(pat, name):
            body = inner_loop(ins)
            tot, hist, other = cost(body, verbose)
            regs = subprocess.run(['cuobjdump', '-res-usage', path], capture_output=True, text=True).stdout
            m = re.search(re.escape(name) + r':\s*\n\s*REG:(\d+)', regs)
            print('%s: %d instructions per step pair, model %d cycles (%.1f per step); fp64 by live 64-bit register operands %s, other %d; registers %s'
                  % (name[:60], len(body), tot, tot / 2, hist, other, m.group(1) if m else '?'))
