#!/usr/bin/env python3
"""Static (no GPU) instruction census of one kernel of libm2venc.so: SASS instructions per '// ---- ' banner region of
csrc/m2v_kernels.cu, split into ALU-pipe opcodes and the rest.  Straight-line regions execute once per macroblock, so
the static count tracks the dynamic one that tools/ncu_summary.py reads from an ncu capture; loops and divergent
branches are where the two differ.  Used to iterate on K1's instruction diet between GPU calls.

usage: sass_static.py [kernel-substring, default k1_mb_encodeILi3ELb1E] [--lines]
"""
import collections
import os
import re
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, 'fpga-mpeg2-encoder_b200', 'csrc', 'm2v_kernels.cu')
LIB = os.path.join(ROOT, 'fpga-mpeg2-encoder_b200', 'libm2venc.so')
ALU = {'VABSDIFF4', 'SHF', 'IADD3', 'ISETP', 'LOP3', 'VIMNMX', 'VIADD', 'SEL', 'PRMT', 'LEA', 'PLOP3', 'IABS', 'VIADDMNMX', 'MOV',
       'FLO', 'POPC', 'BREV', 'SGXT', 'BMSK', 'VIMNMX3', 'P2R', 'R2P', 'IADD', 'LOP'}


def run(cmd, **kw):
    return subprocess.run(cmd, capture_output=True, text=True, **kw).stdout


def main():
    args = [a for a in sys.argv[1:] if not a.startswith('--')]
    kern = args[0] if args else 'k1_mb_encodeILi3ELb1E'
    lib = args[1] if len(args) > 1 else LIB
    tmp = tempfile.mkdtemp()
    subprocess.run(['cuobjdump', '-xelf', 'all', lib], cwd=tmp, capture_output=True)
    seq, cur, line = [], None, None
    for cub in sorted(os.listdir(tmp)):
        for l in run(['nvdisasm', '-g', '-c', os.path.join(tmp, cub)]).splitlines():
            m = re.match(r'\s*\.text\.(\S+):', l)
            if m:
                cur = m.group(1); continue
            m = re.search(r'//## File "(.*?)", line (\d+)', l)
            if m:
                line = int(m.group(2)) if m.group(1).endswith('m2v_kernels.cu') else None
                continue
            if cur and kern in cur:
                m = re.match(r'\s+/\*([0-9a-f]{4,})\*/\s+(.*?);', l)
                if m:
                    seq.append((line, m.group(2).strip()))
    src = open(SRC).read().split('\n')
    banners = [(i + 1, l.strip()[8:70]) for i, l in enumerate(src) if l.strip().startswith('// ---- ')]
    first_banner = banners[0][0] if banners else 0
    state = {'last': 'prologue'}

    def region(ln):
        if ln is None or ln < first_banner:
            return state['last']
        name = 'prologue'
        for b, t in banners:
            if ln >= b:
                name = 'L%d %s' % (b, t)
        state['last'] = name
        return name
    tot, alu = collections.Counter(), collections.Counter()
    regops = collections.defaultdict(collections.Counter)
    perline = collections.defaultdict(collections.Counter)
    for ln, txt in seq:
        op = re.sub(r'^@!?U?P\w+\s+', '', txt).split()[0].split('.')[0]
        if op == 'NOP':
            continue
        reg = region(ln)
        tot[reg] += 1; regops[reg][op] += 1; perline[ln][op] += 1
        if op in ALU:
            alu[reg] += 1
    print('static census of %s: %d instructions, %d ALU-pipe' % (kern, sum(tot.values()), sum(alu.values())))
    for reg in sorted(tot, key=lambda k: int(re.match(r'L(\d+)', k).group(1)) if k[0] == 'L' else 0):
        print('   %-72s %5d instr %5d alu' % (reg, tot[reg], alu[reg]))
        print('       ' + ', '.join('%s %d' % (k, v) for k, v in regops[reg].most_common(12)))
    if '--lines' in sys.argv:
        for ln in sorted(perline, key=lambda k: (k is None, k)):
            c = perline[ln]; t = sum(c.values())
            text = src[ln - 1].strip()[:70] if ln else '(inlined helper / other file)'
            print('     %5s %5d  %-70s %s' % (ln, t, text, ', '.join('%s %d' % (k, v) for k, v in c.most_common(6))))


if __name__ == '__main__':
    main()
