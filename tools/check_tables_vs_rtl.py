#!/usr/bin/env python3
"""Cross-check tools/gen_tables.py (ISO 13818-2 values, standard form) against the constant
tables of the reference RTL (RTL/mpeg2encoder.v:102-112,130-138,155-163,184-739).

Runs only where /root/reference exists (the build container); exit 0 and print 'skipped' elsewhere.
The RTL stores every VLC code as its low 5/6/9/10 bits plus a length; a standard code must
therefore equal  rtl_bits  zero-extended to  rtl_len  bits.
"""
import os, re, sys
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import gen_tables as G

RTL = '/root/reference/RTL/mpeg2encoder.v'


def parse_rtl():
    txt = open(RTL).read()
    tabs = {}
    pat = re.compile(r"assign\s+(\w+)((?:\[\d+\])+)\s*=\s*(-?)\s*(?:\d+'[sS]?([hdHD]))?\s*([0-9a-fA-F]+)\s*;")
    for m in pat.finditer(txt):
        name, idx, neg, base, val = m.groups()
        idx = tuple(int(x) for x in re.findall(r'\[(\d+)\]', idx))
        v = int(val, 16 if (base or 'd').lower() == 'h' else 10)
        if neg:
            v = -v
        tabs.setdefault(name, {})[idx] = v
    return tabs


def main():
    if not os.path.exists(RTL):
        print('skipped: %s not present' % RTL)
        return 0
    t = parse_rtl()
    bad = 0

    def chk(cond, msg):
        nonlocal bad
        if not cond:
            bad += 1
            print('MISMATCH', msg)

    for i in range(8):
        for j in range(8):
            chk(t['DCTM'][(i, j)] == G.DCTM[i][j], 'DCTM %d %d' % (i, j))
            chk(t['INTRA_Q'][(i, j)] == G.INTRA_Q[i][j], 'INTRA_Q %d %d' % (i, j))
            chk(t['ZIGZAG'][(i, j)] == G.ZIGZAG[i][j], 'ZIGZAG %d %d' % (i, j))

    def chk_vlc(bits_name, lens_name, idx, code, what):
        b, l = t[bits_name][idx], t[lens_name][idx]
        chk(l == len(code) and b == int(code, 2), '%s %s: rtl(bits=%x,len=%d) vs %s' % (what, idx, b, l, code))

    for i, c in enumerate(G.MOTION):
        chk_vlc('BITS_MOTION_VECTOR', 'LENS_MOTION_VECTOR', (i,), c, 'motion')
    chk(t['LENS_NZ_FLAGS'][(0,)] == 0, 'cbp0 len')
    for i in range(1, 64):
        chk_vlc('BITS_NZ_FLAGS', 'LENS_NZ_FLAGS', (i,), G.CBP[i], 'cbp')
    for i in range(12):
        chk_vlc('BITS_DC_Y', 'LENS_DC_Y', (i,), G.DC_Y[i], 'dcY')
        chk_vlc('BITS_DC_UV', 'LENS_DC_UV', (i,), G.DC_C[i], 'dcC')
    ac = G.build_ac()
    # RTL put_AC (RTL:2535-2544) selection ranges
    n03 = {0: 40, 1: 18, 2: 5, 3: 4}
    covered = set()
    for run in range(4):
        for m in range(n03[run]):
            code, ln = ac[run][m]
            chk(ln > 0, 'missing B14 (%d,%d)' % (run, m + 1))
            chk(t['LENS_AC_0_3'][(run, m)] == ln and t['BITS_AC_0_3'][(run, m)] == code,
                'AC03 (%d,%d)' % (run, m + 1))
            covered.add((run, m + 1))
    for run in range(4, 32):
        nm = 3 if run <= 6 else 2 if run <= 16 else 1
        for m in range(nm):
            code, ln = ac[run][m]
            chk(ln > 0, 'missing B14 (%d,%d)' % (run, m + 1))
            chk(t['LENS_AC_4_31'][(run, m)] == ln and t['BITS_AC_4_31'][(run, m)] == code,
                'AC431 (%d,%d)' % (run, m + 1))
            covered.add((run, m + 1))
    chk(covered == set(G.B14.keys()), 'B14 key set differs from RTL put_AC ranges: %s' %
        (covered ^ set(G.B14.keys())))
    print('tables vs RTL: %d mismatches' % bad)
    return 1 if bad else 0


if __name__ == '__main__':
    sys.exit(main())
