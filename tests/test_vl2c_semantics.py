"""Known-answer tests of oracle/vl2c.py, the Verilog-2001 -> C++ translator that executes the reference RTL (the oracle's
pin, tests/test_rtl_pin.py, is only as good as this translator).  Two small modules written for this purpose
(tests/vl2c_kat/*.v - not reference code) exercise the language rules the RTL leans on; every expected value below was
derived BY HAND from IEEE 1364-2001 (4.4 expression bit lengths, 4.5 signed expressions, 9.2 blocking / non-blocking
assignment), not by running any simulator:

  comb.v  context-determined widths (carry kept or lost), zero- vs sign-extension by the propagated type, signed x
          unsigned products, >>> on signed / unsigned operands, signed / unsigned comparisons, unsized literals widening
          the context to 32 bits, truncating signed division and modulo, concatenation, replication, reductions, +: selects
  wide.v  vectors wider than 64 bits as the RTL's bit packer uses them (RTL:2879-2994): wide concatenations, OR, a variable
          left shift, an indexed part select, a concatenation on the left-hand side, wide equality, byte reversal (relational operators on wide vectors are
          outside the translator's subset - it refuses them - and the RTL does not use them).
          Expected values: Python integers applying the definitions (concatenation = shift and or, truncation = mod 2^n).
  seq.v   non-blocking swap, blocking temporaries inside a clocked block, counter wrap, concatenation on the left-hand
          side, an array with negative bounds read before written, a blocking for-loop accumulation, case, functions with
          sized arguments / local regs / signed clipping, asynchronous reset
"""
import os
import subprocess
import sys

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
VL2C = os.path.join(ROOT, 'oracle', 'vl2c.py')


def _run(tmp_path, name, params, harness):
    hpp = tmp_path / (name + '.hpp')
    subprocess.check_call([sys.executable, VL2C, os.path.join(HERE, 'vl2c_kat', name + '.v'), str(hpp)] + params,
                          stdout=subprocess.DEVNULL)
    cpp = tmp_path / (name + '.cpp')
    cpp.write_text('#include "%s.hpp"\n#include <stdio.h>\n%s' % (name, harness))
    exe = tmp_path / name
    subprocess.check_call(['g++', '-O1', '-std=c++17', '-o', str(exe), str(cpp)], cwd=str(tmp_path))
    return subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout


def test_expression_sizing_and_signedness(tmp_path):
    # a = b = 200 (8'hC8), sa = -3 (8'shFD), sel = 1
    want = dict(
        sum9=400,            # 9-bit context keeps the carry
        avg8=72,             # (a + b) evaluated in 8 bits: 400 mod 256 = 144, >> 1
        avg9=200,            # {1'b0, a} widens the context to 9 bits: 401 >> 1
        mixmul=50600,        # one unsigned operand makes the product unsigned: sa is ZERO-extended to 16 bits, 253 * 200
        smul=168,            # both signed: (-3) * $signed(8'd200) = (-3) * (-56)
        sshr=254,            # -3 >>> 1 = -2
        ushr=100,            # >>> on an unsigned operand is a logical shift
        mix_shr=242,         # $signed(200) = -56, >>> 2 = -14
        wide_sshr=65522,     # the same in a 16-bit signed context: sign-extended first, 16'hFFF2
        cmp_s=1,             # -3 < 0, signed comparison (unsized decimal literals are signed)
        cmp_mixed=0,         # sa < b with b unsigned: 253 < 200 is false
        cmp_ss=1,            # -3 > -56
        tern=400,            # both arms of ?: take the 9-bit context
        cat=51400, rep=170, rand_=0, ror_=1, rxor_=1,
        ps=2,                # a[2 +: 4] = bits 5..2 of 8'b11001000
        lit32=100,           # the unsized 1 is 32 bits wide: 201 >> 1
        lit8=99,             # all operands 8 bits: (200 + 255) mod 256 = 199, >> 1
        lit_mixed=228,       # ... + 1 widens to 32 bits: 456 >> 1
        sdiv=255, smod=255,  # -3 / 2 = -1 and -3 % 2 = -1 (truncation toward zero, sign of the dividend)
        udiv=66, neg=56, bnot=55,
        mul8=64,             # 40000 mod 256
        shl8=128, shl16=3200,
        lnot_and=6,          # !a = 0, (a && b) = 1 -> bit 1, (a[0] || sel) = 1 -> bit 2
    )
    names = list(want)
    harness = 'int main() { Sim s; s.v_a = 200; s.v_b = 200; s.v_sa = 0xFD; s.v_sel = 1; s.init(); s.comb();\n' + \
              ''.join('  printf("%s %%llu\\n", (unsigned long long)s.v_%s);\n' % (n, n) for n in names) + '  return 0; }\n'
    got = dict((l.split()[0], int(l.split()[1])) for l in _run(tmp_path, 'comb', [], harness).splitlines())
    assert got == want


def test_clocked_semantics(tmp_path):
    # din = 10, 20, ...; idx (signed 4 bit) = -2, 2, -2, 0, 3, 3, -3, -3, -3, 1; we = 1 for the first four clocks.
    # mem has bounds [-2:2]; idx = +-3 is outside them, so `rd` is not checked on those clocks (X in a 4-state simulator).
    X = None
    want = [  # fmean r1 r2 r3 cnt hi   lo  rd  total csel  sacc
        (6,  2, 1, 4, 1, 10,  206, 0,  0,  0xA0, 254),   # swap; t = old r1 + 1 -> r3 = 4; {hi,lo} = {din,8'hCD}+1; sacc = -2
        (11, 1, 2, 6, 2, 20,  206, 0,  10, 0xA1, 0),     # total sums the OLD memory (mem[-2] = 10 written last clock)
        (16, 2, 1, 4, 3, 30,  206, 10, 30, 0xAF, 254),   # rd = old mem[-2] (read before write)
        (21, 1, 2, 6, 4, 40,  206, 0,  50, 0xAF, 254),
        (26, 2, 1, 4, 5, 50,  206, X,  90, 0xA0, 1),     # we = 0 from here on: memory frozen at 30 + 40 + 20
        (31, 1, 2, 6, 6, 60,  206, X,  90, 0xA1, 4),     # 1 + 3 = 4: not clipped (v > 4 is false)
        (36, 2, 1, 4, 7, 70,  206, X,  90, 0xAF, 1),
        (41, 1, 2, 6, 0, 80,  206, X,  90, 0xAF, 254),   # cnt wraps 7 -> 0
        (46, 2, 1, 4, 1, 90,  206, X,  90, 0xA0, 252),   # -2 - 3 = -5 -> clipped to -4
        (51, 1, 2, 6, 2, 100, 206, 0,  90, 0xA1, 253),   # rd = mem[1] = 0
    ]
    harness = r'''
int main() {
    Sim s; s.init();
    s.v_rstn = 1; s.clock(); s.v_rstn = 0; s.clock(); s.v_rstn = 1;
    int idxs[] = {-2, 2, -2, 0, 3, 3, -3, -3, -3, 1};
    for (int k = 0; k < 10; k++) {
        s.v_din = 10 * (k + 1); s.v_idx = (unsigned)idxs[k] & 15; s.v_we = (k < 4);
        s.comb();
        printf("%llu", (unsigned long long)s.v_fmean);
        s.clock();
        printf(" %llu %llu %llu %llu %llu %llu %llu %llu %llu %llu\n", (unsigned long long)s.v_r1, (unsigned long long)s.v_r2,
               (unsigned long long)s.v_r3, (unsigned long long)s.v_cnt, (unsigned long long)s.v_hi, (unsigned long long)s.v_lo,
               (unsigned long long)s.v_rd, (unsigned long long)s.v_total, (unsigned long long)s.v_csel, (unsigned long long)s.v_sacc);
    }
    return 0;
}
'''
    rows = [tuple(int(x) for x in l.split()) for l in _run(tmp_path, 'seq', ['DEPTH=2'], harness).splitlines()]
    assert len(rows) == len(want)
    for k, (g, w) in enumerate(zip(rows, want)):
        assert all(b is None or a == b for a, b in zip(g, w)), (k, g, w)


def test_wide_vectors(tmp_path):
    harness = r'''
static void pw(const char *n, const Big &v, int bits) { printf("%s ", n); for (int i = bits - 4; i >= 0; i -= 4) printf("%llx", (unsigned long long)v.slice(i, 4)); printf("\n"); }
int main() {
    Sim s; s.init();
    s.v_rstn = 1; s.clock(); s.v_rstn = 0; s.clock(); s.v_rstn = 1;
    for (int k = 1; k <= 27; k++) { s.v_din = (unsigned)(k * 7 + 3) & 255; s.clock(); }
    s.v_x64 = 0xDEADBEEF00C0FFEEull; s.v_sh = 140; s.v_sel = 19; s.comb();
    pw("acc", s.v_acc, 200); pw("w", s.v_w, 256); pw("hi128", s.v_hi128, 128); pw("lo72", s.v_lo72, 72); pw("le", s.v_le, 256);
    printf("wbyte %llx\neq %llu\n", (unsigned long long)s.v_wbyte, (unsigned long long)s.v_eq);
    s.v_sh = 300; s.comb();                                   // shifted entirely out of the 256-bit context
    pw("w300", s.v_w, 256);
    return 0;
}
'''
    out = dict(l.split() for l in _run(tmp_path, 'wide', [], harness).splitlines())
    acc = prev = 0
    for k in range(1, 28):                                        # 27 bytes through a 25-byte register
        prev = acc
        acc = ((acc << 8) | ((k * 7 + 3) & 255)) & ((1 << 200) - 1)
    x = 0xDEADBEEF00C0FFEE
    w = ((acc << 56) | (x << 140)) & ((1 << 256) - 1)
    le = int.from_bytes(w.to_bytes(32, 'big')[::-1], 'big')
    assert int(out['acc'], 16) == acc
    assert int(out['w'], 16) == w
    assert int(out['hi128'], 16) == prev >> 72 and int(out['lo72'], 16) == prev & ((1 << 72) - 1)   # the OLD acc
    assert int(out['wbyte'], 16) == (w >> (8 * 19)) & 255
    assert int(out['eq']) == int(prev == acc) == 0
    assert int(out['le'], 16) == le
    assert int(out['w300'], 16) == (acc << 56) & ((1 << 256) - 1)
