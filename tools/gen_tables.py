#!/usr/bin/env python3
"""Generate the MPEG-2 VLC / matrix tables used by the oracle and by the CUDA product.

The values are ISO/IEC 13818-2 facts (Tables B-9, B-10, B-12, B-13, B-14, the default intra
quantiser matrix and the zig-zag scan) plus the 8-bit integer transform matrix the reference uses
(RTL/mpeg2encoder.v:102-112).  They are stated here in the standard's own form (bit strings), NOT in
the reference's 6-bit-suffix form; tools/check_tables_vs_rtl.py cross-checks them against the RTL
(RTL/mpeg2encoder.v:184-739) in the build container.

Outputs (both committed):
  oracle/m2v_tables.h                              - plain C, test infrastructure
  fpga-mpeg2-encoder_b200/csrc/m2v_tables.cuh      - product (host+device constant tables)
"""
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

# Table B-10 motion_code, index = |motion_code| 0..16, code WITHOUT the sign bit
MOTION = ['1', '01', '001', '0001', '000011', '0000101', '0000100', '0000011',
          '000001011', '000001010', '000001001', '0000010001', '0000010000',
          '0000001111', '0000001110', '0000001101', '0000001100']

# Table B-9 coded_block_pattern, index = cbp 1..63 (4:2:0); cbp 0 has no code in 4:2:0
CBP = {
    60: '111', 4: '1101', 8: '1100', 16: '1011', 32: '1010',
    12: '10011', 48: '10010', 20: '10001', 40: '10000', 28: '01111', 44: '01110', 52: '01101',
    56: '01100', 1: '01011', 61: '01010', 2: '01001', 62: '01000',
    24: '001111', 36: '001110', 3: '001101', 63: '001100',
    5: '0010111', 9: '0010110', 17: '0010101', 33: '0010100', 6: '0010011', 10: '0010010',
    18: '0010001', 34: '0010000',
    7: '00011111', 11: '00011110', 19: '00011101', 35: '00011100', 13: '00011011',
    49: '00011010', 21: '00011001', 41: '00011000', 14: '00010111', 50: '00010110',
    22: '00010101', 42: '00010100', 15: '00010011', 51: '00010010', 23: '00010001',
    43: '00010000', 25: '00001111', 37: '00001110', 26: '00001101', 38: '00001100',
    29: '00001011', 45: '00001010', 53: '00001001', 57: '00001000', 30: '00000111',
    46: '00000110', 54: '00000101', 58: '00000100',
    31: '000000111', 47: '000000110', 55: '000000101', 59: '000000100', 27: '000000011',
    39: '000000010',
}

# Table B-12 dct_dc_size_luminance, B-13 dct_dc_size_chrominance, index = size 0..11
DC_Y = ['100', '00', '01', '101', '110', '1110', '11110', '111110', '1111110', '11111110',
        '111111110', '111111111']
DC_C = ['00', '01', '10', '110', '1110', '11110', '111110', '1111110', '11111110', '111111110',
        '1111111110', '1111111111']

# Table B-14 (DCT coefficients table zero): (run, level) -> code WITHOUT the sign bit.
# (0,1) is the "not first coefficient" form '11'.
B14 = {
    (0, 1): '11', (1, 1): '011', (0, 2): '0100', (2, 1): '0101', (0, 3): '00101',
    (3, 1): '00111', (4, 1): '00110', (1, 2): '000110', (5, 1): '000111', (6, 1): '000101',
    (7, 1): '000100', (0, 4): '0000110', (2, 2): '0000100', (8, 1): '0000111', (9, 1): '0000101',
    (0, 5): '00100110', (0, 6): '00100001', (1, 3): '00100101', (3, 2): '00100100',
    (10, 1): '00100111', (11, 1): '00100011', (12, 1): '00100010', (13, 1): '00100000',
    (0, 7): '0000001010', (1, 4): '0000001100', (2, 3): '0000001011', (4, 2): '0000001111',
    (5, 2): '0000001001', (14, 1): '0000001110', (15, 1): '0000001101', (16, 1): '0000001000',
    (0, 8): '000000011101', (0, 9): '000000011000', (0, 10): '000000010011',
    (0, 11): '000000010000', (1, 5): '000000011011', (2, 4): '000000010100',
    (3, 3): '000000011100', (4, 3): '000000010010', (6, 2): '000000011110',
    (7, 2): '000000010101', (8, 2): '000000010001', (17, 1): '000000011111',
    (18, 1): '000000011010', (19, 1): '000000011001', (20, 1): '000000010111',
    (21, 1): '000000010110',
    (0, 12): '0000000011010', (0, 13): '0000000011001', (0, 14): '0000000011000',
    (0, 15): '0000000010111', (1, 6): '0000000010110', (1, 7): '0000000010101',
    (2, 5): '0000000010100', (3, 4): '0000000010011', (5, 3): '0000000010010',
    (9, 2): '0000000010001', (10, 2): '0000000010000', (22, 1): '0000000011111',
    (23, 1): '0000000011110', (24, 1): '0000000011101', (25, 1): '0000000011100',
    (26, 1): '0000000011011',
    (6, 3): '0000000000010100', (11, 2): '0000000000011010', (12, 2): '0000000000011001',
    (13, 2): '0000000000011000', (14, 2): '0000000000010111', (15, 2): '0000000000010110',
    (16, 2): '0000000000010101', (27, 1): '0000000000011111', (28, 1): '0000000000011110',
    (29, 1): '0000000000011101', (30, 1): '0000000000011100', (31, 1): '0000000000011011',
}
for i, lvl in enumerate(range(16, 32)):      # (0,16)..(0,31): 14 bits, 0b11111 downto 0b10000
    B14[(0, lvl)] = format(0b11111 - i, '014b')
for i, lvl in enumerate(range(32, 41)):      # (0,32)..(0,40): 15 bits, 0b11000 downto 0b10000
    B14[(0, lvl)] = format(0b11000 - i, '015b')
for i, lvl in enumerate(range(8, 15)):       # (1,8)..(1,14): 15 bits, 0b11111 downto 0b11001
    B14[(1, lvl)] = format(0b11111 - i, '015b')
for i, lvl in enumerate(range(15, 19)):      # (1,15)..(1,18): 16 bits, 0b10011 downto 0b10000
    B14[(1, lvl)] = format(0b10011 - i, '016b')

INTRA_Q = [
    [8, 16, 19, 22, 26, 27, 29, 34], [16, 16, 22, 24, 27, 29, 34, 37],
    [19, 22, 26, 27, 29, 34, 34, 38], [22, 22, 26, 27, 29, 34, 37, 40],
    [22, 26, 27, 29, 32, 35, 40, 48], [26, 27, 29, 32, 35, 40, 48, 58],
    [26, 27, 29, 34, 38, 46, 56, 69], [27, 29, 35, 38, 46, 56, 69, 83]]

def zigzag_scan():
    """ISO 13818-2 Figure 7-2 (scan[0]): returns pos[i][j] = index in scan order."""
    pos = [[0] * 8 for _ in range(8)]
    i = j = 0
    for k in range(64):
        pos[i][j] = k
        if (i + j) % 2 == 0:            # moving up-right
            if j == 7: i += 1
            elif i == 0: j += 1
            else: i -= 1; j += 1
        else:                           # moving down-left
            if i == 7: j += 1
            elif j == 0: i += 1
            else: i += 1; j -= 1
    return pos

ZIGZAG = zigzag_scan()

# 8-bit integer DCT basis (the HEVC core-transform 8x8 matrix scaled to 8 bits), RTL:102-112
_c = [64, 89, 84, 75, 64, 50, 35, 18]
import math
def dct_matrix():
    m = []
    for i in range(8):
        row = []
        for k in range(8):
            if i == 0:
                row.append(64)
            else:
                # sign and magnitude follow cos((2k+1) i pi / 16); magnitude class by folded index
                ang = (2 * k + 1) * i
                ang %= 32
                sgn = 1
                if ang > 16: ang = 32 - ang
                if ang > 8: ang = 16 - ang; sgn = -1
                row.append(sgn * _c[ang])
        m.append(row)
    return m
DCTM = dct_matrix()


def pack(code):
    return int(code, 2), len(code)


def build_ac():
    """AC table indexed [run 0..31][level-1 0..39] -> (code,len) excluding sign; len 0 = escape."""
    t = [[(0, 0)] * 40 for _ in range(32)]
    for (run, lvl), code in B14.items():
        t[run][lvl - 1] = pack(code)
    return t


def c_array(name, ctype, vals, per_line=16):
    out = ['static const %s %s[%d] = {' % (ctype, name, len(vals))]
    for i in range(0, len(vals), per_line):
        out.append('    ' + ', '.join(str(v) for v in vals[i:i + per_line]) + ',')
    out.append('};')
    return '\n'.join(out)


def emit(qual, guard, header_note):
    ac = build_ac()
    L = []
    L.append('/* GENERATED by tools/gen_tables.py - do not edit. %s */' % header_note)
    L.append('#ifndef %s\n#define %s\n#include <stdint.h>' % (guard, guard))
    L.append('/* each VLC entry packs (len << 16) | code ; code excludes any sign bit */')
    def vlc(name, pairs):
        L.append(c_array(name, qual + 'uint32_t', [(l << 16) | c for c, l in pairs], 8))
    vlc('M2V_VLC_MOTION', [pack(c) for c in MOTION])                 # Table B-10
    vlc('M2V_VLC_CBP', [(0, 0)] + [pack(CBP[i]) for i in range(1, 64)])  # Table B-9
    vlc('M2V_VLC_DC_Y', [pack(c) for c in DC_Y])                     # Table B-12
    vlc('M2V_VLC_DC_C', [pack(c) for c in DC_C])                     # Table B-13
    flat = []
    for run in range(32):
        flat += ac[run]
    vlc('M2V_VLC_AC', flat)                                          # Table B-14 [run][level-1]
    L.append('#define M2V_AC_LEVELS 40')
    L.append(c_array('M2V_INTRA_Q', qual + 'uint8_t', sum(INTRA_Q, []), 8))
    L.append(c_array('M2V_ZIGZAG', qual + 'uint8_t', sum(ZIGZAG, []), 8))
    L.append(c_array('M2V_DCTM', qual + 'int8_t', sum(DCTM, []), 8))
    L.append('#endif')
    return '\n'.join(L) + '\n'


def main():
    with open(os.path.join(ROOT, 'oracle', 'm2v_tables.h'), 'w') as f:
        f.write(emit('', 'M2V_ORACLE_TABLES_H', 'Oracle copy (test infrastructure).'))
    with open(os.path.join(ROOT, 'fpga-mpeg2-encoder_b200', 'csrc', 'm2v_tables.cuh'), 'w') as f:
        f.write(emit('', 'M2V_PRODUCT_TABLES_CUH', 'Product copy (host side; uploaded to __constant__).'))


if __name__ == '__main__':
    main()
