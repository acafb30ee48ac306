"""A from-the-standard parser of the MPEG-2 video syntax subset the encoder emits (ISO/IEC 13818-2
6.2: sequence / GOP / picture / slice / macroblock / block layers, frame pictures, frame_pred_frame_dct=1,
intra_vlc_format=0, 4:2:0).  TEST INFRASTRUCTURE: it walks a stream back into per-macroblock
(type, motion vector, coded_block_pattern, quantised levels) so that the entropy layer of the oracle can be
checked against the standard's syntax independently of the code that wrote it.  No reconstruction.
"""
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tools'))
import gen_tables as G


class Bits:
    def __init__(self, data):
        self.d, self.p = data, 0

    def get(self, n):
        v = 0
        for _ in range(n):
            v = (v << 1) | ((self.d[self.p >> 3] >> (7 - (self.p & 7))) & 1)
            self.p += 1
        return v

    def peek(self, n):
        p = self.p; v = self.get(n); self.p = p
        return v

    def align(self):
        while self.p & 7:
            assert self.get(1) == 0, 'non-zero stuffing'

    def next_start_code(self):
        self.align()
        while self.peek(24) != 1:
            assert self.get(8) == 0, 'garbage before start code'


def _decoder(table):
    """table: {symbol: bitstring} -> function(Bits) -> symbol"""
    inv = {v: k for k, v in table.items()}
    mx = max(len(v) for v in inv)

    def dec(b):
        s = ''
        for _ in range(mx):
            s += str(b.get(1))
            if s in inv:
                return inv[s]
        raise ValueError('bad VLC ' + s)
    return dec


DEC_MOTION = _decoder({i: c for i, c in enumerate(G.MOTION)})
DEC_CBP = _decoder(dict(G.CBP))
DEC_DC_Y = _decoder({i: c for i, c in enumerate(G.DC_Y)})
DEC_DC_C = _decoder({i: c for i, c in enumerate(G.DC_C)})
_AC = {('EOB',): '10', ('ESC',): '000001'}
_AC.update({k: v for k, v in G.B14.items()})
DEC_AC = _decoder(_AC)
MBTYPE_I = _decoder({'intra': '1', 'intra_q': '01'})                       # Table B-2
MBTYPE_P = _decoder({'mc_coded': '1', 'nomc_coded': '01', 'mc_notcoded': '001', 'intra': '00011',
                     'mc_coded_q': '00010', 'nomc_coded_q': '00001', 'intra_q': '000001'})   # Table B-3


def _block(b, intra, is_luma, dc_pred):
    lv = [0] * 64
    i = 0
    if intra:
        size = (DEC_DC_Y if is_luma else DEC_DC_C)(b)
        diff = 0
        if size:
            v = b.get(size)
            diff = v if v >> (size - 1) else v - (1 << size) + 1
        lv[0] = dc_pred + diff
        i = 1
        first = False
    else:
        first = True
    while True:
        if first and b.peek(1) == 1:                     # first coefficient of a non-intra block: '1s' = +-1
            b.get(1)
            run, level = 0, (-1 if b.get(1) else 1)
        else:
            sym = DEC_AC(b)
            if sym == ('EOB',):
                assert not first, 'EOB as first code of a non-intra block'
                break
            if sym == ('ESC',):
                run = b.get(6); level = b.get(12)
                level = level - 4096 if level >= 2048 else level
                assert level not in (0, -2048)
            else:
                run, level = sym
                if b.get(1):
                    level = -level
        first = False
        i += run
        assert i < 64, 'run past the block'
        lv[i] = level
        i += 1
    return lv


def parse(data):
    """-> dict(width, height, pictures=[dict(type, temporal_reference, gop_header or None, mbs=[...])])"""
    b = Bits(data)
    out = {'pictures': []}
    gop = None
    b.next_start_code()
    assert b.get(32) == 0x1B3
    W, H = b.get(12), b.get(12)
    out['width'], out['height'] = W, H
    b.get(4); out['frame_rate_code'] = b.get(4); b.get(18); assert b.get(1) == 1; b.get(10); b.get(1)
    assert b.get(1) == 0 and b.get(1) == 0               # no custom matrices
    b.next_start_code()
    assert b.get(32) == 0x1B5 and b.get(4) == 1          # sequence_extension
    out['profile_level'] = b.get(8); out['progressive_sequence'] = b.get(1); assert b.get(2) == 1   # 4:2:0
    b.get(2); b.get(2); b.get(12); assert b.get(1) == 1; b.get(8); b.get(1); b.get(2); b.get(5)
    b.next_start_code()
    assert b.get(32) == 0x1B5 and b.get(4) == 2          # sequence_display_extension
    b.get(3); assert b.get(1) == 1; b.get(24); assert b.get(14) == W; assert b.get(1) == 1; assert b.get(14) == H
    mbw, mbh = W // 16, H // 16
    while True:
        b.next_start_code()
        code = b.get(32)
        if code == 0x1B7:                                # sequence_end_code
            break
        if code == 0x1B8:
            tc = b.get(25); closed, broken = b.get(1), b.get(1)
            gop = {'hours': (tc >> 19) & 31, 'minutes': (tc >> 13) & 63, 'seconds': (tc >> 6) & 63, 'pictures': tc & 63,
                   'drop': tc >> 24, 'closed_gop': closed, 'broken_link': broken}
            assert (tc >> 12) & 1 == 1
            continue
        assert code == 0x100, hex(code)
        pic = {'temporal_reference': b.get(10), 'type': b.get(3), 'gop': gop, 'mbs': []}
        gop = None
        b.get(16)
        if pic['type'] == 2:
            assert b.get(1) == 0 and b.get(3) == 7       # full_pel_forward_vector, forward_f_code = 111
        assert b.get(1) == 0                             # extra_bit_picture
        b.next_start_code()
        assert b.get(32) == 0x1B5 and b.get(4) == 8      # picture_coding_extension
        fcodes = [b.get(4) for _ in range(4)]
        pic['f_code'] = fcodes
        assert b.get(2) == 2 and b.get(2) == 3           # intra_dc_precision 10 bit, frame picture
        flags = [b.get(1) for _ in range(10)]            # tff, frame_pred_frame_dct, conceal, q_scale_type, intra_vlc_format, alt_scan, rff, chroma420, progressive_frame, composite
        assert flags[1] == 1 and flags[2] == 0 and flags[3] == 0 and flags[4] == 0 and flags[5] == 0 and flags[9] == 0
        for row in range(mbh):
            b.next_start_code()
            assert b.get(32) == 0x100 + row + 1, 'slice order'
            qsc = b.get(5); assert b.get(1) == 0
            dc = [512, 512, 512]; pmv = [0, 0]
            for col in range(mbw):
                assert b.get(1) == 1, 'macroblock_address_increment != 1'
                t = (MBTYPE_I if pic['type'] == 1 else MBTYPE_P)(b)
                assert not t.endswith('_q') and not t.startswith('nomc')
                mb = {'type': t, 'qsc': qsc}
                intra = t == 'intra'
                if t.startswith('mc'):
                    mv = []
                    for k in range(2):
                        code = DEC_MOTION(b)
                        if code and b.get(1):
                            code = -code
                        v = pmv[k] + code                # f_code 1: no residual, range [-16, 15]
                        v = v + 32 if v < -16 else v - 32 if v > 15 else v
                        pmv[k] = v; mv.append(v)
                    mb['mv'] = mv
                else:
                    pmv = [0, 0]
                cbp = 63 if intra else (DEC_CBP(b) if t == 'mc_coded' else 0)
                mb['cbp'] = cbp
                levels = []
                for blk in range(6):
                    if not (cbp >> (5 - blk)) & 1:
                        levels.append([0] * 64); continue
                    c = 0 if blk < 4 else blk - 3
                    lv = _block(b, intra, blk < 4, dc[c])
                    if intra:
                        dc[c] = lv[0]
                    levels.append(lv)
                if not intra:
                    dc = [512, 512, 512]
                mb['levels'] = levels
                pic['mbs'].append(mb)
        out['pictures'].append(pic)
    b.align()
    assert all(x == 0 for x in data[b.p >> 3:]), 'non-zero bytes after the end code'
    return out
