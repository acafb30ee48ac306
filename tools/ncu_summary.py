#!/usr/bin/env python3
"""Summarise an ncu capture (.ncu-rep made with --set full --import-source on) into the text that
gets committed under profiles/: key metrics per launch, and for one kernel a per-region table
(instructions executed / ALU-pipe instructions / stall samples) where the regions are the
'// ---- ' banner comments of csrc/m2v_kernels.cu.  Needs ncu, cuobjdump and nvdisasm (no GPU).

usage: ncu_summary.py <report.ncu-rep> [kernel-substring-for-region-table, e.g. k1_mb_encodeILi3ELb1E]
"""
import collections
import csv
import os
import re
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
# the capture must be read with the library (and source) it was taken from: M2V_LIB / M2V_SRC point at an older build
SRC = os.environ.get('M2V_SRC', os.path.join(ROOT, 'fpga-mpeg2-encoder_b200', 'csrc', 'm2v_kernels.cu'))
LIB = os.environ.get('M2V_LIB', os.path.join(ROOT, 'fpga-mpeg2-encoder_b200', 'libm2venc.so'))
METRICS = ['gpu__time_duration.sum', 'launch__grid_size', 'launch__registers_per_thread', 'launch__occupancy_limit_shared_mem',
           'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum',
           'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active',
           'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active',
           'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
           'lts__t_bytes.sum', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum']
ALU = {'VABSDIFF4', 'SHF', 'IADD3', 'ISETP', 'LOP3', 'VIMNMX', 'VIADD', 'SEL', 'PRMT', 'LEA', 'PLOP3', 'IABS', 'VIADDMNMX', 'MOV',
       'FLO', 'POPC', 'BREV', 'SGXT', 'BMSK', 'VIMNMX3', 'P2R', 'R2P', 'IADD', 'LOP'}


def run(cmd):
    return subprocess.run(cmd, capture_output=True, text=True).stdout


def main():
    rep = sys.argv[1]
    kern = sys.argv[2] if len(sys.argv) > 2 else None
    raw = list(csv.reader(run(['ncu', '-i', rep, '--page', 'raw', '--csv']).splitlines()))
    hdr, units = raw[0], raw[1]
    print('# %s' % os.path.basename(rep))
    for r in raw[2:]:
        print('## launch %s  %s' % (r[hdr.index('ID')], r[hdr.index('Kernel Name')][:70]))
        for m in METRICS:
            if m in hdr:
                print('  %-72s %s %s' % (m, r[hdr.index(m)], units[hdr.index(m)]))
    if not kern:
        return
    # SASS <-> source line map from the in-tree library
    tmp = tempfile.mkdtemp()
    subprocess.run(['cuobjdump', '-xelf', 'all', LIB], cwd=tmp, capture_output=True)
    seq, cur, line = [], None, None
    for cub in sorted(os.listdir(tmp)):
        for l in run(['nvdisasm', '-g', '-c', os.path.join(tmp, cub)]).splitlines():
            m = re.match(r'\s*\.text\.(\S+):', l)
            if m:
                cur = m.group(1); continue
            m = re.search(r'//## File "(.*?)", line (\d+)', l)
            if m:
                # lines of other files (CUDA intrinsic headers) are treated like inlined helpers
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
        # inlined helpers (defined above the first banner) carry their own line numbers; attribute them
        # to the region of the nearest preceding instruction that belongs to the kernel body
        if ln is None or ln < first_banner:
            return state['last']
        name = 'prologue'
        for b, t in banners:
            if ln >= b:
                name = 'L%d %s' % (b, t)
        state['last'] = name
        return name
    sass = list(csv.reader(run(['ncu', '-i', rep, '--page', 'source', '--csv']).splitlines()))
    # first kernel block whose name matches
    blocks, curb = [], None
    for r in sass:
        if r and r[0] == 'Kernel Name':
            curb = {'name': r[1], 'rows': []}; blocks.append(curb)
        elif curb is not None:
            curb['rows'].append(r)
    want = kern.replace('ILi', '<').split('<')[0].replace('_Z12', '')
    if 'k1' in kern:
        pf = 'Lb1' in kern or 'Lb' not in kern                   # P-frame instantiation unless the mangled name says <.., false>
        tag = ('true', '1>') if pf else ('false', '0>')
        blk = next((b for b in blocks if 'k1_mb_encode' in b['name'] and any(t in b['name'] for t in tag)), blocks[0])
    else:
        base = kern.split('I')[0].lstrip('_Z0123456789')
        tag = '(bool)1' if 'Lb1' in kern else '(bool)0' if 'Lb0' in kern else ''
        blk = next((b for b in blocks if base in b['name'] and tag in b['name']), blocks[0])
    h = blk['rows'][0]
    ie, ws = h.index('Instructions Executed'), h.index('Warp Stall Sampling (All Samples)')
    rows = [r for r in blk['rows'][1:] if len(r) == len(h)][:len(seq)]
    tot, alu, smp, ops = collections.Counter(), collections.Counter(), collections.Counter(), collections.Counter()
    regops = collections.defaultdict(collections.Counter)
    for (ln, txt), r in zip(seq, rows):
        c = int(r[ie]); reg = region(ln)
        tot[reg] += c; smp[reg] += int(r[ws])
        op = re.sub(r'^@!?U?P\w+\s+', '', txt).split()[0].split('.')[0]
        ops[op] += c
        regops[reg][op] += c
        if op in ALU:
            alu[reg] += c
    first = int(rows[0][ie]) or 1                     # executions of the first instruction = warps launched
    grid_warps = first
    units = collections.Counter(int(r[ie]) for r in rows).most_common(1)[0][0] or 1   # modal execution count = work items (macroblocks)
    T, A, S = sum(tot.values()), sum(alu.values()), max(sum(smp.values()), 1)
    print('\n## region table for %s (%s), SASS instrs matched: %d' % (kern, blk['name'][:60], len(rows)))
    print('   warp-instructions executed: %d total, %d on the ALU pipe; warps: %d; work items (modal count): %d -> %.0f instr, %.0f ALU per item' % (T, A, grid_warps, units, T / units, A / units))
    for reg in sorted(tot, key=lambda k: int(re.match(r'L(\d+)', k).group(1)) if k[0] == 'L' else 0):
        print('   %-72s %5.1f%% instr  %5.1f%% alu  %5.1f%% stall-samples' % (reg, 100 * tot[reg] / T, 100 * alu[reg] / max(A, 1), 100 * smp[reg] / S))
        print('       per work item: ' + ', '.join('%s %.1f' % (k, v / units) for k, v in regops[reg].most_common(10)))
    if os.environ.get('NCU_LINES'):
        # per source line: executed warp-instructions per work item and the opcodes behind them
        perline = collections.defaultdict(collections.Counter)
        for (ln, txt), r in zip(seq, rows):
            op = re.sub(r'^@!?U?P\w+\s+', '', txt).split()[0].split('.')[0]
            perline[ln][op] += int(r[ie])
        print('   per source line (instr per work item):')
        for ln in sorted(perline, key=lambda k: (k is None, k)):
            c = perline[ln]; t = sum(c.values())
            if t / units < 0.5:
                continue
            text = src[ln - 1].strip()[:70] if ln else '(inlined helper / other file)'
            print('     %5s %7.1f  %-70s %s' % (ln, t / units, text, ', '.join('%s %.0f' % (k, v / units) for k, v in c.most_common(5))))
    print('   opcode mix (%% of executed): ' + ', '.join('%s %.1f' % (k, 100 * v / T) for k, v in ops.most_common(16)))


if __name__ == '__main__':
    main()
