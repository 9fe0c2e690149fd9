"""Summarise an .ncu-rep (read on the CPU box): key raw metrics per kernel, opcode histogram and
the hottest SASS lines with their stall samples.  usage: ncu_summary.py rep [kernel-regex]"""
import collections
import csv
import io
import re
import subprocess
import sys

WANT = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'dram__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread',
        'smsp__inst_executed.sum', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
        'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_fmaheavy.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active',
        'sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active',
        'l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed',
        'l1tex__lsu_writeback_active.avg.pct_of_peak_sustained_elapsed',
        'smsp__cycles_active.avg', 'sm__cycles_elapsed.avg', 'lts__t_bytes.sum']


def run(args):
    return subprocess.run(['ncu'] + args, capture_output=True, text=True).stdout


def main():
    rep = sys.argv[1]
    kre = sys.argv[2] if len(sys.argv) > 2 else None
    rows = list(csv.reader(io.StringIO(run(['-i', rep, '--page', 'raw', '--csv']))))
    hdr, units = rows[0], rows[1]
    kn = hdr.index('Kernel Name')
    for r in rows[2:]:
        if kre and not re.search(kre, r[kn]):
            continue
        print('==', r[kn][:70])
        for w in WANT:
            if w in hdr:
                i = hdr.index(w)
                print('   %-78s %s %s' % (w, r[i], units[i]))
        # stall breakdown
        st = [(hdr[i], float(r[i])) for i in range(len(hdr))
              if hdr[i].startswith('smsp__average_warps_issue_stalled') and hdr[i].endswith('_per_issue_active.ratio') and r[i]]
        st.sort(key=lambda x: -x[1])
        for n, v in st[:8]:
            print('   stall %-60s %.2f' % (n.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', ''), v))
    if kre:
        out = run(['-i', rep, '--page', 'source', '--csv', '--kernel-name', 'regex:' + kre])
        rows = list(csv.reader(io.StringIO(out)))
        hd = [i for i, r in enumerate(rows) if r and r[0] == 'Address']
        if not hd:
            return
        h = rows[hd[0]]
        end = hd[1] - 1 if len(hd) > 1 else len(rows)
        ia, isrc, isamp = h.index('Instructions Executed'), h.index('Source'), h.index('# Samples')
        data = [r for r in rows[hd[0] + 1:end] if len(r) > ia and r[ia].isdigit()]
        tot = sum(int(r[ia]) for r in data)
        tsamp = sum(int(r[isamp]) for r in data if r[isamp].isdigit())
        ops = collections.Counter()
        samp = collections.Counter()
        for r in data:
            m = re.match(r'\s*(@!?U?P\d\s+)?([A-Z0-9_.]+)', r[isrc])
            op = m.group(2).split('.')[0] if m else '?'
            ops[op] += int(r[ia])
            samp[op] += int(r[isamp]) if r[isamp].isdigit() else 0
        print('-- executed warp-instructions %d, stall samples %d' % (tot, tsamp))
        for k, v in ops.most_common(22):
            print('   %-10s %6.2f%% inst  %6.2f%% samples' % (k, 100.0 * v / tot, 100.0 * samp[k] / max(tsamp, 1)))
        print('-- hottest lines by samples')
        order = sorted(range(len(data)), key=lambda i: -(int(data[i][isamp]) if data[i][isamp].isdigit() else 0))
        for i in order[:40]:
            r = data[i]
            print('   %5d %6.2f%% smp %6.2f%% inst  %s' % (i, 100.0 * int(r[isamp]) / max(tsamp, 1), 100.0 * int(r[ia]) / tot, r[isrc][:90]))


if __name__ == '__main__':
    main()
