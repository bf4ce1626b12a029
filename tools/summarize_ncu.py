"""Turn the scratch ncu outputs in gpurun_out/ into the small tracked summaries under profiles/.

  python tools/summarize_ncu.py <tag>

* gpurun_out/launches_<tag>.csv  (ncu --metrics gpu__time_duration.sum over `bench.py`)  -> per-kernel launch
  counts, total time and SHARE of the step  -> profiles/launches_<tag>.md (+ the raw csv, it is small)
* gpurun_out/prof_<name>_<tag>.ncu-rep (one `--set full` capture per kernel) -> the metrics the roofline rows
  in DESIGN.md / bench.py quote  -> profiles/ncu_<tag>.md
"""
import csv
import glob
import io
import os
import re
import shutil
import subprocess
import sys
from collections import OrderedDict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, 'gpurun_out')
PROF = os.path.join(ROOT, 'profiles')

KEYS = [
    ('gpu__time_duration.sum', 'duration'),
    ('launch__grid_size', 'grid'),
    ('launch__block_size', 'block'),
    ('launch__registers_per_thread', 'regs/thread'),
    ('launch__occupancy_limit_registers', 'occ. limit (regs), CTAs/SM'),
    ('launch__occupancy_limit_shared_mem', 'occ. limit (smem), CTAs/SM'),
    ('sm__warps_active.avg.pct_of_peak_sustained_active', 'achieved occupancy %'),
    ('sm__throughput.avg.pct_of_peak_sustained_elapsed', 'SM throughput % of peak'),
    ('smsp__issue_active.avg.pct_of_peak_sustained_active', 'issue slots busy %'),
    ('sm__ops_path_tensor_src_fp64.sum.pct_of_peak_sustained_elapsed', 'FP64 tensor (DMMA) pipe % of peak'),
    ('sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active', 'FP64 ALU pipe active %'),
    ('smsp__sass_thread_inst_executed_op_dfma_pred_on.sum', 'DFMA thread-instructions'),
    ('smsp__sass_thread_inst_executed_op_dmul_pred_on.sum', 'DMUL thread-instructions'),
    ('smsp__sass_thread_inst_executed_op_dadd_pred_on.sum', 'DADD thread-instructions'),
    ('l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed', 'L1 / shared-memory data pipe (wavefronts) % of peak'),
    ('l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'shared-memory wavefronts'),
    ('dram__bytes_read.sum', 'DRAM read'),
    ('dram__bytes_write.sum', 'DRAM write'),
    ('dram__throughput.avg.pct_of_peak_sustained_elapsed', 'DRAM throughput % of peak'),
    ('lts__t_sector_hit_rate.pct', 'L2 hit rate %'),
    ('l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'smem bank conflicts'),
    ('smsp__sass_inst_executed_op_local_ld.sum', 'local-memory load instr'),
    ('smsp__sass_inst_executed_op_local_st.sum', 'local-memory store instr'),
    ('smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio', 'stall math_pipe_throttle / issue'),
    ('smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio', 'stall long_scoreboard / issue'),
    ('smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio', 'stall short_scoreboard / issue'),
    ('smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio', 'stall barrier / issue'),
    ('smsp__average_warps_issue_stalled_wait_per_issue_active.ratio', 'stall wait / issue'),
    ('smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio', 'stall lg_throttle / issue'),
]


def raw_page(rep):
    txt = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    if len(rows) < 3:
        return None
    hdr, units, vals = rows[0], rows[1], rows[2]
    return {h: (vals[i], units[i]) for i, h in enumerate(hdr)}


def short_name(full):
    m = re.match(r'(?:void )?(?:b200gp::)?([A-Za-z0-9_]+)', full)
    return m.group(1) if m else full


def launches(tag):
    path = os.path.join(OUT, 'launches_%s.csv' % tag)
    if not os.path.exists(path):
        return None
    lines = [l for l in open(path) if l.startswith('"')]
    rows = list(csv.DictReader(io.StringIO(''.join(lines))))
    agg = OrderedDict()
    for r in rows:
        if r['Metric Name'] != 'gpu__time_duration.sum':
            continue
        k = short_name(r['Kernel Name'])
        ns = float(r['Metric Value'].replace(',', ''))
        a = agg.setdefault(k, [0, 0.0])
        a[0] += 1
        a[1] += ns
    total = sum(a[1] for a in agg.values())
    out = ['# ncu launch list of `python bench.py --steps 2 --warmup 3 --no-cpu --no-fit` (%s)' % tag, '',
           '`ncu --metrics gpu__time_duration.sum --clock-control none`; per-launch times under ncu are serialised and '
           'cold-cache, so read the SHARE column, not the absolute times.  Raw list: `launches_%s.csv`.' % tag, '',
           '| kernel | launches | total ms | share |', '|---|---:|---:|---:|']
    for k, (n, ns) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        out.append('| `%s` | %d | %.2f | %.1f%% |' % (k, n, ns / 1e6, 100 * ns / total))
    out.append('| **all** | %d | %.2f | 100%% |' % (sum(a[0] for a in agg.values()), total / 1e6))
    dm = sum(ns for k, (n, ns) in agg.items() if k in ('chol_update_kernel', 'inv_update_kernel', 'kinv_kernel',
                                                     'chol_update64_kernel', 'inv_update64_kernel', 'kinv64_kernel',
                                                     'chol_update64_tma_kernel', 'inv_update64_tma_kernel', 'kinv64_tma_kernel',
                                                     'trsm64_kernel', 'chol_trsm_kernel', 'inv_trsm_kernel', 'gemm_kernel'))
    out += ['', 'DMMA tile-GEMM kernels together: %.1f%% of the kernel time.' % (100 * dm / total)]
    shutil.copy(path, os.path.join(PROF, 'launches_%s.csv' % tag))
    return '\n'.join(out) + '\n'


def main():
    tag = sys.argv[1]
    os.makedirs(PROF, exist_ok=True)
    txt = launches(tag)
    if txt:
        open(os.path.join(PROF, 'launches_%s.md' % tag), 'w').write(txt)
    out = ['# ncu `--set full --clock-control none` captures (%s)' % tag, '',
           'Commands: `python tools/perf_probe.py 2048 32 1` (32 C2 kernels, N = 2048, one NLML+grad evaluation) for the scoring '
           'kernels, `python tools/perf_c5.py 16384` for `c5upd` (the first K = 512 trailing update of the N = 16384 Cholesky), '
           '`python tools/bench_configs.py --configs c3` for the Hellinger kernels; one launch of each kernel; script '
           '`tools/ncu_r02.sh`.  Reports themselves stay in `gpurun_out/` (scratch, MBs each).', '']
    traffic = {'source': 'ncu --set full --clock-control none, python tools/perf_probe.py 2048 32 1 (32 items per launch), tag %s; '
                         'see profiles/ncu_%s.md' % (tag, tag), 'items_per_launch': 32}
    for rep in sorted(glob.glob(os.path.join(OUT, 'prof_*_%s.ncu-rep' % tag))):
        d = raw_page(rep)
        name = os.path.basename(rep)[5:-len('_%s.ncu-rep' % tag)]
        if not d:
            out += ['## %s: unreadable' % name, '']
            continue
        kname = d.get('Kernel Name', ('?', ''))[0]
        out += ['## %s — `%s`' % (name, short_name(kname)), '', '| metric | value |', '|---|---:|']
        for key, label in KEYS:
            if key in d and d[key][0] != '':
                v, u = d[key]
                out.append('| %s (`%s`) | %s %s |' % (label, key, v, u))
        # derived roofline figures: FP64 ALU flops from the SASS op counts, DRAM bandwidth from the byte counters
        def num(key):
            try:
                v, u = d[key]
                f = float(v.replace(',', ''))
                scale = {'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9, 'Tbyte': 1e12, 'byte': 1.0, 'us': 1e-6, 'ms': 1e-3, 'ns': 1e-9,
                         's': 1.0, 'usecond': 1e-6, 'msecond': 1e-3, 'nsecond': 1e-9, 'second': 1.0}.get(u, 1.0)
                return f * scale
            except Exception:
                return None
        t = num('gpu__time_duration.sum')
        if name in ('kinv', 'grad', 'build') and t:
            pipe = d.get('sm__ops_path_tensor_src_fp64.avg.pct_of_peak_sustained_elapsed' if name == 'kinv'
                         else 'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active', ('nan', ''))[0]
            traffic[{'kinv': 'kinv64_kernel', 'grad': 'poly_grad_kernel', 'build': 'poly_build_kernel'}[name]] = {
                'dram_bytes_read': num('dram__bytes_read.sum'), 'dram_bytes_write': num('dram__bytes_write.sum'),
                'duration_ms': t * 1e3, 'pipe_pct': float(pipe.replace(',', '')) if pipe not in ('', 'nan') else None}
        dfma, dmul, dadd = (num('smsp__sass_thread_inst_executed_op_dfma_pred_on.sum.per_cycle_elapsed'),
                            num('smsp__sass_thread_inst_executed_op_dmul_pred_on.sum.per_cycle_elapsed'),
                            num('smsp__sass_thread_inst_executed_op_dadd_pred_on.sum.per_cycle_elapsed'))
        peak_ipc = num('sm__sass_thread_inst_executed_op_dfma_pred_on.sum.peak_sustained')      # thread-instr / cycle, GPU-wide
        clk = None
        try:
            v, u = d['smsp__cycles_elapsed.avg.per_second']
            clk = float(v.replace(',', '')) * {'cycle/nsecond': 1e9, 'cycle/usecond': 1e6, 'cycle/second': 1.0, 'Ghz': 1e9, 'Mhz': 1e6, 'hz': 1.0}.get(u, 1e9)
        except Exception:
            pass
        rd, wr = num('dram__bytes_read.sum'), num('dram__bytes_write.sum')
        if t:
            if None not in (dfma, dmul, dadd, peak_ipc) and (dfma + dmul + dadd) > 0:
                fl = 2 * dfma + dmul + dadd
                extra = ''
                if clk:
                    extra = ' = %.2f TFLOP/s at %.2f GHz' % (fl * clk / 1e12, clk / 1e9)
                out.append('| **FP64 ALU throughput** (DFMA / DMUL / DADD thread-instructions per cycle: %.0f / %.0f / %.0f of %.0f issue slots) | '
                           '%.0f %% of the FP64 issue slots; %.0f flop/cycle%s |'
                           % (dfma, dmul, dadd, peak_ipc, 100 * (dfma + dmul + dadd) / peak_ipc, fl, extra))
            if None not in (rd, wr):
                out.append('| **DRAM bandwidth** (read + write / duration) | %.0f GB/s = %.1f %% of the 6 544 GB/s measured copy peak |'
                           % ((rd + wr) / t / 1e9, 100 * (rd + wr) / t / 6544e9))
        out.append('')
    open(os.path.join(PROF, 'ncu_%s.md' % tag), 'w').write('\n'.join(out) + '\n')
    if 'kinv64_kernel' in traffic:
        import json
        json.dump(traffic, open(os.path.join(PROF, 'ncu_traffic_%s.json' % tag), 'w'), indent=1)
    print('\n'.join(out[:80]))
    if txt:
        print(txt)


if __name__ == '__main__':
    main()
