"""Summarise an .ncu-rep (read with `ncu -i`): key raw metrics + executed instructions by opcode and by
source line.  Usage: python profiles/ncu_summarize.py gpurun_out/X.ncu-rep [out.json]"""
import collections, csv, io, json, re, subprocess, sys

KEEP = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread',
        'sm__inst_executed.avg.per_cycle_active', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed',
        'smsp__thread_inst_executed_per_inst_executed.ratio', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
        'smsp__inst_executed.sum', 'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem',
        'sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active', 'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'launch__grid_size', 'launch__block_size', 'launch__shared_mem_per_block_dynamic', 'lts__t_bytes.sum',
        'l1tex__t_bytes_pipe_lsu_mem_global_op_ld.sum', 'l1tex__t_bytes_pipe_lsu_mem_global_op_st.sum']


def run(args):
    return subprocess.run(['ncu'] + args, capture_output=True, text=True).stdout


def main():
    rep = sys.argv[1]
    rows = list(csv.reader(io.StringIO(run(['-i', rep, '--page', 'raw', '--csv']))))
    hdr, units, vals = rows[0], rows[1], rows[2]
    out = {h: [v, u] for h, u, v in zip(hdr, units, vals) if h in KEEP or h.startswith('smsp__average_warps_issue_stalled')}
    out['kernel'] = vals[hdr.index('Kernel Name')]
    rows = list(csv.reader(io.StringIO(run(['-i', rep, '--page', 'source', '--csv']))))
    hdr = rows[1]
    ix = {h: i for i, h in enumerate(hdr)}
    byop, thr, tot = collections.Counter(), collections.Counter(), 0
    for r in rows[2:]:
        try:
            n, t = int(r[ix['Instructions Executed']]), int(r[ix['Thread Instructions Executed']])
        except (ValueError, IndexError):
            continue
        toks = r[ix['Source']].split()
        op = (toks[1] if toks[0].startswith('@') else toks[0]).split('.')[0]
        byop[op] += n; thr[op] += t; tot += n
    out['warp_instructions'] = tot
    out['by_opcode'] = {op: {'pct': round(100 * n / tot, 2), 'avg_threads': round(thr[op] / max(n, 1), 1)} for op, n in byop.most_common(24)}
    print(json.dumps(out, indent=1, sort_keys=True))
    if len(sys.argv) > 2:
        json.dump(out, open(sys.argv[2], 'w'), indent=1, sort_keys=True)


if __name__ == '__main__':
    main()
