"""Summarise an `ncu --page raw --csv` dump: one block of key metrics per distinct kernel."""
import csv
import sys

WANT = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size',
        'smsp__inst_executed.sum', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_fma.sum', 'sm__inst_executed_pipe_fp64.sum', 'sm__inst_executed_pipe_lsu.sum',
        'sm__inst_executed_pipe_alu.sum', 'sm__inst_executed_pipe_xu.sum',
        'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum',
        'smsp__thread_inst_executed_per_inst_executed.ratio',
        'sm__cycles_elapsed.avg', 'smsp__cycles_active.avg']
STALL = 'smsp__average_warps_issue_stalled_'


def main(path):
    rows = list(csv.reader(open(path)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    ki = hdr.index('Kernel Name')
    seen = set()
    for r in data:
        k = r[ki][:48]
        if k in seen:
            continue
        seen.add(k)
        print('==', k)
        for w in WANT:
            if w in hdr:
                i = hdr.index(w)
                print(f"   {w:72s} {r[i]:>16s} {units[i]}")
        stalls = []
        for i, h in enumerate(hdr):
            if h.startswith(STALL) and h.endswith('_per_issue_active.ratio'):
                try:
                    stalls.append((float(r[i].replace(',', '')), h[len(STALL):-len('_per_issue_active.ratio')]))
                except ValueError:
                    pass
        stalls.sort(reverse=True)
        print('   top stalls (warps per issue):', ', '.join(f"{n}={v:.2f}" for v, n in stalls[:6]))


if __name__ == '__main__':
    main(sys.argv[1])
