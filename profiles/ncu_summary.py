"""Key counters of one launch from an `ncu --set full` report: python profiles/ncu_summary.py report.ncu-rep out.json [launch]"""
import csv, io, json, subprocess, sys
rep, out = sys.argv[1], sys.argv[2]
which = int(sys.argv[3]) if len(sys.argv) > 3 else 0
txt = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
hdr, units, vals = rows[0], rows[1], rows[2 + which]
want = {'gpu__time_duration.sum': 'duration', 'dram__bytes_read.sum': 'dram_read', 'dram__bytes_write.sum': 'dram_write',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed': 'dram_pct',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active': 'tensor_pipe_pct_active',
        'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active': 'xu_mufu_pct',
        'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active': 'fma_pct',
        'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active': 'alu_pct',
        'smsp__issue_active.avg.pct_of_peak_sustained_active': 'issue_pct',
        'l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed': 'l1_data_pipe_pct',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum': 'smem_bank_conflicts',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum': 'smem_wavefronts',
        'sm__warps_active.avg.pct_of_peak_sustained_active': 'warps_active_pct',
        'launch__registers_per_thread': 'registers', 'launch__grid_size': 'grid', 'launch__block_size': 'block',
        'lts__t_sector_hit_rate.pct': 'l2_hit_pct', 'smsp__inst_executed.sum': 'warp_instructions',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed': 'sm_throughput_pct'}
res = {'report': rep, 'launch': which}
for h, u, v in zip(hdr, units, vals):
    if h in want:
        res[want[h]] = {'value': v, 'unit': u}
    if h == 'Kernel Name':
        res['kernel'] = v.split('(')[0]
json.dump(res, open(out, 'w'), indent=1)
print(json.dumps(res, indent=1))
