"""Summaries kept under profiles/ from the scratch captures in gpurun_out/:
   python profiles/summarise_ncu.py <tag> [kernel-label]    (reads gpurun_out/prof_<tag>.ncu-rep, launches_<tag>.csv,
   bench_<tag>.log; writes profiles/<tag>_ncu_full_<kernel-label>.csv, label default pk_step_kernel)"""
import csv, collections, os, subprocess, sys
tag = sys.argv[1]
label = sys.argv[2] if len(sys.argv) > 2 else 'pk_step_kernel'
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
go = os.path.join(root, 'gpurun_out')
raw = subprocess.run(['ncu', '-i', os.path.join(go, 'prof_%s.ncu-rep' % tag), '--page', 'raw', '--csv'],
                     capture_output=True, text=True).stdout
rows = [r for r in csv.reader(raw.splitlines()) if r]
hdr, units, vals = rows[0], rows[1], rows[2]
keep = ['Kernel Name', 'Block Size', 'Grid Size', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'gpu__time_duration.sum',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'launch__occupancy_limit_registers',
        'launch__occupancy_limit_shared_mem', 'launch__registers_per_thread', 'launch__shared_mem_per_block_dynamic',
        'lts__t_sector_hit_rate.pct', 'sass__inst_executed_local_loads', 'sass__inst_executed_local_stores',
        'sass__inst_executed_shared_loads', 'sass__inst_executed_shared_stores', 'sm__cycles_elapsed.max',
        'sm__inst_issued.avg.pct_of_peak_sustained_active', 'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'smsp__inst_executed.sum', 'sm__icc_request_hit_rate.pct', 'sm__icc_requests.sum.pct_of_peak_sustained_elapsed',
        'gcc__cache_requests_type_instruction.sum', 'gcc__cache_requests_type_instruction.sum.pct_of_peak_sustained_elapsed',
        'smsp__issue_active.avg.per_cycle_active']
with open(os.path.join(root, 'profiles', '%s_ncu_full_%s.csv' % (tag, label)), 'w') as f:
    f.write('metric,unit,value\n')
    for h, u, v in zip(hdr, units, vals):
        if h in keep or (h.startswith('smsp__average_warps_issue_stalled') and h.endswith('per_issue_active.ratio')):
            f.write('%s,%s,%s\n' % (h, u, v.replace(',', ';')))
            print(h, v)
# dram traffic per launch of the headline kernel -> profiles/traffic.json (bench.py's roofline.traffic reads it)
if label == 'pk_step_kernel':
    import json
    scale = {'byte': 1.0, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}
    col = {h: (u, v) for h, u, v in zip(hdr, units, vals)}
    tot = sum(float(col[k][1].replace(',', '')) * scale[col[k][0]] for k in ('dram__bytes_read.sum', 'dram__bytes_write.sum'))
    json.dump({'pk_step_kernel': {'dram_bytes_per_launch': tot, 'source': 'profiles/%s_ncu_full_%s.csv' % (tag, label),
                                  'kernel': col['Kernel Name'][1]}},
              open(os.path.join(root, 'profiles', 'traffic.json'), 'w'), indent=1)
lf = os.path.join(go, 'launches_%s.csv' % tag)
if os.path.exists(lf):
    rows = [r for r in csv.reader(open(lf)) if len(r) > 5]
    hdr = rows[0]; ik = hdr.index('Kernel Name'); iv = hdr.index('Metric Value'); iu = hdr.index('Metric Unit')
    tot = collections.defaultdict(float); cnt = collections.Counter()
    for r in rows[1:]:
        try:
            v = float(r[iv].replace(',', ''))
        except ValueError:
            continue
        name = r[ik].split('(')[0][:60].replace(',', ';')
        tot[name] += v; cnt[name] += 1
    S = sum(tot.values())
    with open(os.path.join(root, 'profiles', '%s_launches_summary.csv' % tag), 'w') as f:
        f.write('kernel,launches,total_%s,share\n' % rows[1][iu])
        for k, v in sorted(tot.items(), key=lambda x: -x[1]):
            f.write('%s,%d,%.1f,%.4f\n' % (k, cnt[k], v, v / S))
    import shutil
    shutil.copy(lf, os.path.join(root, 'profiles', '%s_launches_bench_steps3.csv' % tag))
bf = os.path.join(go, 'bench_%s.log' % tag)
if os.path.exists(bf):
    with open(os.path.join(root, 'profiles', '%s_bench_n1.json' % tag), 'w') as f:
        f.write(open(bf).read().strip().splitlines()[-1] + '\n')
