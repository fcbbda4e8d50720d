#!/usr/bin/env python
"""Record the DRAM traffic of the pairing kernel from an `ncu --set full` report into profiles/traffic.json and write
a text summary next to it:  python tools/ncu_traffic.py <workload> <report.ncu-rep> <round-tag>"""
import csv, io, json, os, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KEYS = ['gpu__time_duration.sum', 'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size',
        'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'smsp__inst_executed.sum', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'smsp__thread_inst_executed_per_inst_executed.ratio', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'l1tex__throughput.avg.pct_of_peak_sustained_elapsed',
        'l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'l1tex__data_pipe_lsu_wavefronts_mem_shared_op_atom.sum',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_atom.sum', 'smsp__inst_executed_op_shared_atom.sum',
        'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active', 'lts__t_sector_hit_rate.pct', 'l1tex__t_sector_hit_rate.pct']


def to_bytes(v, unit):
    mult = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1)
    return float(v) * mult


def main():
    workload, rep, tag = sys.argv[1:4]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, r = rows[0], rows[1], rows[2]
    col = {h: i for i, h in enumerate(hdr)}
    name = r[col['Kernel Name']]
    rd = to_bytes(r[col['dram__bytes_read.sum']], units[col['dram__bytes_read.sum']])
    wr = to_bytes(r[col['dram__bytes_write.sum']], units[col['dram__bytes_write.sum']])
    out = os.path.join(ROOT, "profiles", f"{tag}_{workload}_{name.split('<')[0].split('(')[0].strip().split()[-1]}.txt")
    with open(out, "w") as fh:
        fh.write(f"# ncu --set full --clock-control none, one launch of the pairing kernel of `{workload}` ({os.path.basename(rep)})\n")
        fh.write(f"kernel = {name}\n")
        for k in KEYS:
            if k in col:
                fh.write(f"{k} [{units[col[k]]}] = {r[col[k]]}\n")
        for h in hdr:
            if 'issue_stalled' in h and 'per_issue_active' in h:
                try:
                    if float(r[col[h]]) > 0.3:
                        fh.write(f"{h} = {r[col[h]]}\n")
                except ValueError:
                    pass
    tj = os.path.join(ROOT, "profiles", "traffic.json")
    data = json.load(open(tj)) if os.path.exists(tj) else {}
    data[workload] = {"kernel": name, "dram_bytes": rd + wr, "dram_bytes_read": rd, "dram_bytes_write": wr,
                      "source": os.path.relpath(out, ROOT)}
    json.dump(data, open(tj, "w"), indent=1, sort_keys=True)
    print(open(out).read())


if __name__ == "__main__":
    main()
