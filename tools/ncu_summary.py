"""Summarise an `ncu --set full` report into the text form kept under profiles/.
Usage: python tools/ncu_summary.py gpurun_out/prof.ncu-rep 'header line' > profiles/rNN_ncu_full_summary.txt
(runs `ncu -i <rep> --page raw --csv` here; no GPU needed)."""
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
header = sys.argv[2] if len(sys.argv) > 2 else ''
# a .csv argument is the already exported raw page (`ncu -i x.ncu-rep --page raw --csv` run on the GPU box: reports with --import-source
# exceed the 64 MiB that travel back)
out = open(rep).read() if rep.endswith('.csv') else subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr, units = rows[0], rows[1]
idx = {h: i for i, h in enumerate(hdr)}
want = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 'lts__throughput.avg.pct_of_peak_sustained_elapsed',
        'l1tex__m_xbar2l1tex_read_bytes.sum', 'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum', 'launch__registers_per_thread', 'launch__grid_size',
        'launch__waves_per_multiprocessor']
print(f'# {header}')
print('# Values from `ncu -i <rep> --page raw --csv`; traffic = dram__bytes_read.sum + dram__bytes_write.sum per launch.')
for r in rows[2:]:
    print()
    print('## ' + r[idx['Kernel Name']][:140])
    for w in want:
        if w in idx:
            print(f'   {w:74s} {r[idx[w]]} {units[idx[w]]}')

    def val(name):
        v = float(r[idx[name]].replace(',', ''))
        u = units[idx[name]].lower()
        return v * {'gbyte': 1e9, 'mbyte': 1e6, 'kbyte': 1e3, 'byte': 1.0}.get(u, 1.0)
    try:
        tr = val('dram__bytes_read.sum') + val('dram__bytes_write.sum')
        t = float(r[idx['gpu__time_duration.sum']].replace(',', ''))
        tu = units[idx['gpu__time_duration.sum']].lower()
        t_us = t * {'us': 1.0, 'usecond': 1.0, 'ms': 1e3, 'msecond': 1e3, 'ns': 1e-3, 'nsecond': 1e-3}.get(tu, 1.0)
        print(f'   {"traffic (read+write)":74s} {tr / 1e6:.1f} Mbyte   -> {tr / t_us / 1e6:.2f} TB/s over the kernel duration')
    except Exception as ex:  # noqa: BLE001
        print(f'   traffic: n/a ({ex})')
